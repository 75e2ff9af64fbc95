#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c24
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/pass_times.log 2>&1
timeout 200 python scripts/pass_times.py 400000 2500 ts >> $O/pass_times.log 2>&1
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rsvd_apply_ts_kernel -s 22 -c 2 -o $O/prof_ts python scripts/pass_times.py 50000 20000 ts > $O/ncu.log 2>&1
