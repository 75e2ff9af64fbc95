#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/final
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ps --no-e2e > $O/bench_short.json 2> $O/bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ps --no-e2e > $O/ncu_launches.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
