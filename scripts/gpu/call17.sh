#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c17
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -s -k "sum_axis0 or deviance" > $O/dev.log 2>&1; echo "rc $?" >> $O/dev.log
timeout 200 python scripts/deviance_times.py 400000 7 > $O/deviance_times.log 2>&1
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ps > $O/bench.json 2> $O/bench.err
