#!/bin/bash
# GPU call 2 (round 2): TS-mode kernel v4 parity + timing, full GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c2
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 200 python scripts/pass_times.py > $O/pass_times.log 2>&1
timeout 200 python scripts/pass_times.py 400000 2500 >> $O/pass_times.log 2>&1
timeout 200 python scripts/pass_times.py 2700 32201 >> $O/pass_times.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ps > $O/bench.json 2> $O/bench.err
