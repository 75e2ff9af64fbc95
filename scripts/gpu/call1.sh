#!/bin/bash
# GPU call 1 (round 2): new parity tests, bench arms, TS draft probe, stage times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2c1
mkdir -p $O
nproc > $O/host.txt; free -g >> $O/host.txt; nvidia-smi -L >> $O/host.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
# TS-mode draft: parity through the embed tests + timing
python scripts/build_variant.py ts --replace scripts/experimental/rsvd_tc_ts.cu > $O/ts_build.log 2>&1
GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/libts.so timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -x -q > $O/ts_pytest.log 2>&1; echo "rc $?" >> $O/ts_pytest.log
GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/libts.so timeout 120 python scripts/pass_times.py > $O/ts_times.log 2>&1
timeout 120 python scripts/pass_times.py > $O/base_times.log 2>&1
timeout 300 python scripts/stage_times.py > $O/stage_times.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc $?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc $?" >> $O/bench_ref.err
