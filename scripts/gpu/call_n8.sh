#!/bin/bash
# 8-GPU call: per-stage table + kernel list at N=8, bench N=8 (parity block), config c5 + subsample check, NCCL dist tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n8c
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29602 scripts/profile_dist.py 50000 20000 $O/profile_n8_rs_ag.txt > $O/profile_n8.log 2>&1
timeout 600 $TR 29604 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 600 $TR 29605 bench.py --gpus 8 --steps 5 --warmup 3 --cells 125000 --genes 30000 --no-e2e --no-ps > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > $O/pytest_dist.log 2>&1; echo "rc $?" >> $O/pytest_dist.log
