#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c20
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29551 scripts/debug_sharded_pca.py 1000000 7552 8 > $O/debug2.log 2>&1
timeout 400 $TR 29552 bench.py --gpus 2 --steps 5 --warmup 3 --cells 500000 --genes 7552 --no-e2e --no-ps > $O/bench2.json 2> $O/bench2.err
