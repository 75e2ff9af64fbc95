#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c23
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29551 scripts/debug_sharded_pca.py 1000000 7552 24 > $O/debug2.log 2>&1
