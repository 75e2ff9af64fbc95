#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c22
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 500 $TR 29551 scripts/debug_sharded_pca.py 1000000 7552 16 > $O/debug2.log 2>&1
GC_PANEL_EXCHANGE=allreduce timeout 500 $TR 29553 scripts/debug_sharded_pca.py 1000000 7552 16 > $O/debug2_ar.log 2>&1
