#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c19
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 200 python scripts/pass_times.py 1000000 3776 ts > $O/pass_times.log 2>&1
timeout 200 python scripts/pass_times.py 1000000 3568 ts >> $O/pass_times.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/debug_sharded_pca.py 1000000 30000 --same-device > $O/debug8.log 2>&1
