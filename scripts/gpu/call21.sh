#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c21
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 400 python scripts/debug_sharded_pca.py 1000000 3776 12 > $O/debug1.log 2>&1
timeout 300 python scripts/debug_sharded_pca.py 50000 20000 30 > $O/debug_c3.log 2>&1
GC_RSVD_MODE=ss timeout 300 python scripts/debug_sharded_pca.py 1000000 3776 8 > $O/debug1_ss.log 2>&1
