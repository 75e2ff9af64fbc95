#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c12
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc $?" >> $O/smoke.log
timeout 300 python scripts/kmeans_times.py > $O/kmeans_times.log 2>&1
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
