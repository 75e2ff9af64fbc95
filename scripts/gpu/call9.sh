#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c9
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
GC_RSVD_MAX_STAGES=12 timeout 200 python scripts/pca_accuracy.py big /tmp/big_ref.npy > $O/accuracy.log 2>&1
for cap in 12 24 48 96 100000; do
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pca_accuracy.py c3 >> $O/accuracy.log 2>&1
  GC_RSVD_MAX_STAGES=$cap timeout 300 python scripts/pca_accuracy.py sub >> $O/accuracy.log 2>&1
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pca_accuracy.py big /tmp/big_ref.npy >> $O/accuracy.log 2>&1
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pass_times.py 50000 20000 ts 2>&1 | sed "s/^/cap $cap /" >> $O/pass_times.log
done
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/kmeans_times.py > $O/kmeans_times.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
