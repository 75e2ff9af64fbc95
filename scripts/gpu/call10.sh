#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c10
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
for cap in 48 32 24 16; do
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pass_times.py 50000 20000 ts 2>&1 | sed "s/^/cap $cap /" >> $O/pass_times.log
done
GC_RSVD_MAX_STAGES=32 timeout 200 python scripts/pass_times.py 400000 2500 ts 2>&1 | sed "s/^/cap 32 /" >> $O/pass_times.log
GC_RSVD_MAX_STAGES=32 timeout 200 python scripts/pass_times.py 125000 3750 ts 2>&1 | sed "s/^/cap 32 /" >> $O/pass_times.log
GC_RSVD_MAX_STAGES=12 timeout 200 python scripts/pca_accuracy.py big /tmp/big_ref.npy > $O/accuracy.log 2>&1
for cap in 48 32 24; do
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pca_accuracy.py c3 >> $O/accuracy.log 2>&1
  GC_RSVD_MAX_STAGES=$cap timeout 200 python scripts/pca_accuracy.py big /tmp/big_ref.npy >> $O/accuracy.log 2>&1
done
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
