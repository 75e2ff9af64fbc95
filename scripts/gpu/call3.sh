#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c3
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for sp in auto 1 2 4 8 16; do
  if [ $sp = auto ]; then timeout 200 python scripts/pass_accuracy.py >> $O/accuracy.log 2>&1;
  else GC_RSVD_SPLIT=$sp timeout 200 python scripts/pass_accuracy.py >> $O/accuracy.log 2>&1; fi
done
timeout 200 python scripts/pass_times.py > $O/pass_times.log 2>&1
timeout 600 python -m pytest tests/test_gpu_subsample.py tests/test_gpu_fullsize.py tests/test_gpu_embed.py -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
GC_RSVD_MODE=ss timeout 600 python -m pytest tests/test_gpu_subsample.py tests/test_gpu_fullsize.py -m gpu -q -s > $O/pytest_ss.log 2>&1; echo "pytest rc $?" >> $O/pytest_ss.log
