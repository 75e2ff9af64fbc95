#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c25
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for v in park epi epiiss all30 all150; do
  echo "== $v" >> $O/variants.log
  GENECLUST_B200_LIB=scripts/microbench/variants/lib$v.so timeout 200 python scripts/pass_times.py 50000 20000 ts >> $O/variants.log 2>&1
done
echo "== in-tree (epi 400 / mma 60 / issuer 100)" >> $O/variants.log
timeout 200 python scripts/pass_times.py 50000 20000 ts >> $O/variants.log 2>&1
timeout 200 python scripts/pass_times.py 400000 2500 ts >> $O/variants.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
