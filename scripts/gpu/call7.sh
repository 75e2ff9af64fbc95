#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c7
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/pass_times.log 2>&1
python scripts/build_variant.py i2f0 -DTS_I2F_WORDS_N=0 >> $O/build.log 2>&1
python scripts/build_variant.py i2f4 -DTS_I2F_WORDS_N=4 >> $O/build.log 2>&1
python scripts/build_variant.py i2f1 -DTS_I2F_WORDS_N=1 >> $O/build.log 2>&1
for v in i2f0 i2f1 i2f4; do
  GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/lib$v.so timeout 200 python scripts/pass_times.py 50000 20000 ts >> $O/pass_times.log 2>&1
done
timeout 200 python scripts/pass_times.py 400000 2500 ts >> $O/pass_times.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/kmeans_times.py > $O/kmeans_times.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rsvd_apply_ts_kernel -s 22 -c 2 -o $O/prof_ts python scripts/pass_times.py 50000 20000 ts > $O/ncu.log 2>&1
