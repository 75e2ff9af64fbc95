#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c8
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py tests/test_gpu_kmeans.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/pass_times.log 2>&1
python scripts/build_variant.py w200 -DTS_WAIT_SLEEP_NS_N=200 >> $O/build.log 2>&1
python scripts/build_variant.py w500 -DTS_WAIT_SLEEP_NS_N=500 >> $O/build.log 2>&1
python scripts/build_variant.py g5w200 -DTS_WAIT_SLEEP_NS_N=200 -DTS_GROUPS_N=5 >> $O/build.log 2>&1
for v in w200 w500 g5w200; do
  GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/lib$v.so timeout 200 python scripts/pass_times.py 50000 20000 ts >> $O/pass_times.log 2>&1
done
timeout 200 python scripts/pass_times.py 400000 2500 ts >> $O/pass_times.log 2>&1
timeout 300 python scripts/kmeans_times.py > $O/kmeans_times.log 2>&1
GC_KPP_PER_ROUND=1 timeout 300 python scripts/kmeans_times.py > $O/kmeans_times_perround.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
