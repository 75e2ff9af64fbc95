#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c5
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/pass_times.log 2>&1
python scripts/build_variant.py g5s5 -DTS_GROUPS_N=5 -DTS_A_STAGES_N=5 >> $O/build.log 2>&1
python scripts/build_variant.py g3s5 -DTS_GROUPS_N=3 -DTS_A_STAGES_N=5 >> $O/build.log 2>&1
python scripts/build_variant.py g4s4 -DTS_GROUPS_N=4 -DTS_A_STAGES_N=4 >> $O/build.log 2>&1
for v in g5s5 g3s5 g4s4; do
  GENECLUST_B200_LIB=$PWD/scripts/microbench/variants/lib$v.so timeout 200 python scripts/pass_times.py 50000 20000 ts >> $O/pass_times.log 2>&1
done
timeout 200 python scripts/pass_times.py 400000 2500 ts >> $O/pass_times.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_subsample.py tests/test_gpu_dist.py -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 300 python scripts/kmeans_times.py > $O/kmeans_times.log 2>&1
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ps > $O/bench.json 2> $O/bench.err
timeout 200 python scripts/pass_times.py 50000 20000 ts > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rsvd_apply_ts_kernel -s 22 -c 2 -o $O/prof_ts python scripts/pass_times.py 50000 20000 ts > $O/ncu.log 2>&1
