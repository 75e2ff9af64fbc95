#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c18
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_embed.py -m gpu -q -x > $O/embed.log 2>&1; echo "rc $?" >> $O/embed.log
timeout 300 python scripts/pass_times.py 1000000 3750 ts > $O/pass_times.log 2>&1
timeout 300 python scripts/profile_dist.py 1000000 3750 $O/profile_1m.txt > $O/profile_1m.log 2>&1
GC_RSVD_MODE=ss timeout 300 python scripts/profile_dist.py 1000000 3750 $O/profile_1m_ss.txt > $O/profile_1m_ss.log 2>&1
timeout 300 python scripts/profile_dist.py 50000 20000 $O/profile_n1.txt > $O/profile_n1.log 2>&1
