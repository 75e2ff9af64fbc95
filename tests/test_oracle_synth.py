"""CPU checks of the oracle's input generator and of the CPU arm of bench.py (no GPU needed)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_generator_equals_numpy_twin_and_product_tables():
    from oracle import synth as osy
    from scgeneclust_b200.synth import synth_params
    p_prod, p_orc = synth_params(3000, 3), osy.synth_params(3000, 3)
    for k in ("gene_mean", "type_fc", "gene_module", "size_levels", "module_levels"):
        assert np.array_equal(getattr(p_prod, k), getattr(p_orc, k)), k
    a = osy.synth_counts(p_orc, 1234, 24, 100, 700)
    b = osy.synth_counts_c(p_orc, 1234, 24, 100, 700)
    assert a.dtype == b.dtype == np.uint16 and np.array_equal(a, b)
    # a sub-block of a bigger call is the same stream (counter based)
    big = osy.synth_counts_c(p_orc, 1200, 100, 0, 3000)
    assert np.array_equal(big[34:58, 100:800], b)


def test_reference_arm_line_small():
    """`bench.py --impl reference` on a tiny matrix: one JSON line with the contract's keys, produced without loading the
    product's shared library."""
    env = dict(os.environ, OMP_NUM_THREADS="1")          # as torchrun would export it: the arm must undo it
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "600",
                          "--genes", "800", "--clusters", "12", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "geneclust_fast_selection_throughput"
    assert line["config"]["cells"] == 600 and "SAMPLED" not in line["config"]["workload"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["host_threads"]["inherited_env"] == {"OMP_NUM_THREADS": "1"}
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["steps"] == 1 and line["product_so_mapped"] is False
