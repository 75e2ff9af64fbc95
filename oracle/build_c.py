"""Build the C pieces of the oracle with gcc (TEST/BENCH INFRASTRUCTURE, see oracle/__init__.py):

    python -m oracle.build_c        ->  oracle/_build/liboracle_synth.so

The reference itself has no native code, so there is nothing to compile into oracle/_ref;
the only C here is the fast twin of the synthetic-count generator (oracle/synth_c.c).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_synth.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "synth_c.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the oracle's C generator")
    os.makedirs(OUT_DIR, exist_ok=True)
    # -ffp-contract=off: no fused multiply-add, the float64 recurrences must round like numpy's
    cmd = [gcc, "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed on synth_c.c:\n{res.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
