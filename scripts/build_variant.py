"""Build a variant of libgeneclust_b200.so with extra -D flags on rsvd_tc.cu, or with another translation unit in its
place (developer tool, kernel experiments):
    python scripts/build_variant.py <name> [-DFLAG ...] [--replace other.cu]   ->  scripts/microbench/variants/lib<name>.so"""
import os
import subprocess
import sys

sys.path.insert(0, ".")
from scgeneclust_b200.csrc import build as B

name, flags = sys.argv[1], sys.argv[2:]
source = os.path.join(B.HERE, "rsvd_tc.cu")
if "--replace" in flags:
    i = flags.index("--replace")
    source = flags[i + 1]
    flags = flags[:i] + flags[i + 2:] + ["-I", B.HERE]
B.build()
out_dir = os.path.join("scripts", "microbench", "variants")
os.makedirs(out_dir, exist_ok=True)
obj_dir = os.path.join(B.HERE, "build")
obj = os.path.join(out_dir, f"rsvd_tc_{name}.o")
subprocess.run([B._nvcc(), *B.COMMON, *flags, "-c", source, "-o", obj], check=True)
objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s, _ in B.SOURCES if s != "rsvd_tc.cu"] + [obj]
lib = os.path.join(out_dir, f"lib{name}.so")
subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(lib)
