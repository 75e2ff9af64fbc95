"""Build a variant of libgeneclust_b200.so with extra -D flags on rsvd_tc.cu (developer tool, kernel experiments):
    python scripts/build_variant.py <name> [-DFLAG ...]   ->  scripts/microbench/variants/lib<name>.so"""
import os
import subprocess
import sys

sys.path.insert(0, ".")
from scgeneclust_b200.csrc import build as B

name, flags = sys.argv[1], sys.argv[2:]
B.build()
out_dir = os.path.join("scripts", "microbench", "variants")
os.makedirs(out_dir, exist_ok=True)
obj_dir = os.path.join(B.HERE, "build")
obj = os.path.join(out_dir, f"rsvd_tc_{name}.o")
subprocess.run([B._nvcc(), *B.COMMON, *flags, "-c", os.path.join(B.HERE, "rsvd_tc.cu"), "-o", obj], check=True)
objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s, _ in B.SOURCES if s != "rsvd_tc.cu"] + [obj]
lib = os.path.join(out_dir, f"lib{name}.so")
subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
print(lib)
