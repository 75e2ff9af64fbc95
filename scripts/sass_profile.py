"""profiles/r2_sass_hist.txt: SASS mnemonic counts per kernel of the built library (the mnemonics that prove the tensor-core,
tensor-memory and TMA paths, plus the FP64 / conversion ones the rooflines talk about).
    python scripts/sass_profile.py > profiles/r2_sass_hist.txt"""
import collections
import re
import subprocess

LIB = "scgeneclust_b200/libgeneclust_b200.so"
KEEP = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDUX", "MUFU", "I2F", "DFMA",
        "DADD", "DMUL", "DSETP", "HMMA"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
print(f"SASS mnemonic counts per kernel of {LIB} (cuobjdump -sass; round 2, final)")
print("UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = tensor-map TMA load, UBLKCP = cp.async.bulk, "
      "UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDUX = warp reduction\n")
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0].strip()
    ops = collections.Counter(m.group(1).split('.')[0] for m in
                              re.finditer(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', f))
    kept = {k: ops[k] for k in KEEP if ops.get(k)}
    if kept:
        print(f"{name}  total {sum(ops.values())}  {kept}")
