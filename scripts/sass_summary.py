#!/usr/bin/env python
"""profiles/sass_summary.txt: per compiled object of emrt_b200/csrc, how many Blackwell-native SASS instructions it holds
(cuobjdump -sass of emrt_b200/build/*.o): UTCHMMA = tcgen05.mma (bf16 kind::f16), UTMALDG / UTMASTG = TMA tensor
load / store (cp.async.bulk.tensor), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,
plus the gather's LDS.128 / HFMA2.BF16 and the backward's shared / global reductions.  Regenerate after every kernel change:
  python scripts/sass_summary.py > profiles/sass_summary.txt     (runs here: no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emrt_b200 import build  # noqa: E402

build.build()
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "LDS.128", "LDS.64",
            "STS.128", "HFMA2.BF16", "FHFMA", "HFMA2", "ATOMS", "REDG", "RED.", "LDG.E.128", "STG.E.128", "SHFL", "MUFU.EX2"]
print("# SASS mnemonic counts per object (cuobjdump -sass, sm_100a).  Columns: mnemonic prefix -> static instruction count.")
for src in build.SOURCES:
    obj = os.path.join(build.OBJ, src.replace(".cu", ".o"))
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    kernels = re.findall(r"Function : (\S+)", out)
    cnt = collections.Counter()
    mnems = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", out, flags=re.M)
    for m in mnems:
        for p in PATTERNS:
            if m.startswith(p):
                cnt[p] += 1
    print(f"\n{src}  ({', '.join(sorted(arch))}; {len(kernels)} kernels, {len(mnems)} instructions)")
    shown = [f"{p} {cnt[p]}" for p in PATTERNS if cnt[p]]
    print("  " + ("  ".join(shown) if shown else "(none of the listed mnemonics)"))
    tma_dims = collections.Counter(re.findall(r"UTMALDG\.(\dD)", out))
    if tma_dims:
        print("  UTMALDG by box rank: " + "  ".join(f"{k} {v}" for k, v in sorted(tma_dims.items())))
    two = {k: len(re.findall(k + r"[A-Z0-9_.]*\.2CTA", out)) for k in ("UTCHMMA", "UTMALDG", "UTCBAR")}
    if any(two.values()):       # CTA-pair forms: tcgen05.mma.cta_group::2, the .cta_group::2 TMA load, multicast commit
        print("  of which .2CTA (cta_group::2): " + "  ".join(f"{k} {v}" for k, v in two.items() if v))
