"""Regenerate profiles/sass_evidence.md: per-kernel counts of the Blackwell-specific SASS mnemonics in the built library.

    python scripts/sass_evidence.py > profiles/sass_evidence.md

(no GPU needed: cuobjdump reads the cubin embedded in infomesh_b200/_native/libinfomesh_b200.so)"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "infomesh_b200" / "_native" / "libinfomesh_b200.so"
KEEP = re.compile(r"^(UTC[A-Z]*MMA|UTCCP|UTCBAR|UTCATOMSWS|UTCSHIFT|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTMACCTL|UBLKCP|UBLKRED|LDGMC|"
                  r"SYNCS|ELECT|ACQBULK|FENCE\.VIEW\.ASYNC|MEMBAR\.[A-Z.]+|REDG|ATOMG|ATOMS|MUFU\.[A-Z0-9]+|FFMA2|FMUL2|FADD2|"
                  r"F2FP[A-Z0-9.]*|HMMA|IMMA|QMMA|STG\.E\.MC|RED\.[A-Z.]*MC|ST\.MC|MULTIMEM|UCGABAR[A-Z_]*|STAS[A-Z0-9.]*|CCTL)")
COLLAPSE = ("SYNCS", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMACCTL", "LDTM", "REDG", "ATOMG", "ATOMS", "UBLKCP", "LDGMC", "UTCCP",
            "UTCQMMA", "UTCHMMA", "UTCOMMA", "UTCBAR", "F2FP", "ELECT", "UTMAREDG", "STAS", "UCGABAR")

out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        if KEEP.match(op):
            for c in COLLAPSE:
                if op.startswith(c):
                    op = c
                    break
            per[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
rows = {}
for mangled, name in zip(per, names):
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    c = per[mangled]
    if c:
        rows.setdefault(short, collections.Counter()).update(c)
total = collections.Counter()
for c in rows.values():
    total.update(c)
print("# SASS evidence (`python scripts/sass_evidence.py`: cuobjdump -sass of infomesh_b200/_native/libinfomesh_b200.so, sm_100a)\n")
print("UTCHMMA = tcgen05.mma kind::f16 · **UTCQMMA** = tcgen05.mma kind::f8f6f4 / kind::mxf8f6f4.block_scale (fp8 and block-scaled MMAs) · "
      "**UTCCP** = tcgen05.cp (scale factors smem -> TMEM) · LDTM = tcgen05.ld · UTMALDG/UTMASTG = TMA tensor load/store · "
      "**UBLKCP** = cp.async.bulk (1-D bulk copy: scale-factor chunks) · UTCBAR = tcgen05.commit -> mbarrier · UTCATOMSWS = TMEM "
      "alloc/dealloc · SYNCS = mbarrier ops · ELECT = elect.sync · **LDGMC** = multimem.ld_reduce (NVLS in-switch reduction) · "
      "REDG/STG on multicast addresses = multimem.red / multimem.st · **STAS** = st.async.shared::cluster with mbarrier complete_tx "
      "(distributed-shared-memory exchange of the fused LayerNorm epilogue) · UCGABAR = barrier.cluster · F2FP = packed fp32 -> e4m3/bf16 conversion (fused quantisers) · "
      "FFMA2/FMUL2/FADD2 = packed fp32x2 math · MUFU.* = SFU approximations · MEMBAR.*.SYS + REDG/ATOMG = system-scope release "
      "for peer flags.  HMMA / IMMA / QMMA (legacy mma.sync tensor-core paths) do not appear anywhere: "
      f"{'NONE FOUND' if not any(k in total for k in ('HMMA', 'IMMA', 'QMMA')) else 'PRESENT (!)'}.\n")
print("Library totals: " + ", ".join(f"{k}×{v}" for k, v in sorted(total.items()) if k.startswith(("UTC", "UTMA", "UBLK", "LDGMC", "LDTM"))) + "\n")
print("| kernel | instruction counts |\n|---|---|")
for name in sorted(rows):
    print(f"| `{name}` | " + ", ".join(f"{k}×{v}" for k, v in sorted(rows[name].items())) + " |")
