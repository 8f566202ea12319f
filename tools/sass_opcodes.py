"""Per-kernel counts of the Blackwell-only SASS opcodes in libdcvit.so (cuobjdump -sass): UTCHMMA (tcgen05.mma, .2CTA =
cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), MUFU.  Usage: python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "diverse_channel_vit_b200" / "libdcvit.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA2"]
print(f"# {lib.name}: SASS opcode counts per kernel (sm_100a); only kernels that use tensor memory, TMA or mbarriers are listed")
print("# " + " ".join(f"{o:>12s}" for o in OPS) + "   instr  kernel")
tot = collections.Counter()
for blk in re.split(r"\n\s*Function : ", txt)[1:]:
    name = blk.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(anonymous namespace\)::|dcv::", "", dem)
    dem = re.sub(r"\(CUtensorMap_st.*", "", dem)[:100]
    ins = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\w+\s+)?", "", l).split()[0].rstrip(";") for l in blk.split("\n")
           if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
    c = collections.Counter()
    for i in ins:
        base = i.split(".")[0]
        if base in OPS:
            c[base] += 1
        if i.startswith("UTCHMMA.2CTA"):
            c["UTCHMMA.2CTA"] += 1
    if not any(c[o] for o in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS")):
        continue
    tot.update(c)
    print("  " + " ".join(f"{c[o]:12d}" for o in OPS) + f" {len(ins):7d}  {dem}")
print("# total")
print("  " + " ".join(f"{tot[o]:12d}" for o in OPS))
