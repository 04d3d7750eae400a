"""Per-kernel census of the Blackwell tensor-core / TMA / TMEM instructions in the built library (profiles/r02_sass_*.txt):
    python scripts/sass_excerpt.py > profiles/r02_sass_tcgen05_tma.txt
UTCHMMA = tcgen05.mma (kind::tf32), UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UTMACCTL = tensor-map proxy fence, SYNCS = mbarrier ops, UTCATOMSWS = TMEM allocation."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "rlrep_b200" / "librlrep_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTMACCTL|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|MUFU|HMMA|IMMA|RED|ATOMG|MEMBAR)\b[\w.]*")
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    if kern and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        total[kern] += 1
        for mm in pat.finditer(line):
            counts[kern][mm.group(0).split(".")[0]] += 1
print(f"# cuobjdump -sass {lib.name} (sm_100a): instructions per kernel; only kernels that use tensor cores / TMA / TMEM are listed")
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
for (k, c), name in zip(counts.items(), demangle):
    if not any(c[x] for x in ("UTCHMMA", "UTMALDG", "LDTM")):
        continue
    short = name.replace("(anonymous namespace)::", "").replace("rlrep::", "").replace("tc::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", short)
    print(f"{short[:70]:70s} sass={total[k]:6d}  " + "  ".join(f"{x}={c[x]}" for x in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UTMACCTL", "UTCATOMSWS", "SYNCS", "RED", "ATOMG") if c[x]))
