"""SASS evidence from the built library (no GPU needed): per-kernel counts of the mnemonics that prove the Blackwell
path (tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA loads -> UTMALDG, TMA stores / add-reductions ->
UTMASTG / UTMAREDG, tcgen05.commit -> UTCBAR, packed fp32x2 -> FFMA2 ...; legacy mma.sync would show as HMMA) and a
short excerpt of each tensor-core kernel around its first MMAs.
usage: python profiles/sass_report.py flash_attention_metal_b200/libflash_attn_b200.so > profiles/r2_sass.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "MUFU.EX2", "FFMA2", "FADD2", "FMUL2", "FFMA",
        "LDGSTS", "LDS", "STS", "USETMAXREG", "UCGABAR", "SYNCS", "RED", "ATOM", "HMMA", "STL", "LDL"]

sass = subprocess.check_output(["cuobjdump", "-sass", sys.argv[1]], text=True)
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        kernels[cur].append(line)


def demangle(name):
    try:
        out = subprocess.check_output(["cu++filt", name], text=True).strip()
    except Exception:
        out = name
    out = re.sub(r"\(anonymous namespace\)::|fa::|<unnamed>::|\(int\)|^void ", "", out)
    return re.sub(r"\(.*", "", out)[:60]


print("# per-kernel SASS mnemonic counts, libflash_attn_b200.so (cuobjdump -sass, sm_100a); RED / ATOM = global atomics, HMMA = legacy mma.sync")
print("# kernel".ljust(62) + " ".join(k[:8].rjust(8) for k in KEYS))
total = Counter()
for name, lines in kernels.items():
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", l).split(" ")[0].rstrip(";") for l in lines]
    c = Counter()
    for op in ops:
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("LDS", "STS", "RED", "ATOM", "STL", "LDL") and op.split(".")[0] == k):
                c[k] += 1
                break
    total.update(c)
    print(demangle(name).ljust(62) + " ".join(str(c[k]).rjust(8) for k in KEYS))
print("TOTAL".ljust(62) + " ".join(str(total[k]).rjust(8) for k in KEYS))
print()
for name, lines in kernels.items():
    idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
    if not idx or "ILi128ELi1" not in name:
        continue
    print(f"## {demangle(name)}: first MMAs and the TMA / TMEM instructions of the kernel (excerpt)")
    for key, cap in (("UTMALDG", 4), ("UTCHMMA", 6), ("UTCBAR", 3), ("LDTM", 3), ("STTM", 2), ("UTMASTG", 2), ("UTMAREDG", 2), ("USETMAXREG", 2)):
        hits = [l for l in lines if key in l][:cap]
        for l in hits:
            print("   " + re.sub(r"\s+", " ", l.strip())[:150])
    print()
