"""profiles/<round>_sass_evidence.txt: which kernels of the library contain tcgen05 / TMEM / TMA / mbarrier instructions.

    cuobjdump -sass llamole_b200/libllamole_b200.so | python tools/sass_evidence.py r1
"""
import collections, os, re, subprocess, sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|LDTM|STTM|HMMA|LDSM|SYNCS|REDUX)\b")
cur, counts = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k in pat.findall(line):
            counts[cur][k] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
out = ["# SASS evidence (cuobjdump -sass llamole_b200/libllamole_b200.so, sm_100a): instruction counts per kernel for the",
       "# B200_PROFILING.md mnemonics: tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld / st = LDTM / STTM,",
       "# TMA = UTMALDG / UTMASTG / UTMAPF / UBLKCP / UBLKPF, mbarrier = SYNCS, mma.sync = HMMA, ldmatrix = LDSM, redux.sync = REDUX.",
       "# Kernels with none of them (row kernels, CSR build, step kernel, aggregation) are omitted.", ""]
for k, d in zip(names, dem):
    c = counts[k]
    if not c:
        continue
    d = d.replace("(anonymous namespace)::", "").replace("llb::", "").replace("void ", "")
    d = re.sub(r"\((bool|int|unsigned int)\)", "", d)
    name = re.sub(r"\(.*", "", d)
    out.append(f"{name[:100]:100s} " + " ".join(f"{a}={b}" for a, b in sorted(c.items())))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{rnd}_sass_evidence.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out) - 5, "kernels")
