"""Dev tool (no GPU needed): counts of the SASS mnemonics that show which hardware paths each kernel of the built library uses
(`cuobjdump -sass`): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tile loads, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, 256-bit LDG/STG = full-sector streaming accesses, HMMA = the legacy mma.sync path (expected: none).
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tangram_b200", "libtangram_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"),
                                ("SYNCS", r"\bSYNCS"), ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("STG.256", r"\bSTG\.[A-Z0-9.]*256"),
                                ("f32x2", r"\b(FFMA2|FMUL2|FADD2)\b"), ("MUFU.EX2", r"MUFU\.EX2"), ("HMMA", r"\bHMMA")])
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
    elif cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':78s} " + " ".join(f"{k:>8s}" for k in pats))
tot = collections.Counter()
for (k, c), nm in zip(counts.items(), names):
    tot.update(c)
    nm = re.sub(r"\(.*", "", nm).replace("void ", "").replace("tgb::", "")
    print(f"{nm[:78]:78s} " + " ".join(f"{c[p]:8d}" for p in pats))
print(f"{'total':78s} " + " ".join(f"{tot[p]:8d}" for p in pats))
