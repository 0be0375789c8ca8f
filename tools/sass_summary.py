"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
    python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt
UTCHMMA = tcgen05.mma (f16 kind), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor
load / store, UBLKCP = bulk copy, SYNCS = mbarrier ops."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dream_b200", "libdreamb200.so")
WANT = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "DADD", "DMUL")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_n"] += 1
        for w in WANT:
            if op.startswith(w):
                per[cur][w] += 1
names = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print("%-92s %7s " % ("kernel", "instr") + " ".join("%8s" % w for w in WANT))
tot = collections.Counter()
for (k, c), n in zip(per.items(), names):
    n = re.sub(r"\(.*", "", n).replace("void db200::", "").replace("db200::", "")
    print("%-92s %7d " % (n[:92], c["_n"]) + " ".join("%8d" % c[w] for w in WANT))
    tot.update(c)
print("%-92s %7d " % ("TOTAL", tot["_n"]) + " ".join("%8d" % tot[w] for w in WANT))
