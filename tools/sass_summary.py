"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (tensor-map TMA), UTCBAR (tcgen05.commit), MUFU.
usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magnet_b200", "lib", "libmagnet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "FFMA", "LDG", "STG", "RED", "ATOM"]
cur, counts, size, arch = None, collections.OrderedDict(), {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter(); size[cur] = 0
        continue
    m = re.search(r"\.target\s+(\S+)|arch = (\S+)", line)
    if m and cur is None:
        arch["a"] = m.group(1) or m.group(2)
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        size[cur] += 1
        op = m.group(2)
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
print("# cuobjdump -sass %s  (arch %s)" % (os.path.relpath(lib), arch.get("a", "sm_100a")))
print("%-64s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in keys))
for name, c in counts.items():
    print("%-64s %7d " % (name[-64:], size[name]) + " ".join("%7d" % c[k] for k in keys))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("%-64s %7d " % ("TOTAL", sum(size.values())) + " ".join("%7d" % tot[k] for k in keys))
