"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, collections, io, os

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg, tot = "?", None, {}, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r
        ia = hdr.index("Address"); isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= isamp or r[ia] != "-":
        continue
    try:
        s = int(r[isamp] or 0); n = int(r[iinst] or 0)
    except ValueError:
        continue
    if s == 0 and n == 0:
        continue
    st = collections.Counter()
    for i, h in stall_cols:
        try:
            v = int(r[i] or 0)
        except ValueError:
            v = 0
        if v:
            st[h[6:]] += v
    key = (fname, r[0], r[1].strip())
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += s; a[1] += n; a[2].update(st); tot += s
tin = sum(a[1] for a in agg.values())
print(f"total samples {tot}  total warp-instructions {tin}")
for (f, ln, src), (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k}:{v}" for k, v in st.most_common(3))
    print(f"{100*s/max(tot,1):5.1f}% inst={100*n/max(tin,1):4.1f}%  {f}:{ln:>4} {src[:80]:80s} | {tops}")
