"""Attribute the warp-stall samples / executed instructions of a warp-specialised kernel to its roles.
Roles are source-line ranges of the kernel's .cu file; SASS that comes from inlined headers is attributed to the role
of the nearest preceding SASS address that maps into the .cu file.
usage: python tools/ncu_roles.py report.ncu-rep file.cu name:first_line [name:first_line ...]"""
import csv, subprocess, io, collections, os, sys, bisect
rep, cu = sys.argv[1], sys.argv[2]
bounds = sorted((int(a.split(":")[1]), a.split(":")[0]) for a in sys.argv[3:])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, line, sass = "?", 0, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if not r or r[0] in ("Line No", "Function Name") or len(r) < 8:
        continue
    if r[2] == "-":
        try: line = int(r[0])
        except ValueError: pass
        continue
    if not r[2].startswith("0x"):
        continue
    try: s, n = int(r[6] or 0), int(r[7] or 0)
    except ValueError: continue
    sass.append((int(r[2], 16), fname, line, s, n, r[3].split()[0] if r[3].split() else "?"))
sass.sort(key=lambda t: (t[0], t[1] != cu))       # an instruction inlined from a header is listed under both files: keep one row
dedup, seen = [], set()
for t in sass:
    if t[0] in seen: continue
    seen.add(t[0]); dedup.append(t)
sass = dedup
def role_of(ln):
    i = bisect.bisect_right([b[0] for b in bounds], ln) - 1
    return bounds[i][1] if i >= 0 else "pre"
agg, inst, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
cur = "pre"
for addr, f, ln, s, n, op in sass:
    if f == cu: cur = role_of(ln)
    agg[cur] += s; inst[cur] += n; ops[cur][op] += n
tot, ti = sum(agg.values()) or 1, sum(inst.values()) or 1
print(f"samples {tot}, warp-instructions {ti}")
for k, v in agg.most_common():
    top = ", ".join(f"{o} {100 * c / max(inst[k], 1):.0f}%" for o, c in ops[k].most_common(6))
    print(f"{k:12s} samples {100 * v / tot:5.1f}%  inst {100 * inst[k] / ti:5.1f}%   {top}")
