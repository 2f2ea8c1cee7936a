"""Summarise an ncu --set full report into profiles/<name>.json (one entry per captured launch) and, for the
dominant kernel, profiles/dominant_kernel_ncu.json (read by bench.py for roofline.traffic).
usage: python tools/ncu_summary.py report.ncu-rep out_name [dominant_kernel_regex]"""
import csv, io, json, os, re, subprocess, sys

rep, name = sys.argv[1], sys.argv[2]
dom = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
res = []
for r in rows[2:]:
    e = {"kernel": r[idx["Kernel Name"]]}
    for w in want:
        if w in idx:
            try:
                v = float(r[idx[w]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[w]]
            if u in scale:
                v *= scale[u]
                u = "s" if u in ("ns", "us", "ms", "s") else "byte"
            e[w] = {"value": v, "unit": u}
    res.append(e)
os.makedirs("profiles", exist_ok=True)
json.dump({"report": os.path.basename(rep), "launches": res}, open(f"profiles/{name}.json", "w"), indent=1)
print(f"profiles/{name}.json: {len(res)} launches")
if dom:
    for e in res:
        if re.search(dom, e["kernel"]):
            d = {"kernel": e["kernel"], "source_report": os.path.basename(rep),
                 "dram_bytes_per_launch": e["dram__bytes_read.sum"]["value"] + e["dram__bytes_write.sum"]["value"],
                 "dram_bytes_read": e["dram__bytes_read.sum"]["value"], "dram_bytes_write": e["dram__bytes_write.sum"]["value"],
                 "duration_under_ncu_s": e["gpu__time_duration.sum"]["value"],
                 "tensor_pipe_active_pct": e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", {}).get("value"),
                 "issue_active_pct": e.get("smsp__issue_active.avg.pct_of_peak_sustained_active", {}).get("value")}
            json.dump(d, open("profiles/dominant_kernel_ncu.json", "w"), indent=1)
            print("profiles/dominant_kernel_ncu.json", d)
            break
