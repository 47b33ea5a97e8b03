"""Summarise an ncu launch list of one proof (scripts/prove_probe2.py) by kernel: launches, total ms, DRAM GB, GB/s."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0].replace("lg::", "").replace("void ", "")
    metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    a = agg.setdefault(name, {"launch_ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["launch_ids"].add(r[ix["ID"]])
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    if metric == "gpu__time_duration.sum":
        a["ns"] += val * scale
    elif metric == "dram__bytes_read.sum":
        a["rd"] += val * scale
    elif metric == "dram__bytes_write.sum":
        a["wr"] += val * scale
tot = sum(a["ns"] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# kernel, launches, total ms, share of device time, DRAM read GB, DRAM write GB, DRAM GB/s")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    ms = a["ns"] / 1e6
    gb = (a["rd"] + a["wr"]) / 1e9
    print(f"{name:58s} {len(a['launch_ids']):4d} {ms:9.3f} {100 * a['ns'] / tot:5.1f}% {a['rd'] / 1e9:8.2f} {a['wr'] / 1e9:8.2f} {gb / (ms * 1e-3) if ms else 0:8.0f}")
print(f"# total device time {tot / 1e6:.2f} ms over {sum(len(a['launch_ids']) for a in agg.values())} launches")
