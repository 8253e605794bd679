"""Summarise an ncu --csv launch list (gpu__time_duration [+ dram bytes]) per kernel name."""
import collections
import csv
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0]
    m = r["Metric Name"]
    v = float(r["Metric Value"].replace(",", ""))
    a = agg.setdefault(k, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
    if m == "gpu__time_duration.sum":
        a["n"] += 1
        a["t"] += v / 1e6 if r["Metric Unit"] == "ns" else (v / 1e3 if r["Metric Unit"] == "us" else v)
    elif m == "dram__bytes_read.sum":
        a["rd"] += v * UNIT[r["Metric Unit"]]
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * UNIT[r["Metric Unit"]]
tot = sum(a["t"] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} rows, {tot:.3f} ms in kernels (cold-cache, serialised: compare shares)")
print("kernel,launches,total_ms,share_pct,dram_read_MB,dram_write_MB,dram_GBps")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    bw = (a["rd"] + a["wr"]) / 1e9 / (a["t"] / 1e3) if a["t"] > 0 else 0.0
    print(f"{k},{a['n']},{a['t']:.3f},{100 * a['t'] / tot:.2f},{a['rd'] / 1e6:.1f},{a['wr'] / 1e6:.1f},{bw:.0f}")
