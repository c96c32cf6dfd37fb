"""Summarise `ncu --metrics ... --csv` per-launch metrics of ONE head step (tools/profile_round.sh) per kernel/grid shape."""
import collections
import csv
import re
import sys


def main(path, per_launch=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    L = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = L.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("grl::", ""), "grid": row["Grid Size"]})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d[row["Metric Name"]] = v
    T = sum(d["gpu__time_duration.sum"] for d in L.values())
    print("# %s: %d launches in one step, %.2f ms summed kernel time (serialised, cold cache)" % (path, len(L), T / 1e3))
    if per_launch:
        print("%-4s %-34s %-12s %9s %9s %8s %7s" % ("id", "kernel", "grid", "time_us", "dram_MB", "tensor%", "L2hit%"))
        for i, d in L.items():
            print("%-4s %-34s %-12s %9.1f %9.1f %8.1f %7.1f" % (i, d["name"][:34], d["grid"].replace(" ", ""), d["gpu__time_duration.sum"],
                  (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / 1e6,
                  d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], d["lts__t_sector_hit_rate.pct"]))
        return
    agg = collections.OrderedDict()
    for d in L.values():
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        t = d["gpu__time_duration.sum"]
        a["n"] += 1
        a["t"] += t
        a["dram"] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        a["tens_t"] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * t
    print("%-36s %5s %10s %8s %7s %10s %9s %8s" % ("kernel", "n", "total_us", "avg_us", "share", "dram_MB/l", "GB/s", "tensor%"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        print("%-36s %5d %10.1f %8.1f %6.1f%% %10.1f %9.0f %8.1f" % (k[:36], a["n"], a["t"], a["t"] / a["n"], 100 * a["t"] / T,
              a["dram"] / a["n"] / 1e6, a["dram"] / (a["t"] * 1e-6) / 1e9, a["tens_t"] / a["t"]))


def write_json(path, out_json, out_traffic, source):
    """Per-kernel DRAM bytes per launch (bench.py: hbm_kernels) and the mean over the GEMM launches (bench.py: roofline.traffic)."""
    import json
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    L = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = L.setdefault(row["ID"], {"name": re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "").replace("grl::", "")})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d[row["Metric Name"]] = v
    agg = collections.OrderedDict()
    for d in L.values():
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        a["n"] += 1
        a["t"] += d["gpu__time_duration.sum"]
        a["dram"] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        a["tens_t"] += d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * d["gpu__time_duration.sum"]
    kernels = {}
    for k, a in agg.items():
        gbs = a["dram"] / (a["t"] * 1e-6) / 1e9
        tensor = a["tens_t"] / a["t"]
        # streaming kernels: no tensor-pipe activity and a DRAM rate that says the time goes into HBM traffic
        bound = "tensor" if tensor > 5 else ("hbm" if gbs > 1500 else "latency")
        kernels[k] = {"launches_per_step": int(a["n"]), "dram_bytes_per_launch": a["dram"] / a["n"], "ncu_avg_us": a["t"] / a["n"],
                      "ncu_gbs": gbs, "bound": bound}
    with open(out_json, "w") as f:
        json.dump({"source": source, "kernels": kernels}, f, indent=1)
    g = [a for k, a in agg.items() if "gemm" in k]
    n = sum(a["n"] for a in g)
    with open(out_traffic, "w") as f:
        json.dump({"source": source, "gemm_launches": int(n), "dram_bytes_per_launch": sum(a["dram"] for a in g) / n,
                   "dram_bytes_per_step_all_kernels": sum(a["dram"] for a in agg.values())}, f, indent=1)


if __name__ == "__main__":
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        write_json(sys.argv[1], sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3])
    else:
        main(sys.argv[1], "--launches" in sys.argv)
