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


if __name__ == "__main__":
    main(sys.argv[1], "--launches" in sys.argv)
