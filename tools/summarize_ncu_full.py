"""Summarise one kernel of an `ncu --set full` report (run here: `ncu -i x.ncu-rep --page raw --csv | python tools/summarize_ncu_full.py`):
duration, DRAM bytes, tensor-pipe / L2 / DRAM utilisation, registers, shared memory, and the issue-stall breakdown."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum"]


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        get = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("n/a", "")
        print("kernel: %s   grid %s block %s" % (get("Kernel Name")[0], get("Grid Size")[0], get("Block Size")[0]))
        for k in KEYS:
            v, u = get(k)
            print("  %-72s %16s %s" % (k, v, u))
        stalls = []
        for i, k in enumerate(hdr):
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp issue-stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))


if __name__ == "__main__":
    main()
