"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the small tables kept under profiles/.
   python tools/ncu_summary.py raw.csv out.csv [--per-kernel]     (--per-kernel: one row per launch, a fixed metric set)
Without --per-kernel: every metric of the FIRST launch whose name matches the interesting families (one metric per row)."""
import csv
import sys

PER_KERNEL = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
FAMILIES = ("gpu__time_duration.sum", "dram__bytes", "gpu__dram_throughput", "sm__throughput", "sm__pipe_tensor", "sm__inst_executed_pipe_fma.",
            "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu.avg", "sm__warps_active", "smsp__issue_active", "launch__", "hit_rate",
            "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_uniform",
            "smsp__inst_executed_pipe_tensor", "sm__sass_inst_executed_op_shared", "l1tex__data_bank_conflicts")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        if "--per-kernel" in sys.argv:
            keep = [k for k in PER_KERNEL if k in idx]
            w.writerow(keep)
            w.writerow([units[idx[k]] for k in keep])
            for r in data:
                w.writerow([r[idx[k]][:100] for k in keep])
        else:
            r = data[0]
            w.writerow(["metric", "unit", "value", "kernel: " + r[idx["Kernel Name"]][:120]])
            for h in hdr:
                if any(h.startswith(p) or p in h for p in FAMILIES) and not any(s in h for s in (".max.", ".min.", ".sum.pct", ".sum.per_second")):
                    v = r[idx[h]]
                    if v not in ("", "0", "0.000000"):
                        w.writerow([h, units[idx[h]], v])


if __name__ == "__main__":
    main()
