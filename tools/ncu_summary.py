"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small table for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("##", r[idx["Kernel Name"]].split("(")[0], "grid", r[idx.get("launch__grid_size", 0)])
        for k in KEYS:
            if k in idx:
                print(f"  {k} = {r[idx[k]]} {units[idx[k]]}")
        st = []
        for h, i in idx.items():
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    st.append((float(r[i] or 0), h))
                except ValueError:
                    pass
        for v, h in sorted(st, reverse=True)[:6]:
            print(f"  stall {h.split('issue_stalled_')[-1]} = {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
