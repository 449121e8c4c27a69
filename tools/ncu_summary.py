#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX — condense an `ncu --set full` report into the evidence committed under
profiles/: OUT_PREFIX_metrics.csv (selected raw metrics, one row per captured launch) and OUT_PREFIX_summary.md."""
import csv, io, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = ["Kernel Name"] + [k for k in KEEP if k in ix]
    with open(out + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[ix[c]] for c in cols])
        for r in data:
            w.writerow([r[ix[c]] for c in cols])

    def val(r, k, default=float("nan")):
        try:
            return float(r[ix[k]])
        except Exception:
            return default

    def to_bytes(r, k):
        v, u = val(r, k), units[ix[k]].lower() if k in ix else ""
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    def to_ms(r):
        v, u = val(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]].lower()
        return v * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)

    total = sum(to_ms(r) for r in data)
    with open(out + "_summary.md", "w") as f:
        f.write(f"ncu --set full --clock-control none, report {rep.split('/')[-1]}: {len(data)} launches, {total:.3f} ms summed (cold-cache, serialised)\n\n")
        f.write("| kernel | ms | share | regs | grid x block | fp64 pipe % | tensor pipe % | smem wavefronts % | issue % | DRAM rd+wr MB | DRAM GB/s | L2 hit % |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in data:
            ms = to_ms(r)
            dram = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
            f.write(f"| `{name}` | {ms:.3f} | {100 * ms / total:.1f}% | {val(r, 'launch__registers_per_thread'):.0f} | {val(r, 'launch__grid_size'):.0f} x {val(r, 'launch__block_size'):.0f} | "
                    f"{val(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'):.1f} | {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{dram / 1e6:.1f} | {dram / 1e9 / (ms * 1e-3):.0f} | {val(r, 'lts__t_sector_hit_rate.pct'):.1f} |\n")
    print(open(out + "_summary.md").read())


if __name__ == "__main__":
    main()
