"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py <rep> [--source N]"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        # SURVEY.md §8d evidence list: L2 sectors by operation, hit/miss, persisting (evict_last) vs normal lines, atomics
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_write_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_atom_dot_alu_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_atom_evict_last_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_atom_evict_last_lookup_miss.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_global_red.sum", "smsp__issue_active.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:80])
        for w in WANT:
            if w in hdr:
                print(f"  {w:95s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if "Source" in r and "# Samples" in " ".join(r) or (hdr is None and "Source" in r):
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if not hdr:
        print(out[:2000]); return
    cs = [i for i, h in enumerate(hdr) if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)"]
    si = hdr.index("Source")
    key = cs[0] if cs else None
    print("columns:", hdr)
    def val(r):
        try: return float(r[key])
        except Exception: return 0.0
    tot = sum(val(r) for r in data) or 1.0
    for r in sorted(data, key=val, reverse=True)[:top]:
        print(f"{val(r):9.0f} {100*val(r)/tot:5.1f}%  L{r[0]:>4s}  {r[si][:150]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))
    else:
        raw(rep)
