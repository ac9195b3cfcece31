"""Prints the metrics the design notes quote from an ncu report: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__cluster_max_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if flt and not flt.search(r[ki]):
        continue
    print("kernel:", r[ki][:150])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-86s %s %s" % (w, r[i], units[i]))
    stalls = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h) or \
            re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
        if m:
            try:
                stalls.append((float(r[i].replace(",", "")), m.group(1)))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  warp stall reasons (cycles per issued instruction):")
    for v, n in stalls[:10]:
        print("    %-60s %.2f" % (n, v))
    print()
