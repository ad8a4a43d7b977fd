"""One-page summary of an `ncu --set full` capture (read here, on the CPU box): duration, DRAM traffic, tensor / issue /
pipe utilisation, L2->SM bytes, registers, top stall reasons.  Usage: python tools/ncu_summary.py file.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"]
print("ncu capture: %s" % rep.split("/")[-1])
for k in want:
    if k in d:
        print("%-72s %s %s" % (k, d[k][1], d[k][0]))
stalls = sorted(((float(v.replace(",", "")), h) for h, (u, v) in d.items()
                 if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v.replace(",", "").replace(".", "").isdigit()), reverse=True)
tot = sum(s for s, _ in stalls) or 1
print("warp-state samples (all warps, incl. the spinning role warps):")
for s, h in stalls[:10]:
    print("  %-40s %6.0f  %4.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), s, 100 * s / tot))
try:
    r, w = float(d["dram__bytes_read.sum"][1].replace(",", "")), float(d["dram__bytes_write.sum"][1].replace(",", ""))
    mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    print("DRAM traffic = %.1f MB" % ((r * mult[d["dram__bytes_read.sum"][0]] + w * mult[d["dram__bytes_write.sum"][0]]) / 1e6))
except Exception:
    pass
