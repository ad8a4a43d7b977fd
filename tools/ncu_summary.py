"""Print the metrics we track from an .ncu-rep (raw page) and, with --source, the hottest source lines by stall samples.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source [N]]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu", "smsp__issue_active.avg.pct",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")])
    for h, v in zip(hdr, r):
        if any(h.startswith(k) for k in KEYS) and not h.endswith((".min", ".max")) and ".min." not in h and ".max." not in h:
            try:
                if float(v.replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            print("  %-90s %s" % (h, v))
if "--source" in sys.argv:
    i = sys.argv.index("--source")
    top = int(sys.argv[i + 1]) if len(sys.argv) > i + 1 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, data = "", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
            d = dict(zip(hdr, r))
            n = int(float(d["# Samples"] or 0))
            if n:
                stalls = sorted(((int(float(v or 0)), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k),
                                reverse=True)[:3]
                data.append((n, fname, r[0], r[1], ", ".join("%s=%d" % (k[6:], v) for v, k in stalls if v)))
    tot = sum(d[0] for d in data) or 1
    print("== source hot spots, total samples %d" % tot)
    for n, fn, ln, src, st in sorted(data, reverse=True)[:top]:
        print("  %6d %5.1f%%  %s:%-4s %-90s %s" % (n, 100.0 * n / tot, fn[:14], ln, src.strip()[:90], st))
