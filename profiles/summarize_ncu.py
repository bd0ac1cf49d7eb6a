"""Turn an .ncu-rep (brought back in gpurun_out/) into the small text summary that is
committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/r1_x.txt ["note"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of %s" % rep.split("/")[-1]]
    if note:
        lines.append("# " + note)
    lines.append("# captured with: ncu --set full --clock-control none --import-source on (one GPU, under gpurun)")
    ik = hdr.index("Kernel Name")
    for n, vals in enumerate(rows[2:]):
        lines.append("")
        lines.append("launch %d: %s" % (n, vals[ik]))
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or "stall" in h and h.endswith("pct") or "tensor" in h and "realtime" in h and "pct" in h:
                lines.append("  %-92s %-10s %s" % (h, u, v))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    try:
        sh = next(r for r in srows if "Source" in r and "# Samples" in r)
        data = srows[srows.index(sh) + 1:]
        i_s, i_n, i_e = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
        data = [r for r in data if len(r) > max(i_s, i_n, i_e) and r[i_n].isdigit()]
        tot = sum(int(r[i_n]) for r in data) or 1
        lines.append("")
        lines.append("top SASS instructions by warp-stall samples (first launch), total samples %d:" % tot)
        for r in sorted(data, key=lambda r: -int(r[i_n]))[:14]:
            lines.append("  %6.2f%%  exec=%-10s %s" % (100.0 * int(r[i_n]) / tot, r[i_e], r[i_s].strip()[:100]))
        ops = {}
        for r in data:
            t = r[i_s].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[i_e] or 0)
        tot_i = sum(ops.values()) or 1
        lines.append("instruction mix (warp instructions): " +
                     ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:12]))
    except StopIteration:
        pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
