#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the handful of numbers DESIGN.md / profiles/*.md quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections, csv, io, re, subprocess, sys

def page(rep, name, extra=()):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout

def main(rep):
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
    for k in keys:
        if k in m: print("%-75s %18s %s" % (k, m[k], u.get(k, "")))
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(m[k]) >= 0.05: print("%-75s %18s" % (k.replace("smsp__average_warps_issue_stalled_", "stall:"), m[k]))
            except ValueError: pass
    rows = list(csv.reader(io.StringIO(page(rep, "source", ["--print-source", "sass"]))))
    hi = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line No", "#")]
    if not hi: return
    hdr = rows[hi[0]]; ci = {h: i for i, h in enumerate(hdr)}
    si = ci.get("Source"); ei = ci.get("Instructions Executed"); pi_ = ci.get("# Samples")
    byop = collections.Counter(); samp = collections.Counter(); tot = 0
    for r in rows[hi[0] + 1:]:
        try: ex = int(r[ei]); sm = int(r[pi_])
        except Exception: continue
        t = r[si].strip().split()
        if not t: continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += ex; samp[op] += sm; tot += ex
    print("instructions executed (warp-level):", tot)
    ts = sum(samp.values()) or 1
    for op, c in byop.most_common(22): print("  %-8s %6.2f%% of instructions   %6.2f%% of stall samples" % (op, 100.0 * c / tot, 100.0 * samp[op] / ts))

if __name__ == "__main__":
    main(sys.argv[1])
