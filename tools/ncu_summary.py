"""Summarise an ncu report of ransac_pairs_kernel into a text file under profiles/:
key raw metrics + per-device-function shares of executed instructions and stall samples.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt [n_pairs]"""
import bisect, csv, io, re, subprocess, sys, os

rep, out = sys.argv[1], sys.argv[2]
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else None
lib = os.path.abspath(sys.argv[4]) if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pydegensac_b200", "libdegensac_b200.so")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
lines = ["ncu report: %s" % os.path.basename(rep), "kernel: %s" % (vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"), ""]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append("%-70s %-12s %s" % (w, units[i], vals[i]))
lines.append("")
lines.append("warp stall reasons per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
for i, h in enumerate(hdr):
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
    if m and float(vals[i] or 0) >= 0.05:
        lines.append("  %-28s %s" % (m.group(1), vals[i]))
if npairs and "smsp__inst_executed.sum" in hdr:
    lines.append("")
    lines.append("warp instructions per image pair: %.2f M" % (float(vals[hdr.index("smsp__inst_executed.sum")]) / npairs / 1e6))
    rd = float(vals[hdr.index("dram__bytes_read.sum")]); wr = float(vals[hdr.index("dram__bytes_write.sum")])
    ur = units[hdr.index("dram__bytes_read.sum")]; uw = units[hdr.index("dram__bytes_write.sum")]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = rd * mult.get(ur, 1) + wr * mult.get(uw, 1)
    lines.append("DRAM traffic per launch: %.1f MB  (%.0f B per pair; algorithmic 66088 B per pair)" % (tot / 1e6, tot / npairs))
# per-function breakdown
try:
    tmp = "/tmp/_ncu_sum"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run("cd %s && rm -f *.cubin && cuobjdump -xelf all %s > /dev/null 2>&1" % (tmp, lib), shell=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and f.startswith("degensac_b200.")][0]
    sym = subprocess.run(["readelf", "-sW", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    kind = os.environ.get("NCU_KIND") or ("ILi1E" if "(int)1" in kname else "ILi0E")
    syms = []
    for l in sym.splitlines():
        f = l.split()
        if len(f) >= 8 and f[3] == "FUNC" and kind in f[-1]:
            m = re.search(r"\$_ZN2dg\d+([A-Za-z_0-9]+?)E", f[-1])
            syms.append((int(f[1], 16), int(f[2], 0), m.group(1) if m else f[-1][-32:]))
    syms.sort()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h2 = srows[1]; ia = h2.index("Address"); ie = h2.index("Instructions Executed"); isamp = h2.index("# Samples")
    data = srows[2:]
    base = int(data[0][ia], 16)
    starts = [s[0] for s in syms]
    reasons = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_no_inst", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_dispatch", "stall_not_selected", "stall_selected"]
    ridx = [h2.index(x) for x in reasons]
    rs = {}
    agg, smp = {}, {}
    for r in data:
        off = int(r[ia], 16) - base
        k = bisect.bisect_right(starts, off) - 1
        name = "kernel body (staging, dispatch)"
        if k >= 0 and off < syms[k][0] + syms[k][1]:
            name = syms[k][2]
        agg[name] = agg.get(name, 0) + int(r[ie] or 0)
        smp[name] = smp.get(name, 0) + int(r[isamp] or 0)
        v = rs.setdefault(name, [0] * len(reasons))
        for q, ix in enumerate(ridx): v[q] += int(r[ix] or 0)
    tot = sum(agg.values()); ts = max(1, sum(smp.values()))
    lines.append("")
    lines.append("device function                executed warp-instr share   (M per pair)   stall-sample share")
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:26]:
        lines.append("  %-30s %6.2f %%   %10s   %6.1f %%" % (k, 100.0 * v / tot, ("%.2f" % (v / npairs / 1e6)) if npairs else "-", 100.0 * smp[k] / ts))
    lines.append("")
    lines.append("stall samples by reason (%% of the function's samples) and cycles per issued warp-instruction")
    lines.append("  %-30s %s  cyc/instr" % ("function", " ".join("%8s" % x.replace("stall_", "")[:8] for x in reasons)))
    for k, v in sorted(agg.items(), key=lambda x: -smp[x[0]])[:26]:
        t = max(1, sum(rs[k]))
        sel = max(1, rs[k][reasons.index("stall_selected")])
        lines.append("  %-30s %s  %8.1f" % (k, " ".join("%8.1f" % (100.0 * x / t) for x in rs[k]), t / sel))
except Exception as ex:
    import traceback; traceback.print_exc()
    lines.append("(per-function breakdown unavailable: %r)" % (ex,))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
