"""Per-CUDA-source-line hot spots of an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [file-substring] [top-n] [sort: samples|inst]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = sys.argv[4] if len(sys.argv) > 4 else "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            recs.append((cur, int(r[0]), r[1].strip(), int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0),
                         int(d.get("stall_barrier") or 0), int(d.get("stall_long_sb") or 0), int(d.get("stall_wait") or 0)))
        except ValueError:
            pass
ts = sum(x[3] for x in recs) or 1; ti = sum(x[4] for x in recs) or 1
sel = [x for x in recs if filt in x[0]]
sel.sort(key=lambda x: -(x[3] if key == "samples" else x[4]))
print("total samples %d, total warp-instr %d; showing %s" % (ts, ti, filt or "all files"))
print("%-14s %5s %7s %7s %6s %6s %6s  %s" % ("file", "line", "smp%", "inst%", "barr%", "lsb%", "wait%", "source"))
for f, ln, src, smp, ins, sb, sl, sw in sel[:top]:
    print("%-14s %5d %7.2f %7.2f %6.0f %6.0f %6.0f  %s" % (f, ln, 100.0 * smp / ts, 100.0 * ins / ti, 100.0 * sb / max(smp, 1), 100.0 * sl / max(smp, 1), 100.0 * sw / max(smp, 1), src[:110]))
