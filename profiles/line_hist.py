"""Per-CUDA-source-line instruction / stall-sample shares of one kernel from a .ncu-rep.
usage: python profiles/line_hist.py <report.ncu-rep> <kernel regex> [min share %]"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# cuda,sass view: a CUDA line row (Line No, Source, metrics...) followed by its SASS rows; use only rows whose first field is a line number
hdr = None; inst = collections.Counter(); samp = collections.Counter(); thrd = collections.Counter(); src = {}
seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1: break
        continue
    if r[0] in ("Line No", "#", "Line"):
        hdr = r; continue
    if hdr is None or not r[0].isdigit() or len(r) < len(hdr): continue
    try:
        iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iS = hdr.index("# Samples")
    except ValueError:
        continue
    if not (r[iI].isdigit() and r[iS].isdigit() and r[iT].isdigit()): continue
    ln = int(r[0]); inst[ln] += int(r[iI]); samp[ln] += int(r[iS]); thrd[ln] += int(r[iT]); src[ln] = r[1].strip()
tot = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
print("total warp inst", tot, "samples", ts)
for ln in sorted(inst):
    if 100.0 * inst[ln] / tot >= thr or 100.0 * samp[ln] / ts >= thr:
        print("%5d %5.1f%% inst %5.1f%% samp %5.1f thr | %s" % (ln, 100.0 * inst[ln] / tot, 100.0 * samp[ln] / ts, thrd[ln] / max(1, inst[ln]), src[ln][:100]))
