"""Opcode histogram of one kernel from a .ncu-rep (source page, SASS view).
usage: python profiles/sass_hist.py <report.ncu-rep> <kernel regex> [launch index]"""
import csv, collections, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if row and row[0] == "Address": cur["hdr"] = row; continue
    cur["rows"].append(row)
b = blocks[which]
hdr = b["hdr"]; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
ops = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
for row in b["rows"]:
    if len(row) <= iT or not row[iI].isdigit(): continue
    op = row[1].strip().split()
    if op and op[0].startswith("@"): op = op[1:]
    k = op[0].rstrip(",") if op else "?"
    ops[k] += int(row[iI]); samp[k] += int(row[iS]); thr[k] += int(row[iT])
tot, ts = sum(ops.values()), max(1, sum(samp.values()))
print(b["name"]); print("total warp inst %d, samples %d, %d launches matched" % (tot, ts, len(blocks)))
for k, v in ops.most_common(45):
    print("%-24s %6.2f%% inst  %6.2f%% samples  %5.1f thr/inst" % (k, 100 * v / tot, 100 * samp[k] / ts, thr[k] / max(1, v)))
