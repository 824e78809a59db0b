"""Warp-state sample shares per kernel phase from an `ncu --set full --import-source on` capture, without running
anything on the GPU: the SASS listing of `ncu -i X.ncu-rep --page source --csv --print-source sass` is cut at the
named barriers (BAR.SYNC) of the step kernel, and the PC samples and stall reasons between two barriers are summed.
A warp waiting at a barrier is sampled on the instruction AFTER the BAR, so a phase's `barrier` share is the wait for
the slowest warp of the PREVIOUS phase.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/phase_from_ncu.py sass.csv [min_share_percent]
"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
nxt = [i for i, r in enumerate(rows) if i > 0 and r and r[0] == "Kernel Name"]     # several launches: the first one only
if nxt:
    rows = rows[:nxt[0]]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
n_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 0                              # rows of the launch: instructions per row
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
total = sum(float(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", "| samples:", int(total))
start, seg = 0, []
for i, r in enumerate(data):
    if re.search(r"\bBAR\.(SYNC|RED|ARV)", r[ix["Source"]]):
        seg.append((start, i))
        start = i + 1
seg.append((start, len(data) - 1))
for a, b in seg:
    n = sum(float(r[ix["# Samples"]] or 0) for r in data[a:b + 1])
    if 100 * n / total < min_share:
        continue
    c = collections.Counter()
    for r in data[a:b + 1]:
        for h in stalls:
            c[h] += float(r[ix[h]] or 0)
    top = ", ".join("%s %.1f" % (k.replace("stall_", ""), 100 * v / total) for k, v in c.most_common(5))
    inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in data[a:b + 1])
    per_row = "  %5.1f warp-instr/row" % (inst / n_rows) if n_rows else ""
    print("SASS %5d-%5d  %5.1f %%%s  | %s" % (a, b, 100 * n / total, per_row, top))
