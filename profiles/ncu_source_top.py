#!/usr/bin/env python3
"""Per-SASS-instruction hot spots of an .ncu-rep captured with --import-source on:
usage: python profiles/ncu_source_top.py report.ncu-rep [N]  -> top-N by stall samples, by shared wavefronts, opcode totals"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in body) or 1
tot_w = sum(f(r, "L1 Wavefronts Shared") for r in body) or 1
tot_i = sum(f(r, "Instructions Executed") for r in body) or 1
print("total samples %.0f, shared wavefronts %.0f, warp instructions %.0f" % (tot_s, tot_w, tot_i))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("\n== top by stall samples")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:N]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print("%5.2f%%  %-70s %s" % (100 * f(r, "# Samples") / tot_s, r[col["Source"]][:70], " ".join("%s=%.0f" % (s[6:], f(r, s)) for s in top)))
print("\n== top by shared wavefronts (actual / ideal)")
for r in sorted(body, key=lambda r: -f(r, "L1 Wavefronts Shared"))[:N]:
    print("%5.2f%%  %-70s %.0f / %.0f" % (100 * f(r, "L1 Wavefronts Shared") / tot_w, r[col["Source"]][:70], f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal")))
print("\n== by opcode: instructions, samples, shared wavefronts")
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in body:
    t = r[col["Source"]].split()
    op = next((x for x in t if not x.startswith("@")), "?").split(".")[0]
    a = agg[op]; a[0] += f(r, "Instructions Executed"); a[1] += f(r, "# Samples"); a[2] += f(r, "L1 Wavefronts Shared")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print("%-10s inst %5.1f%%  samples %5.1f%%  shared wavefronts %5.1f%%" % (op, 100 * a[0] / tot_i, 100 * a[1] / tot_s, 100 * a[2] / tot_w))
