"""Top source lines by warp-stall samples of the first kernel in an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r): hdr = i; break
if hdr is None: print("no source table; header candidates:", rows[:3]); sys.exit(0)
H = rows[hdr]; si = H.index("Source"); ci = [k for k, c in enumerate(H) if "Sampling (All" in c or c.strip() == "# Samples"]
ci = ci[0] if ci else [k for k, c in enumerate(H) if "Sampl" in c][0]
agg = collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= max(si, ci): continue
    if r and "Source" in r: break        # next kernel
    try: agg[r[si].strip()[:150]] += float(r[ci] or 0)
    except ValueError: pass
tot = sum(agg.values()) or 1
for src, v in agg.most_common(n): print("%6.2f%%  %s" % (100 * v / tot, src))
