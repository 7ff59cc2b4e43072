"""Top CUDA source lines by warp-stall samples for the first kernel of an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.Counter(); inst = collections.Counter(); fname = ""; seen_kernel = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if seen_kernel is None: seen_kernel = r[1]
        elif r[1] != seen_kernel: break
        continue
    if r[0] == "Line No" or len(r) < 8 or not r[0].isdigit(): continue
    try: agg[(fname, int(r[0]), r[1].strip()[:110])] += float(r[4] or 0); inst[(fname, int(r[0]), r[1].strip()[:110])] += float(r[7] or 0)
    except ValueError: pass
tot = sum(agg.values()) or 1; ti = sum(inst.values()) or 1
print("kernel:", (seen_kernel or "")[:100], " total samples", int(tot))
for k, v in agg.most_common(n): print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100 * v / tot, 100 * inst[k] / ti, k[0], k[1], k[2]))
