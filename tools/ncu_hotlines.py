"""Top source lines of a kernel by warp-stall samples from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_hotlines.py <file.ncu-rep> [N] [kernel regex] [launch index among the matches]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sel = (["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []) + (["--launch-skip", sys.argv[4], "-c", "1"] if len(sys.argv) > 4 else [])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []; path = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; wi = r.index("Warp Stall Sampling (All Samples)"); ii = r.index("Instructions Executed"); continue
    if hdr is None or len(r) <= wi or r[2] != "-": continue        # keep the per-source-line aggregate rows only
    try: data.append((int(r[wi]), int(r[ii]), path, r[0], r[1].strip()[:120]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print("total stall samples", tot)
for d in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%%  inst=%-10d %s:%s | %s" % (d[0], 100 * d[0] / tot, d[1], d[2], d[3], d[4]))
