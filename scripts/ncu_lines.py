"""Per-source-line warp instructions / stall samples of one file region of an .ncu-rep.
usage: python scripts/ncu_lines.py REP FILE LO HI"""
import csv, subprocess, sys, io
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h2 = None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h2 = r; ix = {h: i for i, h in enumerate(h2)}; continue
    if h2 and len(r) == len(h2) and r[0].isdigit() and cur == fname:
        try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
        except Exception: continue
        ln = int(r[0])
        if lo <= ln <= hi and (n or s): print("%5d %10d %6d  %s" % (ln, n, s, r[1].strip()[:120]))
