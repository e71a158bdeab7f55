"""Aggregate an .ncu-rep's per-source-line instruction / stall-sample counts into regions of vp_splat.cu (or any file):
usage: python scripts/ncu_regions.py REP FILE 'name:lo-hi,name:lo-hi,...'"""
import csv, subprocess, sys, io, collections
rep, fname, spec = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for part in spec.split(','):
    n, r = part.split(':'); lo, hi = r.split('-'); regions.append((n, int(lo), int(hi)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h2 = None
agg = collections.defaultdict(lambda: [0, 0]); other = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h2 = r; ix = {h: i for i, h in enumerate(h2)}; continue
    if h2 and len(r) == len(h2) and r[0].isdigit():
        try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
        except Exception: continue
        ln = int(r[0])
        if cur == fname:
            for name, lo, hi in regions:
                if lo <= ln <= hi: agg[name][0] += n; agg[name][1] += s; break
            else: agg['(unassigned %s)' % fname][0] += n; agg['(unassigned %s)' % fname][1] += s
        else:
            other[cur][0] += n; other[cur][1] += s
tot = sum(v[0] for v in agg.values()) + sum(v[0] for v in other.values()); ts = sum(v[1] for v in agg.values()) + sum(v[1] for v in other.values())
print("total warp instructions %d, samples %d" % (tot, ts))
for k, v in list(agg.items()) + list(other.items()):
    print("%-34s %6.1f%% inst (%9d)  %6.1f%% samples" % (k, 100 * v[0] / tot, v[0], 100 * v[1] / ts))
