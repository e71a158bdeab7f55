"""Per-kernel source-line shares of an .ncu-rep: python scripts/ncu_lines2.py rep kernel_regex [top_n]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
cur = None; h2 = None; out = []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h2 = r; ix = {h: i for i, h in enumerate(h2)}; continue
    if h2 and len(r) == len(h2) and r[0].isdigit():
        try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
        except Exception: continue
        out.append((n, s, cur, int(r[0]), r[1].strip()[:120]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("total warp instructions %d, samples %d" % (tot, ts))
# shares by 20-line region too
reg = collections.Counter()
for n, s, f, l, t in out: reg[(f, l // 25 * 25)] += n
print("== by 25-line region")
for (f, l), n in sorted(reg.items(), key=lambda kv: -kv[1])[:20]: print("%5.1f%%  %s:%d-%d" % (100 * n / tot, f, l, l + 24))
print("== per line")
for n, s, f, l, t in sorted(out, reverse=True)[:topn]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * n / tot, 100 * s / ts, f, l, t))
