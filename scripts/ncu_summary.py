"""Summarise an .ncu-rep: headline metrics + instruction/stall share per CUDA source line.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warp_latency_issue_stalled_barrier.pct','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, r):
        if h in keys: print("  %-85s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h2 = None; out = []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h2 = r; ix = {h: i for i, h in enumerate(h2)}; continue
    if h2 and len(r) == len(h2) and r[0].isdigit():
        try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
        except Exception: continue
        out.append((n, s, cur, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print("== per source line (share of warp instructions / of stall samples)")
for n, s, f, l, t in sorted(out, reverse=True)[:topn]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * n / tot, 100 * s / ts, f, l, t))
print("== top by samples")
for n, s, f, l, t in sorted(out, key=lambda o: -o[1])[:12]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * n / tot, 100 * s / ts, f, l, t))
