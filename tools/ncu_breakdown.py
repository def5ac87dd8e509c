#!/usr/bin/env python3
"""Summarise an .ncu-rep of the step kernel: key raw metrics + executed instructions / stall samples by source line and stage.
usage: ncu_breakdown.py report.ncu-rep n_envs [top]"""
import bisect, csv, os, subprocess, sys
rep, nenv = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio' , 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
for w in want:
    if w in hdr:
        i = hdr.index(w); print('%-90s %s %s' % (w, rows[2][i], rows[1][i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur, data = None, []
for r in csv.reader(src.splitlines()):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) > 8 and r[0].isdigit():
        try: data.append((cur.split('/')[-1], int(r[0]), r[1], int(r[7]), int(r[6]) if r[6].isdigit() else 0))
        except ValueError: pass
tot = sum(d[3] for d in data); ts = max(1, sum(d[4] for d in data))
print('\ntotal warp instructions %d -> %.0f per env-step' % (tot, tot / nenv))
lines = open(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'csrc', 'step_core.inl')).read().split('\n')
marks = [(i + 1, l.strip()) for i, l in enumerate(lines) if '=================' in l or l.startswith('template') or 'CFN void' in l or l.strip().startswith('// ----')]
agg, samp = {}, {}
for f, ln, s, inst, sm in data:
    if f != 'step_core.inl': name = f
    else:
        k = bisect.bisect_right([m[0] for m in marks], ln) - 1; name = marks[k][1][:90] if k >= 0 else 'top'
    agg[name] = agg.get(name, 0) + inst; samp[name] = samp.get(name, 0) + sm
print('\ninst/env  share  stall-samples  stage')
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:28]: print('%8.0f %5.1f%% %5.1f%%  %s' % (v / nenv, 100 * v / tot, 100 * samp[k] / ts, k))
print('\ntop source lines')
for f, ln, s, inst, sm in sorted(data, key=lambda d: -d[3])[:top]: print('%-14s %5d %8.0f inst/env  samp %4.1f%%  %s' % (f[:14], ln, inst / nenv, 100 * sm / ts, s[:110]))
