#!/usr/bin/env python3
"""Writes the static kernel counts bench.py quotes in its roofline object from an .ncu-rep of one single-tick launch of the step kernel.
usage: ncu_counts.py report.ncu-rep n_envs config [existing.json]  -> JSON on stdout ({config: {warp_inst_per_env_step, dram_bytes_per_launch, ...}})"""
import csv, json, subprocess, sys
rep, nenv, cfg = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else {}
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, unit, val = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(val[i].replace(',', '')); u = unit[i]
    return v * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1}.get(u, 1)
inst = get('smsp__inst_executed.sum')
out[cfg] = {'warp_inst_per_env_step': round(inst / nenv, 1), 'dram_bytes_per_launch': get('dram__bytes_read.sum') + get('dram__bytes_write.sum'), 'envs': nenv,
            'issue_active_pct_under_ncu': get('smsp__issue_active.avg.pct_of_peak_sustained_active'), 'kernel': val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else None,
            'source': 'smsp__inst_executed.sum / envs of %s (ncu --set full, one single-tick launch; profiles/r2_step_kernel_plain_ncu_summary.md)' % rep.split('/')[-1]}
print(json.dumps(out, indent=1))
