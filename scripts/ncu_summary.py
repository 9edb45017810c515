"""Summarise an .ncu-rep: headline metrics + executed instructions per source line (per 32-record step)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1048576.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor']
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"{k:75s} {vals[i]:>18s} {units[i]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 3: print(f"stall {h.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''):40s} {v:8.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
h = rows[hi]; ci = h.index('Instructions Executed'); si = h.index('# Samples')
out = []; tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ci or r[0] == '': continue
    try: n = int(r[ci])
    except ValueError: continue
    out.append((int(r[0]), r[1].strip(), n, int(r[si]) if r[si].isdigit() else 0)); tot += n
print(f"instructions executed {tot}  per step {tot/steps:.1f}  per event {tot/steps/32:.2f}")
for line, s, n, smp in sorted(out):
    if n >= steps * thr: print(f"{line:>4} {n/steps:7.1f} {smp:6d}  {s[:110]}")
