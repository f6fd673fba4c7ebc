#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest SASS regions by stall samples, mapped to source lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_pipe_tensor_op_imma.sum', 'sm__inst_executed_pipe_tensor_op_imma.avg.pct_of_peak_sustained_active']
for v in rows[2:]:
    print("== kernel:", v[h.index('Kernel Name')] if 'Kernel Name' in h else '?')
    for i, k in enumerate(h):
        if k in keys: print("  %-75s %-12s %s" % (k, u[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ins = h.index('Instructions Executed'); smp = h.index('# Samples'); so = h.index('Source')
tot_i = sum(int(r[ins] or 0) for r in data); tot_s = sum(int(r[smp] or 0) for r in data)
print("total warp-instructions %d, samples %d, SASS lines %d" % (tot_i, tot_s, len(data)))
stall_cols = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
agg = collections.Counter()
for r in data:
    for i in stall_cols: agg[h[i]] += int(r[i] or 0)
print("stall reasons (all samples):", ", ".join("%s=%.1f%%" % (k, 100.0 * v / max(1, tot_s)) for k, v in agg.most_common(8)))
# windows of 32 SASS lines
win = 32
w = []
for i in range(0, len(data), win):
    blk = data[i:i + win]
    w.append((sum(int(r[smp] or 0) for r in blk), sum(int(r[ins] or 0) for r in blk), i))
w.sort(reverse=True)
print("hottest %d-line SASS windows: samples%%, inst%%, first line, dominant opcodes / stalls" % win)
for s, n, i in w[:ntop]:
    blk = data[i:i + win]
    ops = collections.Counter(r[so].split()[1 if r[so].strip().startswith('@') else 0].split('.')[0] for r in blk if r[so].strip())
    st = collections.Counter()
    for r in blk:
        for c in stall_cols: st[h[c]] += int(r[c] or 0)
    print("  %5.1f%% %5.1f%% @%5d  %s | %s" % (100.0 * s / max(1, tot_s), 100.0 * n / max(1, tot_i), i,
          " ".join("%s:%d" % kv for kv in ops.most_common(5)), " ".join("%s:%d" % kv for kv in st.most_common(3))))

# ---- per CUDA source line (needs -lineinfo) ----
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit() and len(r) >= len(hdr) - 2:
        try:
            s = int(r[hdr.index('# Samples')] or 0); n = int(r[hdr.index('Instructions Executed')] or 0)
        except ValueError:
            continue
        st = {}
        for i, k in enumerate(hdr):
            if k.startswith('stall_') and 'Not Issued' not in k and i < len(r) and r[i].isdigit() and int(r[i]): st[k] = int(r[i])
        lines.append((s, n, cur, int(r[0]), r[1].strip()[:90], st))
lines.sort(key=lambda x: -x[0])
print("hottest source lines: samples%, inst%, file:line, text | top stalls")
for s, n, f, ln, txt, st in lines[:ntop]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("  %5.1f%% %5.1f%%  %s:%d  %s | %s" % (100.0 * s / max(1, tot_s), 100.0 * n / max(1, tot_i), f, ln, txt, " ".join("%s:%d" % kv for kv in top)))
