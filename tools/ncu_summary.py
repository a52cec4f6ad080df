#!/usr/bin/env python
"""Summarises an .ncu-rep: key raw metrics + executed instructions by opcode per warp-row."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; npart = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
print("kernel:", r[hdr.index('Kernel Name')][:60])
for k in want:
    if k in hdr: print("  %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
for i, h in enumerate(hdr):
    if re.search(r'smsp__average_warps_issue_stalled.*ratio$', h) and float(r[i] or 0) > 0.25:
        print("  stall %-30s %s" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; iS, iE, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
agg = collections.Counter(); nrow = npart / 32
hot = []
for rr in rows[2:end]:
    if len(rr) <= iE or not rr[iE].isdigit(): continue
    op = re.sub(r'^@!?U?P\d+\s+', '', rr[iS].strip()).split()[0].split('.')[0]
    agg[op] += int(rr[iE]); hot.append((int(rr[iE]) / nrow, int(rr[iSm] or 0), rr[iS].strip()))
print("  total warp-instructions per 32-particle row: %.1f" % (sum(agg.values()) / nrow))
print("  " + "  ".join("%s %.1f" % (o, c / nrow) for o, c in agg.most_common(30)))
if len(sys.argv) > 3:
    for e, sm, s in hot:
        if e > 0.4: print("%6.2f %5d  %s" % (e, sm, s))
