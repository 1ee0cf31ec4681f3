#!/usr/bin/env python
"""Turns ncu reports (gpurun_out/*.ncu-rep) and the launch list into the committed text summaries.
usage: python profiles/summarize.py raw <rep> <out.md> | launches <launches.csv> <out.md>"""
import csv, subprocess, sys, collections

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum']


def raw(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none: %s\n\n' % rep.split('/')[-1])
        for r in data:
            f.write('## %s\n\n| metric | value | unit |\n|---|---|---|\n' % r[idx['Kernel Name']].strip())
            for k in KEYS:
                if k in idx and r[idx[k]] != '':
                    f.write('| %s | %s | %s |\n' % (k, r[idx[k]], units[idx[k]]))
            rd = float(r[idx['dram__bytes_read.sum']]); wr = float(r[idx['dram__bytes_write.sum']])
            ur, uw = units[idx['dram__bytes_read.sum']], units[idx['dram__bytes_write.sum']]
            sc = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            tot = rd * sc[ur] + wr * sc[uw]
            t = float(r[idx['gpu__time_duration.sum']]) * {'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1}[units[idx['gpu__time_duration.sum']]]
            f.write('| **traffic = dram read + write** | %.1f | MB |\n| **traffic / duration (under ncu, cold, serialised)** | %.0f | GB/s |\n\n' % (tot / 1e6, tot / t / 1e9))


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); mu = h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[mu], 1)
        name = r[kn].split('(')[0].replace('void ', '').replace('ctr::', '').strip()
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none): %d launches, %.1f us total\n\n' % (sum(a[0] for a in agg.values()), tot))
        f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py\'s "kernels" object, not absolutes.\n\n')
        f.write('| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| %s | %d | %.1f | %.1f | %.3f |\n' % (k, n, t, t / n, t / tot))
        # the train step's own kernels (set-up: table fill, id maps, L2 flush fills, key resolve are not part of a step)
        step = collections.OrderedDict((k, v) for k, v in agg.items() if any(k.startswith(p) for p in
                                       ('k_attn', 'umma::', 'k_head', 'k_hot_apply', 'k_adam', 'k_peer_barrier', 'k_apply_table', 'k_segment', 'k_scatter_keys')))
        if step:
            st = sum(a[1] for a in step.values())
            f.write('\n## share inside the train step (set-up kernels excluded)\n\n| kernel | launches | avg us | share of step |\n|---|---|---|---|\n')
            for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
                f.write('| %s | %d | %.1f | %.3f |\n' % (k, n, t / n, t / st))


if __name__ == '__main__':
    {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
