#!/usr/bin/env python
"""Turns ncu output (run on the B200 box, read here) into the small tracked files under
profiles/:  <tag>_launches.csv (every launch with its device time, from the
gpu__time_duration pass), <tag>_kernels.csv / .md (one row per profiled kernel from the
--set full report) and ncu_traffic.json (dram bytes per launch, read by bench.py).

    python tools/summarize_ncu.py r1 gpurun_out/launches.csv gpurun_out/prof.ncu-rep [fused|materialising]
"""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(REPO, 'profiles')

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_read'),
    ('dram__bytes_write.sum', 'dram_write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('smsp__inst_executed.sum', 'warp_insts'),
]


def to_bytes(v, unit):
    f = float(v.replace(',', ''))
    u = unit.lower()
    return f * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)


def to_ms(v, unit):
    f = float(v.replace(',', ''))
    return f * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1, 'msecond': 1, 'nsecond': 1e-6, 'second': 1e3, 's': 1e3}.get(unit.lower(), 1)


def short(name):
    name = name.replace('void ', '').replace('<unnamed>::', '')
    return name.split('(')[0]


def launches(tag, path, fused_only=True):
    rows = [r for r in csv.reader(open(path, errors='replace')) if r and not r[0].startswith('==')]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    kn, mv, mu, idc = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
    out = os.path.join(PROF, tag + '_launches.csv')
    tot = {}
    with open(out, 'w') as f:
        f.write('id,kernel,duration_ms\n')
        for r in rows[hdr + 1:]:
            if len(r) <= mv:
                continue
            # bench.py ends with a pass over the materialising path (for the stand-alone aggregation
            # kernel); the launch list and the shares are those of the product (fused) path only
            k0 = short(r[kn])
            if fused_only and (k0 == 'dm_tile_origin_kernel' or any(t in k0 for t in ('dm_minmax_rectify', 'umma_kernel<0', 'umma_kernel<(int)0', 'dm_planes_kernel'))):
                break
            ms = to_ms(r[mv], r[mu])
            k = short(r[kn])
            tot.setdefault(k, [0, 0.0])
            tot[k][0] += 1; tot[k][1] += ms
            f.write('%s,%s,%.6f\n' % (r[idc], k, ms))
    with open(os.path.join(PROF, tag + '_launch_shares.md'), 'w') as f:
        total = sum(v[1] for v in tot.values())
        f.write('| kernel | launches | total ms (ncu, cold cache, serialised) | share |\n|---|---|---|---|\n')
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.3f | %.1f %% |\n' % (k, n, ms, 100 * ms / total))
    print('wrote', out)


def kernels(tag, rep, path_name):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {'kernel': short(r[idx['Kernel Name']])}
        for m, k in METRICS:
            if m not in idx:
                continue
            v, u = r[idx[m]], units[idx[m]]
            if k == 'time':
                d[k + '_ms'] = round(to_ms(v, u), 4)
            elif k.startswith('dram_') and not k.endswith('pct'):
                d[k + '_bytes'] = to_bytes(v, u)
            else:
                d[k] = float(v.replace(',', '')) if v else None
        res.append(d)
    keys = ['kernel'] + [k for k in res[0] if k != 'kernel']
    with open(os.path.join(PROF, tag + '_kernels.csv'), 'w') as f:
        w = csv.DictWriter(f, keys)
        w.writeheader(); w.writerows(res)
    with open(os.path.join(PROF, tag + '_kernels.md'), 'w') as f:
        f.write('| kernel | ms | DRAM read GB | DRAM write GB | DRAM % | tensor % | SM % | L1/TEX % | issue % | regs | grid x block |\n|---|---|---|---|---|---|---|---|---|---|---|\n')
        for d in res:
            f.write('| `%s` | %.3f | %.3f | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %d x %d |\n' % (
                d['kernel'], d['time_ms'], d['dram_read_bytes'] / 1e9, d['dram_write_bytes'] / 1e9, d['dram_pct'], d['tensor_pct'],
                d['sm_pct'], d['l1tex_pct'], d['issue_pct'], d['regs'], d['grid'], d['block']))
    tp = os.path.join(PROF, 'ncu_traffic.json')
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    t = traffic.setdefault(path_name, {})
    seen = {}
    for d in res:
        name = d['kernel'].split('<')[0]
        seen[name] = max(seen.get(name, 0), d['dram_read_bytes'] + d['dram_write_bytes'])     # the largest launch of that kernel in THIS report
    t.update(seen)
    json.dump(traffic, open(tp, 'w'), indent=1, sort_keys=True)
    print('wrote', tag + '_kernels.csv/.md and ncu_traffic.json')


if __name__ == '__main__':
    tag = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] != '-':
        launches(tag, sys.argv[2])
    if len(sys.argv) > 3 and sys.argv[3] != '-':
        kernels(tag, sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else 'fused')
