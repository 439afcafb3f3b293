# -*- coding: utf-8 -*-
"""Summarises the ncu exports of profiles/ncu_run.sh:

    python profiles/ncu_parse.py gpurun_out r02

  * launches_<tag>.csv / launches_train_<tag>.csv  ->  profiles/<tag>_launches_*.md : per-kernel time shares of one eager step
  * conv_<tag>_raw.csv / train_<tag>_raw.csv (`ncu -i ... --page raw --csv`)  ->  profiles/<tag>_ncu_full_*.md : duration, DRAM bytes,
    DRAM / tensor-pipe / L1 / L2 utilisation, registers, achieved occupancy per captured launch
  * profiles/ncu_traffic.json : dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels bench.py names in `roofline.kernel`
    (bench.py reads `roofline.traffic` from this file).
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict


def read_csv(path):
    rows, header = [], None
    with open(path, newline='') as f:
        for row in csv.reader(f):
            if header is None:
                if 'Kernel Name' in row:
                    header = row
                continue
            if len(row) == len(header) and (not row[0] or row[0].strip().isdigit()) and row[header.index('Kernel Name')]:
                rows.append(dict(zip(header, row)))      # (the raw page has a units row under the header: no kernel name)
    return rows


def fnum(v):
    try:
        return float(str(v).replace(',', ''))
    except Exception:
        return float('nan')


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name.replace('lfd::', '')


def launches(path, out, title):
    rows = read_csv(path)
    if not rows:
        return
    per = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        ns = fnum(r['Metric Value'])
        unit = r.get('Metric Unit', 'ns')
        ns *= {'ns': 1.0, 'us': 1e3, 'usecond': 1e3, 'nsecond': 1.0, 'ms': 1e6, 'msecond': 1e6}.get(unit, 1.0)
        k = short(r['Kernel Name'])
        e = per.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += ns
        total += ns
    with open(out, 'w') as f:
        f.write('# %s\n\nncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES, not absolutes)\n\n' % title)
        f.write('| kernel | launches | total us | share |\n|---|---|---|---|\n')
        for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f %% |\n' % (k, n, ns / 1e3, 100 * ns / total))
        f.write('\ntotal %.1f us in %d launches\n' % (total / 1e3, sum(n for n, _ in per.values())))


METRICS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
           ('FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'), ('dram__bytes.sum.per_second', 'DRAM B/s'), ('sm__inst_executed_pipe_tensor.sum', 'tensor inst'),
           ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe %'),
           ('SM_A.TriageCompute.l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1 %'), ('LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
           ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
           ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %')]


def full(path, out, title):
    rows = read_csv(path)
    if not rows:
        return []
    table = []
    for r in rows:
        e = OrderedDict(kernel=short(r['Kernel Name']), id=r.get('ID', ''), grid=r.get('Grid Size', ''), block=r.get('Block Size', ''))
        for m, label in METRICS:
            if m in r:
                e[label] = r[m]
        table.append(e)
    with open(out, 'w') as f:
        f.write('# %s\n\nncu --set full --clock-control none, one row per captured launch (values as exported by `ncu --page raw --csv`)\n\n' % title)
        cols = ['kernel', 'grid', 'block'] + [l for _, l in METRICS if any(l in e for e in table)]
        f.write('| ' + ' | '.join(cols) + ' |\n|' + '---|' * len(cols) + '\n')
        for e in table:
            f.write('| ' + ' | '.join(str(e.get(c, '')) for c in cols) + ' |\n')
    return table


def main():
    src, tag = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.abspath(__file__))
    for name, title in (('launches_%s.csv' % tag, 'launch list of one eager inference step (WIDERFACE-S 720p b8)'),
                        ('launches_train_%s.csv' % tag, 'launch list of one eager training step (WIDERFACE-L 640x640, 16 crops)')):
        p = os.path.join(src, name)
        if os.path.exists(p):
            launches(p, os.path.join(here, '%s_%s.md' % (tag, name[:-len('_%s.csv' % tag)])), title)
    traffic_path = os.path.join(here, 'ncu_traffic.json')
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    p = os.path.join(src, 'conv_%s_raw.csv' % tag)
    if os.path.exists(p):
        t = full(p, os.path.join(here, '%s_ncu_full_inference.md' % tag), 'full captures: first conv launches of the inference step (op 0 stem0+tail, op 1 3x3/s2+tail, op 2, op 3)')
        # the launches are captured in plan order: op 0 = the stem conv, op 1 = 3x3/s2 64->64 + tail @180x320 (the dominant kernel)
        names = ['stem0 3x3/s2 3->64 @360x640', 'conv 3x3/s2 64->64 @180x320', 'conv 3x3/s2 64->64 @90x160', 'conv 3x3/s1 64->64 @90x160']
        key = traffic.setdefault('WIDERFACE_S/bf16', {})
        for e, n in zip(t, names):
            if 'dram read' in e:
                key[n] = dict(dram_bytes_read=fnum(e['dram read']), dram_bytes_write=fnum(e['dram write']), source='profiles/%s_ncu_full_inference.md' % tag)
    p = os.path.join(src, 'train_%s_raw.csv' % tag)
    if os.path.exists(p):
        full(p, os.path.join(here, '%s_ncu_full_training.md' % tag), 'full captures: weight-gradient / norm-backward / head-backward / BatchNorm-statistics launches of the training step')
    json.dump(traffic, open(traffic_path, 'w'), indent=1, sort_keys=True)
    print('wrote', traffic_path)


if __name__ == '__main__':
    main()
