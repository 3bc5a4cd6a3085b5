#!/usr/bin/env python
"""Writes profiles/traffic.json -- the `roofline.traffic` figure bench.py attaches to the headline line -- from
`ncu --set full` captures of the two headline kernels: dram__bytes_read.sum + dram__bytes_write.sum per launch.

    python profiles/make_traffic.py <pull.ncu-rep> <push.ncu-rep> [note]

A profiler cannot run inside the timed region of bench.py, so the number is taken from a capture of the same
kernels on the same workload (profiles/time_ops.py: 256^3 fp32 cubic dct2, the bench's make_workload) and
the file records the commit and the reports it came from."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(key)
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
        tot += float(vals[i].replace(',', '')) * scale
    name = vals[hdr.index('Kernel Name')]
    dur = vals[hdr.index('gpu__time_duration.sum')] + ' ' + units[hdr.index('gpu__time_duration.sum')]
    return int(tot), name.split('(')[0], dur


def main():
    pull, push = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    head = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    bp, np_, dp = dram_bytes(pull)
    bs, ns, ds = dram_bytes(push)
    doc = {
        '_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, 256^3 cubic fp32 dct2 '
                    '(profiles/time_ops.py = the workload of bench.py); written by profiles/make_traffic.py',
        'grid_pull': bp, 'grid_push': bs,
        'source': {'commit': head, 'pull_report': os.path.basename(pull), 'pull_kernel': np_.strip(), 'pull_duration_under_ncu': dp,
                   'push_report': os.path.basename(push), 'push_kernel': ns.strip(), 'push_duration_under_ncu': ds, 'note': note},
    }
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == '__main__':
    main()
