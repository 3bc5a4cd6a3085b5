#!/usr/bin/env python
"""Split the SASS of a profiled kernel at barriers / long branches and report
stall samples + executed instructions per segment (phase attribution).
usage: python profiles/ncu_phases.py file.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp, ia = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
iw = hdr.index('L1 Wavefronts Shared'); iwi = hdr.index('L1 Wavefronts Shared Ideal')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seg = dict(samples=0, instr=0, n=0, start=0, stalls={}, wf=0, wfi=0, ops={})
segs = []
def close(i, why):
    global seg
    seg['end'] = i; seg['why'] = why
    segs.append(seg)
    seg = dict(samples=0, instr=0, n=0, start=i + 1, stalls={}, wf=0, wfi=0, ops={})
for i, r in enumerate(rows[2:]):
    if len(r) <= iw: continue
    try: s = int(r[isamp]); n = int(r[ia])
    except ValueError: continue
    seg['samples'] += s; seg['instr'] += n; seg['n'] += 1
    try: seg['wf'] += int(r[iw]); seg['wfi'] += int(r[iwi])
    except ValueError: pass
    for c in stall_cols:
        try: v = int(r[c])
        except ValueError: v = 0
        if v: seg['stalls'][hdr[c]] = seg['stalls'].get(hdr[c], 0) + v
    toks = r[isrc].split()
    op = (toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')).split('.')[0]
    if n: seg['ops'][op] = seg['ops'].get(op, 0) + n
    if 'BAR.SYNC' in r[isrc] or 'EXIT' in r[isrc]:
        close(i, r[isrc].strip()[:30])
tot = sum(s['samples'] for s in segs) or 1
toti = sum(s['instr'] for s in segs) or 1
print('total samples %d, total warp instr %d' % (tot, toti))
for s in segs:
    if s['samples'] < 0.005 * tot and s['instr'] < 0.005 * toti: continue
    top = sorted(s['stalls'].items(), key=lambda kv: -kv[1])[:4]
    ops = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:6]
    print('SASS %5d-%5d  samples %5.1f%%  instr %5.1f%%  smem wf %d (ideal %d) | %s | %s | ends %s' % (
        s['start'], s['end'], 100.0 * s['samples'] / tot, 100.0 * s['instr'] / toti, s['wf'], s['wfi'],
        ' '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in top),
        ' '.join('%s=%.0fk' % (k, v / 1e3) for k, v in ops), s['why']))
