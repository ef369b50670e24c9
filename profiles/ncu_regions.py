"""Stall-sample breakdown of one kernel of an .ncu-rep by SASS region (runs of instructions with equal execution count).
usage: python profiles/ncu_regions.py <file.ncu-rep> <launch_index> [min_share]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ie, src, smp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
stalls = [(k, n) for k, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
# ncu lists the instructions of a function once per inlined copy of the source view: keep the first copy
seen, uniq = set(), []
for r in data:
    if r[0] in seen:
        continue
    seen.add(r[0]); uniq.append(r)
data = uniq
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[smp]) for r in data)
print('kernel', rows[0][1][:80]); print('warp instructions', tot, 'samples', tots)
groups = []
for k, r in enumerate(data):
    c = int(r[ie])
    if groups and groups[-1]['c'] == c:
        g = groups[-1]
    else:
        g = {'c': c, 'n': 0, 'k': k, 's': 0, 'st': {}}
        groups.append(g)
    g['n'] += 1; g['s'] += int(r[smp])
    for (ci, name) in stalls:
        v = int(r[ci]) if r[ci].isdigit() else 0
        if v:
            g['st'][name] = g['st'].get(name, 0) + v
for g in groups:
    if g['c'] * g['n'] > tot * thresh or g['s'] > tots * thresh:
        top = sorted(g['st'].items(), key=lambda kv: -kv[1])[:4]
        print('idx %5d n=%3d exec=%9d instr=%5.1f%% samples=%5.1f%%  %-34s %s' % (
            g['k'], g['n'], g['c'], 100.0 * g['c'] * g['n'] / tot, 100.0 * g['s'] / tots, data[g['k']][src].strip()[:34],
            ' '.join('%s=%.1f%%' % (n[6:], 100.0 * v / tots) for n, v in top)))
