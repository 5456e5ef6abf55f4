"""Split a kernel's SASS (ncu --page source --csv) at BAR.SYNC instructions and print, per segment, the stall
samples, executed warp-instructions and HMMA count (development aid: which phase of a fused kernel costs what)."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[h]
iA, iS, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
seg = dict(s=0, e=0, hmma=0, n=0, st=collections.Counter())
segs = []
for r in rows[h + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    s, e = int(r[iS] or 0), int(r[iE])
    seg['s'] += s; seg['e'] += e; seg['n'] += 1
    if 'HMMA' in r[iA]: seg['hmma'] += e
    for i in stall_cols:
        if r[i].isdigit(): seg['st'][hdr[i]] += int(r[i])
    if 'BAR.SYNC' in r[iA] or 'EXIT' in r[iA]:
        segs.append(seg); seg = dict(s=0, e=0, hmma=0, n=0, st=collections.Counter())
segs.append(seg)
tot = sum(x['s'] for x in segs) or 1
tote = sum(x['e'] for x in segs) or 1
for i, x in enumerate(segs):
    if x['e'] == 0: continue
    top = ', '.join('%s %d' % (k.replace('stall_', ''), v) for k, v in x['st'].most_common(4))
    print('seg %2d: %5.1f%% samples  %5.1f%% inst (%8d, %5d sass)  hmma %7d  | %s' % (i, 100 * x['s'] / tot, 100 * x['e'] / tote, x['e'], x['n'], x['hmma'], top))
