"""Opcode mix of one BAR.SYNC-delimited segment of a kernel (see ncu_phases.py): ncu_segmix.py file.csv SEG"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2])
h = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[h]
iA, iS, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
seg = 0
ops, smp = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    if seg == want:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_\.]+)', r[iA])
        op = m.group(2) if m else '?'
        op = '.'.join(op.split('.')[:2])
        ops[op] += int(r[iE]); smp[op] += int(r[iS] or 0)
    if 'BAR.SYNC' in r[iA] or 'EXIT' in r[iA]:
        seg += 1
tot = sum(ops.values()) or 1
for op, c in ops.most_common(28):
    print('%-22s %9d %5.1f%%  samples %5d' % (op, c, 100 * c / tot, smp[op]))
