"""Opcode mix + stall samples of one kernel from `ncu --page source --csv` output (development aid)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[h]
iA, iS, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ops, samp = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_\.]+)', r[iA])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(r[iE]); samp[op] += int(r[iS] or 0)
tot, tots = sum(ops.values()), sum(samp.values())
print('total warp-inst', tot, 'samples', tots)
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print('%-10s %9d %5.1f%%   samples %6d %5.1f%%' % (op, c, 100 * c / tot, samp[op], 100 * samp[op] / tots))
