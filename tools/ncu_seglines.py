"""Top SASS lines (by stall samples) of one BAR.SYNC-delimited segment: ncu_seglines.py file.csv SEG [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
h = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[h]
iA, iS, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
seg, out, idx = 0, [], 0
for r in rows[h + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    if seg == want:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i])), reverse=True)[:2]
        out.append((int(r[iS] or 0), idx, r[iA].strip()[:70], st))
    idx += 1
    if 'BAR.SYNC' in r[iA] or 'EXIT' in r[iA]:
        seg += 1
for s, i, a, st in sorted(out, reverse=True)[:topn]:
    print('%5d  #%-6d %-70s %s' % (s, i, a, st))
