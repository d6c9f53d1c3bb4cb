"""Summarise an `ncu --page source --csv` dump: executed-instruction and stall-sample share per
opcode, and the hottest instruction ranges.  usage: python scripts/ncu_hot.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
S, E, SRC = ci['# Samples'], ci['Instructions Executed'], ci['Source']
data = []
for r in rows[2:]:
    try:
        data.append((float(r[S]), float(r[E]), r[SRC].strip(), r))
    except (ValueError, IndexError):
        pass
ts, te = sum(d[0] for d in data), sum(d[1] for d in data)
print('instructions', len(data), 'samples', ts, 'warp-inst executed', te)
bs, be = collections.Counter(), collections.Counter()
for s, e, t, _ in data:
    parts = t.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    bs[op] += s; be[op] += e
print('%-10s %9s %9s' % ('opcode', 'exec%', 'samples%'))
for op, v in be.most_common(25):
    print('%-10s %8.2f%% %8.2f%%' % (op, 100 * v / te, 100 * bs[op] / ts))
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {h: sum(float(d[3][ci[h]] or 0) for d in data) for h in stall}
print({k: int(v) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]})
# hottest contiguous regions (by samples), window of 24 instructions
best = sorted(range(0, len(data), 8), key=lambda i: -sum(d[0] for d in data[i:i + 24]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 3]
for i in best:
    w = data[i:i + 24]
    print('--- region @%d: %.1f%% of samples, exec/inst %.0f' % (i, 100 * sum(d[0] for d in w) / ts, w[0][1]))
    for d in w:
        print('   %6d %9d  %s' % (d[0], d[1], d[2][:70]))
