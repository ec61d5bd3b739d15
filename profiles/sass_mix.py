"""Summarise an `ncu --page source --csv --print-source sass` export: warp-instruction mix and stall reasons per element."""
import csv, collections, re, sys
path, nelem = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); stall = collections.Counter()
stallcols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
samples = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == 'Address':
        continue
    try:
        n = int(r[ix['Instructions Executed']] or 0)
    except ValueError:
        continue
    src = r[ix['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else src[:10]
    op = op.split('.')[0]
    byop[op] += n; tot += n
    samples += int(r[ix['# Samples']] or 0)
    for c in stallcols:
        stall[c] += int(r[ix[c]] or 0)
print('total warp instr', tot, 'per element', tot / nelem)
for op, n in byop.most_common(28):
    print(f'{op:12s} {n/nelem:8.1f} {100*n/tot:5.1f}%')
print('samples', samples)
for c, n in stall.most_common(10):
    print(c, n, f'{100*n/max(samples,1):.1f}%')
