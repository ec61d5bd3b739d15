"""Per-source-line samples / instruction counts from `ncu --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
path, nelem = sys.argv[1], float(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = sys.argv[4] if len(sys.argv) > 4 else 'samples'
rows = list(csv.reader(open(path)))
cur = None; agg = collections.Counter(); inst = collections.Counter(); txt = {}
hdr = None
def toi(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        ln = (cur, int(r[0])); txt[ln] = r[1]
        agg[ln] += toi(r[hdr.index('# Samples')]); inst[ln] += toi(r[hdr.index('Instructions Executed')])
tot = sum(agg.values()); ti = sum(inst.values())
print('samples', tot, 'instr/elem', ti / nelem)
src = agg if key == 'samples' else inst
for ln, n in src.most_common(topn):
    print(f'{ln[0][:14]:14s} {ln[1]:4d} samp {100*agg[ln]/max(tot,1):5.1f}% inst {inst[ln]/nelem:7.1f}  {txt[ln].strip()[:100]}')
