"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = collections.Counter(); inst = collections.Counter()
tot = toti = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name'): continue
    if r[0] != '' and r[2] == '-':
        try: s = int(r[4]); ie = int(r[7])
        except ValueError: continue
        key = (cur_file, int(r[0]), r[1][:80])
        agg[key] += s; inst[key] += ie; tot += s; toti += ie
print('total samples', tot, 'total warp-instructions executed', toti)
for k, s in agg.most_common(top):
    print(f'{100*s/tot:5.1f}% smp  {100*inst[k]/toti:5.1f}% inst  {k[0]}:{k[1]}  {k[2]}')
