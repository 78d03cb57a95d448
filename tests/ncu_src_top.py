"""Top stall lines of an `ncu --page source --csv` export:  python tests/ncu_src_top.py file.csv [N]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1][:120])
hdr = rows[1]
iS = hdr.index('Warp Stall Sampling (All Samples)'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
sec = []
for r in rows[2:]:
    if len(r) < 6 or r[0] in ("Kernel Name", "Address"): break
    sec.append(r)
tot = sum(int(r[iS] or 0) for r in sec)
print('total samples', tot, 'lines', len(sec))
top = sorted(enumerate(sec), key=lambda x: -int(x[1][iS] or 0))[:N]
for i, r in sorted(top):
    print(i, r[iS], r[iex], r[isrc][:110])
b = collections.Counter(); bi = collections.Counter()
for r in sec:
    b[r[iex]] += int(r[iS] or 0); bi[r[iex]] += 1
print('samples by execution count (role fingerprint):')
for k, v in sorted(b.items(), key=lambda x: -x[1])[:10]:
    print(' exec', k, 'samples', v, 'lines', bi[k])
