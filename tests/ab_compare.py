"""Compare per-op event logs of a same-box interleaved A/B (tests/run_ab.sh):  python tests/ab_compare.py gpurun_out/ops_TAG"""
import re, sys, glob

def load(f):
    d = {}
    for l in open(f):
        m = re.match(r'OP\s+(\d+) stage\s+(\d+) bn\s+(\d+)\s+([\d.]+) us', l)
        if m:
            d[int(m.group(1))] = float(m.group(4))
    return d

pre = sys.argv[1]
off = [load(f) for f in sorted(glob.glob(pre + '_off_*.log'))]
on = [load(f) for f in sorted(glob.glob(pre + '_on_*.log'))]
ta = tb = 0.0
for k in sorted(off[0]):
    a = min(o[k] for o in off); b = min(o[k] for o in on)
    ta += a; tb += b
    if abs(a - b) / a > 0.04:
        print(f'op {k:3d}  {a:8.1f} -> {b:8.1f} us  {(b - a) / a * 100:+.1f}%')
print(f'sum of per-op minima: {ta / 1e3:.3f} -> {tb / 1e3:.3f} ms')
