"""Group an .ncu-rep's per-source-line warp instructions into named line ranges: python tools/ncu_groups.py rep file:lo-hi=name ..."""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
groups = []
for a in sys.argv[2:]:
    spec, name = a.split('=')
    f, r = spec.split(':')
    lo, hi = map(int, r.split('-'))
    groups.append((f, lo, hi, name))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 10 and r[0]:
        try:
            ln = int(r[0]); ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); sm = hdr.index('# Samples')
            wi, ti, ss = int(r[ie]), int(r[te]), int(r[sm])
        except Exception:
            continue
        name = f'other:{cur}'
        for f, lo, hi, n in groups:
            if f == cur and lo <= ln <= hi: name = n; break
        agg[name][0] += wi; agg[name][1] += ti; agg[name][2] += ss
tw = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:28s} warp-inst {v[0]/tw*100:6.2f}%  lanes {v[1]/max(v[0],1):5.1f}  thread-inst {v[1]:.3e}  stall-samples {v[2]/max(ts,1)*100:6.2f}%')
print(f'total warp-inst {tw:.4e} thread-inst {sum(v[1] for v in agg.values()):.4e}')
