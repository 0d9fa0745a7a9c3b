"""Hottest SASS lines (warp-stall samples) of the first kernel in an exported ncu source page.
    python tools/ncu_hot.py gpurun_out/q_conv_fwd_d.source.csv.gz [top] [kernel_index]"""
import csv, gzip, sys
rows = list(csv.reader(gzip.open(sys.argv[1], 'rt')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
a, b = starts[which], starts[which + 1]
print(rows[a][1][:150])
hdr = rows[a + 1]; ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[a + 2:b] if len(r) > 5]
S = lambda r: int(r[ix['# Samples']] or 0)
tot = sum(S(r) for r in body)
print('total samples', tot)
hot = sorted(range(len(body)), key=lambda k: -S(body[k]))[:top]
for k in sorted(hot):
    r = body[k]
    print(f"{k:5d} {S(r):6d} {100.0 * S(r) / tot:5.1f}% {r[ix['Instructions Executed']]:>8} {r[1][:100]}")
