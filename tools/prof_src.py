#!/usr/bin/env python
"""Stall samples of one kernel by phase (between barriers) from `ncu --page source --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[1]]
edges = [0] + [b + 1 for b in bars] + [len(data)]
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("kernel:", rows[0][1][:90], " samples:", tot)
for a, b in zip(edges[:-1], edges[1:]):
    s = sum(int(r[ix['# Samples']]) for r in data[a:b])
    st = {k[6:]: sum(int(r[ix[k]] or 0) for r in data[a:b]) for k in keys}
    st = {k: v for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v > 0.03 * max(s, 1)}
    wf = sum(int(r[ix['L1 Wavefronts Shared']] or 0) for r in data[a:b])
    tg = sum(int(r[ix['L1 Tag Requests Global']] or 0) for r in data[a:b])
    ins = sum(int(r[ix['Instructions Executed']] or 0) for r in data[a:b])
    print(f"[{a:4d},{b:4d}) {100*s/tot:5.1f}%  inst={ins/1e6:7.1f}M smem_wf={wf/1e6:6.1f}M tagreq={tg/1e6:6.1f}M  {st}")
if len(sys.argv) > 2:
    thr = float(sys.argv[2])
    for i, r in enumerate(data):
        s = int(r[ix['# Samples']])
        if s >= thr * tot:
            print(f"{i:4d} {100*s/tot:5.1f}% {r[1].strip()[:80]}")
