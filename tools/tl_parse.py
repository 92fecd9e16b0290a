"""Dev tool: summarise the backward timeline captured by tools/prof_timeline.py (per-role (tag, cycle) lists of CTA 0)."""
import sys
txt = open(sys.argv[1]).read()
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bwd = txt[txt.index('BWD'):].split('\n')
roles = {}
i = 1
while i < len(bwd) - 1:
    w = bwd[i].split()
    if w and w[0] in ('TMA', 'MMA', 'EPI', 'PROD'):
        roles[w[0]] = [tuple(map(int, x.split(':'))) for x in bwd[i + 1].split()]
        i += 2
    else:
        i += 1
mma = roles['MMA']
starts = [k for k, (t, c) in enumerate(mma) if t == 1]
a, b = starts[tile], starts[tile + 1]
ev = mma[a:b + 1]
t0 = ev[0][1]
print("MMA tile", tile, "period", ev[-1][1] - t0)
prev = t0
for t, c in ev:
    print(f"  tag {t:4d} @ {c - t0:6d}  (+{c - prev})")
    prev = c
for r in ('TMA', 'EPI', 'PROD'):
    ev = [(t, c - t0) for t, c in roles[r] if t0 - 3000 <= c <= t0 + 40000]
    print(r, ' '.join(f"{t}:{c}" for t, c in ev))
