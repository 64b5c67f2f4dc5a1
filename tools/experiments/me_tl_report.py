"""Summarise a DC_ME_TIMELINE dump (debug build -DDC_ME_TIMELINE; see music_encoder_tc.cuh ME_TL)."""
import collections
import sys

lines = open(sys.argv[1]).read().split("\n")
idx = [i for i, l in enumerate(lines) if l.startswith("[me_tl] kernel 0")][-1]
cur, d = None, {}
for l in lines[idx:]:
    if l.startswith("[me_tl] kernel"):
        cur = int(l.split()[-1])
        d[cur] = []
    elif l.startswith("  ") and cur is not None:
        a, b = l.split()
        d[cur].append((int(a), int(b)))
for k, ev in d.items():
    seg, prev = collections.defaultdict(list), None
    for s, t in ev:
        if prev is not None:
            seg[(prev[0], s)].append(t - prev[1])
        prev = (s, t)
    bands = [t for s, t in ev if s == 1]
    per = (bands[-1] - bands[1]) / (len(bands) - 2) if len(bands) > 2 else 0
    print("kernel", k, "band period", int(per), " ".join(f"{a}>{b}:{sum(v[1:] or v) // len(v[1:] or v)}x{len(v)}" for (a, b), v in sorted(seg.items())))
