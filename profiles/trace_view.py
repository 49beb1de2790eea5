"""Render the STIF_TRACE clock64 timeline (tags: 1 tile start, 2 gather done, 3 barrier, 4 stage-B gather done,
10+i chunk i ready, 20+i epilogue i done, 30+i barrier i passed)."""
import sys
lines = [l for l in open(sys.argv[1])]
def parse(kern, warp):
    ls = [l for l in lines if l.startswith(f"{kern} warp {warp}:")]
    l = ls[-1].split(':', 1)[1].split()
    return [(int(x.split(':')[0]), int(x.split(':')[1])) for x in l]
for kern in ("K1", "K2"):
    ev = parse(kern, 0); t0 = ev[0][1]
    starts = [i for i, e in enumerate(ev) if e[0] == 1]
    print(kern, "tile starts (clk):", [ev[i][1] - t0 for i in starts][:8])
    a, b = starts[3], starts[4]
    prev = ev[a][1]; out = []
    for tag, t in ev[a:b + 1]:
        out.append(f"{tag}:+{t - prev}"); prev = t
    print(" ".join(out))
    for w in (0, 4, 8, 12):
        e = parse(kern, w); st = [x for x in e if x[0] == 1]
        print("  warp", w, "tile-start clocks rel:", [x[1] - t0 for x in st[:6]])
