"""Two-lane view of the STIF_TRACE clock64 timeline: one epilogue warp of WG0 (slot 0) and one of WG1 (slot 8) of CTA 0 side
by side on a common clock, a few tiles into the launch -- shows how the two workgroups' phases overlap.
    python profiles/trace_lanes.py gpurun_out/trace.txt [K1|K2] [first_tile] [n_tiles]"""
import sys
lines = [l for l in open(sys.argv[1])]
kern = sys.argv[2] if len(sys.argv) > 2 else "K1"
first = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ntile = int(sys.argv[4]) if len(sys.argv) > 4 else 2
def parse(warp):
    ls = [l for l in lines if l.startswith(f"{kern} warp {warp}:")]
    l = ls[-1].split(':', 1)[1].split()
    return [(int(x.split(':')[0]), int(x.split(':')[1])) for x in l]
NAMES = {1: "tile", 2: "L0done", 3: "bar", 4: "gatherB", 50: "ACQ-F2", 51: "REL-F2", 52: "ACQ-L2", 53: "REL-L2"}
def name(tag):
    if tag in NAMES: return NAMES[tag]
    if 10 <= tag < 20: return f"rdy{tag - 10}"
    if 20 <= tag < 30: return f"epi{tag - 20}"
    return str(tag)
e0, e1 = parse(0), parse(8)
s0 = [i for i, e in enumerate(e0) if e[0] == 1]
t0 = e0[s0[first]][1]; t1 = e0[s0[first + ntile]][1]
ev = sorted([(t - t0, 0, tag) for tag, t in e0 if t0 <= t <= t1] + [(t - t0, 1, tag) for tag, t in e1 if t0 - 2000 <= t <= t1])
last = [None, None]
for t, wg, tag in ev:
    d = "" if last[wg] is None else f"+{t - last[wg]}"
    last[wg] = t
    print(f"{t:7d} " + ("" if wg == 0 else " " * 28) + f"{name(tag):8s} {d}")
