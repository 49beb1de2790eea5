"""Three-lane view of the rotation kernel's clock64 timeline (one epilogue warp of each workgroup of CTA 0):
    python profiles/trace_lanes3.py gpurun_out/trace_RT.txt [t_begin] [t_len]"""
import sys
lines = [l for l in open(sys.argv[1])]
t_begin = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
t_len = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
def parse(warp, kern="K2"):
    ls = [l for l in lines if l.startswith(f"{kern} warp {warp}:")]
    l = ls[-1].split(':', 1)[1].split()
    return [(int(x.split(':')[0]), int(x.split(':')[1])) for x in l]
NAMES = {1: "tile", 2: "G-done", 3: "slot!"}
def name(tag):
    if tag in NAMES: return NAMES[tag]
    if 10 <= tag < 20: return f"rdy{tag - 10}"
    if 20 <= tag < 30: return f"epi{tag - 20}"
    return str(tag)
lanes = [parse(0), parse(8), parse(16)]
t0 = min(e[0][1] for e in lanes)
ev = sorted((t - t0, i, tag) for i, e in enumerate(lanes) for tag, t in e if t_begin <= t - t0 <= t_begin + t_len)
last = [None] * 3
for t, wg, tag in ev:
    d = "" if last[wg] is None else f"+{t - last[wg]}"
    last[wg] = t
    print(f"{t:7d} " + " " * (26 * wg) + f"{name(tag):7s}{d}")
