"""Issuer-warp view of the STIF_TRACE timeline: tag 40 = begin issuing a chunk's MMAs, 41 = committed."""
import sys
lines = [l for l in open(sys.argv[1])]
def parse(kern, warp):
    ls = [l for l in lines if l.startswith(f"{kern} warp {warp}:")]
    l = ls[-1].split(':', 1)[1].split()
    return [(int(x.split(':')[0]), int(x.split(':')[1])) for x in l]
for kern in ("K1", "K2"):
    ev0 = parse(kern, 0); st = [i for i, e in enumerate(ev0) if e[0] == 1]; t0 = ev0[st[3]][1]; t1 = ev0[st[4]][1]
    for w in (16, 17):
        ev = [(tag, t - t0) for tag, t in parse(kern, w) if t0 - 500 <= t <= t1 + 500]
        print(kern, "issuer", w, " ".join(f"{tag}:{t}" for tag, t in ev))
