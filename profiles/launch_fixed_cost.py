"""Fixed cost per K1 / K2 launch: time row bands of config 2 of several heights (stif_decode_rows, halo 8: stage A+B covers rows + 8) and fit t = a + b * rows."""
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, numpy as np, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
out = torch.empty((1, 1, 3, 1080, 1920), device="cuda")
rows_list = [64, 136, 272, 544, 808, 1080]
res = []
for rows in rows_list:
    kw = dict(rows=(0, rows), halo=8) if rows < 1080 else {}
    for _ in range(3): dec.decode_stacked(lat, fr, [0.5], (1080, 1920), out=out, **kw)
    torch.cuda.synchronize(); dec.profile(True); dec.profile_read()
    for _ in range(10): dec.decode_stacked(lat, fr, [0.5], (1080, 1920), out=out, **kw)
    p = dec.profile_read(); dec.profile(False)
    ms = [m / max(c, 1) for m, c in zip(p["ms"], p["count"])]
    res.append(ms)
    print(f"rows {rows:5d}: K0 {ms[0]*1e3:7.1f} us  K1 {ms[1]*1e3:7.1f} us  K2 {ms[2]*1e3:7.1f} us")
r = np.array(rows_list, float)
for k, name in ((1, "K1"), (2, "K2")):
    t = np.array([x[k] for x in res]) * 1e3
    b, a = np.polyfit(np.minimum(r + 8, 1080) if k == 1 else r, t, 1)
    print(f"{name}: t = {a:.1f} us + {b:.4f} us/row  (full raster {a + b * 1080:.1f} us; fixed share {a / (a + b * 1080) * 100:.1f} %)")
