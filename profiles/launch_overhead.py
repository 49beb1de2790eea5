"""Fit per-launch fixed overhead of K1/K2: time launches whose tile counts are just under whole numbers of waves."""
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
for rows in (59, 118, 177, 355, 710, 1065, 1080):
    out = torch.empty((2, 1, 3, rows, 1920), device="cuda")
    for _ in range(3): dec.decode_stacked(lat, fr, [0.0, 0.5], (rows, 1920), out=out)
    torch.cuda.synchronize(); dec.profile(True); dec.profile_read()
    for _ in range(10): dec.decode_stacked(lat, fr, [0.0, 0.5], (rows, 1920), out=out)
    p = dec.profile_read(); dec.profile(False)
    k1, k2 = p["ms"][1] / p["count"][1], p["ms"][2] / p["count"][2]
    w1 = rows * 15 / 296; w2 = ((rows + 7) // 8) * 120 / 296
    print(f"rows {rows:5d}: K1 {k1*1e3:7.1f} us  waves {w1:6.2f}   K2 {k2*1e3:7.1f} us  waves {w2:6.2f}")
