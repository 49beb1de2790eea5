"""Quick per-kernel timing of the bf16 path at config-2 size (used while tuning; prints ms per launch)."""
import sys, os
sys.path.insert(0, "stif-continuous-video-representation_b200"); sys.path.insert(0, ".")
import torch, numpy as np, stif_b200
from oracle import synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, len(sys.argv) > 1 and sys.argv[1] == "stress"))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
out = torch.empty((2, 1, 3, 1080, 1920), device="cuda")
for _ in range(3): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
torch.cuda.synchronize(); dec.profile(True); dec.profile_read()
for _ in range(10): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
p = dec.profile_read()
print("ms/launch K0,K1,K2:", [round(m / max(c, 1), 4) for m, c in zip(p["ms"], p["count"])],
      "step ms", round(sum(p["ms"]) / 10, 4), "checksum", float(out.double().abs().mean()))
