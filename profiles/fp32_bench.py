"""STIF_MODE_FP32 (high-precision mode) throughput at config 2 and error vs the SIMT anchor (STIF_FP32_SIMT=1)."""
import os, sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200"); sys.path.insert(0, ".")
import torch, numpy as np, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="fp32"); dec.load_weights(synth.make_weights(0, True))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
out = torch.empty((2, 1, 3, 1080, 1920), device="cuda")
for _ in range(2): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print(f"fp32 mode ({'SIMT' if os.environ.get('STIF_FP32_SIMT') else 'split-bf16 tcgen05'}): {dt * 1e3:.2f} ms per config-2 step = {4147200 / dt:.3e} q/s, checksum {float(out.double().abs().mean()):.10f}")
