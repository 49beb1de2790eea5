"""Where the band-major host pipeline's compute time goes: per-kernel event sums of stif_decode_host (config 2) next to the device-buffer path."""
import sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat_h, fr_h = torch.from_numpy(lat).pin_memory(), torch.from_numpy(fr).pin_memory()
out_h = torch.empty((2, 1, 3, 1080, 1920)).pin_memory()
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
for _ in range(3): dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h)
dec.profile(True); dec.profile_read()
n = 10
t0 = time.perf_counter()
for _ in range(n): dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h)
wall = (time.perf_counter() - t0) / n * 1e3
p = dec.profile_read()
print("host path : K0/K1/K2 ms per call", [round(m / n, 4) for m in p["ms"]], "launches per call", [c // n for c in p["count"]],
      "sum", round(sum(p["ms"]) / n, 4), "wall (profiling on)", round(wall, 3))
latd, frd = lat_h.cuda(), fr_h.cuda()
out = torch.empty((2, 1, 3, 1080, 1920), device="cuda")
for _ in range(3): dec.decode_stacked(latd, frd, [0.0, 0.5], (1080, 1920), out=out)
torch.cuda.synchronize(); dec.profile_read()
for _ in range(n): dec.decode_stacked(latd, frd, [0.0, 0.5], (1080, 1920), out=out)
p = dec.profile_read()
print("device path: K0/K1/K2 ms per call", [round(m / n, 4) for m in p["ms"]], "launches per call", [c // n for c in p["count"]],
      "sum", round(sum(p["ms"]) / n, 4))
