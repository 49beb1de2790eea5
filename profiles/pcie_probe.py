"""PCIe probe for the host-buffer entry point: raw pinned H2D / D2H rates at config-2 sizes and the e2e call time."""
import sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, numpy as np, stif_b200
from stif_b200 import synthetic as synth

def ev_ms(fn, n=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n

lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat_h, fr_h = torch.from_numpy(lat).pin_memory(), torch.from_numpy(fr).pin_memory()
lat_d = torch.empty_like(lat_h, device="cuda")
out_d = torch.empty((2, 1, 3, 1080, 1920), device="cuda"); out_h = torch.empty(out_d.shape).pin_memory()
ms = ev_ms(lambda: lat_d.copy_(lat_h, non_blocking=True)); print("H2D %.1f MB: %.3f ms = %.1f GB/s" % (lat_h.nbytes / 1e6, ms, lat_h.nbytes / ms / 1e6))
ms = ev_ms(lambda: out_h.copy_(out_d, non_blocking=True)); print("D2H %.1f MB: %.3f ms = %.1f GB/s" % (out_h.nbytes / 1e6, ms, out_h.nbytes / ms / 1e6))
s2 = torch.cuda.Stream()
def both():
    lat_d.copy_(lat_h, non_blocking=True)
    with torch.cuda.stream(s2): out_h.copy_(out_d, non_blocking=True)
t0 = time.perf_counter(); [both() for _ in range(10)]; torch.cuda.synchronize(); print("both directions, 10x: %.3f ms each" % ((time.perf_counter() - t0) * 100))
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
for _ in range(3): dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h); ts.append((time.perf_counter() - t0) * 1e3)
print("decode_host ms: min %.3f median %.3f" % (min(ts), sorted(ts)[5]))
