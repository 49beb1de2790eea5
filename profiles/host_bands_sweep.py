"""e2e time of stif_decode_host (config 2, pinned buffers) vs the number of LR row bands of the band-major pipeline."""
import sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat_h, fr_h = torch.from_numpy(lat).pin_memory(), torch.from_numpy(fr).pin_memory()
out_h = torch.empty((2, 1, 3, 1080, 1920)).pin_memory()
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
for bands in ((0,) if len(sys.argv) > 1 and sys.argv[1] == "auto" else (0, 4, 5, 6, 7, 8, 9, 10)):   # 0 = the library's own choice (cost model)
    if bands: dec.host_pipeline(bands=bands)
    for _ in range(3): dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h)
    ts = []
    for _ in range(15):
        t0 = time.perf_counter(); dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"bands {bands:2d}: min {min(ts):.3f} median {sorted(ts)[7]:.3f} ms")
