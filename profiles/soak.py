"""Soak: many decodes over assorted shapes / timestep counts (device and host entry points), to shake out rare hangs or
nondeterminism in the persistent kernels' barrier protocols.  Run under `timeout`."""
import sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200")
import numpy as np, torch, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16")
shapes = [((540, 960), (2160, 3840), 8, 6), ((270, 480), (1080, 1920), 2, 30), ((270, 480), (1755, 3120), 3, 8),
          ((64, 64), (416, 416), 8, 30), ((37, 53), (301, 97), 5, 30), ((96, 40), (384, 163), 6, 30)]
t_all = time.time()
for stress in (False, True):
    dec.load_weights(synth.make_weights(5, stress))
    for (H, W), (HH, WW), T, reps in shapes:
        lat, fr = synth.make_inputs(H + W, 1, H, W, 1.0 if stress else 0.05)
        L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
        times = [i / T for i in range(T)]
        ref = None
        t0 = time.time()
        for r in range(reps):
            out = dec.decode_stacked(L, F, times, (HH, WW))
            torch.cuda.synchronize()
            if ref is None: ref = out.clone()
            else: assert torch.equal(out, ref), ("device path not deterministic", H, W, r)
            del out
        host = dec.decode_host(lat, fr, times, (HH, WW))
        assert torch.equal(host, ref.cpu()), ("host path differs", H, W)
        print(f"stress={stress} {H}x{W}->{HH}x{WW} T={T}: {reps} decodes ok ({time.time() - t0:.1f}s), respins so far {dec.host_pipeline()}", flush=True)
        del ref, host, L, F
print(f"soak ok in {time.time() - t_all:.0f}s")
