"""Device-resident step time with per-kernel event profiling off vs on (events between launches defeat the
programmatic-dependent-launch overlap) -- config 2, 20 steps each."""
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
out = torch.empty((2, 1, 3, 1080, 1920), device="cuda")
def run(n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for _ in range(5): dec.decode_stacked(lat, fr, [0.0, 0.5], (1080, 1920), out=out)
torch.cuda.synchronize()
for rep in range(2):
    dec.profile(False); t_off = run()
    dec.profile(True); dec.profile_read(); t_on = run(); dec.profile_read()
    print(f"ms/step: profiling off {t_off:.4f}  on {t_on:.4f}")
