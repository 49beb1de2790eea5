import sys, time
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat_h, fr_h = torch.from_numpy(lat).pin_memory(), torch.from_numpy(fr).pin_memory()
out_h = torch.empty((2, 1, 3, 1080, 1920)).pin_memory()
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
for _ in range(4): dec.decode_host(lat_h, fr_h, [0.0, 0.5], (1080, 1920), out=out_h)
