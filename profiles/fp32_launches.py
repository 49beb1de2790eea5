import sys
sys.path.insert(0, "stif-continuous-video-representation_b200"); sys.path.insert(0, ".")
import torch, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="fp32"); dec.load_weights(synth.make_weights(0, True))
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
lat, fr = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
for _ in range(2): dec.decode_stacked(lat, fr, [0.5], (1080, 1920))
torch.cuda.synchronize()
