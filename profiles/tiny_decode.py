import sys
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from stif_b200 import synthetic as synth
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(synth.make_weights(0, False))
lat, fr = synth.make_inputs(1, 1, 16, 16, 0.05)
out = dec.decode_stacked(torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda(), [0.5], (64, 64))
torch.cuda.synchronize(); print("ok", float(out.abs().mean()))
