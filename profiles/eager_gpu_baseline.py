"""The reference's execution model on the same B200 ("the kernel to beat on the same box"): the
eager PyTorch op sequence of LunaTokis.decoding (oracle/port_torch.py, fp32, TF32 off) timed with CUDA events at config 2,
next to this library's device-resident decode of the same inputs.  Measurement tooling, not product code."""
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200"); sys.path.insert(0, ".")
import torch
import stif_b200
from stif_b200 import synthetic as synth
from oracle import port_torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
w = synth.make_weights(0, False)
lat, fr = synth.make_inputs(100, 1, 270, 480, 0.05)
times, size = [0.0, 0.5], (1080, 1920)
nq = len(times) * size[0] * size[1]


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


L, F = torch.from_numpy(lat).cuda(), torch.from_numpy(fr).cuda()
wd = {k: torch.from_numpy(v).cuda() for k, v in w.items()}
ms_eager, ref = timed(lambda: port_torch.decode(L, F, wd, times, size, device="cuda"), 3)
dec = stif_b200.STIFQueryDecoder(0, mode="bf16"); dec.load_weights(w)
ms_ours, out = timed(lambda: dec.decode_stacked(L, F, times, size), 20)
dec32 = stif_b200.STIFQueryDecoder(0, mode="fp32"); dec32.load_weights(w)
ms_ours32, out32 = timed(lambda: dec32.decode_stacked(L, F, times, size), 2)
print(f"eager PyTorch fp32 on the B200 : {ms_eager:9.2f} ms  {nq / ms_eager * 1e3:.3e} q/s   peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
print(f"this library, bf16 tensor-core  : {ms_ours:9.2f} ms  {nq / ms_ours * 1e3:.3e} q/s   ({ms_eager / ms_ours:.1f}x)   max-abs vs eager {float((out - ref).abs().max()):.2e}")
print(f"this library, fp32 kernels      : {ms_ours32:9.2f} ms  {nq / ms_ours32 * 1e3:.3e} q/s   ({ms_eager / ms_ours32:.1f}x)   max-abs vs eager {float((out32 - ref).abs().max()):.2e}")
