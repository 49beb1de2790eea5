"""stif_dcn_v2_forward vs torchvision.ops.deform_conv2d at the encoder's three pyramid levels (64 -> 64, 3x3, dg = 8)."""
import sys
sys.path.insert(0, "stif-continuous-video-representation_b200")
import torch, stif_b200
from torchvision.ops import deform_conv2d
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def ev(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
for H, W in ((272, 480), (136, 240), (68, 120)):
    x = torch.randn(1, 64, H, W, device="cuda"); w = torch.randn(64, 64, 3, 3, device="cuda") / 24; b = torch.randn(64, device="cuda")
    off = 2 * torch.randn(1, 144, H, W, device="cuda"); m = torch.sigmoid(torch.randn(1, 72, H, W, device="cuda"))
    t_tv = ev(lambda: deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m))
    t_us = ev(lambda: stif_b200.dcn_v2_forward(x, w, b, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 8))
    err = float((stif_b200.dcn_v2_forward(x, w, b, off, m, 3, 3, 1, 1, 1, 1, 1, 1, 8) - deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m)).abs().max())
    print(f"{H}x{W}: torchvision {t_tv * 1e3:.1f} us, stif_dcn_v2_forward {t_us * 1e3:.1f} us ({t_tv / t_us:.1f}x), max-abs diff {err:.2e}")
