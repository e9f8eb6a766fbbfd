"""Yardstick only (not used by the product): cuFFT Z2Z time via torch.fft for the Bluestein lengths."""
import torch

for log2l in (17, 21, 24):
    x = torch.randn(1 << log2l, dtype=torch.complex128, device="cuda:0")
    for _ in range(3):
        y = torch.fft.fft(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = torch.fft.fft(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"cuFFT Z2Z 2^{log2l}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
