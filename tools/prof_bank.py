"""Tiny driver for ncu: a few bank-builder (K3) launches at BASELINE config 4's bank (1024 x 3 x 256 x 256; no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ramdsir_b200 as R
N, C, H, W = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1024,3,256,256").split(","))
g = torch.Generator(device="cuda").manual_seed(0)
imgs = torch.randint(0, 256, (N, C, H, W), device="cuda", generator=g).float()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    bank = R.build_amp_bank(imgs)
torch.cuda.synchronize()
if len(sys.argv) > 3:     # "time": CUDA-event timing of n more builds (kernel + launch; the images are far larger than L2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for _ in range(n):
            bank = R.build_amp_bank(imgs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    print("bank %s %s: %.5f ms, %.1f GB/s read" % (sys.argv[3], sys.argv[2], best, imgs.numel() * 4 / best / 1e6))
print("done", tuple(bank.shape))
