"""Tiny driver for ncu: a few fused mix launches at the bench configuration (no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import ramdsir_b200 as R
cfg = sys.argv[1] if len(sys.argv) > 1 else "fundus"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
u8 = len(sys.argv) > 3 and sys.argv[3] == "u8"       # uint8-resident sources (ramdsir_mix_dual_u8)
B, C, H, W = {"fundus": (256, 3, 256, 256), "prostate": (512, 1, 384, 384), "prostate3": (128, 3, 384, 384)}[cfg]
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(0, 256, (B, C, H, W), device="cuda", generator=g).float()
bank = R.build_amp_bank(torch.randint(0, 256, (64, C, H, W), device="cuda", generator=g).float())
rs = np.random.RandomState(0)
pidx = torch.from_numpy(rs.randint(0, 64, B)).cuda(); rk = torch.from_numpy(rs.randint(1, 11, B).astype(np.int32)).cuda()
out = torch.empty_like(src)
if u8:
    src = src.to(torch.uint8)
for _ in range(n):
    R.ram_mix_batch(src, bank, pidx, rk, clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0, out=out)
torch.cuda.synchronize()
print("done", cfg, n)
