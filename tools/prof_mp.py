"""Tiny driver for ncu: a few multi-pass (K2) launches at BASELINE config 5a (no timing claims)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import ramdsir_b200 as R
B, C, H, W = 64, 3, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(0, 256, (B, C, H, W), device="cuda", generator=g).float()
bank = R.build_amp_bank(torch.randint(0, 256, (8, C, H, W), device="cuda", generator=g).float())
rs = np.random.RandomState(0)
pidx = torch.from_numpy(rs.randint(0, 8, B)).cuda(); rk = torch.from_numpy(rs.randint(1, 11, B).astype(np.int32)).cuda()
out = torch.empty_like(src)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    R.ram_mix_batch(src, bank, pidx, rk, clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0, out=out)
torch.cuda.synchronize()
print("done")
