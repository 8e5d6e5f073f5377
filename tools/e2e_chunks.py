import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import ramdsir_b200 as R
from ramdsir_b200.host import HostPipeline
B, C, H, W = 256, 3, 256, 256
g = torch.Generator(device="cuda").manual_seed(0)
src = torch.randint(0, 256, (B, C, H, W), device="cuda", generator=g).float()
bank = R.build_amp_bank(src[:64].contiguous())
src_h = src.cpu().pin_memory(); out_h = torch.empty_like(src_h).pin_memory()
pidx = np.arange(B) % 64; rk = np.full(B, 3, np.int32)
for chunk in (8, 16, 32, 64, 128, 256):
    with HostPipeline(chunk_images=chunk, C=C, H=H, W=W) as pipe:
        for _ in range(3): pipe.mix(src_h, bank, pidx, rk, out_h, epilogue=R.FUNDUS_EPILOGUE)
        t0 = time.perf_counter()
        for _ in range(20): pipe.mix(src_h, bank, pidx, rk, out_h, epilogue=R.FUNDUS_EPILOGUE)
        dt = (time.perf_counter() - t0) / 20
    print("chunk %3d: %.3f ms/step  %.0f img/s  (%.1f GB/s each way)" % (chunk, dt * 1e3, B / dt, src_h.numel() * 4 / dt / 1e9))
# plain one-shot copies for reference
d = torch.empty_like(src)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(src_h, non_blocking=True); torch.cuda.synchronize()
print("H2D alone %.1f GB/s" % (src_h.numel() * 4 * 10 / (time.perf_counter() - t0) / 1e9))
t0 = time.perf_counter()
for _ in range(10): out_h.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("D2H alone %.1f GB/s" % (src_h.numel() * 4 * 10 / (time.perf_counter() - t0) / 1e9))
