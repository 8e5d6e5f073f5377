#!/usr/bin/env python
"""(dev) Kernel-only timing of the fused RAM launch: python tools/kbench.py [--workload fundus] [--nsrc 2] [--steps 200]
Prints ms/step and the HBM roofline fraction for serial and overlapped launches, float32 and uint8 sources."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ramdsir_b200 as R

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="256,3,256,256")
ap.add_argument("--nsrc", type=int, default=2)
ap.add_argument("--nout", type=int, default=3)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--nbank", type=int, default=768)
ap.add_argument("--epi", default="fundus")
ap.add_argument("--dual", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
B, C, H, W = map(int, a.shape.split(","))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
fundus = a.epi == "fundus"
mk = (lambda s: torch.randint(0, 256, s, device=dev, generator=g).float()) if fundus else (lambda s: torch.rand(s, device=dev, generator=g) * 2 - 1)
srcs = [mk((B, C, H, W)) for _ in range(a.nsrc)]
outs = [torch.empty_like(srcs[0]) for _ in range(a.nout)]
cleans = [torch.empty_like(srcs[0]) for _ in range(a.nout)] if a.dual else None
bank = R.build_amp_bank(mk((a.nbank, C, H, W)))
rs = np.random.RandomState(0)
pidx = torch.from_numpy(rs.randint(0, a.nbank, B)).to(dev)
rk = torch.from_numpy(rs.randint(1, 11, B).astype(np.int32)).to(dev)
epi = R.FUNDUS_EPILOGUE if fundus else R.PROSTATE_EPILOGUE
b = R.band_halfwidth(H, W, 0.1)
abytes = (2 * C * H * W * 4 + C * (2 * b + 1) ** 2 * 4) * B
peak = 6545.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def run(ss, **kw):
    def one(i):
        R.ram_mix_batch(ss[i % len(ss)], bank, pidx, rk, clip=epi[:2], scale=epi[2], shift=epi[3], out=outs[i % a.nout],
                        clean_out=None if cleans is None else cleans[i % a.nout], **kw)
    for i in range(10):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for i in range(a.steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    return best


res = {"tag": a.tag, "shape": a.shape, "nsrc": a.nsrc}
for name, kw in (("serial", dict(serial=True)), ("auto", {})):
    ms = run(srcs, **kw)
    res[name] = {"ms": round(ms, 5), "frac": round(abytes / (ms * 1e-3) / 1e9 / peak, 4)}
if fundus and R.u8_supported(C, H, W, b):
    s8 = [s.to(torch.uint8) for s in srcs]
    for name, kw in (("u8_serial", dict(serial=True)), ("u8_auto", {})):
        res[name] = {"ms": round(run(s8, **kw), 5)}
print(json.dumps(res), flush=True)
