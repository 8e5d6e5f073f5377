"""(dev) A few small launches of every kernel family, for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanity_run.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ramdsir_b200 as R

g = torch.Generator(device="cuda").manual_seed(0)


def run(shape, epi, paths=(None,), u8=False, dual=False):
    B, C, H, W = shape
    src = torch.randint(0, 256, shape, device="cuda", generator=g).float()
    par = torch.randint(0, 256, (4, C, H, W), device="cuda", generator=g).float()
    for path in paths:
        bank = R.build_amp_bank(par, path=path)
        outs = [torch.empty_like(src) for _ in range(3)]
        clean = torch.empty_like(src) if dual else None
        for i in range(5):
            s = src.to(torch.uint8) if (u8 and i % 2) else src
            R.ram_mix_batch(s, bank, np.arange(B) % 4, np.arange(B) % 9 + 1, clip=epi[:2], scale=epi[2], shift=epi[3],
                            out=outs[i % 3], clean_out=clean, path=path)
        torch.cuda.synchronize()
        print(shape, path, "u8" if u8 else "", "dual" if dual else "", float(outs[0].abs().mean()))


run((3, 3, 256, 256), R.FUNDUS_EPILOGUE, paths=(None, "runtime", "generic"))
run((3, 3, 256, 256), R.FUNDUS_EPILOGUE, u8=True, dual=True)
run((2, 1, 384, 384), R.PROSTATE_EPILOGUE, u8=False, dual=True)
run((2, 2, 64, 128), R.FUNDUS_EPILOGUE, paths=(None, "generic"))
run((1, 1, 31, 45), R.FUNDUS_EPILOGUE)
run((1, 1, 512, 512), R.FUNDUS_EPILOGUE)
run((1, 1, 1024, 1024), R.FUNDUS_EPILOGUE)
