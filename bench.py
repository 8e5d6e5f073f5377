#!/usr/bin/env python
"""bench.py -- RAM images/sec on B200 (BASELINE.json metric) with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one RAM pass over one batch of synthetic fundus-shaped images (BASELINE.json configs[1]:
256 x 3 x 256 x 256 fp32, L = 0.1 -> b = 25, fundus clip/normalise epilogue) per GPU.  Weak scaling:
every rank mixes its own 256-image shard against a bank that was built sharded and all-gathered once
before the timed region.  One JSON line is printed by rank 0.

value    = images/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
e2e      = images/s through the host-buffer C-ABI entry (ramdsir_mix_host_f32): pinned host src ->
           H2D -> kernel -> D2H -> pinned host out, every step, wall-clocked around the blocking call
roofline = algorithmic bytes (2*C*H*W*4 + C*(2b+1)^2*4 per image) / measured kernel time
           against MEASURED_PEAKS.json's HBM copy bandwidth
cpu_baseline = the numpy oracle (a port of the reference's numpy path) on all host cores, bounded sample
--impl reference = the same metric for the reference's CPU implementation (oracle port; the
           reference itself is Python under /root/reference and does not exist on the GPU box)
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1337                       # reference default (train.py:60)
L = 0.1
# BASELINE.json configs[1] is the headline (default); the others are selectable for development runs.
WORKLOADS = {
    "fundus": dict(C=3, H=256, W=256, BATCH=256, N_BANK=768, N_DOMAINS=3, EPI="fundus", LO=0, HI=256,
                   NAME="fundus RAM batch 256x3x256x256 fp32, L=0.1 (b=25), 3 source domains, clip[0,255]/127.5-1"),
    "fundus1024": dict(C=3, H=256, W=256, BATCH=1024, N_BANK=768, N_DOMAINS=3, EPI="fundus", LO=0, HI=256,
                       NAME="fundus RAM, per-rank shard of BASELINE configs[3] (8192 images / 8 GPUs): 1024x3x256x256 fp32, L=0.1"),
    "prostate": dict(C=1, H=384, W=384, BATCH=512, N_BANK=640, N_DOMAINS=5, EPI="prostate", LO=-1, HI=1,
                     NAME="prostate RAM batch 512x1x384x384 fp32, L=0.1 (b=38), 5 source domains, clip[-1,1]"),
    "prostate3": dict(C=3, H=384, W=384, BATCH=128, N_BANK=320, N_DOMAINS=5, EPI="prostate", LO=-1, HI=1,
                      NAME="prostate RAM batch 128x3x384x384 fp32 (real 3-slice layout), L=0.1 (b=38), clip[-1,1]"),
    "hires": dict(C=3, H=1024, W=1024, BATCH=64, N_BANK=48, N_DOMAINS=3, EPI="fundus", LO=0, HI=256,
                  NAME="high-res fundus RAM batch 64x3x1024x1024 fp32, L=0.1 (b=102), clip[0,255]/127.5-1"),
}
C, H, W = 3, 256, 256
BATCH = 256                       # per GPU (configs[1])
N_BANK = 768                      # source images in the amplitude bank (3 domains x 256; real fundus: 789)
N_DOMAINS = 3
NSRC = 2                          # distinct source batches the timed steps rotate over
EPI_NAME, LO, HI = "fundus", 0, 256
WORKLOAD = WORKLOADS["fundus"]["NAME"]


def select_workload(name):
    global C, H, W, BATCH, N_BANK, N_DOMAINS, WORKLOAD, EPI_NAME, LO, HI
    w = WORKLOADS[name]
    C, H, W, BATCH, N_BANK, N_DOMAINS = w["C"], w["H"], w["W"], w["BATCH"], w["N_BANK"], w["N_DOMAINS"]
    WORKLOAD, EPI_NAME, LO, HI = w["NAME"], w["EPI"], w["LO"], w["HI"]


def synth(torch, shape, dev, gen):
    """Synthetic images: uint8-valued floats for fundus (fundus.py:212), uniform [-1, 1] for prostate."""
    if EPI_NAME == "fundus":
        return torch.randint(0, 256, shape, device=dev, generator=gen).float()
    return torch.rand(shape, device=dev, generator=gen) * 2.0 - 1.0


def algorithmic_bytes_per_image(c, h, w, b):
    return 2 * c * h * w * 4 + c * (2 * b + 1) ** 2 * 4


def nominal_fft_flops_per_image(c, h, w):
    return c * 2 * 5 * h * w * np.log2(h * w)


def draw_selection(batch, n_bank, seed):
    """Partner rows and ratio numerators with the reference's RNG calls (fundus.py:205,208,35)."""
    from ramdsir_b200 import rng
    rng.seed_worker_streams(seed)
    per = n_bank // N_DOMAINS
    names = ["Domain%d" % (i + 1) for i in range(N_DOMAINS)]
    index = rng.BankIndex([(n, per) for n in names])
    sizes = {n: per for n in names}
    pidx = np.empty(batch, np.int64)
    rk = np.empty(batch, np.int32)
    for i in range(batch):
        dom, j = rng.draw_partner(names, sizes)
        pidx[i] = index.row(dom, j)
        rk[i] = rng.draw_ratio_k()
    return pidx, rk


# ---------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's numpy path)
_CPU = {}


def _cpu_init(seed):
    from oracle import ram_oracle as O
    rs = np.random.RandomState(seed)
    _CPU["O"] = O
    if EPI_NAME == "fundus":
        _CPU["src"] = rs.randint(0, 256, (8, C, H, W)).astype(np.float32)
        _CPU["par"] = rs.randint(0, 256, (8, C, H, W)).astype(np.float32)
    else:
        _CPU["src"] = (rs.rand(8, C, H, W) * 2 - 1).astype(np.float32)
        _CPU["par"] = (rs.rand(8, C, H, W) * 2 - 1).astype(np.float32)
    _CPU["epi"] = O.FUNDUS_EPILOGUE if EPI_NAME == "fundus" else O.PROSTATE_EPILOGUE


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _cpu_work(args):
    """One image exactly as Fundus_Multi.__getitem__ does it (fundus.py:213-223): partner amplitude,
    mix, clip, normalise.  float64 path (= numpy 1.x behaviour, the parity oracle)."""
    i, pidx, rk = args[:3]
    precision = args[3] if len(args) > 3 else "f64"
    O = _CPU["O"]
    y = O.mix_batch(_CPU["src"][i % 8][None], _CPU["par"], [pidx % 8], [rk], L, epi=_CPU["epi"], precision=precision)
    return float(y[0, 0, 0, 0])


def cpu_images_per_s(n_images, cores, reps=1, precision="f64"):
    """Images/s of the oracle on `cores` worker processes (the reference's DataLoader worker model,
    train.py:558-559); pool start-up excluded.  Returns (best img/s, list of per-rep seconds)."""
    import multiprocessing as mp
    pidx, rk = draw_selection(n_images, N_BANK, SEED)
    jobs = [(i, int(pidx[i]), int(rk[i]), precision) for i in range(n_images)]
    times = []
    if cores == 1:
        _cpu_init(SEED)
        for _ in range(reps):
            t0 = time.perf_counter()
            for j in jobs:
                _cpu_work(j)
            times.append(time.perf_counter() - t0)
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores, initializer=_cpu_init, initargs=(SEED,)) as pool:
            pool.map(_cpu_work, jobs[:cores])            # warm the workers
            for _ in range(reps):
                t0 = time.perf_counter()
                pool.map(_cpu_work, jobs, chunksize=max(1, n_images // (cores * 4)))
                times.append(time.perf_counter() - t0)
    return n_images / min(times), times


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(cores * 4, 64)                           # images per step (bounded sample of the 256-image batch)
    import multiprocessing as mp
    pidx, rk = draw_selection(sample, N_BANK, SEED)
    jobs = [(i, int(pidx[i]), int(rk[i])) for i in range(sample)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(SEED,)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_work, jobs, chunksize=max(1, sample // (cores * 4)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_work, jobs, chunksize=max(1, sample // (cores * 4)))
        dt = time.perf_counter() - t0
    ips = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "RAM images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference numpy path (oracle port) on host cores; "
                   "each step = %d images of the batch" % sample},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d images/step x %d steps, float64 numpy path, multiprocessing.Pool(%d)"
                                   % (sample, args.steps, cores)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "numpy": np.__version__, "cpu_model": cpu_model(),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            text, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            return out
        clocks, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in text.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                clocks.append(float(f[0]))
                mx = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if clocks:
            # samples under load = the upper half (the sampler also sees idle gaps between regions)
            hi = sorted(clocks)[len(clocks) // 2:]
            out.update(sm_mhz=statistics.median(hi), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(clocks),
                       power_w_max=max(power) if power else None)
        return out


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["mix_fused_bytes_per_launch"]
    except Exception:
        return None


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback). Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ramdsir_b200 as R
    from ramdsir_b200 import _lib, sharded
    from ramdsir_b200.host import HostPipeline

    b = R.band_halfwidth(H, W, L)
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    # A loader pipeline keeps several batches in flight: consecutive steps read DIFFERENT source batches (so no step
    # finds its input in L2 because the previous step just read it) and write one of three rotating output buffers.
    srcs = [synth(torch, (BATCH, C, H, W), dev, g) for _ in range(NSRC)]
    src = srcs[0]
    outs = [torch.empty_like(src) for _ in range(3)]
    out = outs[0]

    # amplitude bank: built sharded (each rank its rows), replicated once by an all-gather
    lo, hi = sharded.shard_range(N_BANK, rank, world)
    gb = torch.Generator(device=dev).manual_seed(SEED * 7 + 1)
    bank_src_all = None
    torch.cuda.synchronize()
    bank_imgs = synth(torch, (hi - lo, C, H, W), dev, gb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    bank = sharded.build_amp_bank_sharded(bank_imgs, N_BANK, L)
    e1.record()
    torch.cuda.synchronize()
    bank_ms = e0.elapsed_time(e1)
    del bank_imgs, bank_src_all

    pidx_h, rk_h = draw_selection(BATCH, N_BANK, SEED + rank)
    pidx, rk = torch.from_numpy(pidx_h).to(dev), torch.from_numpy(rk_h).to(dev)
    epi = R.FUNDUS_EPILOGUE if EPI_NAME == "fundus" else R.PROSTATE_EPILOGUE

    # Steps are independent batches (own output buffer each), as in a loader pipeline that keeps two
    # batches in flight: the launch is flagged RAMDSIR_FLAG_INDEPENDENT so it may start on SMs that the
    # previous launch has already vacated (--serial disables that).
    state = {"i": 0}

    def step():
        i = state["i"]
        state["i"] += 1
        R.ram_mix_batch(srcs[i % NSRC], bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=outs[i % 3],
                        serial=args.serial)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    # the same K steps with plain stream-ordered launches (reported next to the headline for transparency)
    serial_ms = None
    if not args.serial:
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            R.ram_mix_batch(srcs[i % NSRC], bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=outs[i % 3],
                            serial=True)
        e1.record()
        torch.cuda.synchronize()
        serial_ms = e0.elapsed_time(e1) / args.steps
    # uint8-resident sources (what a fundus loader actually holds, fundus.py:205-214): the kernel reads the bytes
    # itself, 5 instead of 8 algorithmic bytes per pixel.  Same K steps, same launch mode; reported beside the headline.
    u8_src = None
    if EPI_NAME == "fundus" and R.u8_supported(C, H, W, b):
        src_u8d = src.to(torch.uint8)
        o8 = outs                                   # reuse the rotating output buffers
        want0 = torch.empty_like(src)
        R.ram_mix_batch(src, bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=want0, serial=True)
        for i in range(3):
            R.ram_mix_batch(src_u8d, bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=o8[i % 3],
                            serial=args.serial)
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            R.ram_mix_batch(src_u8d, bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=o8[i % 3],
                            serial=args.serial)
        e1.record()
        torch.cuda.synchronize()
        u8_ms = e0.elapsed_time(e1) / args.steps
        u8_bytes = (C * H * W * 5 + C * (2 * b + 1) ** 2 * 4) * BATCH
        u8_src = {"value": BATCH / (u8_ms * 1e-3), "unit": "images/s", "ms_per_step": u8_ms,
                  "algorithmic_bytes_per_launch": u8_bytes, "achieved": u8_bytes / (u8_ms * 1e-3) / 1e9,
                  "bit_identical_to_f32_source": bool(torch.equal(o8[0], want0)),
                  "note": "rank 0; ramdsir_mix_dual_u8: sources resident in HBM as uint8 [B,C,H,W], float32 output"}
        del src_u8d, o8, want0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = BATCH * world / (ms_per_step * 1e-3)

    # end-to-end through the host-buffer C-ABI entry (pinned host -> device -> pinned host)
    src_h = src.cpu().pin_memory()
    out_h = torch.empty_like(src_h).pin_memory()
    e2e_steps = 1 if args.no_e2e else max(3, min(args.steps, 30))
    with HostPipeline(chunk_images=max(1, BATCH // 8), C=C, H=H, W=W, L=L) as pipe:
        for _ in range(3):
            pipe.mix(src_h, bank, pidx_h, rk_h, out_h, epilogue=epi)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.mix(src_h, bank, pidx_h, rk_h, out_h, epilogue=epi)
        e2e_s = time.perf_counter() - t0
        # same leg with uint8 host images (what a fundus loader holds, fundus.py:212): H2D bytes / 4
        e2e_u8 = None
        if EPI_NAME == "fundus":
            src_u8 = src_h.to(torch.uint8).pin_memory()
            for _ in range(2):
                pipe.mix(src_u8, bank, pidx_h, rk_h, out_h, epilogue=epi)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pipe.mix(src_u8, bank, pidx_h, rk_h, out_h, epilogue=epi)
            e2e_u8 = BATCH * e2e_steps / (time.perf_counter() - t0)
    R.ram_mix_batch(src, bank, pidx, rk, L=L, clip=epi[:2], scale=epi[2], shift=epi[3], out=outs[0], serial=True)
    torch.cuda.synchronize()
    same = bool(torch.equal(out_h, outs[0].cpu()))
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = BATCH * world * e2e_steps / float(t.item())
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = load_peak()
        abytes = algorithmic_bytes_per_image(C, H, W, b) * BATCH
        kernel_s = (ms / args.steps) * 1e-3                     # rank 0's own launches, one kernel per step
        achieved = abytes / kernel_s / 1e9
        line = {
            "metric": "RAM images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "bank_images": N_BANK,
                       "bank_build_allgather_ms": bank_ms, "seed": SEED,
                       "l2": "inputs %.0f MB + outputs %.0f MB per step exceed the 126 MB L2 (no flush needed)"
                             % (src.numel() * 4 / 1e6, src.numel() * 4 / 1e6),
                       "parallelism": "batch-sharded x%d, bank all-gather once" % world,
                       "launch_overlap": "off (--serial)" if args.serial else
                       "consecutive steps are independent batches (alternating output buffers) launched with "
                       "RAMDSIR_FLAG_INDEPENDENT: a step may start on SMs the previous step has vacated",
                       "kernel_path": {1: "fused", 2: "multipass", 3: "generic"}.get(R.kernel_path(C, H, W, b))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_s * 1e3},
            "serial": None if serial_ms is None else {
                "ms_per_step": serial_ms, "value": BATCH / (serial_ms * 1e-3),
                "frac": abytes / (serial_ms * 1e-3) / 1e9 / peak,
                "note": "rank 0, same K steps launched strictly stream-ordered (no overlap between steps)"},
            "u8_source": None if u8_src is None else dict(u8_src, frac=u8_src["achieved"] / peak),
            "fft_tflops_nominal": nominal_fft_flops_per_image(C, H, W) * value / 1e12,
            "e2e": {"value": e2e_value, "unit": "images/s",
                    "h2d_bytes_per_step": int(src_h.numel() * 4 + pidx_h.nbytes + rk_h.nbytes),
                    "d2h_bytes_per_step": int(out_h.numel() * 4), "steps": e2e_steps,
                    "matches_device_path": same, "timing": "wall clock around the blocking C-ABI call",
                    "uint8_input_value": e2e_u8,
                    "uint8_input_note": "rank 0, same call with uint8 host images (ramdsir_mix_host_u8): H2D bytes / 4"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n = max(64, cores * 8)
            ips, times = cpu_images_per_s(n, cores, reps=2)
            ips1, _ = cpu_images_per_s(16, 1, reps=1)
            ips32, _ = cpu_images_per_s(n, cores, reps=1, precision="native")
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "%d images of the workload x2 reps (best), float64 numpy path, "
                                              "multiprocessing.Pool(%d)" % (n, cores),
                                    "single_core_images_per_s": ips1,
                                    "native_float32_path_images_per_s": ips32,   # what numpy >= 2 does with float32 input
                                    "cpu_model": cpu_model(), "numpy": np.__version__}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--serial", action="store_true", help="plain stream-ordered launches (no overlap between steps)")
    ap.add_argument("--no-e2e", action="store_true", help="(dev) shorten the host-buffer end-to-end leg to one step")
    ap.add_argument("--workload", default="fundus", choices=sorted(WORKLOADS), help="(dev) default = BASELINE configs[1]")
    args = ap.parse_args()
    select_workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
