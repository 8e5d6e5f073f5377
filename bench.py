#!/usr/bin/env python
"""bench.py -- RAM images/sec on B200 (BASELINE.json metric) with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one RAM pass over one batch of synthetic fundus-shaped images (BASELINE.json configs[1]:
256 x 3 x 256 x 256 fp32, L = 0.1 -> b = 25, fundus clip/normalise epilogue) per GPU.  Weak scaling:
every rank mixes its own 256-image shard against a bank that was built sharded and all-gathered once
before the timed region.  One JSON line is printed by rank 0.

value    = images/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks).  The steps
           rotate over distinct source batches and three output buffers, like a loader that keeps batches in flight;
           the library overlaps the tail of one launch with the head of the next by itself when their buffers are
           disjoint (include/ramdsir.h, "Overlap of consecutive launches"); "serial" is the same loop with that off.
e2e      = images/s through the host-buffer C-ABI entry (ramdsir_mix_host_f32): pinned host src ->
           H2D -> kernel -> D2H -> pinned host out, every step, wall-clocked around the blocking call
roofline = algorithmic bytes (2*C*H*W*4 + C*(2b+1)^2*4 per image) / measured kernel time
           against MEASURED_PEAKS.json's HBM copy bandwidth
configs  = the other BASELINE.json configurations (prostate 512x1x384x384, hi-res 64x3x1024x1024, the 8192-image
           batch of configs[3] sharded over the N ranks, the amplitude-bank builder), each with its own roofline
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
NSRC = 2                          # distinct source batches the timed steps rotate over
NOUT = 3                          # rotating output buffers (a launch may overlap its predecessor, never the one before)
SPINUP_MS = float(os.environ.get("RAMDSIR_BENCH_SPINUP_MS", "30"))   # device time of back-to-back warm-up launches in front of every timed region (MixRun.timed)
CONFIG4_IMAGES = 8192             # BASELINE configs[3]: global batch, sharded over the ranks
CONFIG4_BANK = 1024

WORKLOADS = {
    "fundus": dict(C=3, H=256, W=256, BATCH=256, N_BANK=768, N_DOMAINS=3, EPI="fundus",
                   NAME="fundus RAM batch 256x3x256x256 fp32, L=0.1 (b=25), 3 source domains, clip[0,255]/127.5-1"),
    "prostate": dict(C=1, H=384, W=384, BATCH=512, N_BANK=640, N_DOMAINS=5, EPI="prostate",
                     NAME="prostate RAM batch 512x1x384x384 fp32, L=0.1 (b=38), 5 source domains, clip[-1,1]"),
    "prostate3": dict(C=3, H=384, W=384, BATCH=128, N_BANK=320, N_DOMAINS=5, EPI="prostate",
                      NAME="prostate RAM batch 128x3x384x384 fp32 (real 3-slice layout), L=0.1 (b=38), clip[-1,1]"),
    "hires": dict(C=3, H=1024, W=1024, BATCH=64, N_BANK=48, N_DOMAINS=3, EPI="fundus",
                  NAME="high-res fundus RAM batch 64x3x1024x1024 fp32, L=0.1 (b=102), clip[0,255]/127.5-1"),
}


class WL:
    """One workload (shape, batch, bank, epilogue)."""

    def __init__(self, name, batch=None, n_bank=None, label=None):
        w = WORKLOADS[name]
        self.key = name
        self.C, self.H, self.W = w["C"], w["H"], w["W"]
        self.BATCH = w["BATCH"] if batch is None else batch
        self.N_BANK = w["N_BANK"] if n_bank is None else n_bank
        self.N_DOMAINS = w["N_DOMAINS"]
        self.EPI = w["EPI"]
        self.NAME = w["NAME"] if label is None else label


def synth(torch, wl, n, dev, gen):
    """Synthetic images: uint8-valued floats for fundus (fundus.py:212), uniform [-1, 1] for prostate."""
    shape = (n, wl.C, wl.H, wl.W)
    if wl.EPI == "fundus":
        return torch.randint(0, 256, shape, device=dev, generator=gen).float()
    return torch.rand(shape, device=dev, generator=gen) * 2.0 - 1.0


def algorithmic_bytes_per_image(c, h, w, b):
    return 2 * c * h * w * 4 + c * (2 * b + 1) ** 2 * 4


def bank_bytes_per_image(c, h, w, b):
    return c * h * w * 4 + c * (2 * b + 1) ** 2 * 4


def nominal_fft_flops_per_image(c, h, w):
    return c * 2 * 5 * h * w * np.log2(h * w)


def draw_selection(wl, batch, seed):
    """Partner rows and ratio numerators with the reference's RNG calls (fundus.py:205,208,35)."""
    from ramdsir_b200 import rng
    rng.seed_worker_streams(seed)
    per = wl.N_BANK // wl.N_DOMAINS
    names = ["Domain%d" % (i + 1) for i in range(wl.N_DOMAINS)]
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


def _cpu_init(seed, key):
    from oracle import ram_oracle as O
    wl = WL(key)
    rs = np.random.RandomState(seed)
    _CPU["O"] = O
    shape = (8, wl.C, wl.H, wl.W)
    if wl.EPI == "fundus":
        _CPU["src"] = rs.randint(0, 256, shape).astype(np.float32)
        _CPU["par"] = rs.randint(0, 256, shape).astype(np.float32)
    else:
        _CPU["src"] = (rs.rand(*shape) * 2 - 1).astype(np.float32)
        _CPU["par"] = (rs.rand(*shape) * 2 - 1).astype(np.float32)
    _CPU["epi"] = O.FUNDUS_EPILOGUE if wl.EPI == "fundus" else O.PROSTATE_EPILOGUE


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


def cpu_images_per_s(wl, n_images, cores, reps=1, precision="f64"):
    """Images/s of the oracle on `cores` worker processes (the reference's DataLoader worker model,
    train.py:558-559); pool start-up excluded.  Returns (best img/s, list of per-rep seconds)."""
    import multiprocessing as mp
    pidx, rk = draw_selection(wl, n_images, SEED)
    jobs = [(i, int(pidx[i]), int(rk[i]), precision) for i in range(n_images)]
    times = []
    if cores == 1:
        _cpu_init(SEED, wl.key)
        for _ in range(reps):
            t0 = time.perf_counter()
            for j in jobs:
                _cpu_work(j)
            times.append(time.perf_counter() - t0)
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores, initializer=_cpu_init, initargs=(SEED, wl.key)) as pool:
            pool.map(_cpu_work, jobs[:cores])            # warm the workers
            for _ in range(reps):
                t0 = time.perf_counter()
                pool.map(_cpu_work, jobs, chunksize=max(1, n_images // (cores * 4)))
                times.append(time.perf_counter() - t0)
    return n_images / min(times), times


def run_reference_arm(args, rank, world):
    """The reference's CPU implementation of the path (oracle port) on all host cores, on OUR arm's config: every
    step is the full 256-image batch of configs[1]."""
    if rank != 0:
        return
    wl = WL(args.workload)
    cores = os.cpu_count() or 1
    sample = wl.BATCH
    import multiprocessing as mp
    pidx, rk = draw_selection(wl, sample, SEED)
    jobs = [(i, int(pidx[i]), int(rk[i])) for i in range(sample)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(SEED, wl.key)) as pool:
        for _ in range(max(min(args.warmup, 2), 1)):
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
        "config": {"workload": wl.NAME, "batch_per_gpu": wl.BATCH, "global_batch": wl.BATCH, "bank_images": wl.N_BANK,
                   "seed": SEED,
                   "note": "reference numpy path (oracle port of fundus.py:13-61 + caller epilogue) on the host cores of "
                           "rank 0; each step = the full %d-image batch" % sample},
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


def load_traffic(key="mix_fused_bytes_per_launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
    except Exception:
        return None


def pin_to_cores(local_rank, world):
    """Give every rank its own slice of the host cores (copy-engine feeding and the oracle pool stay off each other)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


class MixRun:
    """Device-resident buffers of one workload on this rank + the timed loops over them."""

    def __init__(self, torch, R, sharded, dist, wl, dev, rank, world, nsrc=NSRC, nout=NOUT):
        self.torch, self.R, self.wl, self.dev, self.rank, self.world, self.dist = torch, R, wl, dev, rank, world, dist
        self.b = R.band_halfwidth(wl.H, wl.W, L)
        g = torch.Generator(device=dev).manual_seed(SEED + rank)
        self.srcs = [synth(torch, wl, wl.BATCH, dev, g) for _ in range(nsrc)]
        self.outs = [torch.empty_like(self.srcs[0]) for _ in range(nout)]
        # amplitude bank: built sharded (each rank its rows), replicated once by an all-gather
        lo, hi = sharded.shard_range(wl.N_BANK, rank, world)
        gb = torch.Generator(device=dev).manual_seed(SEED * 7 + 1 + 1000 * rank)
        self.bank_imgs = synth(torch, wl, hi - lo, dev, gb)
        self.sharded = sharded
        self.bank = None
        self.pidx_h, self.rk_h = draw_selection(wl, wl.BATCH, SEED + rank)
        self.pidx, self.rk = torch.from_numpy(self.pidx_h).to(dev), torch.from_numpy(self.rk_h).to(dev)
        self.epi = R.FUNDUS_EPILOGUE if wl.EPI == "fundus" else R.PROSTATE_EPILOGUE
        self.i = 0
        self.calls = {}              # bound launches of the (source, output) rotation, see step()
        from ramdsir_b200 import _lib
        self.lib = _lib

    def _sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def build_bank(self):
        """Cold first call (tables, allocator, NCCL connect) is discarded; the timed calls are warm.  Returns
        (kernel-only ms of this rank's shard, ms of build + all-gather, max over ranks done by the caller)."""
        torch = self.torch
        wl = self.wl
        self.calls = {}
        self.bank = self.sharded.build_amp_bank_sharded(self.bank_imgs, wl.N_BANK, L)      # cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best_k = best_all = 1e9
        local = torch.empty((self.bank_imgs.shape[0], wl.C, 2 * self.b + 1, 2 * self.b + 1), device=self.dev)
        for _ in range(3):
            self._sync()
            e0.record()
            self.R.build_amp_bank(self.bank_imgs, L, out=local)
            e1.record()
            torch.cuda.synchronize()
            best_k = min(best_k, e0.elapsed_time(e1))
            self._sync()
            e0.record()
            self.bank = self.sharded.build_amp_bank_sharded(self.bank_imgs, wl.N_BANK, L)
            e1.record()
            torch.cuda.synchronize()
            best_all = min(best_all, e0.elapsed_time(e1))
        del local
        return best_k, best_all

    def bank_bit_identical(self):
        """Rebuild the WHOLE bank on this rank from all ranks' source images and compare with the gathered one."""
        torch, wl = self.torch, self.wl
        parts = []
        for r in range(self.world):
            lo, hi = self.sharded.shard_range(wl.N_BANK, r, self.world)
            gb = torch.Generator(device=self.dev).manual_seed(SEED * 7 + 1 + 1000 * r)
            parts.append(self.R.build_amp_bank(synth(torch, wl, hi - lo, self.dev, gb), L))
        ok = bool(torch.equal(torch.cat(parts, 0), self.bank))
        t = torch.tensor([1 if ok else 0], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def step(self, srcs=None, serial=False):
        """One pass over one batch.  The (source, output) pairs of the rotation are bound once as `MixCall`s (the
        package's prepared form of ram_mix_batch: same checks, same C-ABI call, bit-identical results), so that the
        host side of a step is a single ramdsir_mix_dual_* call and the launch queue stays ahead of 88-us kernels even
        with 8 ranks sharing the box's cores."""
        i = self.i
        self.i += 1
        ss = self.srcs if srcs is None else srcs
        key = (id(ss), i % len(ss), i % len(self.outs), bool(serial))
        call = self.calls.get(key)
        if call is None:
            e = self.epi
            call = self.calls[key] = self.R.MixCall(ss[key[1]], self.bank, self.pidx, self.rk, self.outs[key[2]], L=L,
                                                    clip=e[:2], scale=e[2], shift=e[3], serial=serial)
        call()

    def timed(self, steps, warmup, srcs=None, launches_at=None, **kw):
        """ms per step on this rank (CUDA events on the launching stream), after at least `warmup` untimed steps.

        The warm-up is sized in TIME as well as in steps: the ranks are aligned first (a rank that finished its setup
        early would otherwise idle through the others'), then `warmup` steps run and are timed, then as many more as
        fill SPINUP_MS of device time, back to back.  A B200 that sat idle through host-side setup needs tens of
        milliseconds of load to return to its full clocks (nvidia-smi right after a run shows the idle GPUs of a box at
        700-1000 MHz): at 8 GPUs the 1024^2 leg measured 0.694 ms per step without this and 0.604 with it.  (More than
        ~50 ms of uninterrupted load moves the GPU into its power-capped regime instead -- sw_power_cap, -1.4 % -- which is
        not the regime MEASURED_PEAKS.json's best-of-10 copy bandwidth was taken in; 30 ms stays clear of it.)"""
        torch = self.torch
        self._sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(int(warmup), 1)
        e0.record()
        for _ in range(n):
            self.step(srcs, **kw)
        e1.record()
        torch.cuda.synchronize()
        per = max(e0.elapsed_time(e1) / n, 1e-3)
        for _ in range(int(min(SPINUP_MS / per, 4000))):
            self.step(srcs, **kw)
        self._sync()
        if launches_at is not None:
            launches_at.append(self.lib.launch_count())
        e0.record()
        for _ in range(steps):
            self.step(srcs, **kw)
        e1.record()
        if launches_at is not None:
            launches_at.append(self.lib.launch_count())
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        return e0.elapsed_time(e1) / steps

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world == 1:
            return [float(x)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def free(self):
        self.calls = {}
        self.srcs = self.outs = self.bank = self.bank_imgs = None
        self.torch.cuda.empty_cache()


def config_line(run, ms_rank, ms_max, peak, serial_ms=None, extra=None):
    wl = run.wl
    ab = algorithmic_bytes_per_image(wl.C, wl.H, wl.W, run.b) * wl.BATCH
    ach = ab / (ms_rank * 1e-3) / 1e9
    d = {"workload": wl.NAME, "batch_per_gpu": wl.BATCH, "value": wl.BATCH * run.world / (ms_max * 1e-3), "unit": "images/s",
         "ms_per_step": ms_max,
         "kernel_path": {1: "fused", 2: "multipass", 3: "generic"}.get(run.R.kernel_path(wl.C, wl.H, wl.W, run.b)),
         "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "algorithmic_bytes_per_launch": ab, "kernel_ms": ms_rank}}
    if serial_ms is not None:
        d["serial"] = {"ms_per_step": serial_ms, "frac": ab / (serial_ms * 1e-3) / 1e9 / peak}
    if extra:
        d.update(extra)
    return d


def train_step_leg(steps=4):
    """BASELINE configs[4], second half: one `train.py --ram --rec` step (tools/e2e_step.py restates train.py:223-296 around
    a small consumer network) fed by the GPU-RAM loader (Fundus_Multi(ram="deferred") + RamBatchHook) and, as the CPU
    baseline of the same step, by per-sample numpy RAM (the oracle) on identical samples and RNG draws.  Part of the
    cpu_baseline leg: rank 0, N = 1 only."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import e2e_step as E
    import fixture_data as FX
    from oracle import ram_oracle as O

    def oracle_ram(src, partner, k):
        return O.mix_batch(src[None], partner[None], [0], [k], L, epi=O.FUNDUS_EPILOGUE)[0]

    import contextlib
    import io
    with tempfile.TemporaryDirectory(prefix="ramdsir_e2e_") as root, contextlib.redirect_stdout(io.StringIO()):
        FX.make_fundus(root, counts=E.E2E_FIXTURE)
        res = E.compare(root, FX.ResizeJitter(), oracle_ram, steps=steps, num_workers=0)
    d0 = res["diff"][0]
    return {"batch": res["batch"], "steps": len(res["gpu"]),
            "loss_step0": res["gpu"][0]["loss"], "loss_rel_diff_step0": d0["loss_rel"], "grad_norm_rel_diff_step0": d0["grad_norm_rel"],
            "loss_rel_diff_max": max(d["loss_rel"] for d in res["diff"]),
            "loader_ms_per_step": {"gpu_ram": res["gpu_timing"]["data_ms_per_step"], "cpu_ram": res["cpu_timing"]["data_ms_per_step"]},
            "model_ms_per_step": res["gpu_timing"]["model_ms_per_step"],
            "note": "B = 3 + 6 + 7 fundus images per step (train.py:35) from a synthetic on-disk fixture; loader time = sample "
                    "load + transforms + RAM + H2D for one step, in the main process (num_workers=0); the first step includes "
                    "one-off table / allocator warm-up"}


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback). Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cores = pin_to_cores(local_rank, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import ramdsir_b200 as R
    from ramdsir_b200 import _lib, sharded
    from ramdsir_b200.host import HostPipeline

    peak, peak_src = load_peak()
    wl = WL(args.workload)
    W = max(args.warmup, 3)
    run = MixRun(torch, R, sharded, dist, wl, dev, rank, world)
    b = run.b
    bank_kernel_ms, bank_all_ms = run.build_bank()
    bank_kernel_ms, bank_all_ms = run.max_over_ranks(bank_kernel_ms), run.max_over_ranks(bank_all_ms)
    bank_same = run.bank_bit_identical() if world > 1 else None

    # ---- headline: K steps, device-resident, automatic tail overlap (or --serial) ----
    for _ in range(W):
        run.step(serial=args.serial)
    run._sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches_at = []
    ms = run.timed(args.steps, W, serial=args.serial, launches_at=launches_at)
    launches = launches_at[1] - launches_at[0]                 # launches inside the timed region only
    ms_max = run.max_over_ranks(ms)
    ms_ranks = run.all_ranks(ms)
    value = wl.BATCH * world / (ms_max * 1e-3)
    serial_ms = None if args.serial else run.timed(args.steps, 2, serial=True)
    # A K-step region starts and ends with the device drained, so it pays the fill and the drain of the launch
    # pipeline once (one tail nobody overlaps + the launches it takes the overlap to settle: ~60 us, i.e. 3.5 % of
    # 20 steps, 0.4 % of 200).  The steady-state rate of the same loop, over a region long enough to make that
    # negligible, is reported beside the contract's K-step value.
    STEADY_STEPS = 400
    steady_ms = steady_ms_max = None
    if not args.serial and args.steps < STEADY_STEPS:
        steady_ms = run.timed(STEADY_STEPS, 2, serial=False)
        steady_ms_max = run.max_over_ranks(steady_ms)

    # uint8-resident sources (what a fundus loader actually holds, fundus.py:205-214): the kernel reads the bytes
    # itself, 5 instead of 8 algorithmic bytes per pixel.  Same K steps, same launch mode; reported beside the headline.
    u8_src = None
    if wl.EPI == "fundus" and R.u8_supported(wl.C, wl.H, wl.W, b):
        s8 = [s.to(torch.uint8) for s in run.srcs]
        want0 = torch.empty_like(run.srcs[0])
        e = run.epi
        R.ram_mix_batch(run.srcs[0], run.bank, run.pidx, run.rk, L=L, clip=e[:2], scale=e[2], shift=e[3], out=want0, serial=True)
        got0 = torch.empty_like(want0)
        R.ram_mix_batch(s8[0], run.bank, run.pidx, run.rk, L=L, clip=e[:2], scale=e[2], shift=e[3], out=got0, serial=True)
        u8_ms = run.timed(args.steps, 3, srcs=s8, serial=args.serial)
        u8_bytes = (wl.C * wl.H * wl.W * 5 + wl.C * (2 * b + 1) ** 2 * 4) * wl.BATCH
        u8_src = {"value": wl.BATCH / (u8_ms * 1e-3), "unit": "images/s", "ms_per_step": u8_ms,
                  "algorithmic_bytes_per_launch": u8_bytes, "achieved": u8_bytes / (u8_ms * 1e-3) / 1e9,
                  "bit_identical_to_f32_source": bool(torch.equal(got0, want0)),
                  "note": "rank 0; ramdsir_mix_dual_u8: sources resident in HBM as uint8 [B,C,H,W], float32 output"}
        run.calls = {k: v for k, v in run.calls.items() if k[0] != id(s8)}     # the bound launches hold the uint8 copies
        del s8, want0, got0

    # ---- end-to-end through the host-buffer C-ABI entry (pinned host -> device -> pinned host) ----
    # The fundus loader holds uint8 images at this boundary (PIL -> np.array, BEFORE fundus.py:212's astype(float32)), so the
    # primary number feeds uint8 host images (ramdsir_mix_host_u8: the kernel widens them, bit-identical to the float32
    # call); the same call with float32 host images is reported beside it.  Prostate slices are float32 (.npy): f32 only.
    src_h = run.srcs[0].cpu().pin_memory()
    out_h = torch.empty_like(src_h).pin_memory()
    e2e_steps = 1 if args.no_e2e else max(3, min(args.steps, 30))
    chunk = args.e2e_chunk or max(1, min(32, wl.BATCH))          # 32-image chunks measured best (tools/e2e_chunks.py)
    link = None
    use_u8 = wl.EPI == "fundus"
    e = run.epi
    R.ram_mix_batch(run.srcs[0], run.bank, run.pidx, run.rk, L=L, clip=e[:2], scale=e[2], shift=e[3], out=run.outs[0], serial=True)
    torch.cuda.synchronize()
    want_h = run.outs[0].cpu()
    with HostPipeline(chunk_images=chunk, C=wl.C, H=wl.H, W=wl.W, L=L) as pipe:
        def leg(src):
            for _ in range(3):
                pipe.mix(src, run.bank, run.pidx_h, run.rk_h, out_h, epilogue=run.epi)
            run._sync()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pipe.mix(src, run.bank, run.pidx_h, run.rk_h, out_h, epilogue=run.epi)
            dt = run.max_over_ranks(time.perf_counter() - t0)
            return wl.BATCH * world * e2e_steps / dt, bool(torch.equal(out_h, want_h))
        e2e_f32, same_f32 = leg(src_h)
        e2e_u8, same_u8 = (None, None)
        if use_u8:
            src_u8 = src_h.to(torch.uint8).pin_memory()
            e2e_u8, same_u8 = leg(src_u8)
    e2e_value = e2e_u8 if use_u8 else e2e_f32
    same = same_u8 if use_u8 else same_f32
    h2d_bytes = int(src_h.numel() * (1 if use_u8 else 4) + run.pidx_h.nbytes + run.rk_h.nbytes)
    if not args.no_e2e:
        # what the host link gives this rank while every rank copies at once: the ceiling of any host-buffer path
        d_in, d_out = run.srcs[0], run.outs[0]
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        run._sync()
        t0 = time.perf_counter()
        for _ in range(3):
            with torch.cuda.stream(s1):
                d_in.copy_(src_h, non_blocking=True)
            with torch.cuda.stream(s2):
                out_h.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = run.max_over_ranks(time.perf_counter() - t0)
        gbs = 3 * src_h.numel() * 4 / dt / 1e9
        run._sync()
        t0 = time.perf_counter()
        for _ in range(3):
            out_h.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt1 = run.max_over_ranks(time.perf_counter() - t0)
        link = {"both_ways_gbs_per_rank_each_direction": gbs, "d2h_alone_gbs_per_rank": 3 * src_h.numel() * 4 / dt1 / 1e9,
                "ceiling_images_per_s_float32_in": wl.BATCH * world * 3 / dt,
                "ceiling_images_per_s_uint8_in": wl.BATCH * world * 3 / dt1,
                "note": "pinned host <-> device copies of one float32 batch, all ranks at once: both directions together "
                        "(bounds the float32-in call) and device -> host alone (bounds the uint8-in call, whose upload is 4x "
                        "smaller); ceiling = the rate at which batches can cross the host link at all"}
    clocks = sampler.stop() if sampler else None

    # ---- the other BASELINE configurations, each a few steps ----
    configs = {}
    if not args.no_configs:
        ksteps = max(5, min(args.steps, 40))
        configs["bank_builder_" + wl.key] = {
            "workload": "amplitude-bank builder (K3), %d source images of the %s shape per rank, warm" % (run.bank_imgs.shape[0], wl.key),
            "kernel_ms": bank_kernel_ms, "build_plus_allgather_ms": bank_all_ms,
            "value": run.bank_imgs.shape[0] * world / (bank_kernel_ms * 1e-3), "unit": "images/s",
            "roofline": (lambda ab: {"bound": "hbm", "achieved": ab / (bank_kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": ab / (bank_kernel_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": ab})(
                bank_bytes_per_image(wl.C, wl.H, wl.W, b) * run.bank_imgs.shape[0])}
        run.free()
        extra = [("prostate_512x1x384x384", WL("prostate")), ("hires_64x3x1024x1024", WL("hires")),
                 ("config4_8192_fundus_sharded", WL("fundus", batch=(CONFIG4_IMAGES + world - 1) // world, n_bank=CONFIG4_BANK,
                                                    label="BASELINE configs[3]: %d fundus images batch-sharded over %d rank(s), "
                                                          "bank of %d source images built sharded + all-gathered"
                                                          % (CONFIG4_IMAGES, world, CONFIG4_BANK)))]
        for key, w2 in extra:
            if w2.key == args.workload and w2.BATCH == wl.BATCH:
                continue
            r2 = MixRun(torch, R, sharded, dist, w2, dev, rank, world)
            bk, ball = r2.build_bank()
            bk, ball = r2.max_over_ranks(bk), r2.max_over_ranks(ball)
            same2 = r2.bank_bit_identical() if (world > 1 and key.startswith("config4")) else None
            m = r2.timed(ksteps, 3)
            ms2 = r2.timed(ksteps, 2, serial=True)
            x = {"steps": ksteps, "bank_images": w2.N_BANK, "bank_kernel_ms": bk, "bank_build_plus_allgather_ms": ball}
            if same2 is not None:
                x["bank_bit_identical"] = same2
            configs[key] = config_line(r2, m, r2.max_over_ranks(m), peak, serial_ms=ms2, extra=x)
            r2.free()
            del r2

    if rank == 0:
        abytes = algorithmic_bytes_per_image(wl.C, wl.H, wl.W, b) * wl.BATCH
        kernel_s = ms * 1e-3                                       # rank 0's own launches, one kernel per step
        achieved = abytes / kernel_s / 1e9
        line = {
            "metric": "RAM images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_max, "ms_per_step_ranks": ms_ranks, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.NAME, "batch_per_gpu": wl.BATCH, "global_batch": wl.BATCH * world, "bank_images": wl.N_BANK,
                       "bank_build_allgather_ms": bank_all_ms, "bank_kernel_ms": bank_kernel_ms, "seed": SEED,
                       "l2": "steps rotate over %d source batches (%.0f MB each) and %d output buffers (%.0f MB each): far "
                             "beyond the 126 MB L2, no step finds its input cached by the previous one (no flush needed)"
                             % (NSRC, wl.BATCH * wl.C * wl.H * wl.W * 4 / 1e6, NOUT, wl.BATCH * wl.C * wl.H * wl.W * 4 / 1e6),
                       "parallelism": "batch-sharded x%d, bank all-gather once" % world,
                       "launch_overlap": "off (--serial)" if args.serial else
                       "automatic: the library overlaps a launch with the tail of the previous one when their buffers are "
                       "disjoint (no flag from the caller); the later launch never completes before the earlier one",
                       "host_cores_per_rank": host_cores,
                       "warmup_note": "every timed region is preceded by the W warm-up steps plus %.0f ms of back-to-back launches "
                                      "(ranks aligned first), so that it is taken at the GPUs' steady clocks" % SPINUP_MS,
                       "launch_call": "ramdsir_b200.MixCall: each (source, output) pair of the rotation validated and bound "
                                      "once, a step = one ramdsir_mix_dual_f32 C-ABI call",
                       "kernel_path": {1: "fused", 2: "multipass", 3: "generic"}.get(R.kernel_path(wl.C, wl.H, wl.W, b))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": kernel_s * 1e3},
            "serial": None if serial_ms is None else {
                "ms_per_step": serial_ms, "value": wl.BATCH / (serial_ms * 1e-3),
                "frac": abytes / (serial_ms * 1e-3) / 1e9 / peak,
                "note": "rank 0, same K steps with RAMDSIR_FLAG_SERIAL (no overlap between steps)"},
            "steady_state": None if steady_ms is None else {
                "steps": STEADY_STEPS, "ms_per_step": steady_ms_max, "value": wl.BATCH * world / (steady_ms_max * 1e-3),
                "frac": abytes / (steady_ms * 1e-3) / 1e9 / peak,
                "note": "the same loop over one %d-step region (max over ranks; frac from rank 0): a K-step region pays the "
                        "fill and drain of the launch pipeline once, which is %.1f us per step at K = %d"
                        % (STEADY_STEPS, (ms_max - steady_ms_max) * 1e3, args.steps)},
            "u8_source": None if u8_src is None else dict(u8_src, frac=u8_src["achieved"] / peak),
            "fft_tflops_nominal": nominal_fft_flops_per_image(wl.C, wl.H, wl.W) * value / 1e12,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(out_h.numel() * 4), "steps": e2e_steps, "chunk_images": chunk,
                    "host_input_dtype": "uint8" if use_u8 else "float32",
                    "matches_device_path": same, "timing": "wall clock around the blocking C-ABI call, max over ranks",
                    "float32_input_value": e2e_f32, "float32_input_matches_device_path": same_f32,
                    "float32_input_h2d_bytes_per_step": int(src_h.numel() * 4 + run.pidx_h.nbytes + run.rk_h.nbytes),
                    "note": "ramdsir_mix_host_u8 / _f32: pinned host images -> chunked H2D -> RAM kernel -> chunked D2H -> "
                            "pinned host float32 output, every step",
                    "host_link": link},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if bank_same is not None:
            line["bank_bit_identical"] = bank_same
        if configs:
            line["configs"] = configs
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n = wl.BATCH
            ips, times = cpu_images_per_s(wl, n, cores, reps=2)
            ips1, _ = cpu_images_per_s(wl, 16, 1, reps=1)
            ips32, _ = cpu_images_per_s(wl, n, cores, reps=1, precision="native")
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "the %d-image batch x2 reps (best), float64 numpy path, "
                                              "multiprocessing.Pool(%d)" % (n, cores),
                                    "single_core_images_per_s": ips1,
                                    "native_float32_path_images_per_s": ips32,   # what numpy >= 2 does with float32 input
                                    "cpu_model": cpu_model(), "numpy": np.__version__}
            if not args.no_train_step:
                line["train_step"] = train_step_leg()
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
    ap.add_argument("--serial", action="store_true", help="RAMDSIR_FLAG_SERIAL on every step (no overlap between steps)")
    ap.add_argument("--no-e2e", action="store_true", help="(dev) shorten the host-buffer end-to-end leg to one step")
    ap.add_argument("--no-configs", action="store_true", help="(dev) skip the extra BASELINE configurations")
    ap.add_argument("--no-train-step", action="store_true", help="(dev) skip the config-5b train-step leg")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="(dev) images per chunk of the host pipeline")
    ap.add_argument("--workload", default="fundus", choices=sorted(WORKLOADS), help="(dev) default = BASELINE configs[1]")
    args = ap.parse_args()
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
