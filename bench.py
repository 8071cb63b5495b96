#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric on this repo's hot path.

Metric: image-pairs/sec (and matched-kpts/sec) on a synthetic 4096x3072 pair at ~30k keypoints per image
(BASELINE.json configs[2], "C3": MSER + HessAff, tensor-core NN, DEGENSAC-H), one pair = both images through
HessianAffine detection (Gaussian pyramid, Hessian response, 3x3x3 NMS, localisation, Baumberg) AND MSER detection
(component tree of both polarities, stability thresholds, region moments) -> dominant orientation -> RootSIFT
description, then exact FGINN matching per detector (tcgen05), duplicate filtering and LO-RANSAC homography + LAF
check over the union -- one iteration of mods.cpp's loop with the identity view of both detectors (synthesised
views are not built yet: see DESIGN.md "scope").  --no-mser gives the HessianAffine-only variant.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size WxH]

N > 1 (torchrun, one rank per GPU): every rank matches its own pairs (weak scaling, no collective on
the data path -- pairs are independent), time = max over ranks.

One JSON line on rank 0.  `value` = pairs/s with the images already in HBM; `e2e` = the same call with
pinned HOST images (H2D inside the timed region; results always come back to the host).  Timing rules
of the contract: W >= 3 warm-ups, CUDA events on the library's stream, inputs cycle through 3 different
pairs (300 MB of images, working set of the pyramid ~1.3 GB/image >> 126 MB L2), nvidia-smi clocks
sampled during the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOB_DENSITY = 1.5e-3   # blobs per pixel: calibrated so that HessianAffine finds ~30k keypoints at 4096x3072
N_PAIRS = 3
# ONE metric string for both arms (the driver only forms vs_reference ratios when metric, unit and direction are identical)
METRIC = "image-pairs/sec (4096x3072 synthetic pair, ~30k HessAff kpts/image; matched-kpts/sec in `matched_kpts_per_s`)"
GENERATOR = {"impl": "mods_b200/synth.py (numpy PCG64; SURVEY 8d asked for SplitMix64 in C++ -- same images for both arms, so ratios are unaffected)",
             "blob_density_per_px": BLOB_DENSITY, "calibration": "SURVEY 8d's 5e-3 blobs/px gave far more than 30k +-15 % HessianAffine keypoints on the "
             "identity view at 4096x3072; N_b alone was scaled to 1.5e-3 (~31.5k keypoints, oracle count)", "seeds": "pair i: A = 1 + 16 i, B = A + 1 (rank r adds 1000 r)"}


def workload_string(w, h, no_mser, wxbs=False):
    if wxbs:
        return ("C5: independent %dx%d synthetic pairs streamed through the dataset call (two pairs in flight per GPU), WxBS descriptor lists "
                "(Descriptors=RootSIFT,HalfRootSIFT, iters_mods_cviu_wxbs.ini: regions oriented modulo pi, both descriptors, four (descriptor, detector) "
                "matching groups), HessianAffine(FixedTh 5.3333)+Baumberg and MSER, identity view, FGINN 0.8, duplicate filter 2px, LO-RANSAC-H 3px + LAF "
                "check; pairs sharded over the ranks" % (w, h))
    return ("C3: %dx%d synthetic pair, HessianAffine(FixedTh 5.3333)+Baumberg and %s, orientation+RootSIFT, identity view, "
            "FGINN 0.8 exact NN per detector, duplicate filter 2px, LO-RANSAC-H 3px + LAF check"
            % (w, h, "HessianAffine only (--no-mser)" if no_mser else "MSER(min_margin 8, min_size 30, max_area 0.05)"))


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


def make_pairs(w, h, n_pairs, seed0=1):
    from mods_b200 import synth
    pairs = []
    for i in range(n_pairs):
        A = synth.blob_image(w, h, seed=seed0 + 16 * i, n_blobs=int(BLOB_DENSITY * w * h))
        B = synth.warp_image(A, synth.gt_homography(w, h), seed=seed0 + 16 * i + 1)
        pairs.append((A, B))
    return pairs


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md's clocks line).  Sampled in-process through NVML (nvidia_ml_py)
    every 100 ms.  A spawned `nvidia-smi -lms 100` attaches to the driver ~0.5 s after the fork, i.e. in the middle of the timed arms: one
    round-2 run measured the end-to-end arm at half speed that way (15.6 against 31.1 pairs/s device-resident); with the sampler already up
    (tools/e2e_probe.py) neither source costs anything.  `nvidia-smi` stays as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, force_smi=False, period_s=0.1):
        self.rows = []      # (sm_mhz, sm_max_mhz, [reason names])
        self.proc = None
        self.nvml = None
        self.gpu = gpu_index
        self.period = period_s
        self.stop_flag = threading.Event()
        self.source = None
        if not force_smi:
            try:
                import pynvml
                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = gpu_index
                if vis:
                    try:
                        phys = int(vis.split(",")[gpu_index])
                    except Exception:
                        phys = gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
                self.nvml = pynvml
            except Exception:
                self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        bits = ((n.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"), (n.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"), (n.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap"))
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((sm, self.sm_max, [name for bit, name in bits if r & bit]))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        if self.nvml is not None:
            self.source = "nvml"
            self.t = threading.Thread(target=self._sample_nvml, daemon=True)
            self.t.start()
            return
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            c = [v.strip() for v in line.split(",")]
            try:
                self.rows.append((float(c[1]), float(c[2]), [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8])
                                                             if v.lower().startswith("active")]))
            except Exception:
                pass

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no clock source (NVML and nvidia-smi unavailable)"])
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons.update(r[2])
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm), source=self.source)


# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import mods_b200 as mb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h = args.size
    pairs = make_pairs(w, h, N_PAIRS, seed0=1 + 1000 * rank)
    ctx = mb.Context(local_rank)
    cfg = mb.PairConfig.default()
    cfg.use_mser = 0 if args.no_mser else 1
    wxbs = args.workload == "c5"
    if wxbs:
        cfg.halfRootSIFT = 1   # Descriptors=RootSIFT,HalfRootSIFT (the WxBS tiers)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev_pairs = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in pairs]
    pin_pairs = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ver_cap = [0]   # capacity of the verified-rows output of the un-pipelined call (parity check of pair 0)

    def timed(buffers, steps, pipelined=True, collective=True):
        """K steps = K pairs.  collective=False: rank-local (the profiling pass runs on rank 0 only -- no barrier, no all-reduce).  pipelined: ONE mb2_mods_pairs call over the K pairs (the dataset entry point: verification of
        pair k overlaps detection of pair k+1); otherwise K separate mb2_mods_pair calls (per-pair latency)."""
        results = []
        barrier() if collective else torch.cuda.synchronize()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        if pipelined:
            results, _ = ctx.mods_pairs([buffers[s % len(buffers)] for s in range(steps)], cfg, shapes=[((h, w), (h, w))] * steps)
        else:
            for s in range(steps):
                a, b = buffers[s % len(buffers)]
                res, ver = ctx.mods_pair(a, b, cfg, shape1=(h, w), shape2=(h, w), capacity=ver_cap[0])
                results.append(res)
                if ver_cap[0] and s == 0:
                    results[0].ver_rows = ver
        e1.record(stream)
        barrier() if collective else torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1 and collective:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, results, ctx.launches - l0

    # warm-up (allocations, module load) then the two timed arms
    timed(dev_pairs, max(3, args.warmup))
    timed(pin_pairs, max(3, args.warmup))   # the end-to-end arm gets its own warm-up (staging buffers of the upload path)
    timed(dev_pairs, 2, pipelined=False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev, res_dev, launches = timed(dev_pairs, args.steps)
    ms_e2e, wall_e2e, res_e2e, _ = timed(pin_pairs, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_lat, _, res_lat, _ = timed(dev_pairs, min(args.steps, 2 * N_PAIRS), pipelined=False)
    ver_cap[0] = 1 << 17
    _, _, res_par, _ = timed(dev_pairs, 1, pipelined=False, collective=False)   # pair 0 once more, verified rows kept (parity check below)
    ver_cap[0] = 0

    # per-kernel durations (CUDA events inside the library, same stream) on extra steps
    prof = None
    if rank == 0 and hasattr(ctx, "profile_begin"):
        ctx.profile_begin()
        timed(dev_pairs, min(args.steps, N_PAIRS), pipelined=False, collective=False)
        prof = ctx.profile_end()

    if world > 1:   # every rank: the other ranks wait here while rank 0 finishes its (rank-local) profiling pass
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = load_peaks()
    K = args.steps
    verified = float(np.mean([r.verified for r in res_dev]))
    regions = float(np.mean([r.regions1 + r.regions2 for r in res_dev])) / 2
    tent = float(np.mean([r.tentatives for r in res_dev]))
    pairs_s = world * K / (ms_dev / 1e3)
    e2e_s = world * K / (ms_e2e / 1e3)
    out = {
        "metric": METRIC,
        "value": pairs_s, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (pyramid/patches), f64 (SIFT sums, RANSAC), bf16->f32 tcgen05 (NN, exact on u8 descriptors)",
        "data": "synthetic (numpy PCG64 blob images + ground-truth homography warp, mods_b200/synth.py)",
        "config": {"workload": workload_string(w, h, args.no_mser, wxbs), "generator": GENERATOR,
                   "mser_regions_per_image": float(np.mean([r.mser_regions1 + r.mser_regions2 for r in res_dev])) / 2,
                   "mser_tentatives": float(np.mean([r.mser_tentatives for r in res_dev])),
                   "pairs_in_rotation": N_PAIRS, "l2": "inputs cycle through %d pairs (%.0f MB) and the per-image pyramid working set (~1.3 GB) exceeds the 126 MB L2"
                   % (N_PAIRS, N_PAIRS * 2 * w * h * 4 / 1e6),
                   "regions_per_image": regions, "tentatives": tent, "verified": verified, "parallelism": "pairs sharded over ranks, no collective",
                   "pipelining": "one mb2_mods_pairs call per timed region: duplicate filter + LO-RANSAC of pair k run on a second host thread / stream "
                                 "while pair k+1 is detected; `latency_ms_per_pair` is the un-pipelined mb2_mods_pair call"},
        "matched_kpts_per_s": verified * pairs_s,
        "e2e": {"value": e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": 2 * w * h * 4,
                "d2h_bytes_per_step": int(2 * regions * (2 * 72 + 128) + tent * 56 + 64),
                "ms_per_step": ms_e2e / K, "matched_kpts_per_s": verified * e2e_s,
                "entry": "mb2_mods_pairs (libmods_host.so) over the K pairs, pinned host images, results (regions, tentatives, H, verified) on the host"},
        "latency_ms_per_pair": ms_lat / max(1, len(res_lat)),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stage_ms": {k: float(np.mean([getattr(r, k) for r in res_lat])) for k in ("ms_detect_describe", "ms_match", "ms_duplicate", "ms_ransac", "ms_total")},
    }
    if prof:
        out.update(roofline_from_profile(prof, w, h, regions, peaks, steps=min(args.steps, N_PAIRS),
                                         mser_regions=float(np.mean([r.mser_regions1 + r.mser_regions2 for r in res_dev])) / 2))
    else:
        out["roofline"] = None
    if world == 1:
        try:
            out["micro"] = micro_benchmarks(ctx)
        except Exception as e:   # the micro-benchmarks never take the headline down
            out["micro"] = {"error": repr(e)}
    # results must be plausible: a failing kernel must not show up as a faster pairs/s figure
    for r in res_dev + res_e2e + res_lat:
        if not (r.regions1 > 0 and r.regions2 > 0 and r.tentatives > 0):
            raise SystemExit("bench.py: a pair came back with no regions / tentatives (regions %d / %d, tentatives %d)" % (r.regions1, r.regions2, r.tentatives))
    out["cpu_baseline"] = None
    out["parity"] = {"checked": False}
    if not args.no_cpu_baseline and world == 1 and not wxbs:   # rank 0 at N=1 only (the reference arm has no WxBS descriptor-list leg: C5 reports no CPU baseline): one full-size pair on the host cores, its outputs compared with the GPU's
        out["cpu_baseline"], cpu = cpu_baseline(pairs[0], w, h, cfg_values(cfg), with_mser=not args.no_mser)
        try:
            out["parity"] = parity_check(local_rank, pairs[0], w, h, cfg, cpu, res_par[0], getattr(res_par[0], "ver_rows", None))
        except Exception as e:
            out["parity"] = {"checked": True, "ok": False, "error": repr(e)}
    emit_json_line(out)


C4_WORKLOAD = ("C4: %dx%d synthetic pair, full SIFT view tiers of iters_mods_cviu.ini ([HessianAffine4..6] 61 views + [MSER2..3] 27 views per image), "
               "views sharded over the ranks, one NCCL all-gather of 184-byte region records, row-sharded exact FGINN, verification on rank 0")


def run_ours_c4(args, rank, world, local_rank):
    """BASELINE config 4 through the view-sharded driver (libmods_host.so): K pairs through mb2_views_sharded_pairs (throughput; strong scaling: the
    same K pairs on every world size) and single pairs through mb2_views_sharded_pair (latency)."""
    import torch
    import torch.distributed as dist
    import mods_b200 as mb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h = args.size
    A, B = make_pairs(w, h, 1)[0]
    ctx = mb.Context(local_rank)
    cfg = mb.PairConfig.default(); cfg.use_mser = 1; cfg.mserMatchRatio = 0.85   # iters_mods_cviu.ini:36 ([MSER2] FGINNThreshold)
    hess, mser = mb.iters_mods_cviu_views("c4")
    cfg.set_views(hess, mser)
    comm = ctx.dist_comm_create(rank, world) if world > 1 else None
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev = (torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
    pin = (torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory())

    step_ms = []   # host clock around every call (diagnostics: warm-up, resident-image steps, host-image steps, profiled step)

    def timed(imgs, steps, dataset=False):
        """dataset: ONE mb2_views_sharded_pairs call over `steps` pairs (pair k verified by rank k % world on a helper thread while all ranks go on
        with pair k + 1); otherwise `steps` separate mb2_views_sharded_pair calls (latency of one pair, verification on rank 0)."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = []
        if dataset:
            t0 = time.perf_counter()
            rs, vs, ds, ss = ctx.views_sharded_pairs([imgs] * steps, cfg, comm, rank, world, shapes=[((h, w), (h, w))] * steps, capacity=1 << 18)
            out = [(rs[k], vs[k], ds[k], ss[k]) for k in range(steps)]
            step_ms.append(round((time.perf_counter() - t0) * 1e3, 1))
        for _ in range(0 if dataset else steps):
            t0 = time.perf_counter()
            out.append(ctx.views_sharded_pair(imgs[0], imgs[1], cfg, comm, rank, world, shape1=(h, w), shape2=(h, w), capacity=1 << 18))
            step_ms.append(round((time.perf_counter() - t0) * 1e3, 1))
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, out, ctx.launches - l0

    timed(dev, max(3, args.warmup))
    timed(pin, max(3, args.warmup), dataset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, out_dev, launches = timed(dev, args.steps, dataset=True)
    ms_e2e, out_e2e, _ = timed(pin, args.steps, dataset=True)
    clocks = sampler.stop() if rank == 0 else None
    ms_one, out_one, _ = timed(dev, min(args.steps, 3))   # latency of one pair: separate calls, verification on rank 0
    prof = None
    if rank == 0:
        ctx.profile_begin()
    timed(dev, 1)           # collective: every rank takes part, rank 0 with per-kernel events
    if rank == 0:
        prof = ctx.profile_end()
    digs = [out_dev[-1][2]]
    if world > 1:
        g = [None] * world; dist.all_gather_object(g, out_dev[-1][2]); digs = g
        dist.barrier(); ctx.dist_comm_destroy(comm); dist.destroy_process_group()
    if rank != 0:
        return
    res, ver, dig, st = out_one[-1]
    if not (res.regions1 > 0 and res.tentatives > 0 and res.verified > 0):
        raise SystemExit("bench.py: the C4 pair came back empty")
    f = lambda r: (r.regions1, r.regions2, r.tentatives, r.unique_tentatives, r.ransac_inliers, r.verified)
    for o in out_dev + out_e2e:   # the dataset call gives every pair the result of the single call, on every rank
        if f(o[0]) != f(res) or o[2] != dig:
            raise SystemExit("bench.py: the dataset call differs from the single-pair call: %r vs %r" % (f(o[0]), f(res)))
    peaks = load_peaks()
    K = args.steps
    v = K / (ms_dev / 1e3); e = K / (ms_e2e / 1e3)
    out = {"metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": max(3, args.warmup), "ms_per_step": ms_dev / K,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 (pyramid/patches), f64 (SIFT sums, RANSAC), bf16->f32 tcgen05 (NN, exact on u8 descriptors)",
           "data": "synthetic (numpy PCG64 blob images + ground-truth homography warp, mods_b200/synth.py)",
           "config": {"workload": C4_WORKLOAD % (w, h), "generator": GENERATOR, "views_per_image": [len(hess), len(mser)], "regions": [res.regions1, res.regions2],
                      "tentatives": res.tentatives, "unique": res.unique_tentatives, "verified": res.verified, "parallelism": "views sharded over ranks (longest "
                      "processing time first), ONE ncclAllGather of device-resident region records + count / tentative exchanges; `value` / `e2e`: ONE "
                      "mb2_views_sharded_pairs call over the K pairs -- pair k is verified by rank k % world on a helper thread while all ranks go on with pair "
                      "k + 1; `single_pair`: separate mb2_views_sharded_pair calls, verification on rank 0",
                      "single_pair": {"ms_per_pair": ms_one / max(1, len(out_one)), "pairs_per_s": len(out_one) / (ms_one / 1e3)},
                      "l2": "176 view pipelines per step, working set far beyond the 126 MB L2",
                      "digest": ["%016x" % d for d in dig], "digest_identical_on_all_ranks": all(d == dig for d in digs),
                      "allgather_bytes_per_rank": st["allgather_bytes_per_rank"], "verify_ms_rank0": {"duplicate_filter": res.ms_duplicate, "lo_ransac_and_laf": res.ms_ransac}, "step_ms_rank0": step_ms, "steps_rank0_ms": [[round(o[3][k], 1) for k in ("ms_views", "ms_gather", "ms_match", "ms_tentative_gather", "ms_verify")] for o in out_dev + out_one], "rank0_ms": {k: st[k] for k in ("ms_views", "ms_gather", "ms_match", "ms_tentative_gather", "ms_verify")}},
           "matched_kpts_per_s": res.verified * v,
           "e2e": {"value": e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * w * h * 4, "d2h_bytes_per_step": int(res.tentatives * 56 + (res.regions1 + res.regions2) * 184 + res.verified * 32),
                   "ms_per_step": ms_e2e / K, "entry": "mb2_views_sharded_pairs (libmods_host.so), pinned host images on every rank, results on every rank, verified lists on the verifying rank"},
           "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None, "parity": {"checked": False, "note": "C4 parity: tests/test_gpu_fullsize.py "
                      "(cat pair, 11-view tier vs the compiled reference) and the digest, identical for every world size"}}
    if prof:
        out.update(roofline_from_profile(prof, w, h, (res.regions1 + res.regions2) / 2, peaks, steps=1, mser_regions=(res.mser_regions1 + res.mser_regions2) / 2))
    emit_json_line(out)


def roofline_from_profile(prof, w, h, regions, peaks, steps, mser_regions=0.0):
    """prof: {kernel name: (launches, total ms)} over `steps` pairs (2 images each)."""
    prof = dict(prof)
    gather_bytes = prof.pop("__extract_gather_bytes__", (1, 0.0))[1]
    tot = sum(v[1] for v in prof.values()) or 1.0
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    kernels = [{"kernel": k, "launches": v[0], "ms_per_pair": v[1] / steps, "share": v[1] / tot} for k, v in top[:16]]
    name, (n_launch, ms) = top[0]
    tree_ms = sum(v[1] for k, v in prof.items() if "k_mtree" in k)
    if tree_ms >= ms and tree_ms > 0:     # the component-tree pass (five kernels) is the dominant logical kernel even when one other kernel tops its parts
        name, (n_launch, ms) = max(((k, v) for k, v in prof.items() if "k_mtree" in k), key=lambda kv: kv[1][1])
    avg_s = ms / 1e3 / max(1, n_launch)
    rl = None
    if "k_extract" in name:
        # gather bytes per region = 4*((P+2)^2*4 taps + 1681*4) + 128 out; P from the mean scale is not
        # known here, so we use the measured scratch-plan average written by the library into the profile (bytes)
        gb = gather_bytes / max(1, n_launch)   # algorithmic bytes per launch
        rl = {"bound": "hbm", "kernel": name, "achieved": gb / 1e9 / avg_s if gb else None, "peak": peaks["hbm"], "unit": "GB/s",
              "frac": (gb / 1e9 / avg_s / peaks["hbm"]) if gb else None, "traffic": None,
              "note": "L2-gather / FP32-ALU kernel: algorithmic bytes = bilinear taps read + patch written (SURVEY 8d), peak = %s HBM copy" % peaks["src"]}
    elif "k_mtree" in name:
        # The dominant work is the component tree of the MSER detector, now five kernels per image (tiles in shared memory, local sums,
        # tile borders x 2, fix, gap walk): reported as ONE pass against SURVEY 8d's figure for it, ~30 B/px per polarity (u8 read + sorted
        # offset + label R/W + boundary pass), both polarities of one image per launch group.
        tree = {k: v for k, v in prof.items() if "k_mtree" in k}
        groups = max(v[0] for k, v in tree.items() if "k_mtree_tiles" in k)          # launch groups (= images) in the profiled steps
        ms_group = sum(v[1] for v in tree.values()) / max(1, groups)
        gb = 30.0 * 2 * w * h
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2c_traffic.json")) as f:
                traffic = json.load(f).get("component_tree_%dx%d" % (w, h))
        except Exception:
            pass
        rl = {"bound": "hbm", "kernel": "component tree of one image, both polarities: " + " + ".join(sorted(k.split("<")[0] for k in tree)),
              "achieved": gb / 1e9 / (ms_group / 1e3), "peak": peaks["hbm"], "unit": "GB/s", "frac": gb / 1e9 / (ms_group / 1e3) / peaks["hbm"],
              "traffic": traffic, "traffic_src": "dram__bytes_read.sum + dram__bytes_write.sum of the five kernels, ncu --set full (profiles/r2c_full_tree.md, profiles/r2c_traffic.json)",
              "algorithmic_bytes": gb, "ms_per_launch_group": ms_group, "launches_per_group": sum(v[0] for v in tree.values()) / max(1, groups),
              "top_kernel": {"name": name, "ms_per_launch": ms / max(1, n_launch), "algorithmic_bytes": 14.0 * w * h,
                             "achieved_GBps": 14.0 * w * h / 1e9 / (ms / 1e3 / max(1, n_launch)),
                             "note": "k_mtree_tiles alone: reads the f32 image (4 B/px), writes level + tree word of both polarities (10 B/px); issue bound "
                                     "(67 pct issue slots, 7 of 32 lanes active: divergent pointer walks in shared memory), not bandwidth bound"},
              "note": "lock-free merge of per-tile trees (mser_tree_build.cuh); round 1's level-synchronous k_mser_tree moved 10.3x its algorithmic bytes at 1.0 pct "
                      "of the HBM peak; peak = " + peaks["src"] + " HBM copy"}
    elif "k_blur_hess" in name:
        # per image: every octave runs 4 incremental blurs with the Hessian fused (read 4 B + write blur 4 B + write response 4 B per pixel
        # of the octave, octaves sum to 4/3 W H) plus the initial blur of octave 0 (read 4 B + write 4 B): DESIGN.md, kernel table
        gb = (12.0 * 4 * (4.0 / 3.0) + 8.0) * w * h * 2 * steps   # bytes over all launches of the profiled steps (2 images per pair)
        rl = {"bound": "hbm", "kernel": name, "achieved": gb / 1e9 / (ms / 1e3), "peak": peaks["hbm"], "unit": "GB/s",
              "frac": gb / 1e9 / (ms / 1e3) / peaks["hbm"], "traffic": None, "algorithmic_bytes_per_launch": gb / max(1, n_launch)}
    elif "k_nms" in name:
        gb = 12.0 * 3 * (4.0 / 3.0) * w * h * 2 * steps            # 3 detection levels per octave, 3 response planes read per level
        rl = {"bound": "hbm", "kernel": name, "achieved": gb / 1e9 / (ms / 1e3), "peak": peaks["hbm"], "unit": "GB/s",
              "frac": gb / 1e9 / (ms / 1e3) / peaks["hbm"], "traffic": None, "algorithmic_bytes_per_launch": gb / max(1, n_launch)}
    elif "k_nn_tc" in name:
        rl = {"bound": "tensor", "kernel": name, "achieved": None, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": None, "traffic": None}
    if rl is None:   # a per-region gather kernel on top (orientation / Baumberg / SIFT): L2-gather / FP32 bound, no HBM model -- say so instead of omitting the key
        rl = {"bound": "hbm", "kernel": name, "achieved": None, "peak": peaks["hbm"], "unit": "GB/s", "frac": None, "traffic": None,
              "note": "per-region gather kernel (L2 / FP32 bound); see roofline_pyramid and roofline_nn for the two kernels the north star names"}
    # the two kernels the north star names, always reported
    extra = {}
    pyr = [(k, v) for k, v in prof.items() if "k_blur_hess" in k or "k_nms" in k or "k_hessian" in k or "k_resize_half" in k]   # k_blur_hess_tma, k_nms_finish included
    if pyr:
        ms_pyr = sum(v[1] for _, v in pyr) / (2 * steps)  # per image
        bytes_pyr = 110.7 * w * h                          # level-granular model, B/px
        extra["roofline_pyramid"] = {"bound": "hbm", "achieved": bytes_pyr / 1e9 / (ms_pyr / 1e3), "peak": peaks["hbm"], "unit": "GB/s",
                                     "frac": bytes_pyr / 1e9 / (ms_pyr / 1e3) / peaks["hbm"], "ms_per_image": ms_pyr,
                                     "algorithmic_bytes": bytes_pyr, "kernels": "k_blur_hess_tma*+k_hessian+k_resize_half+k_nms_finish (whole scale space of one image; TMA-staged tiles, in-level extremum test fused)",
                                     "compulsory_bytes_if_fully_fused": 30.7 * w * h,
                                     "note": "issue bound, not bandwidth bound: the parity contract forbids FMA (2 FP32 instructions per tap) and packed f32x2 "
                                             "arithmetic gives no extra lane throughput on B200 (tools/micro/f32x2.cu: 36 T lane-ops/s either way)"}
    nn = [(k, v) for k, v in prof.items() if "k_nn_tc" in k]
    if nn:
        ms_nn = sum(v[1] for _, v in nn) / steps
        nh = regions - mser_regions                        # matching is per detector: N_hess^2 + N_mser^2 distance pairs
        flop = 2.0 * 2.0 * (nh * nh + mser_regions * mser_regions) * 128   # two streaming passes over each N1 x N2 x 128 contraction
        extra["roofline_nn"] = {"bound": "tensor", "achieved": flop / 1e12 / (ms_nn / 1e3), "peak": peaks["bf16"], "unit": "TFLOP/s",
                                "frac": flop / 1e12 / (ms_nn / 1e3) / peaks["bf16"], "ms_per_pair": ms_nn, "flop": flop,
                                "note": "2 passes (NN, then FGINN statistics); epilogue on CUDA cores is fused (distances never leave the SM)"}
    return dict(roofline=rl, kernels=kernels, **extra)


def micro_benchmarks(ctx):
    """SURVEY 8d micro-benchmarks of the verification rows (a17-a19): the batched scorer on n = 8192 correspondences x K = 4096
    hypotheses (kernel time from the library's CUDA events), and the F-matrix LO-RANSAC driver on 30k tentatives."""
    import mods_b200.synth as synth
    rng = np.random.default_rng(1)
    n, K = 8192, 4096
    u = np.zeros((n, 6)); u[:, 0:2] = rng.random((n, 2)) * 1000; u[:, 2] = 1; u[:, 5] = 1
    Hgt = synth.gt_homography(1000, 1000)
    p = (Hgt @ u[:, 0:3].T).T; u[:, 3:5] = p[:, :2] / p[:, 2:3] + rng.normal(size=(n, 2))
    u[int(0.6 * n):, 3:5] = rng.random((n - int(0.6 * n), 2)) * 1000
    M0 = np.linalg.inv(Hgt).T.ravel()
    models = np.stack([M0 * (1 + 1e-3 * rng.normal(size=9)) for _ in range(K)])
    out = {}
    fp64 = ctx.fp64_peak()   # measured with a DFMA micro-kernel in this run (MEASURED_PEAKS.json has no FP64 figure)
    for which, name, flop in ((0, "HDs", 120.0), (3, "FDs", 40.0)):
        ctx.score_models(which, u, models, 9.0)
        ctx.profile_begin()
        for _ in range(5):
            ctx.score_models(which, u, models, 9.0)
        prof = ctx.profile_end()
        ms = prof.get("k_score", (1, 0.0))[1] / 5
        out["scorer_" + name] = {"n": n, "K": K, "kernel_ms": ms, "achieved": flop * n * K / 1e12 / (ms / 1e3) if ms else None, "unit": "TFLOP/s (f64)",
                                 "peak": fp64, "peak_src": "measured in this run: DFMA micro-kernel of the library (mb2_debug_fp64_peak), 8 chains per thread",
                                 "frac": (flop * n * K / 1e12 / (ms / 1e3) / fp64) if ms else None, "flop_per_point_model": flop}
    # F driver at C3 size: 30k tentatives, 40 % outliers, inlLimit 0 as LORANSACFiltering calls it
    X = np.c_[rng.random((30000, 2)) * 6 - 3, rng.random(30000) * 10 + 2]
    Km = np.array([[800, 0, 400], [0, 800, 300], [0, 0, 1.0]]); a = 0.15
    Rm = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t = np.array([0.8, 0.1, 0.2])
    x1 = (Km @ X.T).T; x1 = x1[:, :2] / x1[:, 2:]
    x2 = (Km @ (Rm @ X.T + t[:, None])).T; x2 = x2[:, :2] / x2[:, 2:]
    uf = np.ones((30000, 6)); uf[:, 0:2] = x1 + rng.normal(size=(30000, 2)) * 0.5; uf[:, 3:5] = x2 + rng.normal(size=(30000, 2)) * 0.5
    uf[18000:, 3:5] = rng.random((12000, 2)) * 800
    uf = np.ascontiguousarray(uf[rng.permutation(30000)])
    ctx.ransac_f(uf, seed=9, inlLimit=0)
    t0 = time.perf_counter(); r = ctx.ransac_f(uf, seed=9, inlLimit=0); dt = time.perf_counter() - t0
    out["ransac_f"] = {"tentatives": 30000, "ms": 1e3 * dt, "inliers": int(r["I"]), "samples": r["samples"], "lo": r["lo"], "launches": r["launches"]}
    try:   # CPU-baseline leg of the micro-benchmark: the compiled reference timed beside the driver on the same tentatives (checker / baseline only)
        from oracle import pyoracle
        if pyoracle.have_reference():
            R = pyoracle.Reference()
            t0 = time.perf_counter(); rr = R.exp_ransacF(uf, seed=9, inlLimit=0); dr = time.perf_counter() - t0
            out["ransac_f"].update({"reference_cpu_ms": 1e3 * dr, "identical_to_reference": bool(np.array_equal(rr["inl"], r["inl"]) and rr["samples"] == r["samples"])})
    except Exception:
        pass
    return out


def cfg_values(cfg):
    """The scalar settings of an mb2_pair_config the CPU arm needs (plain dict: the reference arm must not load the GPU library)."""
    if cfg is None:   # mb2_pair_config_default (config_iter_mods_cviu.ini / iters_mods_cviu.ini)
        return dict(matchRatio=0.8, mserMatchRatio=0.8, contradDist=30.0, duplicateDist=2.0, err_threshold=3.0, confidence=0.99, HLAFCoef=12.0, LAFCoef=2.0,
                    max_samples=100000, errorType=0, doSymmCheck=1, seed=1, useF=0, localOptimization=1)
    return {k: getattr(cfg, k) for k in ("matchRatio", "mserMatchRatio", "contradDist", "duplicateDist", "err_threshold", "confidence", "HLAFCoef", "LAFCoef",
                                         "max_samples", "errorType", "doSymmCheck", "seed", "useF", "localOptimization")}


def cpu_pair(pair, with_mser=True, cv=None):
    """ONE full-size pair through the reference's CPU path, un-cropped and un-extrapolated (one iteration of mods.cpp:229-415):
    detect -> orient -> describe of (image, detector) on one thread each -- 2 images x 2 detectors, as mods.cpp:255-271 +
    imagerepresentation.cpp:612 parallelise them -- by the reference's own sources compiled in place (oracle/_ref); then the reference's
    own matching.cpp: MatchFlannFGINN over ALL queries per detector (exact linear k-NN on all host threads standing in for FLANN, which is
    not in the reference tree), DuplicateFiltering, LORANSACFiltering (exp_ransacHcustom + NaiveHCheck + H_LAF_check).
    Falls back to the oracle port (views + FGINN only) when oracle/_ref is absent."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())   # torchrun pins it to 1; set before the OpenMP runtime of the CPU code starts
    from oracle import pyoracle
    cv = cv or cfg_values(None)
    use_ref = pyoracle.have_reference()
    L = pyoracle.Reference() if use_ref else pyoracle.Oracle()
    A, B = pair
    dets = (0, 3) if with_mser else (0,)
    views = {}
    t0 = time.perf_counter()
    th = [threading.Thread(target=lambda k=(i, d), im=im: views.__setitem__(k, L.view_pipeline(im, detector=k[1]))) for i, im in enumerate((A, B)) for d in dets]
    [t.start() for t in th]; [t.join() for t in th]
    t_views = time.perf_counter() - t0
    groups = [(views[(0, d)][1], views[(0, d)][2], views[(1, d)][1], views[(1, d)][2], cv["matchRatio"] if d == 0 else cv["mserMatchRatio"]) for d in dets]
    t0 = time.perf_counter()
    if use_ref:
        back = L.pair_back(groups, contradDist=cv["contradDist"], duplicateDist=cv["duplicateDist"], useF=cv["useF"], err_threshold=cv["err_threshold"],
                           confidence=cv["confidence"], max_samples=cv["max_samples"], localOptimization=cv["localOptimization"], LAFCoef=cv["LAFCoef"],
                           HLAFCoef=cv["HLAFCoef"], errorType=cv["errorType"], doSymmCheck=cv["doSymmCheck"], seed=cv["seed"])
    else:
        tent = [np.c_[np.full(len(m), g), m] for g, m in enumerate(L.match_fginn(gr[1], gr[3], np.ascontiguousarray(gr[2][:, :2]), ratio=gr[4],
                                                                                 contradDist=cv["contradDist"]) for gr in groups)]
        back = dict(tent=np.concatenate(tent), kept=None, verified=None, H=None, counts=[sum(len(t) for t in tent), 0, 0, 0])
    t_back = time.perf_counter() - t0
    return dict(t_views=t_views, t_back=t_back, t_pair=t_views + t_back, views=views, groups=groups, back=back, kind="reference" if use_ref else "port",
                threads_views=len(th), dets=dets)


def cpu_sample_text(r, w, h):
    return ("ONE full %dx%d pair, nothing cropped or extrapolated: %s view pipelines (detect -> orient -> describe) of %d (image, detector) units on %d threads "
            "%.2f s; %s %.2f s -> %d tentatives, %d unique, %d RANSAC inliers, %d verified"
            % (w, h, "oracle/_ref (the reference's sources compiled in place)" if r["kind"] == "reference" else "oracle port", r["threads_views"], r["threads_views"],
               r["t_views"], "matching.cpp MatchFlannFGINN over all queries (exact linear k-NN on all %d host threads in place of FLANN) + DuplicateFiltering + "
               "LORANSACFiltering" % os.cpu_count() if r["kind"] == "reference" else "FGINN port over all queries (duplicate filter / RANSAC not run)",
               r["t_back"], *r["back"]["counts"]))


def cpu_baseline(pair, w, h, cv, with_mser=True):
    r = cpu_pair(pair, with_mser, cv)
    return {"value": 1.0 / r["t_pair"], "unit": "pairs/s", "cores": os.cpu_count(), "kind": r["kind"], "sample": cpu_sample_text(r, w, h),
            "ms_views": 1e3 * r["t_views"], "ms_match_verify": 1e3 * r["t_back"]}, r


def parity_check(local_rank, pair, w, h, cfg, cpu, res0, ver0):
    """GPU outputs of pair 0 against the CPU reference run of the SAME pair in the SAME process (the cpu_baseline leg's outputs):
    regions (det_kp, reproj_kp: 9 doubles each) and descriptors of both images and both detectors, tentatives per detector
    (q, idx0, idxJ, idx1, d0, dJ, d1), then the counts and the verified list of the whole mb2_mods_pair call -- all bit-exact."""
    import mods_b200 as mb
    out = {"checked": True, "config": "C3 %dx%d, pair 0" % (w, h), "against": "oracle/_ref (compiled reference)" if cpu["kind"] == "reference" else "oracle port", "details": {}}
    ok = True
    ctx = mb.Context(local_rank)
    try:
        A, B = pair
        for d in cpu["dets"]:
            name = "HessianAffine" if d == 0 else "MSER"
            det = mb.HessaffParams.default() if d == 0 else mb.MserParams.default()
            for i, im in enumerate((A, B)):
                g = ctx.detect_describe_view(im, det=det, ori=cfg.ori, desc=cfg.desc, slot=i)
                o = cpu["views"][(i, d)]
                same = (len(g[0]) == len(o[0]) and np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]) and np.array_equal(g[2].astype(np.float32), o[2]))
                out["details"]["%s_image%d" % (name, i + 1)] = {"regions": int(len(g[0])), "reference_regions": int(len(o[0])), "bit_exact": bool(same)}
                ok &= bool(same)
            gi = list(cpu["dets"]).index(d)
            gm = ctx.match_slots(0, 1, ratio=cpu["groups"][gi][4], contradDist=cfg.contradDist)
            om = cpu["back"]["tent"]; om = om[om[:, 0] == gi][:, 1:]
            same = len(gm) == len(om) and np.array_equal(gm, om)
            out["details"]["%s_tentatives" % name] = {"n": int(len(gm)), "reference_n": int(len(om)), "bit_exact": bool(same)}
            ok &= bool(same)
        if cpu["back"]["kept"] is not None:
            c = cpu["back"]["counts"]
            mine = [res0.tentatives, res0.unique_tentatives, res0.ransac_inliers, res0.verified]
            rows = cpu["back"]["tent"][cpu["back"]["verified"]]
            exp = np.array([np.r_[cpu["groups"][int(r[0])][0][int(r[1]), :2], cpu["groups"][int(r[0])][2][int(r[2]), :2]] for r in rows]).reshape(-1, 4)
            same_list = ver0 is not None and len(ver0) == len(exp) and np.array_equal(ver0, exp)
            Hdiff = float(np.max(np.abs(np.array(res0.H) / res0.H[8] - cpu["back"]["H"] / cpu["back"]["H"][8]))) if c[3] else None
            out["details"]["pair_chain"] = {"tentatives_unique_inliers_verified": mine, "reference": list(map(int, c)), "verified_list_identical": bool(same_list),
                                            "H_max_abs_diff": Hdiff, "seed": int(cfg.seed)}
            ok &= mine == list(map(int, c)) and bool(same_list)
    finally:
        ctx.close()
    out["ok"] = bool(ok)
    return out


# --------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on this box's host cores, on the full-size workload (cpu_pair): every step is
    one whole 4096x3072 pair.  --steps / --warmup are honoured while the run fits MB2_REF_BUDGET_S (default 240 s of wall clock, a step
    takes ~10 s); otherwise warm-ups drop to one and steps to what fits -- the line reports the steps actually run."""
    if rank != 0:
        return
    w, h = args.size
    t_start = time.perf_counter()
    budget = float(os.environ.get("MB2_REF_BUDGET_S", "240"))
    pair = make_pairs(w, h, 1)[0]
    cv = cfg_values(None)
    first = cpu_pair(pair, not args.no_mser, cv)          # warm-up step 1 (also sizes the rest of the run)
    t_step = first["t_pair"]
    left = budget - (time.perf_counter() - t_start)
    warm = max(1, args.warmup)
    steps = max(1, args.steps)
    if (warm - 1 + steps) * t_step > left:
        warm = 1
        steps = max(1, min(steps, int(left / t_step)))
    for _ in range(warm - 1):
        cpu_pair(pair, not args.no_mser, cv)
    t0 = time.perf_counter()
    runs = [cpu_pair(pair, not args.no_mser, cv) for _ in range(steps)]
    t_tot = time.perf_counter() - t0
    t_pair = t_tot / steps
    v = 1.0 / t_pair
    r = runs[-1]
    verified = r["back"]["counts"][3]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_pair, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64 (CPU)", "data": "synthetic (same generator and seeds as the GPU arm, pair 0)",
           "config": {"workload": workload_string(w, h, args.no_mser), "generator": GENERATOR, "requested_steps": args.steps, "requested_warmup": args.warmup,
                      "budget_s": budget, "regions_per_image": float(np.mean([len(r["views"][k][0]) for k in r["views"]])) * len(r["dets"]),
                      "tentatives": int(r["back"]["counts"][0]), "verified": int(verified)},
           "matched_kpts_per_s": verified * v,
           "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": os.cpu_count(), "kind": r["kind"], "sample": cpu_sample_text(r, w, h),
                            "ms_views": 1e3 * float(np.mean([x["t_views"] for x in runs])), "ms_match_verify": 1e3 * float(np.mean([x["t_back"] for x in runs]))},
           "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json_line(out)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries write to fd 1 during the run (NCCL prints its version banner there at communicator creation) goes to stderr,
    so that the ONE JSON line is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(out):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(out), flush=True)


def main():
    quiet_stdout()
    if os.environ.get("MB2_DUMP_AFTER"):   # diagnostics: Python stacks of all threads if the run is still alive after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["MB2_DUMP_AFTER"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="4096x3072", type=lambda s: tuple(int(v) for v in s.lower().split("x")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mser", action="store_true", help="HessianAffine-only variant of the workload")
    ap.add_argument("--workload", default=os.environ.get("MB2_BENCH_WORKLOAD", "c3"), choices=["c3", "c4", "c5"],
                    help="c3 (default): BASELINE config 3, pairs sharded over the ranks (weak scaling).  c4: BASELINE config 4, the full view tiers of "
                         "iters_mods_cviu.ini on 4096x3072 pairs, views sharded over the ranks + one NCCL all-gather (strong scaling).  c5: BASELINE "
                         "config 5, independent 1920x1080 pairs with the WxBS descriptor lists, pairs sharded over the ranks (weak scaling)")
    args = ap.parse_args()
    if args.workload == "c5" and "--size" not in sys.argv:
        args.size = (1920, 1080)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl != "reference":
        # torchrun pins OMP_NUM_THREADS to 1; the host legs of the library (duplicate filter, libm legs, LO-RANSAC set-up) use OpenMP:
        # give every rank its share of the host cores (set before any OpenMP runtime is loaded)
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, world)))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "c4":
        run_ours_c4(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
