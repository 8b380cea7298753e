#!/usr/bin/env python
"""bench.py -- LK iterations/sec of the B200 hot path (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--scaling weak|strong]

Without --config the line is the headline (config 2, weak scaling) and carries the other BASELINE configurations, run with
fewer steps, under "other_configs" (benchlib.py: config 3 GridTracker cells, config 4 the 8192-patch MI batch sharded over
the ranks, config 5 the particle filter); --config N prints that configuration's line alone.

Workload (config.workload): BASELINE.json configs[1] -- FCLK + SSD + Homography, 1024 independent 50x50 patches
per GPU on 1024x1024 synthetic frames (mtf_b200/synth.py), fixed 30 Gauss-Newton iterations per patch per frame
(epsilon = 0 disables the early exit so that every step does identical work).  One STEP = one frame: the whole
batch tracked for 30 iterations = P * 30 LK iterations.

Three arms are timed in one process: the headline arm (--precision f32 --f32-solve reference: fp32 per-pixel arithmetic with
bit-exact sampling indices, the arithmetic north_star specifies, and the REFERENCE's column-pivoted QR solve in the reference's
parameters -- the arm parity is claimed for) fills the contract's keys; the F32 arm with the opt-in patch-local solve and the
F64 arm are reported under "other_arms" (value, e2e, roofline fraction, distance of their corners from the headline arm's and
from the ground truth).

  value      whole-job LK iterations/sec, frames already resident in HBM, CUDA-event timed per step on the
             launching stream, L2 flushed between steps, max over ranks
  e2e        the same through the reference-facing API with HOST buffers: pinned frame -> H2D -> update ->
             D2H of the P x 8 corners inside the timed region
  e2e_raw_u8 the same from the RAW uint8 frame MTF's pre-processor would receive: 1 MB upload, gray + 5 x 5 Gaussian on the
             device (mtfb_set_image_u8), update, D2H
  roofline   algorithmic HBM bytes ((8N + 432) per patch-iteration) / kernel time vs the measured
             copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference, kind "port") on the box's host cores, bounded sample

--impl reference times the oracle's reference-faithful loop (oracle/, all host threads) on bounded samples of
the same workload; rank 0 only.
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

RES = 50
N_PIX = RES * RES
ITERS = 30
P_PER_GPU = 1024
IMG = 1024
N_FRAMES = 8
ALG_BYTES_PER_ITER = 8 * N_PIX + 432          # I_t footprint + I_0 (fp32 each) + W in + J,H,f out
METRIC = "LK iters/sec (50x50 SSD+Homography)"
# ncu --set full, this workload: the frame and the template once, everything else stays on chip
NCU_DRAM_BYTES_PER_LAUNCH = {"f64": 24466688, "f32": 15141888}
NCU_TRAFFIC_SOURCE = {"f64": "profiles/r01_ncu_f64_summary.txt", "f32": "profiles/r02_ncu_mom_final_summary.txt"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), one sample every 20 ms with its arrival time;
    summary() keeps the samples that fall inside the timed windows."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def summary(self, windows):
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons, n_all = [], 0.0, set(), 0
        for t, r in self.rows:
            try:
                clk, mx = float(r[0]), float(r[1])
            except Exception:
                continue
            n_all += 1
            smax = max(smax, mx)
            if not any(a <= t <= b for a, b in windows):
                continue
            sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": n_all}


def workload(seed_offset=0):
    from mtf_b200 import synth
    cache = "/tmp/mtfb_bench_seq_%d_%d.npz" % (N_FRAMES, IMG)        # synthesis takes a few seconds per process
    global TRUE_WARPS
    try:
        z = np.load(cache)
        frames, TRUE_WARPS = list(z["frames"]), list(z["warps"])
    except Exception:
        frames, TRUE_WARPS = synth.make_sequence(N_FRAMES, IMG, IMG, seed=1234, walk_seed=5678, sigma=1.0)
        try:
            tmp = cache + ".%d.npz" % os.getpid()
            np.savez(tmp, frames=np.stack(frames), warps=np.stack(TRUE_WARPS))
            os.replace(tmp, cache)
        except Exception:
            pass
    corners = synth.make_patches(P_PER_GPU, 49.0, IMG, IMG, seed=42 + seed_offset)
    # ping-pong over frames 1 .. N_FRAMES-1 keeps consecutive frames one random-walk step apart for any number of steps.
    # Frame 0 only initialises ("frames 1..T"): tracking the template's own frame converges to the exact
    # identity, where an integer-aligned 49 px box puts every sample ON the pixel lattice -- the reference's straddling
    # finite difference for all 2500 pixels, a synthetic worst case timed separately (profiles/README.md)
    order = list(range(1, N_FRAMES)) + list(range(N_FRAMES - 2, 1, -1))
    return frames, corners, order


TRUE_WARPS = None


def truth_error(final, corners, frame_index):
    """|tracked corner - ground-truth corner| (px) per patch: the synthetic frames are frame 0 under known homographies"""
    from mtf_b200 import synth
    gt = synth.warp_corners(TRUE_WARPS[frame_index], corners)
    d = np.abs(final - gt).max(axis=(1, 2))
    return {"median": float(np.median(d)), "p99": float(np.percentile(d, 99)), "max": float(d.max())}


def oracle_params(fast_sums=1):
    from oracle import oracle_lib as O
    return O.make_params("ssd", "homography", "fclk", max_iters=ITERS, epsilon=0.0, grad_mode=0, fast_sums=fast_sums)


def cpu_sample(frames, corners, n_patches, n_frames, threads, fast_sums=1):
    """oracle batch driver (OpenMP over patches, GridTracker.cc:253-256) -> (iterations, seconds).
    fast_sums 1: vectorised dot products for J^T J (all S^2 entries, as Eigen's general product computes them);
    2: a cache-blocked SYRK on the upper triangle (less work than Eigen does: an upper bound on the reference's speed)"""
    from oracle import oracle_lib as O
    total, secs, _, _ = O.batch_track(oracle_params(fast_sums), frames[:n_frames + 1], corners[:n_patches], n_threads=threads)
    return total, secs


STAGES = ["updatePixVals", "updateSimilarity", "update*Grad", "updatePixGrad", "cmpt*PixJacobian", "cmpt*Jacobian", "cmpt*Hessian",
          "solve", "compositionalUpdate"]          # record_event labels of NT/FCLK.cc:190-321 / NT/ESM.cc:190-290


def cpu_config1(frames, corners):
    """BASELINE.json configs[0]: ESM + SSD + Homography, ONE 50 x 50 patch, 30 iterations per frame at most (epsilon 1e-4),
    one thread -- the reference's own CPU-runnable case -- with the per-stage wall clock of the oracle's update()"""
    from oracle import oracle_lib as O
    prm = O.make_params("ssd", "homography", "esm", max_iters=ITERS, epsilon=1e-4, grad_mode=0, fast_sums=1, hess_type=2, jac_type=1)
    t = O.OracleTracker(prm)
    t.set_image(frames[0]); t.initialize(corners[0])
    iters, secs, stages = 0, 0.0, np.zeros(9)
    for rep in range(3):
        for f in frames[1:]:
            t.set_image(f)
            t0 = time.perf_counter(); t.update(); secs += time.perf_counter() - t0
            iters += t.n_iters; stages += np.asarray(t.stage_times())
        t.set_image(frames[0]); t.set_region(corners[0])
    return {"workload": "ESM+SSD+Homography, 1 patch 50x50, <= 30 iters/frame (epsilon=1e-4), 1 thread", "value": iters / secs,
            "unit": "iters/s", "iterations": int(iters), "us_per_iteration": 1e6 * secs / iters,
            "stage_us_per_iteration": {k: 1e6 * float(v) / iters for k, v in zip(STAGES, stages)}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames, corners, _ = workload()
    # bounded sample: 2 patches per core x 1 frame x 30 iterations per step
    n_patches = min(P_PER_GPU, 2 * cores)
    best = None
    by_kernel = {}
    for fs, name in ((1, "dot_products_full_SxS"), (2, "blocked_syrk_upper_triangle")):
        for _ in range(args.warmup):
            cpu_sample(frames, corners, n_patches, 1, cores, fs)
        iters = secs = 0.0
        for _ in range(args.steps):
            a, b = cpu_sample(frames, corners, n_patches, 1, cores, fs)
            iters += a; secs += b
        by_kernel[name] = iters / secs
        if best is None or iters / secs > best[0]:
            best = (iters / secs, secs, name)
    v, secs, kernel = best
    sample = "%d patches x 1 frame x %d iterations per step, %d steps" % (n_patches, ITERS, args.steps)
    if args.reference_brief:
        print(json.dumps({"value": v, "by_hessian_kernel": by_kernel}))
        return
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "FCLK+SSD+Homography 50x50, %d iters/frame, CPU oracle (restatement of MTF, not MTF)" % ITERS,
                   "sample": sample},
        # value = the faster of the two J^T J kernels (Eigen's own GEMM lies between them: it is vectorised like the first and
        # computes all S^2 entries like the first; the second skips the lower triangle)
        "cpu_baseline": {"value": v, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample, "hessian_kernel": kernel,
                         "by_hessian_kernel": by_kernel, "flags": "-O3 -march=native -ffp-contract=off -fopenmp"},
        "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}
    try:
        out["config1_single_thread"] = cpu_config1(frames, corners)
    except Exception as e:
        out["config1_single_thread"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:
        # the same sample with the oracle built like the reference's makefile builds MTF (plain -O3, no -march=native)
        env = dict(os.environ, MTF_ORACLE_GENERIC="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--reference-brief", "--steps", str(args.steps),
                            "--warmup", str(args.warmup)], env=env, capture_output=True, text=True, timeout=600)
        out["generic_build_O3"] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        out["generic_build_O3"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(out))


def measure(args, precision, f32_solve, env, n_total, frames, corners, order, d_frames, pinned, raw_pinned):
    """one arm (precision 'f32' | 'f64', F32 solve 'reference' | 'local'): device-resident timing, then end to end from pinned
    host frames.  corners: this rank's patches; n_total: patches of the whole job (the all-gather's row count).
    -> dict(ms, kms, e2e_ms, launches, status, finite, windows)"""
    import torch
    from mtf_b200 import api, sharding
    dev, world, dist, flush = env.dev, env.world, env.dist, env.flush
    P = corners.shape[0]
    prm = api.make_params("ssd", "homography", "fclk", n_patches=P, max_iters=ITERS, epsilon=0.0, device=env.local_rank,
                          threads_per_patch=args.threads, occupancy=args.occ, precision=precision, f32_solve=f32_solve)
    tr = api.BatchTracker(prm)
    stream = torch.cuda.current_stream(dev)
    tr.set_stream(stream.cuda_stream)
    # zero-copy torch view of the library's P x 8 result array: what the all-gather reads (no host hop)
    d_corners = sharding.device_view(tr.device_results()[0], (P, 8), dev)

    gathered = torch.empty((n_total, 8), dtype=torch.float64, device=dev) if world > 1 else None
    peer = world > 1 and args.collective == "peer"
    peer_note = None
    if peer:
        # the library's own exchange over NVLink peer memory: the update kernel stores each patch's corners into the gathered
        # arrays of all ranks (CUDA IPC mappings), mtfb_peer_gather signals / waits -- no collective library on the path.
        # Should the mapping fail on ANY rank (no peer access between two of the GPUs, IPC disabled in the container), every
        # rank falls back to NCCL's all-gather together, and the line says so
        ok, why = 1, ""
        try:
            handle = tr.peer_export(n_total)
        except Exception as e:
            ok, why, handle = 0, "export: %s" % e, np.zeros(api.PEER_HANDLE_BYTES, dtype=np.uint8)
        handles = sharding.exchange_peer_handles(handle)
        if ok:
            try:
                if os.environ.get("MTFB_BENCH_PEER_FAIL") == str(env.rank):       # (test hook: exercise the fallback)
                    raise RuntimeError("forced by MTFB_BENCH_PEER_FAIL")
                tr.peer_attach(env.rank, world, sharding.shard_range(n_total, world, env.rank)[0], handles)
            except Exception as e:
                ok, why = 0, "attach: %s" % e
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = False
            peer_note = "peer-memory exchange unavailable (%s): NCCL all-gather instead" % (why or "another rank")
            tr.close()
            tr = api.BatchTracker(prm)
            tr.set_stream(stream.cuda_stream)
            d_corners = sharding.device_view(tr.device_results()[0], (P, 8), dev)

    def gather():
        # north_star: one all-gather of the per-patch results over NVLink (per frame: LK iterations of different patches
        # never interact)
        if peer:
            tr.peer_gather()
        elif world > 1:
            sharding.all_gather_rows(d_corners, n_total, out=gathered)

    tr.initialize(corners, d_frames[0])
    tr.synchronize()

    def step_device(i):
        tr.setImage(d_frames[order[i % len(order)]])
        tr.update()
        gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident timing
    for i in range(args.warmup):
        step_device(i)
    barrier()
    if peer:
        # the peer exchange against NCCL's all-gather of the same frame: identical bytes on every rank
        ptr, n = tr.peer_gathered_ptr()
        sharding.all_gather_rows(d_corners, n_total, out=gathered)
        same = bool(torch.equal(sharding.device_view(ptr, (n, 8), dev), gathered))
        if not same:
            raise RuntimeError("rank %d: the peer-memory exchange differs from NCCL's all-gather" % env.rank)
    launches0 = tr.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    win0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                                           # evict L2 between timed steps
        ev[i][0].record(stream)
        tr.setImage(d_frames[order[(args.warmup + i) % len(order)]])
        kev[i][0].record(stream)
        tr.update()
        kev[i][1].record(stream)
        gather()
        ev[i][1].record(stream)
    barrier()
    windows = [(win0, time.time())]
    launches = tr.launch_count - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kms = sum(a.elapsed_time(b) for a, b in kev)
    status = tr.patch_status()
    final = tr.getRegion()
    last_frame = order[(args.warmup + args.steps - 1) % len(order)]

    # ------------------------------------------------------------------ end to end (host buffers)
    # every step: one frame H2D from pinned host memory, the update, the P x 8 corners D2H, the caller waits for them.
    # "serial": the frame of step i is uploaded in front of update(i) on the same stream (mtfb_set_image);
    # "pipelined" (the e2e figure): frame i + 1 is handed over right after update(i) has been enqueued and uploads on the
    # library's copy stream while update(i) runs (mtfb_set_image_async, double-buffered) -- what a video pipeline does
    host_out = torch.empty((P, 8), dtype=torch.float64).pin_memory()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)

    def frame_ptr(i, raw=False):
        return (raw_pinned if raw else pinned)[order[i % len(order)]].data_ptr()

    def e2e_loop(raw, pipelined):
        def upload(i, prefetch):
            if raw:
                (tr.prefetch_raw_image_pinned if prefetch else tr.set_raw_image_pinned)(frame_ptr(i, True), IMG, IMG, IMG, 1)
            else:
                (tr.prefetch_image_pinned if prefetch else tr.set_image_pinned)(frame_ptr(i), IMG, IMG, IMG)

        def run(n, first):
            for i in range(first, first + n):
                if pipelined:
                    tr.update()                      # samples frame i (prefetched during step i - 1)
                    upload(i + 1, True)              # H2D of the next frame overlaps this update
                else:
                    upload(i, False)
                    tr.update()
                gather()
                host_out.copy_(d_corners, non_blocking=True)
                stream.synchronize()                 # the caller reads the corners
        if raw:
            tr.set_raw_image_pinned(raw_pinned[0].data_ptr(), IMG, IMG, IMG, 1)
            tr.initialize(corners)
        else:
            tr.initialize(corners, frames[0])
        if pipelined:
            upload(0, True)
        run(args.warmup, 0)
        barrier()
        w0 = time.time()
        wall0 = time.perf_counter()
        t0.record(stream)
        run(args.steps, args.warmup)
        t1.record(stream)
        barrier()
        tr.update()                                  # consumes the last prefetched frame (pipelined) before the next arm
        gather()
        tr.synchronize()
        return max(t0.elapsed_time(t1), 1e3 * (time.perf_counter() - wall0)), (w0, time.time())

    e2e_serial_ms, _ = e2e_loop(False, False)
    e2e_ms, win = e2e_loop(False, True)
    windows.append(win)
    # end to end from RAW frames: the host hands over the uint8 frame MTF's pre-processor would receive; gray
    # conversion and the 5 x 5 Gaussian run on the device behind a 1 MB upload (instead of a host cv::GaussianBlur + 4 MB)
    raw_ms, _ = e2e_loop(True, True)
    if world > 1:
        t = torch.tensor([ms, kms, e2e_ms, raw_ms, e2e_serial_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kms, e2e_ms, raw_ms, e2e_serial_ms = [float(x) for x in t.tolist()]
    tr.close()
    return dict(ms=ms, kms=kms, e2e_ms=e2e_ms, raw_ms=raw_ms, e2e_serial_ms=e2e_serial_ms, launches=int(launches), status=status, final=final, windows=windows,
                peer=peer, peer_note=peer_note,
                truth=truth_error(final, corners, last_frame))


def config2_line(args, env, sampler, strong):
    """the headline: FCLK + SSD + Homography, 1024 patches per GPU (weak) or 1024 in total (strong)"""
    import torch
    from mtf_b200 import sharding
    world, rank, dev = env.world, env.rank, env.dev
    if strong:
        frames, corners_all, order = workload(seed_offset=0)
        lo, hi = sharding.shard_range(P_PER_GPU, world, rank)
        corners, n_total = corners_all[lo:hi], P_PER_GPU
    else:
        frames, corners, order = workload(seed_offset=rank)
        n_total = P_PER_GPU * world
    P = corners.shape[0]
    if args.pitch_pad:
        # experiment: device frames with a row pitch of IMG + pad floats (L1 set-conflict study, profiles/README.md)
        d_frames = []
        for f in frames:
            buf = torch.empty((IMG, IMG + args.pitch_pad), dtype=torch.float32, device=dev)
            buf[:, :IMG].copy_(torch.from_numpy(f))
            d_frames.append(buf[:, :IMG])
    else:
        d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    # the raw uint8 frames a camera / video decoder would deliver (the synthetic frames are already smooth; the extra blur
    # only changes what is tracked, not the work)
    raw_pinned = [torch.from_numpy(np.clip(np.rint(f), 0, 255).astype(np.uint8)).pin_memory() for f in frames]
    common = (env, n_total, frames, corners, order, d_frames, pinned, raw_pinned)

    # headline arm: fp32 per-pixel arithmetic + the reference's solve (what parity is claimed for); the other arms are timed
    # in the same process and reported beside it
    arms = [(args.precision, args.f32_solve)]
    if not args.one_arm:
        arms += [a for a in (("f32", "reference"), ("f32", "local"), ("f64", "reference")) if a != arms[0] and not (a[0] == "f64" and arms[0][0] == "f64")]
    res = [measure(args, prec, solve, *common) for prec, solve in arms]
    main = res[0]
    windows = [w for m in res for w in m["windows"]]
    clocks = sampler.summary(windows)
    total_iters = n_total * ITERS * args.steps
    peak, peak_src = peaks()

    def kernel_name(precision):
        if precision == "f64":
            return "ssd_update_kernel<Homography,FCLK>"
        return ("ssd_update_f32_kernel" if os.environ.get("MTFB_F32_KERNEL") == "classic" else "ssd_fclk_mom_kernel") + "<Homography>"

    def roofline(m, precision):
        achieved = ALG_BYTES_PER_ITER * P * ITERS * args.steps / (m["kms"] * 1e-3) / 1e9        # per GPU
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH[precision],
                "traffic_source": NCU_TRAFFIC_SOURCE[precision] + " (dram__bytes_read + write, one launch)",
                "kernel": kernel_name(precision), "kernel_ms_per_launch": m["kms"] / args.steps,
                "alg_bytes_per_launch": ALG_BYTES_PER_ITER * P * ITERS, "peak_source": peak_src,
                "note": ("fp64-issue bound" if precision == "f64" else "latency / issue bound") + ", not HBM bound: see DESIGN.md"}

    def dtype(precision):
        return "f64" if precision == "f64" else "f32 per pixel (bit-exact sampling indices), f64 reduction + solve"

    status, final = main["status"], main["final"]
    out = {
        "metric": METRIC, "value": total_iters / (main["ms"] * 1e-3), "unit": "iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": dtype(arms[0][0]), "data": "synthetic",
        "config": {"workload": "FCLK+SSD+Homography, %d patches%s 50x50, %d iters/frame (epsilon=0), %dx%d f32 frames"
                               % (P_PER_GPU, " in total" if strong else "/GPU", ITERS, IMG, IMG), "precision": arms[0][0],
                   "f32_solve": arms[0][1] if arms[0][0] == "f32" else None,
                   "frames": "ping-pong over frames 1..%d of the synthetic sequence; frame 0 initialises" % (N_FRAMES - 1),
                   "l2": "flushed between timed steps (256 MB write)", "threads_per_patch": args.threads or "auto",
                   "occupancy": args.occ if args.threads else "auto",
                   "collective": ("none" if world == 1 else "per frame, fused: the update kernel stores the P x 8 corners into every rank's gathered "
                                  "array over NVLink peer memory (CUDA IPC), one signal / wait kernel; checked against NCCL's all-gather"
                                  if main.get("peer") else (main.get("peer_note") or "NCCL all_gather of the P x 8 corners per frame from the kernel's output buffer"))},
        "e2e": {"value": total_iters / (main["e2e_ms"] * 1e-3), "unit": "iters/s",
                "h2d_bytes_per_step": IMG * IMG * 4, "d2h_bytes_per_step": P * 8 * 8,
                "upload": "frame i + 1 uploads on the library's copy stream while update(i) runs (mtfb_set_image_async, two device "
                          "buffers); every step still moves one frame H2D and the P x 8 corners D2H and waits for them",
                # the same with the frame of step i uploaded in front of update(i) on one stream (mtfb_set_image)
                "serial_upload_value": total_iters / (main["e2e_serial_ms"] * 1e-3)},
        # the same with RAW uint8 frames: upload 1 B / pixel, gray + 5 x 5 Gaussian on the device (mtfb_set_image_u8), update, D2H
        "e2e_raw_u8": {"value": total_iters / (main["raw_ms"] * 1e-3), "unit": "iters/s",
                       "h2d_bytes_per_step": IMG * IMG, "d2h_bytes_per_step": P * 8 * 8, "gpu_launches_per_step": 2},
        "gpu_launches": main["launches"],
        "roofline": roofline(main, arms[0][0]),
        "clocks": clocks,
        "valid": {"finite": bool(np.isfinite(final).all()), "patches_nan": int((status & 1 != 0).sum()),
                  "patches_rank_deficient_H": int((status & 2 != 0).sum()),
                  # after warmup + steps frames of tracking, against the synthetic sequence's ground truth
                  "corner_err_vs_truth_px": main["truth"]},
    }
    if len(res) > 1:
        out["other_arms"] = {}
        for (prec, solve), m in zip(arms[1:], res[1:]):
            d = np.abs(m["final"] - final).max(axis=(1, 2))
            out["other_arms"]["%s_%s_solve" % (prec, solve) if prec == "f32" else prec] = {
                "precision": prec, "f32_solve": solve if prec == "f32" else None, "dtype": dtype(prec),
                "value": total_iters / (m["ms"] * 1e-3), "ms_per_step": m["ms"] / args.steps, "e2e": total_iters / (m["e2e_ms"] * 1e-3),
                "roofline_frac": roofline(m, prec)["frac"], "kernel_ms_per_launch": m["kms"] / args.steps, "gpu_launches": m["launches"],
                "corner_err_vs_truth_px": m["truth"], "patches_rank_deficient_H": int((m["status"] & 2 != 0).sum()),
                # the arms track the same patches through the same frames: how far from the headline arm they end up (px)
                "corner_diff_vs_headline_px": {"median": float(np.median(d)), "p99": float(np.percentile(d, 99)), "max": float(d.max())}}
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_patches = min(P, 4 * cores)
        it, secs = cpu_sample(frames, corners, n_patches, 1, cores)
        reps = 1
        if secs < 5.0:                      # aim for ~10 s of wall clock
            reps = int(min(8, max(1, 5.0 / max(secs, 1e-3))))
            it, secs = cpu_sample(frames, corners, n_patches, min(reps, N_FRAMES - 1), cores)
        it2, secs2 = cpu_sample(frames, corners, n_patches, min(reps, N_FRAMES - 1), cores, fast_sums=2)
        by_kernel = {"dot_products_full_SxS": it / secs, "blocked_syrk_upper_triangle": it2 / secs2}
        out["cpu_baseline"] = {"value": max(by_kernel.values()), "unit": "iters/s", "cores": cores, "kind": "port",
                               "sample": "%d patches, %d LK iterations, OpenMP over patches" % (n_patches, it),
                               # the oracle's J^T J two ways; Eigen's vectorised general product lies between them
                               "by_hessian_kernel": by_kernel}
    return out


def run_ours(args):
    import benchlib
    env = benchlib.Env(args)
    sampler = ClockSampler(env.local_rank); sampler.start()        # nvidia-smi needs ~1 s to produce its first sample
    strong = args.scaling == "strong"
    cpu = not args.no_cpu
    few = max(3, min(args.steps, 10))
    if args.config == 2:
        out = config2_line(args, env, sampler, strong)
        if not args.no_others and not args.one_arm:
            # the other BASELINE configurations, fewer steps each; a failure of one must not take the headline down
            others = {}
            for name, fn in (("config3_res25", lambda: benchlib.run_config3(env, ROOT, 25, few, 3, strong, cpu)),
                             ("config3_res10", lambda: benchlib.run_config3(env, ROOT, 10, few, 3, strong, cpu)),
                             ("config3_res25_f64", lambda: benchlib.run_config3(env, ROOT, 25, few, 3, strong, False, precision="f64")),
                             ("config4", lambda: benchlib.run_config4(env, ROOT, few, 3, cpu=cpu)),
                             ("config5", lambda: benchlib.run_config5(env, ROOT, few, 3, cpu=cpu))):
                try:
                    others[name] = fn()
                except Exception as e:          # reported, not hidden
                    others[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            out["other_configs"] = others
    elif args.config == 3:
        out = benchlib.run_config3(env, ROOT, args.res or 25, args.steps, args.warmup, strong, cpu, precision=args.precision)
    elif args.config == 4:
        out = benchlib.run_config4(env, ROOT, args.steps, args.warmup, cpu=cpu)
    else:
        out = benchlib.run_config5(env, ROOT, args.steps, args.warmup, precision=args.precision, cpu=cpu)
    if args.config != 2:
        out["clocks"] = sampler.summary([(0, time.time() + 1)])
    if env.rank == 0:
        print(json.dumps(out))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--threads", type=int, default=0, help="threads per patch (0 = library default)")
    ap.add_argument("--occ", type=int, default=0, help="occupancy knob of the update kernel (0, 1, 2)")
    ap.add_argument("--precision", default="f32", choices=["f64", "f32"],
                    help="per-pixel arithmetic of the headline arm (include/mtf_b200.h MTFB_PRECISION_*); the other one is "
                         "timed too and reported under other_precision")
    ap.add_argument("--f32-solve", default="reference", choices=["reference", "local"],
                    help="F32 arm: the reference's column-pivoted QR in the reference's basis (default, what parity is claimed "
                         "for) or the Gauss-Jordan solve in the patch-local basis (include/mtf_b200.h MTFB_F32_SOLVE_*)")
    ap.add_argument("--one-arm", action="store_true", help="time only the --precision / --f32-solve arm (and no other configuration)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (benchlib.py)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="configs 2 / 3 at N > 1: patches per GPU fixed (weak) or 1024 patches in total split over the ranks "
                         "(strong); configs 4 and 5 are fixed-size batches, always strong")
    ap.add_argument("--res", type=int, default=0, help="config 3: cell resolution (25 = modules.cfg, 10 = parameters.h default)")
    ap.add_argument("--reference-brief", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-others", action="store_true", help="config 2: do not append the other configurations' lines")
    ap.add_argument("--pitch-pad", type=int, default=0, help="experiment: extra floats per device frame row")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1, config 2: the per-frame exchange of the corners -- the library's own stores into peer memory over "
                         "NVLink (mtfb_peer_*) or NCCL's all-gather of the result buffer")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
