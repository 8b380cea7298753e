"""benchlib.py -- the BASELINE.json configurations 3, 4 and 5 as bench lines (bench.py holds the headline, config 2, and the
command line).  Every runner returns one dict with the same keys as the headline line: metric / value / unit / ms_per_step /
scaling / e2e / roofline / gpu_launches / config (+ cpu_baseline on rank 0 at N = 1).

  config 3  ESM + NCC + Affine, GridTracker 32 x 32 cells, every cell re-initialised on every frame
            (grid_reset_at_each_frame = 1, SM/src/GridTracker.cc:265-285): step = setImage + update + initialize
  config 4  ICLK + MI + Homography, 100 x 100, 8192 patches on 2048 x 2048 frames SHARDED over the ranks (strong scaling by
            definition: BASELINE.json "8192-patch batch sharded across 2/4/8 B200")
  config 5  PF + SSD + Homography, 64 objects x 10 000 particles, objects sharded over the ranks: step = one frame of the
            particle filter (perturb -> evaluate -> weights -> resample -> mean state, SM/src/NT/PF.cc:236-446)

Timing rules as for the headline: >= 3 warm-up steps, L2 flushed between timed steps, CUDA events on the library's stream,
max over ranks; e2e = the same step from pinned HOST buffers with the result read back every step.
Algorithmic bytes per unit: SURVEY.md 8(d).
"""
import os
import time

import numpy as np


def peaks(root):
    import json
    try:
        with open(os.path.join(root, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class Env:
    """one rank's device, stream, process group and L2-flush buffer"""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)          # > 126 MB L2
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(env, n_warm, n_steps, step, kernel=None):
    """warm-up, then n_steps timed steps with the L2 flushed in between -> (total ms, kernel ms or None, wall window).
    step(i) enqueues one step on env.stream; kernel(i), if given, is the part of it timed separately (the update launch):
    step(i) must then be written as pre(i); kernel(i); post(i) by the caller through the returned events -- here the caller
    passes step = (pre, post) tuple instead"""
    torch = env.torch
    pre, post = step if isinstance(step, tuple) else (None, None)
    for i in range(n_warm):
        if pre is None:
            step(i)
        else:
            pre(i); kernel(i); post(i)
    env.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
    env.barrier()
    w0 = time.time()
    for i in range(n_steps):
        env.flush.fill_(i & 0xff)
        ev[i][0].record(env.stream)
        if pre is None:
            step(n_warm + i)
        else:
            pre(n_warm + i)
            kev[i][0].record(env.stream)
            kernel(n_warm + i)
            kev[i][1].record(env.stream)
            post(n_warm + i)
        ev[i][1].record(env.stream)
    env.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kms = sum(a.elapsed_time(b) for a, b in kev) if pre is not None else None
    return ms, kms, (w0, time.time())


def timed_e2e(env, n_warm, n_steps, step, prime=None, drain=None):
    """the same step from host buffers, the caller reading the result every step (stream.synchronize inside step).
    prime(): hands the first frame over before the loop (steps prefetch the NEXT frame: mtfb_set_image_async); drain():
    consumes the frame the last step prefetched"""
    torch = env.torch
    if prime is not None:
        prime()
    for i in range(n_warm):
        step(i)
    env.barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    env.barrier()
    w0 = time.time()
    wall0 = time.perf_counter()
    t0.record(env.stream)
    for i in range(n_steps):
        step(n_warm + i)
    t1.record(env.stream)
    env.barrier()
    out = max(t0.elapsed_time(t1), 1e3 * (time.perf_counter() - wall0)), (w0, time.time())
    if drain is not None:
        drain()
        env.barrier()
    return out


def roofline(alg_bytes_per_unit, units_per_launch, kernel_ms_per_launch, root, kernel, note, traffic=None, traffic_source=None):
    peak, peak_src = peaks(root)
    achieved = alg_bytes_per_unit * units_per_launch / (kernel_ms_per_launch * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_source, "kernel": kernel, "kernel_ms_per_launch": kernel_ms_per_launch,
            "alg_bytes_per_unit": alg_bytes_per_unit, "alg_bytes_per_launch": alg_bytes_per_unit * units_per_launch,
            "peak_source": peak_src, "note": note}


def _shard(env, n_total, strong):
    from mtf_b200 import sharding
    if strong:
        lo, hi = sharding.shard_range(n_total, env.world, env.rank)
        return lo, hi, n_total
    return 0, n_total, n_total * env.world


def _gather(env, sharded, strong):
    """the per-frame exchange: all ranks end up with every patch's corners, read from the kernel's own
    output buffer on the device (ShardedBatchTracker.getRegion -> NCCL all-gather on the current stream)"""
    if env.world > 1:
        return sharded.getRegion(device=env.dev)
    return None


# ------------------------------------------------------------------------------------------------ config 3
def run_config3(env, root, res, n_steps, n_warm, strong=False, cpu=True, precision="f32"):
    """GridTracker<Homography>::update (SM/src/GridTracker.cc:247-285) with 32 x 32 ESM + NCC + Affine cells: the cells' update,
    the warp of the region from their centroids (RANSAC + LM refinement on the device), the cells re-initialised on the frame
    at the regions the new corners give them.  weak scaling: one such grid per GPU (the regions all-gathered per frame);
    strong: the cells of one grid split over the GPUs (their corners all-gathered, every rank estimates)."""
    from mtf_b200 import api, grid, sharding, workloads as W
    torch = env.torch
    frames, _ = W.sequence()
    order = W.frame_order()
    G = W.CONFIG3["grid"]
    n_cells = G * G
    N = res * res
    iters = 30
    split = strong and env.world > 1
    lo, hi, n_job = _shard(env, n_cells, strong)
    P = hi - lo
    size = frames[0].shape[0]
    region = np.array([[24.0, size - 25.0, size - 25.0, 24.0], [24.0, 24.0, size - 25.0, size - 25.0]])

    def make_local(n):
        tr = api.BatchTracker(api.make_params("ncc", "affine", "esm", n_patches=n, resx=res, resy=res, max_iters=iters, epsilon=0.0,
                                              device=env.local_rank, hess_type=W.CONFIG3["hess_type"], jac_type=W.CONFIG3["jac_type"],
                                              precision=precision))
        tr.set_stream(env.stream.cuda_stream)
        return tr

    common = dict(grid_size_x=G, grid_size_y=G, patch_size_x=res, patch_size_y=res, reset_at_each_frame=1, ssm="homography",
                  est_params=api.make_est_params("ransac"), seed=7)
    if split:
        sh = sharding.ShardedBatchTracker(n_cells, make_local)
        tr = sh.local
        gt = grid.GridTracker(None, cells=tr, shard=(sh.lo, sh.hi), gather=lambda: sh.getRegion(device=env.dev),
                              upload=lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev), **common)
    else:
        tr = make_local(n_cells)
        gt = grid.GridTracker(None, cells=tr, **common)
    d_frames = [torch.from_numpy(f).to(env.dev) for f in frames]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    h, w = frames[0].shape
    regions = torch.empty((env.world, 8), dtype=torch.float64, device=env.dev) if (env.world > 1 and not split) else None
    gt.setImage(d_frames[0]); gt.initialize(region); tr.synchronize()
    launches0 = tr.launch_count

    def finish():
        gt.finish_update()
        if regions is not None:          # every rank learns every grid's region
            sharding.all_gather_rows(torch.from_numpy(gt.corners.reshape(1, 8)).to(env.dev), env.world, None, out=regions)

    def pre(i):
        gt.setImage(d_frames[order[i % len(order)]])

    def kernel(i):
        tr.update()

    def post(i):
        finish()
    ms, kms, win = timed_steps(env, n_warm, n_steps, (pre, post), kernel)
    launches = tr.launch_count - launches0
    pre(0); kernel(0); n_it = tr.n_iters(); post(0)          # (the reset zeroes the cells' iteration counts: read them in between)
    est = gt.last_estimate
    finite = bool(np.isfinite(gt.corners).all())
    drift = float(np.abs(gt.corners - region).max())

    def e2e_step(i):
        tr.update()                                                                  # frame i, prefetched during step i - 1
        finish()                                                                     # reads the estimate back: the step's result
        tr.prefetch_image_pinned(pinned[order[(i + 1) % len(order)]].data_ptr(), h, w, w)
        env.stream.synchronize()
    gt.setImage(frames[0]); gt.initialize(region)
    e2e_ms, win2 = timed_e2e(env, n_warm, n_steps, e2e_step, prime=lambda: tr.prefetch_image_pinned(pinned[order[0]].data_ptr(), h, w, w),
                             drain=tr.update)
    ms, kms, e2e_ms = env.max_over_ranks([ms, kms, e2e_ms])
    total_iters = n_job * iters * n_steps
    alg = 16 * N + 432
    out = {
        "metric": "LK iters/sec (%dx%d NCC+Affine ESM, GridTracker 32x32 cells)" % (res, res), "value": total_iters / (ms * 1e-3),
        "unit": "iters/s", "n_gpus": env.world, "steps": n_steps, "warmup": n_warm, "ms_per_step": ms / n_steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64" if precision == "f64" else "f32 per pixel (bit-exact sampling indices), f64 reduction + solve", "data": "synthetic",
        "config": {"workload": "GridTracker<Homography>::update: ESM+NCC+Affine cells, 32x32 = %d cells%s of %dx%d px, %d iters/frame "
                               "(epsilon=0), region warp by RANSAC + LM from the cell centroids on the device, every cell "
                               "re-initialised on every frame where the new region puts it, 1024x1024 f32 frames"
                               % (n_cells, "" if strong else " per GPU", res, res, iters),
                   "precision": precision, "l2": "flushed between timed steps (256 MB write)",
                   "collective": ("none" if env.world == 1 else "all_gather of the cells' P x 8 corners per frame from the kernel's output "
                                  "buffer, every rank estimates" if split else "all_gather of the grids' regions per frame")},
        "e2e": {"value": total_iters / (e2e_ms * 1e-3), "unit": "iters/s", "h2d_bytes_per_step": h * w * 4 + P * 64,
                "d2h_bytes_per_step": 21 * 8 + n_cells},
        "gpu_launches": int(launches),
        "roofline": roofline(alg, P * iters, kms / n_steps, root,
                             "ncc_update_kernel<Affine,ESM>" if precision == "f64" else "ncc_update_f32_kernel<Affine,ESM>",
                             "fp64-issue bound (three sweeps per pass over It kept in shared memory), not HBM bound" if precision == "f64"
                             else "one fp32 sweep per pass; bound by the per-pass serial tail (6 x 6 QR, update), not HBM",
                             traffic=(12614656 if (precision != "f64" and res == 25 and P == 1024) else None),
                             traffic_source=("profiles/r02_ncu_ncc_f32_summary.txt (dram__bytes_read + write, one launch)"
                                             if (precision != "f64" and res == 25 and P == 1024) else None)),
        "stages_ms_per_step": {"cells_update_kernel": kms / n_steps, "estimate_and_reset": (ms - kms) / n_steps},
        "valid": {"finite": finite, "n_iters_per_cell": int(n_it[0]), "estimate_ok": bool(est["ok"]), "inliers": int(est["n_inliers"]),
                  "hypotheses": int(est["drawn"]), "region_drift_px": drift},
    }
    if cpu and env.rank == 0 and env.world == 1:
        from oracle import oracle_lib as O
        cores = os.cpu_count() or 1
        n = min(n_cells, 8 * cores)
        cells_all = gt.cell_corners()
        prm = O.make_params("ncc", "affine", "esm", resx=res, resy=res, max_iters=iters, epsilon=0.0, grad_mode=0, fast_sums=1,
                            hess_type=W.CONFIG3["hess_type"], jac_type=W.CONFIG3["jac_type"])
        it, secs, _, _ = O.batch_track(prm, frames[:3], cells_all[:n], n_threads=cores, reset_each_frame=True)
        # + the estimation on all the centroids, once per frame (single thread, as the reference runs it)
        prev, curr = tr.grid_pts()
        t0 = time.perf_counter()
        for k in range(10):
            O.estimate_warp("homography", prev, curr, O.make_est_params("ransac", seed=7 + k))
        t_est = (time.perf_counter() - t0) / 10
        per_frame = secs / 2 * (n_cells / n) + t_est
        out["cpu_baseline"] = {"value": n_cells * iters / per_frame, "unit": "iters/s", "cores": cores, "kind": "port",
                               "estimate_ms_per_frame": t_est * 1e3,
                               "sample": "%d of the %d cells x 2 frames (%d LK iterations, re-initialised every frame, OpenMP over cells) scaled "
                                         "to the grid, + estimateWarpFromPts (RANSAC + LM, one thread) on the %d centroids" % (n, n_cells, it, n_cells)}
    tr.close()
    return out


# ------------------------------------------------------------------------------------------------ config 4
def run_config4(env, root, n_steps, n_warm, n_patches=8192, cpu=True):
    from mtf_b200 import api, sharding, workloads as W
    torch = env.torch
    size = W.CONFIG4["size"]
    frames, _ = W.sequence(4, size)
    order = [1, 2, 3, 2]
    corners_all = W.config4_patches(n_patches, size)
    res, iters = W.CONFIG4["res"], 30
    N = res * res
    lo, hi = sharding.shard_range(n_patches, env.world, env.rank)
    P = hi - lo

    def make_local(n):
        tr = api.BatchTracker(api.make_params("mi", "homography", "iclk", n_patches=n, resx=res, resy=res, max_iters=iters, epsilon=0.0,
                                              device=env.local_rank, hess_type=W.CONFIG4["hess_type"], mi_n_bins=W.CONFIG4["mi_n_bins"],
                                              mi_pre_seed=W.CONFIG4["mi_pre_seed"], mi_pou=W.CONFIG4["mi_pou"]))
        tr.set_stream(env.stream.cuda_stream)
        return tr
    if env.world > 1:
        sh = sharding.ShardedBatchTracker(n_patches, make_local)
        tr = sh.local
    else:
        sh, tr = None, make_local(P)
    d_frames = [torch.from_numpy(f).to(env.dev) for f in frames]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    d_corners = sharding.device_view(tr.device_results()[0], (P, 8), env.dev)
    host_out = torch.empty((P, 8), dtype=torch.float64).pin_memory()
    tr.initialize(corners_all[lo:hi], d_frames[0]); tr.synchronize()
    launches0 = tr.launch_count

    def pre(i):
        tr.setImage(d_frames[order[i % len(order)]])

    def kernel(i):
        tr.update()

    def post(i):
        if env.world > 1:
            sh.getRegion(device=env.dev)
    ms, kms, win = timed_steps(env, n_warm, n_steps, (pre, post), kernel)
    launches = tr.launch_count - launches0
    final = tr.getRegion()
    status = tr.patch_status()

    def e2e_step(i):
        tr.update()
        tr.prefetch_image_pinned(pinned[order[(i + 1) % len(order)]].data_ptr(), size, size, size)
        if env.world > 1:
            sh.getRegion(device=env.dev)
        host_out.copy_(d_corners, non_blocking=True)
        env.stream.synchronize()
    tr.initialize(corners_all[lo:hi], frames[0])
    e2e_ms, _ = timed_e2e(env, n_warm, n_steps, e2e_step, prime=lambda: tr.prefetch_image_pinned(pinned[order[0]].data_ptr(), size, size, size),
                          drain=tr.update)
    ms, kms, e2e_ms = env.max_over_ranks([ms, kms, e2e_ms])
    total_iters = n_patches * iters * n_steps
    alg = 16 * N + 152
    out = {
        "metric": "LK iters/sec (100x100 MI+Homography ICLK)", "value": total_iters / (ms * 1e-3), "unit": "iters/s", "n_gpus": env.world,
        "steps": n_steps, "warmup": n_warm, "ms_per_step": ms / n_steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ICLK+MI+Homography, %d patches 100x100 in total, sharded over the ranks (%d on this one), %d iters/frame "
                               "(epsilon=0), n_bins=8, pre_seed=10, %dx%d f32 frames" % (n_patches, P, iters, size, size),
                   "l2": "flushed between timed steps (256 MB write)",
                   "collective": "all_gather of P x 8 corners per frame from the kernel's output buffer" if env.world > 1 else "none"},
        "e2e": {"value": total_iters / (e2e_ms * 1e-3), "unit": "iters/s", "h2d_bytes_per_step": size * size * 4, "d2h_bytes_per_step": P * 64},
        "gpu_launches": int(launches),
        "roofline": roofline(alg, P * iters, kms / n_steps, root, "mi_update_kernel<Homography,ICLK>",
                             "fp64 issue + shared-memory histogram updates, not HBM bound",
                             # one launch of the whole 8192-patch batch (fp64 template, its gradient and the pass's pixel values
                             # stream from DRAM on every pass): scaled to this rank's share
                             traffic=int((95087834000 + 19510694000) * P / 8192),
                             traffic_source="profiles/r02_ncu_mi_final_summary.txt (dram__bytes_read + write, one launch of 8192 patches)"),
        "valid": {"finite": bool(np.isfinite(final).all()), "patches_nan": int((status & 1 != 0).sum())},
    }
    if cpu and env.rank == 0 and env.world == 1:
        from oracle import oracle_lib as O
        cores = os.cpu_count() or 1
        n = min(n_patches, 2 * cores)
        prm = O.make_params("mi", "homography", "iclk", resx=res, resy=res, max_iters=iters, epsilon=0.0, grad_mode=0, fast_sums=1,
                            hess_type=W.CONFIG4["hess_type"], mi_n_bins=W.CONFIG4["mi_n_bins"], mi_pre_seed=W.CONFIG4["mi_pre_seed"],
                            mi_pou=W.CONFIG4["mi_pou"])
        it, secs, _, _ = O.batch_track(prm, frames[:2], corners_all[:n], n_threads=cores)
        out["cpu_baseline"] = {"value": it / secs, "unit": "iters/s", "cores": cores, "kind": "port",
                               "sample": "%d patches x 1 frame, %d LK iterations, OpenMP over patches" % (n, it)}
    tr.close()
    return out


# ------------------------------------------------------------------------------------------------ config 5
def run_config5(env, root, n_steps, n_warm, n_objects=64, n_particles=10000, precision="f32", cpu=True):
    from mtf_b200 import api, sharding, workloads as W
    torch = env.torch
    frames, _ = W.sequence()
    order = W.frame_order()
    objs_all = W.config5_objects(n_objects)
    lo, hi = sharding.shard_range(n_objects, env.world, env.rank)
    P = hi - lo
    res = W.CONFIG5["res"]
    N = res * res
    tr = api.PFTracker(api.make_params("ssd", "homography", "pf", n_patches=P, resx=res, resy=res, device=env.local_rank,
                                       precision=precision), n_particles=n_particles, sigma=W.PF_SIGMA_HOM, seed=7 + env.rank)
    tr.set_stream(env.stream.cuda_stream)
    d_frames = [torch.from_numpy(f).to(env.dev) for f in frames]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    h, w = frames[0].shape
    d_corners = sharding.device_view(tr.device_results()[0], (P, 8), env.dev)
    gathered = torch.empty((n_objects, 8), dtype=torch.float64, device=env.dev) if env.world > 1 else None
    host_out = torch.empty((P, 8), dtype=torch.float64).pin_memory()
    tr.initialize(objs_all[lo:hi], d_frames[0]); tr.synchronize()
    launches0 = tr.launch_count

    def pre(i):
        tr.setImage(d_frames[order[i % len(order)]])

    def kernel(i):
        tr.update()

    def post(i):
        if env.world > 1:
            sharding.all_gather_rows(d_corners, n_objects, out=gathered)
    ms, kms, win = timed_steps(env, n_warm, n_steps, (pre, post), kernel)
    launches = tr.launch_count - launches0
    final = tr.getRegion()

    def e2e_step(i):
        tr.update()
        tr.prefetch_image_pinned(pinned[order[(i + 1) % len(order)]].data_ptr(), h, w, w)
        if env.world > 1:
            sharding.all_gather_rows(d_corners, n_objects, out=gathered)
        host_out.copy_(d_corners, non_blocking=True)
        env.stream.synchronize()
    tr.initialize(objs_all[lo:hi], frames[0])
    e2e_ms, _ = timed_e2e(env, n_warm, n_steps, e2e_step, prime=lambda: tr.prefetch_image_pinned(pinned[order[0]].data_ptr(), h, w, w),
                          drain=tr.update)
    ms, kms, e2e_ms = env.max_over_ranks([ms, kms, e2e_ms])
    total = n_objects * n_particles * n_steps
    out = {
        "metric": "PF particle evaluations/sec (50x50 SSD+Homography)", "value": total / (ms * 1e-3), "unit": "particles/s", "n_gpus": env.world,
        "steps": n_steps, "warmup": n_warm, "ms_per_step": ms / n_steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 per pixel (bit-exact sampling indices), f64 weights" if precision == "f32" else "f64", "data": "synthetic",
        "config": {"workload": "PF+SSD+Homography, %d objects x %d particles in total, objects sharded over the ranks (%d on this one), one "
                               "particle-filter frame per step (perturb, evaluate, weights, resample, mean), 1024x1024 f32 frames"
                               % (n_objects, n_particles, P), "precision": precision,
                   "l2": "flushed between timed steps (256 MB write)",
                   "collective": "all_gather of the objects' corners per frame" if env.world > 1 else "none"},
        "e2e": {"value": total / (e2e_ms * 1e-3), "unit": "particles/s", "h2d_bytes_per_step": h * w * 4, "d2h_bytes_per_step": P * 64},
        "gpu_launches": int(launches),
        # 72 B in + 8 B out per particle; the template and the frame tile are per object and stay on chip
        "roofline": roofline(80, P * n_particles, kms / n_steps, root, "pf_evaluate_f32_kernel<Homography> (+ perturb / weights / resample)",
                             "not HBM bound by construction (80 B per particle against 30 N flops): fp32 issue / latency bound"),
        "valid": {"finite": bool(np.isfinite(final).all())},
    }
    if cpu and env.rank == 0 and env.world == 1:
        from oracle import oracle_lib as O
        cores = os.cpu_count() or 1
        n_obj, n_part = min(n_objects, cores), 1000
        prm = O.make_params("ssd", "homography", "fclk", resx=res, resy=res)
        n, secs, _ = O.batch_pf_evaluate(prm, frames[0], frames[1], objs_all[:n_obj], W.config5_states(n_obj, n_part, seed=1), n_threads=cores)
        out["cpu_baseline"] = {"value": n / secs, "unit": "particles/s", "cores": cores, "kind": "port",
                               "sample": "%d objects x %d particles, particle evaluation loop only (NT/PF.cc:303-320), OpenMP" % (n_obj, n_part)}
    tr.close()
    return out
