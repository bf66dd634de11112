#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native optical-flow hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "frame-pairs/sec & Mpix/s ... (classic+nl-fast)", north_star target config):
`classic+nl-fast`, fp64, synthetic 640x480 RGB pairs with a known affine flow (generator below, seeds 3, 4, ...),
B pairs per GPU per step.  A step = one pass of the whole hot path (colour conversion, ROF texture, pyramids, 21
warp / assemble / PCG-solve / occlusion / weighted-median iterations) over one batch.  Frame pairs are independent,
so ranks share nothing: weak scaling, no data-path collective (torch.distributed is used for the barrier and the
max-over-ranks of the timings only).

One JSON line on rank 0:
  value      frame-pairs/s, whole job, inputs already resident in HBM (b200flow_estimate_rgb8_dev), CUDA-event timed
  e2e        same metric through the public API estimate_flow_batch with HOST (pinned) uint8 frames in and host
             float64 flow out, H2D/D2H inside the timed region
  roofline   the PCG solver kernel: algorithmic bytes (120 B per pixel-iteration mixed / 228 B fp64) / CUDA-event
             time of the solves inside the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the NumPy/SciPy oracle port (oracle/flow_oracle.py) of the same preset timed on a bounded sample
--impl reference: the CPU arm -- the oracle port on all host cores (one process per core, one sample pair each).
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
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200"))

METHOD = "classic+nl-fast"
H, W = 480, 640
# algorithmic bytes per pixel per PCG iteration (DESIGN.md section 4 / solve.cu headers):
#   mixed (default): fp32 Krylov vectors + fp32 coefficient copy, phase A 76 B + phase B 44 B
#   fp64           : every vector fp64, phase A 152 B + phase B 76 B
PCG_BYTES = {"mixed": 128, "mixed-jacobi": 120, "fp64": 228}
PCG_KERNEL = {"mixed": "pcg_ic_kernel", "mixed-jacobi": "pcg_mixed_kernel", "fp64": "pcg_kernel"}
PCG_PRECOND = {"mixed": "tile-local block-IC(0)", "mixed-jacobi": "block-Jacobi", "fp64": "block-Jacobi"}
SAMPLE_H, SAMPLE_W = 120, 160         # CPU-baseline sample: centre crop with 1/16 of the pixels


# --------------------------------------------------------------------------------------------------
# synthetic data (SURVEY.md section 8d generator, extended to three colour channels)
# --------------------------------------------------------------------------------------------------
def synth_pair(h, w, seed):
    """RGB uint8 pair with im2(p + uv(p)) = im1(p) for a known affine flow (max |flow| about 2.4 px)."""
    from scipy.ndimage import gaussian_filter, map_coordinates
    rng = np.random.default_rng(seed)
    pad = 64
    base = gaussian_filter(rng.random((h + 2 * pad, w + 2 * pad, 3)), (1.5, 1.5, 0))
    base = (base - base.min()) / (base.max() - base.min()) * 255.0
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    a = 0.004 * 256 / max(h, w)
    b = 0.003 * 256 / max(h, w)
    t = np.array([1.5, -0.8])
    M = np.array([[1 + a, -b], [b, 1 + a]])            # p + uv(p) = c + M (p - c) + t
    Mi = np.linalg.inv(M)
    yy, xx = np.mgrid[0:h, 0:w].astype(float)
    qx, qy = xx - cx - t[0], yy - cy - t[1]
    px = cx + Mi[0, 0] * qx + Mi[0, 1] * qy
    py = cy + Mi[1, 0] * qx + Mi[1, 1] * qy
    im1 = base[pad:pad + h, pad:pad + w]
    im2 = np.stack([map_coordinates(base[:, :, c], [py + pad, px + pad], order=3, mode="nearest") for c in range(3)], axis=2)
    u = t[0] + a * (xx - cx) - b * (yy - cy)
    v = t[1] + b * (xx - cx) + a * (yy - cy)
    q = lambda im: np.clip(np.floor(im + 0.5), 0, 255).astype(np.uint8)  # noqa: E731
    return q(im1), q(im2), np.stack([u, v], axis=2)


def make_batch(B, seed0):
    ims = [synth_pair(H, W, seed0 + k) for k in range(B)]
    return (np.stack([i[0] for i in ims]), np.stack([i[1] for i in ims]), np.stack([i[2] for i in ims]))


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port
# --------------------------------------------------------------------------------------------------
def _oracle_sample(seed):
    """One bounded sample of the workload on one core: classic+nl-fast on the SAMPLE_H x SAMPLE_W centre crop."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import flow_oracle as fo
    im1, im2, _ = synth_pair(H, W, seed)
    y0, x0 = (H - SAMPLE_H) // 2, (W - SAMPLE_W) // 2
    c1 = im1[y0:y0 + SAMPLE_H, x0:x0 + SAMPLE_W].astype(float)
    c2 = im2[y0:y0 + SAMPLE_H, x0:x0 + SAMPLE_W].astype(float)
    t0 = time.perf_counter()
    uv = fo.estimate_flow(c1, c2, METHOD)
    return time.perf_counter() - t0, float(np.abs(uv).max())


SAMPLE_DESC = ("oracle port (NumPy/SciPy restatement of the reference, default solver='backslash' = SuperLU) on the "
               "%dx%d centre crop of synthetic pair seed 3 = 1/16 of a 640x480 pair; value = (1/16 pair) / seconds, "
               "which favours the CPU (its direct solve is superlinear in the pixel count)" % (SAMPLE_W, SAMPLE_H))


def cpu_baseline_single():
    dt, _ = _oracle_sample(3)
    frac = (SAMPLE_H * SAMPLE_W) / float(H * W)
    return {"value": frac / dt, "unit": "frame-pairs/s", "cores": 1, "kind": "port", "sample": SAMPLE_DESC,
            "sample_seconds": dt}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference itself cannot
    travel to the GPU box) on all host cores: one process per core, each step = one bounded sample per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    frac = (SAMPLE_H * SAMPLE_W) / float(H * W)
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):
            pool.map(_oracle_sample, [3 + k for k in range(cores)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_oracle_sample, [3 + k for k in range(cores)])
        dt = time.perf_counter() - t0
    value = args.steps * cores * frac / dt
    line = {
        "impl": "reference", "metric": "frame_pairs_per_sec", "value": value, "unit": "frame-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mpix_per_s": value * H * W / 1e6,
        "config": {"workload": "classic+nl-fast, synthetic 640x480 RGB pairs (seed 3..), fp64", "method": METHOD,
                   "height": H, "width": W, "sample": "%dx%d centre crop per core per step" % (SAMPLE_W, SAMPLE_H)},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": "port",
                         "sample": SAMPLE_DESC + "; %d processes, one sample each per step" % cores},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
_stdout_fd = None


def _emit(line):
    """rank 0's ONE JSON line, on the real stdout"""
    sys.stdout.flush()
    if _stdout_fd is not None:
        os.dup2(_stdout_fd, 1)
    print(json.dumps(line), flush=True)


def main():
    global _stdout_fd
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="frame pairs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solver-precision", default="mixed", choices=["mixed", "mixed-jacobi", "fp64"],
                    help="mixed: fp32 Krylov vectors with fp64 reliable updates, tile-local block-IC(0) preconditioner "
                         "(default); mixed-jacobi: same with the block-Jacobi preconditioner; fp64: all-fp64 PCG (variants)")
    ap.add_argument("--split", type=int, default=None,
                    help="concurrent sub-batches per GPU (b200flow_ctx_set_split); default: the library's default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from optical_flow import _lib, estimate_flow_batch, load_of_method

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the flow path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ["B200FLOW_DEVICE"] = str(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on C-level stdout at the first collective: park fd 1 on stderr until the JSON line
        sys.stdout.flush()
        _stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W_ = max(3, args.warmup)
    B = args.batch

    ims1, ims2, flow_gt = make_batch(B, 3 + rank * B)            # every rank its own pairs (weak scaling)
    ctx = _lib.default_context(local_rank)
    if args.split is not None:
        ctx.set_split(args.split)
    PCG_BYTES_PER_PIXEL_ITER = PCG_BYTES[args.solver_precision]
    ope = load_of_method(METHOD)
    ope.solver_precision = args.solver_precision
    ope.pyramid_levels = ope._auto_pyramid_levels(np.empty((H, W, 2)))
    P = ope._c_params(levels=ope.pyramid_levels)
    ope._apply_solver(P)

    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    d1 = torch.from_numpy(ims1).cuda()
    d2 = torch.from_numpy(ims2).cuda()
    duv = torch.empty((B, H, W, 2), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    st = _lib.Stats()

    def step_resident():
        ctx.call("b200flow_estimate_rgb8_dev", P, B, H, W, d1.data_ptr(), d2.data_ptr(), 1, duv.data_ptr(),
                 _lib.C.byref(st))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, CUDA events on the library's stream ----
    ctx.set_timing(True)
    for _ in range(W_):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solver_ms = pixel_iters = launches = 0
    stage_ms = {"solver_ms": 0.0, "warp_ms": 0.0, "filter_ms": 0.0, "pre_ms": 0.0}
    pcg_iters = 0
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            step_resident()
            solver_ms += st.solver_ms
            pixel_iters += st.pcg_pixel_iters
            launches += st.kernel_launches
            pcg_iters += st.pcg_iters
            for k in stage_ms:
                stage_ms[k] += getattr(st, k)
        e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1)
    not_conv = st.not_converged
    ctx.set_timing(False)

    # ---- e2e: public API, pinned host buffers in, host flow out ----
    h1 = torch.from_numpy(ims1).pin_memory().numpy()
    h2 = torch.from_numpy(ims2).pin_memory().numpy()
    for _ in range(2):
        uv_host = estimate_flow_batch(h1, h2, METHOD, params={"solver_precision": args.solver_precision}, device=local_rank)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        uv_host = estimate_flow_batch(h1, h2, METHOD, params={"solver_precision": args.solver_precision}, device=local_rank)
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()

    # accuracy of the timed workload against the known synthetic flow (sanity: the work is real)
    epe = float(np.sqrt(((uv_host - flow_gt) ** 2).sum(-1))[:, 8:-8, 8:-8].mean())

    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e, solver_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_resident, ms_e2e, solver_ms_max = [float(v) for v in t.tolist()]
        s = torch.tensor([float(pixel_iters), float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        pixel_iters_all, launches_all = [float(v) for v in s.tolist()]
    else:
        solver_ms_max, pixel_iters_all, launches_all = solver_ms, float(pixel_iters), float(launches)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        pairs = B * world * args.steps
        value = pairs / (ms_resident / 1e3)
        e2e_value = pairs / (ms_e2e / 1e3)
        # roofline of the solver kernel: per-rank algorithmic bytes / per-rank solver time
        achieved = (pixel_iters * PCG_BYTES_PER_PIXEL_ITER) / (solver_ms / 1e3) / 1e9 if solver_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, like `achieved`'s
        # numerator); only quoted for the kernel it was captured on
        traffic = traffic_detail = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "pcg_traffic.json")))
            if tj.get("kernel") == PCG_KERNEL[args.solver_precision]:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_detail = {k: tj.get(k) for k in ("capture", "algorithmic_bytes_this_launch",
                                                         "dram_bytes_per_pixel_iter", "algorithmic_bytes_per_pixel_iter")}
        except (OSError, ValueError):
            pass
        line = {
            "metric": "frame_pairs_per_sec", "value": value, "unit": "frame-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": W_, "ms_per_step": ms_resident / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mpix_per_s": value * H * W / 1e6,
            "config": {"workload": "classic+nl-fast, synthetic 640x480 RGB pairs (affine flow, seeds 3..), fp64, "
                                   "%d pairs per GPU per step" % B,
                       "method": METHOD, "height": H, "width": W, "pairs_per_gpu_per_step": B,
                       "parallelism": "independent frame pairs per GPU, no collective",
                       "solver": "%s PCG until the fp64 true residual ||b-Ax|| <= %g ||b|| (stands in for spsolve); %s"
                                 % (PCG_PRECOND[args.solver_precision], P.tol,
                                    "Krylov vectors fp32, solution + residual replacement fp64" if
                                    args.solver_precision != "fp64" else "all vectors fp64"),
                       "solver_precision": args.solver_precision,
                       "l2": "per-step working set ~%.1f GB per GPU >> 126 MB L2; no flush needed" % (B * H * W * 450 / 1e9)},
            "e2e": {"value": e2e_value, "unit": "frame-pairs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(2 * B * H * W * 3), "d2h_bytes_per_step": int(B * H * W * 2 * 8)},
            "gpu_launches": int(launches_all),
            "roofline": {"kernel": "%s (persistent cooperative PCG, solve%s.cu)" %
                                   (PCG_KERNEL[args.solver_precision], "_ic" if args.solver_precision == "mixed" else ""),
                         "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_detail": traffic_detail,
                         "bytes_per_pixel_iter": PCG_BYTES_PER_PIXEL_ITER, "pixel_iters_per_step": pixel_iters / args.steps,
                         "solver_ms_per_step": solver_ms / args.steps,
                         "solver_share_of_step": solver_ms / ms_resident if ms_resident else None},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "pcg_iters_per_step": pcg_iters / args.steps, "pcg_not_converged": int(not_conv),
            "aepe_vs_known_flow_px": epe,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
