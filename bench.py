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
PCG_BYTES = {"mixed": 128, "mixed-jacobi": 120, "fp64": 228, "fp32": 128}
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
# CPU arm: the unmodified reference (baseline/_ref, installed by `pip install --no-deps --target baseline/_ref`; it is
# git-ignored but travels to the GPU box) when present, else the oracle port.  Both are pure NumPy/SciPy, single threaded.
# --------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "tests", "golden", "_stubs")       # 2-file matplotlib stub: lets the reference's __init__ import


def cpu_kind():
    return "reference" if os.path.isdir(os.path.join(REF_DIR, "optical_flow")) else "port"


def _cpu_estimate(kind):
    """estimate_flow of the CPU implementation (called in a process that never imports the B200 package)."""
    if kind == "reference":
        for q in (REF_DIR, STUBS):
            if q not in sys.path:
                sys.path.insert(0, q)
        import optical_flow
        assert os.path.realpath(optical_flow.__file__).startswith(os.path.realpath(REF_DIR)), optical_flow.__file__
        from optical_flow import estimate_flow as ef
        import contextlib
        import io

        def run(c1, c2, method, params=None):
            with contextlib.redirect_stdout(io.StringIO()):
                return ef(c1, c2, method, params)
        return run
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import flow_oracle as fo
    return lambda c1, c2, method, params=None: fo.estimate_flow(c1, c2, method, params)


def _cpu_sample(job):
    """One bounded sample of the workload on one core: classic+nl-fast on the sh x sw centre crop of synthetic pair `seed`."""
    seed, sh, sw, kind, solver = job
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    ef = _cpu_estimate(kind)
    im1, im2, _ = synth_pair(H, W, seed)
    y0, x0 = (H - sh) // 2, (W - sw) // 2
    c1 = im1[y0:y0 + sh, x0:x0 + sw].astype(float)
    c2 = im2[y0:y0 + sh, x0:x0 + sw].astype(float)
    params = {"solver": "pcg", "pcg_rtol": 1e-8} if solver == "pcg" else None
    t0 = time.perf_counter()
    uv = ef(c1, c2, METHOD, params)
    return time.perf_counter() - t0, float(np.abs(uv).max())


def full_size_measured():
    """profiles/cpu_fullsize.json: wall seconds of the unmodified reference on the FULL 640x480 seed-3 pair, measured once on
    the build host by tests/golden/gen_golden_configs.py (one core; the run also produced the golden flow the GPU is checked
    against).  Quoted beside the live sample so that the crop extrapolation can be judged against a measurement."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "cpu_fullsize.json")))
    except (OSError, ValueError):
        return None
    out = {"host": d.get("_host")}
    for k in ("bench640", "bench640_pcg"):
        if k in d:
            out[k] = {"seconds_per_pair": d[k]["seconds"], "pairs_per_s_per_core": 1.0 / d[k]["seconds"], "params": d[k].get("params")}
    return out


def sample_desc(kind, sh, sw, solver):
    return ("%s, solver=%s, on the %dx%d centre crop of synthetic pair seed 3.. = %.4f of a 640x480 pair; value = (that "
            "fraction of a pair) / seconds per core, which favours the CPU (its work is superlinear in the pixel count: "
            "see full_size_measured)" % ("the unmodified reference from baseline/_ref" if kind == "reference" else
                                          "oracle port (NumPy/SciPy restatement of the reference)",
                                          "'backslash' (SuperLU, the preset's default)" if solver != "pcg" else
                                          "'pcg' with pcg_rtol=1e-8 (the reference's own iterative mode, base.py:104-136)",
                                          sw, sh, sh * sw / float(H * W)))


def cpu_baseline_single():
    """cpu_baseline of the B200 line: ONE 320x240 sample (1/4 pair) on one core, in a child process (the reference and the
    drop-in share the package name); beside it the reference's own solver='pcg', pcg_rtol=1e-8 on the same crop (a second
    process, concurrently) and the measured full-size seconds."""
    import multiprocessing as mp
    kind = cpu_kind()
    with mp.get_context("spawn").Pool(2) as pool:
        r_big = pool.apply_async(_cpu_sample, ((3, 240, 320, kind, "backslash"),))
        r_pcg = pool.apply_async(_cpu_sample, ((3, 240, 320, kind, "pcg"),))
        dt, _ = r_big.get()
        dt_pcg, _ = r_pcg.get()
    frac = (240 * 320) / float(H * W)
    return {"value": frac / dt, "unit": "frame-pairs/s", "cores": 1, "kind": kind, "sample": sample_desc(kind, 240, 320, "backslash"),
            "sample_seconds": dt,
            "own_pcg_rtol_1e-8": {"value": frac / dt_pcg, "unit": "frame-pairs/s", "cores": 1, "sample_seconds": dt_pcg,
                                  "sample": sample_desc(kind, 240, 320, "pcg")},
            "full_size_measured": full_size_measured()}


def workload_config(B, solver_precision="mixed", tol=1e-10):
    """the `config` object -- identical in both arms"""
    return {"workload": "classic+nl-fast, synthetic 640x480 RGB pairs (affine flow, seeds 3..), fp64, "
                        "%d pairs per GPU per step" % B,
            "method": METHOD, "height": H, "width": W, "pairs_per_gpu_per_step": B,
            "parallelism": "independent frame pairs per GPU, no collective",
            # timing rule: flush L2 between timed iterations OR use inputs larger than L2 -- the latter holds by construction
            "l2": "no flush needed: one step streams ~%.1f GB of fp64 working set (pyramids, 72 B/pixel linear systems, 64 B/pixel "
                  "Krylov vectors for %d pairs) and ~650 GB of traffic through the 126 MB L2 between two reads of the same input"
                  % (B * H * W * 650e-9, B)}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores -- one process per core (it is
    single threaded), each step = one bounded sample (160x120 centre crop of a different pair) per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    kind = cpu_kind()
    # bounded sample per step: the 160x120 centre crop (18 s per sample for the reference, 5 s for the port, build host);
    # with many steps a 128x96 crop, so that `--steps 20 --warmup 5` still ends within a few minutes
    sh, sw = (SAMPLE_H, SAMPLE_W) if (args.steps <= 6 or kind == "port") else (96, 128)
    frac = (sh * sw) / float(H * W)
    jobs = [(3 + k, sh, sw, kind, "backslash") for k in range(cores)]
    warm = min(args.warmup, 1)
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(warm):
            pool.map(_cpu_sample, jobs)
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_sample, jobs)
        dt = time.perf_counter() - t0
    value = args.steps * cores * frac / dt
    line = {
        "impl": "reference", "metric": "frame_pairs_per_sec", "value": value, "unit": "frame-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mpix_per_s": value * H * W / 1e6,
        "config": workload_config(args.batch),
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": kind,
                         "sample": sample_desc(kind, sh, sw, "backslash") + "; %d processes, one sample each per step" % cores,
                         "full_size_measured": full_size_measured()},
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


def _timed_device_steps(ctx, P, B, d1, d2, duv, steps, warm, stream, torch):
    """K resident steps of b200flow_estimate_rgb8_dev timed with CUDA events on the library's stream"""
    from optical_flow import _lib
    st = _lib.Stats()

    def step():
        ctx.call("b200flow_estimate_rgb8_dev", P, B, H, W, d1.data_ptr(), d2.data_ptr(), 1, duv.data_ptr(), _lib.C.byref(st))
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, st


def run_variants(args, ctx, d1, d2, duv, B, local_rank, peak):
    """The same workload with (i) the strict all-fp64 solver, (ii) the fp32 variant (IC solver entirely in fp32, stopped at
    1e-6 on its own residual: not parity-grade, statistical parity only), (iii) the block-Jacobi mixed solver, (iv) two
    concurrent sub-batches per GPU (b200flow_ctx_set_split): a few steps each, resident inputs, timed like `value`."""
    import torch
    from optical_flow import _lib, load_of_method
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    out = {}
    steps = max(2, min(3, args.steps))

    def params_for(precision):
        ope = load_of_method(METHOD)
        ope.solver_precision = precision
        ope.pyramid_levels = ope._auto_pyramid_levels(np.empty((H, W, 2)))
        P = ope._c_params(levels=ope.pyramid_levels)
        ope._apply_solver(P)
        return P
    for name, precision, split in (("fp64", "fp64", 1), ("fp32", "fp32", 1), ("mixed_jacobi", "mixed-jacobi", 1),
                                   ("concurrent_groups_2", "mixed", 2)):
        if precision == args.solver_precision and split == (args.split or 1):
            continue
        try:
            ctx.set_split(split)
            ctx.set_timing(False)
            ms, st = _timed_device_steps(ctx, params_for(precision), B, d1, d2, duv, steps, 3, stream, torch)
            ctx.set_timing(True)
            ms_t, st = _timed_device_steps(ctx, params_for(precision), B, d1, d2, duv, 1, 1, stream, torch)
            ach = st.pcg_pixel_iters * PCG_BYTES[precision] / (st.solver_ms / 1e3) / 1e9 if st.solver_ms > 0 else 0.0
            out[name] = {"value": B / (ms / 1e3), "unit": "frame-pairs/s", "ms_per_step": ms, "steps": steps,
                         "solver_precision": precision, "concurrent_groups": split, "bytes_per_pixel_iter": PCG_BYTES[precision],
                         "pcg_iters_per_step": int(st.pcg_iters), "solver_ms_per_step": st.solver_ms,
                         "solver_GBps": ach, "solver_frac_of_hbm_peak": ach / peak, "not_converged": int(st.not_converged)}
        except Exception as e:            # a variant must never take the headline down
            out[name] = {"error": str(e)[:300]}
        finally:
            ctx.set_timing(False)
            ctx.set_split(args.split or 1)
    return out


def run_configs(ctx):
    """BASELINE.json configs 1-5 at FULL size with each preset's own defaults, steady state, through the public API
    (host arrays in, host flow out), measured in this same run: pairs/s, Mpix/s, PCG iterations, per-stage ms; config 1
    also AAE / AEPE against the Middlebury ground truth (the committed RubberWhale fixture)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from optical_flow import _lib, estimate_flow, estimate_flow_batch, interface, flow_angular_error
    rows = []
    ctx.set_timing(True)

    def single(im1, im2, preset, params=None):
        holder = {}
        orig = interface.load_of_method

        def spy(name):
            holder["ope"] = orig(name)
            return holder["ope"]
        interface.load_of_method = spy
        try:
            uv = estimate_flow(im1, im2, preset, params)
        finally:
            interface.load_of_method = orig
        return uv, holder["ope"].last_stats

    def timed(fn, reps):
        out, best = None, float("inf")
        for i in range(reps):       # run 1 grows the device arena, run 2 consolidates it; the last one is steady state
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            if i == reps - 1:
                best = dt
        return out, best

    def row(name, pairs, h, w, dt, st, extra=None):
        r = {"config": name, "pairs": pairs, "height": h, "width": w, "seconds": dt, "pairs_per_s": pairs / dt,
             "mpix_per_s": pairs * h * w / dt / 1e6, "solves": st.get("solves"), "pcg_iters": st.get("pcg_iters"),
             "not_converged": st.get("not_converged"),
             "stage_ms": {k: st.get(k) for k in ("pre_ms", "warp_ms", "solver_ms", "filter_ms", "total_ms")}}
        if extra:
            r.update(extra)
        rows.append(r)
    try:
        # 1: RubberWhale frame10/11 (the real frames + .flo ground truth are a committed fixture)
        g = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_10_11.npz"))
        a, b = g["im1"].astype(float), g["im2"].astype(float)
        (uv, st), dt = timed(lambda: single(a, b, "classic+nl-fast"), 3)
        aae, std, aepe = flow_angular_error(g["tu"], g["tv"], uv[:, :, 0], uv[:, :, 1], 0)
        ref = np.load(os.path.join(ROOT, "tests", "golden", "rubberwhale_full.npz"))
        row("1: classic+nl-fast, Middlebury RubberWhale 584x388, single pair", 1, 388, 584, dt, st,
            {"aae_deg": float(aae), "aepe_px": float(aepe), "reference_aae_deg": float(ref["aae"]),
             "reference_aepe_px": float(ref["aepe"]), "max_abs_diff_vs_reference_flow_px": float(np.abs(uv - ref["uv"]).max()),
             "reference_cpu_seconds": float(ref["seconds"])})
        # 2: hs-brightness 1024x1024
        im1, im2, flow = synth.gray_pair(1024, 1024, seed=0)
        (uv, st), dt = timed(lambda: single(im1, im2, "hs-brightness"), 3)
        row("2: hs-brightness, synthetic 1024x1024", 1, 1024, 1024, dt, st, {"aepe_vs_known_flow_px": synth.interior_epe(uv, flow)})
        # 3: ba 1920x1080
        im1, im2, flow = synth.gray_pair(1080, 1920, seed=1, disc=True)
        (uv, st), dt = timed(lambda: single(im1, im2, "ba"), 3)
        row("3: ba (lorentzian GNC + ROF), synthetic 1920x1080", 1, 1080, 1920, dt, st)
        # 4: classic+nl-full, one GPU's share (8 pairs) of the 64-pair Middlebury-size job
        pairs = [synth_pair(388, 584, 20 + k) for k in range(8)]
        ims1 = np.stack([q[0] for q in pairs]); ims2 = np.stack([q[1] for q in pairs])
        (res, dt) = timed(lambda: estimate_flow_batch(ims1, ims2, "classic+nl-full", return_stats=True), 3)
        uv, st = res
        row("4: classic+nl-full, 8 x 584x388 (one GPU's share of the 64-pair job)", 8, 388, 584, dt, st,
            {"aepe_vs_known_flow_px": float(np.mean([synth.interior_epe(uv[k], pairs[k][2]) for k in range(8)]))})
        # 5: classic++ 3840x2160, one pair
        im1, im2, flow = synth.gray_pair(2160, 3840, seed=2)
        (uv, st), dt = timed(lambda: single(im1, im2, "classic++"), 3)
        row("5: classic++ (generalized Charbonnier, bi-cubic), synthetic 3840x2160, single pair", 1, 2160, 3840, dt, st,
            {"aepe_vs_known_flow_px": synth.interior_epe(uv, flow, margin=16)})
    except Exception as e:
        rows.append({"error": str(e)[:300]})
    finally:
        ctx.set_timing(False)
    return rows


def run_rowband(args, rank, local_rank, world):
    """BASELINE config 5, row-band mode: latency of ONE synthetic 3840x2160 pair with 'classic++' (its own defaults: 3 GNC
    stages x 10 warps) when every linear solve of a big level is split into row bands over the N ranks
    (optical_flow/rowband.py).  Every rank holds the frames and computes the cheap stages redundantly; strong scaling.
    One JSON line on rank 0; value = seconds per pair (max over ranks, device-synchronised wall clock around K calls)."""
    global _stdout_fd
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from optical_flow import _lib, interface, estimate_flow
    from optical_flow.rowband import RowBand
    torch.cuda.set_device(local_rank)
    os.environ["B200FLOW_DEVICE"] = str(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        _stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    Hh, Ww = [int(v) for v in args.rowband_size.lower().split("x")]
    im1, im2, flow = synth.gray_pair(Hh, Ww, seed=2)
    ctx = _lib.default_context(local_rank)
    ctx.set_timing(True)
    holder = {}
    orig = interface.load_of_method

    def spy(name):
        holder["ope"] = orig(name)
        return holder["ope"]
    interface.load_of_method = spy

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    steps = max(1, min(args.steps, 5))
    with RowBand.from_torch_distributed(Hh, Ww, device=local_rank):
        for _ in range(2):
            uv = estimate_flow(im1, im2, "classic++")
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            uv = estimate_flow(im1, im2, "classic++")
        barrier()
        dt = (time.perf_counter() - t0) / steps
        st = holder["ope"].last_stats
    epe = synth.interior_epe(uv, flow, margin=16)
    checksum = float(np.abs(uv).sum())
    if world > 1:
        t = torch.tensor([dt, st["solver_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, solver_ms = [float(v) for v in t.tolist()]
        c = torch.tensor([checksum, -checksum], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        same = bool(c[0].item() == -c[1].item())        # max == min over ranks: every rank holds the same flow
    else:
        solver_ms, same = st["solver_ms"], True
    if rank == 0:
        iters = max(1, int(st["pcg_iters"]))
        line = {"metric": "seconds_per_pair_rowband", "value": dt, "unit": "s", "n_gpus": world, "steps": steps, "warmup": 2,
                "ms_per_step": 1e3 * dt, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "mpix_per_s": Hh * Ww / dt / 1e6,
                "config": {"workload": "classic++ (generalized Charbonnier, bi-cubic, 3 GNC stages x 10 warps), ONE synthetic "
                                       "%dx%d pair, row-band split of the linear solves over %d GPU(s)" % (Ww, Hh, world),
                           "method": "classic++", "height": Hh, "width": Ww,
                           "parallelism": "row bands of whole 8-row strips; per PCG iteration one halo row of z and p from each "
                                          "neighbour over NVLink P2P loads + two flag barriers carrying the dot products; the "
                                          "solution bands are exchanged by P2P stores after every solve; other stages replicated"},
                "solves": st["solves"], "pcg_iters": iters, "not_converged": st["not_converged"],
                "solver_ms": solver_ms, "solver_us_per_iteration": 1e3 * solver_ms / iters,
                "stage_ms": {k: st[k] for k in ("pre_ms", "warp_ms", "solver_ms", "filter_ms", "total_ms")},
                "aepe_vs_known_flow_px": epe, "all_ranks_identical_flow": same, "gpu_launches": int(st["kernel_launches"]) * steps}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--mode", default="batch", choices=["batch", "rowband"],
                    help="batch: the headline workload (independent pairs per GPU); rowband: ONE 3840x2160 classic++ pair, its "
                         "linear solves split into row bands over the N GPUs (NVLink P2P halo loads, no NCCL on the data path)")
    ap.add_argument("--rowband-size", default="2160x3840", help="HxW of the row-band pair")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE.json configs 1-5 section")
    ap.add_argument("--no-variants", action="store_true", help="skip the solver-precision / concurrency variants section")
    ap.add_argument("--split", type=int, default=None,
                    help="concurrent sub-batches per GPU (b200flow_ctx_set_split); default: the library's default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.mode == "rowband":
        run_rowband(args, rank, local_rank, world)
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

    # ---- value: inputs resident in HBM, CUDA events on the library's stream; the library's own per-stage events are
    #      OFF in this region (they are ~520 extra event records per step) ----
    ctx.set_timing(False)
    for _ in range(W_):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pixel_iters = launches = pcg_iters = 0
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            step_resident()
            pixel_iters += st.pcg_pixel_iters
            launches += st.kernel_launches
            pcg_iters += st.pcg_iters
        e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1)
    not_conv = st.not_converged
    # ---- the same K steps again with the per-stage CUDA events on (stage split, solver time for the roofline) ----
    ctx.set_timing(True)
    step_resident()
    barrier()
    solver_ms = 0.0
    stage_ms = {"solver_ms": 0.0, "warp_ms": 0.0, "filter_ms": 0.0, "pre_ms": 0.0}
    NK = len(_lib.Stats.KERNEL_GROUPS)
    k_ms, k_bytes, k_calls = [0.0] * NK, [0.0] * NK, [0] * NK
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e2.record()
        for _ in range(args.steps):
            step_resident()
            solver_ms += st.solver_ms
            for k in stage_ms:
                stage_ms[k] += getattr(st, k)
            for i in range(NK):
                k_ms[i] += st.kernel_ms[i]; k_bytes[i] += st.kernel_bytes[i]; k_calls[i] += st.kernel_calls[i]
        e3.record()
    barrier()
    ms_staged = e2.elapsed_time(e3)
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
            "config": workload_config(B),
            "config_detail": {
                       "solver": "%s PCG until the fp64 true residual ||b-Ax|| <= %g ||b|| (stands in for spsolve); %s"
                                 % (PCG_PRECOND[args.solver_precision], P.tol,
                                    "Krylov vectors fp32, solution + residual replacement fp64" if
                                    args.solver_precision != "fp64" else "all vectors fp64"),
                       "solver_precision": args.solver_precision,
                       "concurrent_groups": int(args.split) if args.split else 1,
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
                         "solver_share_of_step": solver_ms / ms_staged if ms_staged else None,
                         "measured": "second pass of the same K steps with the library's per-stage CUDA events on "
                                     "(ms_per_step of that pass: %.2f)" % (ms_staged / args.steps)},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "pcg_iters_per_step": pcg_iters / args.steps, "pcg_not_converged": int(not_conv),
            "aepe_vs_known_flow_px": epe,
            "clocks": clocks,
        }
        # per kernel group: CUDA-event time, ALGORITHMIC bytes (DESIGN.md section 4 per-pixel figures x pixels of every
        # launch), fraction of the HBM peak.  The weighted median is issue-bound, not HBM-bound: its frac is reported
        # for completeness and its issue-slot utilisation comes from the committed ncu capture (profiles/).
        stages = {}
        for i, name in enumerate(_lib.Stats.KERNEL_GROUPS):
            if k_calls[i] == 0:
                continue
            ms_i, gb_i = k_ms[i] / args.steps, k_bytes[i] / args.steps / 1e9
            gbs = gb_i / (ms_i / 1e3) if ms_i > 0 else 0.0
            stages[name] = {"ms_per_step": ms_i, "launch_groups_per_step": k_calls[i] / args.steps, "algorithmic_GB_per_step": gb_i,
                            "GBps": gbs, "frac_of_hbm_peak": gbs / peak, "bound": "issue" if name == "weighted_median" else "hbm"}
        # issue-slot utilisation of the compute-bound weighted median: from the committed ncu --set full capture of this build
        try:
            for ln in open(os.path.join(ROOT, "profiles", "r02_ncu_full_kernels.jsonl")):
                rec = json.loads(ln) if ln.startswith("{") else {}
                if "wmedian" in str(rec.get("kernel")) and "weighted_median" in stages:
                    stages["weighted_median"].update({"issue_slot_pct_ncu": rec.get("issue_active_pct"),
                                                      "fp64_pipe_pct_ncu": rec.get("fp64_pipe_pct"),
                                                      "warp_instructions_per_pixel_ncu": rec.get("warp_instr", 0) / float(16 * H * W),
                                                      "ncu_source": "profiles/r02_ncu_full_kernels.jsonl (one launch, B = 16, 640x480)"})
        except (OSError, ValueError):
            pass
        line["stages"] = stages
        if world == 1 and not args.no_variants:
            line["variants"] = run_variants(args, ctx, d1, d2, duv, B, local_rank, peak)
        if world == 1 and not args.no_configs:
            line["configs"] = run_configs(ctx)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
