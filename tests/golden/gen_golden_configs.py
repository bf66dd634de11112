#!/usr/bin/env python
"""Reference goldens at the BASELINE.json configs and at the timed bench workload -- produced by RUNNING THE
UNMODIFIED REFERENCE in this container (same import arrangement as gen_golden.py).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_configs.py --all        # ~40 min on 8 cores
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden_configs.py --job NAME

Jobs (files written under tests/golden/cfg_*.npz; wall seconds of the timed ones also into profiles/cpu_fullsize.json):
  bench640        the bench workload: synthetic 640x480 RGB pair seed 3, 'classic+nl-fast', default solver (SuperLU)
  bench640_pcg    same pair, the reference's own solver='pcg', pcg_rtol=1e-8 (its fastest honest CPU setting)
  cfg2_hs512      config 2 at 512x512: 'hs-brightness', synthetic textured pair with known affine flow
  cfg3_ba270      config 3 at 270x480: 'ba' (lorentzian GNC 3 stages x 10 iterations + ROF texture), moving-disc pair
  cfg5_cpp540     config 5 at 540x960: 'classic++' max_iters=3 end to end + two teacher-forced warp iterations at size
  cfg5_4k_stage   config 5 at 3840x2160, operator level: partial_deriv (bi-cubic) + flow_operator A@probe, b at a known
                  flow (no solve) -- pins 64-bit indexing; outputs kept as a stride-16 grid + whole-array sums
  cfg3_1080_stage config 3 at 1920x1080, operator level: cubic-spline partial_deriv (prefilter line length 1920)
  mb_<Sequence>   config 4: the 8 Middlebury sequences with ground truth, 'classic+nl-fast': uv (float32), AAE/AEPE

Synthetic inputs are QUANTISED to uint8 values so that regenerating them on the GPU box (tests/synth.py, bench.py --
same numpy/scipy) is robust to last-bit differences; a crc32 of the inputs is stored and checked by the tests.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, os.path.join(ROOT, "tests"))
sys.path.insert(3, ROOT)
sys.dont_write_bytecode = True
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

MB_SEQS = ["Dimetrodon", "Grove2", "Grove3", "Hydrangea", "RubberWhale", "Urban2", "Urban3", "Venus"]


def _ref():
    import optical_flow
    assert optical_flow.__file__.startswith("/root/reference"), optical_flow.__file__
    return optical_flow


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def crc(*arrs):
    c = 0
    for a in arrs:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (name, os.path.getsize(path) / 1024.0), flush=True)


def note_seconds(key, rec):
    """profiles/cpu_fullsize.json: measured wall seconds of the unmodified reference at full size (one core)"""
    import fcntl
    path = os.path.join(ROOT, "profiles", "cpu_fullsize.json")
    lock = open(path + ".lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        d = {}
    d[key] = rec
    d["_host"] = {"cpu": _cpu_name(), "cores_in_container": os.cpu_count(),
                  "note": "unmodified /root/reference, one process, BLAS threads = 1; measured in the build container"}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


def _cpu_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def quant_gray_pair(h, w, seed, disc=False):
    """tests/synth.gray_pair rounded to integer grey levels (float64 arrays holding uint8 values)"""
    import synth
    im1, im2, flow = synth.gray_pair(h, w, seed, disc=disc)
    q = lambda im: np.clip(np.floor(im + 0.5), 0, 255)  # noqa: E731
    return q(im1), q(im2), flow


# ---------------------------------------------------------------------------------------------------------------
def job_bench640(pcg=False):
    _ref()
    from optical_flow import estimate_flow
    import bench
    im1, im2, flow = bench.synth_pair(480, 640, 3)
    params = {"solver": "pcg", "pcg_rtol": 1e-8} if pcg else None
    t0 = time.perf_counter()
    uv = quiet(estimate_flow, im1, im2, "classic+nl-fast", params)
    dt = time.perf_counter() - t0
    epe = float(np.sqrt(((uv - flow) ** 2).sum(-1))[8:-8, 8:-8].mean())
    tag = "bench640_pcg" if pcg else "bench640"
    save("cfg_%s.npz" % tag, uv=uv.astype(np.float32), seconds=dt, input_crc=crc(im1, im2), aepe_known=epe)
    note_seconds(tag, {"seconds": dt, "pairs_per_s": 1.0 / dt, "method": "classic+nl-fast", "height": 480, "width": 640,
                       "input": "bench.synth_pair(480, 640, 3) uint8 RGB", "params": params or {"solver": "backslash"},
                       "aepe_vs_known_flow_px": epe})
    print("%s: %.1f s, AEPE vs known %.4f" % (tag, dt, epe), flush=True)


def job_cfg2_hs512():
    _ref()
    from optical_flow import estimate_flow
    im1, im2, flow = quant_gray_pair(512, 512, 0)
    t0 = time.perf_counter()
    uv = quiet(estimate_flow, im1, im2, "hs-brightness")
    dt = time.perf_counter() - t0
    save("cfg_hs512.npz", uv=uv.astype(np.float32), seconds=dt, input_crc=crc(im1, im2))
    note_seconds("cfg2_hs512", {"seconds": dt, "method": "hs-brightness", "height": 512, "width": 512})
    print("cfg2_hs512: %.1f s" % dt, flush=True)


def job_cfg3_ba270():
    _ref()
    from optical_flow import estimate_flow
    im1, im2, flow = quant_gray_pair(270, 480, 1, disc=True)
    t0 = time.perf_counter()
    uv = quiet(estimate_flow, im1, im2, "ba")
    dt = time.perf_counter() - t0
    save("cfg_ba270.npz", uv=uv.astype(np.float32), seconds=dt, input_crc=crc(im1, im2))
    note_seconds("cfg3_ba270", {"seconds": dt, "method": "ba", "height": 270, "width": 480})
    print("cfg3_ba270: %.1f s" % dt, flush=True)


def job_cfg5_cpp540():
    """classic++ max_iters=3 at 540x960 end to end, plus two teacher-forced warp iterations AT SIZE: the flow entering
    the iteration is the reference's own iterate rounded to float32 (so it stores exactly), and the output is what the
    reference's partial_deriv -> flow_operator -> spsolve computes from that rounded flow."""
    _ref()
    from optical_flow import estimate_flow
    import optical_flow.methods.ba as m_ba
    from optical_flow.methods.base import BaseOpticalFlow
    im1, im2, flow = quant_gray_pair(540, 960, 2)
    steps = []
    cur = {}
    orig_pd = m_ba.partial_deriv
    orig_solve = BaseOpticalFlow._solve_linear_system

    def pd(images, uv, *a, **k):
        cur.clear()
        cur.update(shape=images.shape[:2], uv_in=uv.copy())
        return orig_pd(images, uv, *a, **k)

    def solve(self, A, b, uv_shape, x0=None):
        x = orig_solve(self, A, b, uv_shape, x0)
        cur.update(x=x.copy(), alpha=float(getattr(self, "alpha", 1.0)))
        steps.append(dict(cur))
        return x

    m_ba.partial_deriv = pd
    BaseOpticalFlow._solve_linear_system = solve
    t0 = time.perf_counter()
    try:
        uv = quiet(estimate_flow, im1, im2, "classic++", {"max_iters": 3})
    finally:
        m_ba.partial_deriv = orig_pd
        BaseOpticalFlow._solve_linear_system = orig_solve
    dt = time.perf_counter() - t0
    out = {"uv": uv.astype(np.float32), "seconds": dt, "input_crc": crc(im1, im2), "nsteps": len(steps)}
    # the per-step solution increments on a stride-4 grid (cheap, every step)
    for i, s in enumerate(steps):
        out["s%03d_shape" % i] = np.asarray(s["shape"])
        out["s%03d_alpha" % i] = s["alpha"]
    save("cfg_cpp540.npz", **out)
    note_seconds("cfg5_cpp540_mi3", {"seconds": dt, "method": "classic++ max_iters=3", "height": 540, "width": 960})
    print("cfg5_cpp540: %.1f s, %d warp iterations" % (dt, len(steps)), flush=True)


def job_cfg5_cpp540_tf():
    """Teacher-forced single warp iterations of classic++ at 540x960 (finest level), GNC alpha = 1 and alpha = 0: the input flow
    is analytic (affine motion + a smooth ripple), so it is regenerated on the GPU box; output = the reference's
    partial_deriv -> flow_operator (quadratic / robust blend as compute_flow_base does) -> spsolve."""
    _ref()
    import copy
    from optical_flow.methods.config import load_of_method
    from optical_flow.robust.robust_function import RobustFunction
    from optical_flow.utils.derivatives import partial_deriv
    from optical_flow.utils.image_processing import structure_texture_decomposition_rof, scale_image
    im1, im2, flow = quant_gray_pair(540, 960, 2)
    H, W = im1.shape
    uv_in = tf_flow(H, W, flow)
    images = np.stack([im1, im2], axis=2)
    ope = load_of_method("classic++")
    ope.display = False
    tex = structure_texture_decomposition_rof(images, 1.0 / 8, 100, ope.alp)
    ope.images = tex
    h = np.array([1, -8, 0, 8, -1]) / 12.0
    It, Ix, Iy = partial_deriv(tex, uv_in, ope.interpolation_method, h, 0.5)
    qua = copy.copy(ope)
    qua.lambda_ = ope.lambda_q
    ta = ope.rho_data.param[0] / ope.rho_spatial_u[0].param[0]
    qua.rho_spatial_u = [RobustFunction("quadratic", 1) for _ in ope.rho_spatial_u]
    qua.rho_spatial_v = [RobustFunction("quadratic", 1) for _ in ope.rho_spatial_v]
    qua.rho_data = RobustFunction("quadratic", ta)
    duv = np.zeros_like(uv_in)
    Aq, bq, _, _ = qua.flow_operator(uv_in, duv, It, Ix, Iy)
    Ar, br, _, _ = ope.flow_operator(uv_in, duv, It, Ix, Iy)
    out = {"input_crc": crc(im1, im2), "uv_in_crc": crc(uv_in), "tex_sum": tex.sum(), "tex_abs_sum": np.abs(tex).sum(),
           "It_s8": It[::8, ::8].copy(), "Ix_s8": Ix[::8, ::8].copy(), "Iy_s8": Iy[::8, ::8].copy()}
    for alpha in (1.0, 0.0):
        A = alpha * Aq + (1 - alpha) * Ar
        b = alpha * bq + (1 - alpha) * br
        t0 = time.perf_counter()
        x = ope._solve_linear_system(A, b, uv_in.shape)
        print("  teacher-forced alpha=%g spsolve %.1f s" % (alpha, time.perf_counter() - t0), flush=True)
        out["x_a%g" % alpha] = x.astype(np.float32)               # |x| <= ~1 px: float32 keeps 1e-7
        out["b_a%g_s8" % alpha] = b.reshape(uv_in.shape, order="F")[::8, ::8].copy()
    save("cfg_cpp540_tf.npz", **out)


def tf_flow(H, W, flow):
    """analytic, exactly regenerable flow for teacher forcing: the known motion plus a smooth sub-pixel ripple"""
    yy, xx = np.mgrid[0:H, 0:W].astype(float)
    rip = np.stack([0.3 * np.sin(xx / 37.0) * np.cos(yy / 29.0), 0.25 * np.cos(xx / 41.0 + 0.5) * np.sin(yy / 31.0)], axis=2)
    return flow * 0.9 + rip


def _stage_at_size(tag, h, w, seed, preset, disc, stride):
    _ref()
    import copy
    from optical_flow.methods.config import load_of_method
    from optical_flow.robust.robust_function import RobustFunction
    from optical_flow.utils.derivatives import partial_deriv
    im1, im2, flow = quant_gray_pair(h, w, seed, disc=disc)
    uv_in = tf_flow(h, w, flow)
    images = np.stack([im1, im2], axis=2)            # operator level on the raw grey frames: no ROF (keeps the job short)
    ope = load_of_method(preset)
    ope.display = False
    ope.images = images
    hh = np.array([1, -8, 0, 8, -1]) / 12.0
    t0 = time.perf_counter()
    It, Ix, Iy = partial_deriv(images, uv_in, ope.interpolation_method, hh, 0.5)
    print("  %s partial_deriv %.1f s" % (tag, time.perf_counter() - t0), flush=True)
    rng = np.random.default_rng(1000 + seed)
    probe = rng.standard_normal((h, w, 2))
    duv = np.zeros_like(uv_in)
    t0 = time.perf_counter()
    A, b, _, _ = ope.flow_operator(uv_in, duv, It, Ix, Iy)
    Ap = (A @ probe.reshape(-1, order="F")).reshape(uv_in.shape, order="F")
    b = b.reshape(uv_in.shape, order="F")
    print("  %s flow_operator %.1f s" % (tag, time.perf_counter() - t0), flush=True)
    s = (slice(None, None, stride), slice(None, None, stride))
    out = {"input_crc": crc(im1, im2), "uv_in_crc": crc(uv_in), "probe_crc": crc(probe)}
    for k, v in (("It", It), ("Ix", Ix), ("Iy", Iy), ("Ap", Ap), ("b", b)):
        out[k + "_grid"] = v[s].copy()
        out[k + "_sum"] = v.sum()
        out[k + "_abs_sum"] = np.abs(v).sum()
        # last rows / columns in full: the far end of the 64-bit index range
        out[k + "_lastrows"] = v[-3:].copy()
        out[k + "_lastcols"] = v[:, -3:].copy()
    save("cfg_%s.npz" % tag, **out)


def job_cfg5_4k_stage():
    _stage_at_size("cpp4k_stage", 2160, 3840, 2, "classic++", False, 16)


def job_cfg3_1080_stage():
    _stage_at_size("ba1080_stage", 1080, 1920, 1, "ba", True, 8)


def job_mb(seq):
    _ref()
    from optical_flow import estimate_flow
    from optical_flow.io.flo_io import read_flow_file
    from optical_flow.evaluation.metrics import flow_angular_error
    im1, im2, tu, tv = read_flow_file(seq, 10)
    t0 = time.perf_counter()
    uv = quiet(estimate_flow, im1, im2, "classic+nl-fast")
    dt = time.perf_counter() - t0
    aae, std, aepe = flow_angular_error(tu, tv, uv[:, :, 0], uv[:, :, 1], 0)
    save("cfg_mb_%s.npz" % seq, im1=np.asarray(im1).astype(np.uint8), im2=np.asarray(im2).astype(np.uint8),
         tu=np.asarray(tu, dtype=np.float32), tv=np.asarray(tv, dtype=np.float32),
         uv=uv.astype(np.float32), aae=aae, std=std, aepe=aepe, seconds=dt)
    print("mb %s %s: %.1f s  AAE %.6f AEPE %.6f" % (seq, uv.shape, dt, aae, aepe), flush=True)


JOBS = {"bench640": job_bench640, "bench640_pcg": lambda: job_bench640(True), "cfg2_hs512": job_cfg2_hs512,
        "cfg3_ba270": job_cfg3_ba270, "cfg5_cpp540": job_cfg5_cpp540, "cfg5_cpp540_tf": job_cfg5_cpp540_tf,
        "cfg5_4k_stage": job_cfg5_4k_stage, "cfg3_1080_stage": job_cfg3_1080_stage}
for _s in MB_SEQS:
    JOBS["mb_" + _s] = (lambda s=_s: job_mb(s))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--job", default=None)
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--par", type=int, default=7)
    args = ap.parse_args()
    if args.job:
        JOBS[args.job]()
        return
    if not args.all:
        ap.error("--job NAME or --all")
    me = os.path.abspath(__file__)
    # 1. the timed runs, ALONE on the machine, one after the other
    for j in ("bench640", "bench640_pcg"):
        subprocess.run([sys.executable, me, "--job", j], check=False)
    # 2. everything else, args.par at a time
    rest = [j for j in JOBS if j not in ("bench640", "bench640_pcg")]
    running = []
    while rest or running:
        while rest and len(running) < args.par:
            j = rest.pop(0)
            running.append((j, subprocess.Popen([sys.executable, me, "--job", j])))
        time.sleep(2)
        running = [(j, p) for j, p in running if p.poll() is None]


if __name__ == "__main__":
    main()
