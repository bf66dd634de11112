// solve.cu -- kernel group (4): matrix-free preconditioned conjugate gradients on the coupled 2N x 2N
// five-point system, one persistent cooperative (grid-synchronised) kernel per solve, batched over B
// independent systems.  Replaces scipy.sparse.linalg.spsolve / cg behind BaseOpticalFlow._solve_linear_system
// (base.py:87-136): the sparse matrix is never built.
//
// Per iteration: 2 phases separated by grid.sync(); fp64 vectors; algorithmic bytes per pixel:
//   A  p = z + beta p_old (recomputed on the fly at the 5 stencil points, so the textbook "p update" pass and its
//      grid sync disappear), x += alpha_prev p_old (deferred: p_old is in registers anyway), Ap = A p, partial p.Ap
//        read z 16, p_old 16, x 16, D 16, a12 8, WH 16, WV 16; write p 16, x 16, Ap 16                   = 152
//   B  r -= alpha Ap, z = M^-1 r, partial r.z, r.r
//        read r 16, Ap 16, Minv 12 (fp32 block-Jacobi inverse); write r 16, z 16                           = 76
//                                                                                            total        228 B
// Dot products: per-thread accumulation -> warp shuffle tree -> shared -> one double per (CTA, system),
// then every CTA re-reduces the per-CTA partials of the systems it owns in a fixed order, so all CTAs get
// bit-identical scalars and results are run-to-run deterministic (no floating-point atomics).
#include "solve_shared.cuh"

namespace cg = cooperative_groups;

namespace bf {

struct PcgParams {
  LinSys sys;
  PcgWork w;
  double2 *x;
  double tol2;                   // tol^2
  int maxit;
  int scalar_jacobi;
  int tiles_x, tiles_y, tiles_per_sys, tiles_per_cta;
  long long total_tiles;
};

// p_new = z + beta * p_old at pixel j (computed on the fly: phase C of the textbook algorithm is fused into the matvec)
__device__ __forceinline__ double2 pnew_at(const double2 *z, const double2 *pold, long long j, double beta) {
  double2 zz = z[j], pp = pold[j];
  return make_double2(zz.x + beta * pp.x, zz.y + beta * pp.y);
}

#ifndef PCG_MINB
#define PCG_MINB 3
#endif
__global__ void __launch_bounds__(PCG_THREADS, PCG_MINB) pcg_kernel(PcgParams P) {
  cg::grid_group grid = cg::this_grid();
  const LinSys &S = P.sys;
  const int G = gridDim.x, cta = blockIdx.x;
  const int H = S.H, W = S.W, B = S.B;
  const long long HW = (long long)H * W;
  const long long n_all = (long long)B * HW;
  const int tps = P.tiles_per_sys, tpc = P.tiles_per_cta;
  const long long t0 = (long long)cta * tpc;
  const long long t1 = t0 + tpc < P.total_tiles ? t0 + tpc : P.total_tiles;
  const bool has_work = t0 < t1;
  const int b_first = has_work ? (int)(t0 / tps) : 0;
  const int b_last = has_work ? (int)((t1 - 1) / tps) : -1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

  __shared__ double sm_red[2][PCG_THREADS / 32];
  __shared__ double s_rz[MAXLOC], s_bb[MAXLOC], s_alpha[MAXLOC], s_beta[MAXLOC], s_rr[MAXLOC];
  __shared__ int s_done[MAXLOC];

  double *part_pap = P.w.partial;                      // [B][G]
  double *part_rz = P.w.partial + (long long)B * G;
  double *part_rr = P.w.partial + 2LL * B * G;
  double *part_bb = P.w.partial + 3LL * B * G;
  float *Minv = P.w.Minv;                               // [3][B*H*W] block-Jacobi inverse, fp32 (a preconditioner need not be exact)
  double2 *r = P.w.r, *Ap = P.w.Ap, *z = P.w.z, *x = P.x;
  double2 *pold = P.w.p, *pnew = P.w.p2;                // ping-pong search directions
  int *ndone = P.w.flags;                               // [0]
  int *done_g = P.w.flags + 1;                          // [B]
  int *iters_g = P.w.flags + 1 + B;                     // [B]
  double *relres_g = P.w.scal;                          // [B]

  // tile iteration helper: runs the body for every pixel (i, px, py) of this CTA's tiles of system bsys
#define FOR_TILES_OF(bsys, ...)                                                                  \
  {                                                                                              \
    long long ta = (long long)(bsys) * tps > t0 ? (long long)(bsys) * tps : t0;                  \
    long long tb = (long long)((bsys) + 1) * tps < t1 ? (long long)((bsys) + 1) * tps : t1;      \
    for (long long t = ta; t < tb; ++t) {                                                        \
      int tl = (int)(t - (long long)(bsys) * tps);                                               \
      int px = (tl % P.tiles_x) * TILE_W + tx, py = (tl / P.tiles_x) * TILE_H + ty;              \
      if (px < W && py < H) {                                                                    \
        long long i = (long long)(bsys) * HW + (long long)py * W + px;                           \
        __VA_ARGS__                                                                              \
      }                                                                                          \
    }                                                                                            \
  }

  // ---------------- init: x = 0, r = b, Minv, z = Minv r, p_old = 0; partial r.z and b.b ----------------
  for (int b = b_first; b <= b_last; ++b) {
    double acc_rz = 0.0, acc_bb = 0.0;
    FOR_TILES_OF(b, {
      Stencil s = load_stencil(S, i, px, py);
      double duu = s.d.x + s.wr.x + s.wl.x + s.wd.x + s.wu.x;
      double dvv = s.d.y + s.wr.y + s.wl.y + s.wd.y + s.wu.y;
      double m11, m12, m22;
      double det = duu * dvv - s.a12 * s.a12;
      if (!P.scalar_jacobi && det > 0.0 && det > 1e-14 * fabs(duu * dvv)) {
        double inv = 1.0 / det;
        m11 = dvv * inv; m22 = duu * inv; m12 = -s.a12 * inv;
      } else {
        m11 = fabs(duu) > 1e-12 ? 1.0 / duu : 0.0;
        m22 = fabs(dvv) > 1e-12 ? 1.0 / dvv : 0.0;
        m12 = 0.0;
      }
      float f11 = (float)m11, f12 = (float)m12, f22 = (float)m22;
      Minv[i] = f11; Minv[n_all + i] = f12; Minv[2 * n_all + i] = f22;
      double2 rb = __ldg(&S.rhs[i]);
      double2 zz = make_double2((double)f11 * rb.x + (double)f12 * rb.y, (double)f12 * rb.x + (double)f22 * rb.y);
      x[i] = make_double2(0.0, 0.0);
      r[i] = rb;
      z[i] = zz;
      pold[i] = make_double2(0.0, 0.0);
      acc_rz += rb.x * zz.x + rb.y * zz.y;
      acc_bb += rb.x * rb.x + rb.y * rb.y;
    })
    block_sum2(acc_rz, acc_bb, sm_red);
    if (threadIdx.x == 0) { part_rz[(long long)b * G + cta] = acc_rz; part_bb[(long long)b * G + cta] = acc_bb; }
  }
  grid.sync();
  if (threadIdx.x < 32) {
    for (int b = b_first; b <= b_last; ++b) {
      int c_lo = (int)(((long long)b * tps) / tpc), c_hi = (int)((((long long)(b + 1)) * tps - 1) / tpc);
      double rz = reduce_partials(part_rz, G, b, c_lo, c_hi);
      double bb = reduce_partials(part_bb, G, b, c_lo, c_hi);
      if (threadIdx.x == 0) {
        int l = b - b_first;
        s_rz[l] = rz; s_bb[l] = bb; s_rr[l] = bb; s_alpha[l] = 0.0; s_beta[l] = 0.0;
        int dn = !(bb > 0.0) || !(rz > 0.0);          // zero right-hand side: x = 0 is the solution
        s_done[l] = dn;
        if (dn && cta == c_lo) { done_g[b] = 1; iters_g[b] = 0; relres_g[b] = 0.0; atomicAdd(ndone, 1); }
      }
    }
  }
  __syncthreads();
  grid.sync();     // ndone visible to every CTA

  int k = 0;
  for (; k < P.maxit; ++k) {
    // ---------------- phase A: p = z + beta p_old (on the fly, 5 points), x += alpha_prev p_old, Ap = A p ----------------
    for (int b = b_first; b <= b_last; ++b) {
      const int l = b - b_first;
      if (s_done[l]) continue;
      const double beta = s_beta[l], aprev = s_alpha[l];
      double acc = 0.0, dummy = 0.0;
      FOR_TILES_OF(b, {
        // branch-free gather: every load is issued unconditionally at a clamped (always valid) address so that all
        // ~16 of them are in flight together; out-of-image neighbours are then replaced by the centre value
        // (difference 0) -- their edge weights are 0 by construction anyway (no edge leaves the last column / row)
        const long long jl = i > 0 ? i - 1 : 0, jr = i + 1 < n_all ? i + 1 : n_all - 1;
        const long long ju = i >= W ? i - W : 0, jd = i + W < n_all ? i + W : n_all - 1;
        const double2 zc = z[i], po = pold[i];
        const double2 zl = z[jl], pl = pold[jl], zr = z[jr], pr = pold[jr];
        const double2 zu = z[ju], pu = pold[ju], zd = z[jd], pd = pold[jd];
        const double2 sd = __ldg(&S.D[i]), swr = __ldg(&S.WH[i]), swd = __ldg(&S.WV[i]);
        const double2 swl = __ldg(&S.WH[jl]), swu = __ldg(&S.WV[ju]);
        const double sa12 = __ldg(&S.a12[i]);
        const double2 c = make_double2(zc.x + beta * po.x, zc.y + beta * po.y);
        double2 nl = make_double2(zl.x + beta * pl.x, zl.y + beta * pl.y);
        double2 nr = make_double2(zr.x + beta * pr.x, zr.y + beta * pr.y);
        double2 nu = make_double2(zu.x + beta * pu.x, zu.y + beta * pu.y);
        double2 nd = make_double2(zd.x + beta * pd.x, zd.y + beta * pd.y);
        if (px == 0) nl = c;
        if (px + 1 >= W) nr = c;
        if (py == 0) nu = c;
        if (py + 1 >= H) nd = c;
        double au = sd.x * c.x + sa12 * c.y;
        double av = sa12 * c.x + sd.y * c.y;
        au += swr.x * (c.x - nr.x); av += swr.y * (c.y - nr.y);
        au += swl.x * (c.x - nl.x); av += swl.y * (c.y - nl.y);
        au += swd.x * (c.x - nd.x); av += swd.y * (c.y - nd.y);
        au += swu.x * (c.x - nu.x); av += swu.y * (c.y - nu.y);
        if (k > 0) {                                   // deferred x update of the previous iteration (p_old is in registers)
          double2 xc = x[i];
          x[i] = make_double2(xc.x + aprev * po.x, xc.y + aprev * po.y);
        }
        pnew[i] = c;
        Ap[i] = make_double2(au, av);
        acc += c.x * au + c.y * av;
      })
      block_sum2(acc, dummy, sm_red);
      if (threadIdx.x == 0) part_pap[(long long)b * G + cta] = acc;
    }
    grid.sync();
    // Termination test placed HERE on purpose: ndone is only incremented between the second grid.sync of an
    // iteration and the first grid.sync of the next one, so every CTA reads the same value at this point.
    if (*(volatile int *)ndone >= B) break;
    if (threadIdx.x < 32) {
      for (int b = b_first; b <= b_last; ++b) {
        int l = b - b_first;
        if (s_done[l]) continue;
        int c_lo = (int)(((long long)b * tps) / tpc), c_hi = (int)((((long long)(b + 1)) * tps - 1) / tpc);
        double pap = reduce_partials(part_pap, G, b, c_lo, c_hi);
        if (threadIdx.x == 0) s_alpha[l] = pap > 0.0 ? s_rz[l] / pap : 0.0;   // 0 => breakdown, handled below
      }
    }
    __syncthreads();
    // ---------------- phase B: r -= alpha Ap, z = M^-1 r, partial r.z and r.r ----------------
    for (int b = b_first; b <= b_last; ++b) {
      int l = b - b_first;
      if (s_done[l]) continue;
      double alpha = s_alpha[l];
      double acc_rz = 0.0, acc_rr = 0.0;
      FOR_TILES_OF(b, {
        double2 rc = r[i], ac = Ap[i];
        rc.x -= alpha * ac.x; rc.y -= alpha * ac.y;
        double m11 = (double)Minv[i], m12 = (double)Minv[n_all + i], m22 = (double)Minv[2 * n_all + i];
        double2 zz = make_double2(m11 * rc.x + m12 * rc.y, m12 * rc.x + m22 * rc.y);
        r[i] = rc; z[i] = zz;
        acc_rz += rc.x * zz.x + rc.y * zz.y;
        acc_rr += rc.x * rc.x + rc.y * rc.y;
      })
      block_sum2(acc_rz, acc_rr, sm_red);
      if (threadIdx.x == 0) { part_rz[(long long)b * G + cta] = acc_rz; part_rr[(long long)b * G + cta] = acc_rr; }
    }
    grid.sync();
    if (threadIdx.x < 32) {
      for (int b = b_first; b <= b_last; ++b) {
        int l = b - b_first;
        if (s_done[l]) continue;
        int c_lo = (int)(((long long)b * tps) / tpc), c_hi = (int)((((long long)(b + 1)) * tps - 1) / tpc);
        double rz = reduce_partials(part_rz, G, b, c_lo, c_hi);
        double rr = reduce_partials(part_rr, G, b, c_lo, c_hi);
        if (threadIdx.x == 0) {
          double alpha = s_alpha[l];
          s_beta[l] = s_rz[l] > 0.0 ? rz / s_rz[l] : 0.0;
          s_rz[l] = rz;
          s_rr[l] = rr;
          int dn = (rr <= P.tol2 * s_bb[l]) || !(alpha > 0.0) || !(rz > 0.0) || !(rr == rr);
          if (dn) {
            s_done[l] = 2;                              // 2: finished in this iteration, x still lacks alpha_k p_k
            if (cta == c_lo) {
              done_g[b] = (rr <= P.tol2 * s_bb[l]) ? 1 : 2;
              iters_g[b] = k + 1;
              relres_g[b] = sqrt(rr / s_bb[l]);
              atomicAdd(ndone, 1);
            }
          }
        }
      }
    }
    __syncthreads();
    // final x update of systems that just finished (own pixels only: no grid sync needed)
    for (int b = b_first; b <= b_last; ++b) {
      int l = b - b_first;
      if (s_done[l] != 2) continue;
      double alpha = s_alpha[l];
      FOR_TILES_OF(b, {
        double2 xc = x[i], pc = pnew[i];
        x[i] = make_double2(xc.x + alpha * pc.x, xc.y + alpha * pc.y);
      })
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int b = b_first; b <= b_last; ++b)
        if (s_done[b - b_first] == 2) s_done[b - b_first] = 1;
    __syncthreads();
    { double2 *t = pold; pold = pnew; pnew = t; }
  }
  // systems that ran out of iterations: apply the pending x update (p_k now lives in pold) and report
  for (int b = b_first; b <= b_last; ++b) {
    int l = b - b_first;
    if (s_done[l] || k == 0) continue;
    double alpha = s_alpha[l];
    FOR_TILES_OF(b, {
      double2 xc = x[i], pc = pold[i];
      x[i] = make_double2(xc.x + alpha * pc.x, xc.y + alpha * pc.y);
    })
  }
#undef FOR_TILES_OF
  if (threadIdx.x == 0) {
    for (int b = b_first; b <= b_last; ++b) {
      int l = b - b_first;
      int c_lo = (int)(((long long)b * tps) / tpc);
      if (!s_done[l] && cta == c_lo) {
        done_g[b] = 3;
        iters_g[b] = k;
        relres_g[b] = sqrt(s_rr[l] / s_bb[l]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Mixed-precision PCG with reliable updates (the default "exact" solver).
//
// The Krylov vectors r, z, p, Ap and the partial solution y live in fp32, as does a copy of the stencil
// coefficients; the solution x is accumulated in fp64 and, whenever the iterated fp32 residual has dropped by
// `delta` since the last update (or claims convergence), the TRUE residual r = b - A x is recomputed in fp64 from
// the fp64 coefficients and replaces the iterated one ("reliable update": the search direction is kept, so CG does
// not restart).  Convergence is only ever declared on that fp64 true residual, so the result satisfies the same
// ||b - A x|| <= tol ||b|| criterion as the all-fp64 kernel above; on the Classic+NL systems the iteration counts
// are identical (scripts/mp_proto.py: 470 vs 470 iterations on RubberWhale full resolution, alpha = 0).
//
// Algorithmic bytes per pixel-iteration (fp32 working set):
//   A  read z 8, p_old 8, y 8, D 8, a12 4, WH 8, WV 8; write p 8, y 8, Ap 8                              = 76
//   B  read r 8, Ap 8, Minv 12; write r 8, z 8                                                            = 44
//                                                                                            total        120 B
// plus, per reliable update (every 20-100 iterations), ONE extra phase and grid barrier: r = b - A (x + y + alpha p)
// evaluated on the fly at the five stencil points (x 16, y 8, p 8, fp64 coefficients 72, b 16, Minv 12; write r, z 16
// = 148 B); the flush x += y + alpha p itself rides on the next phase A.
//
// Every CTA keeps the scalars of ALL systems of the batch (reduced redundantly from the per-CTA partials in a fixed
// order), so every control decision -- reliable update now? system finished? loop done? -- is taken identically by
// every CTA without flags, atomics or extra grid barriers.  The same shared knowledge lets the tiles of the systems that
// are still ACTIVE be re-dealt over the whole grid each time a system finishes, so a batch whose systems need different
// iteration counts keeps every SM streaming until the last one is done.
// ------------------------------------------------------------------------------------------------
// cold paths of the mixed kernel (initialisation, reliable update)
__device__ __forceinline__ double2 mix_init_pixel(const MixParams &P, long long i, int px, int py, long long n_all,
                                               float2 *pold) {
  const LinSys &S = P.sys;
  Stencil s = load_stencil(S, i, px, py);
  double duu = s.d.x + s.wr.x + s.wl.x + s.wd.x + s.wu.x;
  double dvv = s.d.y + s.wr.y + s.wl.y + s.wd.y + s.wu.y;
  double m11, m12, m22;
  double det = duu * dvv - s.a12 * s.a12;
  if (det > 0.0 && det > 1e-14 * fabs(duu * dvv)) {
    double inv = 1.0 / det;
    m11 = dvv * inv; m22 = duu * inv; m12 = -s.a12 * inv;
  } else {
    m11 = fabs(duu) > 1e-12 ? 1.0 / duu : 0.0;
    m22 = fabs(dvv) > 1e-12 ? 1.0 / dvv : 0.0;
    m12 = 0.0;
  }
  float f11 = (float)m11, f12 = (float)m12, f22 = (float)m22;
  float *Minv = P.w.Minv;
  Minv[i] = f11; Minv[n_all + i] = f12; Minv[2 * n_all + i] = f22;
  P.m.D[i] = make_float2((float)s.d.x, (float)s.d.y);
  P.m.a12[i] = (float)s.a12;
  P.m.WH[i] = make_float2((float)s.wr.x, (float)s.wr.y);
  P.m.WV[i] = make_float2((float)s.wd.x, (float)s.wd.y);
  double2 rb = __ldg(&S.rhs[i]);
  float2 rs = make_float2((float)rb.x, (float)rb.y);
  float2 zz = make_float2(f11 * rs.x + f12 * rs.y, f12 * rs.x + f22 * rs.y);
  P.x[i] = make_double2(0.0, 0.0);
  P.m.r[i] = rs; P.m.z[i] = zz;
  pold[i] = make_float2(0.f, 0.f);
  P.m.y[i] = make_float2(0.f, 0.f);
  return make_double2((double)rs.x * (double)zz.x + (double)rs.y * (double)zz.y, rb.x * rb.x + rb.y * rb.y);
}

// true residual at pixel i of the solution x + y + alpha p (evaluated on the fly at the five stencil points)
__device__ __forceinline__ double2 mix_reliable_pixel(const MixParams &P, long long i, int px, int py, long long n_all,
                                                   const float2 *pnew, double alpha) {
  const LinSys &S = P.sys;
  const int W = S.W, H = S.H;
  const double2 *x = P.x;
  const float2 *y = P.m.y;
  const long long jl = px > 0 ? i - 1 : i, jr = px + 1 < W ? i + 1 : i;
  const long long ju = py > 0 ? i - W : i, jd = py + 1 < H ? i + W : i;
#define XTRUE(j, out)                                                           \
  {                                                                             \
    double2 xx = x[j]; float2 yy = y[j], pp = pnew[j];                          \
    out = make_double2(xx.x + ((double)yy.x + alpha * (double)pp.x),            \
                       xx.y + ((double)yy.y + alpha * (double)pp.y));           \
  }
  double2 c, nl, nr, nu, nd;
  XTRUE(i, c) XTRUE(jl, nl) XTRUE(jr, nr) XTRUE(ju, nu) XTRUE(jd, nd)
#undef XTRUE
  const double2 sd = __ldg(&S.D[i]), swr = __ldg(&S.WH[i]), swd = __ldg(&S.WV[i]);
  const double2 swl = __ldg(&S.WH[jl]), swu = __ldg(&S.WV[ju]);   // multiplied by a zero difference when jl == i / ju == i
  const double sa12 = __ldg(&S.a12[i]);
  double au = sd.x * c.x + sa12 * c.y;
  double av = sa12 * c.x + sd.y * c.y;
  au += swr.x * (c.x - nr.x); av += swr.y * (c.y - nr.y);
  au += swl.x * (c.x - nl.x); av += swl.y * (c.y - nl.y);
  au += swd.x * (c.x - nd.x); av += swd.y * (c.y - nd.y);
  au += swu.x * (c.x - nu.x); av += swu.y * (c.y - nu.y);
  double2 rb = __ldg(&S.rhs[i]);
  double ru = rb.x - au, rv = rb.y - av;
  float2 rs = make_float2((float)ru, (float)rv);
  const float *Minv = P.w.Minv;
  float m11 = Minv[i], m12 = Minv[n_all + i], m22 = Minv[2 * n_all + i];
  float2 zz = make_float2(m11 * rs.x + m12 * rs.y, m12 * rs.x + m22 * rs.y);
  P.m.r[i] = rs; P.m.z[i] = zz;
  return make_double2((double)rs.x * (double)zz.x + (double)rs.y * (double)zz.y, ru * ru + rv * rv);
}

// tuning (B200, round 1, bench workload): 2 CTAs/SM x 128 registers with phase A unrolled 2x beats 3 CTAs x 80 registers
// (which spills): 256.7 vs 295.8 ms of solver time per 16-pair step
#ifndef MIX_MINB
#define MIX_MINB 2
#endif
#ifndef MIX_UA
#define MIX_UA 2
#endif
#ifndef MIX_UB
#define MIX_UB 4
#endif

__global__ void __launch_bounds__(PCG_THREADS, MIX_MINB) pcg_mixed_kernel(MixParams P) {
  cg::grid_group grid = cg::this_grid();
  const LinSys &S = P.sys;
  const int G = gridDim.x, cta = blockIdx.x;
  const int H = S.H, W = S.W, B = S.B;
  const long long HW = (long long)H * W;
  const long long n_all = (long long)B * HW;
  const int tps = P.tiles_per_sys;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  int GW = 32;
  while (GW > 1 && B * GW > PCG_THREADS) GW >>= 1;

  __shared__ double sm_red[2][PCG_THREADS / 32];
  __shared__ double s_rz[MAXB], s_bb[MAXB], s_alpha[MAXB], s_beta[MAXB], s_maxr2[MAXB], s_rzprev[MAXB];
  __shared__ double s_ta[MAXB], s_tb[MAXB];
  __shared__ int s_state[MAXB];          // 0 active, 1 finished, 2 converged: final flush pending, 3 reliable update in progress
  __shared__ int s_bad[MAXB], s_flush[MAXB], s_restarts[MAXB], s_stalls[MAXB];
  __shared__ double s_lasttrue[MAXB];
  __shared__ int s_act[MAXB], s_pos[MAXB];   // compact list of unfinished systems and its inverse
  __shared__ int s_nact, s_tpc;

  double *part_a = P.w.partial;                         // [B][G]  p.Ap
  double *part_b = P.w.partial + (long long)B * G;      // [B][G]  r.z
  double *part_c = P.w.partial + 2LL * B * G;           // [B][G]  r.r
  // reliable-update partials get their own slices: a fast CTA reaches its reliable update while a slow CTA is still
  // summing the phase-B slots (no grid barrier in between), so it must not write part_b / part_c there
  double *part_d = P.w.partial + 3LL * B * G;           // [B][G]  r.z of the replaced residual
  double *part_e = P.w.partial + 4LL * B * G;           // [B][G]  r.r of the replaced residual
  float *Minv = P.w.Minv;
  float2 *r = P.m.r, *z = P.m.z, *Ap = P.m.Ap, *y = P.m.y;
  float2 *pold = P.m.p, *pnew = P.m.p2;
  float2 *Df = P.m.D, *WHf = P.m.WH, *WVf = P.m.WV;
  float *a12f = P.m.a12;
  double2 *x = P.x;
  int *done_g = P.w.flags + 1;
  int *iters_g = P.w.flags + 1 + B;
  double *relres_g = P.w.scal;

  // ---- tile ownership: the tiles of the unfinished systems, in compact order, are dealt to the CTAs in contiguous
  //      chunks of s_tpc tiles; recomputed (identically by every CTA) whenever a system finishes
#define REMAP()                                                                         \
  {                                                                                     \
    __syncthreads();                                                                    \
    if (tid == 0) {                                                                     \
      int n = 0;                                                                        \
      for (int b = 0; b < B; ++b) {                                                     \
        if (s_state[b] != 1) { s_act[n] = b; s_pos[b] = n; ++n; } else s_pos[b] = -1;   \
      }                                                                                 \
      s_nact = n;                                                                       \
      long long tt = (long long)n * tps;                                                \
      s_tpc = tt > 0 ? (int)((tt + G - 1) / G) : 1;                                     \
    }                                                                                   \
    __syncthreads();                                                                    \
  }
  // first / last CTA that owns tiles of system b (b must be unfinished)
#define C_LO(b) ((int)(((long long)s_pos[b] * tps) / s_tpc))
#define C_HI(b) ((int)((((long long)(s_pos[b] + 1)) * tps - 1) / s_tpc))
  // this CTA's range of compact system slots
#define OWN_RANGE()                                                                              \
  const long long t0 = (long long)cta * s_tpc;                                                   \
  const long long tt_ = (long long)s_nact * tps;                                                 \
  const long long t1 = t0 + s_tpc < tt_ ? t0 + s_tpc : tt_;                                      \
  const int a_first = t0 < t1 ? (int)(t0 / tps) : 0;                                             \
  const int a_last = t0 < t1 ? (int)((t1 - 1) / tps) : -1;
#define TILE_RANGE(aslot)                                                                        \
  const int b = s_act[aslot];                                                                    \
  const long long tbase = (long long)(aslot) * tps;                                              \
  const long long ta = tbase > t0 ? tbase : t0;                                                  \
  const long long tb = tbase + tps < t1 ? tbase + tps : t1;                                      \
  const long long base = (long long)b * HW;
#define PIXEL_OF(t, px, py, i, ok)                                                               \
  {                                                                                              \
    int tl = (int)((t) - tbase);                                                                 \
    px = (tl % P.tiles_x) * TILE_W + tx;                                                         \
    py = (tl / P.tiles_x) * TILE_H + ty;                                                         \
    ok = (t) < tb && px < W && py < H;                                                           \
    i = ok ? base + (long long)py * W + px : base;                                               \
  }
  // every CTA reduces the partials of every system in `want` state, all systems at once: a group of GW threads
  // (GW = 2..32, the largest power of two with B*GW <= 256) per system, fixed summation order
#define REDUCE_ALL(pa, pb, want)                                                        \
  {                                                                                     \
    const int b = tid / GW, gl = tid % GW;                                              \
    double va = 0.0, vb = 0.0;                                                          \
    if (b < B && s_state[b] == (want)) {                                                \
      const int c_lo = C_LO(b), c_hi = C_HI(b);                                         \
      const volatile double *qa = (pa) + (long long)b * G;                              \
      const volatile double *qb = (pb) ? (pb) + (long long)b * G : qa;                  \
      for (int c = c_lo + gl; c <= c_hi; c += GW) { va += qa[c]; vb += qb[c]; }         \
    }                                                                                   \
    for (int o = GW >> 1; o > 0; o >>= 1) {                                             \
      va += __shfl_xor_sync(0xffffffffu, va, o);                                        \
      vb += __shfl_xor_sync(0xffffffffu, vb, o);                                        \
    }                                                                                   \
    if (b < B && gl == 0) { s_ta[b] = va; s_tb[b] = vb; }                               \
    __syncthreads();                                                                    \
  }

  // ---------------- init ----------------
  for (int b = tid; b < MAXB; b += PCG_THREADS) { s_state[b] = b < B ? 0 : 1; s_bad[b] = 0; s_flush[b] = 0; s_restarts[b] = 0; s_stalls[b] = 0; s_lasttrue[b] = 1e300; }
  REMAP()
  {
    OWN_RANGE()
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      double acc_rz = 0.0, acc_bb = 0.0;
      for (long long t = ta; t < tb; ++t) {
        int px, py; long long i; bool ok;
        PIXEL_OF(t, px, py, i, ok)
        if (!ok) continue;
        double2 d = mix_init_pixel(P, i, px, py, n_all, pold);
        acc_rz += d.x; acc_bb += d.y;
      }
      block_sum2(acc_rz, acc_bb, sm_red);
      if (tid == 0) { part_b[(long long)b * G + cta] = acc_rz; part_c[(long long)b * G + cta] = acc_bb; }
    }
  }
  grid.sync();
  REDUCE_ALL(part_b, part_c, 0)
  if (tid < B) {
    const int b = tid;
    double rz = s_ta[b], bb = s_tb[b];
    s_rz[b] = rz; s_bb[b] = bb; s_maxr2[b] = bb; s_alpha[b] = 0.0; s_beta[b] = 0.0; s_rzprev[b] = rz;
    if (!(bb > 0.0) || !(rz > 0.0)) {                    // zero right-hand side: x = 0 is the solution
      if (cta == C_LO(b)) { done_g[b] = 1; iters_g[b] = 0; relres_g[b] = 0.0; }
      s_state[b] = 1;
    }
  }
  REMAP()
  int n_active = s_nact;

  int k = 0;
  for (; k < P.maxit && n_active > 0; ++k) {
    OWN_RANGE()
    // ---------------- phase A: p = z + beta p_old (on the fly), y += alpha_prev p_old, Ap = A p ----------------
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      const float beta = (float)s_beta[b], aprev = (float)s_alpha[b];
      const double aprev_d = s_alpha[b];
      const bool flush = s_flush[b] != 0;       // a reliable update happened: fold y + alpha p into the fp64 solution now
      double acc = 0.0, dummy = 0.0;
      for (long long t = ta; t < tb; t += MIX_UA) {
        int px[MIX_UA], py[MIX_UA]; long long i[MIX_UA]; bool ok[MIX_UA];
        float2 zc[MIX_UA], po[MIX_UA], zl[MIX_UA], pl[MIX_UA], zr[MIX_UA], pr[MIX_UA], zu[MIX_UA], pu[MIX_UA], zd[MIX_UA],
            pd[MIX_UA], sd[MIX_UA], swr[MIX_UA], swd[MIX_UA], swl[MIX_UA], swu[MIX_UA], yc[MIX_UA];
        float sa12[MIX_UA];
#pragma unroll
        for (int u = 0; u < MIX_UA; ++u) {
          PIXEL_OF(t + u, px[u], py[u], i[u], ok[u])
          // branch-free gather at clamped (always valid) addresses; out-of-image neighbours are replaced by the
          // centre value below (their edge weights are 0 by construction anyway)
          const long long ii = i[u];
          const long long jl = ii > 0 ? ii - 1 : 0, jr = ii + 1 < n_all ? ii + 1 : n_all - 1;
          const long long ju = ii >= W ? ii - W : 0, jd = ii + W < n_all ? ii + W : n_all - 1;
          zc[u] = z[ii]; po[u] = pold[ii];
          zl[u] = z[jl]; pl[u] = pold[jl]; zr[u] = z[jr]; pr[u] = pold[jr];
          zu[u] = z[ju]; pu[u] = pold[ju]; zd[u] = z[jd]; pd[u] = pold[jd];
          sd[u] = Df[ii]; swr[u] = WHf[ii]; swd[u] = WVf[ii];
          swl[u] = WHf[jl]; swu[u] = WVf[ju];
          sa12[u] = a12f[ii];
          yc[u] = y[ii];
        }
#pragma unroll
        for (int u = 0; u < MIX_UA; ++u) {
          if (!ok[u]) continue;
          const float2 c = make_float2(zc[u].x + beta * po[u].x, zc[u].y + beta * po[u].y);
          float2 nl = make_float2(zl[u].x + beta * pl[u].x, zl[u].y + beta * pl[u].y);
          float2 nr = make_float2(zr[u].x + beta * pr[u].x, zr[u].y + beta * pr[u].y);
          float2 nu = make_float2(zu[u].x + beta * pu[u].x, zu[u].y + beta * pu[u].y);
          float2 nd = make_float2(zd[u].x + beta * pd[u].x, zd[u].y + beta * pd[u].y);
          if (px[u] == 0) nl = c;
          if (px[u] + 1 >= W) nr = c;
          if (py[u] == 0) nu = c;
          if (py[u] + 1 >= H) nd = c;
          float au = sd[u].x * c.x + sa12[u] * c.y;
          float av = sa12[u] * c.x + sd[u].y * c.y;
          au += swr[u].x * (c.x - nr.x); av += swr[u].y * (c.y - nr.y);
          au += swl[u].x * (c.x - nl.x); av += swl[u].y * (c.y - nl.y);
          au += swd[u].x * (c.x - nd.x); av += swd[u].y * (c.y - nd.y);
          au += swu[u].x * (c.x - nu.x); av += swu[u].y * (c.y - nu.y);
          if (flush) {
            double2 xc = x[i[u]];
            x[i[u]] = make_double2(xc.x + ((double)yc[u].x + aprev_d * (double)po[u].x),
                                   xc.y + ((double)yc[u].y + aprev_d * (double)po[u].y));
            y[i[u]] = make_float2(0.f, 0.f);
          } else {
            y[i[u]] = make_float2(yc[u].x + aprev * po[u].x, yc[u].y + aprev * po[u].y);
          }
          pnew[i[u]] = c;
          Ap[i[u]] = make_float2(au, av);
          acc += (double)c.x * (double)au + (double)c.y * (double)av;
        }
      }
      block_sum2(acc, dummy, sm_red);
      if (tid == 0) part_a[(long long)b * G + cta] = acc;
    }
    grid.sync();
    REDUCE_ALL(part_a, (const double *)nullptr, 0)
    if (tid < B && s_state[tid] == 0) {
      double pap = s_ta[tid];
      s_alpha[tid] = pap > 0.0 ? s_rz[tid] / pap : 0.0;      // 0 => breakdown, resolved by the reliable update below
      s_flush[tid] = 0;
    }
    __syncthreads();
    // ---------------- phase B: r -= alpha Ap, z = M^-1 r, partial r.z and r.r ----------------
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      const float alpha = (float)s_alpha[b];
      double acc_rz = 0.0, acc_rr = 0.0;
      for (long long t = ta; t < tb; t += MIX_UB) {
        long long i[MIX_UB]; bool ok[MIX_UB];
        float2 rc[MIX_UB], ac[MIX_UB];
        float m11[MIX_UB], m12[MIX_UB], m22[MIX_UB];
#pragma unroll
        for (int u = 0; u < MIX_UB; ++u) {
          int px, py;
          PIXEL_OF(t + u, px, py, i[u], ok[u])
          rc[u] = r[i[u]]; ac[u] = Ap[i[u]];
          m11[u] = Minv[i[u]]; m12[u] = Minv[n_all + i[u]]; m22[u] = Minv[2 * n_all + i[u]];
        }
#pragma unroll
        for (int u = 0; u < MIX_UB; ++u) {
          if (!ok[u]) continue;
          float2 rn = make_float2(rc[u].x - alpha * ac[u].x, rc[u].y - alpha * ac[u].y);
          float2 zz = make_float2(m11[u] * rn.x + m12[u] * rn.y, m12[u] * rn.x + m22[u] * rn.y);
          r[i[u]] = rn; z[i[u]] = zz;
          acc_rz += (double)rn.x * (double)zz.x + (double)rn.y * (double)zz.y;
          acc_rr += (double)rn.x * (double)rn.x + (double)rn.y * (double)rn.y;
        }
      }
      block_sum2(acc_rz, acc_rr, sm_red);
      if (tid == 0) { part_b[(long long)b * G + cta] = acc_rz; part_c[(long long)b * G + cta] = acc_rr; }
    }
    grid.sync();
    REDUCE_ALL(part_b, part_c, 0)
    int rel = 0;
    if (tid < B && s_state[tid] == 0) {
      const int b = tid;
      double rz = s_ta[b], rr = s_tb[b], alpha = s_alpha[b];
      int bad = !(alpha > 0.0) || !(rz > 0.0) || !(rr == rr);
      rel = bad || rr <= P.tol2 * s_bb[b] || rr < P.delta2 * s_maxr2[b] || k + 1 == P.maxit;
      if (rel) {
        s_state[b] = 3; s_bad[b] = bad; s_rzprev[b] = s_rz[b];
      } else {
        s_beta[b] = rz / s_rz[b];
        s_rz[b] = rz;
      }
    }
    if (__syncthreads_or(rel)) {
      // ---------------- reliable update: r = b - A (x + y + alpha p) in fp64, z = M^-1 r ----------------
      for (int a = a_first; a <= a_last; ++a) {
        TILE_RANGE(a)
        if (s_state[b] != 3) continue;
        const double alpha = s_alpha[b];
        double acc_rz = 0.0, acc_rr = 0.0;
        for (long long t = ta; t < tb; ++t) {
          int px, py; long long i; bool ok;
          PIXEL_OF(t, px, py, i, ok)
          if (!ok) continue;
          double2 d = mix_reliable_pixel(P, i, px, py, n_all, pnew, alpha);
          acc_rz += d.x; acc_rr += d.y;
        }
        block_sum2(acc_rz, acc_rr, sm_red);
        if (tid == 0) { part_d[(long long)b * G + cta] = acc_rz; part_e[(long long)b * G + cta] = acc_rr; }
      }
      grid.sync();
      REDUCE_ALL(part_d, part_e, 3)
      int fin = 0;
      if (tid < B && s_state[tid] == 3) {
        const int b = tid;
        double rz = s_ta[b], rr = s_tb[b];
        int conv = rr <= P.tol2 * s_bb[b];
        // An fp32 breakdown (p.Ap <= 0, r.z <= 0: the iterated quantities have lost their meaning close to the fp32
        // floor) is not the end: the fp64 residual and z = M^-1 r just computed are sound, so CG RESTARTS from them
        // (beta = 0).  Given up after 8 restarts, or when three replacements in a row failed to halve the true residual
        // (the fp64 floor of an ill-conditioned system: attainable accuracy reached).
        const int fatal = !(rz > 0.0) || !(rr == rr) || k + 1 == P.maxit;
        if (s_bad[b]) s_restarts[b] += 1;
        s_stalls[b] = rr > 0.25 * s_lasttrue[b] ? s_stalls[b] + 1 : 0;
        s_lasttrue[b] = rr;
        if (conv || fatal || s_restarts[b] > 8 || s_stalls[b] >= 3) {
          s_state[b] = 2;
          fin = 1;
          if (cta == C_LO(b)) {
            // a solve that stagnates within a factor 4 of the target has reached what fp64 can deliver for this system
            // (seen on 17 x 30 Lorentzian levels: 1.2e-12 against 1e-12); its relres is reported as it is
            const int floor_ok = s_stalls[b] >= 3 && rr <= 16.0 * P.tol2 * s_bb[b];
            done_g[b] = (conv || floor_ok) ? 1 : (k + 1 == P.maxit && !s_bad[b] ? 3 : 2);
            iters_g[b] = k + 1;
            relres_g[b] = sqrt(rr / s_bb[b]);
          }
        } else {
          s_state[b] = 0;
          s_flush[b] = 1;                                  // x += y + alpha_k p_k rides on the next phase A
          s_maxr2[b] = rr;
          s_beta[b] = s_bad[b] ? 0.0 : rz / s_rzprev[b];   // restart after a breakdown: p = z
          s_rz[b] = rz;
        }
      }
      if (__syncthreads_or(fin)) {
        // systems that just finished: fold the pending y + alpha p into x (own pixels only), then re-deal the tiles
        for (int a = a_first; a <= a_last; ++a) {
          TILE_RANGE(a)
          if (s_state[b] != 2) continue;
          const double alpha = s_alpha[b];
          for (long long t = ta; t < tb; ++t) {
            int px, py; long long i; bool ok;
            PIXEL_OF(t, px, py, i, ok)
            if (!ok) continue;
            double2 xc = x[i];
            float2 yc = y[i], pc = pnew[i];
            x[i] = make_double2(xc.x + ((double)yc.x + alpha * (double)pc.x), xc.y + ((double)yc.y + alpha * (double)pc.y));
          }
        }
        __syncthreads();
        if (tid < B && s_state[tid] == 2) s_state[tid] = 1;
        REMAP()
        n_active = s_nact;
      }
    }
    { float2 *t = pold; pold = pnew; pnew = t; }
  }
#undef REMAP
#undef C_LO
#undef C_HI
#undef OWN_RANGE
#undef REDUCE_ALL
#undef TILE_RANGE
#undef PIXEL_OF
}

// ------------------------------------------------------------------------------------------------
// Legacy solver='sor' (base.py:138-172): Gauss-Seidel SOR, omega 1.9, in the reference's unknown order (all u
// column-major, then all v column-major), from x = 0, until ||x - x_old|| < tol ||x|| or max_iters sweeps.
// A cell (y, x) depends only on the already-updated cells (y-1, x) and (y, x-1) of its own component, so the cells of
// one anti-diagonal x + y = d are independent: one CTA per system sweeps the H + W - 1 anti-diagonals in order, which
// reproduces the reference's LEXICOGRAPHIC iterates (and therefore its stopping sweep) instead of the different
// iterates a red-black ordering would give.  Latency-bound by design (one block barrier per anti-diagonal); this is
// the reference's approximate legacy mode, kept for `params={'solver': 'sor'}`, not a fast path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) sor_kernel(LinSys S, double2 *x2, double omega, double tol, int max_iters,
                                                   int *flags, double *relres_g) {
  const int b = blockIdx.x, B = S.B, H = S.H, W = S.W;
  const long long HW = (long long)H * W, base = (long long)b * HW;
  double *x = reinterpret_cast<double *>(x2);
  __shared__ double red[2][32];
  __shared__ int s_stop;
  for (long long i = threadIdx.x; i < HW; i += blockDim.x) x2[base + i] = make_double2(0.0, 0.0);
  __syncthreads();
  int it = 0;
  double rel = 0.0;
  for (; it < max_iters; ++it) {
    double dx2 = 0.0, xx2 = 0.0;
    for (int comp = 0; comp < 2; ++comp) {
      for (int d = 0; d < H + W - 1; ++d) {
        const int y_lo = d - W + 1 > 0 ? d - W + 1 : 0, y_hi = d < H - 1 ? d : H - 1;
        for (int c = threadIdx.x; c <= y_hi - y_lo; c += blockDim.x) {
          const int y = y_lo + c, xq = d - y;
          const long long i = base + (long long)y * W + xq;
          const double2 dd = S.D[i], wr = S.WH[i], wd = S.WV[i];
          const double2 wl = xq > 0 ? S.WH[i - 1] : make_double2(0.0, 0.0);
          const double2 wu = y > 0 ? S.WV[i - W] : make_double2(0.0, 0.0);
          const double a = comp ? dd.y : dd.x, kr = comp ? wr.y : wr.x, kd = comp ? wd.y : wd.x;
          const double kl = comp ? wl.y : wl.x, ku = comp ? wu.y : wu.x;
          const double f_old = x[2 * i + comp], other = x[2 * i + (1 - comp)];
          double sig = S.a12[i] * other;
          if (xq > 0) sig -= kl * x[2 * (i - 1) + comp];
          if (y > 0) sig -= ku * x[2 * (i - W) + comp];
          if (y + 1 < H) sig -= kd * x[2 * (i + W) + comp];
          if (xq + 1 < W) sig -= kr * x[2 * (i + 1) + comp];
          const double dg = a + (((kr + kd) + kl) + ku);
          double f_new = f_old;
          if (fabs(dg) >= 1e-15) {
            const double rhs = comp ? S.rhs[i].y : S.rhs[i].x;
            f_new = (1.0 - omega) * f_old + omega * (rhs - sig) / dg;
            x[2 * i + comp] = f_new;
          }
          dx2 += (f_new - f_old) * (f_new - f_old);
          xx2 += f_new * f_new;
        }
        __syncthreads();
      }
    }
    // ||x - x_old|| and ||x|| over the system
    dx2 = warp_sum(dx2);
    xx2 = warp_sum(xx2);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = dx2; red[1][w] = xx2; }
    __syncthreads();
    if (w == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      dx2 = l < nw ? red[0][l] : 0.0;
      xx2 = l < nw ? red[1][l] : 0.0;
      dx2 = warp_sum(dx2);
      xx2 = warp_sum(xx2);
      if (l == 0) {
        s_stop = sqrt(dx2) < tol * sqrt(xx2);
        red[0][0] = xx2 > 0.0 ? sqrt(dx2 / xx2) : 0.0;
      }
    }
    __syncthreads();
    rel = red[0][0];
    const int stop = s_stop;
    __syncthreads();
    if (stop) { ++it; break; }
  }
  if (threadIdx.x == 0) {
    flags[1 + b] = it < max_iters || rel < tol ? 1 : 3;
    flags[1 + B + b] = it;
    relres_g[b] = rel;
  }
}

// tiny epilogue: fold the per-system outcome of one solve into the running device statistics
__global__ void pcg_stats_kernel(const int *flags, int B, long long hw, long long *stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long nc = 0, it = 0, mx = 0;
    for (int b = 0; b < B; ++b) {
      if (flags[1 + b] != 1) nc++;
      it += flags[1 + B + b];
      if (flags[1 + B + b] > mx) mx = flags[1 + B + b];
    }
    stats[0] += mx;    // iterations of this solve = slowest system of the batch
    stats[1] += nc;    // systems that did not reach tol
    stats[2] += it;    // sum over systems
    stats[3] += it * hw;   // pixel-iterations (roofline accounting)
  }
}

size_t pcg_work_bytes(const b200flow_ctx *ctx, int B, int H, int W) {
  size_t n = (size_t)B * H * W;
  return n * (5 * sizeof(double2) + 3 * sizeof(float) + sizeof(float4) + sizeof(unsigned)) + 4 * (size_t)B * ctx->num_sms * 8 * 8 + 64 * B + 4096;
}

static int pcg_grid(b200flow_ctx *ctx, int *grid_out, int *grid_mixed_out) {
  if (ctx->grid_pcg == 0) {          // per context: distinct contexts may live on different devices / host threads
    int nb = 0, nm = 0;
    BF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pcg_kernel, PCG_THREADS, 0));
    BF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nm, pcg_mixed_kernel, PCG_THREADS, 0));
    if (nb < 1 || nm < 1) return set_err(ctx, B200FLOW_ECUDA, "pcg kernels cannot be made resident");
    ctx->grid_pcg = nb * ctx->num_sms;
    ctx->grid_mixed = nm * ctx->num_sms;
  }
  *grid_out = ctx->grid_pcg;
  *grid_mixed_out = ctx->grid_mixed;
  if (ctx->solver_ctas_per_sm > 0) {  // split context: leave room for the sibling groups' kernels
    const int cap = ctx->solver_ctas_per_sm * ctx->num_sms;
    if (*grid_out > cap) *grid_out = cap;
    if (*grid_mixed_out > cap) *grid_mixed_out = cap;
  }
  return 0;
}

int pcg_work_alloc(b200flow_ctx *ctx, int B, int H, int W, PcgWork *w) {
  size_t n = (size_t)B * H * W;
  int G, Gm;
  BF_TRY(pcg_grid(ctx, &G, &Gm));
  w->grid = G;
  w->grid_mixed = Gm;
  BF_TRY(pcg_ic_grid(ctx, &w->grid_ic));
  if (Gm > G) G = Gm;
  if (w->grid_ic > G) G = w->grid_ic;
  BF_TRY(arena_alloc(ctx, &w->r, n));
  BF_TRY(arena_alloc(ctx, &w->p, n));
  BF_TRY(arena_alloc(ctx, &w->p2, n));
  BF_TRY(arena_alloc(ctx, &w->z, n));
  BF_TRY(arena_alloc(ctx, &w->Ap, n));
  BF_TRY(arena_alloc(ctx, &w->Minv, 3 * n));
  BF_TRY(arena_alloc(ctx, &w->ic_c0, n));
  BF_TRY(arena_alloc(ctx, &w->ic_cw, n));
  BF_TRY(arena_alloc(ctx, &w->partial, (size_t)5 * B * G));   // five slices; G = the largest of the three kernels' grids
  BF_TRY(arena_alloc(ctx, &w->scal, (size_t)B));
  BF_TRY(arena_alloc(ctx, &w->flags, (size_t)(1 + 2 * B)));
  return 0;
}

int k_pcg_solve(b200flow_ctx *ctx, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int mode,
                int *iters_host, double *relres_host, bool sync_results) {
  const bool mixed = mode == PCG_MODE_MIXED || mode == PCG_MODE_MIXED_IC || mode == PCG_MODE_FP32_IC;
  const int tiles_x = (int)cdiv(sys.W, TILE_W), tiles_y = (int)cdiv(sys.H, TILE_H);
  const int tiles_per_sys = tiles_x * tiles_y;
  const long long total_tiles = (long long)tiles_per_sys * sys.B;
  int G = mixed ? w.grid_mixed : w.grid;
  if ((long long)G > total_tiles) G = (int)total_tiles;
  if (G < 1) G = 1;
  const int tiles_per_cta = (int)cdiv(total_tiles, G);
  BF_CUDA(ctx, cudaMemsetAsync(w.flags, 0, sizeof(int) * (1 + 2 * sys.B), ctx->stream));
  if (mode == PCG_MODE_SOR) {
    int len = sys.H < sys.W ? sys.H : sys.W;
    int threads = len >= 1024 ? 1024 : ((len + 31) / 32) * 32;
    BF_LAUNCH(ctx, sor_kernel, sys.B, threads, 0, sys, x, 1.9, tol, maxit, w.flags, w.scal);
  } else if (mixed) {
    if (sys.B > MAXB)
      return set_err(ctx, B200FLOW_EINVAL, "batch of %d systems exceeds %d per solve; split the batch", sys.B, MAXB);
    MixParams P;
    P.sys = sys; P.w = w; P.x = x;
    P.tol2 = tol * tol;
    {
      double delta = mode == PCG_MODE_MIXED_IC ? PCG_RELIABLE_DELTA_IC : PCG_RELIABLE_DELTA;
      if (mode == PCG_MODE_FP32_IC) delta = 0.0;             // no residual replacement on the way
#ifdef B200FLOW_TUNING                                        // tuning builds only (scripts/build_variant.sh ... -DB200FLOW_TUNING)
      if (const char *dl = getenv("B200FLOW_RELIABLE_DELTA")) delta = atof(dl);
#endif
      P.delta2 = delta * delta;
    }
    P.maxit = maxit;
    P.tiles_x = tiles_x; P.tiles_y = tiles_y; P.tiles_per_sys = tiles_per_sys;
    P.debug = 0;
    P.fp32_only = mode == PCG_MODE_FP32_IC;
    P.w.grid = G;
    // the fp32 working set (76 B / pixel) is carved out of the five fp64 vectors (80 B / pixel) of the work area
    const size_t n = (size_t)sys.B * sys.H * sys.W;
    P.m.r = reinterpret_cast<float2 *>(w.r);   P.m.z = P.m.r + n;
    P.m.p = reinterpret_cast<float2 *>(w.p);   P.m.p2 = P.m.p + n;
    P.m.Ap = reinterpret_cast<float2 *>(w.p2); P.m.y = P.m.Ap + n;
    P.m.D = reinterpret_cast<float2 *>(w.z);   P.m.WH = P.m.D + n;
    P.m.WV = reinterpret_cast<float2 *>(w.Ap); P.m.a12 = reinterpret_cast<float *>(P.m.WV + n);
    const bool ic = mode == PCG_MODE_MIXED_IC || mode == PCG_MODE_FP32_IC;
    if (mode == PCG_MODE_MIXED_IC && ctx->band.world > 1 && sys.B == 1 &&
        (long long)sys.H * sys.W >= ctx->band.min_pixels && (sys.H + 7) / 8 >= ctx->band.world) {
      // row-band mode: this rank iterates on its rows only, then the solution bands are exchanged (solve_ic.cu)
      BF_TRY(k_pcg_ic_band_launch(ctx, P, w.grid_ic));
      ctx->launches++;
      BF_TRY(k_band_exchange_x(ctx, x, sys.H, sys.W));
      ctx->launches--;                 // (the common exit below counts the solver launch)
    } else if (ic) {
      BF_TRY(k_pcg_ic_launch(ctx, P, w.grid_ic));
    } else {
      void *args[] = {&P};
      int dyn = 0;
#ifdef B200FLOW_TUNING   // experiment behind profiles/r01_summary.md "L1 matters for phase A": shrink L1 with an unused dynamic carve-out
      if (const char *dbg = getenv("B200FLOW_MIX_SMEM")) dyn = atoi(dbg);
      if (dyn > 0) {
        cudaFuncSetAttribute(pcg_mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        const char *cv = getenv("B200FLOW_MIX_CARVEOUT");
        cudaFuncSetAttribute(pcg_mixed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared);
      }
#endif
      if (ctx->plain_solver_launch) return set_err(ctx, B200FLOW_EINVAL, "concurrent sub-batches need solver_precision='mixed' (pcg_ic_kernel)");
      BF_CUDA(ctx, cudaLaunchCooperativeKernel((void *)pcg_mixed_kernel, dim3(G), dim3(PCG_THREADS), args, dyn, ctx->stream));
    }
  } else {
    PcgParams P;
    P.sys = sys; P.w = w; P.x = x;
    P.tol2 = tol * tol;
    P.maxit = maxit;
    P.scalar_jacobi = mode == PCG_MODE_JACOBI_F64;
    P.tiles_x = tiles_x; P.tiles_y = tiles_y; P.tiles_per_sys = tiles_per_sys;
    P.total_tiles = total_tiles; P.tiles_per_cta = tiles_per_cta;
    // keep the number of systems one CTA may touch within the shared-memory table
    if (P.tiles_per_cta / P.tiles_per_sys + 2 > MAXLOC)
      return set_err(ctx, B200FLOW_EINVAL, "batch of %d systems of %dx%d is too fine-grained for one solve; split the batch",
                     sys.B, sys.H, sys.W);
    P.w.grid = G;
    void *args[] = {&P};
    if (ctx->plain_solver_launch) return set_err(ctx, B200FLOW_EINVAL, "concurrent sub-batches need solver_precision='mixed' (pcg_ic_kernel)");
    BF_CUDA(ctx, cudaLaunchCooperativeKernel((void *)pcg_kernel, dim3(G), dim3(PCG_THREADS), args, 0, ctx->stream));
  }
  if (mode != PCG_MODE_SOR) ctx->launches++;
  if (sync_results) {
    std::vector<int> fl(1 + 2 * sys.B);
    std::vector<double> rr(sys.B);
    if (ctx->band.world > 1) BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // see download() in common.cuh
    BF_CUDA(ctx, cudaMemcpyAsync(fl.data(), w.flags, sizeof(int) * fl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA(ctx, cudaMemcpyAsync(rr.data(), w.scal, sizeof(double) * sys.B, cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = 0;
    for (int b = 0; b < sys.B; ++b) {
      if (iters_host) iters_host[b] = fl[1 + sys.B + b];
      if (relres_host) relres_host[b] = rr[b];
      if (fl[1 + b] != 1) rc = B200FLOW_ENOCONV;
    }
    if (rc) set_err(ctx, rc, "%s stopped before reaching tol=%g (maxit=%d)", mode == PCG_MODE_SOR ? "SOR" : "PCG", tol, maxit);
    return rc;
  }
  return 0;
}

// pipeline variant: no host sync; outcome accumulated into device statistics
int k_pcg_solve_async(b200flow_ctx *ctx, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int mode,
                      long long *stats_dev) {
  BF_TRY(k_pcg_solve(ctx, sys, w, x, tol, maxit, mode, nullptr, nullptr, false));
  if (stats_dev) BF_LAUNCH(ctx, pcg_stats_kernel, 1, 32, 0, w.flags, sys.B, (long long)sys.H * sys.W, stats_dev);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// A @ x and diag(A) (parity tests of the assembled operator; not on the hot path)
// ------------------------------------------------------------------------------------------------
__global__ void operator_apply_kernel(LinSys S, const double2 *x, double2 *Ax, double2 *diag) {
  int px = blockIdx.x * blockDim.x + threadIdx.x;
  int py = blockIdx.y * blockDim.y + threadIdx.y;
  if (px >= S.W || py >= S.H) return;
  long long off = (long long)blockIdx.z * S.H * S.W;
  long long i = off + (long long)py * S.W + px;
  Stencil s = load_stencil(S, i, px, py);
  if (Ax) Ax[i] = apply_stencil(s, x, i, px, py, S.H, S.W);
  if (diag) diag[i] = make_double2(s.d.x + s.wr.x + s.wl.x + s.wd.x + s.wu.x, s.d.y + s.wr.y + s.wl.y + s.wd.y + s.wu.y);
}

int k_operator_apply(b200flow_ctx *ctx, LinSys sys, const double2 *x, double2 *Ax, double2 *diag) {
  dim3 blk(32, 8), grd((unsigned)cdiv(sys.W, 32), (unsigned)cdiv(sys.H, 8), sys.B);
  BF_LAUNCH(ctx, operator_apply_kernel, grd, blk, 0, sys, x, Ax, diag);
  return 0;
}

}  // namespace bf
