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
#include <cooperative_groups.h>
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace bf {

constexpr int PCG_THREADS = 256;
constexpr int TILE_W = 32, TILE_H = 8;
constexpr int MAXLOC = 32;       // systems one CTA may touch

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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two accumulators; result valid in thread 0
__device__ __forceinline__ void block_sum2(double &a, double &b, double (*sm)[PCG_THREADS / 32]) {
  a = warp_sum(a);
  b = warp_sum(b);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // protect sm reuse
  if (l == 0) { sm[0][w] = a; sm[1][w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < PCG_THREADS / 32 ? sm[0][l] : 0.0;
    b = l < PCG_THREADS / 32 ? sm[1][l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// fixed-order reduction of the per-CTA partials of system b; executed by warp 0, result in all its lanes
__device__ __forceinline__ double reduce_partials(const double *part, int G, int b, int c_lo, int c_hi) {
  const volatile double *q = part + (long long)b * G;
  double s = 0.0;
  for (int c = c_lo + (threadIdx.x & 31); c <= c_hi; c += 32) s += q[c];
  return warp_sum(s);
}

struct Stencil {
  double2 d; double a12; double2 wr, wl, wd, wu;   // own right/down edges, left/up neighbours' edges
};

__device__ __forceinline__ Stencil load_stencil(const LinSys &S, long long i, int x, int y) {
  Stencil s;
  s.d = __ldg(&S.D[i]);
  s.a12 = __ldg(&S.a12[i]);
  s.wr = __ldg(&S.WH[i]);
  s.wd = __ldg(&S.WV[i]);
  s.wl = x > 0 ? __ldg(&S.WH[i - 1]) : make_double2(0.0, 0.0);
  s.wu = y > 0 ? __ldg(&S.WV[i - S.W]) : make_double2(0.0, 0.0);
  return s;
}

__device__ __forceinline__ double2 apply_stencil(const Stencil &s, const double2 *v, long long i, int x, int y, int H,
                                                 int W) {
  double2 c = v[i];
  double au = s.d.x * c.x + s.a12 * c.y;
  double av = s.a12 * c.x + s.d.y * c.y;
  if (x + 1 < W) { double2 n = v[i + 1]; au += s.wr.x * (c.x - n.x); av += s.wr.y * (c.y - n.y); }
  if (x > 0)     { double2 n = v[i - 1]; au += s.wl.x * (c.x - n.x); av += s.wl.y * (c.y - n.y); }
  if (y + 1 < H) { double2 n = v[i + W]; au += s.wd.x * (c.x - n.x); av += s.wd.y * (c.y - n.y); }
  if (y > 0)     { double2 n = v[i - W]; au += s.wu.x * (c.x - n.x); av += s.wu.y * (c.y - n.y); }
  return make_double2(au, av);
}

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
  return n * (5 * sizeof(double2) + 3 * sizeof(float)) + 4 * (size_t)B * ctx->num_sms * 8 * 8 + 64 * B + 4096;
}

static int pcg_grid(b200flow_ctx *ctx, int *grid_out) {
  static int cached_dev = -1, cached = 0;
  if (cached_dev != ctx->device) {
    int nb = 0;
    BF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pcg_kernel, PCG_THREADS, 0));
    if (nb < 1) return set_err(ctx, B200FLOW_ECUDA, "pcg_kernel cannot be made resident");
    cached = nb * ctx->num_sms;
    cached_dev = ctx->device;
  }
  *grid_out = cached;
  return 0;
}

int pcg_work_alloc(b200flow_ctx *ctx, int B, int H, int W, PcgWork *w) {
  size_t n = (size_t)B * H * W;
  int G;
  BF_TRY(pcg_grid(ctx, &G));
  w->grid = G;
  BF_TRY(arena_alloc(ctx, &w->r, n));
  BF_TRY(arena_alloc(ctx, &w->p, n));
  BF_TRY(arena_alloc(ctx, &w->p2, n));
  BF_TRY(arena_alloc(ctx, &w->z, n));
  BF_TRY(arena_alloc(ctx, &w->Ap, n));
  BF_TRY(arena_alloc(ctx, &w->Minv, 3 * n));
  BF_TRY(arena_alloc(ctx, &w->partial, (size_t)4 * B * G));
  BF_TRY(arena_alloc(ctx, &w->scal, (size_t)B));
  BF_TRY(arena_alloc(ctx, &w->flags, (size_t)(1 + 2 * B)));
  return 0;
}

int k_pcg_solve(b200flow_ctx *ctx, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int scalar_jacobi,
                int *iters_host, double *relres_host, bool sync_results) {
  PcgParams P;
  P.sys = sys; P.w = w; P.x = x;
  P.tol2 = tol * tol;
  P.maxit = maxit;
  P.scalar_jacobi = scalar_jacobi;
  P.tiles_x = (int)cdiv(sys.W, TILE_W);
  P.tiles_y = (int)cdiv(sys.H, TILE_H);
  P.tiles_per_sys = P.tiles_x * P.tiles_y;
  P.total_tiles = (long long)P.tiles_per_sys * sys.B;
  int G = w.grid;
  if ((long long)G > P.total_tiles) G = (int)P.total_tiles;
  if (G < 1) G = 1;
  P.tiles_per_cta = (int)cdiv(P.total_tiles, G);
  // keep the number of systems one CTA may touch within the shared-memory table
  if (P.tiles_per_cta / P.tiles_per_sys + 2 > MAXLOC)
    return set_err(ctx, B200FLOW_EINVAL, "batch of %d systems of %dx%d is too fine-grained for one solve; split the batch",
                   sys.B, sys.H, sys.W);
  P.w.grid = G;
  BF_CUDA(ctx, cudaMemsetAsync(w.flags, 0, sizeof(int) * (1 + 2 * sys.B), ctx->stream));
  void *args[] = {&P};
  BF_CUDA(ctx, cudaLaunchCooperativeKernel((void *)pcg_kernel, dim3(G), dim3(PCG_THREADS), args, 0, ctx->stream));
  ctx->launches++;
  if (sync_results) {
    std::vector<int> fl(1 + 2 * sys.B);
    std::vector<double> rr(sys.B);
    BF_CUDA(ctx, cudaMemcpyAsync(fl.data(), w.flags, sizeof(int) * fl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA(ctx, cudaMemcpyAsync(rr.data(), w.scal, sizeof(double) * sys.B, cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = 0;
    for (int b = 0; b < sys.B; ++b) {
      if (iters_host) iters_host[b] = fl[1 + sys.B + b];
      if (relres_host) relres_host[b] = rr[b];
      if (fl[1 + b] != 1) rc = B200FLOW_ENOCONV;
    }
    if (rc) set_err(ctx, rc, "PCG stopped before reaching tol=%g (maxit=%d)", tol, maxit);
    return rc;
  }
  return 0;
}

// pipeline variant: no host sync; outcome accumulated into device statistics
int k_pcg_solve_async(b200flow_ctx *ctx, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int scalar_jacobi,
                      long long *stats_dev) {
  BF_TRY(k_pcg_solve(ctx, sys, w, x, tol, maxit, scalar_jacobi, nullptr, nullptr, false));
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
