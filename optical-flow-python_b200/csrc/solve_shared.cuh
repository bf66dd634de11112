// solve_shared.cuh -- definitions shared by the persistent PCG kernels (solve.cu, solve_ic.cu)
#pragma once
#include <cooperative_groups.h>
#include "kernels.cuh"

namespace bf {

constexpr int PCG_THREADS = 256;
constexpr int TILE_W = 32, TILE_H = 8;
constexpr int MAXLOC = 32;       // systems one CTA may touch
constexpr double PCG_RELIABLE_DELTA = 0.01;  // mixed precision: fp64 residual replacement when |r| fell 100x
// IC-preconditioned kernel: sqrt(eps_fp32) -- the usual reliable-update threshold; a replacement costs ~1.5 iterations of
// traffic and the IC solves need only 25-200 iterations (measured: 0.01 -> 157.6 ms, 1e-3 -> 154.1, 1e-4 -> 148.7 ms per bench step)
constexpr double PCG_RELIABLE_DELTA_IC = 2.5e-4;


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

constexpr int MAXB = 128;        // systems per mixed-precision solve (one scalar-update thread per system)

struct MixWork {
  float2 *r, *z, *p, *p2, *Ap, *y;
  float2 *D, *WH, *WV;
  float *a12;
};

struct MixParams {
  LinSys sys;
  PcgWork w;
  MixWork m;
  double2 *x;
  double tol2, delta2;
  int maxit;
  int tiles_x, tiles_y, tiles_per_sys;
  int debug;                     // tuning builds (-DB200FLOW_TUNING) only; always 0 on the product path
  int fp32_only;                 // fp32 variant: no residual replacement, convergence on the iterated fp32 residual
};

// solve_ic.cu
int pcg_ic_grid(b200flow_ctx *ctx, int *grid_out);
int k_pcg_ic_launch(b200flow_ctx *ctx, MixParams P, int grid_max);
int k_pcg_ic_band_launch(b200flow_ctx *ctx, MixParams P, int grid_max);   // row-band mode: this rank's rows of one system
int k_band_exchange_x(b200flow_ctx *ctx, double2 *x, int H, int W);       // all-gather of the solution bands over P2P stores
int k_band_error(b200flow_ctx *ctx, unsigned long long *err_host);

}  // namespace bf
