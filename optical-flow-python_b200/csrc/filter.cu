// filter.cu -- kernel group (5): median filters of the flow, occlusion confidence, and the colour- and
// occlusion-weighted median of the Classic+NL non-local term.  Replaces scipy.ndimage.median_filter call
// sites (hs.py:96-97,139-140; ba.py:198-199), occlusion.py:6-56 and weighted_median.py:5-112.
// Outputs of both medians are always one of the window's input samples, selected with compare/exchange
// networks in registers (no arithmetic on the samples), so selections are bit-exact.
#include <mutex>
#include "kernels.cuh"

namespace bf {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double clip1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

__device__ __forceinline__ void ce(double &a, double &b) {   // compare-exchange: a <= b afterwards
  double lo = fmin(a, b), hi = fmax(a, b);
  a = lo; b = hi;
}

// moves the minimum of v[0..K) to v[0] and the maximum to v[K-1]
template <int K, int CAP>
__device__ __forceinline__ void minmax_ends(double (&v)[CAP]) {
#pragma unroll
  for (int i = 0; i < K / 2; ++i) ce(v[i], v[K - 1 - i]);
#pragma unroll
  for (int i = 1; i < (K + 1) / 2; ++i) ce(v[0], v[i]);
#pragma unroll
  for (int i = K / 2; i < K - 1; ++i) ce(v[i], v[K - 1]);
}

// "forgetful selection" of the median of N = 2m+1 samples: keep m+2 candidates, repeatedly drop the
// extremes and take in one more sample.  ~1.5 K compare-exchanges per step, all in registers.
template <int K, int CAP, typename Next>
__device__ __forceinline__ double forget_select(double (&v)[CAP], Next next, int taken) {
  minmax_ends<K, CAP>(v);
  if constexpr (K == 3) {
    return v[1];
  } else {
    v[0] = v[K - 2];          // compact: drop min (slot 0) and max (slot K-1)
    v[K - 2] = next(taken);
    return forget_select<K - 1, CAP>(v, next, taken + 1);
  }
}

// ------------------------------------------------------------------------------------------------
// KxK median of u and v with scipy 'reflect' boundary.  Candidate field = base (+ clip(x)).
//   assign_direct = 0: out = base + (median - base)   (ba.py:186-204: duv = filtered - uv0; uv = uv0 + duv)
//   assign_direct = 1: out = median                    (hs.py:134-140, 95-97)
//   active[b] == 0  : out = base                       (Horn-Schunck early exit, hs.py:126-127)
// algorithmic bytes per pixel: read 16 (+16 for x) , write 16
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) median_kernel(const double2 *__restrict__ base, const double2 *__restrict__ x,
                                                     int limit_update, const int *__restrict__ active,
                                                     double2 *__restrict__ out, int H, int W, int assign_direct) {
  constexpr int R = K / 2, TW = 32, TH = 8, SW = TW + 2 * R, SH = TH + 2 * R, N = K * K, CAP = N / 2 + 2;
  __shared__ double2 tile[SH * SW];
  const int b = blockIdx.z;
  const long long off = (long long)b * H * W;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const bool act = active ? active[b] != 0 : true;
  for (int t = threadIdx.x; t < SH * SW; t += 256) {
    int sy = t / SW, sx = t % SW;
    int gy = reflect_idx(y0 + sy - R, H), gx = reflect_idx(x0 + sx - R, W);
    long long gi = off + (long long)gy * W + gx;
    double2 v = base[gi];
    if (x && act) {
      double2 d = x[gi];
      if (limit_update) { d.x = clip1(d.x); d.y = clip1(d.y); }
      v.x += d.x; v.y += d.y;
    }
    tile[t] = v;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int px = x0 + lx, py = y0 + ly;
  if (px >= W || py >= H) return;
  long long gi = off + (long long)py * W + px;
  double2 bs = base[gi];
  if (!act) { out[gi] = bs; return; }
  double2 med;
#pragma unroll
  for (int comp = 0; comp < 2; ++comp) {
    auto sample = [&](int e) -> double {
      int dy = e / K, dx = e - dy * K;
      double2 s = tile[(ly + dy) * SW + lx + dx];
      return comp == 0 ? s.x : s.y;
    };
    double v[CAP];
#pragma unroll
    for (int e = 0; e < CAP; ++e) v[e] = sample(e);
    double m = forget_select<CAP, CAP>(v, sample, CAP);
    if (comp == 0) med.x = m; else med.y = m;
  }
  if (assign_direct) out[gi] = med;
  else out[gi] = make_double2(bs.x + (med.x - bs.x), bs.y + (med.y - bs.y));
}

int k_median_uv(b200flow_ctx *ctx, const double2 *base, const double2 *x, int limit_update, const int *active,
                double2 *out, int B, int H, int W, int kh, int kw, int assign_direct) {
  if (kh != kw || (kh != 3 && kh != 5 && kh != 7))
    return set_err(ctx, B200FLOW_EINVAL, "median_filter_size [%d,%d] unsupported (square 3, 5 or 7)", kh, kw);
  dim3 blk(256), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
  if (kh == 3) BF_LAUNCH(ctx, median_kernel<3>, grd, blk, 0, base, x, limit_update, active, out, H, W, assign_direct);
  else if (kh == 5) BF_LAUNCH(ctx, median_kernel<5>, grd, blk, 0, base, x, limit_update, active, out, H, W, assign_direct);
  else BF_LAUNCH(ctx, median_kernel<7>, grd, blk, 0, base, x, limit_update, active, out, H, W, assign_direct);
  return 0;
}

// out = uv + clip(x)   (classic_nl.py:252-261), gated per batch item
__global__ void clip_add_kernel(const double2 *__restrict__ uv, const double2 *__restrict__ x, int limit_update,
                                const int *__restrict__ active, double2 *__restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long gi = (long long)blockIdx.y * n + i;
  double2 v = uv[gi];
  if (!active || active[blockIdx.y]) {
    double2 d = x[gi];
    if (limit_update) { d.x = clip1(d.x); d.y = clip1(d.y); }
    v.x += d.x; v.y += d.y;
  }
  out[gi] = v;
}

int k_clip_add(b200flow_ctx *ctx, const double2 *uv, const double2 *x, int limit_update, const int *active,
               double2 *out, long long n_per_item, int B) {
  BF_LAUNCH(ctx, clip_add_kernel, dim3((unsigned)cdiv(n_per_item, 256), B), 256, 0, uv, x, limit_update, active, out,
            n_per_item);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Horn-Schunck early exit: active[b] &= (||x_b||_2 >= 1e-3)   (hs.py:126-127), two-stage fixed-order sum
// ------------------------------------------------------------------------------------------------
constexpr int NORM_BLOCKS = 64;

__global__ void norm_partial_kernel(const double2 *__restrict__ x, long long n, double *__restrict__ part) {
  const double2 *p = x + (long long)blockIdx.y * n;
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 v = p[i];
    s += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    part[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void norm_gate_kernel(const double *__restrict__ part, int nblk, int *__restrict__ active) {
  int b = blockIdx.x;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nblk; ++i) t += part[b * nblk + i];
    if (sqrt(t) < 1e-3) active[b] = 0;
  }
}

int k_hs_norm_gate(b200flow_ctx *ctx, const double2 *x, int B, long long n, int *active, double *scratch) {
  BF_LAUNCH(ctx, norm_partial_kernel, dim3(NORM_BLOCKS, B), 256, 0, x, n, scratch);
  BF_LAUNCH(ctx, norm_gate_kernel, B, 32, 0, scratch, NORM_BLOCKS, active);
  return 0;
}

// display=True log (classic_nl.py:255-256, ba.py:189-190, hs.py:123-124): ||clip(x) - duv||_2^2 of the first pair of the batch,
// one block, fixed order; only launched when the per-iteration log is switched on (b200flow_ctx_set_log)
__global__ void __launch_bounds__(1024) delta_norm_kernel(const double2 *__restrict__ x, const double2 *__restrict__ sub,
                                                           int clip, long long n, double *__restrict__ out) {
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    double2 v = x[i];
    if (clip) { v.x = clip1(v.x); v.y = clip1(v.y); }
    if (sub) { const double2 d = sub[i]; v.x -= d.x; v.y -= d.y; }
    s += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += sm[w];
    *out = t;
  }
}

int k_delta_norm(b200flow_ctx *ctx, const double2 *x, const double2 *sub, int clip, long long n, double *out) {
  BF_LAUNCH(ctx, delta_norm_kernel, 1, 1024, 0, x, sub, clip, n, out);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// occlusion confidence (occlusion.py:6-56)      bytes/pixel: read uv 16 + im1 8 + im2 8, write 8 = 40
// ------------------------------------------------------------------------------------------------
// frames: [B][2*NC][H][W]; multi-channel frames average |warp - frame 1| over the channels (occlusion.py:47-54)
__global__ void occlusion_kernel(const double2 *__restrict__ uv, const double *__restrict__ frames, long long bstride,
                                 int NC, int H, int W, double sigma_d, double sigma_i, double *__restrict__ occ) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const long long HW = (long long)H * W;
  long long off = (long long)blockIdx.z * HW;
  uv += off; frames += (long long)blockIdx.z * bstride;
  long long i = (long long)y * W + x;
  double2 f = uv[i];
  double div = 0.0;
  if (x > 0) div += f.x - uv[i - 1].x;
  if (y > 0) div += f.y - uv[i - W].y;
  double x2 = (double)x + f.x, y2 = (double)y + f.y;
  double wm = (double)(W - 1), hm = (double)(H - 1);
  x2 = x2 < 0.0 ? 0.0 : (x2 > wm ? wm : x2);      // map_coordinates mode='nearest' == clamp the coordinate
  y2 = y2 < 0.0 ? 0.0 : (y2 > hm ? hm : y2);
  int fx = (int)floor(x2), fy = (int)floor(y2);
  double tx = x2 - (double)fx, ty = y2 - (double)fy;
  int x1 = min(fx + 1, W - 1), y1 = min(fy + 1, H - 1);
  double it = 0.0;
  for (int c = 0; c < NC; ++c) {
    const double *im1 = frames + (long long)c * HW, *im2 = im1 + (long long)NC * HW;
    double w2 = (1.0 - ty) * ((1.0 - tx) * im2[(long long)fy * W + fx] + tx * im2[(long long)fy * W + x1]) +
                ty * ((1.0 - tx) * im2[(long long)y1 * W + fx] + tx * im2[(long long)y1 * W + x1]);
    it += fabs(w2 - im1[i]);
  }
  if (NC > 1) it /= (double)NC;
  occ[off + i] = exp(-(div * div) / (2.0 * (sigma_d * sigma_d))) * exp(-(it * it) / (2.0 * (sigma_i * sigma_i)));
}

int k_occlusion(b200flow_ctx *ctx, const double2 *uv, const double *frames, long long bstride, int NC, int B,
                int H, int W, double sigma_d, double sigma_i, double *occ) {
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
  BF_LAUNCH(ctx, occlusion_kernel, grd, blk, 0, uv, frames, bstride, NC, H, W, sigma_d, sigma_i, occ);
  return 0;
}

__global__ void sub_kernel(const double2 *__restrict__ a, const double2 *__restrict__ b, double2 *__restrict__ out,
                           long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double2(a[i].x - b[i].x, a[i].y - b[i].y);
}

int k_sub(b200flow_ctx *ctx, const double2 *a, const double2 *b, double2 *out, long long n) {
  BF_LAUNCH(ctx, sub_kernel, (unsigned)cdiv(n, 256), 256, 0, a, b, out, n);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// colour/occlusion weighted median over a (2 hsz+1)^2 window (weighted_median.py:24-112)
//   one WARP per output pixel; the window's n samples (u, v, weight) live NPL per lane in registers.
//   The reference sorts the window and walks the cumulative weights; the value it returns is
//       t* = min { x_k : S(x_k) >= total/2 },   S(t) = sum of the weights of the samples <= t,
//   which is found here WITHOUT sorting, by bisection on the sample values: keep an element-valued bracket
//   [lo, hi] with S(hi) >= total/2 and S(x) < total/2 for every sample x < lo; evaluate S at the midpoint
//   together with the largest sample <= pivot and the smallest sample > pivot (one pass over the lane's
//   registers + three warp reductions) and snap the bracket to those samples.  The bracket shrinks by at
//   least one distinct sample per step and typically halves (~8-10 steps for 225 samples); u and v are
//   advanced together for instruction-level parallelism.  The result is always one of the window's samples.
//   Tile of colour / occ / flow staged in shared memory with the NumPy 'reflect' (mirror, no edge repeat)
//   boundary.  Compute-bound (fp64 compare/add/min/max + shuffles), ~64 B/pixel of HBM traffic.
// ------------------------------------------------------------------------------------------------
#ifndef WM_WARPS_N
#define WM_WARPS_N 8
#endif
constexpr int WM_TW = 16, WM_TH = WM_WARPS_N, WM_WARPS = WM_WARPS_N;   // one warp per tile row

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exp(a) for a <= 0, used only as max(exp(a) * occ, 1e-10): the power-of-two scaling is clamped at 2^-1000 (such weights
// sit on the 1e-10 floor whatever the exponential's exact value).  Cody-Waite reduction + degree-13 Taylor polynomial on |r| <= ln2/2
// (truncation error 1.7e-16), coefficients read as constant-bank operands: ~24 instructions against ~46 + call for exp().
__constant__ double WM_EXPC[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                   1.0 / 362880.0,     1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,
                                   1.0 / 120.0,        1.0 / 24.0,        1.0 / 6.0,        0.5};
__constant__ double WM_FLOOR = 1e-10;      // np.maximum(w, 1e-10): a constant-bank operand (an fp64 literal costs two moves per use)
__device__ __forceinline__ double exp_nonpos(double a) {
  const double t = fma(a, 1.4426950408889634, 6755399441055744.0);
  const int k = max(__double2loint(t), -1000);                // a < -693: any value below the 1e-10 floor will do
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, a);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = WM_EXPC[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = fma(p, r, WM_EXPC[i]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));   // p * 2^k, k in [-1000, 0], p in [0.7, 1.42]
}

// largest double strictly below a finite x
__device__ __forceinline__ double next_below(double x) {
  long long b = __double_as_longlong(x);
  if (x > 0.0) return __longlong_as_double(b - 1);
  if (x < 0.0) return __longlong_as_double(b + 1);
  return -4.9406564584124654e-324;
}

// exact S(p) = sum of the weights of the samples <= p, fp64 (all lanes return it)
template <int NPL>
__device__ __forceinline__ double exact_cum_weight(const double (&x)[NPL], const double (&w)[NPL], double p) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NPL; ++k) s += x[k] <= p ? w[k] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Fixed-point shadow of the weights for the bracketing phase: pk = (floor(w * 2^22 / total) << 9) | 1.  One predicated
// integer add per sample accumulates BOTH the weight below the pivot (high 23 bits; the sum over a window is <= 2^22)
// and the number of samples below it (low 9 bits, <= 256), and ONE REDUX.SUM reduces both over the warp -- against
// DSETP + DADD + 2 FSEL + IADD per sample and 25 instructions of fp64 shuffle tree.  Truncation loses < 1 unit per
// sample, so with F = the reduced high field and c = the count the true S(p) * 2^22 / total lies in [F, F + c]: the
// comparison with total / 2 (= 2^21) is decided in fixed point whenever it is safe by a margin, and falls back to the
// exact fp64 sum otherwise (ties and near-ties only).
constexpr unsigned WM_HALF_FIX = 1u << 21;
__device__ __forceinline__ unsigned wm_pack(double w, double scale) {
  return (__double2uint_rz(w * scale) << 9) | 1u;
}

// Weighted median of the n samples x[] with weights w[] held NPL per lane in the warp's registers (all lanes return it):
//   t* = min { x_k : S(x_k) >= half },  S(t) = sum of w_k over x_k <= t.
// Padding slots hold x = +inf, w = 0, pk = 0, so they are never counted, weighed or bracketed.
// Stage 1 bisects the VALUE range (L, R] -- invariant S(L) < half <= S(R) -- with steps that only need to know on which
//   side of half S(p) lies and the count n(p): the fixed-point shadow above (pk), exact fp64 only when that is not
//   decisive; a step that separates nothing (ties / clusters) snaps the bracket to the extreme samples inside it.
//   S(L) itself is evaluated once, exactly, when the bracket is final.
// Stage 2, as soon as at most 32 samples are left inside the bracket (about 4-5 steps for 225 samples), compacts them one
//   per lane through `scratch` (warp-private shared memory).
// Stage 3 keeps bisecting on that one-sample-per-lane set (a step is now ~25 instructions) down to <= 8 survivors and
//   finishes exactly: every survivor evaluates S at its own value (S(L) + an all-pairs pass over the survivors) and
//   the smallest one with S >= half is the answer.  The result is always one of the window's samples.
template <int NPL, int NFULL>
__device__ __forceinline__ double weighted_select(const double (&x)[NPL], const double (&w)[NPL],
                                                  const unsigned (&pk)[NPL], int n, double half, double2 *scratch,
                                                  int lane, float flo, float fhi) {
  // flo < every sample <= fhi: fp32 STRICT lower / upper bounds of the window's values (window_ranges: a separable min / max
  // over the staged tile widened by one fp32 ulp, ~10 warp instructions per pixel instead of a 55-instruction reduction over
  // the registers per selection)
  double L = (double)flo, R = (double)fhi;               // open-closed value bracket: S(L) = 0 < half <= S(R) = total
  int nL = 0, nR = n;                                    // samples <= L, <= R
  unsigned FL = 0u;                                      // fixed-point lower bound of S(L): S(L) * 2^22 / total in [FL, FL + nL]
  while (nR - nL > 32) {
    double p = 0.5 * (L + R);
    bool snap = !(p > L && p < R);
    if (!snap) {
      unsigned acc = 0u;
#pragma unroll
      for (int k = 0; k < NPL; ++k)                      // DSETP + predicated IADD, spelled out: the compiler's own choice is
        asm("{\n\t.reg .pred q;\n\tsetp.le.f64 q, %1, %2;\n\t@q add.u32 %0, %0, %3;\n\t}"      // SEL + 3-input adds + predicate
            : "+r"(acc) : "d"(x[k]), "d"(p), "r"(pk[k]));                                    // spills: 28 instead of 17 per step
      acc = __reduce_add_sync(0xffffffffu, acc);
      const int c = (int)(acc & 511u);
      const unsigned f = acc >> 9;
      bool ge;                                             // S(p) >= half ?
      if (f >= WM_HALF_FIX + 4u) ge = true;
      else if (f + (unsigned)c + 4u < WM_HALF_FIX) ge = false;
      else ge = exact_cum_weight<NPL>(x, w, p) >= half;    // too close to call in fixed point
      if (ge) { snap = c == nR; R = p; nR = c; }
      else { snap = c == nL; L = p; nL = c; FL = f; }
    }
    if (snap) {
      // a = smallest sample > L, b = largest sample <= R  (both exist: the bracket holds weight)
      double a = INFINITY, b = -INFINITY;
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        const double xk = x[k];
        if (xk > L) a = xk < a ? xk : a;
        if (xk <= R) b = xk > b ? xk : b;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double t0 = __shfl_xor_sync(0xffffffffu, a, o), t1 = __shfl_xor_sync(0xffffffffu, b, o);
        a = t0 < a ? t0 : a; b = t1 > b ? t1 : b;
      }
      if (!(a < b)) return b;              // every sample left in the bracket has the same value
      L = next_below(a); R = b;            // no sample lies in (old L, new L]: S(L) and the counts are unchanged
    }
  }
  // stage 2: compact the survivors (L < x <= R), one per lane, each with its fixed-point weight
  int cnt = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    const bool in = x[k] > L && x[k] <= R;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) scratch[cnt + __popc(m & lt)] = make_double2(x[k], __hiloint2double(0, (int)pk[k]));
    cnt += __popc(m);
  }
  __syncwarp();
  bool valid = lane < cnt;
  const double2 me = valid ? scratch[lane] : make_double2(INFINITY, 0.0);
  const unsigned mypk = (unsigned)__double2loint(me.y);
  __syncwarp();
  // stage 3a: bisection on the compacted set, still in fixed point (S(p) * 2^22 / total in [FL + f, FL + f + nL + c])
  while (cnt > 8) {
    const double p = 0.5 * (L + R);
    if (!(p > L && p < R)) break;
    const bool le = valid && me.x <= p;
    const unsigned acc = __reduce_add_sync(0xffffffffu, le ? mypk : 0u);
    const int c = (int)(acc & 511u);
    if (c == 0 || c == cnt) break;         // nothing separated (cluster): let the all-pairs pass sort it out
    const unsigned f = FL + (acc >> 9);
    bool ge;
    if (f >= WM_HALF_FIX + 4u) ge = true;
    else if (f + (unsigned)(nL + c) + 4u < WM_HALF_FIX) ge = false;
    else ge = exact_cum_weight<NPL>(x, w, p) >= half;
    if (ge) { R = p; valid = le; cnt = c; }
    else { L = p; FL = f; nL += c; valid = valid && !le; cnt -= c; }
  }
  // stage 3b: S at every survivor from an all-pairs pass over the survivors, in fixed point; a survivor whose S is too
  // close to half to call (true value in [f, f + count]) gets the exact fp64 sum over the whole window instead
  // (the survivors go back to the scratch row, packed, and every lane reads them all: one broadcast shared load per pair
  // instead of three shuffles + a bit scan)
  unsigned acc = 0u;
  {
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const int nv = __popc(vm);
    if (valid) scratch[__popc(vm & lt)] = make_double2(me.x, __hiloint2double(0, (int)mypk));
    __syncwarp();
    for (int j = 0; j < nv; ++j) {
      const double2 e = scratch[j];
      if (e.x <= me.x) acc += (unsigned)__double2loint(e.y);
    }
    __syncwarp();
  }
  const unsigned fi = FL + (acc >> 9), ci = (unsigned)nL + (acc & 511u);
  bool ge = fi >= WM_HALF_FIX + 4u;
  const bool unsure = valid && !ge && !(fi + ci + 4u < WM_HALF_FIX);
  for (unsigned m = __ballot_sync(0xffffffffu, unsure); m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const double sj = exact_cum_weight<NPL>(x, w, __shfl_sync(0xffffffffu, me.x, j));
    if (lane == j) ge = sj >= half;
  }
  const bool c_ok = ge;                                    // S(own value) >= half
  double ans = (valid && c_ok) ? me.x : INFINITY;          // some survivor qualifies: S(largest survivor) = S(R) >= half
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double t = __shfl_xor_sync(0xffffffffu, ans, o);
    ans = t < ans ? t : ans;
  }
  return ans;
}

// Value range of every output pixel's window, for both flow components: separable min / max over the staged tile in fp32
// (values rounded to nearest on staging, the result widened by one fp32 ulp on each side, so [lo, hi] always contains the
// window).  s_f: [SH][SW] staged {u, v}; s_rr: [SH][WM_TW] row pass; s_rng: [WM_TH][WM_TW] {ulo, uhi, vlo, vhi}.
// Called by all threads of the CTA; the caller synchronises before s_f is read and after s_rng is written.
__device__ __forceinline__ float f32_below(float v) {
  const int b = __float_as_int(v);
  return v > 0.f ? __int_as_float(b - 1) : (v < 0.f ? __int_as_float(b + 1) : -1.4e-45f);
}
__device__ __forceinline__ float f32_above(float v) {
  const int b = __float_as_int(v);
  return v > 0.f ? __int_as_float(b + 1) : (v < 0.f ? __int_as_float(b - 1) : 1.4e-45f);
}
__device__ __forceinline__ void window_ranges(const float2 *s_f, float4 *s_rr, float4 *s_rng, int SW, int SH, int wsz,
                                              int TW, int TH) {
  for (int t = threadIdx.x; t < SH * TW; t += blockDim.x) {
    const int sy = t / TW, lx = t - sy * TW;
    const float2 *row = s_f + sy * SW + lx;
    float2 e = row[0];
    float ulo = e.x, uhi = e.x, vlo = e.y, vhi = e.y;
    for (int dx = 1; dx < wsz; ++dx) {
      e = row[dx];
      ulo = fminf(ulo, e.x); uhi = fmaxf(uhi, e.x); vlo = fminf(vlo, e.y); vhi = fmaxf(vhi, e.y);
    }
    s_rr[t] = make_float4(ulo, uhi, vlo, vhi);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TH * TW; t += blockDim.x) {
    float4 r = s_rr[t];
    for (int dy = 1; dy < wsz; ++dy) {
      const float4 q = s_rr[t + dy * TW];
      r.x = fminf(r.x, q.x); r.y = fmaxf(r.y, q.y); r.z = fminf(r.z, q.z); r.w = fmaxf(r.w, q.w);
    }
    s_rng[t] = make_float4(f32_below(r.x), f32_above(r.y), f32_below(r.z), f32_above(r.w));
  }
}

#ifndef WM_MINB
#define WM_MINB 3                // CTAs per SM the register budget is cut for: 3 (80 registers) 106 ms of filter time per bench step, 2 (128) 116 ms, 4 (64) 126 ms
#endif
// NFULL = number of slots k that hold a real sample in every lane (n / 32, or 0 for the generic instantiation): those slots
// need no padding logic at all -- for the 15 x 15 window of the presets that is 7 of the 8 slots.
template <int NPL, int NFULL>
__global__ void __launch_bounds__(WM_WARPS * 32, WM_MINB) wmedian_kernel(const double2 *__restrict__ cand,
                                                                  const double2 *__restrict__ base,
                                                                  const double *__restrict__ color, int C,
                                                                  const double *__restrict__ occ, int H, int W, int hsz,
                                                                  double inv2s2, double2 *__restrict__ out) {
  extern __shared__ double smem[];
  const int wsz = 2 * hsz + 1, n = wsz * wsz;
  const int SW = WM_TW + 2 * hsz, SH = WM_TH + 2 * hsz, SN = SW * SH;
  // staged tile: flow (u, v) and one 32-byte record {colour 0..2 (missing channels 0), occ} per pixel, so that a
  // sample's weight needs one address and two 16-byte shared loads
  double2 *s_uv = reinterpret_cast<double2 *>(smem);                      // [SN]
  double4 *s_cw = reinterpret_cast<double4 *>(smem + 2 * SN);             // [SN]
  double2 *s_scr = reinterpret_cast<double2 *>(smem + 6 * SN) + 32 * (threadIdx.x >> 5);   // [WM_WARPS][32] compaction scratch
  float2 *s_f = reinterpret_cast<float2 *>(reinterpret_cast<double2 *>(smem + 6 * SN) + 32 * WM_WARPS);   // [SN] fp32 copy of the flow
  float4 *s_rr = reinterpret_cast<float4 *>(s_f + SN);                    // [SH][WM_TW] row pass of window_ranges
  float4 *s_rng = s_rr + SH * WM_TW;                                      // [WM_TH][WM_TW] value ranges
  const int b = blockIdx.z;
  const long long HW = (long long)H * W, off = (long long)b * HW;
  const int x0 = blockIdx.x * WM_TW, y0 = blockIdx.y * WM_TH;
  for (int t = threadIdx.x; t < SN; t += blockDim.x) {
    int sy = t / SW, sx = t - sy * SW;
    int gy = mirror_idx(y0 + sy - hsz, H), gx = mirror_idx(x0 + sx - hsz, W);
    long long gi = (long long)gy * W + gx;
    double2 f = cand[off + gi];
    s_uv[t] = make_double2(f.x + 0.0, f.y + 0.0);              // canonicalise -0.0
    s_f[t] = make_float2(__double2float_rn(f.x), __double2float_rn(f.y));
    const double *cp = color + (long long)b * C * HW + gi;
    s_cw[t] = make_double4(cp[0], C > 1 ? cp[HW] : 0.0, C > 2 ? cp[2 * HW] : 0.0, occ[off + gi]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // window offsets of this lane's NPL samples relative to the window's top-left corner (pixel independent);
  // padding slots (e >= n) point at the window centre and are given zero weight / x = +inf below
  int qoff[NPL];
  unsigned vmask = 0u;
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    int e = k * 32 + lane;
    int dy = e / wsz, dx = e - dy * wsz;
    qoff[k] = e < n ? dy * SW + dx : hsz * SW + hsz;
    vmask |= e < n ? 1u << k : 0u;
  }
  __syncthreads();
  window_ranges(s_f, s_rr, s_rng, SW, SH, wsz, WM_TW, WM_TH);
  __syncthreads();
  const int py = y0 + warp;
  if (py >= H) return;
  for (int lx = 0; lx < WM_TW; ++lx) {
    const int px = x0 + lx;
    if (px >= W) break;
    const int org = warp * SW + lx;
    const double4 cc = s_cw[org + hsz * SW + hsz];
    const float4 rng = s_rng[warp * WM_TW + lx];
    double x[NPL], w[NPL];
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const double4 cq = s_cw[org + qoff[k]];
      const double d0 = cq.x - cc.x, d1 = cq.y - cc.y, d2 = cq.z - cc.z;
      double cd = d0 * d0;
      cd += d1 * d1;
      cd += d2 * d2;
      const double wk = exp_nonpos(-cd * inv2s2) * cq.w;
      // np.maximum(w, 1e-10); padding slots get weight 0
      const double wfl = wk > WM_FLOOR ? wk : WM_FLOOR;
      w[k] = (k < NFULL || ((vmask >> k) & 1u)) ? wfl : 0.0;
      tot += w[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const double half = tot / 2.0;
    unsigned pk[NPL];
    {
      const double scale = 4194304.0 / tot;                  // 2^22 / total
#pragma unroll
      for (int k = 0; k < NPL; ++k) pk[k] = (k < NFULL || ((vmask >> k) & 1u)) ? wm_pack(w[k], scale) : 0u;
    }
    // padding slots: x = +inf with zero weight (never counted, weighed or bracketed)
#pragma unroll
    for (int k = 0; k < NPL; ++k) x[k] = (k < NFULL || ((vmask >> k) & 1u)) ? s_uv[org + qoff[k]].x : INFINITY;
    const double mu = weighted_select<NPL, NFULL>(x, w, pk, n, half, s_scr, lane, rng.x, rng.y);
#pragma unroll
    for (int k = 0; k < NPL; ++k) x[k] = (k < NFULL || ((vmask >> k) & 1u)) ? s_uv[org + qoff[k]].y : INFINITY;
    const double mv = weighted_select<NPL, NFULL>(x, w, pk, n, half, s_scr, lane, rng.z, rng.w);
    if (lane == 0) {
      long long gi = off + (long long)py * W + px;
      if (base) {
        double2 bs = base[gi];
        out[gi] = make_double2(bs.x + (mu - bs.x), bs.y + (mv - bs.y));
      } else {
        out[gi] = make_double2(mu, mv);
      }
    }
  }
}

template <int NPL, int NFULL>
static int launch_wmedian(b200flow_ctx *ctx, const double2 *cand, const double2 *base, const double *color, int C,
                          const double *occ, int B, int H, int W, int hsz, double sigma_i, double2 *out) {
  int SW = WM_TW + 2 * hsz, SH = WM_TH + 2 * hsz;
  size_t smem = (size_t)SW * SH * 6 * sizeof(double) + WM_WARPS * 32 * sizeof(double2)          // staged tile + compaction scratch
                + (size_t)SW * SH * sizeof(float2) + (size_t)(SH + WM_TH) * WM_TW * sizeof(float4);   // window_ranges: fp32 flow, row pass, ranges
  if (smem > 200 * 1024) return set_err(ctx, B200FLOW_EINVAL, "weighted median window hsz=%d needs %zu B of shared memory", hsz, smem);
  {
    // function attributes once per process, device and window size (mutex-guarded): setting an attribute of a kernel that is
    // running on another stream blocks the host until it ends, which would serialise concurrent sub-batches
    static std::mutex mu;
    static size_t smem_set[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    if (smem_set[ctx->device & 63] < smem) {
      BF_CUDA(ctx, cudaFuncSetAttribute(wmedian_kernel<NPL, NFULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // same shared-memory configuration as the persistent solver (solve_ic.cu IC_CARVEOUT_PCT): with concurrent
      // sub-batches CTAs of both kernels share an SM, which they only can under one L1 / shared-memory split
      BF_CUDA(ctx, cudaFuncSetAttribute(wmedian_kernel<NPL, NFULL>, cudaFuncAttributePreferredSharedMemoryCarveout, 64));
      smem_set[ctx->device & 63] = smem;
    }
  }
  dim3 grd((unsigned)cdiv(W, WM_TW), (unsigned)cdiv(H, WM_TH), B);
  double inv2s2 = 1.0 / (2.0 * (sigma_i * sigma_i));
  BF_LAUNCH(ctx, (wmedian_kernel<NPL, NFULL>), grd, WM_WARPS * 32, smem, cand, base, color, C, occ, H, W, hsz, inv2s2, out);
  return 0;
}

int k_weighted_median(b200flow_ctx *ctx, const double2 *cand, const double2 *base, const double *color, int C,
                      const double *occ, int B, int H, int W, int hsz, double sigma_i, double2 *out) {
  if (hsz < 0) return set_err(ctx, B200FLOW_EINVAL, "area_hsz must be >= 0");
  if (C < 1 || C > 3) return set_err(ctx, B200FLOW_EINVAL, "weighted median supports 1..3 colour channels, got %d", C);
  int n = (2 * hsz + 1) * (2 * hsz + 1);
  if (n <= 32) return launch_wmedian<1, 0>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);
  if (n <= 64) return launch_wmedian<2, 0>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);
  if (n <= 128) return launch_wmedian<4, 0>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);
  if (n == 225) return launch_wmedian<8, 7>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);   // the presets' 15 x 15 window
  if (n <= 256) return launch_wmedian<8, 0>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);
  if (n <= 512) return launch_wmedian<16, 0>(ctx, cand, base, color, C, occ, B, H, W, hsz, sigma_i, out);
  return set_err(ctx, B200FLOW_EINVAL, "area_hsz=%d (window %d samples) unsupported (max 10)", hsz, n);
}

}  // namespace bf
