// pre.cu -- kernel group (1): layout changes, colour conversion, min/max scaling, ROF structure-texture
// decomposition, Gaussian pyramid, flow resampling.  Replaces interface.py:74-141, image_processing.py:6-136,
// pyramid.py:6-73, warping.py:6-45 of the reference.  All kernels are one-thread-per-output-element
// streaming stencils (HBM-bound, coalesced along W); compiled with -fmad=false so that the integer-valued
// decisions (gray quantisation, resize sample indices) round exactly like the NumPy reference.
#include <cuda.h>
#include <mutex>
#include "kernels.cuh"

namespace bf {

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
__global__ void deinterleave_kernel(const double *__restrict__ src, double *__restrict__ dst, long long total,
                                    long long HW, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // index into dst (B,C,HW)
  if (i >= total) return;
  long long px = i % HW;
  long long t = i / HW;
  int c = (int)(t % C);
  long long b = t / C;
  dst[i] = src[(b * HW + px) * C + c];
}

__global__ void interleave_kernel(const double *__restrict__ src, double *__restrict__ dst, long long total,
                                  long long HW, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // index into dst (B,HW,C)
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  long long px = t % HW;
  long long b = t / HW;
  dst[i] = src[(b * C + c) * HW + px];
}

int k_deinterleave(b200flow_ctx *ctx, const double *src, double *dst, int B, long long HW, int C) {
  long long total = (long long)B * HW * C;
  BF_LAUNCH(ctx, deinterleave_kernel, (unsigned)cdiv(total, 256), 256, 0, src, dst, total, HW, C);
  return 0;
}
int k_interleave(b200flow_ctx *ctx, const double *src, double *dst, int B, long long HW, int C) {
  long long total = (long long)B * HW * C;
  BF_LAUNCH(ctx, interleave_kernel, (unsigned)cdiv(total, 256), 256, 0, src, dst, total, HW, C);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// min/max scaling (scale_image, image_processing.py:6-26): exact, order-independent reduction
// ------------------------------------------------------------------------------------------------
constexpr int MM_BLOCKS = 64;   // partial blocks per item

__global__ void minmax_partial_kernel(const double *__restrict__ in, long long n, double2 *__restrict__ part) {
  const double *p = in + (long long)blockIdx.y * n;
  double lo = INFINITY, hi = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = p[i];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double slo[32], shi[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { slo[w] = lo; shi[w] = hi; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    lo = l < nw ? slo[l] : INFINITY;
    hi = l < nw ? shi[l] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (l == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = make_double2(lo, hi);
  }
}

__global__ void minmax_final_kernel(const double2 *__restrict__ part, int nblk, double2 *__restrict__ mm) {
  int item = blockIdx.x;
  double lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < nblk; i += 32) {
    double2 v = part[item * nblk + i];
    lo = fmin(lo, v.x);
    hi = fmax(hi, v.y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (threadIdx.x == 0) mm[item] = make_double2(lo, hi);
}

__global__ void scale_apply_kernel(const double *__restrict__ in, double *__restrict__ out, long long n,
                                   const double2 *__restrict__ mm, double lo, double hi) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int item = blockIdx.y;
  double2 m = mm[item];
  long long o = (long long)item * n + i;
  double v;
  if (m.y == m.x) v = (lo + hi) / 2.0;
  else v = (in[o] - m.x) / (m.y - m.x) * (hi - lo) + lo;
  out[o] = v;
}

static int minmax_items(b200flow_ctx *ctx, const double *in, int items, long long n, double2 **mm_out) {
  double2 *part, *mm;
  int nblk = (int)std::min<long long>(MM_BLOCKS, std::max<long long>(1, cdiv(n, 1024)));
  BF_TRY(arena_alloc(ctx, &part, (size_t)items * nblk));
  BF_TRY(arena_alloc(ctx, &mm, (size_t)items));
  BF_LAUNCH(ctx, minmax_partial_kernel, dim3(nblk, items), 256, 0, in, n, part);
  BF_LAUNCH(ctx, minmax_final_kernel, items, 32, 0, part, nblk, mm);
  *mm_out = mm;
  return 0;
}

int k_minmax_scale(b200flow_ctx *ctx, const double *in, double *out, int items, long long n, double lo, double hi) {
  if (items <= 0 || n <= 0) return 0;
  double2 *mm;
  BF_TRY(minmax_items(ctx, in, items, n, &mm));
  BF_LAUNCH(ctx, scale_apply_kernel, dim3((unsigned)cdiv(n, 256), items), 256, 0, in, out, n, mm, lo, hi);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ROF structure-texture decomposition (image_processing.py:52-136)
//   per iteration:  u = im + theta*div p ;  p += delta*grad u ;  p /= max(1,|p|)
//   algorithmic bytes per pixel per iteration: read im 8 + p 16, write p 16 = 40 B
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rof_div(const double2 *__restrict__ p, int y, int x, int W) {
  // backward differences with the reference's boundary rule: first column / row use p itself
  double2 c = p[(long long)y * W + x];
  double dx = x > 0 ? c.x - p[(long long)y * W + x - 1].x : c.x;
  double dy = y > 0 ? c.y - p[(long long)(y - 1) * W + x].y : c.y;
  return dx + dy;
}

// reprojection p /= max(1, |p|) (image_processing.py:118-122).  |p| <= 1 leaves p untouched (dividing by 1.0 is exact); the
// other branch multiplies by rsqrt(|p|^2) instead of sqrt + two divides: the dual iteration is fp64-pipe bound on the tiled
// kernel and the divides were half of its arithmetic.  <= 2 ulp from the reference's rounding, contractive iteration: the
// texture image agrees with the reference's to 1e-12 (tests: ROF goldens at 7 and 100 iterations, tolerance 1e-9).
__device__ __forceinline__ void rof_project(double2 &p) {
  const double n2 = p.x * p.x + p.y * p.y;
  if (n2 > 1.0) {
    const double r = rsqrt(n2);
    p.x = p.x * r;
    p.y = p.y * r;
  }
}

__global__ void rof_iter_kernel(const double *__restrict__ im, const double2 *__restrict__ pin,
                                double2 *__restrict__ pout, int H, int W, double theta, double delta) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  long long off = (long long)blockIdx.z * H * W;
  im += off; pin += off; pout += off;
  long long i = (long long)y * W + x;
  double u = im[i] + theta * rof_div(pin, y, x, W);
  double gx = 0.0, gy = 0.0;
  if (x < W - 1) gx = (im[i + 1] + theta * rof_div(pin, y, x + 1, W)) - u;
  if (y < H - 1) gy = (im[i + W] + theta * rof_div(pin, y + 1, x, W)) - u;
  double2 p = pin[i];
  p.x = p.x + delta * gx;
  p.y = p.y + delta * gy;
  rof_project(p);
  pout[i] = p;
}

__global__ void rof_finish_kernel(const double *__restrict__ im, const double2 *__restrict__ p,
                                  double *__restrict__ out, int H, int W, double theta, double alp) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  long long off = (long long)blockIdx.z * H * W;
  long long i = (long long)y * W + x;
  double s = im[off + i] + theta * rof_div(p + off, y, x, W);
  out[off + i] = im[off + i] - alp * s;
}

// ---- ROF on shared-memory tiles, RT_K dual iterations per launch (temporal blocking), tiles loaded by TMA ---------------
// A CTA loads the (RT_TW + 2 RT_K) x (RT_TH + 2 RT_K) neighbourhood of its RT_TW x RT_TH output tile -- the normalised image and
// the dual field p -- with two cp.async.bulk.tensor loads (one elected thread, completion on an mbarrier), runs RT_K
// iterations on the tile in shared memory (the region that is still exact shrinks by one pixel per iteration), and writes
// the interior of p.  The TMA unit zero-fills everything outside the image, which IS the reference's boundary rule for p
// (div p at the first column / row uses p itself = p - 0); the forward differences of u are switched off at the last
// column / row by global coordinates.  Arithmetic is operation-for-operation that of rof_iter_kernel (bit-identical
// results); HBM traffic per pixel and iteration drops from 40 B to (24 x 1.69 + 16) / RT_K = 14 B.
// Every thread keeps a fixed set of seven staged pixels for all iterations (no index arithmetic in the loop, seven independent
// dependency chains).  Measured (B200, 32 planes of 640x480): see DESIGN.md section 4 -- the kernel is bound by the fp64
// pipe, not by HBM, which is why a version that marched down columns to evaluate u once per pixel (fewer operations, but a
// serial chain per warp) was slower.
// Needs an even image width (TMA row pitch must be a multiple of 16 bytes); other widths keep rof_iter_kernel.
constexpr int RT_K = 4, RT_TW = 64, RT_TH = 16, RT_BW = RT_TW + 2 * RT_K, RT_BH = RT_TH + 2 * RT_K, RT_N = RT_BW * RT_BH;
constexpr int RT_THREADS = 256;
constexpr size_t RT_SMEM = 128 + (size_t)RT_N * (sizeof(double) + 2 * sizeof(double2));   // image + two copies of p (ping-pong)

__device__ __forceinline__ double rof_div_s(const double2 *sp, int l) {   // l = ly * RT_BW + lx, all four neighbours staged
  const double2 c = sp[l];
  return (c.x - sp[l - 1].x) + (c.y - sp[l - RT_BW].y);
}

__global__ void __launch_bounds__(RT_THREADS, 3) rof_tile_kernel(const __grid_constant__ CUtensorMap tm_im,
                                                                 const __grid_constant__ CUtensorMap tm_p,
                                                                 double2 *__restrict__ pout, int H, int W, double theta,
                                                                 double delta, int iters) {
  extern __shared__ unsigned char rt_raw[];
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(rt_raw) + 127) & ~uintptr_t(127));
  double *s_im = reinterpret_cast<double *>(base);                                    // [RT_BH][RT_BW]
  double2 *s_p0 = reinterpret_cast<double2 *>(base + (size_t)RT_N * sizeof(double));  // [RT_BH][RT_BW] (x, y) interleaved
  double2 *s_p1 = s_p0 + RT_N;
  __shared__ __align__(8) unsigned long long bar;
  const int x0 = blockIdx.x * RT_TW - RT_K, y0 = blockIdx.y * RT_TH - RT_K, plane = blockIdx.z;
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // make the initialised barrier visible to the TMA unit
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(RT_N * (sizeof(double) + sizeof(double2)));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    // coordinates innermost first; anything outside [0, W) x [0, H) arrives as zeros
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(s_im)), "l"(reinterpret_cast<unsigned long long>(&tm_im)), "r"(x0), "r"(y0),
                   "r"(plane), "r"(bar_a) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(s_p0)), "l"(reinterpret_cast<unsigned long long>(&tm_p)), "r"(2 * x0), "r"(y0),
                   "r"(plane), "r"(bar_a) : "memory");
  }
  {
    unsigned done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar_a) : "memory");
  }
  double2 *src = s_p0, *dst = s_p1;
  constexpr int NPX = (RT_N + RT_THREADS - 1) / RT_THREADS;      // 7 staged pixels per thread
  int lxs[NPX], lys[NPX];
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    const int idx = threadIdx.x + i * RT_THREADS;
    lys[i] = idx < RT_N ? idx / RT_BW : -100;                     // a slot past the tile never passes the region test
    lxs[i] = idx - (idx / RT_BW) * RT_BW;
  }
  for (int j = 0; j < iters; ++j) {
    const int m = j + 1;
#pragma unroll
    for (int i = 0; i < NPX; ++i) {
      const int lx = lxs[i], ly = lys[i];
      if (lx < m || lx >= RT_BW - m || ly < m || ly >= RT_BH - m) continue;
      const int gx_ = x0 + lx, gy_ = y0 + ly, l = ly * RT_BW + lx;
      double2 p = make_double2(0.0, 0.0);
      if (gx_ >= 0 && gx_ < W && gy_ >= 0 && gy_ < H) {
        const double u = s_im[l] + theta * rof_div_s(src, l);
        double gx = 0.0, gy = 0.0;
        if (gx_ < W - 1) gx = (s_im[l + 1] + theta * rof_div_s(src, l + 1)) - u;
        if (gy_ < H - 1) gy = (s_im[l + RT_BW] + theta * rof_div_s(src, l + RT_BW)) - u;
        p = src[l];
        p.x = p.x + delta * gx;
        p.y = p.y + delta * gy;
        rof_project(p);
      }
      dst[l] = p;
    }
    __syncthreads();
    double2 *t = src; src = dst; dst = t;
  }
  pout += (long long)plane * H * W;
  for (int idx = threadIdx.x; idx < RT_TW * RT_TH; idx += RT_THREADS) {
    const int ly = RT_K + idx / RT_TW, lx = RT_K + idx % RT_TW;
    const int gx_ = x0 + lx, gy_ = y0 + ly;
    if (gx_ < W && gy_ < H) pout[(long long)gy_ * W + gx_] = src[ly * RT_BW + lx];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D tensor map over P planes of H rows of `row_doubles` doubles, box bw x bh x 1
static int make_plane_map(b200flow_ctx *ctx, CUtensorMap *tm, const void *ptr, int P, int H, long long row_doubles, int bw, int bh) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    BF_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return set_err(ctx, B200FLOW_ECUDA, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t dims[3] = {(cuuint64_t)row_doubles, (cuuint64_t)H, (cuuint64_t)P};
  cuuint64_t strides[2] = {(cuuint64_t)row_doubles * 8, (cuuint64_t)row_doubles * 8 * (cuuint64_t)H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, estr[3] = {1u, 1u, 1u};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(ctx, B200FLOW_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int k_rof_texture(b200flow_ctx *ctx, const double *img, double *out, int B, int C, int H, int W, double theta,
                  int iters, double alp) {
  long long HW = (long long)H * W;
  int P = B * C;
  double *norm;
  double2 *pa, *pb;
  BF_TRY(arena_alloc(ctx, &norm, (size_t)P * HW));
  BF_TRY(arena_alloc(ctx, &pa, (size_t)P * HW));
  BF_TRY(arena_alloc(ctx, &pb, (size_t)P * HW));
  BF_TRY(k_minmax_scale(ctx, img, norm, B, (long long)C * HW, -1.0, 1.0));   // joint over the channels of a pair
  BF_CUDA(ctx, cudaMemsetAsync(pa, 0, sizeof(double2) * P * HW, ctx->stream));
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), P);
  double delta = 1.0 / (4.0 * theta);
  const bool tiled = W % 2 == 0 && W >= 16 && H >= 8 && getenv("B200FLOW_ROF_SCALAR") == nullptr;
  if (tiled) {
    static std::mutex mu;
    static bool attr_done[64] = {false};
    {
      std::lock_guard<std::mutex> lock(mu);
      if (!attr_done[ctx->device & 63]) {
        BF_CUDA(ctx, cudaFuncSetAttribute(rof_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM));
        BF_CUDA(ctx, cudaFuncSetAttribute(rof_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));      // three 69 KB tiles per SM
        attr_done[ctx->device & 63] = true;
      }
    }
    CUtensorMap tm_im, tm_pa, tm_pb;
    BF_TRY(make_plane_map(ctx, &tm_im, norm, P, H, W, RT_BW, RT_BH));
    BF_TRY(make_plane_map(ctx, &tm_pa, pa, P, H, 2LL * W, 2 * RT_BW, RT_BH));
    BF_TRY(make_plane_map(ctx, &tm_pb, pb, P, H, 2LL * W, 2 * RT_BW, RT_BH));
    dim3 tg((unsigned)cdiv(W, RT_TW), (unsigned)cdiv(H, RT_TH), P);
    bool a_is_src = true;
    for (int it = 0; it < iters; it += RT_K) {
      const int n = iters - it < RT_K ? iters - it : RT_K;
      BF_LAUNCH(ctx, rof_tile_kernel, tg, RT_THREADS, RT_SMEM, tm_im, a_is_src ? tm_pa : tm_pb, a_is_src ? pb : pa, H, W, theta,
                delta, n);
      a_is_src = !a_is_src;
    }
    if (!a_is_src) std::swap(pa, pb);           // pa = the latest p
  } else {
    for (int it = 0; it < iters; ++it) {
      BF_LAUNCH(ctx, rof_iter_kernel, grd, blk, 0, norm, pa, pb, H, W, theta, delta);
      std::swap(pa, pb);
    }
  }
  BF_LAUNCH(ctx, rof_finish_kernel, grd, blk, 0, norm, pa, (double *)pb, H, W, theta, alp);
  BF_TRY(k_minmax_scale(ctx, (double *)pb, out, B, (long long)C * HW, 0.0, 255.0));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Gaussian pyramid level: correlate(prev, G, 'reflect') then MATLAB-convention bilinear resize
// (pyramid.py:11-73).  One thread per OUTPUT pixel: 4 smoothed taps x ks^2 MACs.
// ------------------------------------------------------------------------------------------------
struct Taps { double k[81]; int ks; };

__device__ __forceinline__ double smooth_at(const double *__restrict__ src, int H, int W, int y, int x, const Taps &t) {
  int c = t.ks / 2;
  double acc = 0.0;
  for (int a = 0; a < t.ks; ++a) {
    const double *row = src + (long long)reflect_idx(y + a - c, H) * W;
    for (int b = 0; b < t.ks; ++b) acc += t.k[a * t.ks + b] * row[reflect_idx(x + b - c, W)];
  }
  return acc;
}

__device__ __forceinline__ double resize_coord(int o, int n_out, int n_in) {
  double scale = (double)n_out / (double)n_in;
  double c = ((double)o + 0.5) / scale - 0.5;
  double hi = (double)(n_in - 1);
  return c < 0.0 ? 0.0 : (c > hi ? hi : c);
}

__global__ void gauss_resize_kernel(const double *__restrict__ src, double *__restrict__ dst, int H, int W, int Hn,
                                    int Wn, Taps t) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Wn || y >= Hn) return;
  src += (long long)blockIdx.z * H * W;
  dst += (long long)blockIdx.z * Hn * Wn;
  double r = resize_coord(y, Hn, H), c = resize_coord(x, Wn, W);
  int r0 = (int)floor(r), c0 = (int)floor(c);
  double tr = r - (double)r0, tc = c - (double)c0;
  int r1 = min(r0 + 1, H - 1), c1 = min(c0 + 1, W - 1);
  double z00 = smooth_at(src, H, W, r0, c0, t), z01 = smooth_at(src, H, W, r0, c1, t);
  double z10 = smooth_at(src, H, W, r1, c0, t), z11 = smooth_at(src, H, W, r1, c1, t);
  dst[(long long)y * Wn + x] = (1.0 - tr) * ((1.0 - tc) * z00 + tc * z01) + tr * ((1.0 - tc) * z10 + tc * z11);
}

// Separable, shared-memory tiled version for rank-one kernels (every Gaussian the drivers build): a CTA stages the raw
// neighbourhood of the smoothed samples its 32 x 8 output tile interpolates between ('reflect' indexing), filters rows
// then columns in shared memory (2 ks multiply-adds per smoothed sample instead of ks^2 per TAP, four taps per output) and
// gathers the bilinear combination.  ~40 instead of ~100 multiply-adds and 4 instead of 100 global loads per output at
// spacing 2.  The rounding differs from the 2-D sum in the last bits (different summation order): the pyramids agree with
// the reference's to 1e-13 (tests: pyramid goldens, tolerance 1e-9).
struct Taps1 { double g[9]; int ks; };

__global__ void gauss_resize_tile_kernel(const double *__restrict__ src, double *__restrict__ dst, int H, int W, int Hn,
                                         int Wn, Taps1 t, int nw_max, int nh_max) {
  extern __shared__ double gr_smem[];
  const int c = t.ks / 2;
  src += (long long)blockIdx.z * H * W;
  dst += (long long)blockIdx.z * Hn * Wn;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * 8;
  const int ox1 = min(ox0 + 31, Wn - 1), oy1 = min(oy0 + 7, Hn - 1);
  // smoothed samples needed: columns cx0 .. cx1, rows ry0 .. ry1
  const int cx0 = (int)floor(resize_coord(ox0, Wn, W)), cx1 = min((int)floor(resize_coord(ox1, Wn, W)) + 1, W - 1);
  const int ry0 = (int)floor(resize_coord(oy0, Hn, H)), ry1 = min((int)floor(resize_coord(oy1, Hn, H)) + 1, H - 1);
  const int nw = cx1 - cx0 + 1, nh = ry1 - ry0 + 1;            // <= nw_max, nh_max by construction of the launcher
  const int rw = nw + 2 * c, rh = nh + 2 * c;
  double *raw = gr_smem;                                        // [rh][rw]
  double *rowf = raw + (nh_max + 2 * c) * (nw_max + 2 * c);     // [rh][nw]
  double *sm = rowf + (nh_max + 2 * c) * nw_max;                // [nh][nw]
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < rh * rw; i += 256) {
    const int y = i / rw, x = i - y * rw;
    raw[i] = src[(long long)reflect_idx(ry0 - c + y, H) * W + reflect_idx(cx0 - c + x, W)];
  }
  __syncthreads();
  for (int i = tid; i < rh * nw; i += 256) {
    const int y = i / nw, x = i - y * nw;
    double acc = 0.0;
    for (int b = 0; b < t.ks; ++b) acc += t.g[b] * raw[y * rw + x + b];
    rowf[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < nh * nw; i += 256) {
    const int y = i / nw, x = i - y * nw;
    double acc = 0.0;
    for (int a = 0; a < t.ks; ++a) acc += t.g[a] * rowf[(y + a) * nw + x];
    sm[i] = acc;
  }
  __syncthreads();
  const int x = ox0 + threadIdx.x, y = oy0 + threadIdx.y;
  if (x >= Wn || y >= Hn) return;
  const double r = resize_coord(y, Hn, H), cc = resize_coord(x, Wn, W);
  const int r0 = (int)floor(r), c0 = (int)floor(cc);
  const double tr = r - (double)r0, tc = cc - (double)c0;
  const int r1 = min(r0 + 1, H - 1), c1 = min(c0 + 1, W - 1);
  const double z00 = sm[(r0 - ry0) * nw + (c0 - cx0)], z01 = sm[(r0 - ry0) * nw + (c1 - cx0)];
  const double z10 = sm[(r1 - ry0) * nw + (c0 - cx0)], z11 = sm[(r1 - ry0) * nw + (c1 - cx0)];
  dst[(long long)y * Wn + x] = (1.0 - tr) * ((1.0 - tc) * z00 + tc * z01) + tr * ((1.0 - tc) * z10 + tc * z11);
}

int k_gauss_resize(b200flow_ctx *ctx, const double *src, double *dst, int P, int H, int W, int Hn, int Wn,
                   const double *taps, int ks) {
  if (ks > 9 || ks < 1 || (ks & 1) == 0)
    return set_err(ctx, B200FLOW_EINVAL, "pyramid smoothing kernel size %d unsupported (odd, <= 9)", ks);
  dim3 blk(32, 8), grd((unsigned)cdiv(Wn, 32), (unsigned)cdiv(Hn, 8), P);
  // rank one?  k[a][b] k[c][c] == k[a][c] k[c][b] (centre c): then k = g g^T with g = the row sums (sum of g = sum of k = 1)
  const int c = ks / 2;
  const double kcc = taps[c * ks + c];
  bool sep = kcc > 0.0 && getenv("B200FLOW_PYRAMID_2D") == nullptr;
  double sum = 0.0;
  for (int i = 0; i < ks * ks; ++i) sum += taps[i];
  for (int a = 0; a < ks && sep; ++a)
    for (int b = 0; b < ks; ++b)
      if (std::fabs(taps[a * ks + b] * kcc - taps[a * ks + c] * taps[c * ks + b]) > 1e-14 * kcc * kcc) { sep = false; break; }
  // smoothed samples one tile can need: the tile spans 31 (7) output steps of n_in / n_out source pixels, plus the bilinear partner
  const int nw_max = (int)std::floor(31.0 * W / Wn) + 3, nh_max = (int)std::floor(7.0 * H / Hn) + 3;
  const size_t smem = sizeof(double) * ((size_t)(nh_max + 2 * c) * (nw_max + 2 * c) + (size_t)(nh_max + 2 * c) * nw_max +
                                        (size_t)nh_max * nw_max);
  if (sep && smem <= 46 * 1024 && std::fabs(sum - 1.0) < 1e-12) {
    Taps1 t1;
    t1.ks = ks;
    for (int a = 0; a < ks; ++a) {
      double g = 0.0;
      for (int b = 0; b < ks; ++b) g += taps[a * ks + b];
      t1.g[a] = g;
    }
    BF_LAUNCH(ctx, gauss_resize_tile_kernel, grd, blk, smem, src, dst, H, W, Hn, Wn, t1, nw_max, nh_max);
    return 0;
  }
  Taps t;
  t.ks = ks;
  for (int i = 0; i < ks * ks; ++i) t.k[i] = taps[i];
  BF_LAUNCH(ctx, gauss_resize_kernel, grd, blk, 0, src, dst, H, W, Hn, Wn, t);
  return 0;
}

// base.py:185-188 + image_processing.py:29-49
void gaussian_taps(double spacing, double *taps, int *ks_out) {
  double sigma = std::sqrt(spacing) / std::sqrt(2.0);
  int ks = 2 * (int)std::nearbyint(1.5 * sigma) + 1;   // Python round() = half-to-even = nearbyint in default mode
  if (ks > 9) ks = 9;                                     // spacing <= 8 gives ks <= 7
  double r = (ks - 1) / 2.0, mx = 0.0, sum = 0.0;
  for (int a = 0; a < ks; ++a)
    for (int b = 0; b < ks; ++b) {
      double y = a - r, x = b - r;
      double v = std::exp(-(x * x + y * y) / (2 * sigma * sigma));
      taps[a * ks + b] = v;
      mx = std::max(mx, v);
    }
  for (int i = 0; i < ks * ks; ++i) {
    if (taps[i] < 2.220446049250313e-16 * mx) taps[i] = 0.0;
    sum += taps[i];
  }
  if (sum != 0.0)
    for (int i = 0; i < ks * ks; ++i) taps[i] /= sum;
  *ks_out = ks;
}

int level_size(int n, double ratio) {
  int v = (int)std::floor(n * ratio + 0.5);
  return v < 1 ? 1 : v;
}

int auto_levels(int H, int W, double spacing) {
  int m = H < W ? H : W;
  return 1 + (int)std::floor(std::log(m / 16.0) / std::log(spacing));
}

// ------------------------------------------------------------------------------------------------
// resample_flow (warping.py:6-45): bilinear resize, both components times the HEIGHT ratio
// ------------------------------------------------------------------------------------------------
__global__ void resample_flow_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, int h, int w, int H,
                                     int W) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  in += (long long)blockIdx.z * h * w;
  out += (long long)blockIdx.z * H * W;
  double2 res;
  if (h == H && w == W) {
    res = in[(long long)y * w + x];
  } else {
    double r = resize_coord(y, H, h), c = resize_coord(x, W, w);
    int r0 = (int)floor(r), c0 = (int)floor(c);
    double tr = r - (double)r0, tc = c - (double)c0;
    int r1 = min(r0 + 1, h - 1), c1 = min(c0 + 1, w - 1);
    double2 z00 = in[(long long)r0 * w + c0], z01 = in[(long long)r0 * w + c1];
    double2 z10 = in[(long long)r1 * w + c0], z11 = in[(long long)r1 * w + c1];
    double ratio = (double)H / (double)h;
    res.x = ((1.0 - tr) * ((1.0 - tc) * z00.x + tc * z01.x) + tr * ((1.0 - tc) * z10.x + tc * z11.x)) * ratio;
    res.y = ((1.0 - tr) * ((1.0 - tc) * z00.y + tc * z01.y) + tr * ((1.0 - tc) * z10.y + tc * z11.y)) * ratio;
  }
  out[(long long)y * W + x] = res;
}

int k_resample_flow(b200flow_ctx *ctx, const double2 *in, double2 *out, int B, int h, int w, int H, int W) {
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
  BF_LAUNCH(ctx, resample_flow_kernel, grd, blk, 0, in, out, h, w, H, W);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// colour conversion (interface.py:74-141)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double gray_from_q(double r, double g, double b) {
  // ((0.2989 R + 0.5870 G) + 0.1140 B), then round half up -- same association as the NumPy expression
  double s = 0.2989 * r + 0.5870 * g;
  s = s + 0.1140 * b;
  return floor(s + 0.5);
}

__device__ __forceinline__ double quant8(double v) {   // clip(floor(v + .5), 0, 255) then uint8 cast
  double q = floor(v + 0.5);
  return q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
}

__device__ __forceinline__ void lab_from_rgb01(double r, double g, double b, double &L, double &A, double &Bb) {
  // XYZ = MAT @ RGB: each row is a 3-term dot product accumulated left to right
  double X = (0.412453 * r + 0.357580 * g) + 0.180423 * b;
  double Y = (0.212671 * r + 0.715160 * g) + 0.072169 * b;
  double Z = (0.019334 * r + 0.119193 * g) + 0.950227 * b;
  X = X / 0.950456;
  Z = Z / 1.088754;
  const double T = 0.008856, third = 1.0 / 3.0, lin = 16.0 / 116.0;
  double Y3 = pow(Y, third);
  double fX = X > T ? pow(X, third) : 7.787 * X + lin;
  double fY = Y > T ? Y3 : 7.787 * Y + lin;
  double fZ = Z > T ? pow(Z, third) : 7.787 * Z + lin;
  L = Y > T ? 116.0 * Y3 - 16.0 : 903.3 * Y;
  A = 500.0 * (fX - fY);
  Bb = 200.0 * (fY - fZ);
}

__global__ void rgb8_max_kernel(const unsigned char *__restrict__ rgb, long long n_per_item, int *__restrict__ mx) {
  const unsigned char *p = rgb + (long long)blockIdx.y * n_per_item;
  int m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_per_item; i += (long long)gridDim.x * blockDim.x)
    m = max(m, (int)p[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&mx[blockIdx.y], m);
}

__global__ void rgb8_convert_kernel(const unsigned char *__restrict__ rgb1, const unsigned char *__restrict__ rgb2,
                                    long long HW, double *__restrict__ gray, double *__restrict__ lab,
                                    const int *__restrict__ mx) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= HW) return;
  int b = blockIdx.y;
  const unsigned char *a = rgb1 + ((long long)b * HW + i) * 3;
  const unsigned char *c = rgb2 + ((long long)b * HW + i) * 3;
  gray[((long long)b * 2 + 0) * HW + i] = gray_from_q(a[0], a[1], a[2]);
  gray[((long long)b * 2 + 1) * HW + i] = gray_from_q(c[0], c[1], c[2]);
  if (lab) {
    double r = a[0], g = a[1], bl = a[2];
    if (mx[b] > 1) { r = r / 255.0; g = g / 255.0; bl = bl / 255.0; }
    double L, A, Bb;
    lab_from_rgb01(r, g, bl, L, A, Bb);
    lab[((long long)b * 3 + 0) * HW + i] = L;
    lab[((long long)b * 3 + 1) * HW + i] = A;
    lab[((long long)b * 3 + 2) * HW + i] = Bb;
  }
}

int k_rgb8_to_gray_lab(b200flow_ctx *ctx, const unsigned char *rgb1, const unsigned char *rgb2, int B, long long HW,
                       double *gray, double *lab) {
  int *mx;
  BF_TRY(arena_alloc(ctx, &mx, (size_t)B));
  BF_CUDA(ctx, cudaMemsetAsync(mx, 0, sizeof(int) * B, ctx->stream));
  if (lab) BF_LAUNCH(ctx, rgb8_max_kernel, dim3(32, B), 256, 0, rgb1, HW * 3, mx);
  BF_LAUNCH(ctx, rgb8_convert_kernel, dim3((unsigned)cdiv(HW, 256), B), 256, 0, rgb1, rgb2, HW, gray, lab, mx);
  return 0;
}

__global__ void rgbf_gray_kernel(const double *__restrict__ rgb, long long HW, double *__restrict__ gray) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= HW) return;
  gray[i] = gray_from_q(quant8(rgb[i * 3]), quant8(rgb[i * 3 + 1]), quant8(rgb[i * 3 + 2]));
}

__global__ void rgbf_lab_kernel(const double *__restrict__ rgb, long long HW, double *__restrict__ lab,
                                const double2 *__restrict__ mm) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= HW) return;
  double r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
  if (mm[0].y > 1.0) { r = r / 255.0; g = g / 255.0; b = b / 255.0; }   // any channel max > 1 (interface.py:103-106)
  double L, A, Bb;
  lab_from_rgb01(r, g, b, L, A, Bb);
  lab[i] = L;
  lab[HW + i] = A;
  lab[2 * HW + i] = Bb;
}

int k_rgbf_to_gray(b200flow_ctx *ctx, const double *rgb, long long HW, double *gray) {
  BF_LAUNCH(ctx, rgbf_gray_kernel, (unsigned)cdiv(HW, 256), 256, 0, rgb, HW, gray);
  return 0;
}

int k_rgbf_to_lab(b200flow_ctx *ctx, const double *rgb, long long HW, double *lab) {
  double2 *mm;
  BF_TRY(minmax_items(ctx, rgb, 1, HW * 3, &mm));
  BF_LAUNCH(ctx, rgbf_lab_kernel, (unsigned)cdiv(HW, 256), 256, 0, rgb, HW, lab, mm);
  return 0;
}

}  // namespace bf
