// eval.cu -- evaluation / export edges of the path kept on the device so that batch benchmarking loops never have to
// bring full flow fields back to the host:
//   flow_error      AAE / std(AE) / AEPE with the Middlebury unknown-flow mask   (evaluation/metrics.py:5-53)
//   flow_to_color   Middlebury 55-bin colour wheel coding -> uint8 RGB            (viz/flow_color.py:5-107)
//   flow_to_flo     the byte image of a Middlebury .flo file (tag, w, h, float32 (u,v) pairs)   (io/flo_io.py:46-63)
// All reductions are two-stage with a fixed summation order (no floating-point atomics): run-to-run deterministic.
#include "kernels.cuh"

namespace bf {

constexpr int EV_BLOCKS = 64, EV_THREADS = 256;

struct ErrAcc { double ae, epe, n, ae2; };

__device__ __forceinline__ double ev_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double ev_block_sum(double v, double *sm) {
  v = ev_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < EV_THREADS / 32 ? sm[threadIdx.x] : 0.0;
    r = ev_warp_sum(r);
  }
  return r;   // valid in thread 0
}

// angular error (degrees) and end-point error of one pixel; false when the ground truth is unknown
__device__ __forceinline__ bool pixel_err(double2 f, double2 g, double &ae, double &epe) {
  if (!(fabs(g.x) < 1e9) || !(fabs(g.y) < 1e9)) return false;
  double n_est = 1.0 / sqrt(f.x * f.x + f.y * f.y + 1.0);
  double n_gt = 1.0 / sqrt(g.x * g.x + g.y * g.y + 1.0);
  double c = (f.x * g.x + f.y * g.y + 1.0) * n_est * n_gt;
  c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
  ae = acos(c) * 180.0 / 3.141592653589793;
  double dx = g.x - f.x, dy = g.y - f.y;
  epe = sqrt(dx * dx + dy * dy);
  return true;
}

// pass 0: per-block sums of AE, EPE and the valid count; pass 1 (mean known): per-block sums of (AE - mean)^2
__global__ void __launch_bounds__(EV_THREADS) flow_error_kernel(const double2 *__restrict__ uv,
                                                                const double2 *__restrict__ gt, int H, int W, int border,
                                                                const double *__restrict__ result, int pass,
                                                                double *__restrict__ partial) {
  __shared__ double sm[EV_THREADS / 32];
  const int b = blockIdx.y;
  const long long off = (long long)b * H * W;
  const int h = H - 2 * border, w = W - 2 * border;
  const long long n = h > 0 && w > 0 ? (long long)h * w : 0;
  const double mean = pass ? result[4 * b] : 0.0;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long k = (long long)blockIdx.x * EV_THREADS + threadIdx.x; k < n; k += (long long)EV_BLOCKS * EV_THREADS) {
    int y = (int)(k / w) + border, x = (int)(k % w) + border;
    long long i = off + (long long)y * W + x;
    double ae, epe;
    if (pixel_err(uv[i], gt[i], ae, epe)) {
      if (pass) { s0 += (ae - mean) * (ae - mean); }
      else { s0 += ae; s1 += epe; s2 += 1.0; }
    }
  }
  s0 = ev_block_sum(s0, sm);
  s1 = ev_block_sum(s1, sm);
  s2 = ev_block_sum(s2, sm);
  if (threadIdx.x == 0) {
    double *p = partial + ((long long)b * EV_BLOCKS + blockIdx.x) * 3;
    p[0] = s0; p[1] = s1; p[2] = s2;
  }
}

// result[b] = {AAE, std(AE), AEPE, valid count}
__global__ void flow_error_final_kernel(const double *__restrict__ partial, int pass, double *__restrict__ result) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < EV_BLOCKS; ++k) {
    const double *p = partial + ((long long)b * EV_BLOCKS + k) * 3;
    s0 += p[0]; s1 += p[1]; s2 += p[2];
  }
  double *r = result + 4 * b;
  if (pass == 0) {
    r[3] = s2;
    r[0] = s2 > 0.0 ? s0 / s2 : nan("");
    r[2] = s2 > 0.0 ? s1 / s2 : nan("");
  } else {
    r[1] = r[3] > 0.0 ? sqrt(s0 / r[3]) : nan("");
  }
}

int k_flow_error(b200flow_ctx *ctx, const double2 *uv, const double2 *gt, int B, int H, int W, int border, double *result) {
  if (border < 0) return set_err(ctx, B200FLOW_EINVAL, "border must be >= 0");
  double *partial;
  BF_TRY(arena_alloc(ctx, &partial, (size_t)B * EV_BLOCKS * 3));
  dim3 grd(EV_BLOCKS, B);
  for (int pass = 0; pass < 2; ++pass) {
    BF_LAUNCH(ctx, flow_error_kernel, grd, EV_THREADS, 0, uv, gt, H, W, border, result, pass, partial);
    BF_LAUNCH(ctx, flow_error_final_kernel, B, 32, 0, partial, pass, result);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Middlebury colour coding
// ------------------------------------------------------------------------------------------------
__constant__ unsigned char COLORWHEEL[55][3];

static void make_colorwheel(unsigned char cw[55][3]) {
  const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
  memset(cw, 0, 55 * 3);
  int col = 0;
  for (int i = 0; i < RY; ++i) { cw[col + i][0] = 255; cw[col + i][1] = (unsigned char)floor(255.0 * i / RY); }
  col += RY;
  for (int i = 0; i < YG; ++i) { cw[col + i][0] = (unsigned char)(255 - floor(255.0 * i / YG)); cw[col + i][1] = 255; }
  col += YG;
  for (int i = 0; i < GC; ++i) { cw[col + i][1] = 255; cw[col + i][2] = (unsigned char)floor(255.0 * i / GC); }
  col += GC;
  for (int i = 0; i < CB; ++i) { cw[col + i][1] = (unsigned char)(255 - floor(255.0 * i / CB)); cw[col + i][2] = 255; }
  col += CB;
  for (int i = 0; i < BM; ++i) { cw[col + i][2] = 255; cw[col + i][0] = (unsigned char)floor(255.0 * i / BM); }
  col += BM;
  for (int i = 0; i < MR; ++i) { cw[col + i][2] = (unsigned char)(255 - floor(255.0 * i / MR)); cw[col + i][0] = 255; }
}

// per-item maximum flow magnitude over the known pixels (two-stage max; exact, order independent)
__global__ void __launch_bounds__(EV_THREADS) max_rad_kernel(const double2 *__restrict__ uv, long long HW,
                                                             double *__restrict__ partial) {
  __shared__ double sm[EV_THREADS / 32];
  const int b = blockIdx.y;
  double m = -1.0;
  for (long long k = (long long)blockIdx.x * EV_THREADS + threadIdx.x; k < HW; k += (long long)EV_BLOCKS * EV_THREADS) {
    double2 f = uv[(long long)b * HW + k];
    if (!(fabs(f.x) > 1e9) && !(fabs(f.y) > 1e9)) {
      double r = sqrt(f.x * f.x + f.y * f.y);
      m = r > m ? r : m;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < EV_THREADS / 32; ++k) m = sm[k] > m ? sm[k] : m;
    partial[(long long)b * EV_BLOCKS + blockIdx.x] = m;
  }
}

__global__ void flow_color_kernel(const double2 *__restrict__ uv, long long HW, double max_flow,
                                  const double *__restrict__ partial, unsigned char *__restrict__ rgb) {
  const int b = blockIdx.y;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= HW) return;
  double max_rad = max_flow;
  if (!(max_flow > 0.0)) {        // auto: largest known magnitude of this item (0 if none is known)
    double m = -1.0;
    for (int j = 0; j < EV_BLOCKS; ++j) { double t = partial[(long long)b * EV_BLOCKS + j]; m = t > m ? t : m; }
    max_rad = m < 0.0 ? 0.0 : m;
  }
  max_rad = max_rad > 1e-8 ? max_rad : 1e-8;
  const double2 f = uv[(long long)b * HW + k];
  unsigned char *o = rgb + 3 * ((long long)b * HW + k);
  if (fabs(f.x) > 1e9 || fabs(f.y) > 1e9) { o[0] = o[1] = o[2] = 0; return; }
  const double u = f.x / max_rad, v = f.y / max_rad;
  const double rad = sqrt(u * u + v * v);
  const double a = atan2(-v, -u) / 3.141592653589793;
  const double fk = (a + 1.0) / 2.0 * 54.0;
  const int k0 = (int)floor(fk);
  const int k1 = k0 + 1 == 55 ? 0 : k0 + 1;
  const double fr = fk - (double)k0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double tmp = (double)COLORWHEEL[k0][i] / 255.0 * (1.0 - fr) + (double)COLORWHEEL[k1][i] / 255.0 * fr;
    tmp = 1.0 - rad * (1.0 - tmp);
    if (rad > 1.0) tmp = tmp * 0.75;
    tmp = tmp < 0.0 ? 0.0 : (tmp > 1.0 ? 1.0 : tmp);
    o[i] = (unsigned char)floor(255.0 * tmp);
  }
}

int k_flow_to_color(b200flow_ctx *ctx, const double2 *uv, int B, int H, int W, double max_flow, unsigned char *rgb) {
  static int wheel_dev = -1;
  if (wheel_dev != ctx->device) {
    unsigned char cw[55][3];
    make_colorwheel(cw);
    BF_CUDA(ctx, cudaMemcpyToSymbol(COLORWHEEL, cw, sizeof cw));
    wheel_dev = ctx->device;
  }
  const long long HW = (long long)H * W;
  double *partial;
  BF_TRY(arena_alloc(ctx, &partial, (size_t)B * EV_BLOCKS));
  if (!(max_flow > 0.0)) BF_LAUNCH(ctx, max_rad_kernel, dim3(EV_BLOCKS, B), EV_THREADS, 0, uv, HW, partial);
  BF_LAUNCH(ctx, flow_color_kernel, dim3((unsigned)cdiv(HW, 256), B), 256, 0, uv, HW, max_flow, partial, rgb);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// .flo byte image: float32 tag 202021.25, int32 width, int32 height, then (u, v) float32 pairs row-major
// ------------------------------------------------------------------------------------------------
__global__ void flow_to_flo_kernel(const double2 *__restrict__ uv, int H, int W, unsigned char *__restrict__ out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W;
  const long long item = 12 + 8 * HW;
  unsigned char *o = out + (long long)blockIdx.y * item;
  if (k == 0) {
    float tag = 202021.25f;
    int w = W, h = H;
    memcpy(o, &tag, 4); memcpy(o + 4, &w, 4); memcpy(o + 8, &h, 4);
  }
  if (k >= HW) return;
  const double2 f = uv[(long long)blockIdx.y * HW + k];
  float2 g = make_float2((float)f.x, (float)f.y);
  memcpy(o + 12 + 8 * k, &g, 8);       // 12-byte header: payload is only 4-byte aligned
}

int k_flow_to_flo(b200flow_ctx *ctx, const double2 *uv, int B, int H, int W, unsigned char *out) {
  const long long HW = (long long)H * W;
  BF_LAUNCH(ctx, flow_to_flo_kernel, dim3((unsigned)cdiv(HW, 256), B), 256, 0, uv, H, W, out);
  return 0;
}

}  // namespace bf
