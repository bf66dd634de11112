// api.cu -- the extern "C" surface declared in include/b200flow.h.  Host-pointer entry points stage their
// arguments through the context arena on the context stream; there is no CPU implementation behind any of
// them: without a usable sm_100 device b200flow_ctx_create fails and nothing else can be called.
#include "kernels.cuh"

namespace bf {
int k_band_error(b200flow_ctx *ctx, unsigned long long *err_host);
int k_band_preload(b200flow_ctx *ctx);
PenaltySet make_penalty_set(const b200flow_params *p, double alpha);
int alloc_linsys(b200flow_ctx *ctx, int B, int H, int W, LinSys *s);
}  // namespace bf

using namespace bf;

static std::string g_create_err;
extern "C" { static void band_release(b200flow_ctx *ctx); }

#define API_BEGIN(ctx)                                                      \
  if (!(ctx)) return B200FLOW_EINVAL;                                       \
  if (cudaSetDevice((ctx)->device) != cudaSuccess) return set_err((ctx), B200FLOW_ECUDA, "cudaSetDevice failed"); \
  arena_reset(ctx);

#define API_SYNC(ctx) BF_CUDA(ctx, cudaStreamSynchronize((ctx)->stream))

extern "C" {

int b200flow_abi_version(void) { return B200FLOW_ABI_VERSION; }

int b200flow_ctx_create(int device, b200flow_ctx **out) {
  if (!out) return B200FLOW_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libb200flow has no CPU fallback)";
    return B200FLOW_ECUDA;
  }
  if (device < 0 || device >= n) {
    g_create_err = "device index out of range";
    return B200FLOW_EINVAL;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return B200FLOW_ECUDA;
  }
  if (prop.major != 10) {
    g_create_err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   "; libb200flow is built for sm_100a (B200) only";
    return B200FLOW_ECUDA;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return B200FLOW_ECUDA;
  }
  b200flow_ctx *ctx = new b200flow_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    delete ctx;
    return B200FLOW_ECUDA;
  }
  if (const char *sp = getenv("B200FLOW_SPLIT")) ctx->nsplit = atoi(sp) > 1 ? atoi(sp) : 1;   // tuning / experiments
  if (const char *sc = getenv("B200FLOW_SOLVER_CTAS")) ctx->solver_ctas_per_sm = atoi(sc) > 0 ? atoi(sc) : 0;
  *out = ctx;
  return 0;
}

void b200flow_ctx_destroy(b200flow_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto *c : ctx->subs) {
    cudaStreamSynchronize(c->stream);
    for (auto &k : c->chunks) cudaFree(k.base);
    cudaStreamDestroy(c->stream);
    if (c->solver_stream) cudaStreamDestroy(c->solver_stream);
    if (c->ev_s0) cudaEventDestroy(c->ev_s0);
    if (c->ev_s1) cudaEventDestroy(c->ev_s1);
    if (c->powtab) cudaFree(c->powtab);
    delete c;
  }
  band_release(ctx);
  for (auto e : ctx->ev_join) cudaEventDestroy(e);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  cudaStreamSynchronize(ctx->stream);
  for (auto &c : ctx->chunks) cudaFree(c.base);
  if (ctx->powtab) cudaFree(ctx->powtab);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *b200flow_last_error(const b200flow_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int b200flow_ctx_set_log(b200flow_ctx *ctx, int enabled) {
  if (!ctx) return B200FLOW_EINVAL;
  ctx->log_on = enabled != 0;
  if (!ctx->log_on) ctx->log.clear();
  return 0;
}

int b200flow_ctx_get_log(b200flow_ctx *ctx, double *rows, int cap_rows, int *n_rows) {
  if (!ctx || !n_rows || cap_rows < 0 || (cap_rows > 0 && !rows)) return set_err(ctx, B200FLOW_EINVAL, "get_log: bad arguments");
  const int n = (int)ctx->log.size();
  *n_rows = n;
  for (int k = 0; k < n && k < cap_rows; ++k) {
    const auto &r = ctx->log[k];
    double *o = rows + 5 * k;
    o[0] = r.gnc; o[1] = r.level; o[2] = r.it; o[3] = r.lin; o[4] = r.v;
  }
  return 0;
}

int b200flow_ctx_set_timing(b200flow_ctx *ctx, int enabled) {
  if (!ctx) return B200FLOW_EINVAL;
  ctx->timing = enabled != 0;
  return 0;
}

int b200flow_ctx_set_split(b200flow_ctx *ctx, int groups, int solver_ctas_per_sm) {
  if (!ctx) return B200FLOW_EINVAL;
  if (groups < 1 || groups > 8 || solver_ctas_per_sm < 0)
    return set_err(ctx, B200FLOW_EINVAL, "concurrent sub-batches: groups %d (1..8), solver CTAs per SM %d (>= 0)", groups, solver_ctas_per_sm);
  ctx->nsplit = groups;
  ctx->solver_ctas_per_sm = solver_ctas_per_sm;
  return 0;
}

// ---- row-band split of one pair over several GPUs ---------------------------------------------------------------------
static void band_release(b200flow_ctx *ctx) {
  for (int r = 0; r < B200FLOW_MAX_BAND_RANKS; ++r) {
    if (ctx->band.opened[r] && ctx->band.base[r]) cudaIpcCloseMemHandle(ctx->band.base[r]);
    ctx->band.opened[r] = false;
    ctx->band.base[r] = nullptr;
  }
  ctx->band.world = 1;
  ctx->band.rank = 0;
  ctx->band.ticket = nullptr;
}

int b200flow_band_init(b200flow_ctx *ctx, int rank, int world, unsigned long long arena_bytes, int same_device) {
  if (!ctx) return B200FLOW_EINVAL;
  if (world < 1 || world > B200FLOW_MAX_BAND_RANKS || rank < 0 || rank >= world)
    return set_err(ctx, B200FLOW_EINVAL, "row-band mode: rank %d of %d (at most %d ranks)", rank, world, B200FLOW_MAX_BAND_RANKS);
  BF_CUDA(ctx, cudaSetDevice(ctx->device));
  BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  band_release(ctx);
  for (auto &c : ctx->chunks) cudaFree(c.base);
  ctx->chunks.clear();
  if (world == 1) return 0;
  if (arena_bytes < ((size_t)64 << 20)) arena_bytes = (size_t)64 << 20;
  char *p = nullptr;
  BF_CUDA(ctx, cudaMalloc(&p, (size_t)arena_bytes));      // ONE block: it is mapped into the peers and must never move
  BF_CUDA(ctx, cudaMemset(p, 0, B200FLOW_BAND_RESERVED));
  ctx->chunks.push_back({p, (size_t)arena_bytes, B200FLOW_BAND_RESERVED});
  ctx->band.rank = rank;
  ctx->band.world = world;
  ctx->band.base[rank] = p;
  ctx->band.ticket = reinterpret_cast<unsigned *>(p + B200FLOW_BAND_RESERVED / 2);
  if (const char *mp = getenv("B200FLOW_BAND_MIN_PIXELS")) ctx->band.min_pixels = atoll(mp);
  BF_TRY(k_band_preload(ctx));
  if (same_device) {        // several ranks emulated on one GPU (tests): every rank's persistent solver must stay resident
    ctx->plain_solver_launch = true;
    ctx->solver_ctas_per_sm = 1;
    if (world > 2) return set_err(ctx, B200FLOW_EINVAL, "same-device emulation holds two ranks");
  }
  return 0;
}

int b200flow_band_export(b200flow_ctx *ctx, void *handle64, void **base) {
  if (!ctx || ctx->band.world < 2) return B200FLOW_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (base) *base = ctx->band.base[ctx->band.rank];
  if (handle64) {
    cudaIpcMemHandle_t h;
    BF_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->band.base[ctx->band.rank]));
    memcpy(handle64, &h, 64);
  }
  return 0;
}

int b200flow_band_connect(b200flow_ctx *ctx, int peer, const void *handle64, void *base) {
  if (!ctx || ctx->band.world < 2 || peer < 0 || peer >= ctx->band.world || peer == ctx->band.rank)
    return ctx ? set_err(ctx, B200FLOW_EINVAL, "row-band connect: bad peer %d", peer) : B200FLOW_EINVAL;
  BF_CUDA(ctx, cudaSetDevice(ctx->device));
  if (base) {               // same process: the peer's block is addressable as it is (same device, or peer access enabled)
    ctx->band.base[peer] = static_cast<char *>(base);
    ctx->band.opened[peer] = false;
    return 0;
  }
  if (!handle64) return set_err(ctx, B200FLOW_EINVAL, "row-band connect: neither a handle nor a base pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  BF_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->band.base[peer] = static_cast<char *>(p);
  ctx->band.opened[peer] = true;
  return 0;
}

int b200flow_band_close(b200flow_ctx *ctx) {
  if (!ctx) return B200FLOW_EINVAL;
  BF_CUDA(ctx, cudaSetDevice(ctx->device));
  BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long err = 0;
  if (ctx->band.world > 1) BF_TRY(k_band_error(ctx, &err));
  band_release(ctx);
  for (auto &c : ctx->chunks) cudaFree(c.base);
  ctx->chunks.clear();
  ctx->plain_solver_launch = false;
  ctx->solver_ctas_per_sm = 0;
  if (err) return set_err(ctx, B200FLOW_ECUDA, "row-band mode: a cross-GPU barrier timed out (barrier %llu): a peer rank stopped", err);
  return 0;
}

int b200flow_ctx_sync(b200flow_ctx *ctx) {
  if (!ctx) return B200FLOW_EINVAL;
  API_SYNC(ctx);
  return 0;
}

void *b200flow_ctx_stream(b200flow_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int b200flow_ctx_num_sms(const b200flow_ctx *ctx) { return ctx ? ctx->num_sms : 0; }

int b200flow_host_alloc(b200flow_ctx *ctx, unsigned long long bytes, void **out) {
  if (!ctx || !out) return B200FLOW_EINVAL;
  *out = nullptr;
  BF_CUDA(ctx, cudaSetDevice(ctx->device));
  BF_CUDA(ctx, cudaHostAlloc(out, bytes ? (size_t)bytes : 1, cudaHostAllocPortable));
  return 0;
}

int b200flow_host_free(b200flow_ctx *ctx, void *ptr) {
  if (!ctx) return B200FLOW_EINVAL;
  if (ptr) BF_CUDA(ctx, cudaFreeHost(ptr));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// whole pipeline
// ------------------------------------------------------------------------------------------------
static int estimate_dev_impl(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int NC, int C,
                             const double *images_dev, const double *color_dev, const double *init_dev,
                             double *uv_out_dev, b200flow_stats *stats) {
  long long HW = (long long)H * W;
  double *gray, *col = nullptr;
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  BF_TRY(arena_alloc(ctx, &gray, (size_t)B * 2 * NC * HW));
  BF_TRY(k_deinterleave(ctx, images_dev, gray, B, HW, 2 * NC));
  if (color_dev && C > 0) {
    BF_TRY(arena_alloc(ctx, &col, (size_t)B * C * HW));
    BF_TRY(k_deinterleave(ctx, color_dev, col, B, HW, C));
  }
  return run_pipeline(ctx, p, B, H, W, NC, C, gray, col, reinterpret_cast<const double2 *>(init_dev),
                      reinterpret_cast<double2 *>(uv_out_dev), stats);
}

int b200flow_estimate_dev(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int C,
                          const double *images_dev, const double *color_dev, const double *init_dev,
                          double *uv_out_dev, b200flow_stats *stats) {
  API_BEGIN(ctx);
  if (!images_dev || !uv_out_dev) return set_err(ctx, B200FLOW_EINVAL, "images / uv_out is NULL");
  BF_TRY(estimate_dev_impl(ctx, p, B, H, W, 1, C, images_dev, color_dev, init_dev, uv_out_dev, stats));
  return 0;
}

int b200flow_estimate(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int C, const double *images,
                      const double *color, const double *init, double *uv_out, b200flow_stats *stats) {
  return b200flow_estimate_mc(ctx, p, B, H, W, 1, C, images, color, init, uv_out, stats);
}

int b200flow_estimate_mc(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int NC, int C,
                         const double *images, const double *color, const double *init, double *uv_out,
                         b200flow_stats *stats) {
  API_BEGIN(ctx);
  if (!images || !uv_out) return set_err(ctx, B200FLOW_EINVAL, "images / uv_out is NULL");
  if (B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad batch/size B=%d H=%d W=%d", B, H, W);
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  size_t N = (size_t)B * H * W;
  double *d_img, *d_col = nullptr, *d_init = nullptr, *d_out;
  BF_TRY(upload(ctx, &d_img, images, 2 * NC * N));
  if (color && C > 0) BF_TRY(upload(ctx, &d_col, color, (size_t)C * N));
  if (init) BF_TRY(upload(ctx, &d_init, init, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_out, 2 * N));
  BF_TRY(estimate_dev_impl(ctx, p, B, H, W, NC, C, d_img, d_col, d_init, d_out, stats));
  BF_TRY(download(ctx, uv_out, d_out, 2 * N));
  API_SYNC(ctx);
  return 0;
}

static int estimate_rgb8_dev_impl(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W,
                                  const unsigned char *rgb1, const unsigned char *rgb2, int use_color,
                                  double *uv_out_dev, b200flow_stats *stats) {
  long long HW = (long long)H * W;
  double *gray, *lab = nullptr, *labs = nullptr;
  BF_TRY(arena_alloc(ctx, &gray, (size_t)B * 2 * HW));
  if (use_color) {
    BF_TRY(arena_alloc(ctx, &lab, (size_t)B * 3 * HW));
    BF_TRY(arena_alloc(ctx, &labs, (size_t)B * 3 * HW));
  }
  BF_TRY(k_rgb8_to_gray_lab(ctx, rgb1, rgb2, B, HW, gray, lab));
  if (use_color) BF_TRY(k_minmax_scale(ctx, lab, labs, B * 3, HW, 0.0, 255.0));   // each Lab channel separately (interface.py:59-60)
  return run_pipeline(ctx, p, B, H, W, 1, use_color ? 3 : 0, gray, labs, nullptr, reinterpret_cast<double2 *>(uv_out_dev),
                      stats);
}

int b200flow_estimate_rgb8_dev(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W,
                               const unsigned char *rgb1_dev, const unsigned char *rgb2_dev, int use_color,
                               double *uv_out_dev, b200flow_stats *stats) {
  API_BEGIN(ctx);
  if (!rgb1_dev || !rgb2_dev || !uv_out_dev) return set_err(ctx, B200FLOW_EINVAL, "NULL argument");
  BF_TRY(estimate_rgb8_dev_impl(ctx, p, B, H, W, rgb1_dev, rgb2_dev, use_color, uv_out_dev, stats));
  return 0;
}

int b200flow_estimate_rgb8(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, const unsigned char *rgb1,
                           const unsigned char *rgb2, int use_color, double *uv_out, b200flow_stats *stats) {
  API_BEGIN(ctx);
  if (!rgb1 || !rgb2 || !uv_out) return set_err(ctx, B200FLOW_EINVAL, "NULL argument");
  if (B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad batch/size B=%d H=%d W=%d", B, H, W);
  size_t N = (size_t)B * H * W;
  unsigned char *d1, *d2;
  double *d_out;
  BF_TRY(upload(ctx, &d1, rgb1, 3 * N));
  BF_TRY(upload(ctx, &d2, rgb2, 3 * N));
  BF_TRY(arena_alloc(ctx, &d_out, 2 * N));
  BF_TRY(estimate_rgb8_dev_impl(ctx, p, B, H, W, d1, d2, use_color, d_out, stats));
  BF_TRY(download(ctx, uv_out, d_out, 2 * N));
  API_SYNC(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// stage-level entry points (host pointers)
// ------------------------------------------------------------------------------------------------
int b200flow_rgb2gray(b200flow_ctx *ctx, const double *rgb, int H, int W, double *gray) {
  API_BEGIN(ctx);
  size_t N = (size_t)H * W;
  double *d_in, *d_out;
  BF_TRY(upload(ctx, &d_in, rgb, 3 * N));
  BF_TRY(arena_alloc(ctx, &d_out, N));
  BF_TRY(k_rgbf_to_gray(ctx, d_in, (long long)N, d_out));
  BF_TRY(download(ctx, gray, d_out, N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_rgb2lab(b200flow_ctx *ctx, const double *rgb, int H, int W, int scale_channels, double *lab) {
  API_BEGIN(ctx);
  size_t N = (size_t)H * W;
  double *d_in, *d_lab, *d_s, *d_out;
  BF_TRY(upload(ctx, &d_in, rgb, 3 * N));
  BF_TRY(arena_alloc(ctx, &d_lab, 3 * N));
  BF_TRY(arena_alloc(ctx, &d_s, 3 * N));
  BF_TRY(arena_alloc(ctx, &d_out, 3 * N));
  BF_TRY(k_rgbf_to_lab(ctx, d_in, (long long)N, d_lab));
  if (scale_channels) BF_TRY(k_minmax_scale(ctx, d_lab, d_s, 3, (long long)N, 0.0, 255.0));
  BF_TRY(k_interleave(ctx, scale_channels ? d_s : d_lab, d_out, 1, (long long)N, 3));
  BF_TRY(download(ctx, lab, d_out, 3 * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_scale_image(b200flow_ctx *ctx, const double *in, long long n, double lo, double hi, double *out) {
  API_BEGIN(ctx);
  if (n < 0) return set_err(ctx, B200FLOW_EINVAL, "n < 0");
  if (n == 0) return 0;
  double *d_in, *d_out;
  BF_TRY(upload(ctx, &d_in, in, (size_t)n));
  BF_TRY(arena_alloc(ctx, &d_out, (size_t)n));
  BF_TRY(k_minmax_scale(ctx, d_in, d_out, 1, n, lo, hi));
  BF_TRY(download(ctx, out, d_out, (size_t)n));
  API_SYNC(ctx);
  return 0;
}

int b200flow_rof_texture(b200flow_ctx *ctx, const double *img, int H, int W, int C, double theta, int iters, double alp,
                         double *out) {
  API_BEGIN(ctx);
  if (H < 1 || W < 1 || C < 1 || iters < 0) return set_err(ctx, B200FLOW_EINVAL, "bad ROF arguments");
  size_t N = (size_t)H * W;
  double *d_in, *d_pl, *d_tex, *d_out;
  BF_TRY(upload(ctx, &d_in, img, C * N));
  BF_TRY(arena_alloc(ctx, &d_pl, C * N));
  BF_TRY(arena_alloc(ctx, &d_tex, C * N));
  BF_TRY(arena_alloc(ctx, &d_out, C * N));
  BF_TRY(k_deinterleave(ctx, d_in, d_pl, 1, (long long)N, C));
  BF_TRY(k_rof_texture(ctx, d_pl, d_tex, 1, C, H, W, theta, iters, alp));
  BF_TRY(k_interleave(ctx, d_tex, d_out, 1, (long long)N, C));
  BF_TRY(download(ctx, out, d_out, C * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_pyramid(b200flow_ctx *ctx, const double *img, int H, int W, int C, int levels, const double *f, int fs,
                     double ratio, double **outs, int *Hs, int *Ws) {
  API_BEGIN(ctx);
  if (H < 1 || W < 1 || C < 1 || levels < 1 || !Hs || !Ws) return set_err(ctx, B200FLOW_EINVAL, "bad pyramid arguments");
  if (!(ratio > 0.0) || ratio > 1.0) return set_err(ctx, B200FLOW_EINVAL, "ratio %g out of range (0, 1]", ratio);
  Hs[0] = H; Ws[0] = W;
  for (int l = 1; l < levels; ++l) { Hs[l] = level_size(Hs[l - 1], ratio); Ws[l] = level_size(Ws[l - 1], ratio); }
  if (!outs) return 0;
  if (levels > 1 && (!f || fs < 1 || fs > 9 || (fs & 1) == 0))
    return set_err(ctx, B200FLOW_EINVAL, "pyramid filter must be an odd square kernel of size <= 9, got %d", fs);
  size_t N = (size_t)H * W;
  double *d_in, *prev;
  BF_TRY(upload(ctx, &d_in, img, C * N));
  BF_TRY(arena_alloc(ctx, &prev, C * N));
  BF_TRY(k_deinterleave(ctx, d_in, prev, 1, (long long)N, C));
  if (outs[0]) BF_TRY(download(ctx, outs[0], d_in, C * N));   // level 0 is an exact copy
  for (int l = 1; l < levels; ++l) {
    double *cur;
    size_t n = (size_t)Hs[l] * Ws[l];
    BF_TRY(arena_alloc(ctx, &cur, C * n));
    BF_TRY(k_gauss_resize(ctx, prev, cur, C, Hs[l - 1], Ws[l - 1], Hs[l], Ws[l], f, fs));
    if (outs[l]) {
      double *il;
      BF_TRY(arena_alloc(ctx, &il, C * n));
      BF_TRY(k_interleave(ctx, cur, il, 1, (long long)n, C));
      BF_TRY(download(ctx, outs[l], il, C * n));
    }
    prev = cur;
  }
  API_SYNC(ctx);
  return 0;
}

int b200flow_resample_flow(b200flow_ctx *ctx, const double *uv, int h, int w, int H, int W, double *out) {
  API_BEGIN(ctx);
  if (h < 1 || w < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad resample sizes");
  double *d_in, *d_out;
  BF_TRY(upload(ctx, &d_in, uv, (size_t)2 * h * w));
  BF_TRY(arena_alloc(ctx, &d_out, (size_t)2 * H * W));
  BF_TRY(k_resample_flow(ctx, (const double2 *)d_in, (double2 *)d_out, 1, h, w, H, W));
  BF_TRY(download(ctx, out, d_out, (size_t)2 * H * W));
  API_SYNC(ctx);
  return 0;
}

int b200flow_partial_deriv(b200flow_ctx *ctx, const double *images, const double *uv, int H, int W, int interp,
                           const double filt[5], double blend, double *It, double *Ix, double *Iy) {
  return b200flow_partial_deriv_mc(ctx, images, uv, H, W, 1, interp, filt, blend, It, Ix, Iy);
}

int b200flow_partial_deriv_mc(b200flow_ctx *ctx, const double *images, const double *uv, int H, int W, int NC, int interp,
                              const double filt[5], double blend, double *It, double *Ix, double *Iy) {
  API_BEGIN(ctx);
  if (interp < 0 || interp > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown interpolation method: %d", interp);
  if (H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad image size");
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  size_t N = (size_t)H * W;
  double *d_img, *d_pl, *d_uv, *I1x, *I1y, *dIt, *dIx, *dIy, *d_il;
  double4 *src2;
  BF_TRY(upload(ctx, &d_img, images, 2 * NC * N));
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_pl, 2 * NC * N));
  BF_TRY(arena_alloc(ctx, &I1x, NC * N));
  BF_TRY(arena_alloc(ctx, &I1y, NC * N));
  BF_TRY(arena_alloc(ctx, &src2, NC * N));
  BF_TRY(arena_alloc(ctx, &dIt, NC * N));
  BF_TRY(arena_alloc(ctx, &dIx, NC * N));
  BF_TRY(arena_alloc(ctx, &dIy, NC * N));
  BF_TRY(arena_alloc(ctx, &d_il, 3 * NC * N));
  BF_TRY(k_deinterleave(ctx, d_img, d_pl, 1, (long long)N, 2 * NC));
  BF_TRY(k_level_prep(ctx, d_pl, 2 * NC * (long long)N, 1, NC, H, W, interp, filt, I1x, I1y, src2));
  PenaltySet ps;
  memset(&ps, 0, sizeof ps);
  LinSys none;
  memset(&none, 0, sizeof none);
  BF_TRY(k_warp_assemble(ctx, d_pl, 2 * NC * (long long)N, NC, I1x, I1y, src2, (const double2 *)d_uv, nullptr, 1, H, W,
                         interp, blend, ps, none, dIt, dIx, dIy));
  // planar [NC][H][W] -> the (H, W, NC) interleaved layout of the reference's arrays
  BF_TRY(k_interleave(ctx, dIt, d_il, 1, (long long)N, NC));
  BF_TRY(k_interleave(ctx, dIx, d_il + NC * N, 1, (long long)N, NC));
  BF_TRY(k_interleave(ctx, dIy, d_il + 2 * NC * N, 1, (long long)N, NC));
  BF_TRY(download(ctx, It, d_il, NC * N));
  BF_TRY(download(ctx, Ix, d_il + NC * N, NC * N));
  BF_TRY(download(ctx, Iy, d_il + 2 * NC * N, NC * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_robust_eval(b200flow_ctx *ctx, b200flow_penalty pen, int d_type, const double *x, long long n, double *y) {
  API_BEGIN(ctx);
  if (n < 0) return set_err(ctx, B200FLOW_EINVAL, "n < 0");
  if (pen.kind < 0 || pen.kind > 9) return set_err(ctx, B200FLOW_EINVAL, "Unknown penalty kind %d", pen.kind);
  if (d_type < 0 || d_type > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown d_type: %d", d_type);
  if (n == 0) return 0;
  double *d_x, *d_y;
  BF_TRY(upload(ctx, &d_x, x, (size_t)n));
  BF_TRY(arena_alloc(ctx, &d_y, (size_t)n));
  BF_TRY(k_robust_eval(ctx, pen, d_type, d_x, n, d_y));
  BF_TRY(download(ctx, y, d_y, (size_t)n));
  API_SYNC(ctx);
  return 0;
}

static int assemble_host(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv, const double *duv,
                         const double *It, const double *Ix, const double *Iy, int H, int W, int NC, LinSys *sys) {
  if (!p) return set_err(ctx, B200FLOW_EINVAL, "params is NULL");
  if (H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad image size");
  if (p->method != B200FLOW_HS && !(alpha >= 0.0 && alpha <= 1.0))
    return set_err(ctx, B200FLOW_EINVAL, "Invalid GNC alpha: %g", alpha);
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  size_t N = (size_t)H * W;
  double *d_uv, *d_duv = nullptr, *dIt, *dIx, *dIy;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  if (duv) BF_TRY(upload(ctx, &d_duv, duv, 2 * N));
  BF_TRY(upload(ctx, &dIt, It, NC * N));
  BF_TRY(upload(ctx, &dIx, Ix, NC * N));
  BF_TRY(upload(ctx, &dIy, Iy, NC * N));
  if (NC > 1) {     // (H, W, NC) interleaved -> planar
    double *pl;
    BF_TRY(arena_alloc(ctx, &pl, 3 * NC * N));
    BF_TRY(k_deinterleave(ctx, dIt, pl, 1, (long long)N, NC));
    BF_TRY(k_deinterleave(ctx, dIx, pl + NC * N, 1, (long long)N, NC));
    BF_TRY(k_deinterleave(ctx, dIy, pl + 2 * NC * N, 1, (long long)N, NC));
    dIt = pl; dIx = pl + NC * N; dIy = pl + 2 * NC * N;
  }
  BF_TRY(alloc_linsys(ctx, 1, H, W, sys));
  PenaltySet ps = make_penalty_set(p, alpha);
  BF_TRY(k_assemble_from_deriv(ctx, dIt, dIx, dIy, NC, (const double2 *)d_uv, (const double2 *)d_duv, 1, H, W, ps, *sys));
  return 0;
}

int b200flow_operator_apply(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv,
                            const double *duv, const double *It, const double *Ix, const double *Iy, int H, int W,
                            const double *x, double *Ax, double *b, double *diag) {
  return b200flow_operator_apply_mc(ctx, p, alpha, uv, duv, It, Ix, Iy, H, W, 1, x, Ax, b, diag);
}

int b200flow_operator_apply_mc(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv,
                               const double *duv, const double *It, const double *Ix, const double *Iy, int H, int W,
                               int NC, const double *x, double *Ax, double *b, double *diag) {
  API_BEGIN(ctx);
  LinSys sys;
  BF_TRY(assemble_host(ctx, p, alpha, uv, duv, It, Ix, Iy, H, W, NC, &sys));
  size_t N = (size_t)H * W;
  double *d_x = nullptr, *d_Ax = nullptr, *d_diag = nullptr;
  if (x && Ax) {
    BF_TRY(upload(ctx, &d_x, x, 2 * N));
    BF_TRY(arena_alloc(ctx, &d_Ax, 2 * N));
  }
  if (diag) BF_TRY(arena_alloc(ctx, &d_diag, 2 * N));
  if (d_Ax || d_diag) BF_TRY(k_operator_apply(ctx, sys, (const double2 *)d_x, (double2 *)d_Ax, (double2 *)d_diag));
  if (d_Ax) BF_TRY(download(ctx, Ax, d_Ax, 2 * N));
  if (d_diag) BF_TRY(download(ctx, diag, d_diag, 2 * N));
  if (b) BF_TRY(download(ctx, b, (const double *)sys.rhs, 2 * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_solve_increment(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv,
                             const double *duv, const double *It, const double *Ix, const double *Iy, int H, int W,
                             double *x, int *iters, double *relres) {
  return b200flow_solve_increment_mc(ctx, p, alpha, uv, duv, It, Ix, Iy, H, W, 1, x, iters, relres);
}

int b200flow_solve_increment_mc(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv,
                                const double *duv, const double *It, const double *Ix, const double *Iy, int H, int W,
                                int NC, double *x, int *iters, double *relres) {
  API_BEGIN(ctx);
  LinSys sys;
  BF_TRY(assemble_host(ctx, p, alpha, uv, duv, It, Ix, Iy, H, W, NC, &sys));
  if (!(p->tol > 0.0) || p->maxit < 1) return set_err(ctx, B200FLOW_EINVAL, "solver tol/maxit invalid");
  if (p->solver < B200FLOW_SOLVER_EXACT || p->solver > B200FLOW_SOLVER_FP32_IC)
    return set_err(ctx, B200FLOW_EINVAL, "Unknown solver: %d", p->solver);
  size_t N = (size_t)H * W;
  PcgWork w;
  double2 *d_x;
  BF_TRY(pcg_work_alloc(ctx, 1, H, W, &w));
  BF_TRY(arena_alloc(ctx, &d_x, N));
  int rc = k_pcg_solve(ctx, sys, w, d_x, p->tol, p->maxit, pcg_mode_of(p->solver), iters, relres, true);
  if (rc < 0 && rc != B200FLOW_ENOCONV) return rc;
  BF_TRY(download(ctx, x, (const double *)d_x, 2 * N));
  API_SYNC(ctx);
  return rc;
}

// the reference's _solve_linear_system(A, b, uv_shape) takes ANY right-hand side (base.py:87-114): same system, caller's b
int b200flow_solve_rhs_mc(b200flow_ctx *ctx, const b200flow_params *p, double alpha, const double *uv, const double *duv,
                          const double *It, const double *Ix, const double *Iy, int H, int W, int NC, const double *rhs,
                          double *x, int *iters, double *relres) {
  API_BEGIN(ctx);
  if (!rhs) return set_err(ctx, B200FLOW_EINVAL, "rhs is NULL");
  LinSys sys;
  BF_TRY(assemble_host(ctx, p, alpha, uv, duv, It, Ix, Iy, H, W, NC, &sys));
  if (!(p->tol > 0.0) || p->maxit < 1) return set_err(ctx, B200FLOW_EINVAL, "solver tol/maxit invalid");
  if (p->solver < B200FLOW_SOLVER_EXACT || p->solver > B200FLOW_SOLVER_FP32_IC)
    return set_err(ctx, B200FLOW_EINVAL, "Unknown solver: %d", p->solver);
  size_t N = (size_t)H * W;
  BF_CUDA(ctx, cudaMemcpyAsync(sys.rhs, rhs, 2 * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));   // after the assembly
  PcgWork w;
  double2 *d_x;
  BF_TRY(pcg_work_alloc(ctx, 1, H, W, &w));
  BF_TRY(arena_alloc(ctx, &d_x, N));
  int rc = k_pcg_solve(ctx, sys, w, d_x, p->tol, p->maxit, pcg_mode_of(p->solver), iters, relres, true);
  if (rc < 0 && rc != B200FLOW_ENOCONV) return rc;
  BF_TRY(download(ctx, x, (const double *)d_x, 2 * N));
  API_SYNC(ctx);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// evaluation / export edges
// ------------------------------------------------------------------------------------------------
int b200flow_flow_error_dev(b200flow_ctx *ctx, const double *uv_dev, const double *gt_dev, int B, int H, int W, int border,
                            double *result /*host [B][4]*/) {
  API_BEGIN(ctx);
  if (!uv_dev || !gt_dev || !result || B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad arguments");
  double *d_res;
  BF_TRY(arena_alloc(ctx, &d_res, (size_t)4 * B));
  BF_TRY(k_flow_error(ctx, (const double2 *)uv_dev, (const double2 *)gt_dev, B, H, W, border, d_res));
  BF_TRY(download(ctx, result, d_res, (size_t)4 * B));
  API_SYNC(ctx);
  return 0;
}

int b200flow_flow_error(b200flow_ctx *ctx, const double *uv, const double *gt, int B, int H, int W, int border,
                        double *result) {
  API_BEGIN(ctx);
  if (!uv || !gt || !result || B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad arguments");
  size_t N = (size_t)B * H * W;
  double *d_uv, *d_gt, *d_res;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(upload(ctx, &d_gt, gt, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_res, (size_t)4 * B));
  BF_TRY(k_flow_error(ctx, (const double2 *)d_uv, (const double2 *)d_gt, B, H, W, border, d_res));
  BF_TRY(download(ctx, result, d_res, (size_t)4 * B));
  API_SYNC(ctx);
  return 0;
}

int b200flow_flow_to_color(b200flow_ctx *ctx, const double *uv, int B, int H, int W, double max_flow, unsigned char *rgb) {
  API_BEGIN(ctx);
  if (!uv || !rgb || B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad arguments");
  size_t N = (size_t)B * H * W;
  double *d_uv;
  unsigned char *d_rgb;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_rgb, 3 * N));
  BF_TRY(k_flow_to_color(ctx, (const double2 *)d_uv, B, H, W, max_flow, d_rgb));
  BF_TRY(download(ctx, rgb, d_rgb, 3 * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_flow_to_flo(b200flow_ctx *ctx, const double *uv, int B, int H, int W, unsigned char *out) {
  API_BEGIN(ctx);
  if (!uv || !out || B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad arguments");
  size_t N = (size_t)B * H * W, bytes = (size_t)B * 12 + 8 * N;
  double *d_uv;
  unsigned char *d_out;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_out, bytes));
  BF_TRY(k_flow_to_flo(ctx, (const double2 *)d_uv, B, H, W, d_out));
  BF_TRY(download(ctx, out, d_out, bytes));
  API_SYNC(ctx);
  return 0;
}

// Diagnostic (bench / ncu only, not on the flow path): time `reps` solves of a synthetic batch of B five-point systems
// of H x W pixels (random SPD coefficients spanning `decades` decades) run for exactly `iters` iterations each.
__global__ void synth_system_kernel(bf::LinSys S, double decades, unsigned seed) {
  long long n = (long long)S.B * S.H * S.W;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x = (int)(i % S.W), y = (int)((i / S.W) % S.H);
  auto rnd = [&](unsigned k) {
    unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed + k * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  };
  double ix = 20.0 * (rnd(0) - 0.5), iy = 20.0 * (rnd(1) - 0.5), d = 1e-7 * pow(10.0, decades * rnd(2));   // weak data term: Poisson-like, hundreds of iterations
  S.D[i] = make_double2(d * ix * ix, d * iy * iy);
  S.a12[i] = d * ix * iy;
  double wh = x + 1 < S.W ? pow(10.0, decades * rnd(3)) : 0.0, wv = y + 1 < S.H ? pow(10.0, decades * rnd(4)) : 0.0;
  S.WH[i] = make_double2(wh, x + 1 < S.W ? pow(10.0, decades * rnd(5)) : 0.0);
  S.WV[i] = make_double2(wv, y + 1 < S.H ? pow(10.0, decades * rnd(6)) : 0.0);
  S.rhs[i] = make_double2(rnd(7) - 0.5, rnd(8) - 0.5);
}

int b200flow_debug_pcg_bench(b200flow_ctx *ctx, int B, int H, int W, int solver, int iters, int reps, double decades,
                             double *ms_per_solve, long long *iters_done) {
  API_BEGIN(ctx);
  if (B < 1 || H < 1 || W < 1 || iters < 1 || reps < 1) return set_err(ctx, B200FLOW_EINVAL, "bad arguments");
  LinSys sys;
  PcgWork w;
  double2 *x;
  long long *dstats;
  BF_TRY(alloc_linsys(ctx, B, H, W, &sys));
  BF_TRY(pcg_work_alloc(ctx, B, H, W, &w));
  BF_TRY(arena_alloc(ctx, &x, (size_t)B * H * W));
  BF_TRY(arena_alloc(ctx, &dstats, 4));
  long long n = (long long)B * H * W;
  BF_LAUNCH(ctx, synth_system_kernel, (unsigned)cdiv(n, 256), 256, 0, sys, decades, 12345u);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  BF_TRY(k_pcg_solve_async(ctx, sys, w, x, 1e-30, iters, pcg_mode_of(solver), nullptr));   // warm-up
  BF_CUDA(ctx, cudaMemsetAsync(dstats, 0, 4 * sizeof(long long), ctx->stream));
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < reps; ++r) BF_TRY(k_pcg_solve_async(ctx, sys, w, x, 1e-30, iters, pcg_mode_of(solver), dstats));
  cudaEventRecord(e1, ctx->stream);
  long long hs[4];
  BF_CUDA(ctx, cudaMemcpyAsync(hs, dstats, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
  BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_per_solve) *ms_per_solve = ms / reps;
  if (iters_done) *iters_done = hs[0] / reps;
  return 0;
}

int b200flow_median_filter(b200flow_ctx *ctx, const double *uv, int H, int W, int kh, int kw, double *out) {
  API_BEGIN(ctx);
  if (H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad image size");
  size_t N = (size_t)H * W;
  double *d_in, *d_out;
  BF_TRY(upload(ctx, &d_in, uv, 2 * N));
  BF_TRY(arena_alloc(ctx, &d_out, 2 * N));
  BF_TRY(k_median_uv(ctx, (const double2 *)d_in, nullptr, 0, nullptr, (double2 *)d_out, 1, H, W, kh, kw, 1));
  BF_TRY(download(ctx, out, d_out, 2 * N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_detect_occlusion(b200flow_ctx *ctx, const double *uv, const double *images, int H, int W, double sigma_d,
                              double sigma_i, double *occ) {
  return b200flow_detect_occlusion_mc(ctx, uv, images, H, W, 1, sigma_d, sigma_i, occ);
}

int b200flow_detect_occlusion_mc(b200flow_ctx *ctx, const double *uv, const double *images, int H, int W, int NC,
                                 double sigma_d, double sigma_i, double *occ) {
  API_BEGIN(ctx);
  if (H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad image size");
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  size_t N = (size_t)H * W;
  double *d_uv, *d_img, *d_pl, *d_occ;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(upload(ctx, &d_img, images, 2 * NC * N));
  BF_TRY(arena_alloc(ctx, &d_pl, 2 * NC * N));
  BF_TRY(arena_alloc(ctx, &d_occ, N));
  BF_TRY(k_deinterleave(ctx, d_img, d_pl, 1, (long long)N, 2 * NC));
  BF_TRY(k_occlusion(ctx, (const double2 *)d_uv, d_pl, 2 * NC * (long long)N, NC, 1, H, W, sigma_d, sigma_i, d_occ));
  BF_TRY(download(ctx, occ, d_occ, N));
  API_SYNC(ctx);
  return 0;
}

int b200flow_weighted_median(b200flow_ctx *ctx, const double *uv, const double *color, const double *occ, int H, int W,
                             int C, int hsz, double sigma_i, double *out) {
  API_BEGIN(ctx);
  if (H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad image size");
  if (C < 1 || C > 3) return set_err(ctx, B200FLOW_EINVAL, "weighted median supports 1..3 colour channels, got %d", C);
  size_t N = (size_t)H * W;
  double *d_uv, *d_col, *d_pl, *d_occ, *d_out;
  BF_TRY(upload(ctx, &d_uv, uv, 2 * N));
  BF_TRY(upload(ctx, &d_col, color, C * N));
  BF_TRY(upload(ctx, &d_occ, occ, N));
  BF_TRY(arena_alloc(ctx, &d_pl, C * N));
  BF_TRY(arena_alloc(ctx, &d_out, 2 * N));
  BF_TRY(k_deinterleave(ctx, d_col, d_pl, 1, (long long)N, C));
  BF_TRY(k_weighted_median(ctx, (const double2 *)d_uv, nullptr, d_pl, C, d_occ, 1, H, W, hsz, sigma_i, (double2 *)d_out));
  BF_TRY(download(ctx, out, d_out, 2 * N));
  API_SYNC(ctx);
  return 0;
}

}  // extern "C"
