// common.cuh -- context, device arena, launch bookkeeping shared by every translation unit of libb200flow.so
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include "../../include/b200flow.h"

#define B200FLOW_MAX_BAND_RANKS 8
#define B200FLOW_BAND_RESERVED 4096      // head of the arena block: cross-GPU barrier state (solve_ic.cu BandSync) + ticket

// row-band split of one pair over several GPUs (b200flow_band_* in include/b200flow.h)
struct b200flow_band {
  int rank = 0, world = 1;
  char *base[B200FLOW_MAX_BAND_RANKS] = {nullptr};   // every rank's arena block as mapped in this process (base[rank] = own)
  bool opened[B200FLOW_MAX_BAND_RANKS] = {false};    // mapped through cudaIpcOpenMemHandle (to be closed)
  unsigned *ticket = nullptr;                        // device counter of band_push_kernel (inside the reserved head)
  long long min_pixels = 1 << 18;                    // levels below this many pixels are solved by every rank on its own
};

struct b200flow_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  std::string err;
  // grow-only device arena (bump allocation, reset per public call)
  struct Chunk { char *base; size_t size, off; };
  std::vector<Chunk> chunks;
  size_t high_water = 0, in_use = 0;
  // pinned staging for host<->device copies
  void *pinned = nullptr;
  size_t pinned_size = 0;
  // statistics
  int launches = 0;
  bool timing = false;
  // resident grid sizes of the persistent solver kernels on this device (filled once by solve.cu; 0 = not yet queried)
  int grid_pcg = 0, grid_mixed = 0, grid_ic = 0;
  int ic_ctas_per_sm = 0;           // occupancy of pcg_ic_kernel as reported by the runtime
  // display=True log of the single-pair drivers: one row per linear solve (GNC stage, pyramid level, warp, linearisation,
  // ||clip(x) - duv||_2), filled by run_pipeline when log_on, read back through b200flow_ctx_get_log
  struct LogRow { int gnc, level, it, lin; double v; };
  bool log_on = false;
  std::vector<LogRow> log;
  // table behind the generalized Charbonnier weight y^(a-1) of the assembly kernels (warp.cu, pow_table_kernel):
  // built on this context's stream whenever the exponent changes
  double *powtab = nullptr;
  double powtab_a = 0.0;
  // Concurrent sub-batches (pipeline.cu, run_pipeline_split): a batch of B pairs is cut into `nsplit` groups that run the
  // whole coarse-to-fine loop on their own stream and arena, so that the bandwidth-bound solver of one group overlaps the
  // issue-bound weighted median of another.  The persistent solver then takes solver_ctas_per_sm CTAs per SM (so that the
  // solvers of all groups can be resident together) and is launched without the cooperative attribute.
  int nsplit = 1;
  int solver_ctas_per_sm = 0;       // 0 = all the runtime allows
  bool plain_solver_launch = false; // children of a split context: <<<>>> instead of cudaLaunchCooperativeKernel
  cudaStream_t solver_stream = nullptr;   // children only: high-priority stream the persistent solver is launched on, so
                                          // that its CTAs are dispatched ahead of a sibling's pending weighted-median CTAs
  cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;
  b200flow_band band;
  b200flow_ctx *parent = nullptr;
  std::vector<b200flow_ctx *> subs; // child contexts (own stream + arena), created on demand
  cudaEvent_t ev_fork = nullptr;
  std::vector<cudaEvent_t> ev_join;
};

namespace bf {

inline int set_err(b200flow_ctx *ctx, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
inline int set_err(b200flow_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define BF_CUDA(ctx, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return bf::set_err((ctx), B200FLOW_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
                         cudaGetErrorString(e__));                                                 \
  } while (0)

#define BF_TRY(expr)             \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ < 0) return rc__;   \
  } while (0)

// kernel launch with bookkeeping; every kernel of this library goes through here
#define BF_LAUNCH(ctx, kern, grid, block, smem, ...)                                              \
  do {                                                                                             \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                 \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaPeekAtLastError();                                                       \
    if (e__ != cudaSuccess)                                                                        \
      return bf::set_err((ctx), B200FLOW_ECUDA, "launch %s failed at %s:%d: %s", #kern, __FILE__,  \
                         __LINE__, cudaGetErrorString(e__));                                       \
  } while (0)

inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- arena -------------------------------------------------------------------------------------
inline void arena_reset(b200flow_ctx *ctx) {
  // consolidate into one chunk sized for the high-water mark so steady state never cudaMallocs
  if (ctx->chunks.size() > 1 && ctx->band.world <= 1) {
    size_t total = 0;
    for (auto &c : ctx->chunks) { total += c.size; cudaFree(c.base); }
    ctx->chunks.clear();
    char *p = nullptr;
    if (cudaMalloc(&p, total) == cudaSuccess) ctx->chunks.push_back({p, total, 0});
  }
  for (auto &c : ctx->chunks) c.off = 0;
  if (ctx->band.world > 1 && !ctx->chunks.empty()) ctx->chunks[0].off = B200FLOW_BAND_RESERVED;   // barrier state lives here
  ctx->in_use = 0;
}

inline void *arena_alloc_raw(b200flow_ctx *ctx, size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes == 0) bytes = 256;
  for (auto &c : ctx->chunks) {
    if (c.off + bytes <= c.size) {
      void *p = c.base + c.off;
      c.off += bytes;
      ctx->in_use += bytes;
      return p;
    }
  }
  if (ctx->band.world > 1) return nullptr;   // row-band mode: the block is mapped into the peers and must not move or grow
  size_t sz = bytes;
  size_t grow = ctx->chunks.empty() ? (size_t(64) << 20) : ctx->chunks.back().size * 2;
  if (sz < grow) sz = grow;
  char *p = nullptr;
  if (cudaMalloc(&p, sz) != cudaSuccess) {
    sz = bytes;
    if (cudaMalloc(&p, sz) != cudaSuccess) return nullptr;
  }
  ctx->chunks.push_back({p, sz, bytes});
  ctx->in_use += bytes;
  return p;
}

template <typename T>
inline int arena_alloc(b200flow_ctx *ctx, T **out, size_t count) {
  *out = static_cast<T *>(arena_alloc_raw(ctx, count * sizeof(T)));
  if (!*out) return set_err(ctx, B200FLOW_ECUDA, "device arena: cudaMalloc of %zu bytes failed", count * sizeof(T));
  return 0;
}

inline int ensure_pinned(b200flow_ctx *ctx, size_t bytes) {
  if (ctx->pinned_size >= bytes) return 0;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_size = 0;
  BF_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
  ctx->pinned_size = bytes;
  return 0;
}

// host -> device through the context stream (pageable source is fine; cudaMemcpyAsync stages it)
template <typename T>
inline int upload(b200flow_ctx *ctx, T **dev, const T *host, size_t count) {
  BF_TRY(arena_alloc(ctx, dev, count));
  BF_CUDA(ctx, cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

template <typename T>
inline int download(b200flow_ctx *ctx, T *host, const T *dev, size_t count) {
  // Row-band mode: a device->host copy into PAGEABLE memory waits inside the driver for the stream's kernels; when two ranks
  // are emulated by two host threads of one process (tests), the other rank's launches then queue up behind that call while
  // this rank's persistent solver spins on the other rank's arrival.  cudaStreamSynchronize waits without that side effect.
  if (ctx->band.world > 1) BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  BF_CUDA(ctx, cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

// ---- small device helpers -----------------------------------------------------------------------
// scipy 'reflect' (half-sample symmetric: d c b a | a b c d | d c b a), any offset
__host__ __device__ inline int reflect_idx(int i, int n) {
  int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i >= n ? p - 1 - i : i;
}
// whole-sample mirror (d c b | a b c d | c b a): numpy.pad 'reflect', scipy spline 'mirror'
__host__ __device__ inline int mirror_idx(int i, int n) {
  if (n == 1) return 0;
  int p = 2 * (n - 1);
  i %= p;
  if (i < 0) i += p;
  return i >= n ? p - i : i;
}
__host__ __device__ inline int clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

}  // namespace bf
