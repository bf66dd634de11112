// pipeline.cu -- the coarse-to-fine control loops of the three method drivers, run natively so that a batch of
// B same-size frame pairs needs no host<->device synchronisation between the upload and the final download.
// Replaces HSOpticalFlow.compute_flow/compute_flow_base (hs.py:49-142), BAOpticalFlow (ba.py:57-206) and
// ClassicNLOpticalFlow (classic_nl.py:89-277).  Every buffer is allocated once at full resolution from the
// context arena and reused by every level / warp iteration.
#include <algorithm>
#include <utility>
#include "kernels.cuh"

namespace bf {

namespace {

struct Pyramid {
  std::vector<int> H, W;
  std::vector<double *> lv;   // lv[l]: [P][H[l]][W[l]] planes
};

int build_pyramid(b200flow_ctx *ctx, double *img, int P, int H, int W, int levels, double spacing, Pyramid *out) {
  double taps[81];
  int ks;
  gaussian_taps(spacing, taps, &ks);
  double ratio = 1.0 / spacing;
  out->H.assign(1, H);
  out->W.assign(1, W);
  out->lv.assign(1, img);
  for (int l = 1; l < levels; ++l) {
    int h = out->H.back(), w = out->W.back();
    int hn = level_size(h, ratio), wn = level_size(w, ratio);
    double *dst;
    BF_TRY(arena_alloc(ctx, &dst, (size_t)P * hn * wn));
    BF_TRY(k_gauss_resize(ctx, out->lv.back(), dst, P, h, w, hn, wn, taps, ks));
    out->H.push_back(hn);
    out->W.push_back(wn);
    out->lv.push_back(dst);
  }
  return 0;
}

__global__ void fill_int_kernel(int *p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// per-stage CUDA-event timing (only when ctx->timing): events are recorded on the stream and resolved once at the end
// as intervals relative to a base event, so that the spans of concurrent sub-batches can be merged (union per category)
struct Interval { int cat; float a, b; };
struct StageTimer {
  b200flow_ctx *ctx = nullptr;
  bool on = false;
  struct Span { cudaEvent_t a, b; int cat; };
  std::vector<Span> spans;
  StageTimer() {}
  void init(b200flow_ctx *c) { ctx = c; on = c->timing; }
  ~StageTimer() { clear(); }
  StageTimer(const StageTimer &) = delete;
  StageTimer &operator=(const StageTimer &) = delete;
  double bytes[B200FLOW_K_COUNT] = {0};   // algorithmic bytes per kernel group (always accumulated)
  int calls[B200FLOW_K_COUNT] = {0};
  void begin(int cat, double nbytes = 0.0) {
    bytes[cat] += nbytes;
    calls[cat] += 1;
    if (!on) return;
    Span s;
    cudaEventCreate(&s.a);
    cudaEventCreate(&s.b);
    s.cat = cat;
    cudaEventRecord(s.a, ctx->stream);
    spans.push_back(s);
  }
  void end() {
    if (!on) return;
    cudaEventRecord(spans.back().b, ctx->stream);
  }
  void resolve(cudaEvent_t base, std::vector<Interval> *out) {
    for (auto &s : spans) {
      float ta = 0.f, tb = 0.f;
      if (on && cudaEventElapsedTime(&ta, base, s.a) == cudaSuccess && cudaEventElapsedTime(&tb, base, s.b) == cudaSuccess)
        out->push_back({s.cat, ta, tb});
    }
    clear();
  }
  void clear() {
    for (auto &s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    spans.clear();
  }
};
// total length of the union of the intervals of one category (ms)
static double union_ms(const std::vector<Interval> &v, const int *cats) {
  std::vector<std::pair<float, float>> w;
  for (auto &i : v) {
    bool take = false;
    for (const int *c = cats; *c >= 0; ++c) take |= i.cat == *c;
    if (take && i.b > i.a) w.push_back({i.a, i.b});
  }
  std::sort(w.begin(), w.end());
  double total = 0.0;
  float ca = 0.f, cb = -1.f;
  for (auto &i : w) {
    if (cb < ca || i.first > cb) { if (cb > ca) total += cb - ca; ca = i.first; cb = i.second; }
    else if (i.second > cb) cb = i.second;
  }
  if (cb > ca) total += cb - ca;
  return total;
}
// the four coarse stages reported since round 1, as unions of kernel groups
static const int STAGE_PRE[] = {B200FLOW_K_ROF, B200FLOW_K_PYRAMID, -1};
static const int STAGE_WARP[] = {B200FLOW_K_RESAMPLE, B200FLOW_K_LEVEL_PREP, B200FLOW_K_WARP_ASSEMBLE, -1};
static const int STAGE_SOLVE[] = {B200FLOW_K_SOLVER, -1};
static const int STAGE_FILTER[] = {B200FLOW_K_CLIP_ADD, B200FLOW_K_OCCLUSION, B200FLOW_K_WMEDIAN, B200FLOW_K_MEDIAN, B200FLOW_K_MISC, -1};

int check_params(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int C) {
  if (!p) return set_err(ctx, B200FLOW_EINVAL, "params is NULL");
  if (B < 1 || H < 1 || W < 1) return set_err(ctx, B200FLOW_EINVAL, "bad batch/size B=%d H=%d W=%d", B, H, W);
  if (p->method < 0 || p->method > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown method %d", p->method);
  if (p->interp < 0 || p->interp > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown interpolation method: %d", p->interp);
  if (p->solver < B200FLOW_SOLVER_EXACT || p->solver > B200FLOW_SOLVER_FP32_IC)
    return set_err(ctx, B200FLOW_EINVAL, "Unknown solver: %d", p->solver);
  if (!(p->pyramid_spacing > 1.0) || p->pyramid_spacing > 8.0)
    return set_err(ctx, B200FLOW_EINVAL, "pyramid_spacing %g out of range (1, 8]", p->pyramid_spacing);
  if (p->method != B200FLOW_HS && (!(p->gnc_pyramid_spacing > 1.0) || p->gnc_pyramid_spacing > 8.0))
    return set_err(ctx, B200FLOW_EINVAL, "gnc_pyramid_spacing %g out of range (1, 8]", p->gnc_pyramid_spacing);
  if (C < 0 || C > 3) return set_err(ctx, B200FLOW_EINVAL, "colour channels C=%d unsupported (0..3)", C);
  if (!(p->tol > 0.0) || p->maxit < 1) return set_err(ctx, B200FLOW_EINVAL, "solver tol/maxit invalid");
  const b200flow_penalty *pens[] = {&p->rho_su[0], &p->rho_su[1], &p->rho_sv[0], &p->rho_sv[1], &p->rho_d,
                                    &p->qua_su[0], &p->qua_su[1], &p->qua_sv[0], &p->qua_sv[1], &p->qua_d};
  for (auto q : pens)
    if (q->kind < 0 || q->kind > 9) return set_err(ctx, B200FLOW_EINVAL, "Unknown penalty kind %d", q->kind);
  return 0;
}

}  // namespace

PenaltySet make_penalty_set(const b200flow_params *p, double alpha) {
  PenaltySet ps;
  for (int i = 0; i < 2; ++i) {
    ps.rho_su[i] = p->rho_su[i]; ps.rho_sv[i] = p->rho_sv[i];
    ps.qua_su[i] = p->qua_su[i]; ps.qua_sv[i] = p->qua_sv[i];
  }
  ps.rho_d = p->rho_d; ps.qua_d = p->qua_d;
  ps.lambda = p->lambda; ps.lambda_q = p->lambda_q; ps.alpha = alpha;
  ps.hs = p->method == B200FLOW_HS;
  ps.hs_w = p->lambda / p->sigmaS2;
  ps.hs_d = 1.0 / p->sigmaD2;
  return ps;
}

int alloc_linsys(b200flow_ctx *ctx, int B, int H, int W, LinSys *s) {
  size_t n = (size_t)B * H * W;
  s->B = B; s->H = H; s->W = W;
  BF_TRY(arena_alloc(ctx, &s->D, n));
  BF_TRY(arena_alloc(ctx, &s->a12, n));
  BF_TRY(arena_alloc(ctx, &s->WH, n));
  BF_TRY(arena_alloc(ctx, &s->WV, n));
  BF_TRY(arena_alloc(ctx, &s->rhs, n));
  return 0;
}

// what pipeline_issue leaves behind for pipeline_finish (events are released by the destructor on every exit path)
struct PipelineRun {
  StageTimer tm;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long *dstats = nullptr;
  double *dlog = nullptr;         // display log: squared norms, one per solve (device)
  int nlog = 0;
  int solves = 0, launches0 = 0, solver_bytes = 0;
  PipelineRun() {}
  PipelineRun(const PipelineRun &) = delete;
  PipelineRun &operator=(const PipelineRun &) = delete;
  ~PipelineRun() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

struct RunResult {
  long long hstats[4] = {0, 0, 0, 0};
  std::vector<Interval> iv;       // kernel-group spans relative to the base event (ms)
  double bytes[B200FLOW_K_COUNT] = {0};
  int calls[B200FLOW_K_COUNT] = {0};
  int solver_bytes_per_pixel_iter = 0;
  float t0 = 0.f, t1 = 0.f;       // begin / end of the run relative to the base event
  int solves = 0, launches = 0;
};

// one kernel-group call under the stage timer (events only when ctx->timing; bytes and calls always)
#define TIMED(K, nbytes, call)       \
  do {                               \
    tm.begin((K), (nbytes));         \
    int rc_t_ = (call);              \
    tm.end();                        \
    if (rc_t_ < 0) return rc_t_;     \
  } while (0)

// Queues the whole coarse-to-fine loop of B pairs on ctx->stream; never synchronises (unless B200FLOW_TRACE is set).
// gray_planar: [B][2*NC][H][W] -- NC channels of frame 1, then NC channels of frame 2 (NC = 1 for gray frames)
static int pipeline_issue(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int NC, int C,
                          const double *gray_planar, const double *color_planar, const double2 *init, double2 *uv_out,
                          PipelineRun *run) {
  BF_TRY(check_params(ctx, p, B, H, W, C));
  if (NC < 1 || NC > 8) return set_err(ctx, B200FLOW_EINVAL, "channels per frame NC=%d unsupported (1..8)", NC);
  const int NP = 2 * NC;                                  // image planes per pair
  const long long HW = (long long)H * W;
  const size_t N = (size_t)B * HW;
  const bool hs = p->method == B200FLOW_HS, cnl = p->method == B200FLOW_CLASSICNL;
  run->launches0 = ctx->launches;
  const bool trace = getenv("B200FLOW_TRACE") != nullptr;
  StageTimer &tm = run->tm;
  tm.init(ctx);
  if (ctx->timing) {
    cudaEventCreate(&run->ev0);
    cudaEventCreate(&run->ev1);
    cudaEventRecord(run->ev0, ctx->stream);
  }

  // ---- pre-processing: texture or [0,255] scaling (joint over the two frames of a pair) ----
  double *pre;
  BF_TRY(arena_alloc(ctx, &pre, NP * N));
  // Horn-Schunck calls structure_texture_decomposition_rof(self.images) with the function's own defaults
  // (hs.py:66-67: theta 1/8, 100 iterations, alp 0.95), whatever self.alp says; BA / Classic+NL pass self.alp
  const int rof_iters = hs ? 100 : p->rof_iters;
  tm.begin(B200FLOW_K_ROF, p->texture > 0 ? (40.0 * rof_iters + 32.0) * NP * N : 32.0 * NP * N);
  if (p->texture > 0)
    BF_TRY(k_rof_texture(ctx, gray_planar, pre, B, NP, H, W, hs ? 1.0 / 8 : p->rof_theta, rof_iters, hs ? 0.95 : p->alp));
  else if (p->texture == 0) BF_TRY(k_minmax_scale(ctx, gray_planar, pre, B, NP * HW, 0.0, 255.0));
  else BF_CUDA(ctx, cudaMemcpyAsync(pre, gray_planar, NP * N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  tm.end();

  int levels = (hs || p->auto_level) ? auto_levels(H, W, p->pyramid_spacing) : p->pyramid_levels;
  if (p->pyramid_levels > 0 && !p->auto_level) levels = p->pyramid_levels;
  Pyramid pyr, gpyr, cpyr, gcpyr;
  tm.begin(B200FLOW_K_PYRAMID);
  BF_TRY(build_pyramid(ctx, pre, NP * B, H, W, levels, p->pyramid_spacing, &pyr));
  const bool use_color = cnl && color_planar != nullptr && C > 0;
  if (!hs) {
    BF_TRY(build_pyramid(ctx, pre, NP * B, H, W, p->gnc_pyramid_levels, p->gnc_pyramid_spacing, &gpyr));
    if (use_color) {
      BF_TRY(build_pyramid(ctx, const_cast<double *>(color_planar), C * B, H, W, levels, p->pyramid_spacing, &cpyr));
      BF_TRY(build_pyramid(ctx, const_cast<double *>(color_planar), C * B, H, W, p->gnc_pyramid_levels,
                           p->gnc_pyramid_spacing, &gcpyr));
    }
  }
  tm.end();
  {  // pyramid: every level reads its source level once and writes itself once (8 B per plane pixel each)
    const Pyramid *all[4] = {&pyr, &gpyr, &cpyr, &gcpyr};
    const int planes[4] = {NP * B, NP * B, C * B, C * B};
    for (int k = 0; k < 4; ++k)
      for (size_t l = 1; l < all[k]->lv.size(); ++l)
        tm.bytes[B200FLOW_K_PYRAMID] += 8.0 * planes[k] * ((double)all[k]->H[l - 1] * all[k]->W[l - 1] + (double)all[k]->H[l] * all[k]->W[l]);
  }

  // ---- work buffers at full resolution, reused by every level ----
  double2 *uvA, *uvB, *x, *cand = nullptr, *duv = nullptr;
  double *occ = nullptr, *I1x, *I1y, *It = nullptr, *Ix = nullptr, *Iy = nullptr, *nscratch = nullptr;
  double4 *src2;
  int *active = nullptr;
  long long *dstats;
  LinSys sys;
  PcgWork work;
  BF_TRY(arena_alloc(ctx, &uvA, N));
  BF_TRY(arena_alloc(ctx, &uvB, N));
  BF_TRY(arena_alloc(ctx, &x, N));
  BF_TRY(arena_alloc(ctx, &I1x, NC * N));
  BF_TRY(arena_alloc(ctx, &I1y, NC * N));
  BF_TRY(arena_alloc(ctx, &src2, NC * N));
  double4 *src2_tmp = nullptr;                            // out-of-place partner of the cubic-spline prefilter
  if (p->interp == B200FLOW_INTERP_CUBIC) BF_TRY(arena_alloc(ctx, &src2_tmp, NC * N));
  BF_TRY(alloc_linsys(ctx, B, H, W, &sys));
  BF_TRY(pcg_work_alloc(ctx, B, H, W, &work));
  BF_TRY(arena_alloc(ctx, &dstats, 4));
  BF_CUDA(ctx, cudaMemsetAsync(dstats, 0, 4 * sizeof(long long), ctx->stream));
  if (cnl) {
    BF_TRY(arena_alloc(ctx, &cand, N));
    BF_TRY(arena_alloc(ctx, &occ, N));
  }
  if (hs) {
    BF_TRY(arena_alloc(ctx, &active, (size_t)B));
    BF_TRY(arena_alloc(ctx, &nscratch, (size_t)B * 64));
  }
  constexpr int LOG_CAP = 4096;
  ctx->log.clear();
  if (ctx->log_on && B == 1) BF_TRY(arena_alloc(ctx, &run->dlog, (size_t)LOG_CAP));
  if (!hs && p->max_linear > 1) {
    BF_TRY(arena_alloc(ctx, &duv, N));
    BF_TRY(arena_alloc(ctx, &It, NC * N));
    BF_TRY(arena_alloc(ctx, &Ix, NC * N));
    BF_TRY(arena_alloc(ctx, &Iy, NC * N));
  }

  double2 *cur = uvA, *nxt = uvB;
  int ch = H, cw = W;     // size of the flow currently held in `cur`
  if (init) BF_CUDA(ctx, cudaMemcpyAsync(cur, init, N * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
  else BF_CUDA(ctx, cudaMemsetAsync(cur, 0, N * sizeof(double2), ctx->stream));

  const int mh = p->median_h, mw = p->median_w;
  const bool have_median = mh > 0 && mw > 0;
  const int pcg_mode = pcg_mode_of(p->solver);
  int solves = 0;

  const int gnc_stages = hs ? 1 : p->gnc_iters;
  double alpha = p->alpha0;
  for (int ignc = 0; ignc < gnc_stages; ++ignc) {
    const Pyramid &ip = (ignc == 0) ? pyr : gpyr;
    const Pyramid &cp = (ignc == 0) ? cpyr : gcpyr;
    const int nl = (int)ip.lv.size() < ((ignc == 0) ? levels : p->gnc_pyramid_levels)
                       ? (int)ip.lv.size() : ((ignc == 0) ? levels : p->gnc_pyramid_levels);
    if (!hs && !(alpha >= 0.0 && alpha <= 1.0)) return set_err(ctx, B200FLOW_EINVAL, "Invalid GNC alpha: %g", alpha);
    for (int l = nl - 1; l >= 0; --l) {
      const int h = ip.H[l], w = ip.W[l];
      const long long hw = (long long)h * w;
      const double *frames = ip.lv[l];
      const long long bstride = NP * hw;
      const double npx = (double)B * hw;   // pixels of this level over the batch: the unit of the per-pixel byte figures
      TIMED(B200FLOW_K_RESAMPLE, 16.0 * B * ((double)ch * cw + (double)hw), k_resample_flow(ctx, cur, nxt, B, ch, cw, h, w));
      std::swap(cur, nxt);
      ch = h; cw = w;
      // Hermite: read im1, im2 16, write I1x, I1y 16 + {Z, DX, DY, DXY} 32; spline: + two in-place prefilter passes of 32 B r/w
      TIMED(B200FLOW_K_LEVEL_PREP, (p->interp == B200FLOW_INTERP_CUBIC ? 64.0 + 128.0 : 64.0) * NC * npx,
            k_level_prep(ctx, frames, bstride, B, NC, h, w, p->interp, p->deriv_filter, I1x, I1y, src2, src2_tmp));
      sys.H = h; sys.W = w;
      if (hs) BF_LAUNCH(ctx, fill_int_kernel, (unsigned)cdiv(B, 128), 128, 0, active, B, 1);
      const int warps = hs ? p->max_warping_iters : p->max_iters;
      const int nlin = hs ? 1 : (ignc == 0 ? 1 : (p->max_linear < 1 ? 1 : p->max_linear));
      PenaltySet ps = make_penalty_set(p, alpha);
      for (int it = 0; it < warps; ++it) {
        const double2 *dcur = nullptr;    // duv of the current linearisation (zero on the first pass)
        for (int j = 0; j < nlin; ++j) {
          if (j == 0)
            TIMED(B200FLOW_K_WARP_ASSEMBLE, (88.0 + 56.0 * NC) * npx,
                  k_warp_assemble(ctx, frames, bstride, NC, I1x, I1y, src2, cur, nullptr, B, h, w, p->interp, p->blend, ps,
                                  sys, nlin > 1 ? It : nullptr, Ix, Iy));
          else
            TIMED(B200FLOW_K_WARP_ASSEMBLE, (104.0 + 24.0 * NC) * npx,
                  k_assemble_from_deriv(ctx, It, Ix, Iy, NC, cur, dcur, B, h, w, ps, sys));
          if (ctx->solver_stream) {          // concurrent sub-batches: the solver runs on the group's high-priority stream
            BF_CUDA(ctx, cudaEventRecord(ctx->ev_s0, ctx->stream));
            BF_CUDA(ctx, cudaStreamWaitEvent(ctx->solver_stream, ctx->ev_s0, 0));
            std::swap(ctx->stream, ctx->solver_stream);
          }
          tm.begin(B200FLOW_K_SOLVER);       // its bytes = pixel-iterations (device counter) x bytes per pixel-iteration
          int rc_solve = k_pcg_solve_async(ctx, sys, work, x, p->tol, p->maxit, pcg_mode, dstats);
          tm.end();
          if (ctx->solver_stream) {
            std::swap(ctx->stream, ctx->solver_stream);
            BF_CUDA(ctx, cudaEventRecord(ctx->ev_s1, ctx->solver_stream));
            BF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_s1, 0));
          }
          BF_TRY(rc_solve);
          if (trace) {   // B200FLOW_TRACE=1 (debug): per-solve time and per-system iteration counts; synchronises
            std::vector<int> fl(1 + 2 * B);
            cudaMemcpyAsync(fl.data(), work.flags, sizeof(int) * fl.size(), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            float ms = 0.f;
            if (tm.on) cudaEventElapsedTime(&ms, tm.spans.back().a, tm.spans.back().b);
            int mn = 1 << 30, mx = 0; long long sum = 0;
            for (int b = 0; b < B; ++b) { int v = fl[1 + B + b]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
            {
              std::vector<double> rr(B);
              cudaMemcpy(rr.data(), work.scal, sizeof(double) * B, cudaMemcpyDeviceToHost);
              for (int b = 0; b < B; ++b)
                if (fl[1 + b] != 1) fprintf(stderr, "[b200flow trace]   system %d NOT converged: status %d, relres %.3e after %d iterations\n", b, fl[1 + b], rr[b], fl[1 + B + b]);
            }
            fprintf(stderr, "[b200flow trace] gnc %d level %d (%dx%d) warp %d: solve %.3f ms, iters min %d mean %.1f max %d, "
                            "%.1f us/iter(max), alg %.0f GB/s\n", ignc, l, h, w, it, ms, mn, (double)sum / B, mx,
                    mx ? 1e3 * ms / mx : 0.0, ms > 0 ? (double)sum * hw * pcg_bytes_per_pixel_iter(pcg_mode) / (ms * 1e6) : 0.0);
          }
          solves++;
          if (run->dlog && run->nlog < LOG_CAP) {      // what the reference prints under display=True (HS: the unclipped norm)
            BF_TRY(k_delta_norm(ctx, x, hs ? nullptr : dcur, hs ? 0 : p->limit_update, hw, run->dlog + run->nlog));
            ctx->log.push_back({ignc, l, it, j, 0.0});
            run->nlog++;
          }
          const double median_bytes = 32.0 * npx, clip_bytes = 48.0 * npx;
          if (hs) {
            TIMED(B200FLOW_K_MISC, 16.0 * npx, k_hs_norm_gate(ctx, x, B, hw, active, nscratch));
            if (have_median && p->mf_iter >= 1) {     // hs.py:137-140: mf_iter passes, none at all when mf_iter < 1
              TIMED(B200FLOW_K_MEDIAN, median_bytes + 16.0 * npx,
                    k_median_uv(ctx, cur, x, p->limit_update, active, nxt, B, h, w, mh, mw, 1));
              for (int m = 1; m < p->mf_iter; ++m) {
                std::swap(cur, nxt);
                TIMED(B200FLOW_K_MEDIAN, median_bytes, k_median_uv(ctx, cur, nullptr, 0, active, nxt, B, h, w, mh, mw, 1));
              }
            } else {
              TIMED(B200FLOW_K_CLIP_ADD, clip_bytes, k_clip_add(ctx, cur, x, p->limit_update, active, nxt, hw, B));
            }
          } else if (cnl && have_median && use_color) {
            TIMED(B200FLOW_K_CLIP_ADD, clip_bytes, k_clip_add(ctx, cur, x, p->limit_update, nullptr, cand, hw, B));
            TIMED(B200FLOW_K_OCCLUSION, (24.0 + 16.0 * NC) * npx,
                  k_occlusion(ctx, cand, frames, bstride, NC, B, h, w, p->occ_sigma_d, p->occ_sigma_i, occ));
            TIMED(B200FLOW_K_WMEDIAN, (56.0 + 8.0 * C) * npx,
                  k_weighted_median(ctx, cand, cur, cp.lv[l], C, occ, B, h, w, p->area_hsz, p->sigma_i, nxt));
          } else if (have_median) {
            // BA (ba.py:197-201) and Classic+NL without a usable colour image (weighted_median.py:42-47: square mfsz[0])
            TIMED(B200FLOW_K_MEDIAN, median_bytes + 16.0 * npx,
                  k_median_uv(ctx, cur, x, p->limit_update, nullptr, nxt, B, h, w, mh, cnl ? mh : mw, 0));
          } else {
            TIMED(B200FLOW_K_CLIP_ADD, clip_bytes, k_clip_add(ctx, cur, x, p->limit_update, nullptr, nxt, hw, B));
          }
          if (j + 1 < nlin) {               // next linearisation pass sees duv = filtered - uv
            TIMED(B200FLOW_K_MISC, 48.0 * npx, k_sub(ctx, nxt, cur, duv, (long long)B * hw));
            dcur = duv;
          }
        }
        std::swap(cur, nxt);                 // uv = uv + duv
      }
    }
    if (!hs && p->gnc_iters > 1) {
      double na = 1.0 - (double)(ignc + 1) / (double)(p->gnc_iters - 1);
      alpha = alpha < na ? alpha : na;
      alpha = alpha > 0.0 ? alpha : 0.0;
    }
  }
  if (ch != H || cw != W) {
    // no level was run (min(H,W) < 16: auto levels <= 0) -- the reference returns init unchanged
    return set_err(ctx, B200FLOW_EINVAL, "internal: final flow size %dx%d != %dx%d", ch, cw, H, W);
  }
  if (hs && have_median && p->final_median) {   // final median (hs.py:95-97)
    TIMED(B200FLOW_K_MEDIAN, 32.0 * N, k_median_uv(ctx, cur, nullptr, 0, nullptr, nxt, B, H, W, mh, mw, 1));
    std::swap(cur, nxt);
  }
  BF_CUDA(ctx, cudaMemcpyAsync(uv_out, cur, N * sizeof(double2), cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->timing) cudaEventRecord(run->ev1, ctx->stream);
  run->dstats = dstats;
  run->solves = solves;
  run->solver_bytes = pcg_bytes_per_pixel_iter(pcg_mode);
  return 0;
}

// Waits for a queued run (only when its outcome is wanted) and collects the device-side statistics and stage spans.
static int pipeline_finish(b200flow_ctx *ctx, PipelineRun *run, bool want, cudaEvent_t base, RunResult *res) {
  res->solves = run->solves;
  res->launches = ctx->launches - run->launches0;
  res->solver_bytes_per_pixel_iter = run->solver_bytes;
  for (int k = 0; k < B200FLOW_K_COUNT; ++k) { res->bytes[k] = run->tm.bytes[k]; res->calls[k] = run->tm.calls[k]; }
  if (want || ctx->timing) {
    if (ctx->band.world > 1) BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // see download() in common.cuh
    BF_CUDA(ctx, cudaMemcpyAsync(res->hstats, run->dstats, sizeof res->hstats, cudaMemcpyDeviceToHost, ctx->stream));
    BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (run->nlog > 0) {
    std::vector<double> sq(run->nlog);
    BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BF_CUDA(ctx, cudaMemcpy(sq.data(), run->dlog, sizeof(double) * run->nlog, cudaMemcpyDeviceToHost));
    for (int k = 0; k < run->nlog; ++k) ctx->log[k].v = sqrt(sq[k]);
  }
  if (ctx->timing) {
    if (!base) base = run->ev0;
    run->tm.resolve(base, &res->iv);
    cudaEventElapsedTime(&res->t0, base, run->ev0);
    cudaEventElapsedTime(&res->t1, base, run->ev1);
  }
  return 0;
}

static void fill_stats(b200flow_stats *stats, const std::vector<RunResult> &rr, bool timing) {
  if (!stats) return;
  stats->solves = 0; stats->pcg_iters = 0; stats->not_converged = 0; stats->pcg_pixel_iters = 0; stats->kernel_launches = 0;
  std::vector<Interval> all;
  float t0 = 0.f, t1 = 0.f;
  bool first = true;
  for (auto &r : rr) {
    stats->solves = r.solves > stats->solves ? r.solves : stats->solves;              // every group runs the same loop
    stats->pcg_iters = r.hstats[0] > stats->pcg_iters ? r.hstats[0] : stats->pcg_iters; // slowest group
    stats->not_converged += (int)r.hstats[1];
    stats->pcg_pixel_iters += r.hstats[3];
    stats->kernel_launches += r.launches;
    all.insert(all.end(), r.iv.begin(), r.iv.end());
    if (first || r.t0 < t0) t0 = r.t0;
    if (first || r.t1 > t1) t1 = r.t1;
    first = false;
  }
  // with concurrent groups the spans of one category overlap in time: report the time during which at least one group
  // was in that stage (for one group this is the plain sum)
  stats->pre_ms = union_ms(all, STAGE_PRE); stats->warp_ms = union_ms(all, STAGE_WARP);
  stats->solver_ms = union_ms(all, STAGE_SOLVE); stats->filter_ms = union_ms(all, STAGE_FILTER);
  for (int k = 0; k < B200FLOW_K_COUNT; ++k) {
    const int one[2] = {k, -1};
    stats->kernel_ms[k] = union_ms(all, one);
    stats->kernel_bytes[k] = 0.0;
    stats->kernel_calls[k] = 0;
    for (auto &r : rr) { stats->kernel_bytes[k] += r.bytes[k]; stats->kernel_calls[k] += r.calls[k]; }
  }
  stats->kernel_bytes[B200FLOW_K_SOLVER] = 0.0;
  for (auto &r : rr) stats->kernel_bytes[B200FLOW_K_SOLVER] += (double)r.hstats[3] * r.solver_bytes_per_pixel_iter;
  stats->total_ms = timing ? (double)(t1 - t0) : 0.0;
}

static int ensure_subs(b200flow_ctx *ctx, int ns) {
  if (!ctx->ev_fork) BF_CUDA(ctx, cudaEventCreate(&ctx->ev_fork));
  while ((int)ctx->subs.size() < ns) {
    b200flow_ctx *c = new b200flow_ctx();
    c->device = ctx->device;
    c->num_sms = ctx->num_sms;
    c->parent = ctx;
    int pri_lo = 0, pri_hi = 0;
    cudaDeviceGetStreamPriorityRange(&pri_lo, &pri_hi);      // numerically lower = higher priority
    const bool prio = getenv("B200FLOW_NO_SOLVER_PRIORITY") == nullptr;
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, pri_lo) != cudaSuccess ||
        (prio && (cudaStreamCreateWithPriority(&c->solver_stream, cudaStreamNonBlocking, pri_hi) != cudaSuccess ||
                  cudaEventCreateWithFlags(&c->ev_s0, cudaEventDisableTiming) != cudaSuccess ||
                  cudaEventCreateWithFlags(&c->ev_s1, cudaEventDisableTiming) != cudaSuccess))) {
      delete c;
      return set_err(ctx, B200FLOW_ECUDA, "cudaStreamCreate for a sub-batch failed");
    }
    ctx->subs.push_back(c);
    cudaEvent_t e;
    BF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->ev_join.push_back(e);
  }
  return 0;
}

int run_pipeline(b200flow_ctx *ctx, const b200flow_params *p, int B, int H, int W, int NC, int C, const double *gray_planar,
                 const double *color_planar, const double2 *init, double2 *uv_out, b200flow_stats *stats) {
  int ns = ctx->nsplit < 1 ? 1 : ctx->nsplit;
  if (ns > B) ns = B;
  if (p && p->solver != B200FLOW_SOLVER_EXACT_IC && p->solver != B200FLOW_SOLVER_FP32_IC) ns = 1;     // only pcg_ic_kernel is sized for co-residency
  if (getenv("B200FLOW_TRACE")) ns = 1;
  std::vector<RunResult> rr(ns);
  if (ns == 1) {
    PipelineRun run;
    BF_TRY(pipeline_issue(ctx, p, B, H, W, NC, C, gray_planar, color_planar, init, uv_out, &run));
    BF_TRY(pipeline_finish(ctx, &run, stats != nullptr, nullptr, &rr[0]));
    fill_stats(stats, rr, ctx->timing);
    return 0;
  }
  // ---- concurrent sub-batches: group g = pairs [g B / ns, (g + 1) B / ns) on its own stream and arena ----
  BF_TRY(ensure_subs(ctx, ns));
  int nb = 0;
  BF_TRY(pcg_ic_grid(ctx, &nb));                              // fills ctx->ic_ctas_per_sm
  const long long HW = (long long)H * W;
  BF_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  std::vector<PipelineRun> runs(ns);
  for (int g = 0; g < ns; ++g) {
    b200flow_ctx *c = ctx->subs[g];
    const int b0 = (int)((long long)g * B / ns), b1 = (int)((long long)(g + 1) * B / ns);
    arena_reset(c);
    c->timing = ctx->timing;
    c->plain_solver_launch = true;
    c->solver_ctas_per_sm = ctx->solver_ctas_per_sm > 0 ? ctx->solver_ctas_per_sm
                                                        : (ctx->ic_ctas_per_sm / ns > 0 ? ctx->ic_ctas_per_sm / ns : 1);
    if (c->solver_ctas_per_sm * ns > ctx->ic_ctas_per_sm)
      return set_err(ctx, B200FLOW_EINVAL, "%d concurrent sub-batches x %d solver CTAs per SM exceed the %d the device holds",
                     ns, c->solver_ctas_per_sm, ctx->ic_ctas_per_sm);
    BF_CUDA(ctx, cudaStreamWaitEvent(c->stream, ctx->ev_fork, 0));
    int rc = pipeline_issue(c, p, b1 - b0, H, W, NC, C, gray_planar + (size_t)b0 * 2 * NC * HW,
                            color_planar ? color_planar + (size_t)b0 * C * HW : nullptr, init ? init + (size_t)b0 * HW : nullptr,
                            uv_out + (size_t)b0 * HW, &runs[g]);
    if (rc < 0) { ctx->err = c->err; return rc; }
    BF_CUDA(ctx, cudaEventRecord(ctx->ev_join[g], c->stream));
    BF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[g], 0));
  }
  for (int g = 0; g < ns; ++g) {
    int rc = pipeline_finish(ctx->subs[g], &runs[g], stats != nullptr, ctx->ev_fork, &rr[g]);
    if (rc < 0) { ctx->err = ctx->subs[g]->err; return rc; }
  }
  fill_stats(stats, rr, ctx->timing);
  return 0;
}

}  // namespace bf
