// kernels.cuh -- device-level (all pointers are DEVICE pointers, every routine is batch-aware) launchers.
// Data layout in HBM:
//   image planes   double [P][H][W]            P = B*C planes, planar (de-interleaved on upload)
//   flow / vectors double2[B][H][W]            (u,v) interleaved = one 16-byte load per pixel
//   warp source    double4[B][H][W]            Hermite: {Z, DX, DY, DXY} of frame 2; spline/bilinear:
//                                              {c(I2), c(I2x), c(I2y), 0} -> one 32-byte sector per tap
//   linear system  D  double2 {a11,a22}, a12 double, WH double2 {wuh,wvh}, WV double2 {wuv,wvv},
//                  rhs double2 {bu,bv}
#pragma once
#include "common.cuh"

namespace bf {

struct PenaltySet {            // everything the assemble kernel needs to form the GNC-blended IRLS weights
  b200flow_penalty rho_su[2], rho_sv[2], rho_d, qua_su[2], qua_sv[2], qua_d;
  double lambda, lambda_q, alpha;
  int hs;                      // 1: Horn-Schunck form (unit weights * lambda/sigmaS2, d = 1/sigmaD2)
  double hs_w, hs_d;
  // generalized Charbonnier weights through a table (warp.cu, pow_tab): filled by the assembly launchers, never by callers
  const double *ptab = nullptr;  // device table built for exponent ptab_a - 1; null = exp/log path
  double ptab_a = 0.0;
  double pc[6] = {0, 0, 0, 0, 0, 0};   // binomial coefficients C(a-1, 1..6)
  int fast = 0;                  // Classic+NL-shaped set: edge weight = eq + er y^(a-1), data weight = dq + dr y^(a-1)
  double eq = 0, er = 0, s2s = 0, dq = 0, dr = 0, s2d = 0;
};

struct LinSys {                // matrix-free 2N x 2N system, B systems of H x W pixels
  int B, H, W;
  double2 *D;                  // {a11, a22}
  double *a12;
  double2 *WH, *WV;            // {wuh, wvh}, {wuv, wvv}
  double2 *rhs;
};

struct PcgWork {               // scratch vectors + reduction buffers for the persistent PCG kernel
  double2 *r, *p, *p2, *z, *Ap; // residual, search direction (ping-pong), preconditioned residual, A p
  float *Minv;                 // 3 planes: m11, m12, m22 (block-Jacobi inverse / IC pivot inverse, fp32) [3][B*H*W]
  float4 *ic_c0;               // IC preconditioner [B*H*W]: {i11, i12, i22, bf16x2 {wuh, wvh}} (inverted pivot block + right edge)
  unsigned *ic_cw;             // IC preconditioner [B*H*W]: bf16x2 {wuv, wvv} (down edge; 0 on every eighth row)
  double *partial;             // [3][B][grid]
  double *scal;                // per-system scalars [8][B]
  int *flags;                  // [0]=ndone, [1..B]=done[b], then iters[b]
  int grid, grid_mixed, grid_ic;   // resident grid sizes of pcg_kernel / pcg_mixed_kernel / pcg_ic_kernel
};

// ---- pre.cu
int k_deinterleave(b200flow_ctx *, const double *src, double *dst, int B, long long HW, int C);
int k_interleave(b200flow_ctx *, const double *src, double *dst, int B, long long HW, int C);
int k_minmax_scale(b200flow_ctx *, const double *in, double *out, int items, long long n, double lo, double hi);
int k_rof_texture(b200flow_ctx *, const double *img, double *out, int B, int C, int H, int W, double theta,
                  int iters, double alp);
int k_gauss_resize(b200flow_ctx *, const double *src, double *dst, int P, int H, int W, int Hn, int Wn,
                   const double *taps, int ks);
int k_resample_flow(b200flow_ctx *, const double2 *in, double2 *out, int B, int h, int w, int H, int W);
int k_rgb8_to_gray_lab(b200flow_ctx *, const unsigned char *rgb1, const unsigned char *rgb2, int B, long long HW,
                       double *gray /*[B][2][HW]*/, double *lab /*[B][3][HW] or null*/);
int k_rgbf_to_gray(b200flow_ctx *, const double *rgb, long long HW, double *gray);
int k_rgbf_to_lab(b200flow_ctx *, const double *rgb, long long HW, double *lab /*[3][HW]*/);
void gaussian_taps(double spacing, double *taps /*>=81*/, int *ks);
int level_size(int n, double ratio);
int auto_levels(int H, int W, double spacing);

// ---- warp.cu
// frames: [B][2*NC][H][W], NC channels of frame 1 then NC channels of frame 2 (bstride = 2*NC*H*W for a dense batch);
// I1x / I1y / src2 / It / Ix / Iy are [B][NC][H][W]
int k_level_prep(b200flow_ctx *, const double *frames, long long bstride, int B, int NC, int H, int W,
                 int interp, const double filt[5], double *I1x, double *I1y, double4 *src2,
                 double4 *tmp = nullptr /* [B][NC][H][W] scratch of the spline prefilter; allocated when null */);
int k_warp_assemble(b200flow_ctx *, const double *frames, long long bstride, int NC, const double *I1x, const double *I1y,
                    const double4 *src2,
                    const double2 *uv, const double2 *duv, int B, int H, int W, int interp, double blend,
                    const PenaltySet &ps, LinSys sys, double *It, double *Ix, double *Iy);
// assemble from given derivative planes (operator_apply / solve_increment entry points, max_linear>1 re-linearisation)
int k_assemble_from_deriv(b200flow_ctx *, const double *It, const double *Ix, const double *Iy, int NC, const double2 *uv,
                          const double2 *duv, int B, int H, int W, const PenaltySet &ps, LinSys sys);
int k_robust_eval(b200flow_ctx *, b200flow_penalty pen, int d_type, const double *x, long long n, double *y);

// ---- solve.cu
size_t pcg_work_bytes(const b200flow_ctx *, int B, int H, int W);
int pcg_work_alloc(b200flow_ctx *, int B, int H, int W, PcgWork *w);
// mode: how `solver` of b200flow_params maps onto the two persistent kernels
enum { PCG_MODE_MIXED = 0,        // block-Jacobi PCG, fp32 Krylov vectors + fp64 reliable updates (B200FLOW_SOLVER_EXACT)
       PCG_MODE_JACOBI_F64 = 1,   // scalar-Jacobi all-fp64 PCG: the reference's own 'pcg' mode (B200FLOW_SOLVER_PCG)
       PCG_MODE_BLOCK_F64 = 2,    // block-Jacobi all-fp64 PCG (B200FLOW_SOLVER_EXACT_F64)
       PCG_MODE_SOR = 3,          // the reference's legacy lexicographic SOR, omega 1.9 (B200FLOW_SOLVER_SOR)
       PCG_MODE_MIXED_IC = 4,     // as MIXED with the tile-local block-IC(0) preconditioner (B200FLOW_SOLVER_EXACT_IC)
       PCG_MODE_FP32_IC = 5 };    // the IC kernel without reliable updates, stopped on the iterated fp32 residual (B200FLOW_SOLVER_FP32_IC)
inline int pcg_mode_of(int solver) { return solver; }
// algorithmic bytes per pixel-iteration of a solver mode (roofline accounting; solve.cu / solve_ic.cu headers)
inline int pcg_bytes_per_pixel_iter(int mode) {
  return (mode == PCG_MODE_MIXED_IC || mode == PCG_MODE_FP32_IC) ? 128 : mode == PCG_MODE_MIXED ? 120 : 228;
}
int k_pcg_solve(b200flow_ctx *, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int mode,
                int *iters_host /*[B] or null*/, double *relres_host /*[B] or null*/, bool sync_results);
int k_pcg_solve_async(b200flow_ctx *, LinSys sys, PcgWork w, double2 *x, double tol, int maxit, int mode,
                      long long *stats_dev);
int k_operator_apply(b200flow_ctx *, LinSys sys, const double2 *x, double2 *Ax, double2 *diag);
int pcg_ic_grid(b200flow_ctx *ctx, int *grid_out);   // solve_ic.cu: resident grid of pcg_ic_kernel (fills ctx->ic_ctas_per_sm)

// ---- filter.cu
// out = base + (median(base + clip(x)) - base) when x != null (BA update, ba.py:186-204), else out = median(base)
int k_median_uv(b200flow_ctx *, const double2 *base, const double2 *x, int limit_update, const int *active,
                double2 *out, int B, int H, int W, int kh, int kw, int assign_direct);
int k_clip_add(b200flow_ctx *, const double2 *uv, const double2 *x, int limit_update, const int *active, double2 *out,
               long long n_per_item, int B);
int k_occlusion(b200flow_ctx *, const double2 *uv, const double *frames, long long bstride, int NC, int B,
                int H, int W, double sigma_d, double sigma_i, double *occ);
int k_sub(b200flow_ctx *, const double2 *a, const double2 *b, double2 *out, long long n);
// out = base + (wmed(cand) - base) when base != null else wmed(cand)
int k_weighted_median(b200flow_ctx *, const double2 *cand, const double2 *base, const double *color, int C,
                      const double *occ, int B, int H, int W, int hsz, double sigma_i, double2 *out);
int k_hs_norm_gate(b200flow_ctx *, const double2 *x, int B, long long n, int *active, double *scratch);
int k_delta_norm(b200flow_ctx *, const double2 *x, const double2 *sub, int clip, long long n, double *out_sq);

// ---- eval.cu
int k_flow_error(b200flow_ctx *, const double2 *uv, const double2 *gt, int B, int H, int W, int border,
                 double *result /*[B][4] = AAE, std, AEPE, valid count*/);
int k_flow_to_color(b200flow_ctx *, const double2 *uv, int B, int H, int W, double max_flow, unsigned char *rgb);
int k_flow_to_flo(b200flow_ctx *, const double2 *uv, int B, int H, int W, unsigned char *out /*[B][12 + 8HW]*/);

// ---- pipeline.cu
int run_pipeline(b200flow_ctx *, const b200flow_params *p, int B, int H, int W, int NC, int C, const double *gray_planar,
                 const double *color_planar, const double2 *init, double2 *uv_out, b200flow_stats *stats);

}  // namespace bf
