// warp.cu -- kernel groups (2)+(3): per-level derivative planes, cubic B-spline prefilter, and the fused
// warp + Ix/Iy/It + robust IRLS weights + five-point-system assembly kernel.
// Replaces derivatives.py:27-296 (partial_deriv, interp2_bicubic), penalties.py (deriv_over_x, all 10
// penalties) and flow_operator (classic_nl.py:279-378, ba.py:208-302, hs.py:144-203) of the reference.
#include "kernels.cuh"

namespace bf {

// ------------------------------------------------------------------------------------------------
// robust penalties, d_type 0/1/2 (penalties.py:18-345; App. B of SURVEY.md)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double pen_weight(const b200flow_penalty &pn, double x) {   // rho'(x)/x
  double s2 = pn.p0 * pn.p0;
  switch (pn.kind) {
    case 0: return 2.0 / s2;
    case 1: return 2.0 / (2.0 * s2 + x * x);
    case 2: { double q = x / s2; return 1.0 / (s2 * sqrt(1.0 + q * q)); }
    case 3: return 2.0 * pn.p1 * pow(s2 + x * x, pn.p1 - 1.0);
    case 4: { double d = s2 + x * x; return 2.0 * s2 / (d * d); }
    case 5: { double ax = fabs(x); return ax <= s2 ? 2.0 : 2.0 * s2 / fmax(ax, 1e-30); }
    case 6: { double om = 1.0 - (x * x) / s2; return fabs(x) <= pn.p0 ? 2.0 * (om * om) / s2 : 0.0; }
    case 7: return 1.0 / s2;
    case 8:
    case 9: return (pn.p0 + 1.0) / (pn.p1 * pn.p1 * pn.p0 + x * x);
  }
  return 0.0;
}

// Same weights for the ASSEMBLY kernels.  The generalized Charbonnier weight 2a (s2 + x^2)^(a-1) is the only penalty that
// needs a transcendental (nine of them per pixel in Classic+NL / classic++).  pow() costs ~200 fp64 instructions,
// exp((a-1) log y) ~200 too once both are inlined (62 % of warp_assemble_kernel's instructions, ncu source page, round 2),
// and B200 issues fp64 at half rate -- the kernel was bound by that pipe, not by HBM.  pow_tab() splits
//     y = 2^e m,  m = (1 + d) / inv_k     (k = the top 7 mantissa bits, inv_k = 1 / bin centre, |d| <= 2^-8)
//     y^c = 2^(c e) * (1/inv_k)^c * (1 + d)^c
// with the first two factors read from a 3 KB table built once per exponent (pow_table_kernel) and the third a degree-6
// binomial series (truncation < 1e-17): 9 fp64 instructions, relative error <= 4 ulp.  Outside 2^-64 <= y < 2^64 (and for
// exponents the table was not built for) the exp/log form remains.  The operator tests hold the assembled A, b to 1e-11
// relative; the public RobustFunction.deriv_over_x (pen_eval below) keeps pow().
constexpr int POWTAB_BINS = 128, POWTAB_EXP = 128, POWTAB_DOUBLES = 2 * POWTAB_BINS + POWTAB_EXP;

__global__ void pow_table_kernel(double *tab, double c) {
  const int k = threadIdx.x;
  if (k < POWTAB_BINS) {
    const double inv = 1.0 / (1.0 + (k + 0.5) / POWTAB_BINS);
    tab[2 * k] = inv;
    tab[2 * k + 1] = pow(1.0 / inv, c);
  }
  if (k < POWTAB_EXP) tab[2 * POWTAB_BINS + k] = exp2(c * (double)(k - POWTAB_EXP / 2));
}

__device__ __forceinline__ double pow_tab(const PenaltySet &ps, double y, double c) {
  const long long bits = __double_as_longlong(y);
  const int e = (int)(bits >> 52) - 1023 + POWTAB_EXP / 2;
  if ((unsigned)e >= (unsigned)POWTAB_EXP) return exp(c * log(y));   // also inf / nan / subnormal
  const int k = (int)(bits >> 45) & (POWTAB_BINS - 1);
  const double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
  const double2 t = __ldg(reinterpret_cast<const double2 *>(ps.ptab) + k);
  const double sc = t.y * __ldg(ps.ptab + 2 * POWTAB_BINS + e);
  const double d = fma(m, t.x, -1.0);
  double p = fma(ps.pc[5], d, ps.pc[4]);
  p = fma(p, d, ps.pc[3]);
  p = fma(p, d, ps.pc[2]);
  p = fma(p, d, ps.pc[1]);
  p = fma(p, d, ps.pc[0]);
  return fma(p * d, sc, sc);
}

__device__ __forceinline__ double pen_weight_asm(const PenaltySet &ps, const b200flow_penalty &pn, double x) {
  if (pn.kind == 3) {
    const double y = pn.p0 * pn.p0 + x * x;
    if (ps.ptab && pn.p1 == ps.ptab_a) return 2.0 * pn.p1 * pow_tab(ps, y, pn.p1 - 1.0);
    return 2.0 * pn.p1 * exp((pn.p1 - 1.0) * log(y));
  }
  return pen_weight(pn, x);
}

// Host side of pow_tab(): pick the exponent the table serves (the first generalized Charbonnier penalty of the set; the
// reference's methods use one `a` throughout), rebuild the table on the context's stream when it changed.
static int attach_pow_table(b200flow_ctx *ctx, PenaltySet *ps) {
  ps->ptab = nullptr;
  ps->fast = 0;
  if (ps->hs || !(ps->alpha <= 1.0)) return 0;
  const b200flow_penalty *cand[5] = {&ps->rho_d, &ps->rho_su[0], &ps->rho_su[1], &ps->rho_sv[0], &ps->rho_sv[1]};
  const b200flow_penalty *pn = nullptr;
  for (auto q : cand)
    if (q->kind == 3) { pn = q; break; }
  if (!pn || getenv("B200FLOW_POW_EXPLOG")) return 0;
  if (ps->alpha < 1.0) {            // alpha == 1 (first GNC stage): the robust penalties carry zero weight, no table is read
    if (!ctx->powtab) BF_CUDA(ctx, cudaMalloc(&ctx->powtab, POWTAB_DOUBLES * sizeof(double)));
    if (ctx->powtab_a != pn->p1) {
      BF_LAUNCH(ctx, pow_table_kernel, 1, POWTAB_BINS, 0, ctx->powtab, pn->p1 - 1.0);
      ctx->powtab_a = pn->p1;
    }
    ps->ptab = ctx->powtab;
  }
  ps->ptab_a = pn->p1;
  double c = pn->p1 - 1.0, b = 1.0;
  for (int k = 1; k <= 6; ++k) { b = b * (c - (k - 1)) / k; ps->pc[k - 1] = b; }
  // the Classic+NL shape of the set: one generalized Charbonnier for the four spatial terms and one for the data term (same
  // exponent), quadratic partners -> the two-constant form of blended_edge<true> / blended_data<true>
  auto same = [](const b200flow_penalty &a, const b200flow_penalty &b) { return a.kind == b.kind && a.p0 == b.p0 && a.p1 == b.p1; };
  const b200flow_penalty &rs = ps->rho_su[0], &qs = ps->qua_su[0];
  bool fast = rs.kind == 3 && rs.p1 == pn->p1 && ps->rho_d.kind == 3 && ps->rho_d.p1 == pn->p1 && qs.kind == 0 && ps->qua_d.kind == 0;
  for (int k = 0; k < 2; ++k)
    fast = fast && same(ps->rho_su[k], rs) && same(ps->rho_sv[k], rs) && same(ps->qua_su[k], qs) && same(ps->qua_sv[k], qs);
  ps->fast = fast && !getenv("B200FLOW_ASM_GENERIC");
  if (ps->fast) {
    const double a = pn->p1, al = ps->alpha;
    ps->eq = al > 0.0 ? al * (ps->lambda_q * (2.0 / (qs.p0 * qs.p0))) : 0.0;
    ps->er = (1.0 - al) * (ps->lambda * (2.0 * a));
    ps->s2s = rs.p0 * rs.p0;
    ps->dq = al > 0.0 ? al * (2.0 / (ps->qua_d.p0 * ps->qua_d.p0)) : 0.0;
    ps->dr = (1.0 - al) * (2.0 * a);
    ps->s2d = ps->rho_d.p0 * ps->rho_d.p0;
  }
  return 0;
}

__device__ double pen_eval(const b200flow_penalty &pn, int d_type, double x, double tdist_const) {
  if (d_type == 2) return pen_weight(pn, x);
  double s2 = pn.p0 * pn.p0;
  switch (pn.kind) {
    case 0: return d_type == 0 ? x * x / s2 : 2.0 * x / s2;
    case 1: return d_type == 0 ? log(1.0 + x * x / (2.0 * s2)) : 2.0 * x / (2.0 * s2 + x * x);
    case 2: { double q = x / s2; double r = sqrt(1.0 + q * q); return d_type == 0 ? s2 * r : x / (s2 * r); }
    case 3: { double base = s2 + x * x;
              return d_type == 0 ? pow(base, pn.p1) : 2.0 * pn.p1 * x * pow(base, pn.p1 - 1.0); }
    case 4: { double d = s2 + x * x; return d_type == 0 ? x * x / d : 2.0 * s2 * x / (d * d); }
    case 5: { double ax = fabs(x); bool in = ax <= s2;
              if (d_type == 0) return in ? x * x : 2.0 * s2 * ax - s2 * s2;
              return in ? 2.0 * x : 2.0 * s2 * (x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0)); }
    case 6: { double om = 1.0 - (x * x) / s2; bool in = fabs(x) <= pn.p0;
              if (d_type == 0) return in ? (1.0 / 3.0) * (1.0 - om * om * om) : 1.0 / 3.0;
              return in ? 2.0 * x * (om * om) / s2 : 0.0; }
    case 7: { double q = x / pn.p0;
              return d_type == 0 ? 0.5 * log(2.0 * 3.141592653589793) + log(pn.p0) + 0.5 * (q * q) : x / s2; }
    case 8:
    case 9: { double s2r = pn.p1 * pn.p1 * pn.p0;
              if (d_type == 0) return (pn.p0 + 1.0) / 2.0 * log(1.0 + x * x / s2r) + tdist_const;
              return (pn.p0 + 1.0) * x / (s2r + x * x); }
  }
  return 0.0;
}

__global__ void robust_eval_kernel(b200flow_penalty pn, int d_type, const double *__restrict__ x, long long n,
                                   double *__restrict__ y, double tdist_const) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = pen_eval(pn, d_type, x[i], tdist_const);
}

int k_robust_eval(b200flow_ctx *ctx, b200flow_penalty pen, int d_type, const double *x, long long n, double *y) {
  if (pen.kind < 0 || pen.kind > 9) return set_err(ctx, B200FLOW_EINVAL, "Unknown penalty kind %d", pen.kind);
  if (d_type < 0 || d_type > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown d_type: %d", d_type);
  double c = 0.0;
  if (pen.kind == 8)   // tdist normalisation constant is a host-side scalar (penalties.py:304-306)
    c = std::lgamma(pen.p0 / 2.0) - std::lgamma((pen.p0 + 1.0) / 2.0) + 0.5 * std::log(pen.p0 * 3.141592653589793) +
        std::log(pen.p1);
  if (n > 0) BF_LAUNCH(ctx, robust_eval_kernel, (unsigned)cdiv(n, 256), 256, 0, pen, d_type, x, n, y, c);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// per-level derivative planes (once per level, not per warp)
//   I1x, I1y             5-tap correlate of frame 1, scipy 'reflect'          (derivatives.py:201-202,258-259)
//   Hermite source       {Z, DX, DY, DXY} of frame 2, DXY = 5x5 outer(h,h)    (derivatives.py:77-86)
//   spline source        {I2, I2x, I2y, 0}, then prefiltered in place          (derivatives.py:244-254)
// algorithmic bytes/pixel: read 16, write 16 + 32 = 64 B
// ------------------------------------------------------------------------------------------------
struct Filt5 { double h[5]; };

// frames: [B][2*NC][H][W] (frame-1 channels, then frame-2 channels; pair stride bstride); outputs [B][NC][H][W]
__global__ void level_prep_kernel(const double *__restrict__ frames, long long bstride, int NC,
                                  int H, int W, int hermite, Filt5 f, double *__restrict__ I1x,
                                  double *__restrict__ I1y, double4 *__restrict__ src2) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const long long HW = (long long)H * W;
  long long off = (long long)blockIdx.z * HW;
  const int pb = blockIdx.z / NC, pc = blockIdx.z - pb * NC;
  const double *im1 = frames + (long long)pb * bstride + (long long)pc * HW;
  const double *im2 = im1 + (long long)NC * HW;
  int xs[5], ys[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { xs[k] = reflect_idx(x + k - 2, W); ys[k] = reflect_idx(y + k - 2, H); }
  double a1x = 0.0, a1y = 0.0, a2x = 0.0, a2y = 0.0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    a1x += f.h[k] * im1[(long long)y * W + xs[k]];
    a1y += f.h[k] * im1[(long long)ys[k] * W + x];
    a2x += f.h[k] * im2[(long long)y * W + xs[k]];
    a2y += f.h[k] * im2[(long long)ys[k] * W + x];
  }
  long long i = off + (long long)y * W + x;
  I1x[i] = a1x;
  I1y[i] = a1y;
  double dxy = 0.0;
  if (hermite) {
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        double k = f.h[a] * f.h[b];
        if (k != 0.0) dxy += k * im2[(long long)ys[a] * W + xs[b]];
      }
  }
  src2[i] = make_double4(im2[(long long)y * W + x], a2x, a2y, dxy);
}

// cubic B-spline prefilter, pole z = sqrt(3)-2, gain 6, whole-sample mirror boundary, on the three live components of
// the double4 (scipy ni_splines.c apply_filter): per line  c+_i = gain c_i + z c+_{i-1}  (causal, exact
// mirror initialisation) then  c_i = z (c_{i+1} - c+_i)  (anticausal, exact end initialisation).
// The recursions are sequential along a line, but |z| = 0.268 forgets its past at 2^-1.9 per sample: a line is cut into
// SEGMENTS, and a segment that does not start at the line's end runs BSP_WARM = 64 samples of warm-up from a zero state
// first (z^64 = 2e-37: what the truncation leaves is 20 orders below the last bit, so the result has the bits of the
// whole-line recursion).  One thread per (line, segment), lines fastest across a warp: the column pass is coalesced and
// the row pass touches exactly one 32-byte sector per thread and step; a 1080-row plane gives 1080 x 15 threads instead
// of 1080.  Out of place (src -> dst) because warm-ups read samples other segments are overwriting.
constexpr int BSP_SEG = 128, BSP_WARM = 64;
__device__ __forceinline__ void d4_scale(double4 &v, double s) { v.x *= s; v.y *= s; v.z *= s; }

struct BspLine { long long base; int n, stride; };
__device__ __forceinline__ bool bsp_line(int item, int H, int W, int along_x, int nlines, BspLine &L, int &seg) {
  const int line = item % nlines;
  seg = item / nlines;
  if (along_x) { L.n = W; L.stride = 1; L.base = (long long)line * W; }                                   // line = plane * H + y
  else { L.n = H; L.stride = W; L.base = (long long)(line / W) * H * W + (line % W); }                    // line = plane * W + x
  return true;
}

__global__ void bspline_causal_kernel(const double4 *__restrict__ src, double4 *__restrict__ dst, int H, int W, int along_x,
                                      int nlines, int nseg) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= nlines * nseg) return;
  BspLine L; int seg;
  bsp_line(item, H, W, along_x, nlines, L, seg);
  const int n = L.n;
  const long long st = L.stride;
  const double4 *p = src + L.base;
  double4 *q = dst + L.base;
  if (n < 2) { if (seg == 0 && n == 1) q[0] = p[0]; return; }
  const double z = -0.2679491924311227064725536584941276330571947461896193719441930205;   // sqrt(3)-2
  const double gain = (1.0 - z) * (1.0 - 1.0 / z);
  const int a = seg * BSP_SEG, b = (seg + 1 == nseg) ? n : (seg + 1) * BSP_SEG;
  double4 prev;
  int i0;
  if (a == 0) {
    // causal initialisation  c0 = [c0 + z^(n-1) c_{n-1} + sum_{i=1}^{n-2} z^i (c_i + z^(n-1) c_{n-1-i})] / (1 - z^(2n-2))
    double zn1 = pow(z, (double)(n - 1));
    double4 first = p[0], last = p[(long long)(n - 1) * st];
    d4_scale(first, gain);
    d4_scale(last, gain);
    double sx = first.x + zn1 * last.x, sy = first.y + zn1 * last.y, sz = first.z + zn1 * last.z;
    double zi = z;
    for (int i = 1; i < n - 1; ++i) {
      double4 u = p[(long long)i * st], v = p[(long long)(n - 1 - i) * st];
      sx += zi * (u.x * gain + zn1 * (v.x * gain));
      sy += zi * (u.y * gain + zn1 * (v.y * gain));
      sz += zi * (u.z * gain + zn1 * (v.z * gain));
      zi *= z;
      if (fabs(zi) < 1e-40 && fabs(zn1) < 1e-40) break;   // what is left is 20 orders below the last bit of the sum
    }
    double den = 1.0 - zn1 * zn1;
    prev = make_double4(sx / den, sy / den, sz / den, p[0].w);
    q[0] = prev;
    i0 = 1;
  } else {
    prev = make_double4(0.0, 0.0, 0.0, 0.0);
    for (int i = a - BSP_WARM; i < a; ++i) {
      const double4 v = p[(long long)i * st];
      prev.x = v.x * gain + z * prev.x;
      prev.y = v.y * gain + z * prev.y;
      prev.z = v.z * gain + z * prev.z;
    }
    i0 = a;
  }
  for (int i = i0; i < b; ++i) {
    double4 v = p[(long long)i * st];
    v.x = v.x * gain + z * prev.x;
    v.y = v.y * gain + z * prev.y;
    v.z = v.z * gain + z * prev.z;
    q[(long long)i * st] = v;
    prev = v;
  }
}

__global__ void bspline_anticausal_kernel(const double4 *__restrict__ src, double4 *__restrict__ dst, int H, int W,
                                          int along_x, int nlines, int nseg) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= nlines * nseg) return;
  BspLine L; int seg;
  bsp_line(item, H, W, along_x, nlines, L, seg);
  const int n = L.n;
  const long long st = L.stride;
  const double4 *p = src + L.base;            // the causal pass' output c+
  double4 *q = dst + L.base;
  if (n < 2) { if (seg == 0 && n == 1) q[0] = p[0]; return; }
  const double z = -0.2679491924311227064725536584941276330571947461896193719441930205;
  const int a = seg * BSP_SEG, b = (seg + 1 == nseg) ? n : (seg + 1) * BSP_SEG;
  double4 prev;
  int i1;                                     // first index (going down) that is stored
  if (b + BSP_WARM >= n) {
    // anticausal initialisation at the line's end, then down to this segment (at most BSP_WARM + BSP_SEG unsaved steps)
    const double4 pm = p[(long long)(n - 2) * st], pl = p[(long long)(n - 1) * st];
    const double k = z / (z * z - 1.0);
    prev = make_double4((z * pm.x + pl.x) * k, (z * pm.y + pl.y) * k, (z * pm.z + pl.z) * k, pl.w);
    if (b == n) q[(long long)(n - 1) * st] = prev;
    for (int i = n - 2; i >= b; --i) {
      const double4 v = p[(long long)i * st];
      prev.x = z * (prev.x - v.x); prev.y = z * (prev.y - v.y); prev.z = z * (prev.z - v.z);
    }
    i1 = b == n ? n - 2 : b - 1;
  } else {
    prev = make_double4(0.0, 0.0, 0.0, 0.0);
    for (int i = b + BSP_WARM - 1; i >= b; --i) {
      const double4 v = p[(long long)i * st];
      prev.x = z * (prev.x - v.x); prev.y = z * (prev.y - v.y); prev.z = z * (prev.z - v.z);
    }
    i1 = b - 1;
  }
  for (int i = i1; i >= a; --i) {
    double4 v = p[(long long)i * st];
    v.x = z * (prev.x - v.x);
    v.y = z * (prev.y - v.y);
    v.z = z * (prev.z - v.z);
    q[(long long)i * st] = v;
    prev = v;
  }
}

static int bspline_prefilter_axis(b200flow_ctx *ctx, double4 *c, double4 *tmp, int planes, int H, int W, int along_x) {
  const int n = along_x ? W : H, nlines = planes * (along_x ? H : W);
  int nseg = n / BSP_SEG;                       // the last segment takes the remainder; short lines are one exact segment
  if (nseg < 1) nseg = 1;
  const long long items = (long long)nlines * nseg;
  BF_LAUNCH(ctx, bspline_causal_kernel, (unsigned)cdiv(items, 128), 128, 0, c, tmp, H, W, along_x, nlines, nseg);
  BF_LAUNCH(ctx, bspline_anticausal_kernel, (unsigned)cdiv(items, 128), 128, 0, tmp, c, H, W, along_x, nlines, nseg);
  return 0;
}

int k_level_prep(b200flow_ctx *ctx, const double *frames, long long bstride, int B, int NC, int H, int W,
                 int interp, const double filt[5], double *I1x, double *I1y, double4 *src2, double4 *tmp) {
  Filt5 f;
  for (int i = 0; i < 5; ++i) f.h[i] = filt[i];
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B * NC);
  BF_LAUNCH(ctx, level_prep_kernel, grd, blk, 0, frames, bstride, NC, H, W,
            interp == B200FLOW_INTERP_BICUBIC ? 1 : 0, f, I1x, I1y, src2);
  if (interp == B200FLOW_INTERP_CUBIC) {
    // scipy spline_filter: axis 0 (columns) first, then axis 1 (rows); tmp: the out-of-place partner of src2
    double4 *t = tmp;
    if (!t) BF_TRY(arena_alloc(ctx, &t, (size_t)B * NC * H * W));
    BF_TRY(bspline_prefilter_axis(ctx, src2, t, B * NC, H, W, 0));
    BF_TRY(bspline_prefilter_axis(ctx, src2, t, B * NC, H, W, 1));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// warp + derivatives (partial_deriv) for one pixel
// ------------------------------------------------------------------------------------------------
struct Deriv { double It, Ix, Iy; };

// 32-byte gather element through the read-only path as two 128-bit loads (one sector)
__device__ __forceinline__ double4 ld4(const double4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}

// L1 prefetch (no destination register): warp_assemble is latency-bound at three CTAs per SM -- the operand streams that do
// not depend on the flow are requested before the flow arrives, the four Hermite corners as soon as it has
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void hermite_basis(double t, double &h0, double &h1, double &g0, double &g1, double &dh0,
                                              double &dh1, double &dg0, double &dg1) {
  // factored forms (14 instead of 27 fp64 instructions per axis; the library is built with -fmad=false, these are explicit):
  // h0 = 1 - t^2 (3 - 2t), h1 = 1 - h0, g0 = t (t-1)^2, g1 = t^2 (t-1), dh0 = 6 t (t-1) = -dh1, dg0 = 1 + t (3t - 4), dg1 = t (3t - 2)
  const double t2 = t * t, u = t - 1.0;
  h0 = fma(t2, fma(2.0, t, -3.0), 1.0);
  h1 = 1.0 - h0;
  g0 = (t * u) * u;
  g1 = t2 * u;
  dh0 = 6.0 * (t * u);
  dh1 = -dh0;
  dg0 = fma(t, fma(3.0, t, -4.0), 1.0);
  dg1 = t * fma(3.0, t, -2.0);
}

// tensor-product cubic Hermite on the unit cell == the 16x16 bcucof/bcuint product of interp2_bicubic
// (derivatives.py:27-145; closed form of SURVEY App. A.4b).  x1,y1 are 1-based sample coordinates.
__device__ __forceinline__ bool warp_hermite(const double4 *__restrict__ s, int H, int W, double x1, double y1,
                                             double &val, double &ddx, double &ddy) {
  double flx = floor(x1), fly = floor(y1);
  // compare in double so that huge / non-finite flows are classified out of bounds instead of overflowing int
  bool oob = !(flx >= 1.0) || !(flx + 1.0 <= (double)W) || !(fly >= 1.0) || !(fly + 1.0 <= (double)H);
  if (oob) { val = ddx = ddy = 0.0; return false; }
  int fx = (int)flx, fy = (int)fly;
  int x0 = fx - 1, xb = fx, y0 = fy - 1, yb = fy;   // 0-based corners (already inside the image)
  double ax = x1 - flx, ay = y1 - fly;
  double hx[2], gx[2], dhx[2], dgx[2], hy[2], gy[2], dhy[2], dgy[2];
  hermite_basis(ax, hx[0], hx[1], gx[0], gx[1], dhx[0], dhx[1], dgx[0], dgx[1]);
  hermite_basis(ay, hy[0], hy[1], gy[0], gy[1], dhy[0], dhy[1], dgy[0], dgy[1]);
  // sum over the four corners of (hx, gx) (x) (hy, gy) against {Z, DX, DY, DXY}, factored: the x-direction products
  // P, Q (value basis) and dP, dQ (derivative basis) are shared by val / ddx / ddy -- 14 fused multiply-adds per corner
  // instead of 36 multiplies + 12 adds (the library is built with -fmad=false; these are explicit: the result is a
  // smooth function of the inputs, no discrete decision depends on its last bit)
  val = ddx = ddy = 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const double4 c = ld4(&s[(long long)(b ? yb : y0) * W + (a ? xb : x0)]);   // {Z, DX, DY, DXY}
      const double P = fma(c.y, gx[a], c.x * hx[a]), Q = fma(c.w, gx[a], c.z * hx[a]);
      const double dP = fma(c.y, dgx[a], c.x * dhx[a]), dQ = fma(c.w, dgx[a], c.z * dhx[a]);
      val = fma(P, hy[b], fma(Q, gy[b], val));
      ddx = fma(dP, hy[b], fma(dQ, gy[b], ddx));
      ddy = fma(P, dhy[b], fma(Q, dgy[b], ddy));
    }
  return true;
}

// scipy map_coordinates(order=3 | 1, mode='constant', cval=nan) on {I2, I2x, I2y}; x0,y0 0-based
__device__ __forceinline__ void warp_spline(const double4 *__restrict__ s, int H, int W, double x0, double y0,
                                            int order, double &val, double &vx, double &vy) {
  double flx = floor(x0), fly = floor(y0);
  int fx = (int)flx, fy = (int)fly;
  double tx = x0 - flx, ty = y0 - fly;
  val = vx = vy = 0.0;
  if (order == 3) {
    double wx[4], wy[4];
    {
      double z = 1.0 - tx;
      wx[1] = (tx * tx * (tx - 2.0) * 3.0 + 4.0) / 6.0;
      wx[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
      wx[0] = z * z * z / 6.0;
      wx[3] = 1.0 - wx[0] - wx[1] - wx[2];
      z = 1.0 - ty;
      wy[1] = (ty * ty * (ty - 2.0) * 3.0 + 4.0) / 6.0;
      wy[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
      wy[0] = z * z * z / 6.0;
      wy[3] = 1.0 - wy[0] - wy[1] - wy[2];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double4 *row = s + (long long)mirror_idx(fy + a - 1, H) * W;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double4 c = ld4(&row[mirror_idx(fx + b - 1, W)]);
        double w = wy[a] * wx[b];
        val += c.x * w;
        vx += c.y * w;
        vy += c.z * w;
      }
    }
  } else {
    int x1 = mirror_idx(fx + 1, W), y1 = mirror_idx(fy + 1, H);
    double4 c00 = ld4(&s[(long long)fy * W + fx]), c01 = ld4(&s[(long long)fy * W + x1]);
    double4 c10 = ld4(&s[(long long)y1 * W + fx]), c11 = ld4(&s[(long long)y1 * W + x1]);
    val = (1.0 - ty) * ((1.0 - tx) * c00.x + tx * c01.x) + ty * ((1.0 - tx) * c10.x + tx * c11.x);
    vx = (1.0 - ty) * ((1.0 - tx) * c00.y + tx * c01.y) + ty * ((1.0 - tx) * c10.y + tx * c11.y);
    vy = (1.0 - ty) * ((1.0 - tx) * c00.z + tx * c01.z) + ty * ((1.0 - tx) * c10.z + tx * c11.z);
  }
}

__device__ __forceinline__ Deriv pixel_deriv(const double *__restrict__ im1, const double *__restrict__ I1x,
                                             const double *__restrict__ I1y, const double4 *__restrict__ src2,
                                             int H, int W, int x, int y, double2 f, int interp, double blend) {
  long long i = (long long)y * W + x;
  double x2 = (double)(x + 1) + f.x, y2 = (double)(y + 1) + f.y;   // 1-based, as the reference's meshgrid
  double val, wx, wy;
  bool ok;
  if (interp == B200FLOW_INTERP_BICUBIC) {
    ok = warp_hermite(src2, H, W, x2, y2, val, wx, wy);
  } else {
    ok = !(x2 > (double)W) && !(x2 < 1.0) && !(y2 > (double)H) && !(y2 < 1.0) && x2 == x2 && y2 == y2;
    if (ok) warp_spline(src2, H, W, x2 - 1.0, y2 - 1.0, interp == B200FLOW_INTERP_CUBIC ? 3 : 1, val, wx, wy);
  }
  Deriv d;
  if (!ok) { d.It = d.Ix = d.Iy = 0.0; return d; }
  d.It = val - im1[i];
  d.Ix = blend * wx + (1.0 - blend) * I1x[i];
  d.Iy = blend * wy + (1.0 - blend) * I1y[i];
  return d;
}

// ------------------------------------------------------------------------------------------------
// IRLS weights + system assembly for one pixel (flow_operator + GNC blend)
// ------------------------------------------------------------------------------------------------
// FAST: the penalty set is the Classic+NL family's -- one generalized Charbonnier penalty for all four spatial terms, one (same
// exponent) for the data term, quadratic GNC partners (attach_pow_table checks it) -- so the per-call dispatch on the penalty kind
// and its constant loads (two thirds of the kernel's instructions once pow became a table) fold into two host-side constants per
// term: w = eq + er * y^(a-1).
template <bool FAST>
__device__ __forceinline__ double blended_edge(const PenaltySet &ps, const b200flow_penalty &rob,
                                               const b200flow_penalty &qua, double delta) {
  if (FAST) {
    if (ps.er == 0.0) return ps.eq;
    return fma(ps.er, pow_tab(ps, fma(delta, delta, ps.s2s), ps.ptab_a - 1.0), ps.eq);
  }
  if (ps.hs) return ps.hs_w;
  double w = 0.0;
  if (ps.alpha > 0.0) w = w + ps.alpha * (ps.lambda_q * pen_weight_asm(ps, qua, delta));
  if (ps.alpha < 1.0) w = w + (1.0 - ps.alpha) * (ps.lambda * pen_weight_asm(ps, rob, delta));
  return w;
}

template <bool FAST>
__device__ __forceinline__ double blended_data(const PenaltySet &ps, double it_lin) {
  if (FAST) {
    if (ps.dr == 0.0) return ps.dq;
    return fma(ps.dr, pow_tab(ps, fma(it_lin, it_lin, ps.s2d), ps.ptab_a - 1.0), ps.dq);
  }
  if (ps.hs) return ps.hs_d;
  double d = 0.0;
  if (ps.alpha > 0.0) d = d + ps.alpha * pen_weight_asm(ps, ps.qua_d, it_lin);
  if (ps.alpha < 1.0) d = d + (1.0 - ps.alpha) * pen_weight_asm(ps, ps.rho_d, it_lin);
  return d;
}

// data-term part of one pixel's 2x2 block and right-hand side
struct DataTerm { double a11, a22, a12, bu, bv; };

// single-channel frames (flow_operator's `else` branch, classic_nl.py:344-351)
template <bool FAST>
__device__ __forceinline__ DataTerm data_term_single(const PenaltySet &ps, Deriv dv, double2 dc) {
  double it_lin = dv.It + dv.Ix * dc.x + dv.Iy * dc.y;
  double d = blended_data<FAST>(ps, it_lin);
  DataTerm t;
  t.a11 = d * dv.Ix * dv.Ix; t.a22 = d * dv.Iy * dv.Iy; t.a12 = d * dv.Ix * dv.Iy;
  t.bu = d * it_lin * dv.Ix; t.bv = d * it_lin * dv.Iy;
  return t;
}

// multi-channel frames: the IRLS weight and every product are AVERAGED over the channels separately and only then
// multiplied (classic_nl.py:330-343, ba.py:254-267, hs.py:176-181)
struct DataAccum {
  double sd = 0.0, ix2 = 0.0, iy2 = 0.0, ixy = 0.0, itx = 0.0, ity = 0.0;
  __device__ __forceinline__ void add(const PenaltySet &ps, Deriv dv, double2 dc) {
    double it_lin = dv.It + dv.Ix * dc.x + dv.Iy * dc.y;
    sd += blended_data<false>(ps, it_lin);
    ix2 += dv.Ix * dv.Ix; iy2 += dv.Iy * dv.Iy; ixy += dv.Ix * dv.Iy;
    itx += it_lin * dv.Ix; ity += it_lin * dv.Iy;
  }
  __device__ __forceinline__ DataTerm finish(int NC) const {
    double n = (double)NC, d = sd / n;
    DataTerm t;
    t.a11 = d * (ix2 / n); t.a22 = d * (iy2 / n); t.a12 = d * (ixy / n);
    t.bu = d * (itx / n); t.bv = d * (ity / n);
    return t;
  }
};

// One pixel of the assembled system.  Every edge weight is computed ONCE, by the pixel that stores it (its right and
// down edges), and handed to the neighbour on the other side through shared memory (sH / sV, the CTA's 32 x 8 tile);
// only the tile's first column / row recompute their left / up edge (the weight is an even function of the flow
// difference, so both sides get the same bits).  8 -> 4.3 penalty evaluations per pixel.  Must be called by every thread
// of the CTA (`in` = the thread has a pixel); SHARE = false is the plain per-pixel form for callers without a tile.
template <bool SHARE, bool FAST>
__device__ __forceinline__ void assemble_pixel(const PenaltySet &ps, const double2 *__restrict__ uv,
                                               const double2 *__restrict__ duv, int H, int W, int x, int y, bool in,
                                               DataTerm dt, const LinSys &sys, long long gi, double2 (*sH)[32],
                                               double2 (*sV)[32], double2 *sL) {
  const long long i = (long long)y * W + x;
  double2 c0 = make_double2(0.0, 0.0), c = c0;
  double whu = 0.0, whv = 0.0, wvu = 0.0, wvv = 0.0;        // own right / down edges (stored)
  double lr_u = 0.0, lr_v = 0.0, ld_u = 0.0, ld_v = 0.0;    // w (uv[p] - uv[q]) of the right / down edge
  auto nb = [&](long long j, double2 &n0, double2 &n) {
    n0 = uv[j];
    double2 nd = duv ? duv[j] : make_double2(0.0, 0.0);
    n = make_double2(n0.x + nd.x, n0.y + nd.y);
  };
  double2 n0, n;
  if (in) {
    c0 = uv[i];                                             // uv (for the rhs Laplacian)
    const double2 dc = duv ? duv[i] : make_double2(0.0, 0.0);
    c = make_double2(c0.x + dc.x, c0.y + dc.y);             // uv + duv (for the weights)
    if (x + 1 < W) {   // right: delta = f[x+1] - f[x]
      nb(i + 1, n0, n);
      whu = blended_edge<FAST>(ps, ps.rho_su[0], ps.qua_su[0], n.x - c.x);
      whv = blended_edge<FAST>(ps, ps.rho_sv[0], ps.qua_sv[0], n.y - c.y);
      lr_u = whu * (c0.x - n0.x);
      lr_v = whv * (c0.y - n0.y);
    }
    if (y + 1 < H) {   // down
      nb(i + W, n0, n);
      wvu = blended_edge<FAST>(ps, ps.rho_su[1], ps.qua_su[1], n.x - c.x);
      wvv = blended_edge<FAST>(ps, ps.rho_sv[1], ps.qua_sv[1], n.y - c.y);
      ld_u = wvu * (c0.x - n0.x);
      ld_v = wvv * (c0.y - n0.y);
    }
  }
  if (SHARE) {
    sH[threadIdx.y][threadIdx.x] = make_double2(whu, whv);
    sV[threadIdx.y][threadIdx.x] = make_double2(wvu, wvv);
    // the left edges of the tile's first column belong to the CTA next door: eight lanes of ONE warp recompute them (lane 0 of
    // every warp doing its own would run the evaluation eight times per CTA at 1/32 utilisation)
    if (threadIdx.y == 1 && threadIdx.x < 8) {
      const int xl = x - (int)threadIdx.x, yl = y - 1 + (int)threadIdx.x;     // pixel (tile x0, tile y0 + lane)
      if (xl > 0 && xl < W && yl < H) {
        const long long il = (long long)yl * W + xl;
        double2 a0, a, b0, b;
        nb(il, a0, a);
        nb(il - 1, b0, b);
        sL[threadIdx.x] = make_double2(blended_edge<FAST>(ps, ps.rho_su[0], ps.qua_su[0], a.x - b.x),
                                       blended_edge<FAST>(ps, ps.rho_sv[0], ps.qua_sv[0], a.y - b.y));
      }
    }
    __syncthreads();
  }
  if (!in) return;
  double lu = 0.0, lv = 0.0;                                // sum_q w_pq (uv[p] - uv[q]) in the order right, left, down, up
  lu += lr_u; lv += lr_v;
  if (x > 0) {       // left edge belongs to pixel x-1: delta = f[x] - f[x-1]
    nb(i - 1, n0, n);
    double wu, wv;
    if (SHARE) { const double2 e = threadIdx.x > 0 ? sH[threadIdx.y][threadIdx.x - 1] : sL[threadIdx.y]; wu = e.x; wv = e.y; }
    else {
      wu = blended_edge<FAST>(ps, ps.rho_su[0], ps.qua_su[0], c.x - n.x);
      wv = blended_edge<FAST>(ps, ps.rho_sv[0], ps.qua_sv[0], c.y - n.y);
    }
    lu += wu * (c0.x - n0.x);
    lv += wv * (c0.y - n0.y);
  }
  lu += ld_u; lv += ld_v;
  if (y > 0) {       // up
    nb(i - W, n0, n);
    double wu, wv;
    if (SHARE && threadIdx.y > 0) { const double2 e = sV[threadIdx.y - 1][threadIdx.x]; wu = e.x; wv = e.y; }
    else {
      wu = blended_edge<FAST>(ps, ps.rho_su[1], ps.qua_su[1], c.x - n.x);
      wv = blended_edge<FAST>(ps, ps.rho_sv[1], ps.qua_sv[1], c.y - n.y);
    }
    lu += wu * (c0.x - n0.x);
    lv += wv * (c0.y - n0.y);
  }
  sys.D[gi] = make_double2(dt.a11, dt.a22);
  sys.a12[gi] = dt.a12;
  sys.WH[gi] = make_double2(whu, whv);
  sys.WV[gi] = make_double2(wvu, wvv);
  sys.rhs[gi] = make_double2(-lu - dt.bu, -lv - dt.bv);
}

// algorithmic bytes per pixel, single-channel frames: read uv 16 + im1,I1x,I1y 24 + gathered source 32,
// write D 16 + a12 8 + WH 16 + WV 16 + rhs 16  = 144 B   (NC channels: 88 + 56 NC)
template <bool MULTI, bool FAST>   // MULTI: NC > 1 (kept out of the single-channel instantiation: it doubles the register count)
__global__ void __launch_bounds__(256, MULTI ? 1 : (FAST ? 4 : 3)) warp_assemble_kernel(const double *__restrict__ frames, long long bstride, int NC,
                                     const double *__restrict__ I1x,
                                     const double *__restrict__ I1y, const double4 *__restrict__ src2,
                                     const double2 *__restrict__ uv, const double2 *__restrict__ duv, int H, int W,
                                     int interp, double blend, PenaltySet ps, LinSys sys, double *__restrict__ It,
                                     double *__restrict__ Ix, double *__restrict__ Iy, int do_assemble) {
  __shared__ double2 sH[8][32], sV[8][32], sL[8];          // the tile's right / down edge weights (assemble_pixel)
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = x < W && y < H;
  if (!in && !do_assemble) return;
  const long long HW = (long long)H * W;
  long long off = (long long)blockIdx.z * HW;
  long long i = (long long)y * W + x;
  DataTerm dt = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (in) {
    if (!MULTI) {
      prefetch_l1(frames + (long long)blockIdx.z * bstride + i);
      prefetch_l1(I1x + off + i);
      prefetch_l1(I1y + off + i);
      if (do_assemble && y + 1 < H) prefetch_l1(uv + off + i + W);
    }
    const double2 f = uv[off + i];
    const double2 dc = duv ? duv[off + i] : make_double2(0.0, 0.0);
    if (!MULTI && interp == B200FLOW_INTERP_BICUBIC) {
      const double fx = floor((double)(x + 1) + f.x), fy = floor((double)(y + 1) + f.y);
      if (fx >= 1.0 && fx + 1.0 <= (double)W && fy >= 1.0 && fy + 1.0 <= (double)H) {
        const double4 *t = src2 + off + (long long)((int)fy - 1) * W + ((int)fx - 1);
        prefetch_l1(t); prefetch_l1(t + 1); prefetch_l1(t + W); prefetch_l1(t + W + 1);
      }
    }
    if (!MULTI) {
      Deriv dv = pixel_deriv(frames + (long long)blockIdx.z * bstride, I1x + off, I1y + off, src2 + off, H, W, x, y, f,
                             interp, blend);
      if (It) { It[off + i] = dv.It; Ix[off + i] = dv.Ix; Iy[off + i] = dv.Iy; }
      dt = data_term_single<FAST>(ps, dv, dc);
    } else {
      DataAccum acc;
      for (int c = 0; c < NC; ++c) {
        const long long coff = ((long long)blockIdx.z * NC + c) * HW;
        Deriv dv = pixel_deriv(frames + (long long)blockIdx.z * bstride + (long long)c * HW, I1x + coff, I1y + coff,
                               src2 + coff, H, W, x, y, f, interp, blend);
        if (It) { It[coff + i] = dv.It; Ix[coff + i] = dv.Ix; Iy[coff + i] = dv.Iy; }
        acc.add(ps, dv, dc);
      }
      dt = acc.finish(NC);
    }
  }
  if (do_assemble)
    assemble_pixel<true, FAST>(ps, uv + off, duv ? duv + off : nullptr, H, W, x, y, in, dt, sys, off + i, sH, sV, sL);
}

// It, Ix, Iy: [B][NC][H][W]
template <bool FAST>
__global__ void __launch_bounds__(256, 3) assemble_from_deriv_kernel(const double *__restrict__ It, const double *__restrict__ Ix,
                                           const double *__restrict__ Iy, int NC, const double2 *__restrict__ uv,
                                           const double2 *__restrict__ duv, int H, int W, PenaltySet ps, LinSys sys) {
  __shared__ double2 sH[8][32], sV[8][32], sL[8];
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = x < W && y < H;
  const long long HW = (long long)H * W;
  long long off = (long long)blockIdx.z * HW;
  long long i = (long long)y * W + x;
  DataTerm dt = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (in) {
    const double2 dc = duv ? duv[off + i] : make_double2(0.0, 0.0);
    if (NC == 1) {
      Deriv dv;
      dv.It = It[off + i]; dv.Ix = Ix[off + i]; dv.Iy = Iy[off + i];
      dt = data_term_single<FAST>(ps, dv, dc);
    } else {
      DataAccum acc;
      for (int c = 0; c < NC; ++c) {
        const long long j = ((long long)blockIdx.z * NC + c) * HW + i;
        Deriv dv;
        dv.It = It[j]; dv.Ix = Ix[j]; dv.Iy = Iy[j];
        acc.add(ps, dv, dc);
      }
      dt = acc.finish(NC);
    }
  }
  assemble_pixel<true, FAST>(ps, uv + off, duv ? duv + off : nullptr, H, W, x, y, in, dt, sys, off + i, sH, sV, sL);
}

int k_warp_assemble(b200flow_ctx *ctx, const double *frames, long long bstride, int NC, const double *I1x, const double *I1y,
                    const double4 *src2, const double2 *uv, const double2 *duv, int B, int H, int W, int interp, double blend,
                    const PenaltySet &ps_in, LinSys sys, double *It, double *Ix, double *Iy) {
  if (interp < 0 || interp > 2) return set_err(ctx, B200FLOW_EINVAL, "Unknown interpolation method: %d", interp);
  PenaltySet ps = ps_in;
  if (sys.D != nullptr) BF_TRY(attach_pow_table(ctx, &ps));
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
  int do_assemble = sys.D != nullptr;
  if (NC == 1 && ps.fast)
    BF_LAUNCH(ctx, (warp_assemble_kernel<false, true>), grd, blk, 0, frames, bstride, NC, I1x, I1y, src2, uv, duv, H, W, interp,
              blend, ps, sys, It, Ix, Iy, do_assemble);
  else if (NC == 1)
    BF_LAUNCH(ctx, (warp_assemble_kernel<false, false>), grd, blk, 0, frames, bstride, NC, I1x, I1y, src2, uv, duv, H, W, interp,
              blend, ps, sys, It, Ix, Iy, do_assemble);
  else
    BF_LAUNCH(ctx, (warp_assemble_kernel<true, false>), grd, blk, 0, frames, bstride, NC, I1x, I1y, src2, uv, duv, H, W, interp,
              blend, ps, sys, It, Ix, Iy, do_assemble);
  return 0;
}

int k_assemble_from_deriv(b200flow_ctx *ctx, const double *It, const double *Ix, const double *Iy, int NC, const double2 *uv,
                          const double2 *duv, int B, int H, int W, const PenaltySet &ps_in, LinSys sys) {
  PenaltySet ps = ps_in;
  BF_TRY(attach_pow_table(ctx, &ps));
  dim3 blk(32, 8), grd((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), B);
  if (NC == 1 && ps.fast)
    BF_LAUNCH(ctx, assemble_from_deriv_kernel<true>, grd, blk, 0, It, Ix, Iy, NC, uv, duv, H, W, ps, sys);
  else
    BF_LAUNCH(ctx, assemble_from_deriv_kernel<false>, grd, blk, 0, It, Ix, Iy, NC, uv, duv, H, W, ps, sys);
  return 0;
}

}  // namespace bf
