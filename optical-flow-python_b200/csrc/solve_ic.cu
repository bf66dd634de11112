// solve_ic.cu -- kernel group (4), default "exact" solver: mixed-precision PCG with reliable updates (as
// pcg_mixed_kernel in solve.cu) preconditioned by a TILE-LOCAL 2x2-BLOCK INCOMPLETE CHOLESKY IC(0) instead of block
// Jacobi.  Replaces scipy.sparse.linalg.spsolve behind BaseOpticalFlow._solve_linear_system (base.py:87-114).
//
// Why: the solver is HBM-bound (solve.cu), so the only way to make a solve much faster is FEWER iterations, and the
// preconditioner may spend on-chip work freely as long as it adds little HBM traffic.  IC(0) of the five-point block
// stencil needs no fill-in storage: M = (P + L) P^-1 (P + L)^T with L the strictly lower part of A itself and
// P_i = A_ii - sum_{j in {left, up}} W_ij P_j^-1 W_ij (W_ij = diag(w_u, w_v) of the edge), so the preconditioner is
// the 12 B/pixel P^-1 that block Jacobi already stores plus a bf16 copy of the edge weights (8 B/pixel).  Couplings
// are cut at the borders of 8 x 8 sub-tiles, which makes the two triangular solves local: every warp stages one STRIP
// (8 rows x 32 columns = four sub-tiles) in shared memory and sweeps it along the anti-diagonals (lane = (sub-tile,
// row), 15 steps forward + 15 back, neighbour rows through warp shuffles, PUSH form so that both sweeps use only the
// pixel's own coefficients).  scripts/ic_proto.py / ic32_proto.py (real Classic+NL systems, fp32 factor, truncated
// bf16 edges, exactly as here): 470 -> 231 iterations (alpha = 0), 60 -> 27 (alpha = 1); bench step 3015 -> 1485.
//
// Algorithmic bytes per pixel-iteration (fp32 working set):
//   A  read z 8, p_old 8, y 8, D 8, a12 4, WH 8, WV 8; write p 8, y 8, Ap 8                              = 76
//   B  read r 8, Ap 8, {P^-1 12, 4 edge weights as truncated bf16 8} by cp.async; write r 8, z 8          = 52
//                                                                                            total        128 B
// Measured per phase at 16 x 480 x 640 (IC_TIMERS build): A 66 us = 5.6 TB/s, B 44 us = 5.8 TB/s, two grid barriers +
// scalar reductions 13 us.  The unit of work dealt to the CTAs is the strip (19 200 strips over 296 CTAs: < 1 %
// imbalance); in phase A a strip is the CTA's 32 x 8 thread tile, in phase B the CTA's strips go round-robin to its
// eight warps, which never synchronise with each other there.
// Everything else (all-system scalar tracking, re-dealing of the active strips, reliable updates on the fp64 true
// residual, determinism) is as in pcg_mixed_kernel.
#include <mutex>
#include "solve_shared.cuh"

namespace cg = cooperative_groups;

namespace bf {

#ifndef IC_THREADS
#define IC_THREADS 256                         // threads per CTA, 2 CTAs / SM (4 CTAs of 128 threads measured: no gain)
#endif
constexpr int IC_NSTRIP = IC_THREADS / 32;     // strips (8 rows x 32 columns) staged at a time: one per warp in phase B
constexpr int IC_ROWS = IC_NSTRIP * 8;         // staged rows
constexpr int IC_PITCH = 34;                   // staged row pitch in elements (see ic_slot)
constexpr int IC_PAD = 8;                      // guard elements before / after each staged array (idle wavefront steps)
constexpr int IC_NSLOT = IC_ROWS * IC_PITCH + 2 * IC_PAD;
static_assert(IC_THREADS == 256, "phase A maps one 32 x 8 thread tile onto one strip");
#ifndef IC_SW
#define IC_SW 8                                // sub-tile = 8 rows x IC_SW columns, 8 or 16 (scripts/ic32_proto.py: 231 / 212 iterations vs 470)
#endif
static_assert(IC_SW == 8 || IC_SW == 16, "sub-tile width");
// anti-diagonals of one sub-tile (IC_SW + 7), rounded up to a multiple of the unroll factor: a surplus step is idle for every lane
constexpr int IC_STEPS = IC_SW == 8 ? 15 : 24;
#ifndef IC_SWEEP_UNROLL_N
#define IC_SWEEP_UNROLL_N (IC_SW == 8 ? 5 : 6)  // of the wavefront steps (a full unroll exhausts the 7 predicate registers)
#endif
constexpr int IC_SWEEP_UNROLL = IC_SWEEP_UNROLL_N;
#ifdef B200FLOW_TUNING
constexpr bool IC_TUNING = true;
#else
constexpr bool IC_TUNING = false;
#endif
// staged per pixel: float4 {i11, i12, i22, bf16x2 {wuh, wvh}} + bf16x2 {wuv, wvv} + float2 r = 28 B
constexpr size_t IC_SMEM = (size_t)IC_NSLOT * (sizeof(float4) + sizeof(unsigned) + sizeof(float2));   // 60 KB at 256 threads

// Shared-memory slot of staged pixel (row, col), row = 8 * strip + j.  A pitch of 34 elements makes both access patterns
// bank-conflict free for 4-, 8- and 16-byte elements without any index arithmetic in the wavefront: a warp touching one
// row (lane = column), and a warp walking the anti-diagonals of its four 8 x 8 sub-tiles (lane = (sub-tile q, row j),
// column 8 q + step - j): slot = const + 33 j + 8 q + step, i.e. (j + 8 q + step) mod 32 -- and `step` enters as a
// compile-time immediate of the unrolled loop.
__device__ __forceinline__ int ic_slot(int row, int col) { return IC_PAD + row * IC_PITCH + col; }

// edge weights are kept as truncated bfloat16 pairs in the preconditioner (never in the operator): rounding TOWARDS
// ZERO keeps the perturbed matrix a diagonally dominant M-matrix, so the incomplete factorisation still exists, and the
// factor is computed from the same truncated values, so M stays symmetric positive definite.  Iteration counts are
// unchanged (scripts/ic32_proto.py: 211 vs 211).
__device__ __forceinline__ unsigned ic_pack(float lo, float hi) {
  return (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u);
}
__device__ __forceinline__ float ic_lo(unsigned p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float ic_hi(unsigned p) { return __uint_as_float(p & 0xffff0000u); }

// ---- shared-memory access by 32-bit shared-space address (computed once per kernel): keeps the generic-to-shared
//      window arithmetic (an S2R of the cluster CTA id per access) out of the wavefront loops
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned lds32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(unsigned a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts32(unsigned a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// global loads of phase B as volatile asm: the compiler otherwise sinks each load down to its first use to save
// registers (the kernel sits at the 128-register cap), which turns one batch of loads per half-tile into several
// dependent round trips to DRAM.  Volatile asm statements keep their program order.
__device__ __forceinline__ float2 ldg_f2(const float2 *p) {
  float2 v;
  asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_nc_f(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_nc_u2(const uint2 *p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// asynchronous global -> shared copies (LDGSTS): the preconditioner's coefficients go straight into the staged tile without
// passing through registers; src_bytes = 0 zero-fills the destination (pixels outside the image)
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const void *src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// block-wide sum of two accumulators over IC_THREADS threads; result valid in thread 0
__device__ __forceinline__ void ic_block_sum2(double &a, double &b, double (*sm)[IC_THREADS / 32]) {
  a = warp_sum(a);
  b = warp_sum(b);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // protect sm reuse
  if (l == 0) { sm[0][w] = a; sm[1][w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < IC_THREADS / 32 ? sm[0][l] : 0.0;
    b = l < IC_THREADS / 32 ? sm[1][l] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

struct IcSmem {            // shared-space byte addresses of the three staged arrays
  unsigned c0, cw, r;      // float4 {i11, i12, i22, bf16x2 wh}, bf16x2 wv, float2 r -> t -> z
};

// ---- wavefront sweeps of one strip (8 rows x 32 columns = four 8 x 8 sub-tiles) by the warp that staged it ----------
// forward (P + L) t = r, backward (I + P^-1 L^T) z = t, in place (r -> t -> z).  Lane (q, j) walks row j of sub-tile q:
// at step s it is at column s - j of the sub-tile, so the rows of a sub-tile form a wavefront and the neighbour row's
// value arrives by a warp shuffle of the previous step.  Phase B is bound by instruction issue, not by latency (all 16
// warps of an SM sweep concurrently), so the loops are fully unrolled (shared-memory offsets become immediates), carry
// no masking arithmetic and no address arithmetic:
//  * idle steps (column outside 0..7) read guard / neighbouring slots and compute garbage that nobody consumes: an
//    active lane only ever reads the shuffle of a lane that was active at the same column one step earlier, and the
//    own-row carries (cu, cv / zu, zv) and the store are predicated on the step being active;
//  * the coupling between vertically adjacent sub-tiles is cut by the down-edge weights {wuv, wvv} of every eighth row
//    being 0 (staged as 0, and forced to 0 on that row's idle steps), so what those rows push down is always 0 (the
//    shuffle rotates, lane 0 <- lane 31 is such a row);
//  * guard slots and the two pad columns are zeroed once per kernel, so idle-step garbage is always finite.
__device__ __forceinline__ void ic_sweeps(const IcSmem sm, int w, int lane) {
  const unsigned full = 0xffffffffu;
  // with 16-wide sub-tiles a strip holds two of them: lanes 16..31 repeat the work (and the identical stores) of lanes 0..15
  const int q = IC_SW == 8 ? lane >> 3 : (lane >> 3) & 1, j = lane & 7;
  const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
  const int slot0 = ic_slot(8 * w + j, IC_SW * q - j);   // slot of step 0 (column -j); step s is slot0 + s
  const unsigned a0 = sm.c0 + slot0 * 16, ar = sm.r + slot0 * 8, aw = sm.cw + slot0 * 4;
  const unsigned actmask = ((1u << IC_SW) - 1u) << j;    // bit s set: step s is inside the sub-tile
  const unsigned keep = j == 7 ? 0u : 0xffffffffu;       // last row of a sub-tile never pushes down / pulls up
  {
    float cu = 0.f, cv = 0.f, pu = 0.f, pv = 0.f;        // pushed from the left (own row) / pushed down to the next row
    float4 A = lds128(a0);
    float2 R = lds64(ar);
    unsigned Wv = lds32(aw);
#pragma unroll IC_SWEEP_UNROLL
    for (int s = 0; s < IC_STEPS; ++s) {
      float4 An = A; float2 Rn = R; unsigned Wvn = Wv;
      if (s + 1 < IC_STEPS) { An = lds128(a0 + (s + 1) * 16); Rn = lds64(ar + (s + 1) * 8); Wvn = lds32(aw + (s + 1) * 4); }
      const float uu = __shfl_sync(full, pu, up_lane), uw = __shfl_sync(full, pv, up_lane);
      unsigned am = actmask;
      asm volatile("" : "+r"(am));                     // keeps the 15 step predicates from being hoisted (7 predicate registers)
      const bool act = (am >> s) & 1u;
      const float su = (R.x + cu) + uu, sv = (R.y + cv) + uw;
      const float tu = fmaf(A.y, sv, A.x * su), tv = fmaf(A.z, sv, A.y * su);
      const unsigned wh = __float_as_uint(A.w);
      const unsigned wv = Wv & keep;
      pu = ic_lo(wv) * tu; pv = ic_hi(wv) * tv;
      if (act) {
        cu = ic_lo(wh) * tu; cv = ic_hi(wh) * tv;
        sts64(ar + s * 8, make_float2(tu, tv));
      }
      A = An; R = Rn; Wv = Wvn;
    }
  }
  {
    float zu = 0.f, zv = 0.f;                            // own z of the last active step = right neighbour, offered to the row above
    float4 A = lds128(a0 + (IC_STEPS - 1) * 16);
    float2 T = lds64(ar + (IC_STEPS - 1) * 8);
    unsigned Wv = lds32(aw + (IC_STEPS - 1) * 4);
#pragma unroll IC_SWEEP_UNROLL
    for (int s = IC_STEPS - 1; s >= 0; --s) {
      float4 An = A; float2 Tn = T; unsigned Wvn = Wv;
      if (s > 0) { An = lds128(a0 + (s - 1) * 16); Tn = lds64(ar + (s - 1) * 8); Wvn = lds32(aw + (s - 1) * 4); }
      const float du = __shfl_sync(full, zu, dn_lane), dv = __shfl_sync(full, zv, dn_lane);
      unsigned am = actmask;
      asm volatile("" : "+r"(am));
      const bool act = (am >> s) & 1u;
      const unsigned wh = __float_as_uint(A.w);
      // z = T + P^-1 (WH z_right + WV z_down), z_right = own previous z
      const unsigned wv = Wv & keep;
      const float au = fmaf(ic_lo(wv), du, ic_lo(wh) * zu), av = fmaf(ic_hi(wv), dv, ic_hi(wh) * zv);
      const float nu = T.x + fmaf(A.y, av, A.x * au), nv = T.y + fmaf(A.z, av, A.y * au);
      if (act) {
        zu = nu; zv = nv;
        sts64(ar + s * 8, make_float2(nu, nv));
      }
      A = An; T = Tn; Wv = Wvn;
    }
  }
}

// incomplete factorisation of the staged strip: c0 holds {a11 + sum w_u, a12, a22 + sum w_v, wh} on entry and the
// inverted pivot blocks {i11, i12, i22, wh} on exit.  Same wavefront as the sweeps (three pushed values per direction);
// runs once per solve, so it keeps a rolled loop and explicit masking.
__device__ __forceinline__ void ic_factor(const IcSmem sm, int w, int lane) {
  const unsigned full = 0xffffffffu;
  const int q = IC_SW == 8 ? lane >> 3 : (lane >> 3) & 1, j = lane & 7;
  const int up_lane = (lane + 31) & 31;
  const int slot0 = ic_slot(8 * w + j, IC_SW * q - j);
  float c11 = 0.f, c12 = 0.f, c22 = 0.f, p11 = 0.f, p12 = 0.f, p22 = 0.f;
#pragma unroll 1
  for (int s = 0; s < IC_STEPS; ++s) {
    const bool act = (unsigned)(s - j) < (unsigned)IC_SW;
    const int idx = slot0 + s;
    const float4 A = lds128(sm.c0 + idx * 16);
    const unsigned wh = __float_as_uint(A.w), wv = lds32(sm.cw + idx * 4);
    const float whu = ic_lo(wh), whv = ic_hi(wh), wvu = ic_lo(wv), wvv = ic_hi(wv);
    const float u11 = __shfl_sync(full, p11, up_lane), u12 = __shfl_sync(full, p12, up_lane),
                u22 = __shfl_sync(full, p22, up_lane);           // 0 from a sub-tile's last row (its wuv, wvv are staged as 0)
    double d11 = (double)A.x - ((double)c11 + (double)u11);
    double d12 = (double)A.y - ((double)c12 + (double)u12);
    double d22 = (double)A.z - ((double)c22 + (double)u22);
    double det = d11 * d22 - d12 * d12;
    if (!(d11 > 0.0 && d22 > 0.0 && det > 1e-10 * d11 * d22)) {      // pivot breakdown: keep the unmodified block
      d11 = (double)A.x; d12 = (double)A.y; d22 = (double)A.z;
      det = d11 * d22 - d12 * d12;
    }
    float i11, i12, i22;
    if (d11 > 0.0 && d22 > 0.0 && det > 1e-14 * d11 * d22) {
      const double inv = 1.0 / det;
      i11 = (float)(d22 * inv); i12 = (float)(-d12 * inv); i22 = (float)(d11 * inv);
    } else {                                                           // as block Jacobi's scalar fallback (solve.cu)
      i11 = d11 > 1e-12 ? (float)(1.0 / d11) : 0.f;
      i22 = d22 > 1e-12 ? (float)(1.0 / d22) : 0.f;
      i12 = 0.f;
    }
    if (act) {
      sts128(sm.c0 + idx * 16, make_float4(i11, i12, i22, A.w));
      c11 = whu * whu * i11; c12 = whu * whv * i12; c22 = whv * whv * i22;
      p11 = wvu * wvu * i11; p12 = wvu * wvv * i12; p22 = wvv * wvv * i22;
    } else {
      c11 = 0.f; c12 = 0.f; c22 = 0.f; p11 = 0.f; p12 = 0.f; p22 = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-band split of ONE system over several GPUs (reference loop: ba.py:140-206, one solve per warp
// iteration).  Rank g owns the pixel rows [y0, y1) (multiples of 8, so the 8 x 8 IC sub-tiles never straddle a band and the
// preconditioner -- hence every iterate -- is the one the single-GPU solve has).  Every rank holds the full coefficient
// arrays (the assembly is replicated); the Krylov vectors live band-local at the SAME offsets of every rank's arena, which
// is one cudaMalloc block mapped into the peers by CUDA IPC: a neighbour's row is read by adding the byte distance between
// the two blocks to the local pointer (direct NVLink P2P loads, no staging, no NCCL).
//   per iteration:  phase A reads ONE row of z and of p_old from each neighbour band (2 x W x 8 B per side);
//                   each of the two barriers carries the rank-local dot products to every peer (2 doubles, pushed by
//                   CTA 0 over NVLink) and waits for the peers' sequence flags -- no host, no collective library.
// All ranks sum the per-rank scalars in rank order, so every rank (and every CTA) takes bit-identical control decisions.
// ------------------------------------------------------------------------------------------------------------------
struct BandSync {                        // at offset 0 of every rank's arena block
  unsigned long long seq;                // barriers this rank has completed (written by its CTA 0 only; persists across kernels)
  unsigned long long release;            // local release: this rank's CTAs may leave barrier `release`
  unsigned long long flag[B200FLOW_MAX_BAND_RANKS];     // flag[r]: last barrier rank r has arrived at (written by rank r, remotely)
  double xs[2][B200FLOW_MAX_BAND_RANKS][2];             // (unused since the sums travel with their sequence number, below)
  // ll[parity][r][i] = {rank r's i-th local sum, the barrier's sequence number}: ONE 16-byte store each, so the value and the
  // flag that validates it arrive together and no fence has to order them (the "LL" scheme of collective libraries)
  unsigned long long ll[2][B200FLOW_MAX_BAND_RANKS][2][2];
  double gs[2][2];                       // the global sums of that barrier (written by the local CTA 0)
  unsigned long long error;              // != 0: a spin loop timed out (a peer died); the result is invalid
};

struct BandParams {
  int rank, world, y0, y1;               // this rank's rows
  long long up_delta, dn_delta;          // byte distance from this rank's block to the block of the rank above / below (0: none)
  BandSync *self;
  BandSync *peer[B200FLOW_MAX_BAND_RANKS];   // every rank's BandSync as mapped here (peer[rank] == self)
};

__device__ __forceinline__ void st_sys_f64(double *p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys_ll(unsigned long long *p, double v, unsigned long long seq) {
  asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(seq) : "memory");
}
__device__ __forceinline__ bool ld_sys_ll(const unsigned long long *p, unsigned long long seq, double &v) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  v = __longlong_as_double((long long)a);
  return b >= seq;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// neighbour-band rows: coherent at system scope, never from this SM's L1
__device__ __forceinline__ float2 ld_sys_f2(const float2 *p) {
  float2 v;
  asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_sys_d2(const double2 *p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
template <typename T>
__device__ __forceinline__ const T *band_shift(const T *p, long long delta) {
  return reinterpret_cast<const T *>(reinterpret_cast<const char *>(p) + delta);
}
constexpr long long BAND_SPIN_LIMIT = 1LL << 25;       // ~ 10-20 s of polling: a dead peer must not hang the GPU

// Barrier + reduction of the band solver.  All CTAs arrive on the local monotonic counter; CTA 0 waits for them, reduces the
// rank-local partials pa / pb (slots c_lo .. c_hi) in a fixed order, pushes the two sums and then its sequence flag to every
// rank, waits for every rank's flag, sums the per-rank values in rank order and releases the local CTAs.  Returns the
// global sums in every thread.
__device__ __forceinline__ void band_barrier_reduce(const BandParams &bp, unsigned *arrive, unsigned &target,
                                                    unsigned long long &seq, const double *pa, const double *pb, int c_lo,
                                                    int c_hi, double &ga, double &gb) {
  __syncthreads();
  seq += 1;
  const int slot = (int)(seq & 1ull);
  BandSync *self = bp.self;
  // once a barrier has timed out every later one gives up at once and hands NaN sums to the solver, which then stops
  const long long limit = *(volatile unsigned long long *)&self->error != 0ull ? 0 : BAND_SPIN_LIMIT;
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(arrive, 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int lane = threadIdx.x;
    long long spins = 0;
    if (lane == 0) {
      while (*(volatile unsigned *)arrive < target && ++spins < limit) {}
      __threadfence();
    }
    __syncwarp();
    double va = 0.0, vb = 0.0;
    for (int c = c_lo + lane; c <= c_hi; c += 32) {
      va += ((const volatile double *)pa)[c];
      if (pb) vb += ((const volatile double *)pb)[c];
    }
    va = warp_sum(va);
    vb = warp_sum(vb);
    // the exchange with the peers runs on one lane per rank: lane r publishes to rank r and then waits for rank r's flag
    // (world <= 8 <= 32), so its latency is that of ONE NVLink round trip, not of `world` of them
    // Ordering without system-scope fences (each costs microseconds): (i) the band's vectors were written to THIS GPU's
    // L2 before the CTAs arrived (their gpu-scope fences) and the neighbours read them there, over NVLink, with loads that
    // bypass their own L1; (ii) each sum travels in one 16-byte store together with the sequence number that validates it.
    const bool act = lane < bp.world;
    double ra = 0.0, rb = 0.0;
    if (act) {
      st_sys_ll(&bp.peer[lane]->ll[slot][bp.rank][0][0], va, seq);
      st_sys_ll(&bp.peer[lane]->ll[slot][bp.rank][1][0], vb, seq);
      bool oka = false, okb = false;
      while (!(oka && okb) && ++spins < limit) {
        oka = ld_sys_ll(&self->ll[slot][lane][0][0], seq, ra);
        okb = ld_sys_ll(&self->ll[slot][lane][1][0], seq, rb);
      }
    }
    const bool timed_out = __any_sync(0xffffffffu, spins >= limit);
    double sa = 0.0, sb = 0.0;
    for (int r = 0; r < bp.world; ++r) {            // rank order: the same bits on every rank
      sa += __shfl_sync(0xffffffffu, ra, r);
      sb += __shfl_sync(0xffffffffu, rb, r);
    }
    if (lane == 0) {
#ifdef BAND_DEBUG
      printf("[band] rank %d barrier %llu: global %.6e %.6e\n", bp.rank, seq, sa, sb);
#endif
      if (timed_out) {                              // timed out (now or earlier): poison the sums, the solve ends as failed
        if (limit > 0) self->error = seq;
        sa = sb = __longlong_as_double(0x7ff8000000000000LL);
      }
      self->gs[slot][0] = sa;
      self->gs[slot][1] = sb;
      self->seq = seq;
      __threadfence();
      *(volatile unsigned long long *)&self->release = seq;
    }
  }
  if (threadIdx.x == 0) {
    long long spins = 0;
    while (*(volatile unsigned long long *)&self->release < seq && ++spins < 4 * BAND_SPIN_LIMIT) {}
    __threadfence();
  }
  __syncthreads();
  ga = ((volatile double *)self->gs[slot])[0];
  gb = ((volatile double *)self->gs[slot])[1];
}

// true residual at pixel i of the solution x + y + alpha p, fp64, evaluated on the fly at the five stencil points
template <bool BAND>
__device__ __forceinline__ double2 ic_true_residual(const MixParams &P, const BandParams &bp, long long i, int px, int py,
                                                    const float2 *pnew, double alpha) {
  const LinSys &S = P.sys;
  const int W = S.W, H = S.H;
  const double2 *x = P.x;
  const float2 *y = P.m.y;
  const long long jl = px > 0 ? i - 1 : i, jr = px + 1 < W ? i + 1 : i;
  const long long ju = py > 0 ? i - W : i, jd = py + 1 < H ? i + W : i;
#define XTRUE(j, out)                                                           \
  {                                                                             \
    double2 xx = x[j]; float2 yy = y[j], pp = pnew[j];                          \
    out = make_double2(xx.x + ((double)yy.x + alpha * (double)pp.x),            \
                       xx.y + ((double)yy.y + alpha * (double)pp.y));           \
  }
  // a row of the neighbour band: the same offsets in the neighbour's arena block, read over NVLink
#define XTRUE_PEER(j, delta, out)                                               \
  {                                                                             \
    double2 xx = ld_sys_d2(band_shift(x + (j), delta));                         \
    float2 yy = ld_sys_f2(band_shift(y + (j), delta)), pp = ld_sys_f2(band_shift(pnew + (j), delta)); \
    out = make_double2(xx.x + ((double)yy.x + alpha * (double)pp.x),            \
                       xx.y + ((double)yy.y + alpha * (double)pp.y));           \
  }
  double2 c, nl, nr, nu, nd;
  XTRUE(i, c) XTRUE(jl, nl) XTRUE(jr, nr)
  if (BAND && py == bp.y0 && bp.up_delta != 0) XTRUE_PEER(ju, bp.up_delta, nu) else XTRUE(ju, nu)
  if (BAND && py + 1 == bp.y1 && bp.dn_delta != 0) XTRUE_PEER(jd, bp.dn_delta, nd) else XTRUE(jd, nd)
#undef XTRUE
#undef XTRUE_PEER
  const double2 sd = __ldg(&S.D[i]), swr = __ldg(&S.WH[i]), swd = __ldg(&S.WV[i]);
  const double2 swl = __ldg(&S.WH[jl]), swu = __ldg(&S.WV[ju]);   // multiplied by a zero difference when jl == i / ju == i
  const double sa12 = __ldg(&S.a12[i]);
  double au = sd.x * c.x + sa12 * c.y;
  double av = sa12 * c.x + sd.y * c.y;
  au += swr.x * (c.x - nr.x); av += swr.y * (c.y - nr.y);
  au += swl.x * (c.x - nl.x); av += swl.y * (c.y - nl.y);
  au += swd.x * (c.x - nd.x); av += swd.y * (c.y - nd.y);
  au += swu.x * (c.x - nu.x); av += swu.y * (c.y - nu.y);
  const double2 rb = __ldg(&S.rhs[i]);
  return make_double2(rb.x - au, rb.y - av);
}

#ifndef IC_UA
#define IC_UA 2
#endif

// Grid barrier of the persistent kernel (the launch is cooperative, so every CTA is resident): a monotonic counter
// (flags[0], zeroed by the launcher) -- barrier number k is passed when the counter reaches k * gridDim.x.  One atomic
// and one spinning thread per CTA; the gpu-scope fences publish the CTA's writes before the arrival and drop the SM's L1
// (CCTL.IVALL) before the CTA reads what the other SMs wrote.  Measured against cooperative_groups' grid.sync()
// (-DIC_CG_BARRIER): 10.4 instead of 14.3 us per iteration at 16 x 30x40, 144.5 instead of 147.5 ms of solver time per
// bench step.
#ifndef IC_CG_BARRIER
__device__ __forceinline__ void ic_grid_barrier(unsigned *counter, unsigned &target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*(volatile unsigned *)counter < target) {}
    __threadfence();
  }
  __syncthreads();
}
#define IC_GRID_SYNC() ic_grid_barrier(bar_counter, bar_target)
#else
#define IC_GRID_SYNC() grid.sync()
#endif


// Cluster mode (small levels): ONE THREAD-BLOCK CLUSTER PER SYSTEM.  A level of 30 x 40 ... 120 x 160 pixels is latency bound:
// its iteration is two grid barriers (atomic counter + polling over up to 296 CTAs, ~6 us each) around a few microseconds of
// work, and all systems of the batch wait for each other.  Here a system's strips are dealt over the CTAs of one cluster, the
// two barriers of an iteration are the hardware cluster barrier (barrier.cluster), and the clusters never synchronise with
// each other: a system that has converged simply ends.  The vectors stay where they are (global memory: a small level lives
// in L2); the body below is the same code with B = 1 and every pointer moved to the cluster's system.
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  // the gpu-scope fences on both sides do for the data what they do in ic_grid_barrier: publish this CTA's rows, and drop the
  // SM's L1 before rows written by the other CTAs of the cluster are read
  __syncthreads();
  if (threadIdx.x == 0) __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) __threadfence();
  __syncthreads();
}

enum { IC_MODE_GRID = 0, IC_MODE_BAND = 1, IC_MODE_CLUSTER = 2 };

template <int MODE>
__device__ __forceinline__ void pcg_ic_body(const MixParams &Pin, const BandParams &bp) {
  constexpr bool BAND = MODE == IC_MODE_BAND, CLUS = MODE == IC_MODE_CLUSTER;
#ifdef IC_CG_BARRIER
  cg::grid_group grid = cg::this_grid();
#endif
  const int Btot = Pin.sys.B;                      // systems of the launch (cluster mode: one of them per cluster)
  const int G = CLUS ? (int)cluster_nctarank() : (int)gridDim.x, cta = CLUS ? (int)cluster_ctarank() : (int)blockIdx.x;
  const int sysid = CLUS ? (int)blockIdx.x / G : 0;
  MixParams Pc;                                    // cluster mode: the parameter block of this cluster's system
  if (CLUS) {
    Pc = Pin;
    const long long o = (long long)sysid * Pin.sys.H * Pin.sys.W;
    Pc.sys.B = 1;
    Pc.sys.D += o; Pc.sys.a12 += o; Pc.sys.WH += o; Pc.sys.WV += o; Pc.sys.rhs += o;
    Pc.x += o;
    Pc.m.r += o; Pc.m.z += o; Pc.m.p += o; Pc.m.p2 += o; Pc.m.Ap += o; Pc.m.y += o;
    Pc.m.D += o; Pc.m.WH += o; Pc.m.WV += o; Pc.m.a12 += o;
    Pc.w.ic_c0 += o; Pc.w.ic_cw += o;
    Pc.w.partial += (long long)sysid * 5 * G;      // five slices of G partials for this system
  }
  const MixParams &P = CLUS ? Pc : Pin;
  const LinSys &S = P.sys;
  const int Y0 = BAND ? bp.y0 : 0;                 // first pixel row of this rank's band
  const int Y1 = BAND ? bp.y1 : S.H;               // one past its last row
  const int H = S.H, W = S.W, B = S.B;
  const long long HW = (long long)H * W;
  const long long n_all = (long long)B * HW;
  const int tps = P.tiles_per_sys;          // strips (8 rows x 32 columns) per system
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  int GW = 32;
  while (GW > 1 && B * GW > IC_THREADS) GW >>= 1;

  extern __shared__ __align__(16) unsigned char ic_smem[];
  IcSmem sm;
  sm.c0 = (unsigned)__cvta_generic_to_shared(ic_smem);             // float4 {i11, i12, i22, bf16x2 {wuh, wvh}}
  asm volatile("mov.u32 %0, %0;" : "+r"(sm.c0));                   // opaque: must live in a register, never rematerialised
  sm.r = sm.c0 + IC_NSLOT * 16;                                    // float2 r -> t -> z
  sm.cw = sm.r + IC_NSLOT * 8;                                     // bf16x2 {wuv, wvv}
  for (int i = threadIdx.x; i < (int)(IC_SMEM / 4); i += IC_THREADS) sts32(sm.c0 + i * 4, 0u);   // guards / pad columns = 0
  __syncthreads();

  __shared__ double sm_red[2][IC_THREADS / 32];
  __shared__ double s_rz[MAXB], s_bb[MAXB], s_alpha[MAXB], s_beta[MAXB], s_maxr2[MAXB], s_rzprev[MAXB];
  __shared__ double s_ta[MAXB], s_tb[MAXB];
  __shared__ int s_state[MAXB];          // 0 active, 1 finished, 2 converged: final flush pending, 3 reliable update in progress
  __shared__ int s_bad[MAXB], s_flush[MAXB], s_restarts[MAXB], s_stalls[MAXB];
  __shared__ double s_lasttrue[MAXB];
  __shared__ int s_act[MAXB], s_pos[MAXB];   // compact list of unfinished systems and its inverse
  __shared__ int s_nact, s_tpc;

  double *part_a = P.w.partial;                         // [B][G]  p.Ap
  double *part_b = P.w.partial + (long long)B * G;      // [B][G]  r.z
  double *part_c = P.w.partial + 2LL * B * G;           // [B][G]  r.r
  // reliable-update partials get their own slices: a fast CTA reaches its reliable update while a slow CTA is still
  // summing the phase-B slots (no grid barrier in between), so it must not write part_b / part_c there
  double *part_d = P.w.partial + 3LL * B * G;           // [B][G]  r.z of the replaced residual
  double *part_e = P.w.partial + 4LL * B * G;           // [B][G]  r.r of the replaced residual
  float4 *C0 = P.w.ic_c0;                               // [n_all] {i11, i12, i22, bf16x2 {wuh, wvh}}: inverted IC pivot blocks + edges
  unsigned *CW = P.w.ic_cw;                             // [n_all] bf16x2 {wuv, wvv}
  float2 *r = P.m.r, *z = P.m.z, *Ap = P.m.Ap, *y = P.m.y;
  float2 *pold = P.m.p, *pnew = P.m.p2;
  float2 *Df = P.m.D, *WHf = P.m.WH, *WVf = P.m.WV;
  float *a12f = P.m.a12;
  double2 *x = P.x;
  int *done_g = P.w.flags + 1 + sysid;
#ifndef IC_CG_BARRIER
  unsigned *bar_counter = reinterpret_cast<unsigned *>(P.w.flags);   // flags[0]: zeroed before every launch
  unsigned bar_target = 0u;
#endif
  unsigned long long band_seq = 0ull;
  if (BAND) band_seq = *(volatile unsigned long long *)&bp.self->seq;   // this rank's barrier count so far (own writes only)
  int *iters_g = P.w.flags + 1 + Btot + sysid;
  double *relres_g = P.w.scal + sysid;

  // ---- strip ownership: the strips of the unfinished systems, in compact order, are dealt to the CTAs in
  //      contiguous chunks of s_tpc; recomputed (identically by every CTA) whenever a system finishes
#define REMAP() MAP_WHERE(s_state[b] != 1)
  // the same dealing restricted to the systems that satisfy `cond` (the reliable update deals the systems it replaces the
  // residual of over the WHOLE grid, see below)
#define MAP_WHERE(cond)                                                                 \
  {                                                                                     \
    __syncthreads();                                                                    \
    if (tid == 0) {                                                                     \
      int n = 0;                                                                        \
      for (int b = 0; b < B; ++b) {                                                     \
        if (cond) { s_act[n] = b; s_pos[b] = n; ++n; } else s_pos[b] = -1;              \
      }                                                                                 \
      s_nact = n;                                                                       \
      long long tt = (long long)n * tps;                                                \
      s_tpc = tt > 0 ? (int)((tt + G - 1) / G) : 1;                                     \
    }                                                                                   \
    __syncthreads();                                                                    \
  }
#define C_LO(b) ((int)(((long long)s_pos[b] * tps) / s_tpc))
#define C_HI(b) ((int)((((long long)(s_pos[b] + 1)) * tps - 1) / s_tpc))
#define OWN_RANGE()                                                                              \
  const long long t0 = (long long)cta * s_tpc;                                                   \
  const long long tt_ = (long long)s_nact * tps;                                                 \
  const long long t1 = t0 + s_tpc < tt_ ? t0 + s_tpc : tt_;                                      \
  const int a_first = t0 < t1 ? (int)(t0 / tps) : 0;                                             \
  const int a_last = t0 < t1 ? (int)((t1 - 1) / tps) : -1;
#define TILE_RANGE(aslot)                                                                        \
  const int b = s_act[aslot];                                                                    \
  const long long tbase = (long long)(aslot) * tps;                                              \
  const long long ta = tbase > t0 ? tbase : t0;                                                  \
  const long long tb = tbase + tps < t1 ? tbase + tps : t1;                                      \
  const long long base = (long long)b * HW;
  // strip number -> strip coordinates.  One GPU: row-major.  Band mode: COLUMN-major, so that the strips of the band's first
  // and last row -- whose phase A waits for a remote load of the neighbour's halo row, ~3 us each -- are spread over all CTAs
  // instead of sitting back to back in the first and the last few (a CTA's chunk of consecutive strips is then a vertical run)
#ifdef IC_COLMAJOR      // tuning: column-major on one GPU too
#define STRIP_X(tl) ((tl) / P.tiles_y)
#define STRIP_Y(tl) ((tl) % P.tiles_y)
#else
#define STRIP_X(tl) (BAND ? (tl) / P.tiles_y : (tl) % P.tiles_x)
#define STRIP_Y(tl) (BAND ? (tl) % P.tiles_y : (tl) / P.tiles_x)
#endif
  // pixel of this thread in strip v of the current system (phase A: one pixel per thread, the strip is the 32 x 8 thread tile)
#define PIXEL_OF(v, px, py, i, ok)                                                               \
  {                                                                                              \
    const int tl = (int)((v) - tbase);                                                           \
    px = STRIP_X(tl) * 32 + tx;                                                                  \
    py = Y0 + STRIP_Y(tl) * 8 + ty;                                                              \
    ok = (v) < tb && px < W && py < Y1;                                                          \
    i = ok ? base + (long long)py * W + px : base;                                               \
  }
  // barrier + reduction of one phase: single GPU = the grid barrier, then every CTA re-reduces the partials; band mode =
  // band_barrier_reduce (one system, CTA 0 reduces and exchanges with the peers)
#ifdef IC_CG_BARRIER
#define SYNC_REDUCE(pa, pb, want) { IC_GRID_SYNC(); REDUCE_ALL(pa, pb, want) }
#else
#define SYNC_REDUCE(pa, pb, want)                                                               \
  if (CLUS) {                                                                                   \
    cluster_barrier();                                                                          \
    REDUCE_ALL(pa, pb, want)                                                                    \
  } else if (BAND) {                                                                            \
    double ga_, gb_;                                                                            \
    band_barrier_reduce(bp, bar_counter, bar_target, band_seq, (pa), (pb), C_LO(0), C_HI(0), ga_, gb_);   \
    if (tid == 0) { s_ta[0] = ga_; s_tb[0] = gb_; }                                             \
    __syncthreads();                                                                            \
  } else {                                                                                      \
    IC_GRID_SYNC();                                                                             \
    REDUCE_ALL(pa, pb, want)                                                                    \
  }
#endif
#define REDUCE_ALL(pa, pb, want)                                                        \
  {                                                                                     \
    const int b = tid / GW, gl = tid % GW;                                              \
    double va = 0.0, vb = 0.0;                                                          \
    if (b < B && s_state[b] == (want)) {                                                \
      const int c_lo = C_LO(b), c_hi = C_HI(b);                                         \
      const volatile double *qa = (pa) + (long long)b * G;                              \
      const volatile double *qb = (pb) ? (pb) + (long long)b * G : qa;                  \
      for (int c = c_lo + gl; c <= c_hi; c += GW) { va += qa[c]; vb += qb[c]; }         \
    }                                                                                   \
    for (int o = GW >> 1; o > 0; o >>= 1) {                                             \
      va += __shfl_xor_sync(0xffffffffu, va, o);                                        \
      vb += __shfl_xor_sync(0xffffffffu, vb, o);                                        \
    }                                                                                   \
    if (b < B && gl == 0) { s_ta[b] = va; s_tb[b] = vb; }                               \
    __syncthreads();                                                                    \
  }

  // ---- phase B of one strip: r -= alpha Ap (skipped when !use_ap), z = M^-1 r by the staged IC sweeps.  Thread
  //      (ty, tx) owns rows 0 .. 7 of column tx of strip ty; a warp stages, sweeps and writes back its strip on its own
  //      (warp barriers only), so the 16 warps of an SM overlap each other's loads and sweeps.  ONE round trip to
  //      memory per tile: the coefficients travel global -> shared asynchronously (cp.async, no registers) while
  //      r and Ap of all eight pixels are in flight in registers.  Accumulates r.z and r.r (fp32 over the thread's
  //      eight pixels, fp64 across tiles).
#define PHASE_B_TILE(t, alpha_f, use_ap, acc_rz, acc_rr)                                                       \
  {                                                                                                            \
    const int tl = (int)((t) - tbase);                                                                         \
    const int px = STRIP_X(tl) * 32 + tx;                                                                      \
    const int py0 = Y0 + STRIP_Y(tl) * 8;                                                                      \
    const long long i0 = base + (long long)py0 * W + px;                                                       \
    const int sl0 = ic_slot(ty * 8, tx);                                                                       \
    float2 rc[8], ac[8];                                                                                       \
    _Pragma("unroll")                                                                                          \
    for (int u = 0; u < 8; ++u) {                                                                              \
      const bool ok = px < W && py0 + u < Y1;                                                                  \
      const long long ii = ok ? i0 + (long long)u * W : base;                                                  \
      cp_async16(sm.c0 + (sl0 + u * IC_PITCH) * 16, C0 + ii, ok ? 16 : 0);                                     \
      cp_async4(sm.cw + (sl0 + u * IC_PITCH) * 4, CW + ii, ok ? 4 : 0);                                        \
      rc[u] = ldg_f2(r + ii);                                                                                  \
      ac[u] = (use_ap) ? ldg_f2(Ap + ii) : make_float2(0.f, 0.f);                                              \
    }                                                                                                          \
    float prr = 0.f, prz = 0.f;                                                                                \
    _Pragma("unroll")                                                                                          \
    for (int u = 0; u < 8; ++u) {                                                                              \
      const bool ok = px < W && py0 + u < Y1;                                                                  \
      float2 v = make_float2(0.f, 0.f);                                                                        \
      if (ok) {                                                                                                \
        v = (use_ap) ? make_float2(fmaf(-(alpha_f), ac[u].x, rc[u].x), fmaf(-(alpha_f), ac[u].y, rc[u].y)) : rc[u]; \
        if (use_ap) r[i0 + (long long)u * W] = v;                                                              \
        prr = fmaf(v.x, v.x, fmaf(v.y, v.y, prr));                                                             \
      }                                                                                                        \
      rc[u] = v;                                                                                               \
      sts64(sm.r + (sl0 + u * IC_PITCH) * 8, v);                                                               \
    }                                                                                                          \
    acc_rr += (double)prr;                                                                                     \
    cp_async_wait_all();                                                                                       \
    __syncwarp();                                                                                              \
    if (!(IC_TUNING && (P.debug & 1))) ic_sweeps(sm, ty, tx);                                                  \
    __syncwarp();                                                                                              \
    _Pragma("unroll")                                                                                          \
    for (int u = 0; u < 8; ++u) {                                                                              \
      if (px < W && py0 + u < Y1) {                                                                            \
        const float2 zz = lds64(sm.r + (sl0 + u * IC_PITCH) * 8);                                              \
        z[i0 + (long long)u * W] = zz;                                                                         \
        prz = fmaf(rc[u].x, zz.x, fmaf(rc[u].y, zz.y, prz));                                                   \
      }                                                                                                        \
    }                                                                                                          \
    acc_rz += (double)prz;                                                                                     \
  }

  // ---------------- init ----------------
  for (int b = tid; b < MAXB; b += IC_THREADS) { s_state[b] = b < B ? 0 : 1; s_bad[b] = 0; s_flush[b] = 0; s_restarts[b] = 0; s_stalls[b] = 0; s_lasttrue[b] = 1e300; }
  REMAP()
  {
    OWN_RANGE()
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      double acc_rz = 0.0, acc_bb = 0.0;
      for (long long t = ta + ty; t < tb; t += IC_NSTRIP) {     // strips of this CTA, dealt round-robin to its warps
        const int tl = (int)(t - tbase);
        const int px = STRIP_X(tl) * 32 + tx;
        const int py0 = Y0 + STRIP_Y(tl) * 8;
#pragma unroll 2
        for (int u = 0; u < 8; ++u) {
          const int py = py0 + u;
          const int sl = ic_slot(ty * 8 + u, tx);
          float2 v = make_float2(0.f, 0.f);
          float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
          unsigned cw = 0u;
          if (px < W && py < Y1) {
            const long long i = base + (long long)py * W + px;
            const Stencil s = load_stencil(S, i, px, py);
            const double duu = s.d.x + s.wr.x + s.wl.x + s.wd.x + s.wu.x;
            const double dvv = s.d.y + s.wr.y + s.wl.y + s.wd.y + s.wu.y;
            const float2 fwh = make_float2((float)s.wr.x, (float)s.wr.y), fwv = make_float2((float)s.wd.x, (float)s.wd.y);
            Df[i] = make_float2((float)s.d.x, (float)s.d.y);
            a12f[i] = (float)s.a12;
            WHf[i] = fwh;
            WVf[i] = fwv;
            const uint2 wp = make_uint2(ic_pack(fwh.x, fwh.y), ic_pack(fwv.x, fwv.y));
            const double2 rb = __ldg(&S.rhs[i]);
            v = make_float2((float)rb.x, (float)rb.y);
            x[i] = make_double2(0.0, 0.0);
            r[i] = v;
            pold[i] = make_float2(0.f, 0.f);
            y[i] = make_float2(0.f, 0.f);
            c0 = make_float4((float)duu, (float)s.a12, (float)dvv, __uint_as_float(wp.x));
            cw = (u == 7) ? 0u : wp.y;
            acc_bb += rb.x * rb.x + rb.y * rb.y;
          }
          sts64(sm.r + sl * 8, v); sts128(sm.c0 + sl * 16, c0); sts32(sm.cw + sl * 4, cw);
        }
        __syncwarp();
        ic_factor(sm, ty, tx);
        __syncwarp();
        ic_sweeps(sm, ty, tx);
        __syncwarp();
#pragma unroll 2
        for (int u = 0; u < 8; ++u) {
          const int py = py0 + u;
          if (px < W && py < Y1) {
            const long long i = base + (long long)py * W + px;
            const int sl = ic_slot(ty * 8 + u, tx);
            const float2 zz = lds64(sm.r + sl * 8), rv = r[i];
            z[i] = zz;
            C0[i] = lds128(sm.c0 + sl * 16);             // {i11, i12, i22, bf16x2 {wuh, wvh}}
            CW[i] = lds32(sm.cw + sl * 4);               // bf16x2 {wuv, wvv}, 0 on every eighth row
            acc_rz += (double)rv.x * (double)zz.x + (double)rv.y * (double)zz.y;
          }
        }
      }
      ic_block_sum2(acc_rz, acc_bb, sm_red);
      if (tid == 0) { part_b[(long long)b * G + cta] = acc_rz; part_c[(long long)b * G + cta] = acc_bb; }
    }
  }
  SYNC_REDUCE(part_b, part_c, 0)
  if (tid < B) {
    const int b = tid;
    double rz = s_ta[b], bb = s_tb[b];
    s_rz[b] = rz; s_bb[b] = bb; s_maxr2[b] = bb; s_alpha[b] = 0.0; s_beta[b] = 0.0; s_rzprev[b] = rz;
    if (!(bb > 0.0) || !(rz > 0.0)) {                    // zero right-hand side: x = 0 is the solution
      if (cta == C_LO(b)) { done_g[b] = 1; iters_g[b] = 0; relres_g[b] = 0.0; }
      s_state[b] = 1;
    }
  }
  REMAP()
  int n_active = s_nact;

#ifdef IC_TIMERS                                                 // tuning build: cycles per phase, printed by three CTAs
  long long tmA = 0, tmS1 = 0, tmB = 0, tmS2 = 0, tm0 = 0;
#define IC_TICK(acc) { const long long now = clock64(); acc += now - tm0; tm0 = now; }
#else
#define IC_TICK(acc)
#endif
  int k = 0;
  for (; k < P.maxit && n_active > 0; ++k) {
    OWN_RANGE()
#ifdef IC_TIMERS
    tm0 = clock64();
#endif
    // ---------------- phase A: p = z + beta p_old (on the fly), y += alpha_prev p_old, Ap = A p ----------------
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      const float beta = (float)s_beta[b], aprev = (float)s_alpha[b];
      const double aprev_d = s_alpha[b];
      const bool flush = s_flush[b] != 0;       // a reliable update happened: fold y + alpha p into the fp64 solution now
      double acc = 0.0, dummy = 0.0;
      for (long long v = ta; v < tb; v += IC_UA) {
        int px[IC_UA], py[IC_UA]; long long i[IC_UA]; bool ok[IC_UA];
        float2 zc[IC_UA], po[IC_UA], zl[IC_UA], pl[IC_UA], zr[IC_UA], pr[IC_UA], zu[IC_UA], pu[IC_UA], zd[IC_UA],
            pd[IC_UA], sd[IC_UA], swr[IC_UA], swd[IC_UA], swl[IC_UA], swu[IC_UA], yc[IC_UA];
        float sa12[IC_UA];
#pragma unroll
        for (int u = 0; u < IC_UA; ++u) {
          PIXEL_OF(v + u, px[u], py[u], i[u], ok[u])
          const long long ii = i[u];
          const long long jl = ii > 0 ? ii - 1 : 0, jr = ii + 1 < n_all ? ii + 1 : n_all - 1;
          const long long ju = ii >= W ? ii - W : 0, jd = ii + W < n_all ? ii + W : n_all - 1;
          zc[u] = z[ii]; po[u] = pold[ii];
          zl[u] = z[jl]; pl[u] = pold[jl]; zr[u] = z[jr]; pr[u] = pold[jr];
          if (BAND && py[u] == Y0 && bp.up_delta != 0) {          // the row above lives in the neighbour's block
            zu[u] = ld_sys_f2(band_shift(z + ju, bp.up_delta)); pu[u] = ld_sys_f2(band_shift(pold + ju, bp.up_delta));
          } else { zu[u] = z[ju]; pu[u] = pold[ju]; }
          if (BAND && py[u] + 1 == Y1 && bp.dn_delta != 0) {
            zd[u] = ld_sys_f2(band_shift(z + jd, bp.dn_delta)); pd[u] = ld_sys_f2(band_shift(pold + jd, bp.dn_delta));
          } else { zd[u] = z[jd]; pd[u] = pold[jd]; }
          sd[u] = Df[ii]; swr[u] = WHf[ii]; swd[u] = WVf[ii];
          swl[u] = WHf[jl];
          if (BAND && py[u] == Y0) {                              // the fp32 copy only covers the band: take the fp64 edge
            const double2 e = __ldg(&S.WV[ju]);
            swu[u] = make_float2((float)e.x, (float)e.y);
          } else swu[u] = WVf[ju];
          sa12[u] = a12f[ii];
          yc[u] = y[ii];
        }
#pragma unroll
        for (int u = 0; u < IC_UA; ++u) {
          if (!ok[u]) continue;
          const float2 c = make_float2(zc[u].x + beta * po[u].x, zc[u].y + beta * po[u].y);
          float2 nl = make_float2(zl[u].x + beta * pl[u].x, zl[u].y + beta * pl[u].y);
          float2 nr = make_float2(zr[u].x + beta * pr[u].x, zr[u].y + beta * pr[u].y);
          float2 nu = make_float2(zu[u].x + beta * pu[u].x, zu[u].y + beta * pu[u].y);
          float2 nd = make_float2(zd[u].x + beta * pd[u].x, zd[u].y + beta * pd[u].y);
          if (px[u] == 0) nl = c;
          if (px[u] + 1 >= W) nr = c;
          if (py[u] == 0) nu = c;
          if (py[u] + 1 >= H) nd = c;
          float au = sd[u].x * c.x + sa12[u] * c.y;
          float av = sa12[u] * c.x + sd[u].y * c.y;
          au += swr[u].x * (c.x - nr.x); av += swr[u].y * (c.y - nr.y);
          au += swl[u].x * (c.x - nl.x); av += swl[u].y * (c.y - nl.y);
          au += swd[u].x * (c.x - nd.x); av += swd[u].y * (c.y - nd.y);
          au += swu[u].x * (c.x - nu.x); av += swu[u].y * (c.y - nu.y);
          if (flush) {
            double2 xc = x[i[u]];
            x[i[u]] = make_double2(xc.x + ((double)yc[u].x + aprev_d * (double)po[u].x),
                                   xc.y + ((double)yc[u].y + aprev_d * (double)po[u].y));
            y[i[u]] = make_float2(0.f, 0.f);
          } else {
            y[i[u]] = make_float2(yc[u].x + aprev * po[u].x, yc[u].y + aprev * po[u].y);
          }
          pnew[i[u]] = c;
          Ap[i[u]] = make_float2(au, av);
          acc += (double)c.x * (double)au + (double)c.y * (double)av;
        }
      }
      ic_block_sum2(acc, dummy, sm_red);
      if (tid == 0) part_a[(long long)b * G + cta] = acc;
    }
    IC_TICK(tmA)
    SYNC_REDUCE(part_a, (const double *)nullptr, 0)
    IC_TICK(tmS1)
    if (tid < B && s_state[tid] == 0) {
      double pap = s_ta[tid];
      s_alpha[tid] = pap > 0.0 ? s_rz[tid] / pap : 0.0;      // 0 => breakdown, resolved by the reliable update below
      s_flush[tid] = 0;
    }
    __syncthreads();
    // ---------------- phase B: r -= alpha Ap, z = M^-1 r (tile-local IC sweeps), partial r.z and r.r ----------------
    for (int a = a_first; a <= a_last; ++a) {
      TILE_RANGE(a)
      const float alpha = (float)s_alpha[b];
      double acc_rz = 0.0, acc_rr = 0.0;
      for (long long t = ta + ty; t < tb; t += IC_NSTRIP) PHASE_B_TILE(t, alpha, true, acc_rz, acc_rr)
      ic_block_sum2(acc_rz, acc_rr, sm_red);
      if (tid == 0) { part_b[(long long)b * G + cta] = acc_rz; part_c[(long long)b * G + cta] = acc_rr; }
    }
    IC_TICK(tmB)
    SYNC_REDUCE(part_b, part_c, 0)
    IC_TICK(tmS2)
    int rel = 0;
    if (tid < B && s_state[tid] == 0) {
      const int b = tid;
      double rz = s_ta[b], rr = s_tb[b], alpha = s_alpha[b];
      int bad = !(alpha > 0.0) || !(rz > 0.0) || !(rr == rr);
      rel = bad || rr <= P.tol2 * s_bb[b] || rr < P.delta2 * s_maxr2[b] || k + 1 == P.maxit;
      if (rel) {
        s_state[b] = 3; s_bad[b] = bad; s_rzprev[b] = s_rz[b];
      } else {
        s_beta[b] = rz / s_rz[b];
        s_rz[b] = rz;
      }
    }
    if (__syncthreads_or(rel)) {
      // ---------------- reliable update: r = b - A (x + y + alpha p) in fp64, then z = M^-1 r ----------------
      // Only a few systems replace their residual in a given iteration, and every CTA waits at the barrier that follows:
      // with the regular dealing the 1/B of the CTAs that own such a system did ~1.5 iterations' worth of work while all
      // the others idled (ncu stall samples: that barrier alone was 10 % of the kernel).  For the update the strips of
      // exactly those systems are therefore RE-DEALT over the whole grid, and the regular dealing is restored afterwards
      // (every CTA derives both from the same shared state, so nothing has to be communicated).
      MAP_WHERE(s_state[b] == 3)
      {
        OWN_RANGE()
        for (int a = a_first; a <= a_last; ++a) {
          TILE_RANGE(a)
          const double alpha = s_alpha[b];
          double acc_rz = 0.0, acc_rr = 0.0, acc_dummy = 0.0;
          for (long long v = ta; v < tb; ++v) {
            int px, py; long long i; bool ok;
            PIXEL_OF(v, px, py, i, ok)
            if (!ok) continue;
            const double2 rt = ic_true_residual<BAND>(P, bp, i, px, py, pnew, alpha);
            r[i] = make_float2((float)rt.x, (float)rt.y);
            acc_rr += rt.x * rt.x + rt.y * rt.y;
          }
          __syncthreads();                       // r of the own strips is complete (same ownership in both passes)
          for (long long t = ta + ty; t < tb; t += IC_NSTRIP) PHASE_B_TILE(t, 0.f, false, acc_rz, acc_dummy)
          ic_block_sum2(acc_rz, acc_rr, sm_red);
          if (tid == 0) { part_d[(long long)b * G + cta] = acc_rz; part_e[(long long)b * G + cta] = acc_rr; }
        }
      }
      SYNC_REDUCE(part_d, part_e, 3)
      int fin = 0;
      if (tid < B && s_state[tid] == 3) {
        const int b = tid;
        double rz = s_ta[b], rr = s_tb[b];
        int conv = rr <= P.tol2 * s_bb[b];
        // fp32 variant: the iterated fp32 residual has reached its target (that is what brought us here, unless it was a
        // breakdown or the iteration cap); the fp64 residual just computed is reported, not enforced
        if (P.fp32_only && !s_bad[b] && k + 1 < P.maxit && rr == rr) conv = 1;
        // An fp32 breakdown (p.Ap <= 0, r.z <= 0: the iterated quantities have lost their meaning close to the fp32
        // floor) is not the end: the fp64 residual and z = M^-1 r just computed are sound, so CG RESTARTS from them
        // (beta = 0).  Given up after 8 restarts, or when three replacements in a row failed to halve the true residual
        // (the fp64 floor of an ill-conditioned system: attainable accuracy reached).
        const int fatal = !(rz > 0.0) || !(rr == rr) || k + 1 == P.maxit;
        if (s_bad[b]) s_restarts[b] += 1;
        s_stalls[b] = rr > 0.25 * s_lasttrue[b] ? s_stalls[b] + 1 : 0;
        s_lasttrue[b] = rr;
        if (conv || fatal || s_restarts[b] > 8 || s_stalls[b] >= 3) {
          s_state[b] = 2;
          fin = 1;
          if (cta == C_LO(b)) {
            // a solve that stagnates within a factor 4 of the target has reached what fp64 can deliver for this system
            // (seen on 17 x 30 Lorentzian levels: 1.2e-12 against 1e-12); its relres is reported as it is
            const int floor_ok = s_stalls[b] >= 3 && rr <= 16.0 * P.tol2 * s_bb[b];
            done_g[b] = (conv || floor_ok) ? 1 : (k + 1 == P.maxit && !s_bad[b] ? 3 : 2);
            iters_g[b] = k + 1;
            relres_g[b] = sqrt(rr / s_bb[b]);
          }
        } else {
          s_state[b] = 0;
          s_flush[b] = 1;                                  // x += y + alpha_k p_k rides on the next phase A
          s_maxr2[b] = rr;
          s_beta[b] = s_bad[b] ? 0.0 : rz / s_rzprev[b];   // restart after a breakdown: p = z
          s_rz[b] = rz;
        }
      }
      const int any_fin = __syncthreads_or(fin);
      if (any_fin) {
        // systems that just finished: fold the pending y + alpha p into x -- their pixels dealt over the whole grid like
        // the update above (a pass over one system by the few CTAs that own it would show up as skew at the next
        // barrier) -- then re-deal the strips of the systems that are left
        MAP_WHERE(s_state[b] == 2)
        {
          OWN_RANGE()
          for (int a = a_first; a <= a_last; ++a) {
            TILE_RANGE(a)
            const double alpha = s_alpha[b];
            for (long long v = ta; v < tb; ++v) {
              int px, py; long long i; bool ok;
              PIXEL_OF(v, px, py, i, ok)
              if (!ok) continue;
              double2 xc = x[i];
              float2 yc = y[i], pc = pnew[i];
              x[i] = make_double2(xc.x + ((double)yc.x + alpha * (double)pc.x), xc.y + ((double)yc.y + alpha * (double)pc.y));
            }
          }
        }
        __syncthreads();
        if (tid < B && s_state[tid] == 2) s_state[tid] = 1;
        REMAP()
        n_active = s_nact;
      } else {
        REMAP()                                // back to the regular dealing (same systems as at the top of the iteration)
      }
    }
    { float2 *t = pold; pold = pnew; pnew = t; }
  }
#ifdef IC_TIMERS
  if (tid == 0 && (cta == 0 || cta == G / 2 || cta == G - 2))
    printf("[ic timers] cta %d of %d, %d iterations: phase A %.1f, sync %.1f, phase B %.1f, sync %.1f kcycles/iter\n", cta, G, k,
           1e-3 * tmA / k, 1e-3 * tmS1 / k, 1e-3 * tmB / k, 1e-3 * tmS2 / k);
#endif
#undef IC_TICK
#undef REMAP
#undef MAP_WHERE
#undef C_LO
#undef C_HI
#undef OWN_RANGE
#undef REDUCE_ALL
#undef SYNC_REDUCE
#undef TILE_RANGE
#undef PIXEL_OF
#undef STRIP_X
#undef STRIP_Y
#undef PHASE_B_TILE
}

__global__ void __launch_bounds__(IC_THREADS, 512 / IC_THREADS) pcg_ic_kernel(MixParams P) {
  BandParams none;
  pcg_ic_body<IC_MODE_GRID>(P, none);
}

#ifndef IC_CG_BARRIER
__global__ void __launch_bounds__(IC_THREADS, 512 / IC_THREADS) pcg_ic_cluster_kernel(MixParams P) {
  BandParams none;
  pcg_ic_body<IC_MODE_CLUSTER>(P, none);
}

__global__ void __launch_bounds__(IC_THREADS, 512 / IC_THREADS) pcg_ic_band_kernel(MixParams P, BandParams bp) {
  pcg_ic_body<IC_MODE_BAND>(P, bp);
}
#endif

// two CTAs x (60 KB staged strips + 11 KB scalars + 1 KB reserved) = 144 KB -> the 164 KB shared-memory configuration, which
// leaves 92 KB of L1 for phase A's stencil reuse (measured: below ~90 KB of L1 the matvec phase loses 35 %)
constexpr int IC_CARVEOUT_PCT = 64;

int pcg_ic_grid(b200flow_ctx *ctx, int *grid_out) {
  if (ctx->grid_ic == 0) {
    // Function attributes are per device and set ONCE per process (mutex-guarded): changing an attribute of a kernel that is
    // running on another stream blocks until that kernel ends -- which a peer rank's persistent solver, spinning on this
    // rank's arrival (row-band emulation on one GPU), never does.
    static std::mutex mu;
    static bool attr_done[64] = {false};
    {
      std::lock_guard<std::mutex> lock(mu);
      if (!attr_done[ctx->device & 63]) {
        BF_CUDA(ctx, cudaFuncSetAttribute(pcg_ic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IC_SMEM));
        int carve = IC_CARVEOUT_PCT;
#ifdef B200FLOW_TUNING
        if (const char *cv = getenv("B200FLOW_IC_CARVEOUT")) carve = atoi(cv);
#endif
        BF_CUDA(ctx, cudaFuncSetAttribute(pcg_ic_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
#ifndef IC_CG_BARRIER
        BF_CUDA(ctx, cudaFuncSetAttribute(pcg_ic_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IC_SMEM));
        BF_CUDA(ctx, cudaFuncSetAttribute(pcg_ic_band_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
#endif
        attr_done[ctx->device & 63] = true;
      }
    }
    int nb = 0;
    BF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pcg_ic_kernel, IC_THREADS, IC_SMEM));
    if (nb < 1) return set_err(ctx, B200FLOW_ECUDA, "pcg_ic_kernel cannot be made resident");
    ctx->ic_ctas_per_sm = nb;
    ctx->grid_ic = nb * ctx->num_sms;
  }
  *grid_out = ctx->grid_ic;
  if (ctx->solver_ctas_per_sm > 0 && *grid_out > ctx->solver_ctas_per_sm * ctx->num_sms)
    *grid_out = ctx->solver_ctas_per_sm * ctx->num_sms;      // split context: the sibling groups' solvers are resident too
  return 0;
}

#ifndef IC_CG_BARRIER
// small levels: one cluster of `cs` CTAs per system (pcg_ic_body<IC_MODE_CLUSTER>); 0 = not available
static int ic_cluster_size(b200flow_ctx *ctx) {
  static std::mutex mu;
  static int cs_of[64] = {0};
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lock(mu);
  const int d = ctx->device & 63;
  if (!done[d]) {
    done[d] = true;
    int want = 16;
    if (const char *e = getenv("B200FLOW_CLUSTER_SIZE")) want = atoi(e);
    if (want >= 2 && cudaFuncSetAttribute(pcg_ic_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IC_SMEM) == cudaSuccess &&
        cudaFuncSetAttribute(pcg_ic_cluster_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, IC_CARVEOUT_PCT) == cudaSuccess) {
      if (want > 8) cudaFuncSetAttribute(pcg_ic_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      for (int cs = want; cs >= 2; cs >>= 1) {           // the largest cluster the device schedules with this footprint
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(IC_THREADS); cfg.dynamicSmemBytes = IC_SMEM;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, pcg_ic_cluster_kernel, &cfg) == cudaSuccess && n >= 1) { cs_of[d] = cs; break; }
      }
    }
    cudaGetLastError();
  }
  return cs_of[d];
}
#endif

int k_pcg_ic_launch(b200flow_ctx *ctx, MixParams P, int grid_max) {
  const LinSys &sys = P.sys;
  if (sys.B > MAXB)
    return set_err(ctx, B200FLOW_EINVAL, "batch of %d systems exceeds %d per solve; split the batch", sys.B, MAXB);
  P.tiles_x = (int)cdiv(sys.W, 32);      // strips of 8 rows x 32 columns: the unit that is dealt to the CTAs
  P.tiles_y = (int)cdiv(sys.H, 8);
  P.tiles_per_sys = P.tiles_x * P.tiles_y;
#ifndef IC_CG_BARRIER
  {
    // OFF by default (B200FLOW_CLUSTER_MAXPIX=24000 switches it on for levels up to 120 x 160).  Measured on B200, 16 systems,
    // us per iteration, grid barrier / clusters of 8 / 16 CTAs: 30x40 10.2 / 9.8 / 18.8, 60x80 12.1 / 12.2 / 20.9, 120x160
    // 16.6 / 22.7 / 29.8 (only eight 16-CTA clusters of this footprint are resident at a time: two waves).  The ~9 us floor
    // of a tiny iteration is the chain of dependent global-memory round trips inside the two phases, not the barrier.
    long long maxpix = 0;
    if (const char *e = getenv("B200FLOW_CLUSTER_MAXPIX")) maxpix = atoll(e);
    const int cs = (long long)sys.H * sys.W <= maxpix ? ic_cluster_size(ctx) : 0;
    if (cs >= 2 && (long long)sys.B * 5 * cs <= (long long)5 * sys.B * grid_max) {
      P.debug = 0;
      P.w.grid = cs;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3((unsigned)(sys.B * cs)); cfg.blockDim = dim3(IC_THREADS); cfg.dynamicSmemBytes = IC_SMEM;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      BF_CUDA(ctx, cudaLaunchKernelEx(&cfg, pcg_ic_cluster_kernel, P));
      return 0;
    }
  }
#endif
  const long long total = (long long)P.tiles_per_sys * sys.B;
  int G = grid_max;
  if ((long long)G > total) G = (int)total;
  if (G < 1) G = 1;
  P.w.grid = G;
  P.debug = 0;
#ifdef B200FLOW_TUNING                     // bit 0: skip the sweeps (z = staged r): what do they cost?
  if (const char *dbg = getenv("B200FLOW_IC_DEBUG")) P.debug = atoi(dbg);
#endif
  void *args[] = {&P};
  if (ctx->plain_solver_launch) {
    // split context (pipeline.cu): the groups' solver grids together never exceed what the device can hold and every
    // other kernel of the library terminates on its own, so all CTAs become resident without the cooperative attribute
    // (the kernel uses its own counter barrier, not cooperative_groups' grid.sync())
#ifdef IC_CG_BARRIER
    return set_err(ctx, B200FLOW_EINVAL, "IC_CG_BARRIER builds cannot run concurrent sub-batches");
#else
    pcg_ic_kernel<<<dim3(G), dim3(IC_THREADS), IC_SMEM, ctx->stream>>>(P);
    BF_CUDA(ctx, cudaPeekAtLastError());
#endif
  } else {
    BF_CUDA(ctx, cudaLaunchCooperativeKernel((void *)pcg_ic_kernel, dim3(G), dim3(IC_THREADS), args, IC_SMEM, ctx->stream));
  }
  return 0;
}

#ifndef IC_CG_BARRIER
// ---- row-band mode: launcher, and the exchange of the solution bands after a solve ------------------------------------
static BandParams make_band_params(const b200flow_ctx *ctx, int y0, int y1) {
  BandParams bp;
  memset(&bp, 0, sizeof bp);
  const b200flow_band &bd = ctx->band;
  bp.rank = bd.rank; bp.world = bd.world; bp.y0 = y0; bp.y1 = y1;
  bp.self = reinterpret_cast<BandSync *>(bd.base[bd.rank]);
  for (int r = 0; r < bd.world; ++r) bp.peer[r] = reinterpret_cast<BandSync *>(bd.base[r]);
  return bp;
}

void band_rows(int H, int rank, int world, int *y0, int *y1) {
  const int strips = (H + 7) / 8;                  // bands are whole 8-row strips: the IC sub-tiles never straddle a band
  const int s0 = (int)((long long)strips * rank / world), s1 = (int)((long long)strips * (rank + 1) / world);
  *y0 = s0 * 8;
  *y1 = s1 * 8 < H ? s1 * 8 : H;
}

int k_pcg_ic_band_launch(b200flow_ctx *ctx, MixParams P, int grid_max) {
  static_assert(sizeof(BandSync) <= B200FLOW_BAND_RESERVED, "BandSync must fit the reserved head of the arena block");
  const LinSys &sys = P.sys;
  if (sys.B != 1) return set_err(ctx, B200FLOW_EINVAL, "row-band mode solves one system at a time (B = %d)", sys.B);
  const b200flow_band &bd = ctx->band;
  int y0, y1;
  band_rows(sys.H, bd.rank, bd.world, &y0, &y1);
  BandParams bp = make_band_params(ctx, y0, y1);
  int ty0, ty1;
  if (bd.rank > 0) {                               // a neighbour exists only if its band is not empty
    band_rows(sys.H, bd.rank - 1, bd.world, &ty0, &ty1);
    if (ty1 > ty0) bp.up_delta = bd.base[bd.rank - 1] - bd.base[bd.rank];
  }
  if (bd.rank + 1 < bd.world) {
    band_rows(sys.H, bd.rank + 1, bd.world, &ty0, &ty1);
    if (ty1 > ty0) bp.dn_delta = bd.base[bd.rank + 1] - bd.base[bd.rank];
  }
  for (int r = 0; r < bd.world; ++r) {             // every band must be non-empty and start where the previous one ends
    band_rows(sys.H, r, bd.world, &ty0, &ty1);
    if (ty1 <= ty0) return set_err(ctx, B200FLOW_EINVAL, "row-band mode: %d rows do not split over %d ranks", sys.H, bd.world);
  }
  P.tiles_x = (int)cdiv(sys.W, 32);
  P.tiles_y = (int)cdiv(y1 - y0, 8);
  P.tiles_per_sys = P.tiles_x * P.tiles_y;
  int G = grid_max;
  if ((long long)G > P.tiles_per_sys) G = P.tiles_per_sys;
  if (G < 1) G = 1;
  P.w.grid = G;
  P.debug = 0;
  void *args[] = {&P, &bp};
  if (ctx->plain_solver_launch) {                  // in-process emulation of several ranks on one GPU (tests): all grids must be co-resident
    pcg_ic_band_kernel<<<dim3(G), dim3(IC_THREADS), IC_SMEM, ctx->stream>>>(P, bp);
    BF_CUDA(ctx, cudaPeekAtLastError());
  } else {
    BF_CUDA(ctx, cudaLaunchCooperativeKernel((void *)pcg_ic_band_kernel, dim3(G), dim3(IC_THREADS), args, IC_SMEM, ctx->stream));
  }
  return 0;
}

// After a band solve every rank holds rows [y0, y1) of x: each rank stores its rows into every peer's x (same offset in the
// peer's block) and the last CTA to finish runs one flag barrier, so that the kernel only completes once every peer's rows
// have arrived here.
__global__ void band_push_kernel(const double2 *__restrict__ x, long long first, long long count, BandParams bp,
                                 unsigned *ticket) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride) {
    const double2 v = x[first + i];
    for (int r = 0; r < bp.world; ++r) {
      if (r == bp.rank) continue;
      double2 *dst = const_cast<double2 *>(band_shift(x + first + i, reinterpret_cast<const char *>(bp.peer[r]) -
                                                                      reinterpret_cast<const char *>(bp.self)));
      asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(v.x), "d"(v.y) : "memory");
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;                                  // ready for the next push
    BandSync *self = bp.self;
    const unsigned long long seq = *(volatile unsigned long long *)&self->seq + 1ull;
    __threadfence_system();
    for (int r = 0; r < bp.world; ++r) st_relaxed_sys_u64(&bp.peer[r]->flag[bp.rank], seq);
    long long spins = 0;
    for (int r = 0; r < bp.world; ++r)
      while (ld_relaxed_sys_u64(&self->flag[r]) < seq && ++spins < BAND_SPIN_LIMIT) {}
    __threadfence_system();
    if (spins >= BAND_SPIN_LIMIT) self->error = seq;
    self->seq = seq;
    self->release = seq;
    __threadfence_system();
  }
}

int k_band_exchange_x(b200flow_ctx *ctx, double2 *x, int H, int W) {
  const b200flow_band &bd = ctx->band;
  int y0, y1;
  band_rows(H, bd.rank, bd.world, &y0, &y1);
  BandParams bp = make_band_params(ctx, y0, y1);
  const long long count = (long long)(y1 - y0) * W;
  int grid = (int)cdiv(count > 0 ? count : 1, 256 * 4);
  if (grid > 4 * ctx->num_sms) grid = 4 * ctx->num_sms;
  BF_LAUNCH(ctx, band_push_kernel, grid, 256, 0, x, (long long)y0 * W, count, bp, ctx->band.ticket);
  return 0;
}

// CUDA loads kernels lazily, at their first launch, and a load can need the context to be idle: a first launch issued
// while a peer rank's persistent solver is spinning on this rank would wait for a kernel that is waiting for it.  Row-band
// mode therefore loads its kernels up front.
int k_band_preload(b200flow_ctx *ctx) {
  cudaFuncAttributes a;
  BF_CUDA(ctx, cudaFuncGetAttributes(&a, pcg_ic_band_kernel));
  BF_CUDA(ctx, cudaFuncGetAttributes(&a, pcg_ic_kernel));
  BF_CUDA(ctx, cudaFuncGetAttributes(&a, band_push_kernel));
  int g = 0;
  BF_TRY(pcg_ic_grid(ctx, &g));
  return 0;
}

int k_band_error(b200flow_ctx *ctx, unsigned long long *err_host) {
  BandSync *self = reinterpret_cast<BandSync *>(ctx->band.base[ctx->band.rank]);
  BF_CUDA(ctx, cudaMemcpyAsync(err_host, &self->error, sizeof *err_host, cudaMemcpyDeviceToHost, ctx->stream));
  BF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
#endif

}  // namespace bf
