// DNA (4 states) x Gamma-4 specialisation of the likelihood kernels: one thread per pattern owns
// all 16 entries (4 categories x 4 states) of a CLV column.
//
// Why a second formulation next to the generic (pattern, category)-per-thread kernels of lik.cu:
// ncu on the generic root evaluation (profiles/r1a_*.txt) showed it issue-bound, not HBM-bound --
// 10.5e9 warp instructions per launch, a third of them the FP64 log() executed by warps in which
// only 8 of 32 lanes hold a site sum.  Here a warp covers 32 patterns: 16 coalesced 256-byte loads
// per operand, P-matrix rows read from shared memory as warp-wide broadcasts, one log() per lane.
//
// The CLV update can carry a fused evaluation (edge insertion: the partial N just computed is
// joined with the far-side CLV and the pruned subtree's CLV while still in registers), which
// removes one of the six CLV transfers per SPR neighbour.
#pragma once

namespace plg {

// Block size / residency chosen by the sweep in profiles/tune_d4.sh (B200, 64 taxa x 10k sites):
// 128 x {5,6,7} blocks per SM plateau at ~5.1 TB/s algorithmic; 4 blocks (128 registers) gives 4.6.
#ifndef PLG_D4_PREFETCH
#define PLG_D4_PREFETCH 0
#endif
#ifndef PLG_D4_THREADS
#define PLG_D4_THREADS 128
#endif
#ifndef PLG_D4_MINB
#define PLG_D4_MINB 6
#endif
constexpr int D4_THREADS = PLG_D4_THREADS;

__device__ __forceinline__ void d4_load16(const double *__restrict__ base, int64_t Ppad, double (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = __ldg(base + (int64_t)i * Ppad);
}

// stage P[4][16] (global stride 17 per category) or the 16x16 tip table into shared memory
__device__ __forceinline__ void d4_stage(double *sm, bool tip, int pm, const double *__restrict__ pmat,
                                         const double *__restrict__ tiptab) {
  if (tip) {
    const double *src = tiptab + (int64_t)pm * 256;
    for (int e = threadIdx.x; e < 256; e += D4_THREADS) sm[e] = src[e];
  } else {
    const double *src = pmat + (int64_t)pm * 68;
    for (int e = threadIdx.x; e < 64; e += D4_THREADS) sm[e] = src[(e >> 4) * 17 + (e & 15)];
  }
}

// v[r*4+k] (*)= sum_j P[r][k][j] x[r*4+j]; P rows are warp-uniform shared-memory broadcasts
template <bool MUL>
__device__ __forceinline__ void d4_apply(const double *sm, const double (&x)[16], double (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 p01 = *reinterpret_cast<const double2 *>(sm + r * 16 + k * 4);
      const double2 p23 = *reinterpret_cast<const double2 *>(sm + r * 16 + k * 4 + 2);
      double s = p01.x * x[r * 4];
      s = fma(p01.y, x[r * 4 + 1], s);
      s = fma(p23.x, x[r * 4 + 2], s);
      s = fma(p23.y, x[r * 4 + 3], s);
      if (MUL) v[r * 4 + k] *= s; else v[r * 4 + k] = s;
    }
}

template <bool MUL>
__device__ __forceinline__ void d4_tip(const double *sm, int code, double (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const double2 t = *reinterpret_cast<const double2 *>(sm + code * 16 + i);
    if (MUL) { v[i] *= t.x; v[i + 1] *= t.y; } else { v[i] = t.x; v[i + 1] = t.y; }
  }
}

// One operand of a product: tip -> table row; inner -> P x.  `pm < 0` = identity.
template <bool MUL>
__device__ __forceinline__ void d4_operand(const double *sm, int id, int pm, int n_rows, int64_t p, int64_t Ppad,
                                           const unsigned char *__restrict__ codes,
                                           const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                           const int *__restrict__ scl, double (&v)[16], int &sc) {
  if (id < n_rows) {
    const int code = codes[(int64_t)id * Ppad + p];
    if (pm >= 0) {
      d4_tip<MUL>(sm, code, v);
    } else {
      const unsigned int m = masks[code];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const double ind = (m >> (i & 3) & 1) ? 1.0 : 0.0;
        if (MUL) v[i] *= ind; else v[i] = ind;
      }
    }
  } else {
    double x[16];
    d4_load16(clv + (int64_t)(id - n_rows) * 16 * Ppad + p, Ppad, x);
    sc += scl[(int64_t)(id - n_rows) * Ppad + p];
    if (pm >= 0) {
      d4_apply<MUL>(sm, x, v);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (MUL) v[i] *= x[i]; else v[i] = x[i];
    }
  }
}

// w_p * (log site likelihood + scaler correction), summed over the block in a fixed order
__device__ __forceinline__ void d4_site_sum(const double (&v)[16], int sc, int64_t p,
                                            const double *__restrict__ weights,
                                            const ModelDev *__restrict__ model, double *warp_sums,
                                            double *__restrict__ out) {
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) term = fma(model->freqs[k], (v[k] + v[4 + k]) + (v[8 + k] + v[12 + k]), term);
  const double w = weights[p];
  double val = 0.0;
  if (w != 0.0) val = w * (log(term * 0.25) + sc * LOG_MINLH);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D4_THREADS / 32; i++) s += warp_sums[i];
    *out = s;
  }
}

__global__ void __launch_bounds__(D4_THREADS, PLG_D4_MINB)
clv4_kernel(const LikOpF *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
            const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
            const double *__restrict__ weights, const ModelDev *__restrict__ model,
            const double *__restrict__ pmat, const double *__restrict__ tiptab, double *__restrict__ clv,
            int *__restrict__ scl, double *__restrict__ partial, int pstride) {
  __shared__ __align__(16) double smA[256];
  __shared__ __align__(16) double smB[256];
  __shared__ __align__(16) double smE[64 + 256 + 256];
  __shared__ double warp_sums[D4_THREADS / 32];
  const int op_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - op_idx * blocks_per_op;
  const LikOpF op = ops[op_idx];
  const int64_t p = (int64_t)pb * D4_THREADS + threadIdx.x;
#if PLG_D4_PREFETCH
  // Later operands are fetched after dependent compute phases; start pulling their lines now.
  {
    auto pf = [&](int id) {
      if (id >= n_rows && (threadIdx.x & 15) == 0) {
        const double *src = clv + (int64_t)(id - n_rows) * 16 * Ppad + p;
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (int64_t)i * Ppad));
      }
    };
    if (op.b >= 0) pf(op.b);
    if (op.tree >= 0) { pf(op.e_b); pf(op.e_c); }
  }
#endif
  d4_stage(smA, op.a < n_rows, op.pa, pmat, tiptab);
  if (op.b >= 0) d4_stage(smB, op.b < n_rows, op.pb, pmat, tiptab);
  if (op.tree >= 0) {
    d4_stage(smE, false, op.e_pa, pmat, tiptab);
    d4_stage(smE + 64, op.e_b < n_rows, op.e_pb, pmat, tiptab);
    d4_stage(smE + 320, op.e_c < n_rows, op.e_pc, pmat, tiptab);
  }
  __syncthreads();
  double v[16];
  int sc = 0;
  d4_operand<false>(smA, op.a, op.pa, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  if (op.b >= 0) d4_operand<true>(smB, op.b, op.pb, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(v[i]));
  if (mx < MINLH) {  // every entry below 2^-256: rescale the column, count it
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] *= TWO256;
    sc += 1;
  }
  if (!op.pad) {  // pad = 1: nobody reads this CLV (dead-store elimination on the host)
    double *dst = clv + (int64_t)(op.dst - n_rows) * 16 * Ppad + p;
#pragma unroll
    for (int i = 0; i < 16; i++) dst[(int64_t)i * Ppad] = v[i];
    scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
  }
  if (op.tree >= 0) {
    double e[16];
    d4_apply<false>(smE, v, e);
    d4_operand<true>(smE + 64, op.e_b, op.e_pb, n_rows, p, Ppad, codes, masks, clv, scl, e, sc);
    d4_operand<true>(smE + 320, op.e_c, op.e_pc, n_rows, p, Ppad, codes, masks, clv, scl, e, sc);
    d4_site_sum(e, sc, p, weights, model, warp_sums, partial + (int64_t)op.tree * pstride + pb);
  }
}

__global__ void __launch_bounds__(D4_THREADS, PLG_D4_MINB)
eval4_kernel(const LikEval *__restrict__ evals, int blocks_per_eval, int n_rows, int64_t Ppad,
             const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
             const double *__restrict__ weights, const ModelDev *__restrict__ model,
             const double *__restrict__ pmat, const double *__restrict__ tiptab,
             const double *__restrict__ clv, const int *__restrict__ scl, double *__restrict__ partial, int pstride) {
  __shared__ __align__(16) double sm[3][256];
  __shared__ double warp_sums[D4_THREADS / 32];
  const int ev_idx = blockIdx.x / blocks_per_eval;
  const int pb = blockIdx.x - ev_idx * blocks_per_eval;
  const LikEval ev = evals[ev_idx];
  const int64_t p = (int64_t)pb * D4_THREADS + threadIdx.x;
  if (ev.pa >= 0) d4_stage(sm[0], ev.a < n_rows, ev.pa, pmat, tiptab);
  if (ev.pb >= 0) d4_stage(sm[1], ev.b < n_rows, ev.pb, pmat, tiptab);
  if (ev.c >= 0 && ev.pc >= 0) d4_stage(sm[2], ev.c < n_rows, ev.pc, pmat, tiptab);
  __syncthreads();
  double v[16];
  int sc = 0;
  d4_operand<false>(sm[0], ev.a, ev.pa, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  d4_operand<true>(sm[1], ev.b, ev.pb, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  if (ev.c >= 0) d4_operand<true>(sm[2], ev.c, ev.pc, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  d4_site_sum(v, sc, p, weights, model, warp_sums, partial + (int64_t)ev.tree * pstride + pb);
}

}  // namespace plg

// ---------------------------------------------------------------------------------------------
// Warp-per-category formulation of the same fused op: block = 4 warps = the 4 Gamma categories of
// 32 patterns, a thread owns the 4 states of ONE category.  Four times as many, four times lighter
// threads than clv4_kernel (thread = 16 entries, 80+ registers, 24 warps/SM): ncu showed that one
// latency-bound (issue slots 40 % active, long- and short-scoreboard stalls), not HBM-bound, so
// the lever is thread-level parallelism.  P rows come straight from L1 as warp-uniform loads (no
// shared-memory staging, no barrier before the first global load); only the cross-category
// maximum (rescaling) and the cross-category site sum go through shared memory.
namespace plg {

constexpr int R4_THREADS = 128;
constexpr int R4_PATTERNS = 32;

// v[k] (*)= sum_j P[r][k][j] x[j]  for this warp's category r; tips: table row slice
template <bool MUL>
__device__ __forceinline__ void r4_operand(int id, int pm, int r, int n_rows, int64_t p, int64_t Ppad,
                                           const unsigned char *__restrict__ codes,
                                           const unsigned int *__restrict__ masks, const double *__restrict__ pmat,
                                           const double *__restrict__ tiptab, const double *__restrict__ clv,
                                           double (&v)[4]) {
  if (id < n_rows) {
    const int code = codes[(int64_t)id * Ppad + p];
    if (pm >= 0) {
      const double *row = tiptab + (int64_t)pm * 256 + code * 16 + r * 4;
      const double2 t01 = __ldg(reinterpret_cast<const double2 *>(row));
      const double2 t23 = __ldg(reinterpret_cast<const double2 *>(row + 2));
      if (MUL) { v[0] *= t01.x; v[1] *= t01.y; v[2] *= t23.x; v[3] *= t23.y; }
      else { v[0] = t01.x; v[1] = t01.y; v[2] = t23.x; v[3] = t23.y; }
    } else {
      const unsigned int m = masks[code];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double ind = (m >> k & 1) ? 1.0 : 0.0;
        if (MUL) v[k] *= ind; else v[k] = ind;
      }
    }
  } else {
    const double *src = clv + ((int64_t)(id - n_rows) * 16 + r * 4) * Ppad + p;
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = __ldg(src + (int64_t)j * Ppad);
    if (pm >= 0) {
      const double *P = pmat + (int64_t)pm * 68 + r * 17;  // warp-uniform: broadcast from L1
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double s = __ldg(P + k * 4) * x[0];
        s = fma(__ldg(P + k * 4 + 1), x[1], s);
        s = fma(__ldg(P + k * 4 + 2), x[2], s);
        s = fma(__ldg(P + k * 4 + 3), x[3], s);
        if (MUL) v[k] *= s; else v[k] = s;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (MUL) v[k] *= x[k]; else v[k] = x[k];
    }
  }
}

__global__ void __launch_bounds__(R4_THREADS)
clv4r_kernel(const LikOpF *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
             const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
             const double *__restrict__ weights, const ModelDev *__restrict__ model,
             const double *__restrict__ pmat, const double *__restrict__ tiptab, double *__restrict__ clv,
             int *__restrict__ scl, double *__restrict__ partial, int pstride) {
  __shared__ double xr[4][R4_PATTERNS];  // cross-category exchange: maxima, then site terms
  const int op_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - op_idx * blocks_per_op;
  const LikOpF op = ops[op_idx];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int64_t p = (int64_t)pb * R4_PATTERNS + lane;
  double v[4];
  r4_operand<false>(op.a, op.pa, r, n_rows, p, Ppad, codes, masks, pmat, tiptab, clv, v);
  if (op.b >= 0) r4_operand<true>(op.b, op.pb, r, n_rows, p, Ppad, codes, masks, pmat, tiptab, clv, v);
  int sc = 0;
  if (op.a >= n_rows) sc += scl[(int64_t)(op.a - n_rows) * Ppad + p];
  if (op.b >= n_rows) sc += scl[(int64_t)(op.b - n_rows) * Ppad + p];
  xr[r][lane] = fmax(fmax(fabs(v[0]), fabs(v[1])), fmax(fabs(v[2]), fabs(v[3])));
  __syncthreads();
  const double mx = fmax(fmax(xr[0][lane], xr[1][lane]), fmax(xr[2][lane], xr[3][lane]));
  if (mx < MINLH) {  // all 16 entries of the column below 2^-256
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] *= TWO256;
    sc += 1;
  }
  if (!op.pad) {
    double *dst = clv + ((int64_t)(op.dst - n_rows) * 16 + r * 4) * Ppad + p;
#pragma unroll
    for (int k = 0; k < 4; k++) dst[(int64_t)k * Ppad] = v[k];
    if (r == 0) scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
  }
  if (op.tree >= 0) {
    double e[4];
    {
      const double *P = pmat + (int64_t)op.e_pa * 68 + r * 17;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double s = __ldg(P + k * 4) * v[0];
        s = fma(__ldg(P + k * 4 + 1), v[1], s);
        s = fma(__ldg(P + k * 4 + 2), v[2], s);
        s = fma(__ldg(P + k * 4 + 3), v[3], s);
        e[k] = s;
      }
    }
    r4_operand<true>(op.e_b, op.e_pb, r, n_rows, p, Ppad, codes, masks, pmat, tiptab, clv, e);
    r4_operand<true>(op.e_c, op.e_pc, r, n_rows, p, Ppad, codes, masks, pmat, tiptab, clv, e);
    if (op.e_b >= n_rows) sc += scl[(int64_t)(op.e_b - n_rows) * Ppad + p];
    if (op.e_c >= n_rows) sc += scl[(int64_t)(op.e_c - n_rows) * Ppad + p];
    double term = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) term = fma(model->freqs[k], e[k], term);
    __syncthreads();  // everyone has read the maxima
    xr[r][lane] = term;
    __syncthreads();
    if (r == 0) {
      const double w = weights[p];
      double val = 0.0;
      if (w != 0.0) val = w * (log(((xr[0][lane] + xr[1][lane]) + (xr[2][lane] + xr[3][lane])) * 0.25) + sc * LOG_MINLH);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
      if (lane == 0) partial[(int64_t)op.tree * pstride + pb] = val;
    }
  }
}

}  // namespace plg
