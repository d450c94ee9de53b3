// DNA x Gamma-4 "walk" kernel: a whole SPR search (one pruned subtree, every regraft edge of the
// rest of the tree) is walked depth-first by ONE warp for a slice of 32 patterns, so the partial
// that a search step hands to the next one never leaves the SM.
//
// The level-by-level kernels (visit4b_kernel) write every step's partial N to HBM and read it back
// one launch later: 2 x (C+4) bytes per continuing step on top of the operands, and the kernel sat
// on the L2 fabric at ~50 % of DRAM peak (profiles/r1b_visit4b_kernel_ncu.txt).  Here
//   * Tv (the pruned subtree's CLV) is pushed through its branch once per search and parked in
//     shared memory for every regraft edge of the search;
//   * the partial that continues into a child stays in REGISTERS; when both children continue, the
//     walk descends into the smaller subtree first and parks the other partial in a per-warp stack
//     (global memory, reused by every search the warp runs, <= log2(n) levels deep, so it lives in
//     L2 and is overwritten before it is ever written back);
//   * work items are ordered slice-major: the ~3(n-2) searches of one pattern slice run at the same
//     time on different warps, so the base tree's directional CLVs of that slice (3(n-2) x 4 KB)
//     are fetched from HBM once and then served by L2.
// What is left per neighbour is the far-side operand D (C+4 bytes, mostly an L2 hit) and the FP64
// arithmetic: the step is FP64-pipe / L2 bound instead of HBM bound.
//
// One persistent grid; warps draw (slice, search) items from an atomic counter, longest searches
// first.  Warps never synchronise with each other (only __syncwarp), matrices are staged per warp.
#pragma once

namespace plg {

constexpr int WK_WARPS = 4;
constexpr int WK_THREADS = 32 * WK_WARPS;
#ifndef PLG_WK_MINB
#define PLG_WK_MINB 3
#endif
constexpr int WK_SMEM_DOUBLES_PER_WARP = 2 * 5 * 64 + 3 * 16 * 32;  // matrices double-buffered + 3 parked vectors
constexpr int WK_SMEM_BYTES = WK_WARPS * WK_SMEM_DOUBLES_PER_WARP * 8;

// stage one 4 x (4x4) P-matrix (global stride 17 per category) into this warp's shared memory
__device__ __forceinline__ void wk_stage_mat(double *sm, int pm, const double *__restrict__ pmat, int lane) {
  const double *src = pmat + (int64_t)pm * 68;
  sm[lane] = __ldg(src + (lane >> 4) * 17 + (lane & 15));
  sm[lane + 32] = __ldg(src + ((lane + 32) >> 4) * 17 + (lane & 15));
}
// the same without a register round trip (cp.async, 8 bytes per copy): issued one visit ahead
__device__ __forceinline__ void wk_stage_mat_async(double *sm, int pm, const double *__restrict__ pmat, int lane) {
  const double *src = pmat + (int64_t)pm * 68;
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(sm + lane);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src + (lane >> 4) * 17 + (lane & 15)) : "memory");
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d + 256), "l"(src + ((lane + 32) >> 4) * 17 + (lane & 15)) : "memory");
}
// One visit descriptor = 12 ints: lane l < 12 keeps int l, fields are broadcast at the point of use
__device__ __forceinline__ int wk_desc_load(const LikWalkVisit *__restrict__ visits, int vi, int lane) {
  return lane < 12 ? __ldg(reinterpret_cast<const int *>(visits + vi) + lane) : 0;
}

// y = P x for an operand that is either a tip row (table lookup, the table already holds P e_code)
// or an inner CLV in global memory: TWO branch lengths of the same operand at once (the full
// branch toward the other neighbour's partial, half the branch for this neighbour's own far side),
// so that every far-side CLV is read exactly once per search step.
__device__ __forceinline__ void wk_operand2(const double *mat_t, const double *mat_h, int id, int pm_t, int pm_h,
                                            int n_rows, int code, int64_t p, int64_t Ppad,
                                            const double *__restrict__ tiptab, const double *__restrict__ clv,
                                            double (&yt)[16], double (&yh)[16]) {
  if (id < n_rows) {
    const double *rt = tiptab + (int64_t)pm_t * 256 + code * 16, *rh = tiptab + (int64_t)pm_h * 256 + code * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const double2 a = __ldg(reinterpret_cast<const double2 *>(rt + i));
      const double2 b = __ldg(reinterpret_cast<const double2 *>(rh + i));
      yt[i] = a.x; yt[i + 1] = a.y; yh[i] = b.x; yh[i + 1] = b.y;
    }
  } else {
    double x[16];
    d4_load16(clv + (int64_t)(id - n_rows) * 16 * Ppad + p, Ppad, x);
    v4_apply_reg<false>(mat_t, x, yt);
    v4_apply_reg<false>(mat_h, x, yh);
  }
}

template <bool MUL>
__device__ __forceinline__ void wk_operand(const double *mat, int id, int pm_tab, int n_rows, int code, int64_t p,
                                           int64_t Ppad, const double *__restrict__ tiptab,
                                           const double *__restrict__ clv, double (&v)[16]) {
  if (id < n_rows) {
    const double *row = tiptab + (int64_t)pm_tab * 256 + code * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(row + i));
      if (MUL) { v[i] *= t.x; v[i + 1] *= t.y; } else { v[i] = t.x; v[i + 1] = t.y; }
    }
  } else {
    double x[16];
    d4_load16(clv + (int64_t)(id - n_rows) * 16 * Ppad + p, Ppad, x);
    v4_apply_reg<MUL>(mat, x, v);
  }
}

// all 16 entries (non-negative) below 2^-256 <=> the largest high word is below that of 2^-256
// (whose low word is zero): multiply by 2^256 and count (libpll's per-site scaling)
__device__ __forceinline__ int wk_rescale(double (&n)[16]) {
  int hi = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) hi = max(hi, __double2hiint(n[i]));
  if (hi >= 0x2FF00000) return 0;  // high word of 2^-256 (biased exponent 1023 - 256 = 0x2FF)
#pragma unroll
  for (int i = 0; i < 16; i++) n[i] *= TWO256;
  return 1;
}

#ifndef PLG_WK_PIPE
#define PLG_WK_PIPE 1  // 0: visit descriptors, matrices and codes fetched at the top of each visit (first version)
#endif
#if PLG_WK_PIPE
__global__ void __launch_bounds__(WK_THREADS, PLG_WK_MINB)
walk4_kernel(const LikWalkGroup *__restrict__ groups, int n_groups, const LikWalkVisit *__restrict__ visits,
             int64_t n_items, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
             const double *__restrict__ weights, const ModelDev *__restrict__ model,
             const double *__restrict__ pmat, const double *__restrict__ tiptab, const double *__restrict__ clv,
             const int *__restrict__ scl, double *__restrict__ stack_clv, int *__restrict__ stack_scl,
             int stack_levels, double *__restrict__ partial, int pstride, unsigned long long *counter) {
  extern __shared__ __align__(16) double wk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *mats0 = wk_smem + warp * WK_SMEM_DOUBLES_PER_WARP;                                     // [2][5][64]
  double(*park)[16][32] = reinterpret_cast<double(*)[16][32]>(mats0 + 2 * 5 * 64);  // [0] a = P_in X, [1] tvv = P_v Tv, [2] h1 = P_h1 D1
  const int64_t gw = (int64_t)blockIdx.x * WK_WARPS + warp;
  double *stk = stack_clv + gw * stack_levels * 16 * 32 + lane;
  int *stks = stack_scl + gw * stack_levels * 32 + lane;
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((int64_t)item >= n_items) break;
    const int slice = (int)(item / (unsigned)n_groups);
    const LikWalkGroup g = groups[item - (unsigned long long)slice * n_groups];
    const int64_t p = (int64_t)slice * 32 + lane;
    const unsigned char *cp = codes + p;
    const int *sp = scl + p;
    const double w = weights[p];
    int st = 0;
    // software pipeline over the visits of this search: descriptor two visits ahead, matrices
    // (cp.async, double-buffered), tip codes and scaling counters one visit ahead
    const int v_end = g.first + g.count;
    int dcur = wk_desc_load(visits, g.first, lane);
    int dnxt = g.count > 1 ? wk_desc_load(visits, g.first + 1, lane) : 0;
    __syncwarp();
    {
      double *m = mats0;  // buffer 0: first visit (slot 0 doubles as the pruned subtree's matrix first)
      wk_stage_mat_async(m + 1 * 64, __shfl_sync(0xffffffffu, dcur, 3), pmat, lane);
      wk_stage_mat_async(m + 2 * 64, __shfl_sync(0xffffffffu, dcur, 4), pmat, lane);
      wk_stage_mat_async(m + 3 * 64, __shfl_sync(0xffffffffu, dcur, 6), pmat, lane);
      wk_stage_mat_async(m + 4 * 64, __shfl_sync(0xffffffffu, dcur, 7), pmat, lane);
    }
    int c1, c2, s1, s2;
    {
      const int d1 = __shfl_sync(0xffffffffu, dcur, 2), d2 = __shfl_sync(0xffffffffu, dcur, 5);
      c1 = d1 < n_rows ? cp[(int64_t)d1 * Ppad] : 0;
      c2 = d2 < n_rows ? cp[(int64_t)d2 * Ppad] : 0;
      s1 = d1 < n_rows ? 0 : sp[(int64_t)(d1 - n_rows) * Ppad];
      s2 = d2 < n_rows ? 0 : sp[(int64_t)(d2 - n_rows) * Ppad];
    }
    {  // the pruned subtree seen through its branch, once per search
      if (g.tv >= n_rows) { wk_stage_mat(mats0, g.p_v, pmat, lane); st = sp[(int64_t)(g.tv - n_rows) * Ppad]; }
      const int ct = g.tv < n_rows ? cp[(int64_t)g.tv * Ppad] : 0;
      __syncwarp();
      double t[16];
      wk_operand<false>(mats0, g.tv, g.p_v, n_rows, ct, p, Ppad, tiptab, clv, t);
#pragma unroll
      for (int i = 0; i < 16; i++) park[1][i][lane] = t[i];
      __syncwarp();
      wk_stage_mat_async(mats0, __shfl_sync(0xffffffffu, dcur, 1), pmat, lane);
    }
    double nk[16];  // the partial that continues into the next visit
    int sck = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) nk[i] = 0.0;
    for (int vi = g.first; vi < v_end; vi++) {
      double *mats = mats0 + ((vi - g.first) & 1) * 5 * 64;
      const int x_in = __shfl_sync(0xffffffffu, dcur, 0), p_in = __shfl_sync(0xffffffffu, dcur, 1);
      const int d1 = __shfl_sync(0xffffffffu, dcur, 2), p_t1 = __shfl_sync(0xffffffffu, dcur, 3);
      const int p_h1 = __shfl_sync(0xffffffffu, dcur, 4), d2 = __shfl_sync(0xffffffffu, dcur, 5);
      const int p_t2 = __shfl_sync(0xffffffffu, dcur, 6), p_h2 = __shfl_sync(0xffffffffu, dcur, 7);
      const int tree1 = __shfl_sync(0xffffffffu, dcur, 8), tree2 = __shfl_sync(0xffffffffu, dcur, 9);
      const int keep = __shfl_sync(0xffffffffu, dcur, 10), push = __shfl_sync(0xffffffffu, dcur, 11);
      // this visit's matrices have landed; start the next visit's copies into the other buffer
      asm volatile("cp.async.wait_all;\n" ::: "memory");
      __syncwarp();
      int dnn = 0, nc1 = 0, nc2 = 0, ns1 = 0, ns2 = 0;
      if (vi + 1 < v_end) {
        if (vi + 2 < v_end) dnn = wk_desc_load(visits, vi + 2, lane);
        double *mn = mats0 + ((vi + 1 - g.first) & 1) * 5 * 64;
        wk_stage_mat_async(mn + 0 * 64, __shfl_sync(0xffffffffu, dnxt, 1), pmat, lane);
        wk_stage_mat_async(mn + 1 * 64, __shfl_sync(0xffffffffu, dnxt, 3), pmat, lane);
        wk_stage_mat_async(mn + 2 * 64, __shfl_sync(0xffffffffu, dnxt, 4), pmat, lane);
        wk_stage_mat_async(mn + 3 * 64, __shfl_sync(0xffffffffu, dnxt, 6), pmat, lane);
        wk_stage_mat_async(mn + 4 * 64, __shfl_sync(0xffffffffu, dnxt, 7), pmat, lane);
        const int e1 = __shfl_sync(0xffffffffu, dnxt, 2), e2 = __shfl_sync(0xffffffffu, dnxt, 5);
        nc1 = e1 < n_rows ? cp[(int64_t)e1 * Ppad] : 0;
        nc2 = e2 < n_rows ? cp[(int64_t)e2 * Ppad] : 0;
        ns1 = e1 < n_rows ? 0 : sp[(int64_t)(e1 - n_rows) * Ppad];
        ns2 = e2 < n_rows ? 0 : sp[(int64_t)(e2 - n_rows) * Ppad];
      }
      int sx;
      {  // a = P_in X: X from global (first step of a side), from the stack, or still in registers
        double a[16];
        if (x_in >= 0) {
          const int cx = x_in < n_rows ? codes[(int64_t)x_in * Ppad + p] : 0;
          sx = x_in < n_rows ? 0 : scl[(int64_t)(x_in - n_rows) * Ppad + p];
          wk_operand<false>(mats, x_in, p_in, n_rows, cx, p, Ppad, tiptab, clv, a);
        } else {
          if (x_in <= -2) {
            const int lvl = -2 - x_in;
#pragma unroll
            for (int i = 0; i < 16; i++) nk[i] = stk[(lvl * 16 + i) * 32];
            sck = stks[lvl * 32];
          }
          sx = sck;
          v4_apply_reg<false>(mats, nk, a);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) park[0][i][lane] = a[i];
      }
      // D1 once: the partial toward d2 (N2 = a o P_t1 D1) and the far side of edge 1 (h1 = P_h1 D1)
      double n2[16];
      {
        double h1[16];
        wk_operand2(mats + 1 * 64, mats + 2 * 64, d1, p_t1, p_h1, n_rows, c1, p, Ppad, tiptab, clv, n2, h1);
#pragma unroll
        for (int i = 0; i < 16; i++) { n2[i] *= park[0][i][lane]; park[2][i][lane] = h1[i]; }
      }
      int sc2 = sx + s1;
      sc2 += wk_rescale(n2);
      // D2 once: N1 = a o P_t2 D2 and h2 = P_h2 D2
      double n1[16], h2[16];
      wk_operand2(mats + 3 * 64, mats + 4 * 64, d2, p_t2, p_h2, n_rows, c2, p, Ppad, tiptab, clv, n1, h2);
#pragma unroll
      for (int i = 0; i < 16; i++) n1[i] *= park[0][i][lane];
      int sc1 = sx + s2;
      sc1 += wk_rescale(n1);
      // neighbour 2 (insertion into edge x--c2): (P_h2 N2) o h2 o tvv
      if (tree2 >= 0) {
        double e[16];
        v4_apply_reg<false>(mats + 4 * 64, n2, e);
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] *= h2[i] * park[1][i][lane];
        double val = v4_site(e, sc2 + s2 + st, w, model);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
        if (lane == 0) partial[(int64_t)tree2 * pstride + slice] = val;
      }
      // neighbour 1 (insertion into edge x--c1): (P_h1 N1) o h1 o tvv
      if (tree1 >= 0) {
        double e[16];
        v4_apply_reg<false>(mats + 2 * 64, n1, e);
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] *= park[2][i][lane] * park[1][i][lane];
        double val = v4_site(e, sc1 + s1 + st, w, model);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
        if (lane == 0) partial[(int64_t)tree1 * pstride + slice] = val;
      }
      if (push >= 0) {  // the partial that does not continue right away waits on the stack
        if (keep == 1) {
#pragma unroll
          for (int i = 0; i < 16; i++) stk[(push * 16 + i) * 32] = n2[i];
          stks[push * 32] = sc2;
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) stk[(push * 16 + i) * 32] = n1[i];
          stks[push * 32] = sc1;
        }
      }
      if (keep == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) nk[i] = n1[i];
        sck = sc1;
      } else if (keep == 2) {
#pragma unroll
        for (int i = 0; i < 16; i++) nk[i] = n2[i];
        sck = sc2;
      }
      dcur = dnxt; dnxt = dnn;
      c1 = nc1; c2 = nc2; s1 = ns1; s2 = ns2;
    }
  }
}

#else
__global__ void __launch_bounds__(WK_THREADS, PLG_WK_MINB)
walk4_kernel(const LikWalkGroup *__restrict__ groups, int n_groups, const LikWalkVisit *__restrict__ visits,
             int64_t n_items, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
             const double *__restrict__ weights, const ModelDev *__restrict__ model,
             const double *__restrict__ pmat, const double *__restrict__ tiptab, const double *__restrict__ clv,
             const int *__restrict__ scl, double *__restrict__ stack_clv, int *__restrict__ stack_scl,
             int stack_levels, double *__restrict__ partial, int pstride, unsigned long long *counter) {
  extern __shared__ __align__(16) double wk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *mats = wk_smem + warp * WK_SMEM_DOUBLES_PER_WARP;  // [5][64] (+ an unused second buffer)                                      // [5][64]
  double(*park)[16][32] = reinterpret_cast<double(*)[16][32]>(mats + 2 * 5 * 64);  // [0] a = P_in X, [1] tvv = P_v Tv, [2] h1 = P_h1 D1
  const int64_t gw = (int64_t)blockIdx.x * WK_WARPS + warp;
  double *stk = stack_clv + gw * stack_levels * 16 * 32 + lane;
  int *stks = stack_scl + gw * stack_levels * 32 + lane;
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((int64_t)item >= n_items) break;
    const int slice = (int)(item / (unsigned)n_groups);
    const LikWalkGroup g = groups[item - (unsigned long long)slice * n_groups];
    const int64_t p = (int64_t)slice * 32 + lane;
    const double w = weights[p];
    int st = 0;
    __syncwarp();
    {  // the pruned subtree seen through its branch, once per search
      if (g.tv >= n_rows) { wk_stage_mat(mats, g.p_v, pmat, lane); st = scl[(int64_t)(g.tv - n_rows) * Ppad + p]; }
      const int ct = g.tv < n_rows ? codes[(int64_t)g.tv * Ppad + p] : 0;
      __syncwarp();
      double t[16];
      wk_operand<false>(mats, g.tv, g.p_v, n_rows, ct, p, Ppad, tiptab, clv, t);
#pragma unroll
      for (int i = 0; i < 16; i++) park[1][i][lane] = t[i];
    }
    double nk[16];  // the partial that continues into the next visit
    int sck = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) nk[i] = 0.0;
    for (int vi = g.first; vi < g.first + g.count; vi++) {
      const int4 *vp = reinterpret_cast<const int4 *>(visits + vi);
      const int4 q0 = __ldg(vp), q1 = __ldg(vp + 1), q2 = __ldg(vp + 2);
      const int x_in = q0.x, p_in = q0.y, d1 = q0.z, p_t1 = q0.w;
      const int p_h1 = q1.x, d2 = q1.y, p_t2 = q1.z, p_h2 = q1.w;
      const int tree1 = q2.x, tree2 = q2.y, keep = q2.z, push = q2.w;
      __syncwarp();  // everyone is done with the previous step's matrices
      wk_stage_mat(mats + 0 * 64, p_in, pmat, lane);
      wk_stage_mat(mats + 1 * 64, p_t1, pmat, lane);
      wk_stage_mat(mats + 2 * 64, p_h1, pmat, lane);
      wk_stage_mat(mats + 3 * 64, p_t2, pmat, lane);
      wk_stage_mat(mats + 4 * 64, p_h2, pmat, lane);
      const int c1 = d1 < n_rows ? codes[(int64_t)d1 * Ppad + p] : 0;
      const int c2 = d2 < n_rows ? codes[(int64_t)d2 * Ppad + p] : 0;
      const int s1 = d1 < n_rows ? 0 : scl[(int64_t)(d1 - n_rows) * Ppad + p];
      const int s2 = d2 < n_rows ? 0 : scl[(int64_t)(d2 - n_rows) * Ppad + p];
      int sx;
      __syncwarp();
      {  // a = P_in X: X from global (first step of a side), from the stack, or still in registers
        double a[16];
        if (x_in >= 0) {
          const int cx = x_in < n_rows ? codes[(int64_t)x_in * Ppad + p] : 0;
          sx = x_in < n_rows ? 0 : scl[(int64_t)(x_in - n_rows) * Ppad + p];
          wk_operand<false>(mats, x_in, p_in, n_rows, cx, p, Ppad, tiptab, clv, a);
        } else {
          if (x_in <= -2) {
            const int lvl = -2 - x_in;
#pragma unroll
            for (int i = 0; i < 16; i++) nk[i] = stk[(lvl * 16 + i) * 32];
            sck = stks[lvl * 32];
          }
          sx = sck;
          v4_apply_reg<false>(mats, nk, a);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) park[0][i][lane] = a[i];
      }
      // D1 once: the partial toward d2 (N2 = a o P_t1 D1) and the far side of edge 1 (h1 = P_h1 D1)
      double n2[16];
      {
        double h1[16];
        wk_operand2(mats + 1 * 64, mats + 2 * 64, d1, p_t1, p_h1, n_rows, c1, p, Ppad, tiptab, clv, n2, h1);
#pragma unroll
        for (int i = 0; i < 16; i++) { n2[i] *= park[0][i][lane]; park[2][i][lane] = h1[i]; }
      }
      int sc2 = sx + s1;
      sc2 += wk_rescale(n2);
      // D2 once: N1 = a o P_t2 D2 and h2 = P_h2 D2
      double n1[16], h2[16];
      wk_operand2(mats + 3 * 64, mats + 4 * 64, d2, p_t2, p_h2, n_rows, c2, p, Ppad, tiptab, clv, n1, h2);
#pragma unroll
      for (int i = 0; i < 16; i++) n1[i] *= park[0][i][lane];
      int sc1 = sx + s2;
      sc1 += wk_rescale(n1);
      // neighbour 2 (insertion into edge x--c2): (P_h2 N2) o h2 o tvv
      if (tree2 >= 0) {
        double e[16];
        v4_apply_reg<false>(mats + 4 * 64, n2, e);
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] *= h2[i] * park[1][i][lane];
        double val = v4_site(e, sc2 + s2 + st, w, model);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
        if (lane == 0) partial[(int64_t)tree2 * pstride + slice] = val;
      }
      // neighbour 1 (insertion into edge x--c1): (P_h1 N1) o h1 o tvv
      if (tree1 >= 0) {
        double e[16];
        v4_apply_reg<false>(mats + 2 * 64, n1, e);
#pragma unroll
        for (int i = 0; i < 16; i++) e[i] *= park[2][i][lane] * park[1][i][lane];
        double val = v4_site(e, sc1 + s1 + st, w, model);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
        if (lane == 0) partial[(int64_t)tree1 * pstride + slice] = val;
      }
      if (push >= 0) {  // the partial that does not continue right away waits on the stack
        if (keep == 1) {
#pragma unroll
          for (int i = 0; i < 16; i++) stk[(push * 16 + i) * 32] = n2[i];
          stks[push * 32] = sc2;
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) stk[(push * 16 + i) * 32] = n1[i];
          stks[push * 32] = sc1;
        }
      }
      if (keep == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) nk[i] = n1[i];
        sck = sc1;
      } else if (keep == 2) {
#pragma unroll
        for (int i = 0; i < 16; i++) nk[i] = n2[i];
        sck = sc2;
      }
    }
  }
}

#endif

}  // namespace plg
