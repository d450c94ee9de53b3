// EXPERIMENT (not compiled into the library; measured slower): to try it again, copy next to
// lik_walkm.cuh, include it after that header in lik.cu and launch walk4m2_kernel in place of walk4m_kernel
// with n_sm * PLG_W2_MINB blocks of W2_THREADS threads, (W2_WARPS / 2) * W2_PAIR_BYTES of dynamic shared
// memory and one stack region per PAIR.
//
// DNA x Gamma-4 SPR search walk, the four rate categories of a slice split over a WARP PAIR.
//
// Same walk, same visit records, same layout V as walk4m_kernel (lik_walkm.cuh), but a warp carries two
// of the four categories of its 32 patterns: 8 instead of 16 doubles per vector and lane, 32 instead of
// 64 DMMAs per visit on each warp's dependent chain.  What couples the categories of a pattern is small
// and goes through shared memory with a 64-thread named barrier: the rescaling decision (largest high
// word per pattern group: 4 ints per lane, twice per visit) and the site sums (8 doubles per lane, once
// per visit; the even warp of the pair finishes and emits both neighbours).  Why: ncu on walk4m_kernel
// (profiles/r2q_walk4m_ncu.txt) shows the FP64 pipe 63 % busy with 12 warps per SM at 168 registers, and
// removing a quarter of its tensor instructions made it slower -- it is short of
// independent work per scheduler, not of pipe cycles.  Half the state per warp buys twice the warps.
#pragma once

namespace plg {

#ifndef PLG_W2_WARPS
#define PLG_W2_WARPS 8
#endif
#ifndef PLG_W2_MINB
#define PLG_W2_MINB 3
#endif
constexpr int W2_WARPS = PLG_W2_WARPS;  // per block (even): pairs (0,1), (2,3), ...
constexpr int W2_THREADS = 32 * W2_WARPS;
constexpr int W2_V = 8;                 // doubles per vector and lane: [category rl of this warp][group j] = v[rl * 4 + j]
// shared memory per pair (bytes): parked vectors 2 warps x (tvv, h1) x 8 x 32 doubles; high words 2 buffers x 2 warps x
// 4 x 32 ints; site sums 8 x 32 doubles; item slot
constexpr int W2_PARK_DOUBLES = 2 * W2_V * 32;  // per warp
constexpr int W2_PAIR_BYTES = 2 * W2_PARK_DOUBLES * 8 + 2 * 2 * 4 * 32 * 4 + 8 * 32 * 8 + 16;

__device__ __forceinline__ void w2_bar(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

__device__ __forceinline__ void w2_pair(const double (&x)[W2_V], const double (&bf)[2], double (&t)[W2_V], double (&h)[W2_V]) {
#pragma unroll
  for (int rl = 0; rl < 2; rl++)
#pragma unroll
    for (int j = 0; j < 4; j++) wm_dmma(t[rl * 4 + j], h[rl * 4 + j], x[rl * 4 + j], bf[rl]);
}

__device__ __forceinline__ void w2_frag(const double *__restrict__ pfrag, int pm, int lane, int rbase, double (&bf)[2]) {
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];\n" : "=d"(bf[0]), "=d"(bf[1]) : "l"(pfrag + ((int64_t)pm * 32 + lane) * 4 + rbase));
}

__device__ __forceinline__ void w2_load_raw(int id, int n_rows, int64_t Ppad, int64_t p0, int lane, int rbase, int out_off,
                                            const unsigned char *__restrict__ codes, const double *__restrict__ clv,
                                            const int *__restrict__ scl, double (&x)[W2_V], int &s) {
  const int g2 = (lane >> 2) * 2, q = lane & 3;
  if (id < n_rows) {
    const unsigned char *c = codes + (int64_t)id * Ppad + p0 + g2;
    unsigned int m[4];
    const unsigned int lo = *reinterpret_cast<const unsigned short *>(c);
    const unsigned int hi = *reinterpret_cast<const unsigned short *>(c + 16);
    m[0] = lo & 255u; m[1] = lo >> 8;
    m[2] = hi & 255u; m[3] = hi >> 8;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned int mask = m[j] ? m[j] : 15u;
      const double ind = (mask >> q & 1u) ? 1.0 : 0.0;
      x[j] = ind; x[4 + j] = ind;
    }
    s = 0;
  } else {
    const double *b = clv + (int64_t)(id - n_rows) * 16 * Ppad + p0 + g2;
#pragma unroll
    for (int rl = 0; rl < 2; rl++) {
      const int r = rbase + rl;
      const double2 u = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad));
      const double2 v = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad + 16));
      x[rl * 4] = u.x; x[rl * 4 + 1] = u.y; x[rl * 4 + 2] = v.x; x[rl * 4 + 3] = v.y;
    }
    s = __ldg(scl + (int64_t)(id - n_rows) * Ppad + p0 + out_off);
  }
}

// all 16 entries of a pattern below 2^-256 -> times 2^256, counted (libpll's rule).  The pattern's other
// two categories live in the partner warp: high words through xhi[wid][j][lane], one pair barrier.
__device__ __forceinline__ void w2_rescale(double (&n)[W2_V], int &sc, int lane, int wid, int (*xhi)[4][32], int bar_id) {
  int hi[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    hi[j] = max(__double2hiint(n[j]), __double2hiint(n[4 + j]));
    xhi[wid][j][lane] = hi[j];
  }
  w2_bar(bar_id);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    hi[j] = max(hi[j], xhi[wid ^ 1][j][lane]);
    const unsigned int big = __ballot_sync(0xffffffffu, hi[j] >= 0x2FF00000);
    if (((big >> (lane & 28)) & 15u) == 0u) {
      n[j] *= TWO256; n[4 + j] *= TWO256;
      sc += (j == (lane & 3)) ? 1 : 0;
    }
  }
}

template <bool TO_REGS>
__device__ __forceinline__ void w2_chain(const double (&n)[W2_V], const double (&bf)[2], const double (&h)[W2_V],
                                         const double *tvv, int lane, double (&a_out)[W2_V], double *stk_lv, bool push,
                                         double (&u)[4]) {
#pragma unroll
  for (int rl = 0; rl < 2; rl++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int i = rl * 4 + j;
      double a, g;
      wm_dmma(a, g, n[i], bf[rl]);
      if (TO_REGS) a_out[i] = a;
      if (push) stk_lv[i * 32 + lane] = a;
      const double ht = h[i] * tvv[i * 32 + lane];
      u[j] = rl == 0 ? g * ht : fma(g, ht, u[j]);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void w2_search(const LikWalkGroup g, const LikWalkVisit *__restrict__ visits, int slice,
                                          int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
                                          double w_out, double pi_q, int out_off, const double *__restrict__ pmat,
                                          const double *__restrict__ clv, const int *__restrict__ scl,
                                          double *__restrict__ stk, int *__restrict__ stks, double *park,
                                          int (*xhi)[2][4][32], double (*xu)[32], int wid, int rbase, int bar_id,
                                          double *__restrict__ partial, int *__restrict__ partial_e, int pstride, int lane) {
  const int64_t p0 = (int64_t)slice * WM_PATTERNS;
  const int v_end = g.first + g.count;
  int dcur = wm_desc_load(visits, g.first, lane);
  int dnxt = g.count > 1 ? wm_desc_load(visits, g.first + 1, lane) : 0;
  double bf1[2], bf2[2];
  double *tvv = park, *ph1 = park + W2_V * 32;
  int st;
  {
    double x[W2_V], bv[2], t[W2_V], dump[W2_V];
    w2_load_raw(g.tv, n_rows, Ppad, p0, lane, rbase, out_off, codes, clv, scl, x, st);
    w2_frag(pmat, g.p_v, lane, rbase, bv);
    w2_pair(x, bv, t, dump);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < W2_V; i++) tvv[i * 32 + lane] = t[i] * pi_q;
  }
  int sk = 0;
  double ak[W2_V];
#pragma unroll
  for (int i = 0; i < W2_V; i++) ak[i] = 0.0;
  for (int vi = g.first; vi < v_end; vi++) {
    const int x_in = __shfl_sync(0xffffffffu, dcur, 0);
    const int d1 = __shfl_sync(0xffffffffu, dcur, 2), d2 = __shfl_sync(0xffffffffu, dcur, 5);
    const int tree1 = __shfl_sync(0xffffffffu, dcur, 8), tree2 = __shfl_sync(0xffffffffu, dcur, 9);
    const int push = __shfl_sync(0xffffffffu, dcur, 11);
    int dnn = 0;
    if (vi + 2 < v_end) dnn = wm_desc_load(visits, vi + 2, lane);
    w2_frag(pmat, __shfl_sync(0xffffffffu, dcur, 3), lane, rbase, bf1);
    w2_frag(pmat, __shfl_sync(0xffffffffu, dcur, 6), lane, rbase, bf2);
    if (x_in >= 0) {
      double x[W2_V], bi[2], dump[W2_V];
      w2_load_raw(x_in, n_rows, Ppad, p0, lane, rbase, out_off, codes, clv, scl, x, sk);
      w2_frag(pmat, __shfl_sync(0xffffffffu, dcur, 1), lane, rbase, bi);
      w2_pair(x, bi, ak, dump);
    } else if (x_in <= -2) {
      const int lvl = -2 - x_in;
      const double *lv = stk + (int64_t)lvl * WM_V * 32 + rbase * 4 * 32;
#pragma unroll
      for (int i = 0; i < W2_V; i++) ak[i] = lv[i * 32 + lane];
      sk = stks[lvl * 4 * 32];
    }
    double n2[W2_V];
    int s1, sc2;
    {
      double x[W2_V], h1[W2_V];
      w2_load_raw(d1, n_rows, Ppad, p0, lane, rbase, out_off, codes, clv, scl, x, s1);
      w2_pair(x, bf1, n2, h1);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < W2_V; i++) {
        n2[i] *= ak[i];
        ph1[i * 32 + lane] = h1[i];
      }
    }
    sc2 = sk + s1;
    w2_rescale(n2, sc2, lane, wid, xhi[0], bar_id);
    double n1[W2_V], h2[W2_V];
    int s2, sc1;
    {
      double x[W2_V];
      w2_load_raw(d2, n_rows, Ppad, p0, lane, rbase, out_off, codes, clv, scl, x, s2);
      w2_pair(x, bf2, n1, h2);
#pragma unroll
      for (int i = 0; i < W2_V; i++) n1[i] *= ak[i];
    }
    sc1 = sk + s2;
    w2_rescale(n1, sc1, lane, wid, xhi[1], bar_id);
    double u1[4], u2[4];
    {
      double dummy[W2_V];
      w2_chain<false>(n2, bf2, h2, tvv, lane, dummy, stk + (int64_t)(push >= 0 ? push : 0) * WM_V * 32 + rbase * 4 * 32,
                      push >= 0, u2);
      if (push >= 0) stks[push * 4 * 32] = sc2;
    }
    {
      double h1[W2_V];
#pragma unroll
      for (int i = 0; i < W2_V; i++) h1[i] = ph1[i * 32 + lane];
      w2_chain<true>(n1, bf1, h1, tvv, lane, ak, nullptr, false, u1);
    }
    sk = sc1;
    // the other two categories' share of the site sums: the odd warp hands them over, the even one emits
    if (wid) {
#pragma unroll
      for (int j = 0; j < 4; j++) { xu[j][lane] = u1[j]; xu[4 + j][lane] = u2[j]; }
    }
    w2_bar(bar_id);
    if (!wid) {
#pragma unroll
      for (int j = 0; j < 4; j++) { u1[j] += xu[j][lane]; u2[j] += xu[4 + j][lane]; }
      if (tree2 >= 0) {
        const int64_t o = (int64_t)tree2 * pstride + slice;
        wm_emit<MODE == 1>(wm_site_L(u2, lane), sc2 + s2 + st, w_out, lane, partial + o, partial_e + o);
      }
      if (tree1 >= 0) {
        const int64_t o = (int64_t)tree1 * pstride + slice;
        wm_emit<MODE == 1>(wm_site_L(u1, lane), sc1 + s1 + st, w_out, lane, partial + o, partial_e + o);
      }
    }
    dcur = dnxt; dnxt = dnn;
  }
}

__global__ void __launch_bounds__(W2_THREADS, PLG_W2_MINB)
walk4m2_kernel(const LikWalkGroup *__restrict__ groups, int n_groups, const LikWalkVisit *__restrict__ visits,
               int64_t n_items, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
               const double *__restrict__ weights, const ModelDev *__restrict__ model,
               const double *__restrict__ pmat, const double *__restrict__ clv, const int *__restrict__ scl,
               double *__restrict__ stack_clv, int *__restrict__ stack_scl, int stack_levels,
               double *__restrict__ partial, int *__restrict__ partial_e, int pstride, unsigned long long *counter) {
  extern __shared__ __align__(16) unsigned char w2_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3;
  const int pair = warp >> 1, wid = warp & 1, rbase = wid * 2, bar_id = 1 + pair;
  unsigned char *pb = w2_smem + (size_t)pair * W2_PAIR_BYTES;
  double *park = reinterpret_cast<double *>(pb) + wid * W2_PARK_DOUBLES;
  int(*xhi)[2][4][32] = reinterpret_cast<int(*)[2][4][32]>(pb + 2 * W2_PARK_DOUBLES * 8);
  double(*xu)[32] = reinterpret_cast<double(*)[32]>(pb + 2 * W2_PARK_DOUBLES * 8 + 2 * 2 * 4 * 32 * 4);
  unsigned long long *slot = reinterpret_cast<unsigned long long *>(pb + 2 * W2_PARK_DOUBLES * 8 + 2 * 2 * 4 * 32 * 4 + 8 * 32 * 8);
  const int64_t gp = (int64_t)blockIdx.x * (W2_WARPS / 2) + pair;
  double *stk = stack_clv + gp * stack_levels * WM_V * 32;
  int *stks = stack_scl + gp * stack_levels * 4 * 32 + lane;
  const double pi_q = model->freqs[q];
  const int out_off = 16 * (q >> 1) + 2 * (lane >> 2) + (q & 1);
  for (;;) {
    if (!wid && lane == 0) *slot = atomicAdd(counter, 1ull);
    w2_bar(bar_id);
    const unsigned long long item = *slot;
    w2_bar(bar_id);  // (the slot is rewritten by the even warp only after both have read it)
    if ((int64_t)item >= n_items) break;
    const int slice = (int)(item / (unsigned)n_groups);
    const LikWalkGroup g = groups[item - (unsigned long long)slice * n_groups];
    const double w_out = weights[(int64_t)slice * WM_PATTERNS + out_off];
    const bool unit = __all_sync(0xffffffffu, w_out == 1.0 || w_out == 0.0);
    if (unit)
      w2_search<1>(g, visits, slice, n_rows, Ppad, codes, w_out, pi_q, out_off, pmat, clv, scl, stk, stks, park, xhi, xu,
                   wid, rbase, bar_id, partial, partial_e, pstride, lane);
    else
      w2_search<0>(g, visits, slice, n_rows, Ppad, codes, w_out, pi_q, out_off, pmat, clv, scl, stk, stks, park, xhi, xu,
                   wid, rbase, bar_id, partial, partial_e, pstride, lane);
  }
}

}  // namespace plg
