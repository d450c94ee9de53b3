// DNA x Gamma-4 SPR search walk on the FP64 tensor pipe (DMMA m8n8k4).
//
// One warp carries a 32-pattern slice through a whole SPR search (all regraft positions of one
// pruned subtree); partials never leave the SM.  The 4x4 products run as mma.sync.m8n8k4.f64: the
// matrix is a FRAGMENT (one double per lane and category, fetched with one 32-byte load from the
// pair table of pairfrag4_kernel), the vectors are the other fragment, and one instruction does 256 FMAs.  B200 measurement (plg_fp64_peak, bench.py): DMMA and DFMA share one
// pipe -- so this buys no flops, it removes 7/8 of the issue slots and all of the P-row traffic
// of a scalar formulation.
//
// Layout ("V"): lane = 4g + q holds STATE q of pattern g of each of four 8-pattern groups j, for
// the four categories r: v[r*4+j].  pattern(j, g) = 16 (j >> 1) + 2 g + (j & 1), so a lane reads
// two adjacent patterns with one 16-byte load.  With A = the vector (8 patterns x 4 states) and
// B[k][n] = (n odd ? P_h : P_t)[n >> 1][k] -- two matrices of one branch pair interleaved -- the
// accumulator fragment comes out as d0 = (P_t x)[q], d1 = (P_h x)[q] for pattern g: again layout V,
// no transposition between chained products.  Every product of a visit is such a pair:
//     D1 -> (P_t1 D1, P_h1 D1)      N1 = a o P_t2 D2 -> (a' toward c1 = P_t1 N1, P_h1 N1)
//     D2 -> (P_t2 D2, P_h2 D2)      N2 = a o P_t1 D1 -> (a' toward c2 = P_t2 N2, P_h2 N2)
// 64 DMMAs per visit, none wasted when the search continues; what a visit hands on (registers or
// stack) is a' = P_in N, already pushed through the next incoming branch.
//
// Site terms: a neighbour's lnL over the slice is sum_p w_p (log L_p + scalers).  With unit
// weights (the common case: distinct patterns of a long alignment) that is log prod_p L_p, so the
// kernel takes NO logarithm: every lane splits its L into mantissa and exponent (integer ops),
// the warp multiplies the 32 mantissas (< 2^32) and adds the exponents (scaling counters are
// exponents too: -256 each), and stores (M, E) per (neighbour, slice); reduce_me_kernel turns
// them into log M + E ln 2 -- one FP64 log per 32 site terms instead of one per term, and none on
// the walk's critical path (ncu round 1: the log was 18 % of the walk's instructions and the
// longest dependent chain of a visit).  Slices that hold a weight other than 0 or 1 take the
// per-lane log path and store (sum, WM_PLAIN).
#pragma once

namespace plg {

#ifndef PLG_WM_WARPS
#define PLG_WM_WARPS 4
#endif
constexpr int WM_WARPS = PLG_WM_WARPS;
constexpr int WM_THREADS = 32 * WM_WARPS;
#ifndef PLG_WM_MINB
#define PLG_WM_MINB 3
#endif
constexpr int WM_V = 16;          // doubles per vector per lane: [category r][group j] = v[r * 4 + j]
constexpr int WM_PATTERNS = 32;   // patterns per warp item
constexpr int WM_SMEM_DOUBLES_PER_WARP = 2 * WM_V * 32;  // parked: tvv, h1
constexpr int WM_PLAIN = (int)0x80808080;  // exponent marker: the double is a finished sum of site terms (memset byte 0x80)
constexpr double WM_LN2 = 0.693147180559945309417232121458176568;

__device__ __forceinline__ void wm_dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
      : "=d"(d0), "=d"(d1)
      : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// (t, h) = (P_t x, P_h x) for the branch pair whose interleaved fragment is bf[r]
__device__ __forceinline__ void wm_pair(const double (&x)[WM_V], const double (&bf)[4], double (&t)[WM_V],
                                        double (&h)[WM_V]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int j = 0; j < 4; j++) wm_dmma(t[r * 4 + j], h[r * 4 + j], x[r * 4 + j], bf[r]);
}

// fragment of the branch pair (P(t), P(t/2)) of P-matrix id `pm`: one 32-byte load (pairfrag4_kernel)
__device__ __forceinline__ void wm_frag(const double *__restrict__ pfrag, int pm, int lane, double (&bf)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];\n"
               : "=d"(bf[0]), "=d"(bf[1]), "=d"(bf[2]), "=d"(bf[3])
               : "l"(pfrag + ((int64_t)pm * 32 + lane) * 4));
}

// Operand `id` as it sits at its node, in layout V (tips: indicator of the code's states; one code
// path for both kinds -- a separate tip path doubles the DMMA blocks and costs more in spills than
// the 32 register moves it saves).
__device__ __forceinline__ void wm_load_raw(int id, int n_rows, int64_t Ppad, int64_t p0, int lane, int out_off,
                                            const unsigned char *__restrict__ codes, const double *__restrict__ clv,
                                            const int *__restrict__ scl, double (&x)[WM_V], int &s) {
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
#pragma unroll
      for (int r = 0; r < 4; r++) x[r * 4 + j] = ind;
    }
    s = 0;
  } else {
    const double *b = clv + (int64_t)(id - n_rows) * 16 * Ppad + p0 + g2;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double2 u = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad));
      const double2 v = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad + 16));
      x[r * 4] = u.x; x[r * 4 + 1] = u.y; x[r * 4 + 2] = v.x; x[r * 4 + 3] = v.y;
    }
    // scaling counter of THIS LANE's output pattern only (see wm_rescale)
    s = __ldg(scl + (int64_t)(id - n_rows) * Ppad + p0 + out_off);
  }
}


// Parked vectors (shared memory, per warp) and stack levels (global, per warp): element i = r * 4 + j
// of this lane sits at [i][lane] (8-byte accesses; 16- and 32-byte layouts measured slower: the wide
// register tuples cost more in spills than the instructions they save).
__device__ __forceinline__ void wm_park_store(double *base, int lane, const double (&v)[WM_V]) {
#pragma unroll
  for (int i = 0; i < WM_V; i++) base[i * 32 + lane] = v[i];
}
__device__ __forceinline__ void wm_park_load(const double *base, int lane, double (&v)[WM_V]) {
#pragma unroll
  for (int i = 0; i < WM_V; i++) v[i] = base[i * 32 + lane];
}
__device__ __forceinline__ double wm_park_at(const double *base, int lane, int i) {
  return base[i * 32 + lane];
}
// stack level base `lv` = stack + level * WM_V * 32 (no lane offset)
__device__ __forceinline__ void wm_stack_store(double *lv, int lane, const double (&v)[WM_V]) {
#pragma unroll
  for (int i = 0; i < WM_V; i++) lv[i * 32 + lane] = v[i];
}
__device__ __forceinline__ void wm_stack_load(const double *lv, int lane, double (&v)[WM_V]) {
#pragma unroll
  for (int i = 0; i < WM_V; i++) v[i] = lv[i * 32 + lane];
}

// per pattern: all 16 entries below 2^-256 -> multiply by 2^256 and count (libpll's scaling).
// A pattern's entries sit in the four lanes of a quad: one vote per pattern group.  The four lanes
// of a quad would keep identical counters for their four groups; each lane keeps only the one its
// own site term needs (group jq = lane & 3).  (A warp vote in front of the four ballots -- skip them
// unless some lane holds an all-small group -- measured slower: 5.17 vs 4.90 ms; testing the ballot's
// own result for a zero nibble costs no vote.)
__device__ __forceinline__ void wm_rescale(double (&n)[WM_V], int &sc, int lane) {
  int hi[4];
#pragma unroll
  for (int j = 0; j < 4; j++)
    hi[j] = max(max(__double2hiint(n[j]), __double2hiint(n[4 + j])),
                max(__double2hiint(n[8 + j]), __double2hiint(n[12 + j])));
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned int big = __ballot_sync(0xffffffffu, hi[j] >= 0x2FF00000);
    // `big` is the same in every lane: a quad needs the factor iff its nibble is 0.  The common case -- no
    // nibble is 0 -- is a UNIFORM branch around the multiplies (predicated off they still took an issue
    // slot each: 9 % of the walk's instructions in profiles/r2q_walk4m_ncu.txt)
    if ((big - 0x11111111u) & ~big & 0x88888888u) {
      if (((big >> (lane & 28)) & 15u) == 0u) {
#pragma unroll
        for (int r = 0; r < 4; r++) n[r * 4 + j] *= TWO256;
        sc += (j == (lane & 3)) ? 1 : 0;
      }
    }
  }
}

// Site likelihoods from the per-group sums v[j] = sum_r e[r][j] (this lane's state q; pi_q is already
// inside: wm_search folds it into the pruned subtree's vector once per search): transpose-reduce over the quad, lane q ends with L of pattern (j = q, g).
__device__ __forceinline__ double wm_site_L(const double (&v)[4], int lane) {
  const bool hi = lane & 2, odd = lane & 1;
  double k0 = hi ? v[2] : v[0], k1 = hi ? v[3] : v[1];
  k0 += __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 2);
  k1 += __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 2);
  double L = odd ? k1 : k0;
  L += __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
  return L;
}

// One neighbour's contribution of this slice -> partial / partial_e (see the header comment).
// L = 4 x the site likelihood of this lane's pattern, scq its scaling count, w its weight.
template <bool UNIT>
__device__ __forceinline__ void wm_emit(double L, int scq, double w, int lane, double *__restrict__ dst,
                                        int *__restrict__ dst_e) {
  if (UNIT) {
    int hi = __double2hiint(L);
    int e = -2 - 256 * scq;            // the 1/4 of the category mean, 2^-256 per scaling event
    if (hi < 0x00100000) {             // zero or subnormal (pathological data): renormalise exactly
      L *= TWO256; L *= TWO256;
      hi = __double2hiint(L);
      e -= 512;
    }
    double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(L));
    e += (hi >> 20) - 1023;
    if (hi < 0x00100000) m = 0.0;      // L == 0: the product is 0 and its log -inf
    if (w == 0.0) { m = 1.0; e = 0; }  // padding pattern
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m *= __shfl_down_sync(0xffffffffu, m, d);
    e = __reduce_add_sync(0xffffffffu, e);
    if (lane == 0) { *dst = m; *dst_e = e; }
  } else {
    double val = w != 0.0 ? w * (log(L * 0.25) + scq * LOG_MINLH) : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
    if (lane == 0) { *dst = val; *dst_e = WM_PLAIN; }
  }
}


// One neighbour chain of a visit in straight-line form: (a, g) = (P_t n, P_h n) category by
// category; g is consumed at once into the per-group sums of g o h o tvv, a goes to `a_out`
// (registers) and/or the stack level `stk_lv` -- so neither a nor g is ever live as a whole vector
// (the register allocator spills them otherwise at 168 registers).
template <bool TO_REGS>
__device__ __forceinline__ void wm_chain(const double (&n)[WM_V], const double (&bf)[4], const double (&h)[WM_V],
                                         const double *tvv, int lane, double (&a_out)[WM_V], double *stk_lv,
                                         bool push, double (&u)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    double a4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int i = r * 4 + j;
      double g;
      wm_dmma(a4[j], g, n[i], bf[r]);
      if (TO_REGS) a_out[i] = a4[j];
      if (push) stk_lv[i * 32 + lane] = a4[j];
      const double ht = h[i] * wm_park_at(tvv, lane, i);
      u[j] = r == 0 ? g * ht : fma(g, ht, u[j]);
    }
  }
}

__device__ __forceinline__ int wm_desc_load(const LikWalkVisit *__restrict__ visits, int vi, int lane) {
  int v = 0;
  if (lane < 12) asm volatile("ld.global.nc.b32 %0, [%1];\n" : "=r"(v) : "l"(reinterpret_cast<const int *>(visits + vi) + lane));
  return v;
}

// MODE 1: unit weights (mantissa/exponent records), 0: weighted (per-lane log)
template <int MODE>
__device__ __forceinline__ void wm_search(const LikWalkGroup g, const LikWalkVisit *__restrict__ visits, int slice,
                                          int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
                                          double w_out, double pi_q, int out_off, const double *__restrict__ pmat,
                                          const double *__restrict__ clv, const int *__restrict__ scl,
                                          double *__restrict__ stk, int *__restrict__ stks,
                                          double (*park)[WM_V][32], double *__restrict__ partial,
                                          int *__restrict__ partial_e, int pstride, int lane) {
  const int64_t p0 = (int64_t)slice * WM_PATTERNS;
  const int v_end = g.first + g.count;
  // pipeline: descriptors two visits ahead (one int per lane)
  int dcur = wm_desc_load(visits, g.first, lane);
  int dnxt = g.count > 1 ? wm_desc_load(visits, g.first + 1, lane) : 0;
  double bf1[4], bf2[4];
  int st;  // scaling counters: this lane's output pattern only (wm_rescale)
  {  // the pruned subtree seen through its branch, once per search
    double x[WM_V], bv[4], t[WM_V], dump[WM_V];
    wm_load_raw(g.tv, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, st);
    wm_frag(pmat, g.p_v, lane, bv);
    wm_pair(x, bv, t, dump);
#pragma unroll
    for (int i = 0; i < WM_V; i++) t[i] *= pi_q;  // every site term of the search carries pi_q (x) this vector
    __syncwarp();
    wm_park_store(&park[0][0][0], lane, t);
  }
  int sk = 0;
  double ak[WM_V];  // a = P_in (incoming partial), handed from visit to visit
#pragma unroll
  for (int i = 0; i < WM_V; i++) ak[i] = 0.0;
  for (int vi = g.first; vi < v_end; vi++) {
    const int x_in = __shfl_sync(0xffffffffu, dcur, 0);
    const int d1 = __shfl_sync(0xffffffffu, dcur, 2), d2 = __shfl_sync(0xffffffffu, dcur, 5);
    const int tree1 = __shfl_sync(0xffffffffu, dcur, 8), tree2 = __shfl_sync(0xffffffffu, dcur, 9);
    const int push = __shfl_sync(0xffffffffu, dcur, 11);
    int dnn = 0;
    if (vi + 2 < v_end) dnn = wm_desc_load(visits, vi + 2, lane);
    wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 3), lane, bf1);  // (P_t1, P_h1): the half is implied by the branch
    wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 6), lane, bf2);
    if (x_in >= 0) {  // first step of a side: a = P_in X from a stored operand
      double x[WM_V], bi[4], dump[WM_V];
      wm_load_raw(x_in, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, sk);
      wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 1), lane, bi);
      wm_pair(x, bi, ak, dump);
    } else if (x_in <= -2) {
      const int lvl = -2 - x_in;
      wm_stack_load(stk + (int64_t)lvl * WM_V * 32, lane, ak);
      sk = stks[lvl * 4 * 32];
    }
    // D1 once: N2 = a o P_t1 D1 (the partial toward d2) and h1 = P_h1 D1 (far side of edge 1)
    double n2[WM_V];
    int s1, sc2;
    {
      double x[WM_V], h1[WM_V];
      wm_load_raw(d1, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, s1);
      wm_pair(x, bf1, n2, h1);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < WM_V; i++) n2[i] *= ak[i];
      wm_park_store(&park[1][0][0], lane, h1);
    }
    sc2 = sk + s1;
    wm_rescale(n2, sc2, lane);
    // D2 once: N1 = a o P_t2 D2 and h2 = P_h2 D2
    double n1[WM_V], h2[WM_V];
    int s2, sc1;
    {
      double x[WM_V];
      wm_load_raw(d2, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, s2);
      wm_pair(x, bf2, n1, h2);
#pragma unroll
      for (int i = 0; i < WM_V; i++) n1[i] *= ak[i];
    }
    sc1 = sk + s2;
    wm_rescale(n1, sc1, lane);
    // Visits come in normal form (planners): the partial toward c1 continues in registers, the
    // one toward c2 is at most parked.
    double u1[4], u2[4];
    {  // side 2 first, so that its hand-over is on the stack (and dead) before side 1's lands directly in ak
      double dummy[WM_V];
      wm_chain<false>(n2, bf2, h2, &park[0][0][0], lane, dummy, stk + (int64_t)(push >= 0 ? push : 0) * WM_V * 32, push >= 0, u2);
      if (push >= 0) stks[push * 4 * 32] = sc2;
    }
    if (tree2 >= 0) {  // neighbour 2 = insertion into x--c2
      const int64_t o = (int64_t)tree2 * pstride + slice;
      wm_emit<MODE == 1>(wm_site_L(u2, lane), sc2 + s2 + st, w_out, lane, partial + o, partial_e + o);
    }
    {
      double h1[WM_V];
      wm_park_load(&park[1][0][0], lane, h1);
      wm_chain<true>(n1, bf1, h1, &park[0][0][0], lane, ak, nullptr, false, u1);
    }
    sk = sc1;
    if (tree1 >= 0) {  // neighbour 1 = insertion into x--c1
      const int64_t o = (int64_t)tree1 * pstride + slice;
      wm_emit<MODE == 1>(wm_site_L(u1, lane), sc1 + s1 + st, w_out, lane, partial + o, partial_e + o);
    }
    dcur = dnxt; dnxt = dnn;
  }
}

__global__ void __launch_bounds__(WM_THREADS, PLG_WM_MINB)
walk4m_kernel(const LikWalkGroup *__restrict__ groups, int n_groups, const LikWalkVisit *__restrict__ visits,
              int64_t n_items, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
              const double *__restrict__ weights, const ModelDev *__restrict__ model,
              const double *__restrict__ pmat, const double *__restrict__ clv, const int *__restrict__ scl,
              double *__restrict__ stack_clv, int *__restrict__ stack_scl, int stack_levels,
              double *__restrict__ partial, int *__restrict__ partial_e, int pstride, unsigned long long *counter) {
  extern __shared__ __align__(16) double wm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3;
  double(*park)[WM_V][32] = reinterpret_cast<double(*)[WM_V][32]>(wm_smem + warp * WM_SMEM_DOUBLES_PER_WARP);  // [0] tvv, [1] h1
  const int64_t gw = (int64_t)blockIdx.x * WM_WARPS + warp;
  double *stk = stack_clv + gw * stack_levels * WM_V * 32;  // level l at stk + l * WM_V * 32 (wm_stack_store)
  int *stks = stack_scl + gw * stack_levels * 4 * 32 + lane;
  const double pi_q = model->freqs[q];
  // the pattern whose site term this lane ends up with (wm_site_L): pattern(j = q, g)
  const int out_off = 16 * (q >> 1) + 2 * (lane >> 2) + (q & 1);
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((int64_t)item >= n_items) break;
    const int slice = (int)(item / (unsigned)n_groups);
    const LikWalkGroup g = groups[item - (unsigned long long)slice * n_groups];
    const double w_out = weights[(int64_t)slice * WM_PATTERNS + out_off];
    const bool unit = __all_sync(0xffffffffu, w_out == 1.0 || w_out == 0.0);
    if (unit)
      wm_search<1>(g, visits, slice, n_rows, Ppad, codes, w_out, pi_q, out_off, pmat, clv, scl, stk, stks, park,
                   partial, partial_e, pstride, lane);
    else
      wm_search<0>(g, visits, slice, n_rows, Ppad, codes, w_out, pi_q, out_off, pmat, clv, scl, stk, stks, park,
                   partial, partial_e, pstride, lane);
  }
}

// lnL[t] = sum over the slices of a tree's (M, E) records: log M + E ln 2, or M itself where the
// record is a finished sum (E == WM_PLAIN: weighted slices, evaluations fused into level ops).
// One warp per tree, strided partial sums, then a fixed shuffle tree (deterministic order).
__global__ void reduce_me_kernel(int n_trees, int pstride, const double *__restrict__ partial,
                                 const int *__restrict__ partial_e, double *__restrict__ lnl) {
  const int t = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (t >= n_trees) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < pstride; b += 32) {
    const double m = partial[(int64_t)t * pstride + b];
    const int e = partial_e[(int64_t)t * pstride + b];
    s += e == WM_PLAIN ? m : log(m) + (double)e * WM_LN2;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) lnl[t] = s;
}

}  // namespace plg
