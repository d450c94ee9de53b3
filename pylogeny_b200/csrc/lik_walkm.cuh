// DNA x Gamma-4 SPR search walk on the FP64 tensor pipe (DMMA m8n8k4).
//
// Same walk as walk4_kernel (lik_walk.cuh): one warp carries a 32-pattern slice through a whole
// SPR search, partials never leave the SM.  What changes is WHERE the 4x4 products run.  ncu on
// the scalar walk (profiles/r1c_walk4_kernel_ncu.txt) shows the FP64 pipe 27 % busy and the warps
// waiting on shared-memory broadcasts of P rows (32 LDS.128 per 64 FMAs) -- the matrices are
// warp-uniform but every lane needs every entry.  mma.sync.m8n8k4.f64 turns that around: the
// matrix is a FRAGMENT (one double per lane), the vectors are the other fragment, and one
// instruction does 256 FMAs.  B200 measurement (profiles/r1c_fp64_peak.txt): DMMA 37 TFLOP/s,
// DFMA 34 TFLOP/s, one shared pipe -- so this buys no flops, it removes 7/8 of the issue slots
// and all of the P-row traffic.
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
#ifndef PLG_WM_PREFETCH
#define PLG_WM_PREFETCH 0  // 1: matrix fragments of the next visit fetched during this one (16 more registers: 7.65 vs 7.09 ms)
#endif
#ifndef PLG_WM_PARK_H2
#define PLG_WM_PARK_H2 0   // 1: h2 waits in shared memory instead of registers
#endif
#ifndef PLG_WM_NG
#define PLG_WM_NG 4        // 8-pattern groups per warp: 4 (32 patterns, 16 doubles per vector per lane) or 2
#endif
constexpr int WM_NG = PLG_WM_NG;
constexpr int WM_V = 4 * WM_NG;          // doubles per vector per lane: [category r][group j] = v[r * WM_NG + j]
constexpr int WM_PATTERNS = 8 * WM_NG;   // patterns per warp item
static_assert(WM_NG == 4 || WM_NG == 2, "a warp carries 32 or 16 patterns");
#ifndef PLG_WM_STAGE
#define PLG_WM_STAGE 0     // 1: the far-side operands of the NEXT visit are fetched by cp.async into shared memory during this one
                           //    (B200: 9.2 vs 7.9 ms -- 209 KB of shared memory leave no L1 for the spills and fragments)
#endif
#ifndef PLG_WM_EARLY2
#define PLG_WM_EARLY2 0     // 1: the second operand loaded before the first one is multiplied out: 5.72 vs 5.59 ms
#endif
#ifndef PLG_WM_FUSE2
#define PLG_WM_FUSE2 0     // 1: the D2 stage and neighbour chain 2 in one straight-line loop (h2 never a vector): 5.76 vs 5.58 ms
#endif
#ifndef PLG_WM_REFRAG
#define PLG_WM_REFRAG 0
#endif
#ifndef PLG_WM_PF
#define PLG_WM_PF 0        // 1: ... or pulled into L1 by one prefetch instruction per operand (32 lines, one per lane): 8.03 vs 7.76 ms
#endif
// staging area of one operand: 9 rows of 16 B per lane (8 rows CLV + 1 row scaling counters; a tip uses 8 B of row 0)
constexpr int WM_STAGE_ROWS = 9;
constexpr int WM_STAGE_DOUBLES = PLG_WM_STAGE ? 2 * WM_STAGE_ROWS * 2 * 32 : 0;  // two operands
static_assert(!PLG_WM_STAGE || WM_NG == 4, "operand staging is written for 32 patterns per warp");
constexpr int WM_SMEM_DOUBLES_PER_WARP = (2 + PLG_WM_PARK_H2) * WM_V * 32 + WM_STAGE_DOUBLES;  // parked: tvv, h1 (, h2); staged D1, D2

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
    for (int j = 0; j < WM_NG; j++) wm_dmma(t[r * WM_NG + j], h[r * WM_NG + j], x[r * WM_NG + j], bf[r]);
}

// fragment of a branch pair: lane (g, q) holds (g odd ? P[pm_h] : P[pm_t])[r][g >> 1][q]
__device__ __forceinline__ void wm_frag(const double *__restrict__ pmat, int pm_t, int pm_h, int lane, double (&bf)[4]) {
  const double *src = pmat + (int64_t)((lane >> 2) & 1 ? pm_h : pm_t) * 68 + (lane >> 3) * 4 + (lane & 3);
#pragma unroll
  for (int r = 0; r < 4; r++) asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(bf[r]) : "l"(src + r * 17));
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
    unsigned int m[WM_NG];
    const unsigned int lo = *reinterpret_cast<const unsigned short *>(c);
    m[0] = lo & 255u; m[1] = lo >> 8;
    if constexpr (WM_NG == 4) {
      const unsigned int hi = *reinterpret_cast<const unsigned short *>(c + 16);
      m[2] = hi & 255u; m[3] = hi >> 8;
    }
#pragma unroll
    for (int j = 0; j < WM_NG; j++) {
      const unsigned int mask = m[j] ? m[j] : 15u;
      const double ind = (mask >> q & 1u) ? 1.0 : 0.0;
#pragma unroll
      for (int r = 0; r < 4; r++) x[r * WM_NG + j] = ind;
    }
    s = 0;
  } else {
    const double *b = clv + (int64_t)(id - n_rows) * 16 * Ppad + p0 + g2;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double2 u = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad));
      x[r * WM_NG] = u.x; x[r * WM_NG + 1] = u.y;
      if constexpr (WM_NG == 4) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(b + (int64_t)(r * 4 + q) * Ppad + 16));
        x[r * WM_NG + 2] = v.x; x[r * WM_NG + 3] = v.y;
      }
    }
    // scaling counter of THIS LANE's output pattern only (see wm_rescale)
    s = __ldg(scl + (int64_t)(id - n_rows) * Ppad + p0 + out_off);
  }
}

// Asynchronous copy of operand `id` (as wm_load_raw reads it) into this lane's staging slots:
// every lane fetches exactly the bytes it will consume, so no cross-lane synchronisation is needed
// (cp.async.wait_group is per thread).  sdst = shared address of row 0 of this lane; rows are 512 B apart.
__device__ __forceinline__ void wm_stage(int id, int n_rows, int64_t Ppad, int64_t p0, int lane,
                                         const unsigned char *__restrict__ codes, const double *__restrict__ clv,
                                         const int *__restrict__ scl, unsigned int sdst) {
  const int g2 = (lane >> 2) * 2, q = lane & 3;
  if (id < n_rows) {
    const unsigned char *c = codes + (int64_t)id * Ppad + p0 + (g2 & ~3);  // the aligned word holding codes g2, g2+1
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sdst), "l"(c));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sdst + 4), "l"(c + 16));
  } else {
    const double *b = clv + (int64_t)(id - n_rows) * 16 * Ppad + p0 + g2;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double *src = b + (int64_t)(r * 4 + q) * Ppad;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst + (2 * r) * 512), "l"(src));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst + (2 * r + 1) * 512), "l"(src + 16));
    }
    const int *sp = scl + (int64_t)(id - n_rows) * Ppad + p0 + g2;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst + 8 * 512), "l"(sp));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst + 8 * 512 + 8), "l"(sp + 16));
  }
}

// the staged operand in layout V (what wm_load_raw returns); src = this lane's row 0
__device__ __forceinline__ void wm_unstage(int id, int n_rows, int lane, const double2 *src, double (&x)[WM_V],
                                           int &s) {
  const int q = lane & 3;
  if (id < n_rows) {
    const uint2 w = *reinterpret_cast<const uint2 *>(src);
    const int sh = ((lane >> 2) & 1) * 16;  // (g2 & 2) * 8: which half of the aligned word
    const unsigned int lo = (w.x >> sh) & 0xFFFFu, hi = (w.y >> sh) & 0xFFFFu;
    unsigned int m[4] = {lo & 255u, lo >> 8, hi & 255u, hi >> 8};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned int mask = m[j] ? m[j] : 15u;
      const double ind = (mask >> q & 1u) ? 1.0 : 0.0;
#pragma unroll
      for (int r = 0; r < 4; r++) x[r * 4 + j] = ind;
    }
    s = 0;
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double2 u = src[(2 * r) * 32], v = src[(2 * r + 1) * 32];
      x[r * 4] = u.x; x[r * 4 + 1] = u.y; x[r * 4 + 2] = v.x; x[r * 4 + 3] = v.y;
    }
    const int4 c = *reinterpret_cast<const int4 *>(src + 8 * 32);
    s = q == 0 ? c.x : (q == 1 ? c.y : (q == 2 ? c.z : c.w));
  }
}

// L1 prefetch of operand `id` for this warp's slice: an inner CLV is 16 rows x 2 half-slices of
// 128 B (one line per lane) + one line of scaling counters; a tip is one 32-byte piece of a code row.
__device__ __forceinline__ void wm_prefetch(int id, int n_rows, int64_t Ppad, int64_t p0, int lane,
                                            const unsigned char *__restrict__ codes, const double *__restrict__ clv,
                                            const int *__restrict__ scl) {
  if (id < n_rows) {
    if (lane == 0) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(codes + (int64_t)id * Ppad + p0));
  } else {
    const double *b = clv + (int64_t)(id - n_rows) * 16 * Ppad + p0 + (int64_t)(lane >> 1) * Ppad + (lane & 1) * 16;
    asm volatile("prefetch.global.L1 [%0];\n" ::"l"(b));
    if (lane == 0) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(scl + (int64_t)(id - n_rows) * Ppad + p0));
  }
}

// per pattern: all 16 entries below 2^-256 -> multiply by 2^256 and count (libpll's scaling).
// A pattern's entries sit in the four lanes of a quad: one vote per pattern group.  The four lanes
// of a quad would keep identical counters for their four groups; each lane keeps only the one its
// own site term needs (group jq = lane & 3; lane & 1 with two groups) -- one register instead of four
// for each of the six counters a visit carries.
__device__ __forceinline__ void wm_rescale(double (&n)[WM_V], int &sc, int lane) {
  const int jq = WM_NG == 4 ? (lane & 3) : (lane & 1);
#pragma unroll
  for (int j = 0; j < WM_NG; j++) {
    int hi = max(max(__double2hiint(n[j]), __double2hiint(n[WM_NG + j])),
                 max(__double2hiint(n[2 * WM_NG + j]), __double2hiint(n[3 * WM_NG + j])));
    const unsigned int big = __ballot_sync(0xffffffffu, hi >= 0x2FF00000);
    if (((big >> (lane & 28)) & 15u) == 0u) {
#pragma unroll
      for (int r = 0; r < 4; r++) n[r * WM_NG + j] *= TWO256;
      sc += (j == jq) ? 1 : 0;
    }
  }
}

// sum over the warp's 32 patterns of w (log L + scaling), L = 1/4 sum_r sum_k pi_k e[r][k];
// the result is valid in lane 0
__device__ __forceinline__ double wm_site_sum(const double (&e)[WM_V], int scq, double w_out,
                                              double pi_q, int lane) {
  double v[WM_NG];
#pragma unroll
  for (int j = 0; j < WM_NG; j++) v[j] = pi_q * ((e[j] + e[WM_NG + j]) + (e[2 * WM_NG + j] + e[3 * WM_NG + j]));
  const bool hi = lane & 2, odd = lane & 1;
  double L;
  if constexpr (WM_NG == 4) {
    // transpose-reduce over the quad: lane q ends with the sum of group j = q
    double k0 = hi ? v[2] : v[0], k1 = hi ? v[3] : v[1];
    k0 += __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 2);
    k1 += __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 2);
    L = odd ? k1 : k0;
    L += __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
  } else {
    // two groups: lanes q = 0, 1 end with groups 0, 1 (lanes 2, 3 hold copies and contribute nothing)
    L = odd ? v[1] : v[0];
    L += __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[1], 1);
    L += __shfl_xor_sync(0xffffffffu, L, 2);
    if (hi) w_out = 0.0;
  }
  double val = w_out != 0.0 ? w_out * (log(L * 0.25) + scq * LOG_MINLH) : 0.0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
  return val;
}

// same from the per-group sums u[j] = (e0 + e1) + (e2 + e3) over the categories (not yet weighted by pi)
__device__ __forceinline__ double wm_site_sum_u(const double (&u)[WM_NG], int scq, double w_out,
                                                double pi_q, int lane) {
  double v[WM_NG];
#pragma unroll
  for (int j = 0; j < WM_NG; j++) v[j] = pi_q * u[j];
  const bool hi = lane & 2, odd = lane & 1;
  double L;
  if constexpr (WM_NG == 4) {
    double k0 = hi ? v[2] : v[0], k1 = hi ? v[3] : v[1];
    k0 += __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 2);
    k1 += __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 2);
    L = odd ? k1 : k0;
    L += __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 1);
  } else {
    L = odd ? v[1] : v[0];
    L += __shfl_xor_sync(0xffffffffu, odd ? v[0] : v[1], 1);
    L += __shfl_xor_sync(0xffffffffu, L, 2);
    if (hi) w_out = 0.0;
  }
  double val = w_out != 0.0 ? w_out * (log(L * 0.25) + scq * LOG_MINLH) : 0.0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
  return val;
}

// One neighbour chain of a visit in straight-line form: (a, g) = (P_t n, P_h n) category by
// category; g is consumed at once into the per-group sums of g o h o tvv (same summation order as
// wm_site_sum), a goes to `a_out` (registers) and/or the stack -- so neither a nor g is ever live
// as a whole vector (the register allocator spills them otherwise at 168 registers).
template <bool TO_REGS>
__device__ __forceinline__ void wm_chain(const double (&n)[WM_V], const double (&bf)[4], const double (&h)[WM_V],
                                         const double (*tvv)[32], int lane, double (&a_out)[WM_V], double *stk_dst,
                                         bool push, double (&u)[WM_NG]) {
  double s01[WM_NG], e2[WM_NG];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int j = 0; j < WM_NG; j++) {
      const int i = r * WM_NG + j;
      double a, g;
      wm_dmma(a, g, n[i], bf[r]);
      if (TO_REGS) a_out[i] = a;
      if (push) stk_dst[i * 32] = a;
      const double t = g * (h[i] * tvv[i][lane]);
      if (r == 0) s01[j] = t;
      else if (r == 1) s01[j] += t;
      else if (r == 2) e2[j] = t;
      else u[j] = s01[j] + (e2[j] + t);
    }
}

__device__ __forceinline__ int wm_desc_load(const LikWalkVisit *__restrict__ visits, int vi, int lane) {
  int v = 0;
  if (lane < 12) asm volatile("ld.global.nc.b32 %0, [%1];\n" : "=r"(v) : "l"(reinterpret_cast<const int *>(visits + vi) + lane));
  return v;
}

__global__ void __launch_bounds__(WM_THREADS, PLG_WM_MINB)
walk4m_kernel(const LikWalkGroup *__restrict__ groups, int n_groups, const LikWalkVisit *__restrict__ visits,
              int64_t n_items, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
              const double *__restrict__ weights, const ModelDev *__restrict__ model,
              const double *__restrict__ pmat, const double *__restrict__ clv, const int *__restrict__ scl,
              double *__restrict__ stack_clv, int *__restrict__ stack_scl, int stack_levels,
              double *__restrict__ partial, int pstride, unsigned long long *counter) {
  extern __shared__ __align__(16) double wm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3;
  double(*park)[WM_V][32] = reinterpret_cast<double(*)[WM_V][32]>(wm_smem + warp * WM_SMEM_DOUBLES_PER_WARP);  // [0] tvv, [1] h1
#if PLG_WM_STAGE
  // staged far-side operands of the visit about to run: [operand][row][lane] 16-byte units
  const double2 *stg = reinterpret_cast<const double2 *>(wm_smem + warp * WM_SMEM_DOUBLES_PER_WARP +
                                                         (2 + PLG_WM_PARK_H2) * WM_V * 32) + lane;
  const unsigned int stg_s = (unsigned int)__cvta_generic_to_shared(stg);
#endif
  const int64_t gw = (int64_t)blockIdx.x * WM_WARPS + warp;
  double *stk = stack_clv + gw * stack_levels * WM_V * 32 + lane;
  int *stks = stack_scl + gw * stack_levels * WM_NG * 32 + lane;
  const double pi_q = model->freqs[q];
  // the pattern whose site term this lane ends up with (wm_site_sum): pattern(j = q, g)
  const int out_off = WM_NG == 4 ? 16 * (q >> 1) + 2 * (lane >> 2) + (q & 1) : 2 * (lane >> 2) + (q & 1);
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((int64_t)item >= n_items) break;
    const int slice = (int)(item / (unsigned)n_groups);
    const LikWalkGroup g = groups[item - (unsigned long long)slice * n_groups];
    const int64_t p0 = (int64_t)slice * WM_PATTERNS;
    const double w_out = weights[p0 + out_off];
    const int v_end = g.first + g.count;
    // pipeline: descriptors two visits ahead (one int per lane), matrix fragments one visit ahead
    int dcur = wm_desc_load(visits, g.first, lane);
    int dnxt = g.count > 1 ? wm_desc_load(visits, g.first + 1, lane) : 0;
#if PLG_WM_STAGE
    wm_stage(__shfl_sync(0xffffffffu, dcur, 2), n_rows, Ppad, p0, lane, codes, clv, scl, stg_s);
    wm_stage(__shfl_sync(0xffffffffu, dcur, 5), n_rows, Ppad, p0, lane, codes, clv, scl, stg_s + WM_STAGE_ROWS * 512);
    asm volatile("cp.async.commit_group;\n" ::);
#endif
    double bf1[4], bf2[4];
    wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 3), __shfl_sync(0xffffffffu, dcur, 4), lane, bf1);
    wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 6), __shfl_sync(0xffffffffu, dcur, 7), lane, bf2);
    int st;  // scaling counters: this lane's output pattern only (wm_rescale)
    {  // the pruned subtree seen through its branch, once per search
      double x[WM_V], bv[4], t[WM_V], dump[WM_V];
      wm_load_raw(g.tv, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, st);
      wm_frag(pmat, g.p_v, g.p_v, lane, bv);
      wm_pair(x, bv, t, dump);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < WM_V; i++) park[0][i][lane] = t[i];
    }
    double ak[WM_V];  // a = P_in (incoming partial), handed from visit to visit
    int sk = 0;
#pragma unroll
    for (int i = 0; i < WM_V; i++) ak[i] = 0.0;
    for (int vi = g.first; vi < v_end; vi++) {
      const int x_in = __shfl_sync(0xffffffffu, dcur, 0), p_in = __shfl_sync(0xffffffffu, dcur, 1);
      const int d1 = __shfl_sync(0xffffffffu, dcur, 2), d2 = __shfl_sync(0xffffffffu, dcur, 5);
      const int tree1 = __shfl_sync(0xffffffffu, dcur, 8), tree2 = __shfl_sync(0xffffffffu, dcur, 9);
      const int keep = __shfl_sync(0xffffffffu, dcur, 10), push = __shfl_sync(0xffffffffu, dcur, 11);
      int dnn = 0;
#if PLG_WM_PREFETCH
      double nf1[4] = {0, 0, 0, 0}, nf2[4] = {0, 0, 0, 0};
      if (vi + 1 < v_end) {
        if (vi + 2 < v_end) dnn = wm_desc_load(visits, vi + 2, lane);
        wm_frag(pmat, __shfl_sync(0xffffffffu, dnxt, 3), __shfl_sync(0xffffffffu, dnxt, 4), lane, nf1);
        wm_frag(pmat, __shfl_sync(0xffffffffu, dnxt, 6), __shfl_sync(0xffffffffu, dnxt, 7), lane, nf2);
      }
#else
      if (vi + 2 < v_end) dnn = wm_desc_load(visits, vi + 2, lane);
      if (vi > g.first) {
        wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 3), __shfl_sync(0xffffffffu, dcur, 4), lane, bf1);
        wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 6), __shfl_sync(0xffffffffu, dcur, 7), lane, bf2);
      }
#endif
      if (x_in >= 0) {  // first step of a side: a = P_in X from a stored operand
        double x[WM_V], bi[4], dump[WM_V];
        wm_load_raw(x_in, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, sk);
        wm_frag(pmat, p_in, p_in, lane, bi);
        wm_pair(x, bi, ak, dump);
      } else if (x_in <= -2) {
        const int lvl = -2 - x_in;
#pragma unroll
        for (int i = 0; i < WM_V; i++) ak[i] = stk[(lvl * WM_V + i) * 32];
        sk = stks[lvl * WM_NG * 32];
      }
#if PLG_WM_EARLY2 && !PLG_WM_STAGE
      // the second operand's loads are issued before the first one's products: __syncwarp() and the
      // rescale vote keep the compiler from hoisting them itself
      double x2e[WM_V];
      int s2e;
      wm_load_raw(d2, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x2e, s2e);
#endif
      // D1 once: N2 = a o P_t1 D1 (the partial toward d2) and h1 = P_h1 D1 (far side of edge 1)
      double n2[WM_V];
      int s1, sc2;
      {
        double x[WM_V], h1[WM_V];
#if PLG_WM_STAGE
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        wm_unstage(d1, n_rows, lane, stg, x, s1);
#else
        wm_load_raw(d1, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, s1);
#endif
        wm_pair(x, bf1, n2, h1);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < WM_V; i++) { n2[i] *= ak[i]; park[1][i][lane] = h1[i]; }
      }
      sc2 = sk + s1;
      wm_rescale(n2, sc2, lane);
#if PLG_WM_FUSE2 && !PLG_WM_STAGE
      // D2 once, fused with neighbour chain 2: (P_t2 D2, P_h2 D2) and (P_t2 N2, P_h2 N2) element by
      // element -- N1 = a o P_t2 D2 is the only vector this stage leaves behind; h2 = P_h2 D2 is
      // consumed on the spot by neighbour 2's site sums and never exists as a vector.
      double n1[WM_V];
      int s2, sc1;
      {
        double x[WM_V], u[WM_NG], s01[WM_NG], e2[WM_NG];
        wm_load_raw(d2, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, s2);
        double *stk_dst = stk + (int64_t)(push >= 0 ? push : 0) * WM_V * 32;
        const bool do_push = push >= 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int j = 0; j < WM_NG; j++) {
            const int i = r * WM_NG + j;
            double t, h, a, g;
            wm_dmma(t, h, x[i], bf2[r]);
            n1[i] = t * ak[i];
            wm_dmma(a, g, n2[i], bf2[r]);
            if (do_push) stk_dst[i * 32] = a;
            const double tt = g * (h * park[0][i][lane]);
            if (r == 0) s01[j] = tt;
            else if (r == 1) s01[j] += tt;
            else if (r == 2) e2[j] = tt;
            else u[j] = s01[j] + (e2[j] + tt);
          }
        if (push >= 0) stks[push * WM_NG * 32] = sc2;
        if (tree2 >= 0) {
          const double val = wm_site_sum_u(u, sc2 + s2 + st, w_out, pi_q, lane);
          if (lane == 0) partial[(int64_t)tree2 * pstride + slice] = val;
        }
      }
      sc1 = sk + s2;
      wm_rescale(n1, sc1, lane);
#else
      // D2 once: N1 = a o P_t2 D2 and h2 = P_h2 D2
      double n1[WM_V], h2[WM_V];
      int s2, sc1;
      {
        double x[WM_V];
#if PLG_WM_STAGE
        wm_unstage(d2, n_rows, lane, stg + WM_STAGE_ROWS * 32, x, s2);
        if (vi + 1 < v_end) {  // both operands are in registers: fetch the next visit's into the same slots
          wm_stage(__shfl_sync(0xffffffffu, dnxt, 2), n_rows, Ppad, p0, lane, codes, clv, scl, stg_s);
          wm_stage(__shfl_sync(0xffffffffu, dnxt, 5), n_rows, Ppad, p0, lane, codes, clv, scl,
                   stg_s + WM_STAGE_ROWS * 512);
          asm volatile("cp.async.commit_group;\n" ::);
        }
#else
#if PLG_WM_EARLY2
#pragma unroll
        for (int i = 0; i < WM_V; i++) x[i] = x2e[i];
        s2 = s2e;
#else
        wm_load_raw(d2, n_rows, Ppad, p0, lane, out_off, codes, clv, scl, x, s2);
#endif
#if PLG_WM_PF
        if (vi + 1 < v_end) {
          wm_prefetch(__shfl_sync(0xffffffffu, dnxt, 2), n_rows, Ppad, p0, lane, codes, clv, scl);
          wm_prefetch(__shfl_sync(0xffffffffu, dnxt, 5), n_rows, Ppad, p0, lane, codes, clv, scl);
        }
#endif
#endif
        wm_pair(x, bf2, n1, h2);
#pragma unroll
        for (int i = 0; i < WM_V; i++) n1[i] *= ak[i];
#if PLG_WM_PARK_H2
#pragma unroll
        for (int i = 0; i < WM_V; i++) park[2][i][lane] = h2[i];
#endif
      }
      sc1 = sk + s2;
      wm_rescale(n1, sc1, lane);
      // Visits come in normal form (planners): keep is 0 or 1 -- the partial toward c1 continues in
      // registers, the one toward c2 is at most parked.  Side 2 first, so that its hand-over is on
      // the stack (and dead) before side 1's lands directly in ak.
      {  // side 2: toward c2 (at most parked) and neighbour 2 (insertion into x--c2)
        double u[WM_NG], dummy[WM_V];
#if PLG_WM_PARK_H2
#pragma unroll
        for (int i = 0; i < WM_V; i++) h2[i] = park[2][i][lane];
#endif
        wm_chain<false>(n2, bf2, h2, park[0], lane, dummy, stk + (int64_t)(push >= 0 ? push : 0) * WM_V * 32, push >= 0, u);
        if (push >= 0) stks[push * WM_NG * 32] = sc2;
        if (tree2 >= 0) {
          const double val = wm_site_sum_u(u, sc2 + s2 + st, w_out, pi_q, lane);
          if (lane == 0) partial[(int64_t)tree2 * pstride + slice] = val;
        }
      }
#endif
      {  // side 1: toward c1 (straight into ak) and neighbour 1 (insertion into x--c1)
        double u[WM_NG], h1[WM_V];
#if PLG_WM_REFRAG
        // bf1 is not kept across the D2 stage and chain 2: fetched again (L1) where chain 1 needs it
        wm_frag(pmat, __shfl_sync(0xffffffffu, dcur, 3), __shfl_sync(0xffffffffu, dcur, 4), lane, bf1);
#endif
#pragma unroll
        for (int i = 0; i < WM_V; i++) h1[i] = park[1][i][lane];
        wm_chain<true>(n1, bf1, h1, park[0], lane, ak, nullptr, false, u);
        sk = sc1;
        if (tree1 >= 0) {
          const double val = wm_site_sum_u(u, sc1 + s1 + st, w_out, pi_q, lane);
          if (lane == 0) partial[(int64_t)tree1 * pstride + slice] = val;
        }
      }
      dcur = dnxt; dnxt = dnn;
#if PLG_WM_PREFETCH
#pragma unroll
      for (int r = 0; r < 4; r++) { bf1[r] = nf1[r]; bf2[r] = nf2[r]; }
#endif
    }
  }
}

}  // namespace plg
