// Fitch tree walk: one thread carries one VW-word column (128 patterns for DNA) of all K planes
// through the WHOLE post-order of a tree (treewalk.h).  The running state set lives in registers,
// parked sets in the thread's local-memory stack (L1/L2), so HBM/L2 only supplies the tips:
// the kernel is bound by the integer pipes (about 25 bit-ops per node-op per 32 patterns), not by
// the 3*K/8 bytes per node-op that the level-synchronous kernel has to move (SURVEY.md 8d i).
//
// Semantics per node (binary; fitch.cpp:215-226, 262-264): X = a & b; where X is empty the node
// takes a | b and the pattern costs its weight.  Integer sums only => bit-exact, any order.
#pragma once

namespace plg {

template <int K, int VW>
__device__ __forceinline__ void fw_load_tip(const unsigned int *__restrict__ tips, int row, int64_t vec_words,
                                            int64_t words32, int64_t w0, unsigned int (&x)[K][VW]) {
  const unsigned int *src = tips + (int64_t)row * vec_words + w0;
#pragma unroll
  for (int s = 0; s < K; s++) load_vec<VW>(src + s * words32, x[s]);
}

template <int K, int VW>
__global__ void __launch_bounds__(256)
fitch_walk_kernel(const unsigned int *__restrict__ tips, const WalkOp *__restrict__ ops, int ni, int64_t words32,
                  int blocks_per_tree, const unsigned int *__restrict__ wbits,
                  const unsigned char *__restrict__ chunk_nb, unsigned int *__restrict__ cost) {
  const int tree = blockIdx.x / blocks_per_tree;
  const int cb = blockIdx.x - tree * blocks_per_tree;
  const int64_t col = (int64_t)cb * blockDim.x + threadIdx.x;
  const int64_t w0 = col * VW;
  const int64_t vec_words = (int64_t)K * words32;
  unsigned int total = 0;
  if (w0 < words32) {
    unsigned int stack[WALK_MAX_LEVELS][K][VW];
    unsigned int top[K][VW];
#pragma unroll
    for (int s = 0; s < K; s++)
#pragma unroll
      for (int v = 0; v < VW; v++) top[s][v] = 0;
    const unsigned int nb = chunk_nb[w0 >> 2];
    const int4 *op_p = reinterpret_cast<const int4 *>(ops + (int64_t)tree * ni);
    int4 nxt = __ldg(op_p);
    for (int i = 0; i < ni; i++) {
      const int4 op = nxt;
      if (i + 1 < ni) nxt = __ldg(op_p + i + 1);
      unsigned int xa[K][VW], xb[K][VW];
      if (op.x >= 0) {
        fw_load_tip<K, VW>(tips, op.x, vec_words, words32, w0, xa);
      } else if (op.x == -1) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xa[s][v] = top[s][v];
      } else {
        const int l = -2 - op.x;
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xa[s][v] = stack[l][s][v];
      }
      if (op.y >= 0) {
        fw_load_tip<K, VW>(tips, op.y, vec_words, words32, w0, xb);
      } else if (op.y == -1) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xb[s][v] = top[s][v];
      } else {
        const int l = -2 - op.y;
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xb[s][v] = stack[l][s][v];
      }
      unsigned int miss[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) {
        unsigned int any = 0;
#pragma unroll
        for (int s = 0; s < K; s++) any |= xa[s][v] & xb[s][v];
        miss[v] = ~any;
      }
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) xa[s][v] = (xa[s][v] & xb[s][v]) | (miss[v] & (xa[s][v] | xb[s][v]));
      if (op.z >= 0) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) stack[op.z][s][v] = xa[s][v];
      } else {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) top[s][v] = xa[s][v];
      }
      if (nb == 0xFFu) {
#pragma unroll
        for (int v = 0; v < VW; v++) total += __popc(miss[v]);
      } else {
        for (unsigned int b = 0; b < nb; b++) {
          unsigned int wv[VW];
          load_vec<VW>(wbits + (int64_t)b * words32 + w0, wv);
          unsigned int c = 0;
#pragma unroll
          for (int v = 0; v < VW; v++) c += __popc(miss[v] & wv[v]);
          total += c << b;
        }
      }
    }
  }
  total = __reduce_add_sync(0xffffffffu, total);
  __shared__ unsigned int warp_sums[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = total;
  __syncthreads();
  if (warp == 0) {
    unsigned int v = lane < (blockDim.x + 31) / 32 ? warp_sums[lane] : 0;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0 && v) atomicAdd(cost + tree, v);
  }
}

template <int K, int VW>
static void launch_fitch_walk(const FitchAln &a, const WalkOp *d_ops, int64_t n_trees, int ni, unsigned int *d_cost,
                              cudaStream_t st) {
  const int64_t cols = a.words32 / VW;
  int best_t = 256;
  int64_t best_waste = INT64_MAX;
  for (int t = 256; t >= 64; t -= 32) {
    const int64_t blocks = (cols + t - 1) / t, waste = blocks * t - cols;
    if (waste < best_waste) { best_waste = waste; best_t = t; }
  }
  const int64_t bpt = (cols + best_t - 1) / best_t;
  const int64_t max_trees = std::max<int64_t>(1, ((int64_t)1 << 30) / bpt);
  for (int64_t t0 = 0; t0 < n_trees; t0 += max_trees) {
    const int64_t cnt = std::min(max_trees, n_trees - t0);
    prof_begin(PROF_FITCH_OPS, st);
    fitch_walk_kernel<K, VW><<<(unsigned)(cnt * bpt), best_t, 0, st>>>(a.tips.p, d_ops + t0 * ni, ni, a.words32,
                                                                      (int)bpt, a.wbits.p, a.chunk_nb.p, d_cost + t0);
    count_launch();
    // algorithmic bytes: every node-op moves 3 vectors of K_real planes; requested:
    // the tips (n per tree) -- everything else stays on chip
    const double vec = a.K_real * (double)a.P / 8.0;
    prof_end(PROF_FITCH_OPS, st, (double)cnt * ni * 3.0 * vec, (double)cnt * ni * (double)a.P,
             (double)cnt * (ni + 1) * vec);
  }
}


// ---------------------------------------------------------------------------------------------
// SPR search walk (neighbourhood scoring by edge insertion): a thread carries one column through a
// whole search -- every regraft edge of one pruned subtree, depth first.  At a node x reached with
// the partial X of the reduced tree, with x's other neighbours c1, c2 (directional sets D1, D2):
//     N1 = F(X, D2)   neighbour 1 (insertion into x--c1): patterns where F(N1, D1) misses Tv
//     N2 = F(X, D1)   neighbour 2 (insertion into x--c2): patterns where F(N2, D2) misses Tv
// F = Fitch's binary node rule (intersection, union where empty).  The partial that continues into
// a child stays in registers, the other one waits on the thread's local-memory stack; D1 and D2
// are read once for two neighbours and no partial ever goes to HBM (the level-synchronous ops
// moved 5 vectors per neighbour).
template <int K, int VW>
__device__ __forceinline__ void fw_node(const unsigned int (&a)[K][VW], const unsigned int (&b)[K][VW],
                                        unsigned int (&r)[K][VW]) {
#pragma unroll
  for (int v = 0; v < VW; v++) {
    unsigned int any = 0;
#pragma unroll
    for (int s = 0; s < K; s++) any |= a[s][v] & b[s][v];
#pragma unroll
    for (int s = 0; s < K; s++) r[s][v] = (a[s][v] & b[s][v]) | (~any & (a[s][v] | b[s][v]));
  }
}

template <int K, int VW>
__device__ __forceinline__ void fw_load(const unsigned int *__restrict__ tips, const unsigned int *__restrict__ scratch,
                                        int n_tip_slots, int id, int64_t vec_words, int64_t words32, int64_t w0,
                                        unsigned int (&x)[K][VW]) {
  const unsigned int *src = (id < n_tip_slots ? tips + (int64_t)id * vec_words
                                              : scratch + (int64_t)(id - n_tip_slots) * vec_words) + w0;
#pragma unroll
  for (int s = 0; s < K; s++) load_vec<VW>(src + s * words32, x[s]);
}

// weighted count of the patterns where the Fitch join of (n, d) shares no state with tv
template <int K, int VW>
__device__ __forceinline__ unsigned int fw_join_miss(const unsigned int (&n)[K][VW], const unsigned int (&d)[K][VW],
                                                     const unsigned int (&tv)[K][VW], unsigned int nb,
                                                     const unsigned int *__restrict__ wbits, int64_t words32,
                                                     int64_t w0) {
  unsigned int miss[VW];
#pragma unroll
  for (int v = 0; v < VW; v++) {
    unsigned int any = 0, hit = 0, hit_u = 0;
#pragma unroll
    for (int s = 0; s < K; s++) {
      any |= n[s][v] & d[s][v];
      hit |= n[s][v] & d[s][v] & tv[s][v];
      hit_u |= (n[s][v] | d[s][v]) & tv[s][v];
    }
    miss[v] = ~(hit | (~any & hit_u));
  }
  unsigned int c = 0;
  if (nb == 0xFFu) {
#pragma unroll
    for (int v = 0; v < VW; v++) c += __popc(miss[v]);
  } else {
    for (unsigned int b = 0; b < nb; b++) {
      unsigned int wv[VW];
      load_vec<VW>(wbits + (int64_t)b * words32 + w0, wv);
      unsigned int cb = 0;
#pragma unroll
      for (int v = 0; v < VW; v++) cb += __popc(miss[v] & wv[v]);
      c += cb << b;
    }
  }
  return c;
}

#ifndef PLG_FSW_VW_DNA
#define PLG_FSW_VW_DNA 4  // 32-bit words per plane a thread carries through a 4-plane search walk (4 = 128 patterns; 2: 2.79 vs 2.59 ms per 256x100k step at best)
#endif
#ifndef PLG_FSW_MINB
#define PLG_FSW_MINB 2  // 120 registers, no spills (1: 147 registers, 3.2 vs 2.66 ms per 256x100k step; 3: 80 registers with spills, 3.6)
#endif
template <int K, int VW>
__global__ void __launch_bounds__(256, PLG_FSW_MINB)
fitch_search_walk_kernel(const unsigned int *__restrict__ tips, const unsigned int *__restrict__ scratch,
                         int n_tip_slots, const FitchWalkGroup *__restrict__ groups,
                         const FitchWalkVisit *__restrict__ visits, int64_t words32, int blocks_per_group,
                         const unsigned int *__restrict__ wbits, const unsigned char *__restrict__ chunk_nb,
                         unsigned int *__restrict__ cost) {
  const int gi = blockIdx.x / blocks_per_group;
  const int cb = blockIdx.x - gi * blocks_per_group;
  const int64_t col = (int64_t)cb * blockDim.x + threadIdx.x;
  const bool live = col * VW < words32;
  const int64_t w0 = live ? col * VW : 0;  // idle lanes of the last block shadow column 0 and add nothing
  const int64_t vec_words = (int64_t)K * words32;
  const int4 gq = __ldg(reinterpret_cast<const int4 *>(groups + gi));
  const unsigned int nb = chunk_nb[w0 >> 2];
  unsigned int tv[K][VW], xk[K][VW];
  unsigned int stack[WALK_MAX_LEVELS][K][VW];
  fw_load<K, VW>(tips, scratch, n_tip_slots, gq.x, vec_words, words32, w0, tv);
#pragma unroll
  for (int s = 0; s < K; s++)
#pragma unroll
    for (int v = 0; v < VW; v++) xk[s][v] = 0;
  const int4 *vp = reinterpret_cast<const int4 *>(visits + gq.y);
  int4 q0 = __ldg(vp), q1 = __ldg(vp + 1);
  for (int vi = 0; vi < gq.z; vi++) {
    const int x_in = q0.x, d1 = q0.y, d2 = q0.z, tree1 = q0.w, tree2 = q1.x, keep = q1.y, push = q1.z;
    if (vi + 1 < gq.z) { q0 = __ldg(vp + 2 * vi + 2); q1 = __ldg(vp + 2 * vi + 3); }
    unsigned int D1[K][VW], D2[K][VW];
    fw_load<K, VW>(tips, scratch, n_tip_slots, d1, vec_words, words32, w0, D1);
    fw_load<K, VW>(tips, scratch, n_tip_slots, d2, vec_words, words32, w0, D2);
    if (x_in >= 0) {
      fw_load<K, VW>(tips, scratch, n_tip_slots, x_in, vec_words, words32, w0, xk);
    } else if (x_in <= -2) {
      const int l = -2 - x_in;
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) xk[s][v] = stack[l][s][v];
    }
    unsigned int n1[K][VW], n2[K][VW];
    fw_node<K, VW>(xk, D2, n1);
    fw_node<K, VW>(xk, D1, n2);
    if (tree1 >= 0) {
      unsigned int c = fw_join_miss<K, VW>(n1, D1, tv, nb, wbits, words32, w0);
      c = __reduce_add_sync(0xffffffffu, live ? c : 0u);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(cost + tree1, c);
    }
    if (tree2 >= 0) {
      unsigned int c = fw_join_miss<K, VW>(n2, D2, tv, nb, wbits, words32, w0);
      c = __reduce_add_sync(0xffffffffu, live ? c : 0u);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(cost + tree2, c);
    }
    if (push >= 0) {
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) stack[push][s][v] = keep == 1 ? n2[s][v] : n1[s][v];
    }
    if (keep == 1) {
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) xk[s][v] = n1[s][v];
    } else if (keep == 2) {
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) xk[s][v] = n2[s][v];
    }
  }
}

template <int K, int VW>
static void launch_fitch_search_walk(const FitchAln &a, const unsigned int *scratch, const FitchWalkVisit *d_visits,
                                     const FitchWalkGroup *d_groups, int n_groups, unsigned int *d_cost,
                                     cudaStream_t st) {
  const int64_t cols = a.words32 / VW;
  int best_t = 256;
  int64_t best_waste = INT64_MAX;
  for (int t = 256; t >= 64; t -= 32) {
    const int64_t blocks = (cols + t - 1) / t, waste = blocks * t - cols;
    if (waste < best_waste) { best_waste = waste; best_t = t; }
  }
  const int64_t bpg = (cols + best_t - 1) / best_t;
  fitch_search_walk_kernel<K, VW><<<(unsigned)(n_groups * bpg), best_t, 0, st>>>(
      a.tips.p, scratch, a.n_cols, d_groups, d_visits, a.words32, (int)bpg, a.wbits.p, a.chunk_nb.p, d_cost);
}

}  // namespace plg
