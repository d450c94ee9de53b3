// Fitch small parsimony on sm_100a: bit-plane state sets, level-synchronous op programs.
//
// Replaces cost() and the string-set helpers of the reference
// (/root/reference/pylogeny/fitch.cpp:197-272).  Data layout in HBM:
//
//   state-set vector of one node = K bit-planes over the P patterns,
//   [plane][W4] uint4, W4 = ceil(P/128) (every plane is a whole number of 128-bit
//   words so all accesses are 16-byte vectorised and coalesced).  Bit i of plane s
//   is set iff state s is in the node's set at pattern i.  Tips (one vector per
//   alignment row) are read-only; internal vectors live in a scratch pool.
//
//   Patterns are re-ordered by weight (descending).  Weights are bit-sliced:
//   wbits[b][W4] holds bit b of every pattern weight, so a weighted miss count is
//   sum_b popc(miss & wbits[b]) << b -- integer, order-independent, hence
//   bit-exact with fitch.cpp:268 including its 32-bit wrap-around.  128-pattern
//   chunks whose weights are all 1 (the long tail after sorting) skip the loads.
//
// One "op" = one internal node: fold over its children with the reference's
// restarting intersection (fitch.cpp:215-226), union on empty (:262-264), cost
// (k-1) per missed pattern.  Ops of the same level are independent and run in one
// launch; a thread owns one 128/64/32-pattern column of all K planes.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "fitch.h"
#include "hostpool.h"
#include "treewalk.h"

namespace plg {

// ----------------------------------------------------------------------------
// device
// ----------------------------------------------------------------------------
template <int VW>
struct VecT;
template <>
struct VecT<4> { using type = uint4; };
template <>
struct VecT<2> { using type = uint2; };
template <>
struct VecT<1> { using type = unsigned int; };

template <int VW>
__device__ __forceinline__ void load_vec(const unsigned int *p, unsigned int (&v)[VW]) {
  if constexpr (VW == 4) {
    uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VW == 2) {
    uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

template <int VW>
__device__ __forceinline__ void store_vec(unsigned int *p, const unsigned int (&v)[VW]) {
  if constexpr (VW == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]);
  else if constexpr (VW == 2) *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
  else p[0] = v[0];
}

// words32: 32-bit words per plane (= 4*W4).  cols_per_op = words32 / VW.
#ifndef PLG_FITCH_MINB
#define PLG_FITCH_MINB 4  // residency sweep in profiles/tune_fitch.sh: 4 is best (6.8 ms vs 8.0 for C3)
#endif
template <int K, int VW>
__global__ void __launch_bounds__(256, PLG_FITCH_MINB)
fitch_level_kernel(const unsigned int *__restrict__ tips, unsigned int *__restrict__ scratch,
                   int n_tip_slots, const FitchOp *__restrict__ ops,
                   const int *__restrict__ children, int64_t words32, int blocks_per_op,
                   const unsigned int *__restrict__ wbits, const unsigned char *__restrict__ chunk_nb,
                   unsigned int *__restrict__ cost) {
  const int op_idx = blockIdx.x / blocks_per_op;
  const int cb = blockIdx.x - op_idx * blocks_per_op;
  const FitchOp op = ops[op_idx];
  const int64_t col = (int64_t)cb * blockDim.x + threadIdx.x;  // VW-word column
  const int64_t w0 = col * VW;                                   // first 32-bit word
  const int64_t vec_words = (int64_t)K * words32;
  unsigned int partial = 0;
  if (w0 < words32) {
    unsigned int acc[K][VW], orr[K][VW];
    {
      const int c0 = children[op.child_off];
      const unsigned int *src = (c0 < n_tip_slots ? tips + (int64_t)c0 * vec_words
                                                  : scratch + (int64_t)(c0 - n_tip_slots) * vec_words) + w0;
#pragma unroll
      for (int s = 0; s < K; s++) {
        load_vec<VW>(src + s * words32, acc[s]);
#pragma unroll
        for (int v = 0; v < VW; v++) orr[s][v] = acc[s][v];
      }
    }
    const int nfold = op.nchild < 0 ? 2 : op.nchild;
    for (int j = 1; j < nfold; j++) {
      const int cj = children[op.child_off + j];
      const unsigned int *src = (cj < n_tip_slots ? tips + (int64_t)cj * vec_words
                                                  : scratch + (int64_t)(cj - n_tip_slots) * vec_words) + w0;
      unsigned int x[K][VW];
#pragma unroll
      for (int s = 0; s < K; s++) load_vec<VW>(src + s * words32, x[s]);
      unsigned int empty[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) {
        unsigned int any = 0;
#pragma unroll
        for (int s = 0; s < K; s++) any |= acc[s][v];
        empty[v] = ~any;
      }
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) {
          // running intersection that restarts from the child when empty (fitch.cpp:221-222)
          acc[s][v] = (acc[s][v] | empty[v]) & x[s][v];
          orr[s][v] |= x[s][v];
        }
    }
    if (op.nchild == FITCH_JOIN3) {
      // edge-insertion test: state set of the first two operands joined (Fitch), intersected
      // with the third; only the miss count is kept
      const int c2 = children[op.child_off + 2];
      const unsigned int *src = (c2 < n_tip_slots ? tips + (int64_t)c2 * vec_words
                                                  : scratch + (int64_t)(c2 - n_tip_slots) * vec_words) + w0;
      unsigned int any[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) {
        any[v] = 0;
#pragma unroll
        for (int s = 0; s < K; s++) any[v] |= acc[s][v];
      }
#pragma unroll
      for (int s = 0; s < K; s++) {
        unsigned int x[VW];
        load_vec<VW>(src + s * words32, x);
#pragma unroll
        for (int v = 0; v < VW; v++) acc[s][v] = (acc[s][v] | (~any[v] & orr[s][v])) & x[v];
      }
    }
    unsigned int miss[VW];
#pragma unroll
    for (int v = 0; v < VW; v++) {
      unsigned int any = 0;
#pragma unroll
      for (int s = 0; s < K; s++) any |= acc[s][v];
      miss[v] = ~any;
    }
    if (op.dst >= 0) {
      unsigned int *dst = scratch + (int64_t)(op.dst - n_tip_slots) * vec_words + w0;
#pragma unroll
      for (int s = 0; s < K; s++) {
#pragma unroll
        for (int v = 0; v < VW; v++) acc[s][v] |= miss[v] & orr[s][v];  // union on empty (:262-263)
        store_vec<VW>(dst + s * words32, acc[s]);
      }
    } else if (op.nchild == FITCH_FUSED) {
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) acc[s][v] |= miss[v] & orr[s][v];
    }
    if (op.nchild == FITCH_FUSED) {
      // edge insertion in one pass: acc now holds N; join it with the far side of the target
      // edge (Fitch), intersect with the pruned subtree's set, keep only that miss count
      const int c2 = children[op.child_off + 2], c3 = children[op.child_off + 3];
      const unsigned int *f = (c2 < n_tip_slots ? tips + (int64_t)c2 * vec_words
                                                : scratch + (int64_t)(c2 - n_tip_slots) * vec_words) + w0;
      const unsigned int *tv = (c3 < n_tip_slots ? tips + (int64_t)c3 * vec_words
                                                 : scratch + (int64_t)(c3 - n_tip_slots) * vec_words) + w0;
      unsigned int any[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) any[v] = 0;
#pragma unroll
      for (int s = 0; s < K; s++) {
        unsigned int x[VW];
        load_vec<VW>(f + s * words32, x);
#pragma unroll
        for (int v = 0; v < VW; v++) {
          orr[s][v] = acc[s][v] | x[v];
          acc[s][v] &= x[v];
          any[v] |= acc[s][v];
        }
      }
      unsigned int hit[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) hit[v] = 0;
#pragma unroll
      for (int s = 0; s < K; s++) {
        unsigned int x[VW];
        load_vec<VW>(tv + s * words32, x);
#pragma unroll
        for (int v = 0; v < VW; v++) hit[v] |= (acc[s][v] | (~any[v] & orr[s][v])) & x[v];
      }
#pragma unroll
      for (int v = 0; v < VW; v++) miss[v] = ~hit[v];
    }
    // weighted miss count
    const unsigned int nb = chunk_nb[w0 >> 2];
    if (nb == 0xFFu) {
#pragma unroll
      for (int v = 0; v < VW; v++) partial += __popc(miss[v]);
    } else {
      for (unsigned int b = 0; b < nb; b++) {
        unsigned int wv[VW];
        load_vec<VW>(wbits + (int64_t)b * words32 + w0, wv);
        unsigned int c = 0;
#pragma unroll
        for (int v = 0; v < VW; v++) c += __popc(miss[v] & wv[v]);
        partial += c << b;
      }
    }
    if (op.nchild > 2) partial *= (unsigned int)(op.nchild - 1);  // numChildren-1 per missed pattern (:264)
    if (op.nchild == 1) partial = 0;
  }
  // block reduction -> one integer atomic per block (order-independent => deterministic)
  partial = __reduce_add_sync(0xffffffffu, partial);
  __shared__ unsigned int warp_sums[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = partial;
  __syncthreads();
  if (warp == 0) {
    unsigned int v = lane < (blockDim.x + 31) / 32 ? warp_sums[lane] : 0;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0 && v) atomicAdd(cost + op.tree, v);
  }
}

// Packs a dense [P][n_cols] symbol matrix (permuted by weight through `order`) into bit-planes on the device.
// A warp owns one 32-pattern word; a lane walks ITS pattern's column of symbols (contiguous bytes) once, relabels
// them by first appearance (a 256-byte table per lane) and the warp turns each row's 32 labels into K words with
// K ballots.  (The first version gave a thread one (row, word): 32 strided byte reads and a rescan of the rows
// above for every symbol -- 1.1 ms at 256 x 100k, more than the upload.)
__global__ void fitch_pack_kernel(const unsigned char *__restrict__ states, const int *__restrict__ order,
                                  int64_t P, int n_cols, int K, int64_t words32,
                                  unsigned int *__restrict__ tips) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= words32) return;
  const int64_t p = w * 32 + lane;
  const bool real = p < P;  // padding patterns: every row in state 0 -> never a miss
  const unsigned char *colp = states + (real ? (int64_t)order[p] * n_cols : 0);
  unsigned char label[256];  // symbol -> 1 + rank of first appearance among the rows (0: not seen yet)
  for (int c = 0; c < 256; c++) label[c] = 0;
  int n_seen = 0;
  for (int row = 0; row < n_cols; row++) {
    int st = 0;
    if (real) {
      const unsigned char c = colp[row];
      int t = label[c];
      if (!t) { t = ++n_seen; label[c] = (unsigned char)t; }
      st = t - 1;
    }
    for (int s = 0; s < K; s++) {
      const unsigned int m = __ballot_sync(0xffffffffu, st == s);
      if (lane == (s & 31)) tips[((int64_t)row * K + s) * words32 + w] = m;
    }
  }
}


}  // namespace plg
#include "fitch_walk.cuh"
namespace plg {

// ----------------------------------------------------------------------------
// host: alignment
// ----------------------------------------------------------------------------
static int pick_template_planes(int K) {
  static const int opts[] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20};
  for (int o : opts)
    if (K <= o) return o;
  return -1;
}

// Largest number of distinct symbols in a pattern (= the number of planes): one warp per pattern, a 256-bit
// set per lane, OR-reduced over the warp.  (On the host this was half of what creating a handle cost at
// 256 x 100k, even on all cores.)
__global__ void fitch_kmax_kernel(const unsigned char *__restrict__ states, int64_t P, int n_cols, int *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  int best = 0;
  for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < P; p += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    unsigned int m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned char *colp = states + p * n_cols;
    for (int r = lane; r < n_cols; r += 32) {
      const unsigned int c = colp[r], bit = 1u << (c & 31), q = c >> 5;
#pragma unroll
      for (int i = 0; i < 8; i++) m[i] |= q == (unsigned int)i ? bit : 0u;
    }
    int k = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) k += __popc(__reduce_or_sync(0xffffffffu, m[i]));
    best = max(best, k);
  }
  if (lane == 0 && best) atomicMax(out, best);
}

FitchAln::FitchAln(int64_t n_prof, int n_cols_, const uint8_t *states, const int64_t *weights) {
  StageTimer tm;
  P = n_prof;
  n_cols = n_cols_;
  if (P < 0 || n_cols <= 0) PLG_FAIL(PLG_ERR_ARG, "alignment needs n_prof >= 0 and n_cols > 0");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  // distinct symbols per pattern -> number of planes: the symbols go up first (the pack kernel needs them
  // anyway) and are counted on the device
  DevBuf<unsigned char> d_states;
  d_states.alloc(std::max<size_t>(1, (size_t)P * n_cols));
  int kmax = 1;
  if (P) {
    PLG_CUDA(cudaMemcpy(d_states.p, states, (size_t)P * n_cols, cudaMemcpyHostToDevice));
    DevBuf<int> d_k;
    d_k.alloc(1);
    PLG_CUDA(cudaMemset(d_k.p, 0, sizeof(int)));
    fitch_kmax_kernel<<<(unsigned)std::min<int64_t>((P + 7) / 8, 148 * 16), 256>>>(d_states.p, P, n_cols, d_k.p);
    count_launch();
    PLG_CUDA(cudaGetLastError());
    PLG_CUDA(cudaMemcpy(&kmax, d_k.p, sizeof(int), cudaMemcpyDeviceToHost));
    kmax = std::max(kmax, 1);
  }
  tm.lap("aln states");
  K_real = kmax;
  K = pick_template_planes(kmax);
  if (K < 0)
    PLG_FAIL(PLG_ERR_STATES, "a pattern has %d distinct states; at most %d are supported", kmax,
             FITCH_MAX_PLANES);
  W4 = std::max<int64_t>(1, (P + 127) / 128);
  words32 = W4 * 4;
  // weight order (descending by low 32 bits: the cost is taken mod 2^32, fitch.cpp:247,268)
  std::vector<int> order(P);
  std::iota(order.begin(), order.end(), 0);
  std::vector<uint32_t> w32(P);
  for (int64_t p = 0; p < P; p++) w32[p] = (uint32_t)(uint64_t)weights[p];
  bool sorted = true;  // (equal weights -- distinct patterns of a long alignment -- are in order already)
  for (int64_t p = 1; p < P && sorted; p++) sorted = w32[p] <= w32[p - 1];
  if (!sorted) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w32[a] > w32[b]; });
  // bit-sliced weights + per-128-chunk plane count
  std::vector<uint8_t> nb(W4, 0xFF);
  int B = 0;
  for (int64_t c = 0; c < W4; c++) {
    uint32_t mx = 0;
    bool unit = true;
    for (int64_t p = c * 128; p < std::min<int64_t>(P, c * 128 + 128); p++) {
      uint32_t w = w32[order[p]];
      mx = std::max(mx, w);
      unit = unit && w == 1;
    }
    if (!unit) {
      int bits = 0;
      while (bits < 32 && (mx >> bits)) bits++;
      nb[c] = (uint8_t)bits;
      B = std::max(B, bits);
    }
  }
  n_wbits = B;
  std::vector<uint32_t> wb((size_t)std::max(B, 1) * words32, 0);
  for (int64_t p = 0; p < P; p++) {
    uint32_t w = w32[order[p]];
    for (int b = 0; b < B; b++)
      if (w >> b & 1) wb[(size_t)b * words32 + (p >> 5)] |= 1u << (p & 31);
  }
  tm.lap("aln weights");
  tips.alloc((size_t)n_cols * K * words32);
  wbits.alloc(wb.size());
  chunk_nb.alloc(W4);
  PLG_CUDA(cudaMemcpy(wbits.p, wb.data(), wb.size() * 4, cudaMemcpyHostToDevice));
  PLG_CUDA(cudaMemcpy(chunk_nb.p, nb.data(), W4, cudaMemcpyHostToDevice));
  // pack on the device
  DevBuf<int> d_order;
  d_order.alloc(std::max<size_t>(1, P));
  if (P) PLG_CUDA(cudaMemcpy(d_order.p, order.data(), (size_t)P * sizeof(int), cudaMemcpyHostToDevice));
  tm.lap("aln upload");
  fitch_pack_kernel<<<(unsigned)((words32 + 3) / 4), 128>>>(d_states.p, d_order.p, P, n_cols, K, words32, tips.p);
  count_launch();
  PLG_CUDA(cudaGetLastError());
  PLG_CUDA(cudaDeviceSynchronize());
  tm.lap("aln pack");
}

// ----------------------------------------------------------------------------
// host: program construction and execution
// ----------------------------------------------------------------------------
FitchWorkspace &fitch_workspace() {
  static FitchWorkspace *ws[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ws[dev & 63]) ws[dev & 63] = new FitchWorkspace();  // intentionally never freed (process lifetime)
  return *ws[dev & 63];
}

void FitchProgram::clear() {
  ops.clear();
  children.clear();
  level_off.clear();
  walk_visits.clear();
  walk_groups.clear();
  spr_edges.clear();
  spr_searches.clear();
  spr_n_visits = spr_n_unique = 0;
  spr_lo = spr_hi = 0;
  spr_first_acc = 0;
  spr_vecs = spr_req_vecs = 0;
  n_scratch_slots = 0;
  n_costs = 0;
}

// Lays out the search walks of a full SPR neighbourhood on the device: one thread per search
// (pruned subtree, side) runs the depth-first traversal the host planner
// (neighbourhood.cpp build_fitch_nb_walk_program) would, and writes the FitchWalkVisit list the
// walk kernel executes next on the same stream.  A visit's two moves get the unique ids of the
// host enumeration (enumerate_spr): moves are numbered in pre-order with the lower neighbour
// slot first, two per visit (higher slot's move first); only the first two moves of a search can
// repeat an earlier topology (dup bits).  The walk itself descends into the smaller subtree first
// and parks the other partial, so the stack never exceeds log2(n) + 1 frames.
constexpr size_t SPR_PLAN_SMEM_MAX = 200 << 10;  // edge records of trees up to ~4,000 (parsimony) / ~2,000 (likelihood) taxa
__global__ void __launch_bounds__(64)
spr_plan_kernel(const int4 *__restrict__ edges_g, int n_edge_recs, const int4 *__restrict__ searches, int n_search,
                int first_acc, int id_lo, int id_hi, int4 *__restrict__ visits, int4 *__restrict__ out_moves) {
  // a search is one long chain of dependent edge-record lookups: keep the records in shared
  // memory when they fit (n_edge_recs > 0), 155 -> 30 us per 256-taxon neighbourhood
  extern __shared__ int4 spr_edges_s[];
  const int4 *edges = edges_g;
  if (n_edge_recs > 0) {
    for (int i = threadIdx.x; i < n_edge_recs; i += blockDim.x) spr_edges_s[i] = __ldg(edges_g + i);
    __syncthreads();
    edges = spr_edges_s;
  }
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_search) return;
  const int4 r0 = __ldg(searches + 2 * s), r1 = __ldg(searches + 2 * s + 1);
  const int move_cnt = r1.y;
  if (move_cnt == 0) return;
  const int uniq_base = r1.x, dup = r1.z & 3, complement = (r1.z >> 2) & 1, pe = r1.w;
  const int dup0 = dup & 1, dup01 = dup0 + (dup >> 1);
  constexpr int LV = WALK_MAX_LEVELS + 2;
  int sx[LV], sfrom[LV], spre[LV], sin[LV], sdep[LV];
  int sp = 0, park = 0;
  int64_t at = r0.w;
  sx[0] = r0.x; sfrom[0] = r0.y; spre[0] = 0; sin[0] = r0.z; sdep[0] = 1;
  sp = 1;
  while (sp > 0) {
    sp--;
    const int x = sx[sp], from = sfrom[sp], pre = spre[sp], x_in = sin[sp], dep = sdep[sp];
    if (x_in <= -2) park--;
    const int4 e0 = edges[3 * x], e1 = edges[3 * x + 1], e2 = edges[3 * x + 2];
    // the two neighbours other than `from`: E1 = higher slot (move 2*pre), E2 = lower slot (2*pre + 1)
    const int4 E1 = e2.x == from ? e1 : e2;
    const int4 E2 = e0.x == from ? e1 : e0;
    const int i1 = 2 * pre, i2 = i1 + 1;
    int uid1, uid2;
    if (pre == 0) {
      uid1 = dup0 ? -1 : uniq_base;
      uid2 = (dup >> 1) ? -1 : uniq_base + 1 - dup0;
    } else {
      uid1 = uniq_base + i1 - dup01;
      uid2 = uniq_base + i2 - dup01;
    }
    const int sub1 = E1.z, sub2 = E2.z;  // visits below: 0 = the neighbour is a tip
    const int pre2 = pre + 1, pre1 = pre + 1 + sub2;
    int keep = 0, push = -1;
    if (sub1 > 0 && sub2 > 0) {
      const bool first1 = sub1 <= sub2;
      keep = first1 ? 1 : 2;
      push = park;
      sx[sp] = first1 ? E2.x : E1.x; sfrom[sp] = x; spre[sp] = first1 ? pre2 : pre1; sin[sp] = -2 - park; sdep[sp] = dep + 1;
      sp++;
      sx[sp] = first1 ? E1.x : E2.x; sfrom[sp] = x; spre[sp] = first1 ? pre1 : pre2; sin[sp] = -1; sdep[sp] = dep + 1;
      sp++;
      park++;
    } else if (sub1 > 0) {
      keep = 1;
      sx[sp] = E1.x; sfrom[sp] = x; spre[sp] = pre1; sin[sp] = -1; sdep[sp] = dep + 1;
      sp++;
    } else if (sub2 > 0) {
      keep = 2;
      sx[sp] = E2.x; sfrom[sp] = x; spre[sp] = pre2; sin[sp] = -1; sdep[sp] = dep + 1;
      sp++;
    }
    // accumulators exist for the scored range of unique ids only (a shard scores [id_lo, id_hi))
    const int t1 = uid1 >= id_lo && uid1 < id_hi ? first_acc + (uid1 - id_lo) : -1;
    const int t2 = uid2 >= id_lo && uid2 < id_hi ? first_acc + (uid2 - id_lo) : -1;
    visits[2 * at] = make_int4(x_in, E1.y, E2.y, t1);
    visits[2 * at + 1] = make_int4(t2, keep, push, 0);
    at++;
    if (out_moves) {
      if (uid1 >= 0) out_moves[uid1] = make_int4(pe, complement, E1.w, dep);
      if (uid2 >= 0) out_moves[uid2] = make_int4(pe, complement, E2.w, dep);
    }
  }
}

// Full re-traversal of trees [t0, t1) of a batch: every internal node is one op.
void build_full_program(const TreeBatch &b, int64_t t0, int64_t t1, int n_tip_slots, FitchProgram &prog) {
  prog.clear();
  const int64_t n0 = b.node_off[t0], n1 = b.node_off[t1];
  const int64_t n_ops = n1 - n0;
  std::vector<int> level(n_ops);
  int max_level = 0;
  for (int64_t t = t0; t < t1; t++) {
    const int64_t base = b.node_off[t];
    for (int64_t x = base; x < b.node_off[t + 1]; x++) {
      int lv = 0;
      for (int64_t c = b.child_off[x]; c < b.child_off[x + 1]; c++) {
        int32_t ch = b.child[c];
        if (ch < 0) lv = std::max(lv, level[base + (-ch - 1) - n0] + 1);
      }
      level[x - n0] = lv;
      max_level = std::max(max_level, lv);
    }
  }
  prog.level_off.assign(max_level + 2, 0);
  for (int64_t i = 0; i < n_ops; i++) prog.level_off[level[i] + 1]++;
  for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cursor(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(n_ops);
  prog.children.resize(b.child_off[n1] - b.child_off[n0]);
  const int64_t c_base = b.child_off[n0];
  for (int64_t t = t0; t < t1; t++) {
    const int64_t base = b.node_off[t], end = b.node_off[t + 1];
    for (int64_t x = base; x < end; x++) {
      FitchOp op;
      op.tree = (int)(t - t0);
      op.nchild = (int)(b.child_off[x + 1] - b.child_off[x]);
      op.child_off = (int)(b.child_off[x] - c_base);
      op.dst = (x == end - 1) ? -1 : (int)(n_tip_slots + (x - n0));  // roots are not stored
      for (int64_t c = b.child_off[x]; c < b.child_off[x + 1]; c++) {
        int32_t ch = b.child[c];
        // (rows beyond the alignment only occur with zero patterns; clamp keeps the address valid)
        prog.children[c - c_base] =
            ch >= 0 ? std::min(ch, n_tip_slots - 1) : (int)(n_tip_slots + (base - n0) + (-ch - 1));
      }
      prog.ops[cursor[level[x - n0]]++] = op;
    }
  }
  prog.n_scratch_slots = n_ops;
  prog.n_costs = t1 - t0;
}

template <int K, int VW>
static void launch_level(const FitchAln &a, FitchWorkspace &ws, int64_t op0, int64_t n_ops, double level_bytes,
                         double level_req, cudaStream_t st) {
  const int64_t cols = a.words32 / VW;
  // block size: multiple of 32 in [64,256] wasting the fewest lanes
  int best_t = 256;
  int64_t best_waste = INT64_MAX;
  for (int t = 256; t >= 64; t -= 32) {
    int64_t blocks = (cols + t - 1) / t;
    int64_t waste = blocks * t - cols;
    if (waste < best_waste) { best_waste = waste; best_t = t; }
  }
  const int64_t bpo = (cols + best_t - 1) / best_t;
  const int64_t max_ops_per_launch = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
  for (int64_t o = 0; o < n_ops; o += max_ops_per_launch) {
    const int64_t cnt = std::min(max_ops_per_launch, n_ops - o);
    prof_begin(PROF_FITCH_OPS, st);
    fitch_level_kernel<K, VW><<<(unsigned)(cnt * bpo), best_t, 0, st>>>(
        a.tips.p, ws.scratch.p, a.n_cols, ws.ops.p + op0 + o, ws.children.p, a.words32, (int)bpo,
        a.wbits.p, a.chunk_nb.p, ws.cost.p);
    count_launch();
    prof_end(PROF_FITCH_OPS, st, level_bytes * (double)cnt / (double)n_ops, (double)cnt * (double)a.P,
             level_req * (double)cnt / (double)n_ops);
  }
}

// one helper stream per device for work that is independent of the level launches
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ready = nullptr, planned = nullptr;
};
static SideStream &side_stream() {
  static SideStream ss[64];
  int dev = 0;
  cudaGetDevice(&dev);
  SideStream &x = ss[dev & 63];
  if (!x.s) {
    PLG_CUDA(cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking));
    PLG_CUDA(cudaEventCreateWithFlags(&x.ready, cudaEventDisableTiming));
    PLG_CUDA(cudaEventCreateWithFlags(&x.planned, cudaEventDisableTiming));
  }
  return x;
}

void run_fitch_program(const FitchAln &a, FitchWorkspace &ws, const FitchProgram &prog, int32_t *out_costs,
                       cudaStream_t st, int32_t *out_moves) {
  const size_t vec_words = (size_t)a.K * a.words32;
  ws.scratch.alloc(std::max<size_t>(1, (size_t)prog.n_scratch_slots * vec_words));
  ws.ops.alloc(std::max<size_t>(1, prog.ops.size()));
  ws.children.alloc(std::max<size_t>(1, prog.children.size()));
  ws.cost.alloc(std::max<size_t>(1, (size_t)prog.n_costs));
  if (!prog.ops.empty()) {
    PLG_CUDA(cudaMemcpyAsync(ws.ops.p, prog.ops.data(), prog.ops.size() * sizeof(FitchOp), cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemcpyAsync(ws.children.p, prog.children.data(), prog.children.size() * sizeof(int),
                             cudaMemcpyHostToDevice, st));
  }
  PLG_CUDA(cudaMemsetAsync(ws.cost.p, 0, std::max<size_t>(1, (size_t)prog.n_costs) * sizeof(unsigned int), st));
  const bool device_plan = !prog.spr_searches.empty() && prog.spr_n_visits > 0;
  SideStream &side = side_stream();
  if (device_plan) {
    // Device-planned SPR neighbourhood: the visit list is written by spr_plan_kernel from O(n)
    // records, not uploaded -- on a side stream, so that its one-warp-per-search latency chain
    // (~0.12 ms at 256 taxa) overlaps the base tree's level launches instead of preceding the walk.
    ws.nbw_visits.alloc((size_t)prog.spr_n_visits * 2);
    ws.spr_edges.alloc(prog.spr_edges.size());
    ws.spr_searches.alloc(prog.spr_searches.size() * 2);
    if (out_moves) { ws.spr_moves.alloc((size_t)prog.spr_n_unique); ws.spr_moves_host.alloc((size_t)prog.spr_n_unique); }
    PLG_CUDA(cudaEventRecord(side.ready, st));  // (the buffers above are stream-ordered allocations of `st`)
    PLG_CUDA(cudaStreamWaitEvent(side.s, side.ready, 0));
    PLG_CUDA(cudaMemcpyAsync(ws.spr_edges.p, prog.spr_edges.data(), prog.spr_edges.size() * sizeof(SprEdgeRec),
                             cudaMemcpyHostToDevice, side.s));
    PLG_CUDA(cudaMemcpyAsync(ws.spr_searches.p, prog.spr_searches.data(),
                             prog.spr_searches.size() * sizeof(SprSearchRec), cudaMemcpyHostToDevice, side.s));
    const int n_search = (int)prog.spr_searches.size();
    const size_t edge_bytes = prog.spr_edges.size() * sizeof(SprEdgeRec);
    const bool in_smem = edge_bytes <= SPR_PLAN_SMEM_MAX && !getenv("PLG_SPR_PLAN_GLOBAL");  // (tests: the fallback for huge trees)
    static bool plan_attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!plan_attr_set[dev & 63]) {
      PLG_CUDA(cudaFuncSetAttribute(spr_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPR_PLAN_SMEM_MAX));
      plan_attr_set[dev & 63] = true;
    }
    spr_plan_kernel<<<(n_search + 63) / 64, 64, in_smem ? edge_bytes : 0, side.s>>>(
        ws.spr_edges.p, in_smem ? (int)prog.spr_edges.size() : 0, ws.spr_searches.p, n_search, prog.spr_first_acc,
        (int)prog.spr_lo, (int)prog.spr_hi, ws.nbw_visits.p, out_moves ? ws.spr_moves.p : nullptr);
    count_launch();
    PLG_CUDA(cudaEventRecord(side.planned, side.s));
    if (out_moves)  // the move table comes back while the walk runs; copied out after the final sync
      PLG_CUDA(cudaMemcpyAsync(ws.spr_moves_host.p, ws.spr_moves.p, (size_t)prog.spr_n_unique * sizeof(int4),
                               cudaMemcpyDeviceToHost, side.s));
  }
  for (size_t l = 0; l + 1 < prog.level_off.size(); l++) {
    const int64_t o0 = prog.level_off[l], cnt = prog.level_off[l + 1] - o0;
    if (!cnt) continue;
    // algorithmic bytes of this launch: every operand vector read once, every stored vector
    // written once, K_real planes of P bits each (3*K_p/8 B per binary node-op)
    double vecs = 0, req_vecs = 0;
    for (int64_t o = o0; o < o0 + cnt; o++) {
      const int nc = prog.ops[o].nchild;
      const int st_ = prog.ops[o].dst >= 0 ? 1 : 0;
      if (nc == FITCH_FUSED) { vecs += 6; req_vecs += 4 + st_; }  // fold (2 in, 1 out) + 3-way join (3 in)
      else if (nc == FITCH_JOIN3) { vecs += 3; req_vecs += 3; }
      else { vecs += nc + 1; req_vecs += nc + st_; }
    }
    const double lb = vecs * a.K_real * (double)a.P / 8.0;
    const double lreq = req_vecs * a.K_real * (double)a.P / 8.0;
    switch (a.K) {
      case 2: launch_level<2, 4>(a, ws, o0, cnt, lb, lreq, st); break;
      case 3: launch_level<3, 4>(a, ws, o0, cnt, lb, lreq, st); break;
      case 4: launch_level<4, 4>(a, ws, o0, cnt, lb, lreq, st); break;
      case 5: launch_level<5, 4>(a, ws, o0, cnt, lb, lreq, st); break;
      case 6: launch_level<6, 4>(a, ws, o0, cnt, lb, lreq, st); break;
      case 8: launch_level<8, 2>(a, ws, o0, cnt, lb, lreq, st); break;
      case 10: launch_level<10, 2>(a, ws, o0, cnt, lb, lreq, st); break;
      case 12: launch_level<12, 1>(a, ws, o0, cnt, lb, lreq, st); break;
      case 16: launch_level<16, 1>(a, ws, o0, cnt, lb, lreq, st); break;
      case 20: launch_level<20, 1>(a, ws, o0, cnt, lb, lreq, st); break;
      default: PLG_FAIL(PLG_ERR_STATES, "no kernel for %d planes", a.K);
    }
  }
  if (!prog.walk_groups.empty()) {
    if (!device_plan) ws.nbw_visits.alloc(prog.walk_visits.size() * 2);
    ws.nbw_groups.alloc(prog.walk_groups.size());
    if (device_plan) {
      // the visit list was written by spr_plan_kernel on the side stream while the levels ran
      PLG_CUDA(cudaStreamWaitEvent(st, side.planned, 0));
    } else {
      PLG_CUDA(cudaMemcpyAsync(ws.nbw_visits.p, prog.walk_visits.data(),
                               prog.walk_visits.size() * sizeof(FitchWalkVisit), cudaMemcpyHostToDevice, st));
    }
    PLG_CUDA(cudaMemcpyAsync(ws.nbw_groups.p, prog.walk_groups.data(), prog.walk_groups.size() * sizeof(FitchWalkGroup),
                             cudaMemcpyHostToDevice, st));
    // algorithmic bytes as for the per-step ops it replaces: a scored neighbour = fold (2 in, 1 out)
    // + 3-way join (3 in); requested: D1, D2 per visit, a stored X at the first step of a side,
    // Tv per search (partials stay on chip)
    double vecs = 0, req_vecs = 0, units = 0;
    if (device_plan) {
      vecs = prog.spr_vecs;
      req_vecs = prog.spr_req_vecs;
      units = 2.0 * (double)prog.spr_n_visits;
    } else {
      for (size_t g = 0; g < prog.walk_groups.size(); g++) req_vecs += 1;
    }
    for (size_t v = 0; !device_plan && v < prog.walk_visits.size(); v++) {
      const FitchWalkVisit &w = prog.walk_visits[v];
      const bool n1 = w.tree1 >= 0 || w.keep == 1 || (w.push >= 0 && w.keep == 2);
      const bool n2 = w.tree2 >= 0 || w.keep == 2 || (w.push >= 0 && w.keep == 1);
      vecs += (n1 ? 3 : 0) + (n2 ? 3 : 0) + (w.tree1 >= 0 ? 3 : 0) + (w.tree2 >= 0 ? 3 : 0);
      req_vecs += 2 + (w.x_in >= 0 ? 1 : 0);
      units += 2;
    }
    const double vb = a.K_real * (double)a.P / 8.0;
    const FitchWalkVisit *dv = reinterpret_cast<const FitchWalkVisit *>(ws.nbw_visits.p);
    const FitchWalkGroup *dg = reinterpret_cast<const FitchWalkGroup *>(ws.nbw_groups.p);
    const int ng = (int)prog.walk_groups.size();
    prof_begin(PROF_FITCH_OPS, st);
    switch (a.K) {
      case 2: launch_fitch_search_walk<2, 4>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 3: launch_fitch_search_walk<3, 4>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 4: launch_fitch_search_walk<4, PLG_FSW_VW_DNA>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 5: launch_fitch_search_walk<5, 4>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 6: launch_fitch_search_walk<6, 4>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 8: launch_fitch_search_walk<8, 2>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 10: launch_fitch_search_walk<10, 2>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 12: launch_fitch_search_walk<12, 1>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 16: launch_fitch_search_walk<16, 1>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      case 20: launch_fitch_search_walk<20, 1>(a, ws.scratch.p, dv, dg, ng, ws.cost.p, st); break;
      default: PLG_FAIL(PLG_ERR_STATES, "no kernel for %d planes", a.K);
    }
    count_launch();
    prof_end(PROF_FITCH_OPS, st, vecs * vb, units * (double)a.P, req_vecs * vb);
  }
  PLG_CUDA(cudaGetLastError());
  if (out_costs && prog.n_costs) {
    PLG_CUDA(cudaMemcpyAsync(out_costs, ws.cost.p, (size_t)prog.n_costs * sizeof(int32_t), out_copy_kind(), st));
    PLG_CUDA(cudaStreamSynchronize(st));
  }
  if (device_plan && out_moves) {
    PLG_CUDA(cudaStreamSynchronize(side.s));
    // pinned staging -> the caller's (pageable) table, on the host cores when it is large
    const size_t bytes = (size_t)prog.spr_n_unique * sizeof(int4);
    const int n_thr = (int)std::max<size_t>(1, std::min<size_t>(HostPool::get().size(), bytes >> 20));
    const char *src = reinterpret_cast<const char *>(ws.spr_moves_host.p);
    char *dst = reinterpret_cast<char *>(out_moves);
    HostPool::get().run(n_thr, [&](int ti, int nt) {
      const size_t b0 = bytes * ti / nt, b1 = bytes * (ti + 1) / nt;
      memcpy(dst + b0, src + b0, b1 - b0);
    });
  }
}

// Scores every tree of a batch by full re-traversal, splitting it so the scratch
// pool stays inside the memory budget.
void score_batch_full(FitchAln &a, const TreeBatch &b, int32_t *out_costs, cudaStream_t st) {
  const int64_t n_trees = b.n_trees();
  if (n_trees == 0) return;
  StageTimer tm;
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  tm.lap("meminfo");
  const size_t vec_bytes = (size_t)a.K * a.words32 * 4;
  // scratch budget: half of what is free (counting the pool we already hold), capped at 48 GiB
  const size_t budget = std::min<size_t>((free_b + fitch_workspace().scratch.n * 4) / 2, (size_t)48 << 30);
  const int64_t max_slots = std::max<int64_t>(1, (int64_t)(budget / vec_bytes));
  static thread_local FitchProgram prog;  // keeps its pinned staging across calls
  int64_t t0 = 0;
  while (t0 < n_trees) {
    int64_t t1 = t0 + 1;  // a single over-budget tree is still attempted; cudaMalloc reports failure
    while (t1 < n_trees && b.node_off[t1 + 1] - b.node_off[t0] <= max_slots) t1++;
    build_full_program(b, t0, t1, a.n_cols, prog);
    tm.lap("build_full");
    run_fitch_program(a, fitch_workspace(), prog, out_costs + t0, st);
    tm.lap("run");
    t0 = t1;
  }
}

// Descriptor batches: tree walks planned on the device (treewalk.h).  The host copies the raw
// descriptors up and the costs down; nothing per node happens on the host.
bool tree_walk_enabled() {
  const char *e = getenv("PLG_TREE_MODE");
  return !(e && !strcmp(e, "levels"));
}

void score_desc_walk(FitchAln &a, int64_t n_trees, int n_tips, const int32_t *desc, const int32_t *tip_rows,
                     int32_t *out_costs, cudaStream_t st) {
  if (n_trees == 0) return;
  StageTimer tm;
  validate_desc(n_trees, n_tips, desc, tip_rows, a.n_cols, a.P > 0);
  tm.lap("validate");
  FitchWorkspace &ws = fitch_workspace();
  const int ni = n_tips - 1;
  int64_t chunk = std::min<int64_t>(n_trees, ((int64_t)512 << 20) / (36 * (int64_t)ni + 8));
  if (const char *e = getenv("PLG_WALK_CHUNK")) chunk = std::min<int64_t>(chunk, atoll(e));  // tests: force several chunks
  chunk = std::max<int64_t>(1, chunk);
  ws.walk_desc.alloc((size_t)chunk * ni * 2);
  ws.walk_ops.alloc((size_t)chunk * ni);
  ws.walk_scratch.alloc((size_t)chunk * (3 * (size_t)ni + 2));
  ws.cost.alloc((size_t)chunk);
  ws.walk_rows.alloc((size_t)n_tips);
  if (tip_rows)
    PLG_CUDA(cudaMemcpyAsync(ws.walk_rows.p, tip_rows, (size_t)n_tips * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  for (int64_t t0 = 0; t0 < n_trees; t0 += chunk) {
    const int64_t cnt = std::min(chunk, n_trees - t0);
    PLG_CUDA(cudaMemcpyAsync(ws.walk_desc.p, desc + (size_t)t0 * ni * 2, (size_t)cnt * ni * 2 * sizeof(int32_t),
                             cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemsetAsync(ws.cost.p, 0, (size_t)cnt * sizeof(unsigned int), st));
    launch_plan_walk(cnt, n_tips, ws.walk_desc.p, tip_rows ? ws.walk_rows.p : nullptr, a.n_cols,
                     reinterpret_cast<WalkOp *>(ws.walk_ops.p), ws.walk_scratch.p, st);
    const WalkOp *ops = reinterpret_cast<const WalkOp *>(ws.walk_ops.p);
    switch (a.K) {
      case 2: launch_fitch_walk<2, 4>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 3: launch_fitch_walk<3, 4>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 4: launch_fitch_walk<4, 4>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 5: launch_fitch_walk<5, 4>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 6: launch_fitch_walk<6, 4>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 8: launch_fitch_walk<8, 2>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 10: launch_fitch_walk<10, 2>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 12: launch_fitch_walk<12, 1>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 16: launch_fitch_walk<16, 1>(a, ops, cnt, ni, ws.cost.p, st); break;
      case 20: launch_fitch_walk<20, 1>(a, ops, cnt, ni, ws.cost.p, st); break;
      default: PLG_FAIL(PLG_ERR_STATES, "no kernel for %d planes", a.K);
    }
    PLG_CUDA(cudaGetLastError());
    PLG_CUDA(cudaMemcpyAsync(out_costs + t0, ws.cost.p, (size_t)cnt * sizeof(int32_t), out_copy_kind(), st));
    PLG_CUDA(cudaStreamSynchronize(st));
    tm.lap("walk chunk");
  }
}

}  // namespace plg

// ----------------------------------------------------------------------------
// C-ABI
// ----------------------------------------------------------------------------
using namespace plg;

static std::vector<uint8_t> gather_profiles(int64_t n_prof, const char *const *profiles, int *n_cols) {
  size_t mn = SIZE_MAX;
  for (int64_t i = 0; i < n_prof; i++) {
    if (!profiles[i]) PLG_FAIL(PLG_ERR_ARG, "profile %lld is NULL", (long long)i);
    mn = std::min(mn, strlen(profiles[i]));
  }
  if (n_prof == 0) mn = 1;  // no patterns: every tree costs 0 (fitch.cpp:249 loop is empty)
  if (mn == 0) PLG_FAIL(PLG_ERR_ARG, "empty profile string");
  *n_cols = (int)mn;
  std::vector<uint8_t> dense((size_t)n_prof * mn);
  for (int64_t i = 0; i < n_prof; i++) memcpy(dense.data() + (size_t)i * mn, profiles[i], mn);
  return dense;
}

extern "C" int plg_fitch_aln_create_dense(int64_t n_prof, int n_cols, const uint8_t *states,
                                          const int64_t *weights, plg_fitch_aln **out) {
  PLG_API_BEGIN_LOCKED
  if (!out || (n_prof > 0 && (!states || !weights))) PLG_FAIL(PLG_ERR_ARG, "null argument");
  *out = new plg_fitch_aln(n_prof, n_cols, states, weights);
  PLG_API_END
}

extern "C" int plg_fitch_aln_create(int64_t n_prof, const char *const *profiles, const int64_t *weights,
                                    plg_fitch_aln **out) {
  PLG_API_BEGIN_LOCKED
  if (!out || (n_prof > 0 && (!profiles || !weights))) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int n_cols = 0;
  std::vector<uint8_t> dense = gather_profiles(n_prof, profiles, &n_cols);
  *out = new plg_fitch_aln(n_prof, n_cols, dense.data(), weights);
  PLG_API_END
}

extern "C" void plg_fitch_aln_destroy(plg_fitch_aln *aln) { delete aln; }

extern "C" int plg_fitch_aln_info(const plg_fitch_aln *aln, int64_t *n_prof, int *n_cols, int *n_planes) {
  PLG_API_BEGIN
  if (!aln) PLG_FAIL(PLG_ERR_ARG, "null alignment handle");
  if (n_prof) *n_prof = aln->impl.P;
  if (n_cols) *n_cols = aln->impl.n_cols;
  if (n_planes) *n_planes = aln->impl.K;
  PLG_API_END
}

static void check_rows(const FitchAln &a, const TreeBatch &b) {
  if (a.P == 0) return;  // no patterns: rows are never read (fitch.cpp:249 loop body never runs)
  for (int32_t c : b.child)
    if (c >= a.n_cols)
      PLG_FAIL(PLG_ERR_ARG, "taxon index %d is outside the profile length %d (fitch.cpp:256 substr would throw)",
               c, a.n_cols);
}

extern "C" int plg_fitch_score_newicks(plg_fitch_aln *aln, int64_t n_trees, const char *const *newicks,
                                       int n_taxa, const char *const *taxa_names, const int32_t *taxa_idx,
                                       int32_t *out_costs, void *stream) {
  PLG_API_BEGIN_LOCKED
  if (!aln || !out_costs || (n_trees > 0 && !newicks) || (n_taxa > 0 && (!taxa_names || !taxa_idx)))
    PLG_FAIL(PLG_ERR_ARG, "null argument");
  TaxaMap taxa(n_taxa, taxa_names, taxa_idx);
  TreeBatch b;
  for (int64_t t = 0; t < n_trees; t++) append_tree(b, parse_newick(newicks[t]), taxa, false);
  check_rows(aln->impl, b);
  score_batch_full(aln->impl, b, out_costs, (cudaStream_t)stream);
  PLG_API_END
}

static void fitch_score_trees_impl(plg_fitch_aln *aln, int64_t n_trees, int n_tips, const int32_t *desc,
                                   const int32_t *tip_rows, int32_t *out_costs, void *stream) {
  if (!aln || !out_costs || (n_trees > 0 && !desc)) PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (tree_walk_enabled() && n_tips >= 2 && n_tips <= (1 << WALK_MAX_LEVELS)) {
    score_desc_walk(aln->impl, n_trees, n_tips, desc, tip_rows, out_costs, (cudaStream_t)stream);
    return;
  }
  TreeBatch b = batch_from_desc(n_trees, n_tips, desc, nullptr, tip_rows);
  check_rows(aln->impl, b);
  score_batch_full(aln->impl, b, out_costs, (cudaStream_t)stream);
}

extern "C" int plg_fitch_score_trees(plg_fitch_aln *aln, int64_t n_trees, int n_tips, const int32_t *desc,
                                     const int32_t *tip_rows, int32_t *out_costs, void *stream) {
  PLG_API_BEGIN_LOCKED
  fitch_score_trees_impl(aln, n_trees, n_tips, desc, tip_rows, out_costs, stream);
  PLG_API_END
}

extern "C" int plg_fitch_score_trees_shard(plg_fitch_aln *aln, int64_t n_trees, int n_tips, const int32_t *desc,
                                           const int32_t *tip_rows, int64_t shard_index, int64_t shard_count,
                                           int32_t *out_costs, int out_where, void *stream) {
  PLG_API_BEGIN_LOCKED
  if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
    PLG_FAIL(PLG_ERR_ARG, "bad shard %lld of %lld", (long long)shard_index, (long long)shard_count);
  if (out_where != PLG_OUT_HOST && out_where != PLG_OUT_DEVICE) PLG_FAIL(PLG_ERR_ARG, "out_where: PLG_OUT_HOST or PLG_OUT_DEVICE");
  if (n_tips < 2) PLG_FAIL(PLG_ERR_ARG, "descriptor trees need at least 2 tips");
  const int64_t lo = n_trees * shard_index / shard_count, hi = n_trees * (shard_index + 1) / shard_count;
  const size_t stride = (size_t)(n_tips - 1) * 2;
  OutDeviceScope scope(out_where == PLG_OUT_DEVICE);
  if (hi > lo)
    fitch_score_trees_impl(aln, hi - lo, n_tips, desc ? desc + lo * stride : nullptr, tip_rows,
                           out_costs ? out_costs + lo : nullptr, stream);
  PLG_API_END
}

extern "C" int plg_fitch_cost(const char *newick, int64_t n_prof, const char *const *profiles,
                              const int64_t *weights, int n_taxa, const char *const *taxa_names,
                              const int32_t *taxa_idx, int32_t *out_cost) {
  PLG_API_BEGIN_LOCKED
  if (!newick || !out_cost) PLG_FAIL(PLG_ERR_ARG, "null argument");
  // parse + resolve labels first: a bad tree must fail before any device work
  TaxaMap taxa(n_taxa, taxa_names, taxa_idx);
  TreeBatch b;
  append_tree(b, parse_newick(newick), taxa, false);
  int n_cols = 0;
  std::vector<uint8_t> dense = gather_profiles(n_prof, profiles, &n_cols);
  FitchAln a(n_prof, n_cols, dense.data(), weights);
  check_rows(a, b);
  score_batch_full(a, b, out_cost, 0);
  PLG_API_END
}
