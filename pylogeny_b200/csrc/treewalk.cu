// Device-side planner of tree-walk programs (see treewalk.h).
#include <vector>

#include "treewalk.h"

namespace plg {

// One thread per tree.  Pass 1 (descriptor order = children first): stack levels each subtree
// needs when its result must end in registers.  Pass 2: iterative post-order from the root with
// the hungrier child first; the first child's result is parked at the current depth while the
// second one is evaluated one level up.
//
// frame = node << 12 | depth << 7 | (dest + 1) << 1 | stage
__global__ void plan_walk_kernel(int64_t n_trees, int n_tips, const int32_t *__restrict__ desc,
                                 const int32_t *__restrict__ tip_rows, int n_rows, WalkOp *__restrict__ ops,
                                 int *__restrict__ scratch) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_trees) return;
  const int ni = n_tips - 1;
  const int32_t *d = desc + t * ni * 2;
  WalkOp *out = ops + t * ni;
  int *need = scratch + t * (3 * (int64_t)ni + 2);  // [ni]
  int *stk = need + ni;                              // [2*ni + 2]
  for (int i = 0; i < ni; i++) {
    const int a = d[2 * i], b = d[2 * i + 1];
    const int na = a < n_tips ? -1 : need[a - n_tips], nb = b < n_tips ? -1 : need[b - n_tips];
    int nd;
    if (na < 0 && nb < 0) nd = 0;
    else if (na < 0 || nb < 0) nd = max(na, nb);
    else nd = max(max(na, nb), 1 + min(na, nb));
    need[i] = nd;
  }
  auto row = [&](int id) {
    const int r = tip_rows ? tip_rows[id] : id;
    return min(max(r, 0), n_rows - 1);
  };
  int sp = 0, n_out = 0;
  stk[sp++] = (ni - 1) << 12;  // root: depth 0, dest -1 (registers), stage 0
  while (sp > 0) {
    const int f = stk[--sp];
    const int i = f >> 12, depth = (f >> 7) & 31, dest = ((f >> 1) & 63) - 1, stage = f & 1;
    const int a = d[2 * i], b = d[2 * i + 1];
    const bool ia = a >= n_tips, ib = b >= n_tips;
    // with two internal children: which one goes first (and waits on the stack)
    const bool a_first = ia && ib && need[a - n_tips] >= need[b - n_tips];
    if (stage == 0 && (ia || ib)) {
      stk[sp++] = f | 1;
      if (ia && ib) {
        const int first = a_first ? a : b, second = a_first ? b : a;
        stk[sp++] = (second - n_tips) << 12 | (depth + 1) << 7;               // result stays in registers
        stk[sp++] = (first - n_tips) << 12 | depth << 7 | (depth + 1) << 1;  // result parked at `depth`
      } else {
        stk[sp++] = ((ia ? a : b) - n_tips) << 12 | depth << 7;
      }
      continue;
    }
    WalkOp o;
    o.a = !ia ? row(a) : (ib && a_first ? -2 - depth : -1);
    o.b = !ib ? row(b) : (ia && !a_first ? -2 - depth : -1);
    o.push = dest;
    o.slot = 2 * i;
    out[n_out++] = o;
  }
}

void launch_plan_walk(int64_t n_trees, int n_tips, const int32_t *d_desc, const int32_t *d_tip_rows, int n_rows,
                      WalkOp *d_ops, int *d_scratch, cudaStream_t st) {
  if (n_trees <= 0) return;
  plan_walk_kernel<<<(unsigned)((n_trees + 63) / 64), 64, 0, st>>>(n_trees, n_tips, d_desc, d_tip_rows, n_rows, d_ops,
                                                                  d_scratch);
  count_launch();
}

void validate_desc(int64_t n_trees, int n_tips, const int32_t *desc, const int32_t *tip_rows, int n_rows,
                   bool rows_matter) {
  if (n_tips < 2) PLG_FAIL(PLG_ERR_ARG, "descriptor trees need at least 2 tips");
  const int ni = n_tips - 1;
  // every tip and every non-root internal node is a child exactly once: 2 * ni slots for 2 * ni ids,
  // so "no id twice" is the whole condition.  A DAG or a forest would make the device planner
  // (plan_walk_kernel) write more ops than the tree's slot holds.
  std::vector<int64_t> used_by((size_t)n_tips + ni, -1);
  for (int64_t t = 0; t < n_trees; t++) {
    const int32_t *d = desc + (size_t)t * ni * 2;
    for (int i = 0; i < ni; i++)
      for (int c = 0; c < 2; c++) {
        const int32_t id = d[2 * i + c];
        if (id < 0 || id >= n_tips + i)
          PLG_FAIL(PLG_ERR_ARG, "descriptor of tree %lld node %d: child id %d is not earlier in post-order",
                   (long long)t, i, id);
        if (used_by[id] == t)
          PLG_FAIL(PLG_ERR_ARG, "descriptor of tree %lld node %d: node %d is a child twice (not a tree)",
                   (long long)t, i, id);
        used_by[id] = t;
      }
  }
  if (rows_matter) {
    if (tip_rows) {
      for (int i = 0; i < n_tips; i++)
        if (tip_rows[i] < 0 || tip_rows[i] >= n_rows)
          PLG_FAIL(PLG_ERR_ARG, "tip row %d is outside the alignment (%d rows)", tip_rows[i], n_rows);
    } else if (n_tips > n_rows) {
      // only an error if such a tip is referenced; ids < n_tips are all legal children
      for (int64_t t = 0; t < n_trees; t++) {
        const int32_t *d = desc + (size_t)t * ni * 2;
        for (int i = 0; i < 2 * ni; i++)
          if (d[i] < n_tips && d[i] >= n_rows)
            PLG_FAIL(PLG_ERR_ARG, "tip row %d is outside the alignment (%d rows)", d[i], n_rows);
      }
    }
  }
}

}  // namespace plg
