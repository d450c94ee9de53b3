/* pylogeny_b200 -- C-ABI of the B200-native tree-scoring engine.
 *
 * This is the drop-in boundary for Pylogeny's two native extension modules:
 *   `fitch`          (/root/reference/pylogeny/fitch.cpp, CPython-2 binding :274-346)
 *   `libpllWrapper`  (/root/reference/pylogeny/libpllWrapper.c, method table :323-339)
 * Plain pointers and sizes only; no exceptions cross the ABI; every function
 * returns a status (PLG_OK == 0) and plg_last_error() holds the message of the
 * last failure on the calling thread.  Handles are opaque and explicitly
 * destroyed.  Device memory is owned by the handles; `stream` arguments are
 * cudaStream_t passed as void* (0 = default stream).  There is NO CPU fallback:
 * without a CUDA device every compute entry point returns PLG_ERR_CUDA.
 *
 * The ctypes shim modules pylogeny_b200/fitch.py and pylogeny_b200/libpllWrapper.py
 * bind exactly these symbols; INTEGRATION.md shows the reference-side stub.
 */
#ifndef PYLOGENY_B200_H
#define PYLOGENY_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLG_OK 0
#define PLG_ERR_ARG 1         /* bad argument, length mismatch (fitch.cpp:285 returns NULL) */
#define PLG_ERR_PARSE 2       /* Newick string could not be parsed */
#define PLG_ERR_LABEL 3       /* leaf label missing from the taxa map (fitch.cpp:255 aborts) */
#define PLG_ERR_CUDA 4        /* CUDA error or no device */
#define PLG_ERR_STATES 5      /* more distinct states in one column than the kernels support */
#define PLG_ERR_UNSUPPORTED 6 /* e.g. non-binary tree handed to a binary-only entry point */
#define PLG_ERR_IO 7          /* file could not be read (libpllWrapper.c:141-146 IOError) */

#define PLG_MOVE_NNI 2 /* rearrangement.py:27 TYPE_SPR, TYPE_NNI, TYPE_TBR = 1, 2, 3 */
#define PLG_MOVE_SPR 1
#define PLG_MOVE_TBR 3
/* SPR moves whose regraft point is at most r branches from the pruning point (the maxdist of
 * getSPRMovesInDistance, libpllWrapper.c:241-262); r = 0: unlimited; r = 1 equals NNI. */
#define PLG_MOVE_SPR_WITHIN(r) (PLG_MOVE_SPR | ((r) << 8))

/* Where a scoring entry point with an `out_where` argument writes its per-tree results. */
#define PLG_OUT_HOST 0
#define PLG_OUT_DEVICE 1 /* device memory of the current device, e.g. a tensor that goes straight into an NCCL collective */

#define PLG_DATA_DNA 4  /* 4 states, tip codes are 4-bit ambiguity masks (A=1,C=2,G=4,T=8) */
#define PLG_DATA_AA 20  /* 20 states, tip codes 0..19, 20=B, 21=Z, 22=X/-/? */

const char *plg_last_error(void);
int plg_version(void);
int plg_device_count(int *count);
int plg_set_device(int device);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t plg_launch_count(void);
/* Per-kernel timing for the roofline report: CUDA events recorded on the launching stream
 * around every launch while enabled.  kind: 0 Fitch node-ops, 1 CLV update, 2 root/3-way
 * evaluation, 3 P-matrices, 4 depth-first likelihood walks (SPR search walks, tree walks).
 * plg_profile_read sums and clears one kind: device ms,
 * algorithmic bytes (SURVEY.md 8d per-node-op figures, unaffected by fusion or operand
 * sharing), node-op x pattern units, launches, and the bytes the kernels actually requested. */
int plg_profile_enable(int on);
int plg_profile_read(int kind, double *ms, double *bytes, double *units, int64_t *launches,
                     double *requested_bytes);

/* Pipe peaks of the current device, measured now (CUDA events, best of three): mode 0 = DFMA on the
 * CUDA cores, 1 = DMMA m8n8k4 (FP64 tensor instructions), 2 = both interleaved -- TFLOP/s; 3 = LOP3
 * (INT32 logic) -- 10^12 lane-ops/s.  The roofline denominators of the likelihood and Fitch walk
 * kernels (bench.py); not in MEASURED_PEAKS.json. */
int plg_fp64_peak(int mode, double *tflops);

/* Diagnostics of the host worker pool the planners run on (no device needed): n slices, slice fail_at
 * (if in [0, n)) raises the planners' error.  Returns PLG_ERR_ARG in that case, PLG_OK otherwise;
 * *ran = slices that completed (n - 1 on failure: a failing slice never takes the process or the
 * other slices down). */
int plg_hostpool_selftest(int n, int fail_at, int *ran);

/* ------------------------------------------------------------------ */
/* Parsimony (replaces fitch.cpp)                                      */
/* ------------------------------------------------------------------ */
typedef struct plg_fitch_aln plg_fitch_aln;

/* Device-resident, bit-plane packed pattern set.  `profiles[i]` is the i-th
 * profile string exactly as scoring.py:134 builds it (one char per alignment
 * row; fitch.cpp:256 reads profile[taxInd]); weights as parsimony.py:28-46.
 * Replaces the per-call copy loop fitch.cpp:293-309. */
int plg_fitch_aln_create(int64_t n_prof, const char *const *profiles,
                         const int64_t *weights, plg_fitch_aln **out);
/* Same from a dense row-major [n_prof][n_cols] byte matrix of state symbols. */
int plg_fitch_aln_create_dense(int64_t n_prof, int n_cols, const uint8_t *states,
                               const int64_t *weights, plg_fitch_aln **out);
void plg_fitch_aln_destroy(plg_fitch_aln *aln);
int plg_fitch_aln_info(const plg_fitch_aln *aln, int64_t *n_prof, int *n_cols, int *n_planes);

/* fitch.calculateCost(newick, profiles, weights, taxa) -> int  (fitch.cpp:274-334).
 * One-shot: packs + uploads the profiles, scores one tree, frees. k-ary nodes
 * follow fitch.cpp:215-226/262-264 exactly; result wraps as 32-bit int (:247,268). */
int plg_fitch_cost(const char *newick, int64_t n_prof, const char *const *profiles,
                   const int64_t *weights, int n_taxa, const char *const *taxa_names,
                   const int32_t *taxa_idx, int32_t *out_cost);
/* Same, against a resident alignment, for many Newick strings at once. */
int plg_fitch_score_newicks(plg_fitch_aln *aln, int64_t n_trees,
                            const char *const *newicks, int n_taxa,
                            const char *const *taxa_names, const int32_t *taxa_idx,
                            int32_t *out_costs, void *stream);
/* Batched post-order descriptor (the hot path): n_trees rooted binary trees over
 * the same n_tips leaves.  desc[t][i][0..1] = children of internal node i of tree
 * t, nodes in post-order; child id < n_tips is leaf (alignment row tip_rows[id]),
 * child id >= n_tips is internal node (id - n_tips) of the same tree; the root is
 * internal node n_tips-2.  Scores to host memory. */
int plg_fitch_score_trees(plg_fitch_aln *aln, int64_t n_trees, int n_tips,
                          const int32_t *desc, const int32_t *tip_rows,
                          int32_t *out_costs, void *stream);

/* Multi-GPU form of the same (BASELINE configs[3]: a list of random trees sharded over the GPUs of a
 * box): `desc` is the WHOLE batch, the call reads and scores only block
 * [n_trees*i/c, n_trees*(i+1)/c) of it and writes that block of out_costs (indexed by tree, host or
 * device memory per out_where).  One process per GPU calls it with its rank as shard_index; the
 * blocks are combined by one all-gather (pylogeny_b200/distributed.py). */
int plg_fitch_score_trees_shard(plg_fitch_aln *aln, int64_t n_trees, int n_tips,
                                const int32_t *desc, const int32_t *tip_rows,
                                int64_t shard_index, int64_t shard_count,
                                int32_t *out_costs, int out_where, void *stream);

/* ---- rearrangement neighbourhoods (host side; no device needed) ----
 * Base tree = rooted binary descriptor as above (n_tips-1 internal nodes, ids n_tips+i, root
 * last).  It is read as the unrooted tree it represents (the root's two child branches merge).
 * Every DISTINCT neighbour topology is produced once -- NNI: 2(n-3), SPR: 2(n-3)(2n-7), TBR: by
 * enumeration (textbook bisection/reconnection; the reference only names TBR,
 * rearrangement.py:27,88-90,661-664) -- in a
 * deterministic order, identified by a canonical topology hash (sum over splits), whereas the
 * reference emits 4(n-3)(n-2) SPR moves with duplicates (rearrangement.py:580-594). */
int plg_neighbourhood_size(int move_type, int n_tips, const int32_t *desc, int64_t *n_moves);
/* Materialise selected neighbours: out_desc/out_brlen [n_sel][n_tips-1][2], rooted on tip 0's
 * branch (rearrangement.py:267-331); neighbour branch lengths follow rearrangement.py:466-497
 * (target branch halved, the two branches left at the pruned node summed).  out_hash[n_sel] =
 * canonical topology hashes (tips identified by tip_rows, NULL: by id). */
int plg_neighbourhood_trees(int move_type, int n_tips, const int32_t *desc, const double *brlen,
                            const int32_t *tip_rows, int64_t n_sel, const int64_t *sel,
                            int32_t *out_desc, double *out_brlen, uint64_t *out_hash);
/* Every (pruned branch, target branch) pair of an NNI / SPR neighbourhood INCLUDING the pairs that repeat
 * a topology -- the move list the reference walks (rearrangement.py:543-594: allSPR emits 4(n-3)(n-2)
 * moves for 2(n-3)(2n-7) topologies).  out_moves (NULL: count only): int32 [*n_all][5] = the four
 * columns of the move table above + the index of the move's topology in the distinct list (the row of
 * out_costs / out_lnl that scores it).  Lets host code keep the reference's per-move rules -- branch
 * locks (rearrangement.py:228-240), enumeration order (heuristic.py:92-95 tie-breaking) -- around one
 * engine call per neighbourhood (pylogeny_b200/explore.py). */
int plg_neighbourhood_all_moves(int move_type, int n_tips, const int32_t *desc, int64_t *n_all,
                                int32_t *out_moves);
/* Canonical hash of the unrooted topology; tips are identified by tip_rows[id] (NULL: by id), so
 * the same tree written with its leaves in another order hashes the same. */
int plg_topology_hash(int n_tips, const int32_t *desc, const int32_t *tip_rows, uint64_t *out);

/* Fitch length of every distinct neighbour in one call, by edge insertion: directional state
 * sets of the base tree once, one partial per search step, one 3-way join per neighbour
 * (exact for binary trees: Fitch length is root-invariant).  *n_moves = total count M.
 * out_moves (may be NULL): int32 [M][4] = (child end of pruned branch, 0 subtree / 1 complement
 * moves, child end of target branch, SPR distance) in base-tree node ids; for TBR = (bisected
 * branch, reconnection edge in part 1 or -1 = original attachment, same for part 2, distance).
 * out_costs (may be
 * NULL): int32 [M]; only the shard [M*i/c, M*(i+1)/c) is written (ranks score disjoint blocks,
 * the host all-gathers).  Replaces the per-neighbour loop landscape.py:724-756. */
int plg_fitch_score_neighbourhood(plg_fitch_aln *aln, int move_type, int n_tips,
                                  const int32_t *desc, const int32_t *tip_rows,
                                  int64_t shard_index, int64_t shard_count,
                                  int64_t *n_moves, int32_t *out_moves,
                                  int32_t *out_costs, void *stream);

/* Same with out_costs in host or device memory (out_where).  Fitch lengths are finished on the
 * host, so a device block is sent up from pinned staging by the library. */
int plg_fitch_score_neighbourhood_out(plg_fitch_aln *aln, int move_type, int n_tips,
                                      const int32_t *desc, const int32_t *tip_rows,
                                      int64_t shard_index, int64_t shard_count,
                                      int64_t *n_moves, int32_t *out_moves,
                                      int32_t *out_costs, int out_where, void *stream);

/* ------------------------------------------------------------------ */
/* Likelihood (replaces libpllWrapper.c + the libpll calls it makes)   */
/* ------------------------------------------------------------------ */
typedef struct plg_lik_aln plg_lik_aln;
typedef struct plg_model plg_model;

/* Tip codes, row-major [n_rows][n_patterns] (1 byte per cell), pattern weights. */
int plg_lik_aln_create(int datatype, int n_rows, int64_t n_patterns, const uint8_t *codes,
                       const double *weights, plg_lik_aln **out);
void plg_lik_aln_destroy(plg_lik_aln *aln);

/* Reversible model + discrete Gamma (Yang 1994 mean categories).  exch: K(K-1)/2
 * exchangeabilities (upper triangle, row-major), freqs: K.  What pllInitModel
 * sets up at libpllWrapper.c:159. */
int plg_model_create(int n_states, const double *exch, const double *freqs, double alpha,
                     int n_cat, plg_model **out);
/* Built-in empirical protein models by the name a partition file carries (pll.py:102-115): "WAG",
 * "JTT" (case-insensitive).  exch[190] (upper triangle, row-major, order ARNDCQEGHILKMFPSTWYV) and
 * freqs[20] as plg_model_create takes them; PLG_ERR_UNSUPPORTED for other names.  Provenance of the
 * tables: csrc/protein_models.cpp. */
int plg_protein_model(const char *name, double *exch, double *freqs);
void plg_model_destroy(plg_model *m);
int plg_model_gamma_rates(const plg_model *m, double *rates /*[n_cat]*/);
/* P(t) on the host from the model's eigensystem, row-major K*K (for tests). */
int plg_model_pmatrix(const plg_model *m, double t, double *out);

/* Batched fixed-branch-length log-likelihood: descriptor as for
 * plg_fitch_score_trees plus brlen[t][i][0..1] = lengths of the two child
 * branches of internal node i.  out_lnl[t] on the host. */
int plg_lik_score_trees(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips,
                        const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                        double *out_lnl, void *stream);

/* Sharded form (see plg_fitch_score_trees_shard): only block i of c of the batch is read and scored,
 * out_lnl is indexed by tree and lives in host or device memory per out_where. */
int plg_lik_score_trees_shard(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips,
                              const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                              int64_t shard_index, int64_t shard_count, double *out_lnl,
                              int out_where, void *stream);

/* lnL of every distinct SPR/NNI neighbour of one base tree at fixed branch lengths, by edge
 * insertion (directional CLVs of the base tree, one CLV update per search step, one 3-way
 * evaluation per neighbour).  Same conventions as plg_fitch_score_neighbourhood; the closest
 * reference analogue is getSPR/NNIMovesInDistance, libpllWrapper.c:241-319. */
int plg_lik_score_neighbourhood(plg_lik_aln *aln, plg_model *m, int move_type, int n_tips,
                                const int32_t *desc, const double *brlen,
                                const int32_t *tip_rows, int64_t shard_index,
                                int64_t shard_count, int64_t *n_moves, int32_t *out_moves,
                                double *out_lnl, void *stream);

/* Same with out_lnl in host or device memory (out_where): the search walk's reduction writes the
 * shard's block straight into the caller's device buffer (no host bounce before a collective). */
int plg_lik_score_neighbourhood_out(plg_lik_aln *aln, plg_model *m, int move_type, int n_tips,
                                    const int32_t *desc, const double *brlen,
                                    const int32_t *tip_rows, int64_t shard_index,
                                    int64_t shard_count, int64_t *n_moves, int32_t *out_moves,
                                    double *out_lnl, int out_where, void *stream);

/* ---- libpllWrapper compat (libpllWrapper.c:97-239): opaque problem handle ---- */
typedef struct plg_pll plg_pll;
/* new(alignment_file, newick, partition_file): PHYLIP file, unrooted Newick, partition file
 * "DNA, p1 = 1-N" | "WAG, p1 = 1-N" (pll.py:112-115).  Input branch lengths are discarded
 * (useDefaultz, :152): every branch starts at z = 0.9.  DNA = GTR with all rates 1 and
 * empirical base frequencies, alpha = 1, 4 Gamma categories; WAG table provenance: see
 * csrc/protein_models.cpp.  The reference makes one instance per tree and re-parses the alignment
 * each time (:138); here the parsed, pattern-compressed, device-resident alignment is shared by
 * all handles on the same file (path, size, mtime, inode, device), most recent four files kept;
 * PLG_PLL_ALN_CACHE=0 turns the sharing off. */
int plg_pll_new(const char *phylip_path, const char *newick, const char *partition_path,
                plg_pll **out);
int plg_pll_destroy(plg_pll *p);
/* getLogLikelihood(handle): 3 branch-length smoothing passes, then lnL (:195-215).  MUTATES the
 * handle's branch lengths like the reference does. */
int plg_pll_loglik(plg_pll *p, double *out);
/* lnL at the current branch lengths without optimisation (pllEvaluateLikelihood, :77-78). */
int plg_pll_evaluate(plg_pll *p, double *out);
/* Batched scoring.getLogLikelihood (scoring.py:41-75: new / getLogLikelihood / getNewickString /
 * destroy per tree): every tree starts at z = 0.9, gets up to `passes` smoothing passes (the
 * reference: 3) and its lnL.  Trees must hold the alignment's taxa.  DNA and WAG (4 Gamma categories)
 * run all trees in lockstep on the device; *newick_bytes = room needed by plg_pll_batch_newicks, which copies the optimised
 * trees of the calling thread's last batch, NUL-separated, in input order. */
int plg_pll_loglik_batch(const char *phylip_path, const char *partition_path, int64_t n_trees,
                         const char *const *newicks, int passes, double *out_lnl,
                         size_t *newick_bytes);
int plg_pll_batch_newicks(char *buf, size_t bytes);
/* getNewickString(handle) (:217-239): unrooted Newick with the current branch lengths.  *len in:
 * buffer size, out: bytes needed incl. NUL; call with buf == NULL to size the buffer. */
int plg_pll_newick(plg_pll *p, char *buf, size_t *len);
/* getSPRMovesInDistance / getNNIMovesInDistance / getAllMovesInDistance(handle, maxdist)
 * (:241-319, list built at :36-93): every distinct neighbour of the handle's current tree whose
 * regraft point is within maxdist branches, with its lnL at the current branch lengths and its
 * Newick string, best lnL first.  which: PLG_MOVE_SPR, PLG_MOVE_NNI or PLG_MOVES_ALL (distance-1
 * moves are labelled NNI, the rest SPR).  compute scores the neighbourhood and keeps the list in
 * the handle; fetch copies it out: kinds[M], lnl[M] (float, like :54), M NUL-terminated strings
 * back to back in newick_buf. */
#define PLG_MOVES_ALL 4
int plg_pll_moves_compute(plg_pll *p, int which, int maxdist, int64_t *n_moves, size_t *newick_bytes);
int plg_pll_moves_fetch(plg_pll *p, int32_t *kinds, float *lnl, char *newick_buf, size_t newick_bytes);
int plg_pll_info(plg_pll *p, int *n_states, int64_t *n_patterns, double *freqs, double *rates);

#ifdef __cplusplus
}
#endif
#endif
