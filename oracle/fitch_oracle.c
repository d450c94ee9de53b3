/* TEST INFRASTRUCTURE ONLY -- not product code.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this.
 *
 * Plain-C restatement of the reference's Fitch small-parsimony path
 * (/root/reference/pylogeny/fitch.cpp).  Each function cites the lines it
 * follows.  Parity is PINNED: tests/test_oracle_cpu.py checks it against the
 * golden vectors in tests/golden/ (generated from the reference's own code by
 * tests/golden/make_golden.py) and, when oracle/_ref/libfitch_ref.so exists,
 * against the compiled reference core on random inputs.
 *
 * Domain fence (SURVEY.md 8a "a7-detail"): Newick strings on which the
 * reference parser terminates -- no length-less subtree after a sibling
 * subtree that carried a length/label (stale `q`, fitch.cpp:154,175), no
 * negative branch lengths (stof("") at fitch.cpp:133).  Inside the fence this
 * parser builds the same child order as fitch.cpp:148-191.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct ONode {
  int first_child, next_sibling, n_children; /* child list in string order (fitch.cpp:40) */
  int label_start, label_len;                /* into the newick string (fitch.cpp:39) */
} ONode;

typedef struct {
  const char *s;
  int len;
  ONode *nodes;
  int n_nodes, cap;
  int err;
} OParser;

static int new_node(OParser *p) {
  if (p->n_nodes == p->cap) {
    p->cap = p->cap ? 2 * p->cap : 64;
    p->nodes = (ONode *)realloc(p->nodes, sizeof(ONode) * (size_t)p->cap);
  }
  ONode *n = &p->nodes[p->n_nodes];
  n->first_child = n->next_sibling = -1;
  n->n_children = 0;
  n->label_start = 0;
  n->label_len = 0;
  return p->n_nodes++;
}

static void add_child(OParser *p, int parent, int child) { /* fitch.cpp:40 push_back */
  ONode *pn = &p->nodes[parent];
  if (pn->first_child < 0) pn->first_child = child;
  else {
    int c = pn->first_child;
    while (p->nodes[c].next_sibling >= 0) c = p->nodes[c].next_sibling;
    p->nodes[c].next_sibling = child;
  }
  pn->n_children++;
}

/* fitch.cpp:118-125 getBalancingBracket */
static int balancing(const OParser *p, int a) {
  int open = 0, i = a;
  while (i + 1 < p->len) {
    if (p->s[i + 1] == '(') ++open;
    else if (p->s[i + 1] == ')' && --open < 0) return i + 1;
    ++i;
  }
  return -1;
}

/* fitch.cpp:127-135 readBranchLength: digits and '.' only; value unused by cost() */
static int skip_length(const OParser *p, int a) {
  int i = a;
  while (i + 1 < p->len) {
    char c = p->s[i + 1];
    if (!((c >= '0' && c <= '9') || c == '.')) break;
    ++i;
  }
  return i;
}

/* fitch.cpp:137-146 readLeafName */
static int read_name(const OParser *p, int a, int *start, int *len) {
  int i = a;
  while (i < p->len) {
    char c = p->s[i];
    if (c == ',' || c == ':' || c == ';' || c == ')') {
      *start = a;
      *len = i - a;
      return i - 1;
    }
    ++i;
  }
  return -1;
}

/* fitch.cpp:148-191 parse(i,j,cursor) */
static void parse_range(OParser *p, int i, int j, int cursor) {
  while (i <= j && !p->err) {
    int k, q = -1; /* reference keeps q across siblings: outside the fence */
    int nn = new_node(p);
    if (cursor >= 0) add_child(p, cursor, nn);
    if (p->s[i] == '(') {
      k = balancing(p, i);
      if (k < 0) { p->err = 1; return; }
      if (k + 1 < p->len && p->s[k + 1] == ':') q = skip_length(p, k + 1);
      else if (k + 1 < p->len && p->s[k + 1] != ',') {
        int st = 0, ln = 0;
        q = read_name(p, k + 1, &st, &ln);
        if (q < 0) { st = k + 1; ln = p->len - (k + 1); q = p->len - 1; }
        p->nodes[nn].label_start = st;
        p->nodes[nn].label_len = ln;
        if (q + 1 < p->len && p->s[q + 1] == ':') q = skip_length(p, q + 1);
      }
      parse_range(p, i + 1, k, nn);
      if (q != -1) k = q;
    } else {
      int st = 0, ln = 0;
      k = read_name(p, i, &st, &ln);
      if (k < 0) { st = i; ln = p->len - i; k = p->len - 1; }
      if (k + 1 <= j && p->s[k + 1] == ':') k = skip_length(p, k + 1);
      p->nodes[nn].label_start = st;
      p->nodes[nn].label_len = ln;
    }
    i = k + 1;
    if (i < j) while (i + 1 <= j && p->s[i] != ',') ++i;
    ++i;
  }
}

/* state sets: the reference holds sorted strings of chars (fitch.cpp:197-213);
 * a 256-bit membership mask is the same set algebra for any char value. */
typedef struct { uint64_t w[4]; } OSet;
static int set_empty(const OSet *a) { return !(a->w[0] | a->w[1] | a->w[2] | a->w[3]); }

/* fitch.cpp:70-74,85-91 post-order: children in list order, then self */
static void post_order(const OParser *p, int n, int *out, int *cnt) {
  for (int c = p->nodes[n].first_child; c >= 0; c = p->nodes[c].next_sibling)
    post_order(p, c, out, cnt);
  out[(*cnt)++] = n;
}

/* Returns 0 ok; 1 parse error; 2 unknown leaf label (reference: map::at throws,
 * fitch.cpp:255); 3 profile shorter than taxon index (reference: substr throws). */
int oracle_fitch_cost(const char *newick, int n_prof, const char *const *profiles,
                      const int64_t *weights, int n_taxa,
                      const char *const *taxa_names, const int32_t *taxa_idx,
                      int32_t *out) {
  OParser p;
  memset(&p, 0, sizeof p);
  p.s = newick;
  p.len = (int)strlen(newick);
  if (p.len == 0) return 1;
  parse_range(&p, 0, p.len - 1, -1); /* fitch.cpp:112-114 */
  if (p.err || p.n_nodes == 0) { free(p.nodes); return 1; }
  int *po = (int *)malloc(sizeof(int) * (size_t)p.n_nodes);
  int cnt = 0;
  post_order(&p, 0, po, &cnt);
  /* leaf -> taxon index (fitch.cpp:255) */
  int *tax = (int *)malloc(sizeof(int) * (size_t)p.n_nodes);
  for (int x = 0; x < p.n_nodes; x++) {
    tax[x] = -1;
    if (p.nodes[x].n_children) continue;
    for (int t = 0; t < n_taxa; t++)
      if ((int)strlen(taxa_names[t]) == p.nodes[x].label_len &&
          !strncmp(taxa_names[t], newick + p.nodes[x].label_start, (size_t)p.nodes[x].label_len)) {
        tax[x] = taxa_idx[t];
        break;
      }
    if (tax[x] < 0) { free(po); free(tax); free(p.nodes); return 2; }
  }
  OSet *data = (OSet *)malloc(sizeof(OSet) * (size_t)p.n_nodes);
  int32_t total = 0; /* fitch.cpp:247 `int total` */
  int rc = 0;
  for (int i = 0; i < n_prof && !rc; i++) { /* fitch.cpp:249 */
    int local = 0;
    size_t plen = strlen(profiles[i]);
    for (int z = 0; z < cnt; z++) {
      int x = po[z];
      ONode *nd = &p.nodes[x];
      if (nd->n_children == 0) { /* fitch.cpp:254-257: one char of the profile */
        if ((size_t)tax[x] >= plen) { rc = 3; break; }
        unsigned char ch = (unsigned char)profiles[i][tax[x]];
        memset(&data[x], 0, sizeof(OSet));
        data[x].w[ch >> 6] |= (uint64_t)1 << (ch & 63);
      } else {
        /* fitch.cpp:215-226 running intersection that RESTARTS when empty */
        OSet acc;
        memset(&acc, 0, sizeof acc);
        for (int c = nd->first_child; c >= 0; c = p.nodes[c].next_sibling) {
          if (set_empty(&acc)) acc = data[c];
          else for (int w = 0; w < 4; w++) acc.w[w] &= data[c].w[w];
        }
        if (set_empty(&acc)) { /* fitch.cpp:262-264 */
          for (int c = nd->first_child; c >= 0; c = p.nodes[c].next_sibling)
            for (int w = 0; w < 4; w++) acc.w[w] |= data[c].w[w]; /* :228-239 */
          local += nd->n_children - 1;
        }
        data[x] = acc;
      }
    }
    /* fitch.cpp:268 `total += local*weiArr[i]` (long product narrowed to int) */
    total = (int32_t)(uint32_t)((uint64_t)(int64_t)total + (uint64_t)((int64_t)local * weights[i]));
  }
  *out = total;
  free(data); free(po); free(tax); free(p.nodes);
  return rc;
}
