// TEST INFRASTRUCTURE ONLY -- not product code.
//
// Thin extern "C" driver around the reference's own Fitch core.  The reference
// translation unit /root/reference/pylogeny/fitch.cpp cannot be compiled whole
// (its CPython-2 binding, fitch.cpp:5 and :274-353, needs Python 2), so the
// Makefile extracts lines 6-272 (Node/Tree/PhylogenyParser/cost, untouched)
// into the git-ignored oracle/_ref/fitch_core.inc at build time and this
// driver #includes it.  No reference source is stored in the repository.
//
// The marshalling below restates what fitch_cost() does at fitch.cpp:280-327:
// copy profiles into string[], weights into long[], the taxa dict into
// std::map<string,int>, parse, call cost().
#include "_ref/fitch_core.inc"

extern "C" int ref_fitch_cost(const char *newick, int n_prof,
                              const char *const *profiles,
                              const long *weights, int n_taxa,
                              const char *const *taxa_names,
                              const int *taxa_idx, int *out) {
  try {
    Tree tree;
    PhylogenyParser parser(&tree, std::string(newick));
    parser.parse();
    std::string *pro = new std::string[n_prof];
    long *wei = new long[n_prof];
    for (int i = 0; i < n_prof; i++) {
      pro[i] = std::string(profiles[i]);
      wei[i] = weights[i];
    }
    std::map<std::string, int> taxmap;
    for (int j = 0; j < n_taxa; j++)
      taxmap.insert(std::pair<std::string, int>(taxa_names[j], taxa_idx[j]));
    *out = cost(&tree, n_prof, pro, wei, &taxmap);
    delete[] pro;
    delete[] wei;
    return 0;
  } catch (...) {
    return 1;  // the reference would abort here (uncaught C++ exception)
  }
}
