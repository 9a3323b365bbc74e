// TEST INFRASTRUCTURE (oracle/_ref). A plain C ABI around the UNMODIFIED reference tree
// engine so tests and the CPU-baseline leg of bench.py can drive it through ctypes without
// Cython.  The reference sources are compiled from where they lie under /root/reference
// (-I core/ctree); like the reference's own ctree.pxd:5,27 this pulls the two .cpp files into
// one translation unit.  Nothing here re-implements tree logic: every call forwards to
// tree::CRoots / tree::cmulti_traverse / tree::cmulti_back_propagate (cnode.cpp).
#include "cminimax.cpp"
#include "cnode.cpp"

#include <cstring>
#include <vector>

namespace {
struct RefTrees {
  tree::CRoots* roots;
  tools::CMinMaxStatsList* minmax;
  tree::CSearchResults* results;
  int num, actions;
};

std::vector<std::vector<float>> rows_f(const float* p, int n, int a) {
  std::vector<std::vector<float>> v(n);
  for (int i = 0; i < n; ++i) v[i].assign(p + (size_t)i * a, p + (size_t)(i + 1) * a);
  return v;
}
std::vector<std::vector<int>> rows_i(const int* p, int n, int a) {
  std::vector<std::vector<int>> v(n);
  for (int i = 0; i < n; ++i) v[i].assign(p + (size_t)i * a, p + (size_t)(i + 1) * a);
  return v;
}
}  // namespace

extern "C" {

// mirrors cytree.Roots.__cinit__ (cytree.pyx:42-45): pool_size = action_num * (tree_nodes + 2)
void* ref_trees_new(int num, int actions, int tree_nodes, float value_delta_max) {
  RefTrees* t = new RefTrees;
  t->num = num;
  t->actions = actions;
  t->roots = new tree::CRoots(num, actions, actions * (tree_nodes + 2));
  t->minmax = new tools::CMinMaxStatsList(num);
  t->minmax->set_delta(value_delta_max);
  t->results = nullptr;
  return t;
}

void ref_trees_free(void* h) {
  RefTrees* t = (RefTrees*)h;
  delete t->results;
  delete t->minmax;
  delete t->roots;
  delete t;
}

// noises == NULL -> prepare_no_noise
void ref_trees_prepare(void* h, float frac, const float* noises, const float* rewards,
                       const float* logits, const int* masks) {
  RefTrees* t = (RefTrees*)h;
  std::vector<float> r(rewards, rewards + t->num);
  auto pol = rows_f(logits, t->num, t->actions);
  auto leg = rows_i(masks, t->num, t->actions);
  if (noises) {
    auto nz = rows_f(noises, t->num, t->actions);
    t->roots->prepare(frac, nz, r, pol, leg);
  } else {
    t->roots->prepare_no_noise(r, pol, leg);
  }
}

void ref_trees_traverse(void* h, int pb_c_base, float pb_c_init, float discount, int* out_ix,
                        int* out_iy, int* out_action) {
  RefTrees* t = (RefTrees*)h;
  delete t->results;
  t->results = new tree::CSearchResults(t->num);
  tree::cmulti_traverse(t->roots, pb_c_base, pb_c_init, discount, t->minmax, *t->results);
  std::memcpy(out_ix, t->results->hidden_state_index_x_lst.data(), sizeof(int) * t->num);
  std::memcpy(out_iy, t->results->hidden_state_index_y_lst.data(), sizeof(int) * t->num);
  std::memcpy(out_action, t->results->last_actions.data(), sizeof(int) * t->num);
}

int ref_trees_path_len(void* h, int i) {
  RefTrees* t = (RefTrees*)h;
  return (int)t->results->search_paths[i].size();
}

void ref_trees_backprop(void* h, int x, float discount, const float* rewards, const float* values,
                        const float* logits) {
  RefTrees* t = (RefTrees*)h;
  std::vector<float> r(rewards, rewards + t->num), v(values, values + t->num);
  auto pol = rows_f(logits, t->num, t->actions);
  tree::cmulti_back_propagate(x, discount, r, v, pol, t->minmax, *t->results);
}

void ref_trees_stats(void* h, int* out_visits, float* out_values, float* out_minmax) {
  RefTrees* t = (RefTrees*)h;
  auto d = t->roots->get_distributions();
  auto v = t->roots->get_values();
  for (int i = 0; i < t->num; ++i) {
    for (int a = 0; a < t->actions; ++a) out_visits[(size_t)i * t->actions + a] = d[i][a];
    out_values[i] = v[i];
    if (out_minmax) {
      out_minmax[2 * i] = t->minmax->stats_lst[i].minimum;
      out_minmax[2 * i + 1] = t->minmax->stats_lst[i].maximum;
    }
  }
}

// root children priors (for checking prepare): child a of root i
void ref_trees_root_priors(void* h, float* out_priors) {
  RefTrees* t = (RefTrees*)h;
  for (int i = 0; i < t->num; ++i)
    for (int a = 0; a < t->actions; ++a)
      out_priors[(size_t)i * t->actions + a] = t->roots->roots[i].get_child(a)->prior;
}

// best-action chains (CRoots::get_trajectories, cnode.cpp:266-274); out is [num][max_len], -1 padded
void ref_trees_trajectories(void* h, int* out, int max_len) {
  RefTrees* t = (RefTrees*)h;
  auto tr = t->roots->get_trajectories();
  for (int i = 0; i < t->num; ++i)
    for (int k = 0; k < max_len; ++k)
      out[(size_t)i * max_len + k] = k < (int)tr[i].size() ? tr[i][k] : -1;
}

}  // extern "C"

// per expanded non-root node of tree i, indexed by hidden_state_index_x - 1:
// (reward, value_sum, visit_count).  Returns the largest x seen.
extern "C" int ref_trees_expanded_stats(void* h, int i, float* out_reward, float* out_value_sum,
                                        int* out_visits, int cap) {
  RefTrees* t = (RefTrees*)h;
  int n_out = 0;
  for (tree::CNode& n : t->roots->node_pools[i]) {
    if (!n.expanded()) continue;
    int x = n.hidden_state_index_x;
    if (x >= 1 && x <= cap) {
      out_reward[x - 1] = n.reward;
      out_value_sum[x - 1] = n.value_sum;
      out_visits[x - 1] = n.visit_count;
      if (x > n_out) n_out = x;
    }
  }
  return n_out;
}
