// union_find.h -- host-side looper::union_find API (reference: looper/union_find.h:57-136 node
// types, :145-172 root_index, :242-284 unify, :307-343 count_root/set_id/copy_id).
//
// The device side (csrc/lq_device.cuh uf_find/uf_union) is the hot path; this header keeps the
// reference's names for host code that post-processes labels (tests, tools, boundary merges on the
// host).  Linking rule: the root with the SMALLER index wins -- the reference's
// LOOPER_USE_DETERMINISTIC_UNIFY convention -- so that host and device produce the same roots.
#pragma once
#include <vector>

namespace looper {
namespace union_find {

// a root stores -(weight) (<= 0), any other node stores parent+1 (> 0)
class node {
public:
  node() : link_(-1), id_(0) {}
  bool is_root() const { return link_ <= 0; }
  int parent() const { return link_ - 1; }
  int weight() const { return -link_; }
  int id() const { return id_; }
  void set_parent(int p) { link_ = p + 1; }
  void set_weight(int w) { link_ = -w; }
  void set_id(int i) { id_ = i; }
private:
  int link_, id_;
};

// 4-byte link: a root stores ~id (< 0), any other node parent+1 (boundary links, flip tables)
class node_noweight {
public:
  node_noweight() : link_(~0) {}
  bool is_root() const { return link_ <= 0; }
  int parent() const { return link_ - 1; }
  int weight() const { return 0; }
  int id() const { return ~link_; }
  void set_parent(int p) { link_ = p + 1; }
  void set_weight(int) { link_ = ~0; }
  void set_id(int i) { link_ = ~i; }
private:
  int link_;
};

template <class T> int add(std::vector<T>& v) { v.push_back(T()); return int(v.size()) - 1; }

template <class T> int root_index(const std::vector<T>& v, int g) {
  while (!v[g].is_root()) g = v[g].parent();
  return g;
}

// find with path halving
template <class T> int root_index_ph(std::vector<T>& v, int g) {
  while (!v[g].is_root()) {
    const int p = v[g].parent();
    if (v[p].is_root()) return p;
    v[g].set_parent(v[p].parent());
    g = p;
  }
  return g;
}

template <class T> int unify(std::vector<T>& v, int g0, int g1) {
  int r0 = root_index_ph(v, g0), r1 = root_index_ph(v, g1);
  if (r0 == r1) return r0;
  if (r1 < r0) { int t = r0; r0 = r1; r1 = t; }
  v[r0].set_weight(v[r0].weight() + v[r1].weight());
  v[r1].set_parent(r0);
  return r0;
}

template <class T> const T& root(const std::vector<T>& v, int g) { return v[root_index(v, g)]; }
template <class T> int cluster_id(const std::vector<T>& v, int g) { return root(v, g).id(); }

template <class T> int count_root(const std::vector<T>& v, int start, int n) {
  int c = 0;
  for (int i = start; i < start + n; ++i) c += v[i].is_root() ? 1 : 0;
  return c;
}
// roots are numbered in array order starting at nc; returns the next free id
template <class T> int set_id(std::vector<T>& v, int start, int n, int nc) {
  for (int i = start; i < start + n; ++i)
    if (v[i].is_root()) v[i].set_id(nc++);
  return nc;
}
template <class T> void copy_id(std::vector<T>& v, int start, int n) {
  for (int i = start; i < start + n; ++i) v[i].set_id(cluster_id(v, i));
}

}  // namespace union_find
}  // namespace looper
