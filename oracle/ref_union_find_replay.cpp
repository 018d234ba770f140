// ref_union_find_replay.cpp -- drives the REFERENCE's own looper/union_find.h (compiled in place
// with -DALPS_INDEP_SOURCE -I$(REF), see oracle/Makefile) through the sequence of
// test/union_find.C:40-74, with std::mt19937 standing in for boost::mt19937 (same generator) and
// eng()/2^32 for boost::uniform_real<>.  Output must equal test/union_find.op.
// TEST INFRASTRUCTURE ONLY; built into oracle/_ref/.
#include <looper/union_find.h>
#include <iostream>
#include <random>
#include <vector>
int main() {
  const int n = 100;
  std::mt19937 eng(29833u);
  auto rng = [&]() { return eng() / 4294967296.0; };
  std::cout << "[[union find test]]\n";
  std::vector<looper::union_find::node> nodes(n);
  std::vector<looper::union_find::node_noweight> nodes_noweight(n);
  std::cout << "\n[making tree]\n";
  for (int i = 0; i < n; i++) {
    int i0 = static_cast<int>(n * rng());
    int i1 = static_cast<int>(n * rng());
    std::cout << "connecting node " << i0 << " to node " << i1 << std::endl;
    looper::union_find::unify(nodes, i0, i1);
    looper::union_find::unify(nodes_noweight, i0, i1);
  }
  std::cout << "\n[results]\n";
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < n; i++) {
      if (nodes[i].is_root()) {
        if (pass == 0) std::cout << "node " << i << " is root and tree size is " << nodes[i].weight() << std::endl;
        else std::cout << "node " << i << " is root\n";
      } else
        std::cout << "node " << i << "'s parent is " << nodes[i].parent() << " and its root is "
                  << looper::union_find::root_index(nodes, i) << std::endl;
    }
}
