// oracle shim (test infrastructure only): F/KSearchTree.h:12-14 only needs these names to
// parse; the k-d tree is never reached from the assembly+solve path (ks_stub.cpp).
#pragma once
namespace CGAL {
template <class T> struct Kernel_traits;
template <class T> class Kd_tree_rectangle;
template <class A, class B, class C, class D> struct Search_traits {};
template <class T> struct Orthogonal_k_neighbor_search { struct Tree {}; };
}
