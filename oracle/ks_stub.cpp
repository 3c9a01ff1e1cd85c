// oracle support (test infrastructure only): link-time stubs for KSearchTree
// (declared F/KSearchTree.h:128-132). Only reached from Mesh::findCommonFaces /
// findCommonNodes (F/Mesh.cpp:877-1120), which the assembly+solve path never calls.
#include "KSearchTree.h"
#include <cstdlib>
KSearchTree::KSearchTree() {}
KSearchTree::KSearchTree(const Vec3DArray&) {}
void KSearchTree::insert(const Vec3D&, const int) { std::abort(); }
void KSearchTree::findNeighbors(const Vec3D&, const int, Array<int>&) { std::abort(); }
