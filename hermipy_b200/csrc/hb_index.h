// Multi-index sets of the Hermite-Galerkin basis (host side).  The enumeration ORDER is part of the
// reference's data format -- coefficient vectors and matrices are laid out in it -- so it is reproduced
// bit-exactly (reference: cpp/src/hermite/iterators.cpp:106-596), but generated here from a declarative
// description of each order (shells + mixed-radix / factorisation blocks) instead of a stateful
// successor function, and indexed by a dense look-up table over the bounding box instead of a
// std::string-keyed hash map (lib.cpp:76-84).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace hb {

enum IndexSetId : int { SET_TRIANGLE = 0, SET_CROSS = 1, SET_CROSS_NC = 2, SET_CUBE = 3, SET_RECTANGLE = 4 };

int index_set_from_name(const char* name);  // -1 if unknown

struct IndexSet {
    int set = 0, dim = 0, degree = 0;
    uint32_t size = 0;
    std::vector<uint32_t> list;  // size x dim, row-major, reference enumeration order
    // Dense inverse map over the box [0, extent_d)^dim, lexicographic with axis 0 slowest; -1 = absent.
    std::vector<uint32_t> extent;
    std::vector<int32_t> lex_to_pos;
    int32_t position(const uint32_t* m) const;  // -1 if m is not in the set
};

// Cached (per process) construction; returns nullptr for invalid arguments
// (unknown set, dim < 1, or cross with degree 0 which the reference rejects by assert).
std::shared_ptr<const IndexSet> get_index_set(int set, int dim, int degree);

// Number of multi-indices, following the reference's s_size (including its quirks); -1 on error.
long long index_size(int set, int dim, int degree);

// Inverse of index_size.  Returns 0 and writes *degree, or nonzero if no degree matches
// (the reference prints "Can't find x" and exits, lib.cpp:93-99).
int index_get_degree(int set, int dim, long long n_polys, int* degree);

// Closed-form positions (reference exposes only these two: module.cpp:145-146).
uint32_t triangle_index(const uint32_t* m, int dim);
uint32_t cube_index(const uint32_t* m, int dim);

}  // namespace hb
