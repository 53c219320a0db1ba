// Multi-index set enumeration, from a declarative description of each order.  Reference behaviour
// being matched (bit-exactly): cpp/src/hermite/iterators.cpp -- Cube :106-217, Triangle :222-321,
// Cross :326-421, Cross_nc :424-486, Rectangle :489-596.  See hb_index.h.
#include "hb_index.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace hb {

int index_set_from_name(const char* name) {
    if (!name) return -1;
    if (!strcmp(name, "triangle")) return SET_TRIANGLE;
    if (!strcmp(name, "cross")) return SET_CROSS;
    if (!strcmp(name, "cross_nc")) return SET_CROSS_NC;
    if (!strcmp(name, "cube")) return SET_CUBE;
    if (!strcmp(name, "rectangle")) return SET_RECTANGLE;
    return -1;
}

namespace {

typedef std::vector<uint32_t> Flat;

// ---- graded (total-degree) order ------------------------------------------------------------
// Shell g = |m| ascending; inside a shell the last coordinate runs from g down to 0, then the
// second-to-last over what is left, ... ; coordinate 0 takes the remainder.
void graded_shell(int pos, uint32_t remaining, std::vector<uint32_t>& m, Flat& out) {
    if (pos == 0) {
        m[0] = remaining;
        out.insert(out.end(), m.begin(), m.end());
        return;
    }
    for (uint32_t v = remaining + 1; v-- > 0;) {
        m[pos] = v;
        graded_shell(pos - 1, remaining - v, m, out);
    }
}

void list_triangle(int dim, int degree, Flat& out) {
    std::vector<uint32_t> m(dim, 0);
    for (int g = 0; g <= degree; g++) graded_shell(dim - 1, (uint32_t)g, m, out);
}

// ---- weighted max-norm shells: cube (weights 1) and rectangle (weights 1,2,2,...) -------------
// Shell value s = max_i w_i m_i ascending.  Inside a shell, blocks by the LAST axis `a` that attains
// the maximum (only axes with w_a | s), and inside a block the other axes form a mixed-radix counter
// with axis 0 fastest: axes before `a` satisfy w_i m_i <= s, axes after it w_i m_i < s.
void list_weighted_max(int dim, int degree, const std::vector<uint32_t>& w, Flat& out) {
    std::vector<uint32_t> m(dim), radix(dim);
    for (uint32_t s = 0; s <= (uint32_t)degree; s++) {
        for (int a = 0; a < dim; a++) {
            if (s % w[a]) continue;
            bool empty = false;
            for (int i = 0; i < dim; i++) {
                radix[i] = (i < a) ? s / w[i] + 1 : (s + w[i] - 1) / w[i];
                if (i != a && radix[i] == 0) empty = true;
            }
            if (empty) continue;
            std::fill(m.begin(), m.end(), 0u);
            m[a] = s / w[a];
            while (true) {
                out.insert(out.end(), m.begin(), m.end());
                int i = 0;
                for (; i < dim; i++) {
                    if (i == a) continue;
                    if (m[i] + 1 < radix[i]) { m[i]++; break; }
                    m[i] = 0;
                }
                if (i == dim) break;
            }
        }
    }
}

// ---- hyperbolic cross: prod max(m_i,1) <= degree ----------------------------------------------
// Shell p = prod max(m_i,1) ascending.  Inside a shell, the ordered factorisations p = f_0...f_{d-1}
// with the last factor descending, then the one before it, ...; every factor equal to 1 stands for
// the two values m_i in {0,1}, expanded as a binary counter whose lowest axis is fastest.
void cross_emit(const std::vector<uint32_t>& f, Flat& out) {
    const int dim = (int)f.size();
    std::vector<int> ones;
    for (int i = 0; i < dim; i++)
        if (f[i] == 1) ones.push_back(i);
    std::vector<uint32_t> m(f);
    const uint64_t n = 1ull << ones.size();
    for (uint64_t mask = 0; mask < n; mask++) {
        for (size_t b = 0; b < ones.size(); b++) m[ones[b]] = (uint32_t)((mask >> b) & 1u);
        out.insert(out.end(), m.begin(), m.end());
    }
}

void cross_factorise(int pos, uint32_t rem, std::vector<uint32_t>& f, Flat& out) {
    if (pos == 0) {
        f[0] = rem;
        cross_emit(f, out);
        return;
    }
    for (uint32_t v = rem; v >= 1; v--) {
        if (rem % v) continue;
        f[pos] = v;
        cross_factorise(pos - 1, rem / v, f, out);
    }
}

void list_cross(int dim, int degree, Flat& out) {
    std::vector<uint32_t> f(dim, 1);
    for (uint32_t p = 1; p <= (uint32_t)degree; p++) cross_factorise(dim - 1, p, f, out);
}

// ---- "non-consistent" cross: graded order restricted to prod max(m_i,1) <= max(degree,1) -------
void list_cross_nc(int dim, int degree, Flat& out) {
    const uint32_t cap = (uint32_t)std::max(degree, 1);
    Flat tri;
    list_triangle(dim, (int)cap + dim - 1, tri);
    for (size_t i = 0; i < tri.size(); i += dim) {
        uint64_t prod = 1;
        for (int d = 0; d < dim; d++) prod *= std::max<uint32_t>(tri[i + d], 1u);
        if (prod <= cap) out.insert(out.end(), tri.begin() + i, tri.begin() + i + dim);
    }
}

bool build_list(int set, int dim, int degree, Flat& out) {
    if (dim < 1 || degree < 0) return false;
    switch (set) {
        case SET_TRIANGLE: list_triangle(dim, degree, out); return true;
        case SET_CUBE: list_weighted_max(dim, degree, std::vector<uint32_t>(dim, 1u), out); return true;
        case SET_RECTANGLE: {
            std::vector<uint32_t> w(dim, 2u);
            w[0] = 1u;
            list_weighted_max(dim, degree, w, out);
            return true;
        }
        case SET_CROSS:
            if (degree < 1) return false;  // reference: assert(degree > 0), iterators.cpp:328
            list_cross(dim, degree, out);
            return true;
        case SET_CROSS_NC: list_cross_nc(dim, degree, out); return true;
    }
    return false;
}

std::mutex g_mutex;
std::map<std::tuple<int, int, int>, std::shared_ptr<const IndexSet>> g_cache;

uint64_t binom(uint64_t n, uint64_t k) {
    if (k > n) return 0;
    k = std::min(k, n - k);
    unsigned __int128 r = 1;
    for (uint64_t i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return (uint64_t)r;
}

}  // namespace

int32_t IndexSet::position(const uint32_t* m) const {
    size_t lex = 0;
    for (int d = 0; d < dim; d++) {
        if (m[d] >= extent[d]) return -1;
        lex = lex * extent[d] + m[d];
    }
    return lex_to_pos[lex];
}

std::shared_ptr<const IndexSet> get_index_set(int set, int dim, int degree) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto key = std::make_tuple(set, dim, degree);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    auto s = std::make_shared<IndexSet>();
    s->set = set; s->dim = dim; s->degree = degree;
    if (!build_list(set, dim, degree, s->list)) return nullptr;
    s->size = (uint32_t)(s->list.size() / dim);
    s->extent.assign(dim, 0);
    for (uint32_t i = 0; i < s->size; i++)
        for (int d = 0; d < dim; d++) s->extent[d] = std::max(s->extent[d], s->list[(size_t)i * dim + d] + 1);
    size_t box = 1;
    for (int d = 0; d < dim; d++) box *= s->extent[d];
    s->lex_to_pos.assign(box, -1);
    for (uint32_t i = 0; i < s->size; i++) {
        size_t lex = 0;
        for (int d = 0; d < dim; d++) lex = lex * s->extent[d] + s->list[(size_t)i * dim + d];
        s->lex_to_pos[lex] = (int32_t)i;
    }
    if (g_cache.size() > 64) g_cache.clear();
    g_cache[key] = s;
    return s;
}

long long index_size(int set, int dim, int degree) {
    if (dim < 1 || degree < 0) return -1;
    switch (set) {
        case SET_TRIANGLE: return (long long)binom((uint64_t)degree + dim, dim);
        case SET_CUBE: {
            long long r = 1;
            for (int d = 0; d < dim; d++) r *= (degree + 1);
            return r;
        }
        case SET_RECTANGLE: {
            long long r = 1;
            for (int d = 0; d < dim; d++) r *= 1 + degree / (d == 0 ? 1 : 2);
            return r;
        }
        case SET_CROSS:
        case SET_CROSS_NC: {
            auto s = get_index_set(set, dim, degree);
            return s ? (long long)s->size : -1;
        }
    }
    return -1;
}

int index_get_degree(int set, int dim, long long n_polys, int* degree) {
    if (dim < 1 || n_polys < 1 || set < 0 || set > SET_RECTANGLE) return 1;
    if (set == SET_CROSS) {
        // Reference (iterators.cpp:411-416): first coordinate of entry n_polys-1 of the enumeration.
        // Shells are appended in ascending product order, so grow the degree until it is covered.
        for (int deg = 1;; deg = deg * 2) {
            auto s = get_index_set(SET_CROSS, dim, deg);
            if (!s) return 1;
            if ((long long)s->size >= n_polys) {
                *degree = (int)s->list[(size_t)(n_polys - 1) * dim];
                return 0;
            }
            if (deg > (1 << 20)) return 1;
        }
    }
    // monotone size -> smallest degree whose size reaches n_polys
    int lo = 0, hi = 1;
    while (index_size(set, dim, hi) < n_polys) {
        lo = hi;
        hi *= 2;
        if (hi > (1 << 24)) return 1;
    }
    if (index_size(set, dim, lo) == n_polys) { *degree = lo; return 0; }
    while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        if (index_size(set, dim, mid) < n_polys) lo = mid; else hi = mid;
    }
    if (index_size(set, dim, hi) != n_polys) return 1;
    *degree = hi;
    return 0;
}

// Position in the graded order: sum over i of C(i + s_i, i + 1), s_i = m_0 + ... + m_i.
uint32_t triangle_index(const uint32_t* m, int dim) {
    uint64_t sum = 0, r = 0;
    for (int i = 0; i < dim; i++) {
        sum += m[i];
        if (sum > 0) r += binom(i + sum, i + 1);
    }
    return (uint32_t)r;
}

// Position in the max-norm shell order: s^d entries precede shell s; then the blocks of the axes
// before `a` (the last axis attaining s), then the mixed-radix offset inside the block.
uint32_t cube_index(const uint32_t* m, int dim) {
    uint64_t s = 0;
    int a = 0;
    for (int i = 0; i < dim; i++)
        if (m[i] >= s) { s = m[i]; a = i; }
    auto ipow = [](uint64_t b, int e) { uint64_t r = 1; while (e-- > 0) r *= b; return r; };
    uint64_t r = ipow(s, dim);
    for (int i = 0; i < a; i++) r += ipow(s, dim - i - 1) * ipow(s + 1, i);
    uint64_t unit = 1;
    for (int i = 0; i < dim; i++) {
        if (i == a) continue;
        r += unit * m[i];
        unit *= (i < a) ? s + 1 : s;
    }
    return (uint32_t)r;
}

}  // namespace hb
