// Plan object: device-resident tables + workspace for one (grid, degree, index set) combination.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <atomic>
#include <vector>

#include "hb_index.h"
#include "hb_internal.h"

namespace hb {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int invalid(const char* where);   // sets "<where>: invalid arguments" and returns HB_ERR_INVALID
void note_launch(int n = 1);
int launch_status();  // counts one launch and maps cudaGetLastError() to a status

#define HB_CUDA(call)                                        \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)
#define HB_TRY(call)            \
    do {                        \
        int rc__ = (call);      \
        if (rc__ != HB_OK) return rc__; \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf();
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    int ensure(size_t need);  // grow-only
    int ensure_zero(size_t need, cudaStream_t s);
    void release();
    double* d() const { return static_cast<double*>(p); }
    int* i() const { return static_cast<int*>(p); }
};

// bumped whenever a DevBuf frees its memory (growth, release, plan destruction): a captured CUDA graph that baked
// device addresses in is valid only while this counter has the value it had at capture
extern std::atomic<unsigned long long> g_devbuf_generation;

inline long long even_up(long long v) { return (v + 1) & ~1LL; }

struct Axis {
    int n = 0, P = 0, Pp = 0, np = 0, fourier = 0;
    std::vector<double> h_x, h_w;
    DevBuf x, w;
    DevBuf plain, wgt;    // [n][Pp]   phi_k(x_i), w_i*phi_k(x_i)
    DevBuf plainT, wgtT;  // [P][np]
    DevBuf psi;           // [n][Pp]   sqrt(w_i)*phi_k(x_i), built on first Gram-form varf
    DevBuf sw;            // optional: sqrt(w_i) supplied in extended range (plan_set_sqrt_weights)
    // dense 2-D varf: slots of the 8x8-blocked pair tables (slot -> basis index, -1 padding).  Hermite axes list
    // the even indices first and start the odd ones on a block boundary: `nb_even` blocks of even indices.
    std::vector<int> slots;
    DevBuf d_slots;
    int nb = 0, nb_even = 0;
    bool symmetric_nodes = false;   // x_i == -x_{n-1-i} and w_i == w_{n-1-i} bit for bit (non-Fourier axis)
    // per-parity operand tables of the transform on a symmetric node set (built on first use), with nh = ceil(n/2)
    // non-negative nodes lo + h and Pe even / Po odd basis functions:
    //   pWe / pWo     [nh][Pe|Po]  w*phi        forward, this axis last  (B operand)
    //   pWeT / pWoT   [Pe|Po][nh]  w*phi        forward, other axes      (A operand)
    //   pPe / pPo     [Pe|Po][nh]  phi          inverse, this axis last  (B operand)
    //   pPhiE / pPhiO [nh][Pe|Po]  phi          inverse, other axes      (A operand)
    DevBuf pWe, pWo, pWeT, pWoT, pPe, pPo, pPhiE, pPhiO;
    bool par_tables = false;
    // the same tables with both parities in ONE buffer of uniform shape (d >= 2 transforms run the even and the odd
    // half of a stage as the two low-level batch entries of one GEMM launch); entry 1 (odd) is zero-padded to the
    // shape of entry 0:  cW / cPhi [2][nh][Pep],  cWT / cPhiT [2][Pe][nhp]
    DevBuf cW, cWT, cPhi, cPhiT;
    bool cpar_tables = false;
};

// A device-resident CSR result and the scratch it is built from (hb_sparse.cu): candidates (keys = row << 32 | column,
// ~0 dropped; vals), their sorted copies (keys_sorted; data doubles as the value array of the result), indptr (int64,
// rows + 1), indices (int32).
struct CsrResult {
    DevBuf keys, vals, keys_sorted, data, indices, indptr, offsets, temp, scalars;
    long long rows = -1, nnz = 0;
};
int csr_from_candidates(long long n_rows, long long n_cand, CsrResult& out, cudaStream_t st);
struct SpTensorizeArgs {
    int k, n_out;
    const long long* indptr[8];
    const int* indices[8];
    const double* data[8];
    const int* map[8];            // full-set position -> row of factor j
    const long long* col_lex[8];  // column of factor j -> its contribution to the lexicographic box offset
    const int* lex_to_pos;        // box offset -> position in the full set (-1 outside)
};
int tensorize_csr(const SpTensorizeArgs& a, CsrResult& out, cudaStream_t st);
int relabel_csr(int n_rows_in, long long nnz_in, const long long* indptr, const int* indices, const double* data,
                const int* row_map, const int* col_map, const double* col_fac, long long n_rows_out, CsrResult& out,
                cudaStream_t st);

}  // namespace hb

struct hb_plan {
    bool varf_natural_axis1 = false;   // dim 2: slots of axis 1 in natural order (see plan_create)
    int dim = 0, degree = 0, set = 0;
    std::shared_ptr<const hb::IndexSet> iset;
    std::vector<hb::Axis> ax;
    long long n_polys = 0, grid_rows = 0 /* n_0..n_{d-2} */, grid_size = 0, box_rows = 0, box_size = 0;
    hb::DevBuf lut_box;  // int32[box_size]: padded lexicographic box -> position in the set, -1 absent
    // Parity-sorted boxes: when the contraction of axis a is split by parity, its coefficient index runs
    // [even indices | odd indices] in every intermediate (the half GEMMs then write / read contiguous blocks);
    // one look-up table per set of sorted axes (bit a of the key), built on first use.
    struct SortedBox { hb::DevBuf lut; long long pitch = 0, size = 0; };
    std::map<int, SortedBox> sorted_boxes;
    hb::DevBuf ws[2];
    hb::DevBuf parA, parB;   // work buffers of the parity stages (even / odd halves)
    hb::DevBuf nodes_cat;  // all axes' nodes, concatenated (discretize)
    // --- varf state (lazy) ---
    hb_plan* plan2 = nullptr;   // transform plan at degree 2*degree (Hf), owned
    int Tdeg = -1, TPp = 0;
    hb::DevBuf T, Tf;           // triple products [2N+1][N+1][TPp] (Hermite / Fourier)
    hb::DevBuf midx, blkpos, blkperm, blkperm2, blkrange, Hf, Hc, G, kept_alpha, kept_val, pair[2], tmp;
    std::vector<hb::DevBuf> pairN;  // d >= 3 Gram varf: psi_m psi_n per axis, [n_a][P_a * Pp_a]
    hb::DevBuf pairB[2], TB[2];     // blocked pair tables (dim 2 dense varf): Gram / triple products per axis
    hb::DevBuf fsplit;              // even / odd parts of f along axis 0 (parity-halved Gram contraction)
    bool pairB_ok[2] = {false, false}, TB_ok[2] = {false, false};
    std::vector<double> h_Hf;
    std::vector<int> h_blkrange;      // host copy of blkrange (tile lists of the stage-2 kernel)
    hb::DevBuf tiles;                 // needed tiles of the last (row range, symmetric) request
    long long tiles_key[3] = {-1, -1, -1};
    int n_tiles = 0;
    int last_n_kept = 0, last_path = 0;
    // reference-algebra varf: speculative zero-fill of the output on a side stream while Hf is transformed,
    // fetched and thresholded on the host (taken when the previous call on this plan ended in the banded assembly)
    cudaStream_t fill_stream = nullptr;
    cudaEvent_t fill_ready = nullptr, fill_done = nullptr;
    double* h_Hf_pinned = nullptr;     // page-locked copy of Hf for the host-side threshold
    size_t h_Hf_pinned_len = 0;
    // optional device timing (CUDA events on the launch stream) of every GEMM launch of the last call
    struct Timed { cudaEvent_t e0 = nullptr, e1 = nullptr; int tag = 0; double flops = 0; };
    bool timing = false;
    std::vector<Timed> timed;
    int n_timed = 0;
    ~hb_plan();
};

namespace hb {
int plan_create(hb_plan** out, int dim, const int* n_pts, int degree, const double* nodes, const double* weights,
                const int* do_fourier, int set);
int plan_forward(hb_plan* pl, const double* d_f, long long f_stride, int batch, double* d_coef,
                 long long coef_stride, cudaStream_t st);
int plan_backward(hb_plan* pl, const double* d_coef, long long coef_stride, int batch, double* d_f,
                  long long f_stride, cudaStream_t st);
// reference-algebra varf as CSR without a dense matrix when the retained alpha leave a narrow band (polynomial f);
// out.rows = -1 when they do not (nothing assembled: use plan_varf + a compaction)
int plan_set_sqrt_weights(hb_plan* pl, int axis, const double* sqrt_w);
int plan_varf_csr(hb_plan* pl, const double* d_f, long long row_begin, long long row_end, CsrResult& out, cudaStream_t st);
int plan_varf(hb_plan* pl, const double* d_f, int mode, double* d_out, long long ldo, long long row_begin,
              long long row_end, cudaStream_t st);
int ensure_triple_products(hb_plan* pl, cudaStream_t st);
int ensure_combined_parity_tables(Axis& A, cudaStream_t st);
// GEMM launch bracketed by events when pl->timing; tag: 1+axis forward, 11+axis backward, 21/22 varf stage 1/2
int timed_gemm(hb_plan* pl, const GemmArgs& g, int tag, cudaStream_t st);
void fourier_triple_products_host(int degree, int Pp, std::vector<double>& out);
int discretize_device(const char* body, int dim, const int* n_pts, const double* d_nodes, const double* translation,
                      const double* dilation, double* d_out, long long pitch_last, cudaStream_t st);
}  // namespace hb
