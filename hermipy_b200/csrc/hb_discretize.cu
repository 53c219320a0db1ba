// discretize: evaluate a function, given as the body of a C expression in v[0..dim-1], on the affinely
// mapped tensor grid.  Reference: cpp/src/hermite/discretize.cpp:33-126 -- it writes /tmp/<hash>.cpp,
// runs `c++ -O3 -Ofast -shared` through system() and dlopen()s the result, then loops over the grid on
// one core.  Here the expression is compiled once per distinct body with NVRTC for sm_100a, loaded with
// cudaLibraryLoadData, and evaluated by one thread per grid point straight into device memory in the
// padded grid layout the transform consumes, so that a Quad.transform('expr') never touches the host.
// NVRTC is dlopen()ed on first use: the library itself does not link against it.
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "hb_kernels.h"
#include "hb_plan.h"

namespace hb {
namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return;
#define HB_SYM(field, sym) n.field = reinterpret_cast<decltype(n.field)>(dlsym(h, sym)); if (!n.field) return;
        HB_SYM(CreateProgram, "nvrtcCreateProgram")
        HB_SYM(CompileProgram, "nvrtcCompileProgram")
        HB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        HB_SYM(GetCUBIN, "nvrtcGetCUBIN")
        HB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        HB_SYM(GetProgramLog, "nvrtcGetProgramLog")
        HB_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef HB_SYM
        n.ok = true;
    });
    return n;
}

struct Compiled { cudaLibrary_t lib = nullptr; cudaKernel_t kernel = nullptr; };
std::mutex g_mutex;
std::map<std::string, Compiled> g_cache;  // key: "<dim>|<body>"

const char* kPrologue =
    "#define M_PI 3.14159265358979323846\n"
    "#define M_E 2.71828182845904523536\n"
    "#define M_LOG2E 1.44269504088896340736\n"
    "#define M_LOG10E 0.434294481903251827651\n"
    "#define M_LN2 0.693147180559945309417\n"
    "#define M_LN10 2.30258509299404568402\n"
    "#define M_PI_2 1.57079632679489661923\n"
    "#define M_PI_4 0.785398163397448309616\n"
    "#define M_1_PI 0.318309886183790671538\n"
    "#define M_2_PI 0.636619772367581343076\n"
    "#define M_2_SQRTPI 1.12837916709551257390\n"
    "#define M_SQRT2 1.41421356237309504880\n"
    "#define M_SQRT1_2 0.707106781186547524401\n"
    // sympy.ccode writes integer literals as function arguments (sqrt(3), pow(2, v[0])); with <cmath> the host
    // compiler of the reference promotes them to double, NVRTC's built-in overload set is ambiguous instead
    "#define HB_PROMOTE1(f) __device__ inline double f(int x) { return f((double)x); } "
    "__device__ inline double f(long long x) { return f((double)x); }\n"
    "HB_PROMOTE1(sqrt) HB_PROMOTE1(cbrt) HB_PROMOTE1(exp) HB_PROMOTE1(exp2) HB_PROMOTE1(expm1) HB_PROMOTE1(log) "
    "HB_PROMOTE1(log2) HB_PROMOTE1(log10) HB_PROMOTE1(log1p) HB_PROMOTE1(sin) HB_PROMOTE1(cos) HB_PROMOTE1(tan) "
    "HB_PROMOTE1(asin) HB_PROMOTE1(acos) HB_PROMOTE1(atan) HB_PROMOTE1(sinh) HB_PROMOTE1(cosh) HB_PROMOTE1(tanh) "
    "HB_PROMOTE1(asinh) HB_PROMOTE1(acosh) HB_PROMOTE1(atanh) HB_PROMOTE1(erf) HB_PROMOTE1(erfc) "
    "HB_PROMOTE1(tgamma) HB_PROMOTE1(lgamma) HB_PROMOTE1(fabs) HB_PROMOTE1(floor) HB_PROMOTE1(ceil)\n"
    "__device__ inline double pow(int a, int b) { return pow((double)a, b); }\n"
    "__device__ inline double pow(int a, double b) { return pow((double)a, b); }\n"
    "__device__ inline double atan2(int a, double b) { return atan2((double)a, b); }\n"
    "__device__ inline double atan2(double a, int b) { return atan2(a, (double)b); }\n"
    "__device__ inline double fmax(int a, double b) { return fmax((double)a, b); }\n"
    "__device__ inline double fmax(double a, int b) { return fmax(a, (double)b); }\n"
    "__device__ inline double fmin(int a, double b) { return fmin((double)a, b); }\n"
    "__device__ inline double fmin(double a, int b) { return fmin(a, (double)b); }\n"
    "struct HbGrid { int dim; int n[8]; long long pitch_last; long long total; double t[8]; double d[64]; };\n"
    "extern \"C\" __global__ void hb_discretize_kernel(const double* __restrict__ nodes, HbGrid g, double* out) {\n"
    "    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
    "    if (i >= g.total) return;\n"
    "    double node[8], v[8];\n"
    "    long long r = i, addr = 0, stride = 1;\n"
    "    int off[8]; off[0] = 0;\n"
    "    for (int a = 1; a < g.dim; a++) off[a] = off[a - 1] + g.n[a - 1];\n"
    "    for (int a = g.dim - 1; a >= 0; a--) {\n"            // row-major grid, last axis fastest
    "        const int k = (int)(r % g.n[a]); r /= g.n[a];\n"
    "        node[a] = nodes[off[a] + k];\n"
    "        addr += k * stride; stride *= (a == g.dim - 1) ? g.pitch_last : g.n[a];\n"
    "    }\n"
    "    for (int a = 0; a < g.dim; a++) {\n"                 // map_point, discretize.cpp:33-48
    "        double m = g.t[a];\n"
    "        for (int b = 0; b < g.dim; b++) m += g.d[a * g.dim + b] * node[b];\n"
    "        v[a] = m;\n"
    "    }\n"
    "    out[addr] = (double)(";
const char* kEpilogue = ");\n}\n";

struct HbGrid { int dim; int n[8]; long long pitch_last; long long total; double t[8]; double d[64]; };

int get_kernel(const std::string& body, cudaKernel_t* kernel) {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_cache.find(body);
    if (it != g_cache.end()) { *kernel = it->second.kernel; return HB_OK; }
    Nvrtc& n = nvrtc();
    if (!n.ok) { set_error("discretize: libnvrtc could not be loaded"); return HB_ERR_CUDA; }
    const std::string src = std::string(kPrologue) + body + kEpilogue;
    nvrtcProgram prog;
    if (n.CreateProgram(&prog, src.c_str(), "hb_discretize.cu", 0, nullptr, nullptr) != 0) {
        set_error("discretize: nvrtcCreateProgram failed");
        return HB_ERR_CUDA;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    const int rc = n.CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.GetProgramLog(prog, &log[0]);
        n.DestroyProgram(&prog);
        set_error("discretize: cannot compile expression: %.380s", log.c_str());
        return HB_ERR_INVALID;   // the reference prints "Error compiling file in /tmp" and then crashes in dlsym
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    Compiled c;
    HB_CUDA(cudaLibraryLoadData(&c.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    HB_CUDA(cudaLibraryGetKernel(&c.kernel, c.lib, "hb_discretize_kernel"));
    if (g_cache.size() > 256) g_cache.clear();   // libraries of evicted entries stay loaded (tiny)
    g_cache[body] = c;
    *kernel = c.kernel;
    return HB_OK;
}

}  // namespace

// d_nodes: concatenated per-axis nodes on the device; out: device, padded layout (pitch_last >= n[dim-1])
int discretize_device(const char* body, int dim, const int* n_pts, const double* d_nodes, const double* translation,
                      const double* dilation, double* d_out, long long pitch_last, cudaStream_t st) {
    if (!body || dim < 1 || dim > 8) return HB_ERR_INVALID;
    cudaKernel_t kernel;
    HB_TRY(get_kernel(body, &kernel));
    HbGrid g;
    g.dim = dim;
    g.total = 1;
    for (int a = 0; a < dim; a++) {
        g.n[a] = n_pts[a];
        g.total *= n_pts[a];
        g.t[a] = translation ? translation[a] : 0.0;
        for (int b = 0; b < dim; b++) g.d[a * dim + b] = dilation ? dilation[a * dim + b] : (a == b ? 1.0 : 0.0);
    }
    g.pitch_last = pitch_last;
    if (g.total == 0) return HB_OK;
    void* args[] = {(void*)&d_nodes, (void*)&g, (void*)&d_out};
    const unsigned blocks = (unsigned)((g.total + 255) / 256);
    HB_CUDA(cudaLaunchKernel((const void*)kernel, dim3(blocks), dim3(256), args, 0, st));
    note_launch(1);
    return HB_OK;
}

}  // namespace hb
