// engine.cu -- plan object, per-equation fused schedule and the C ABI (include/cupss_b200.h).
//
// What evolver::advanceTime (/root/reference/src/evolver.cpp:199-226) does with 4 sweeps of
// field::updateTerms / field::setRHS, each a chain of un-fused kernels, full C2C cuFFTs and
// cudaDeviceSynchronize(), is compiled here ONCE (finalize) into a short list of launches:
//
//   per sweep class (constraint fields first, then dynamic fields -- the reference's Jacobi order):
//     x pass      C2R of the dealiased fields, real-space products, R2C          (kernels_x.cu)
//     y pass      forward, 3-D only                                              (kernels_axis.cu, plain)
//     [all-to-all over NVLink when the grid is slab-partitioned]
//     k stage     last forward pass + update of every field of the sweep + dealias + first inverse pass
//     y pass      inverse, 3-D only
//
// Storage (per rank): Hermitian half spectra, complex64, [sz][sy_local][pitch] with pitch =
// roundup(sx/2+1, 16) so every row is 128-byte aligned; the dealiased real fields exist only as
// "y-inverse done" half spectra W2 = [sz_local][sy][pitch] which the x pass turns into real lines
// on chip.  W2 starts at zero, which reproduces the reference's step-0 behaviour (real_dealiased is
// zero until a field's first setRHS).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <vector>

#include <cuda.h>   // CUtensorMap types only; the encoder is resolved at run time (cudaGetDriverEntryPoint), libcuda is not linked

#include "../../include/cupss_b200.h"
#include "kernels.h"

using namespace cupss;

static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(CUPSS_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKR(call)                          \
    do {                                   \
        int r_ = (call);                   \
        if (r_ != CUPSS_B200_OK) return r_; \
    } while (0)

// ---------------------------------------------------------------- NCCL, bound lazily (the library loads without it)
struct Id128 { char b[128]; };
namespace {
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.h) return CUPSS_B200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail(CUPSS_B200_ERR_COMM, "libnccl.so.2 not found: %s", dlerror());
#define SYM(field, name)                                                         \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                            \
    if (!g_nccl.field) return fail(CUPSS_B200_ERR_COMM, "NCCL symbol %s missing", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return CUPSS_B200_OK;
}
#define NK(call)                                                                                   \
    do {                                                                                           \
        int r_ = (call);                                                                           \
        if (r_ != 0) return fail(CUPSS_B200_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------- run-time compilation of plan-specialised k stages
// A generic sweep (anything outside the q^2-polynomial class) is evaluated by an interpreter over the plan descriptors
// (kstage_point).  For grids where it matters the STRUCTURE of the sweep -- counts, source / destination indices, exponents,
// flags -- is turned into a constexpr plan type and the k-stage kernel is compiled for it with NVRTC from the very same
// headers (kernels_axis.cuh): every descriptor read, loop and branch of the interpreter folds away.  Coefficients, cut-offs
// and noise amplitudes stay run-time arguments, so updateParameter re-uses the compiled kernel.  NVRTC and the driver API
// are bound lazily; if either is missing the interpreter runs instead (same results).
namespace {
struct JitApi {
    bool tried = false, ok = false, nvrtcOk = false;   // ok: compile + load + launch; nvrtcOk: compile only
    void *hn = nullptr, *hc = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, cudaStream_t, void**, void**) = nullptr;
    std::map<std::string, void*> cache;   // source -> CUfunction
    std::string csrcDir;
};
}  // namespace
static JitApi g_jit;

static bool jit_load() {
    if (g_jit.tried) return g_jit.ok;
    g_jit.tried = true;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so"}) { g_jit.hn = dlopen(n, RTLD_NOW); if (g_jit.hn) break; }
    for (const char* n : {"libcuda.so.1", "libcuda.so"}) { g_jit.hc = dlopen(n, RTLD_NOW); if (g_jit.hc) break; }
    if (!g_jit.hn) return false;
#define JSYM(h, field, name) *(void**)(&g_jit.field) = dlsym(g_jit.h, name); if (!g_jit.field) return false;
    JSYM(hn, CreateProgram, "nvrtcCreateProgram") JSYM(hn, CompileProgram, "nvrtcCompileProgram")
    JSYM(hn, GetProgramLogSize, "nvrtcGetProgramLogSize") JSYM(hn, GetProgramLog, "nvrtcGetProgramLog")
    JSYM(hn, GetCUBINSize, "nvrtcGetCUBINSize") JSYM(hn, GetCUBIN, "nvrtcGetCUBIN") JSYM(hn, DestroyProgram, "nvrtcDestroyProgram")
    g_jit.nvrtcOk = true;
    Dl_info info;
    if (!dladdr((void*)&jit_load, &info) || !info.dli_fname) return false;
    std::string lib = info.dli_fname;                       // .../cupss_b200/lib/libcupss_b200.so
    const size_t slash = lib.find_last_of('/');
    g_jit.csrcDir = (slash == std::string::npos ? std::string(".") : lib.substr(0, slash)) + "/../csrc";
    FILE* f = fopen((g_jit.csrcDir + "/kernels_axis.cuh").c_str(), "r");
    if (!f) { g_jit.nvrtcOk = false; return false; }
    fclose(f);
    if (!g_jit.hc) return false;   // compile-only (self test) still possible
    JSYM(hc, ModuleLoadData, "cuModuleLoadData") JSYM(hc, ModuleGetFunction, "cuModuleGetFunction")
    JSYM(hc, FuncSetAttribute, "cuFuncSetAttribute") JSYM(hc, LaunchKernel, "cuLaunchKernel")
#undef JSYM
    g_jit.ok = true;
    return true;
}

// constexpr plan type + entry point for one sweep
static std::string jit_source(const cupss::KStageD& ks, int L) {
    using namespace cupss;
    std::string s = "#include \"kernels_axis.cuh\"\nnamespace cupss {\nstruct JitPlan {\n";
    char b[512];
    int npres = 0, nterm = 0;
    for (int o = 0; o < ks.nout; ++o) {
        nterm = std::max(nterm, ks.out[o].termOff + ks.out[o].nterm);
        npres = std::max(npres, ks.out[o].impOff + ks.out[o].nimp);
    }
    for (int t = 0; t < nterm; ++t) npres = std::max(npres, ks.term[t].presOff + ks.term[t].npres);
    snprintf(b, sizeof b, "    static constexpr int nsrc = %d, nout = %d;\n", ks.nsrc, ks.nout);
    s += b;
    s += "    CUPSS_HD static constexpr PlanOut out(int o) {\n        switch (o) {\n";
    for (int o = 0; o < ks.nout; ++o) {
        const OutD& d = ks.out[o];
        snprintf(b, sizeof b, "            case %d: return PlanOut{%d, %d, %d, %d, %d, %d, %d, %d, %d};\n", o, d.termOff, d.impOff, d.nterm, d.nimp,
                 d.dynamic, d.noisy, d.selfSrc, d.dst, d.inv);
        s += b;
    }
    s += "        }\n        return PlanOut{};\n    }\n    CUPSS_HD static constexpr PlanTerm term(int t) {\n        switch (t) {\n";
    for (int t = 0; t < nterm; ++t) {
        const TermD& d = ks.term[t];
        snprintf(b, sizeof b, "            case %d: return PlanTerm{%d, %d, %d, %d};\n", t, d.presOff, d.npres, d.src, d.mulI);
        s += b;
    }
    s += "        }\n        return PlanTerm{};\n    }\n    CUPSS_HD static constexpr PlanPres pres(int i) {\n        switch (i) {\n";
    for (int i = 0; i < npres; ++i) {
        const PresD& d = ks.pres[i];
        snprintf(b, sizeof b, "            case %d: return PlanPres{%d, %d, %d, %d, %d};\n", i, d.q2n, d.iqx, d.iqy, d.iqz, d.invq);
        s += b;
    }
    s += "        }\n        return PlanPres{};\n    }\n};\n}  // namespace cupss\n";
    // cluster-shared axes: at most 2 CTAs per SM like the library's k-stage kernels (128 registers per thread; measured on Model H
    // 2048^2: 0.097 ms against 0.50 ms with the 85-register cap of 3 CTAs per SM).  Single-CTA axes keep 3 CTAs per SM: the noisy
    // KPZ-3D 512^3 k stage runs 0.81 ms that way against 0.98 ms at 2 (profiles/README.md, r2e / r2f).
    if (cupss::axis_cluster_size(L) > 1) snprintf(b, sizeof b, "extern \"C\" __global__ void __cluster_dims__(cupss::AxisCfg<%d>::CL, 1, 1) __launch_bounds__(cupss::AxisCfg<%d>::THREADS, (cupss::AxisCfg<%d>::MINB > 2 ? 2 : cupss::AxisCfg<%d>::MINB))\n", L, L, L, L);
    else snprintf(b, sizeof b, "extern \"C\" __global__ void __launch_bounds__(cupss::AxisCfg<%d>::THREADS, cupss::AxisCfg<%d>::MINB)\n", L, L);
    s += b;
    s += "jit_kstage(const __grid_constant__ cupss::AxisArgs a, const __grid_constant__ cupss::KStageD ks) {\n";
    snprintf(b, sizeof b, "    cupss::axis_kstage_body<%d, cupss::KS_JIT, -1, cupss::JitPlan>(a, ks);\n}\n", L);
    s += b;
    return s;
}

// the lean evaluator with a signature that is not compiled into the library (a noisy field): same kernel body, SIG a constant
static std::string jit_source_lean(int L, int sig) {
    char b[1024];
    std::string s = "#include \"kernels_axis.cuh\"\n";
    if (cupss::axis_cluster_size(L) > 1) snprintf(b, sizeof b, "extern \"C\" __global__ void __cluster_dims__(cupss::AxisCfg<%d>::CL, 1, 1) __launch_bounds__(cupss::AxisCfg<%d>::THREADS, (cupss::AxisCfg<%d>::MINB > 2 ? 2 : cupss::AxisCfg<%d>::MINB))\n", L, L, L, L);
    else if (const char* mb = getenv("CUPSS_B200_LEAN_NOISE_MINB")) snprintf(b, sizeof b, "extern \"C\" __global__ void __launch_bounds__(cupss::AxisCfg<%d>::THREADS, %d)\n", L, atoi(mb));
    else snprintf(b, sizeof b, "extern \"C\" __global__ void __launch_bounds__(cupss::AxisCfg<%d>::THREADS, cupss::AxisCfg<%d>::MINB)\n", L, L);
    s += b;
    s += "jit_kstage(const __grid_constant__ cupss::AxisArgs a, const __grid_constant__ cupss::KStageD ks) {\n";
    snprintf(b, sizeof b, "    cupss::axis_kstage_body<%d, cupss::KS_SCALAR_Q2, %d, void>(a, ks);\n}\n", L, sig);
    s += b;
    return s;
}

// NVRTC only (no driver API needed): source -> cubin.  Used by the engine and by the GPU-less self test.
static bool jit_compile(const std::string& src, std::vector<char>* cubin, std::string* why);

// Returns the CUfunction of the specialised kernel, or nullptr (with the reason in `why`) if it cannot be built.
static void* jit_kstage_function(const cupss::KStageD& ks, int L, std::string* why, int leanSig = -1) {
    if (!jit_load()) { *why = "NVRTC / driver API / kernel sources not available"; return nullptr; }
    const std::string src = leanSig >= 0 ? jit_source_lean(L, leanSig) : jit_source(ks, L);
    auto it = g_jit.cache.find(src);
    if (it != g_jit.cache.end()) return it->second;
    std::vector<char> cubin;
    if (!jit_compile(src, &cubin, why)) return nullptr;
    void *mod = nullptr, *fn = nullptr;
    if (g_jit.ModuleLoadData(&mod, cubin.data()) != 0 || g_jit.ModuleGetFunction(&fn, mod, "jit_kstage") != 0) { *why = "cuModuleLoadData failed"; return nullptr; }
    int threads = 0, minb = 0;
    size_t smem = 0;
    axis_kstage_geometry(L, &threads, &smem, &minb);
    if (smem > 48 * 1024 && g_jit.FuncSetAttribute(fn, /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/ 8, (int)smem) != 0) { *why = "cuFuncSetAttribute failed"; return nullptr; }
    g_jit.cache[src] = fn;
    return fn;
}

static bool jit_compile(const std::string& src, std::vector<char>* cubin, std::string* why) {
    void* prog = nullptr;
    if (g_jit.CreateProgram(&prog, src.c_str(), "cupss_b200_jit_kstage.cu", 0, nullptr, nullptr) != 0) { *why = "nvrtcCreateProgram failed"; return false; }
    const std::string inc1 = "-I" + g_jit.csrcDir;
    const char* cudaHome = getenv("CUDA_HOME");
    const std::string inc2 = std::string("-I") + (cudaHome ? cudaHome : "/usr/local/cuda") + "/include";
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", inc1.c_str(), inc2.c_str()};
    const int rc = g_jit.CompileProgram(prog, 5, opts);
    if (rc != 0) {
        size_t n = 0;
        g_jit.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_jit.GetProgramLog(prog, &log[0]);
        *why = "nvrtcCompileProgram failed: " + log.substr(0, 1500);
        g_jit.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    g_jit.GetCUBINSize(prog, &n);
    cubin->resize(n);
    g_jit.GetCUBIN(prog, cubin->data());
    g_jit.DestroyProgram(&prog);
    return n > 0;
}

// ---------------------------------------------------------------- TMA descriptors of the strided-axis tile prologues
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
    static bool tried = false;
    static TensorMapEncodeTiledFn fn = nullptr;
    if (!tried) {
        tried = true;
        const char* off = getenv("CUPSS_B200_TMA");
        if (off && off[0] == '0') return nullptr;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// Tensor map of the INPUT array of a strided-axis launch (kernels.h: AxisArgs::tmap).  Leaves tmaOn = 0 (per-thread cp.async
// prologue) when the transform has a single level, the encoder is unavailable or the live row range does not split into boxes.
static void prepare_tma(cupss::AxisArgs& a, int L, bool kstage) {
    a.tmaOn = 0;
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc || L < 32 || !a.in || cupss::axis_cluster_size(L) > 1) return;   // cluster kernels load rows inside their cross level
    const int C = cupss::axis_tile_cols(L);
    const long long rpc = (long long)a.ain.rpcMask + 1, nchunk = L / rpc;
    const bool cyclic = a.ain.locShift > 0;   // row = local * nchunk + chunk: the chunk index runs fastest along the rows
    int B = L < 256 ? L : 256;
    if (!kstage && a.rowCut >= 0) {   // pruned inverse: boxes over [0, cut) and [L - cut, L)
        const int c = a.rowCut;
        if (c > 0) { B = 256; while (B > 1 && (c % B)) B >>= 1; }
        if (c > 0 && B < 16) return;
    }
    struct Dim { unsigned long long size, stride; unsigned box; int role; };
    const long long rs = a.ain.rs, cs = a.ain.cs, bs = a.ain.bs;
    Dim d[3] = {{(unsigned long long)rpc, (unsigned long long)rs * 8ull, (unsigned)(B < rpc ? B : rpc), 0},
                {(unsigned long long)nchunk, (unsigned long long)cs * 8ull, (unsigned)(B < rpc ? 1 : B / rpc), 1},
                {(unsigned long long)a.nbatch, (unsigned long long)bs * 8ull, 1u, 2}};
    if (cyclic) {   // a box of B consecutive rows = min(B, P) chunks x B / that locals, chunk index fastest in shared memory
        d[1].box = (unsigned)(B < nchunk ? B : nchunk);
        d[0].box = (unsigned)(B < nchunk ? 1 : B / nchunk);
    }
    unsigned long long top = 0;
    for (const Dim& x : d) if (x.size > 1) top = std::max(top, x.size * x.stride);
    if (top == 0) top = 128;
    for (Dim& x : d) {
        if (x.size > 1 && (x.stride == 0 || (x.stride & 15ull))) return;
        if (x.size <= 1) { x.size = 1; x.stride = top; top *= 2; }   // degenerate axes: any legal stride, kept monotonic
    }
    if (cyclic) std::swap(d[0], d[1]);   // dimension order = order in the box: chunk, local, batch (strides not monotonic; if the
                                         // encoder refuses that, the cp.async prologue runs)
    else std::stable_sort(d, d + 3, [](const Dim& p, const Dim& q) { return p.stride < q.stride; });
    cuuint64_t gdim[4] = {(cuuint64_t)(2 * a.ncol), d[0].size, d[1].size, d[2].size};
    cuuint64_t gstr[3] = {d[0].stride, d[1].stride, d[2].stride};
    cuuint32_t box[4] = {(cuuint32_t)(2 * C), d[0].box, d[1].box, d[2].box};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float2*>(a.in), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return;
    static_assert(sizeof(CUtensorMap) == sizeof(a.tmap), "tensor map size");
    memcpy(a.tmap, &tm, sizeof tm);
    for (int k = 0; k < 3; ++k) a.tmaSlot[k] = d[k].role;
    a.tmaBoxRows = B;
    a.tmaOn = 1;
}

// ---------------------------------------------------------------- plan data model
namespace {

struct Pres { float pre; int q2n, iqx, iqy, iqz, invq; };
struct Term { std::vector<Pres> pres; std::vector<int> product; };

struct Field {
    std::string name;
    bool dynamic = false;
    std::vector<Pres> implicit;
    std::vector<Term> terms;
    bool noisy = false;
    Pres noise{};
    unsigned long long seed = 0;
    bool needsAlias = false;
    int aliasOrder = 1;
    float2* S = nullptr;    // spectrum
    float2* W2 = nullptr;   // dealiased field, y-inverse done
    int w2cut[3] = {-1, -1, -1};   // cut-offs W2 was pruned with (pruned regions rely on staying zero)
    bool w2FullBand = false;       // a real-space callback wrote W2: the x pass must not stop at the dealias cut-off
};

struct Launch {
    enum Kind { XPASS, AXIS_PLAIN, AXIS_KSTAGE, A2A, BUMP, XBAR } kind;
    char name[64];
    int L = 0, dir = 0, mode = 0;
    AxisArgs ax{};
    KStageD ks{};
    XArgs xa{};
    const float2* send = nullptr;
    float2* recv = nullptr;
    size_t chunk = 0;   // float2 per peer
    double bytes = 0;   // algorithmic HBM bytes (A2A: bytes sent)
    double commBytes = 0;   // bytes this rank sends to other GPUs inside this launch (pushed exchange)
    XBarrier xb{};
    void* jitFn = nullptr;   // plan-specialised k stage compiled at run time (AXIS_KSTAGE), else the library's kernels
    // Lanes: a launch goes to stream `lane` (0 = the plan's main stream).  `waitEv` / `recEv` index the plan's event pool:
    // the launch waits for waitEv before it starts and records recEv when it is queued (cross-lane dependencies of the
    // chunked pipelines).  `joinBefore`: every lane used so far is joined into the main stream first.  The list order is
    // always a valid serial order, so running a list on one stream (profile_step) is correct too.
    int lane = 0, waitEv = -1, recEv = -1;
    bool joinBefore = false;
    int pipe = -1;           // launches of one multi-lane pipeline share an id >= 0; the launch after a pipeline joins the lanes
    bool tmaReady = false;   // AXIS_*: the tensor map of the input was built (lazily, at the first launch)
    int group = -1;          // launches of one pipeline (same id >= 0) are timed as one unit by profile_step
    char groupName[64] = "";
    double groupBytes = 0;   // DRAM-level bytes of the whole pipeline (set on its first launch)
};

int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
int FftLevelsN(int L) { return L <= 16 ? 1 : 2; }   // only "more than one level" matters here (the tile exists)

}  // namespace

struct cupss_b200_plan {
    int sx, sy, sz;
    float dx, dy, dz, dt;
    int dim;
    int rank = 0, nranks = 1;
    int zl, kyl;           // local z planes (real space) / local ky rows (Fourier space)
    // Fourier space: rank r owns the rows ky = r, r + P, r + 2P, ... (cyclic) so that the |ky| <= cut rows a dealiased inverse
    // transform keeps -- and sends back -- are spread evenly over the ranks; with contiguous blocks (CUPSS_B200_KY_BLOCK=1) half of
    // the ranks own none of them at P = 4 and 8 and the others do, and push, twice the share.
    bool kyCyclic = false;
    int ncol, pitch;
    size_t specElems;      // float2 per spectrum-shaped array on this rank
    int dealiasRule = CUPSS_B200_DEALIAS_GPU_RULE;
    std::vector<Field> fields;
    std::vector<Launch> step;
    size_t stageSplit = 0;     // step[0, stageSplit): constraint sweep; [stageSplit, end): dynamic sweep + counter bump
    float2* viewBuf = nullptr; // float2[N] view of a real field handed to user callbacks
    std::vector<float2*> scratch;
    std::map<int, float2*> twiddles;
    float* realBuf = nullptr;
    unsigned int* stepCounter = nullptr;
    cudaStream_t stream = nullptr;
    cudaGraphExec_t graphExec = nullptr;
    bool finalized = false;
    bool useGraph = true;
    bool prune = true;     // skip the parts of inverse transforms that the dealias mask makes identically zero
    void* comm = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // extra streams / events of the chunked pipelines (lane 0 is `stream`); lane 1 has the higher priority: it carries the
    // downstream kernel of a producer -> consumer pair, whose CTAs should get the SM slots the producer's CTAs free
    static constexpr int kMaxLanes = 3;
    cudaStream_t laneStream[kMaxLanes] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> evPool;
    cudaEvent_t forkEv[kMaxLanes] = {nullptr, nullptr, nullptr}, joinEv[kMaxLanes] = {nullptr, nullptr, nullptr};
    int nextEv = 0, nextGroup = 0, nextPipe = 0;
    int zChunk = 0;        // planes per chunk of the x -> forward-y pipeline (0: off)
    int zChunkLanes = 2;
    int xChunks = 1;       // column chunks of the slab-exchange pipeline (1: off, launches run one after the other)
    bool laneFanout = true;   // 1-D / 2-D grids: independent launches of a sweep side by side on the lanes (emit_concurrent)
    // Peer-memory exchange arena (multi-GPU): [header: flags, epochs, error][slot 0][slot 1]...; every rank maps
    // every peer's arena through CUDA IPC, and the y / z pass kernels store their output rows straight into the
    // owner's slot over NVLink (no NCCL, no staging copy on the hot path).
    bool useP2P = true;
    float2* arena = nullptr;
    size_t arenaSlots = 0;
    float2* peerArena[CUPSS_MAX_PEERS] = {nullptr};
    static constexpr size_t kArenaHeader = 4096;   // float2 elements (32 KiB)
    int arenaNext = 0, nextPt = 0;
    int* hostErr = nullptr;       // mapped page-locked word raised by a barrier that timed out (kernels_axis.cu)
    int* hostErrDev = nullptr;
    unsigned long long barrierTimeoutNs = 60ull * 1000000000ull;   // CUPSS_B200_BARRIER_TIMEOUT_S (0: wait for ever)

    // ------------------------------------------------------------ helpers
    int get_twiddle(int L, const float2** out) {
        auto it = twiddles.find(L);
        if (it != twiddles.end()) { *out = it->second; return CUPSS_B200_OK; }
        std::vector<float2> h(L > 0 ? L : 1, make_float2(1.0f, 0.0f));
        if (host_level_twiddles(L, h.data()) <= 0) return fail(CUPSS_B200_ERR_ARG, "unsupported transform length %d", L);
        float2* d = nullptr;
        CK(cudaMalloc(&d, sizeof(float2) * L));
        CK(cudaMemcpyAsync(d, h.data(), sizeof(float2) * L, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        twiddles[L] = d;
        *out = d;
        return CUPSS_B200_OK;
    }
    // Twiddles of a strided-axis pass over L rows: the level table of the transform one CTA runs and, when a cluster shares
    // the axis, the table of the level that couples the CTAs' blocks (key 1000000 + L in the same cache).
    int get_axis_twiddles(int L, AxisArgs& a) {
        a.twX = nullptr;
        const int CL = axis_cluster_size(L);
        if (CL <= 1) return get_twiddle(L, &a.tw);
        CKR(get_twiddle(axis_cta_rows(L), &a.tw));
        auto it = twiddles.find(1000000 + L);
        if (it == twiddles.end()) {
            std::vector<float2> h((size_t)L);
            const int nent = host_cross_twiddles(L, h.data());
            if (nent <= 0) return fail(CUPSS_B200_ERR_ARG, "unsupported transform length %d", L);
            float2* d = nullptr;
            CK(cudaMalloc(&d, sizeof(float2) * (size_t)nent));
            CK(cudaMemcpyAsync(d, h.data(), sizeof(float2) * (size_t)nent, cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
            it = twiddles.emplace(1000000 + L, d).first;
        }
        a.twX = it->second;
        return CUPSS_B200_OK;
    }
    // twiddles of the two-level x kernel (key -sx in the same cache); *out = nullptr when that kernel does not cover sx
    int get_twiddle_x3(const float2** out) {
        *out = nullptr;
        const char* no = getenv("CUPSS_B200_NO_X3");
        if ((no && no[0] == '1') || !xpass3_supported(sx)) return CUPSS_B200_OK;
        auto it = twiddles.find(-sx);
        if (it != twiddles.end()) { *out = it->second; return CUPSS_B200_OK; }
        std::vector<float2> h(sx, make_float2(1.0f, 0.0f));
        if (host_x3_twiddles(sx, h.data()) <= 0) return CUPSS_B200_OK;
        float2* d = nullptr;
        CK(cudaMalloc(&d, sizeof(float2) * sx));
        CK(cudaMemcpyAsync(d, h.data(), sizeof(float2) * sx, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        twiddles[-sx] = d;
        *out = d;
        return CUPSS_B200_OK;
    }
    int get_scratch(size_t idx, float2** out) {
        while (scratch.size() <= idx) scratch.push_back(nullptr);
        if (!scratch[idx]) {
            CK(cudaMalloc(&scratch[idx], specElems * sizeof(float2)));
            CK(cudaMemsetAsync(scratch[idx], 0, specElems * sizeof(float2), stream));
        }
        *out = scratch[idx];
        return CUPSS_B200_OK;
    }
    double spec_bytes() const { return (double)ncol * (double)sy * (double)sz / nranks * 8.0; }

    // Addressing of the three kinds of strided pass.
    void fill_common(AxisArgs& a, int L) {
        a.ncol = ncol;
        a.ncolTiles = (ncol + axis_tile_cols(L) - 1) / axis_tile_cols(L);
        a.ctBase = 0;
        a.sx = sx; a.sy = sy; a.sz = sz;
        a.maskOn = 0; a.cutx = a.cuty = a.cutz = 0;
        a.pruneOn = 0; a.pruneCutX = a.pruneCutY = 0; a.rowCut = -1;
        a.kyBase = kyCyclic ? rank : rank * kyl;
        a.kyStride = kyCyclic ? nranks : 1;
    }
    // last axis of the transform: z in 3-D (rows kz, batch ky_local), y in 2-D, nothing in 1-D
    int make_last_axis(AxisArgs& a, int* L) {
        *L = dim == 3 ? sz : (dim == 2 ? sy : 1);
        fill_common(a, *L);
        AxisAddr n{};
        if (dim == 3) { n.bs = pitch; n.rs = (long long)kyl * pitch; a.nbatch = kyl; a.axis = 2; }
        else if (dim == 2) { n.bs = 0; n.rs = pitch; a.nbatch = 1; a.axis = 1; }
        else { n.bs = 0; n.rs = 0; a.nbatch = 1; a.axis = 0; }
        n.cs = 0; n.rpcShift = ilog2(*L); n.rpcMask = *L - 1; n.chunkMask = 0; n.locShift = 0;
        a.ain = n; a.aout = n;
        return get_axis_twiddles(*L, a);
    }
    // y pass of a 3-D transform: natural side [z_local][y][pitch], exchange side [peer][z_local][ky_local][pitch]
    int make_y_axis(AxisArgs& a, bool forward) {
        fill_common(a, sy);
        a.nbatch = zl; a.axis = 1;
        AxisAddr nat{}, exc{};
        nat.bs = (long long)sy * pitch; nat.rs = pitch; nat.cs = 0; nat.rpcShift = ilog2(sy); nat.rpcMask = sy - 1; nat.chunkMask = 0; nat.locShift = 0;
        exc.bs = (long long)kyl * pitch; exc.rs = pitch; exc.cs = (long long)zl * kyl * pitch;
        exc.rpcMask = kyl - 1; exc.chunkMask = nranks - 1;
        if (kyCyclic) { exc.rpcShift = 0; exc.locShift = ilog2(nranks); }   // ky = ky_local * P + peer
        else { exc.rpcShift = ilog2(kyl); exc.locShift = 0; }              // ky = peer * kyl + ky_local
        if (forward) { a.ain = nat; a.aout = exc; } else { a.ain = exc; a.aout = nat; }
        return get_axis_twiddles(sy, a);
    }

    int add_a2a(std::vector<Launch>& out, const char* nm, const float2* send, float2* recv) {
        Launch l{};
        l.kind = Launch::A2A;
        snprintf(l.name, sizeof l.name, "%s", nm);
        l.send = send; l.recv = recv;
        l.chunk = (size_t)zl * kyl * pitch;
        l.bytes = (double)l.chunk * 8.0 * (nranks - 1);
        out.push_back(l);
        return CUPSS_B200_OK;
    }

    float2* arena_slot_ptr(int peer, int slot) const {
        return peerArena[peer] ? peerArena[peer] + kArenaHeader + (size_t)slot * specElems : nullptr;
    }
    // Collective: (re)allocate the arena with room for `nslots` receive buffers and map every peer's copy.
    int ensure_arena(size_t nslots) {
        if (arena && arenaSlots >= nslots) return CUPSS_B200_OK;
        CK(cudaStreamSynchronize(stream));
        for (int d = 0; d < nranks; ++d) {
            if (d != rank && peerArena[d]) CK(cudaIpcCloseMemHandle(peerArena[d]));
            peerArena[d] = nullptr;
        }
        if (arena) CK(cudaFree(arena));
        arena = nullptr;
        const size_t elems = kArenaHeader + nslots * specElems;
        CK(cudaMalloc(&arena, elems * sizeof(float2)));
        CK(cudaMemsetAsync(arena, 0, elems * sizeof(float2), stream));
        cudaIpcMemHandle_t mine;
        CK(cudaIpcGetMemHandle(&mine, arena));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
        char* dbuf = nullptr;
        CK(cudaMalloc(&dbuf, 64 * (size_t)(nranks + 1)));
        CK(cudaMemcpyAsync(dbuf, &mine, 64, cudaMemcpyHostToDevice, stream));
        NK(g_nccl.AllGather(dbuf, dbuf + 64, 64, /*ncclChar*/ 0, comm, stream));
        std::vector<cudaIpcMemHandle_t> all(nranks);
        CK(cudaMemcpyAsync(all.data(), dbuf + 64, 64 * (size_t)nranks, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaFree(dbuf));
        for (int d = 0; d < nranks; ++d) {
            if (d == rank) { peerArena[d] = arena; continue; }
            void* ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, all[d], cudaIpcMemLazyEnablePeerAccess));
            peerArena[d] = static_cast<float2*>(ptr);
        }
        arenaSlots = nslots;
        return CUPSS_B200_OK;
    }
    // Receive-side addressing of a pushed exchange: slot layout [src rank][z_local][ky_local][pitch].
    void set_push(AxisArgs& a, int slot, bool rowsAreKy) {
        a.pushOn = 1;
        a.pushPeerMask = nranks - 1; a.pushLocShift = 0;
        if (rowsAreKy) {   // y pass: batch = z_local; the row ky goes to the rank that owns it
            a.pushMask = kyl - 1; a.pushRs = pitch; a.pushBs = (long long)kyl * pitch;
            if (kyCyclic) { a.pushShift = 0; a.pushLocShift = ilog2(nranks); } else a.pushShift = ilog2(kyl);
        } else { a.pushShift = ilog2(zl); a.pushMask = zl - 1; a.pushRs = (long long)kyl * pitch; a.pushBs = pitch; }   // z pass: batch = ky_local, z-slabs are contiguous
        a.pushBase = (long long)kArenaHeader + (long long)slot * (long long)specElems + (long long)rank * zl * kyl * pitch;
        for (int d = 0; d < CUPSS_MAX_PEERS; ++d) a.push[d] = d < nranks ? peerArena[d] : nullptr;
    }
    // Cross-GPU barrier after a pushed exchange: a 1-warp kernel (kernels_axis.cu).  Measured alternatives that did not pay
    // (profiles/README.md): flags folded into the producing / consuming kernels, z-chunked overlap with the x pass.
    void add_barrier(std::vector<Launch>& out, const char* nm) {
        Launch l{};
        l.kind = Launch::XBAR;
        snprintf(l.name, sizeof l.name, "%s", nm);
        for (int d = 0; d < CUPSS_MAX_PEERS; ++d) l.xb.flags[d] = d < nranks ? reinterpret_cast<unsigned int*>(peerArena[d]) : nullptr;
        l.xb.epoch = arena ? reinterpret_cast<unsigned int*>(arena) + 1024 : nullptr;
        l.xb.error = arena ? reinterpret_cast<int*>(arena) + 2048 : nullptr;
        l.xb.hostError = hostErrDev;
        l.xb.timeoutNs = barrierTimeoutNs;
        l.xb.rank = rank; l.xb.nranks = nranks; l.xb.pt = nextPt++;
        out.push_back(l);
    }

    int lane_stream(int lane, cudaStream_t* out) {
        if (lane <= 0) { *out = stream; return CUPSS_B200_OK; }
        if (lane >= kMaxLanes) return fail(CUPSS_B200_ERR_STATE, "internal: lane %d", lane);
        if (!laneStream[lane]) {
            int lo = 0, hi = 0;   // lo: least priority (numerically largest), hi: greatest
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&laneStream[lane], cudaStreamNonBlocking, lane == 1 ? hi : lo));
            CK(cudaEventCreateWithFlags(&forkEv[lane], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&joinEv[lane], cudaEventDisableTiming));
        }
        *out = laneStream[lane];
        return CUPSS_B200_OK;
    }
    int pool_event(int idx, cudaEvent_t* out) {
        while ((int)evPool.size() <= idx) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            evPool.push_back(e);
        }
        *out = evPool[idx];
        return CUPSS_B200_OK;
    }
    int join_lanes(bool (&forked)[kMaxLanes]) {
        for (int ln = 1; ln < kMaxLanes; ++ln)
            if (forked[ln]) {
                CK(cudaEventRecord(joinEv[ln], laneStream[ln]));
                CK(cudaStreamWaitEvent(stream, joinEv[ln], 0));
                forked[ln] = false;
            }
        return CUPSS_B200_OK;
    }

    int run_launch(Launch& l) { return run_launch(l, stream); }
    int run_launch(Launch& l, cudaStream_t stream) {
        switch (l.kind) {
            case Launch::XPASS: CK(launch_xpass(sx, l.mode, l.xa, stream)); break;
            case Launch::AXIS_PLAIN:
                if (!l.tmaReady) { prepare_tma(l.ax, l.L, false); l.tmaReady = true; }
                CK(launch_axis_plain(l.L, l.dir, l.ax, stream));
                break;
            case Launch::AXIS_KSTAGE:
                if (!l.tmaReady) { prepare_tma(l.ax, l.L, true); l.tmaReady = true; }
                if (l.jitFn) {
                    int threads = 0, minb = 0;
                    size_t smem = 0;
                    axis_kstage_geometry(l.L, &threads, &smem, &minb);
                    void* params[] = {&l.ax, &l.ks};
                    const unsigned grid = (unsigned)l.ax.ncolTiles * (unsigned)l.ax.nbatch * (unsigned)axis_cluster_size(l.L);
                    const int rc = g_jit.LaunchKernel(l.jitFn, grid, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem, stream, params, nullptr);
                    if (rc != 0) return fail(CUPSS_B200_ERR_CUDA, "cuLaunchKernel of the plan-specialised k stage failed (%d)", rc);
                } else {
                    CK(launch_axis_kstage(l.L, l.ax, l.ks, stream));
                }
                break;
            case Launch::BUMP: CK(launch_bump_counter(stepCounter, stream)); break;
            case Launch::XBAR: CK(launch_xgpu_barrier(l.xb, stream)); break;
            case Launch::A2A: {
                NK(g_nccl.GroupStart());
                for (int d = 0; d < nranks; ++d) {
                    NK(g_nccl.Send(l.send + (size_t)d * l.chunk, l.chunk * 2, /*ncclFloat*/ 7, d, comm, stream));
                    NK(g_nccl.Recv(l.recv + (size_t)d * l.chunk, l.chunk * 2, 7, d, comm, stream));
                }
                NK(g_nccl.GroupEnd());
                break;
            }
        }
        return CUPSS_B200_OK;
    }
    // Launches [b, e) of a list, each on its lane; every lane that was used is joined into the main stream at the end, so the
    // caller (and a stream capture) only ever sees the main stream.
    int run_range(std::vector<Launch>& v, size_t b, size_t e) {
        bool forked[kMaxLanes] = {false, false, false};
        for (size_t i = b; i < e; ++i) {
            Launch& l = v[i];
            if (l.joinBefore) CKR(join_lanes(forked));
            cudaStream_t st;
            CKR(lane_stream(l.lane, &st));
            if (l.lane > 0 && !forked[l.lane]) {   // the lane starts after everything queued on the main stream so far
                CK(cudaEventRecord(forkEv[l.lane], stream));
                CK(cudaStreamWaitEvent(st, forkEv[l.lane], 0));
                forked[l.lane] = true;
            }
            if (l.waitEv >= 0) {
                cudaEvent_t ev;
                CKR(pool_event(l.waitEv, &ev));
                CK(cudaStreamWaitEvent(st, ev, 0));
            }
            CKR(run_launch(l, st));
            if (l.recEv >= 0) {
                cudaEvent_t ev;
                CKR(pool_event(l.recEv, &ev));
                CK(cudaEventRecord(ev, st));
            }
        }
        return join_lanes(forked);
    }
    int run_list(std::vector<Launch>& v) { return run_range(v, 0, v.size()); }

    // ------------------------------------------------------------ full transforms (upload / download; not on the hot path)
    // real [zl][sy][sx] (realBuf) -> spectrum S
    int forward_full(float2* S) {
        std::vector<Launch> v;
        float2 *t0, *t1;
        CKR(get_scratch(0, &t0));
        CKR(get_scratch(1, &t1));
        Launch x{};
        x.kind = Launch::XPASS; x.mode = X_R2C_ONLY;
        x.xa.nIn = 0; x.xa.nOut = 1; x.xa.nMono = 0;
        x.xa.out[0] = dim == 1 ? S : t0;
        x.xa.pitch = pitch; x.xa.nlines = (long long)zl * sy; x.xa.norm = 1.0f; x.xa.realIn = realBuf;
        CKR(get_twiddle(sx, &x.xa.tw));
        v.push_back(x);
        if (dim == 3) {
            Launch y{};
            y.kind = Launch::AXIS_PLAIN; y.L = sy; y.dir = -1;
            CKR(make_y_axis(y.ax, true));
            y.ax.in = t0; y.ax.out = t1;
            v.push_back(y);
            const float2* zin = t1;
            if (nranks > 1) { CKR(add_a2a(v, "a2a", t1, t0)); zin = t0; }
            Launch z{};
            z.kind = Launch::AXIS_PLAIN; z.dir = -1;
            CKR(make_last_axis(z.ax, &z.L));
            z.ax.in = zin; z.ax.out = S;
            v.push_back(z);
        } else if (dim == 2) {
            Launch y{};
            y.kind = Launch::AXIS_PLAIN; y.dir = -1;
            CKR(make_last_axis(y.ax, &y.L));
            y.ax.in = t0; y.ax.out = S;
            v.push_back(y);
        }
        return run_list(v);
    }
    // spectrum S -> real (realBuf), normalised
    int inverse_full(const float2* S) {
        std::vector<Launch> v;
        float2 *t0, *t1;
        CKR(get_scratch(0, &t0));
        CKR(get_scratch(1, &t1));
        const float2* xin = S;
        if (dim == 3) {
            Launch z{};
            z.kind = Launch::AXIS_PLAIN; z.dir = +1;
            CKR(make_last_axis(z.ax, &z.L));
            z.ax.in = S; z.ax.out = t0;
            v.push_back(z);
            const float2* yin = t0;
            if (nranks > 1) { CKR(add_a2a(v, "a2a", t0, t1)); yin = t1; }
            Launch y{};
            y.kind = Launch::AXIS_PLAIN; y.L = sy; y.dir = +1;
            CKR(make_y_axis(y.ax, false));
            float2* yo = yin == t0 ? t1 : t0;
            y.ax.in = yin; y.ax.out = yo;
            v.push_back(y);
            xin = yo;
        } else if (dim == 2) {
            Launch y{};
            y.kind = Launch::AXIS_PLAIN; y.dir = +1;
            CKR(make_last_axis(y.ax, &y.L));
            y.ax.in = S; y.ax.out = t0;
            v.push_back(y);
            xin = t0;
        }
        Launch x{};
        x.kind = Launch::XPASS; x.mode = X_C2R_ONLY;
        x.xa.nIn = 1; x.xa.nOut = 0; x.xa.nMono = 0;
        x.xa.in[0] = xin;
        x.xa.kmax[0] = ncol - 1;
        x.xa.pitch = pitch; x.xa.nlines = (long long)zl * sy;
        x.xa.norm = 1.0f / ((float)sx * (float)sy * (float)sz);
        x.xa.realOut = realBuf;
        CKR(get_twiddle(sx, &x.xa.tw));
        v.push_back(x);
        return run_list(v);
    }

    // ------------------------------------------------------------ schedule construction
    struct Mono { float coef; std::vector<int> fac; };
    struct Group { int field; int firstTerm; std::vector<Pres> pres; std::vector<Mono> monos; };
    static std::vector<int> group_inputs(const Group& g) {   // distinct fields its monomials read
        std::vector<int> ins;
        for (const Mono& m : g.monos)
            for (int fid : m.fac)
                if (std::find(ins.begin(), ins.end(), fid) == ins.end()) ins.push_back(fid);
        return ins;
    }

    static bool proportional(const std::vector<Pres>& a, const std::vector<Pres>& b, float* lambda) {
        if (a.size() != b.size()) return false;
        float lam = 0.0f; bool have = false;
        for (size_t i = 0; i < a.size(); ++i) {
            if (a[i].q2n != b[i].q2n || a[i].iqx != b[i].iqx || a[i].iqy != b[i].iqy || a[i].iqz != b[i].iqz || a[i].invq != b[i].invq) return false;
            if (a[i].pre == 0.0f) { if (b[i].pre != 0.0f) return false; continue; }
            const float l = b[i].pre / a[i].pre;
            if (have && l != lam) return false;
            lam = l; have = true;
        }
        if (!have) return false;
        *lambda = lam;
        return true;
    }

    static void fold_pres(const std::vector<Pres>& in, KStageD& ks, int& npresUsed, TermD& td) {
        const int mulI = in.empty() ? 0 : ((in[0].iqx + in[0].iqy + in[0].iqz) % 2);
        td.presOff = (short)npresUsed; td.npres = (signed char)in.size(); td.mulI = (signed char)mulI;
        for (const Pres& p : in) {
            const int units = p.iqx + p.iqy + p.iqz;
            const int negate = -2 * (((units - mulI) / 2) % 2) + 1;   // term::precomputePrefactors, src/term_init.cpp:155
            PresD d{};
            d.pre = p.pre * (float)negate;
            d.q2n = (signed char)p.q2n; d.iqx = (signed char)p.iqx; d.iqy = (signed char)p.iqy;
            d.iqz = (signed char)p.iqz; d.invq = (signed char)p.invq;
            ks.pres[npresUsed++] = d;
        }
    }

    void cutoffs(int order, short* cx, short* cy, short* cz) const {
        if (dealiasRule == CUPSS_B200_DEALIAS_GPU_RULE) {
            *cx = (short)(sx / (order + 1)); *cy = (short)(sy / (order + 1)); *cz = (short)(sz / (order + 1));
        } else {   // field::dealias CPU loop: third test repeats nj against sz (src/field.cpp:220)
            const int a = sy / (order + 1), b = sz / (order + 1);
            *cx = (short)(sx / (order + 1)); *cy = (short)(a < b ? a : b); *cz = 32767;
        }
    }

    // Masked inverse last-axis pass of field f (dealias mask + innermost inverse axis, src/field.cpp:199-232 + the first part
    // of toReal); the caller sets out / push.
    int lastinv_launch(int f, const char* tag, Launch& z) {
        z = Launch{};
        z.kind = Launch::AXIS_PLAIN; z.dir = +1;
        snprintf(z.name, sizeof z.name, "lastinv_%s", tag);
        CKR(make_last_axis(z.ax, &z.L));
        z.ax.maskOn = 1;
        short cx, cy, cz;
        cutoffs(fields[f].aliasOrder, &cx, &cy, &cz);
        z.ax.cutx = cx; z.ax.cuty = cy; z.ax.cutz = cz;
        if (prune) {
            z.ax.pruneOn = 1; z.ax.pruneCutX = cx; z.ax.pruneCutY = cy;
            z.ax.rowCut = dim == 3 ? (cz < z.L ? cz : -1) : (cy < z.L ? cy : -1);
        }
        z.ax.in = fields[f].S;
        z.bytes = 2.0 * spec_bytes();
        if (prune) {   // live column tiles x live ky rows (whole tiles skipped); of a live tile only the rows up to the cut-off are read
            const int C = axis_tile_cols(z.L);
            const double fx = (double)std::min(ncol, (cx / C + 1) * C) / ncol;
            const double fy = dim == 3 ? std::min(1.0, (2.0 * cy + 1.0) / sy) : 1.0;
            const double fr = z.ax.rowCut >= 0 ? std::min(1.0, (2.0 * z.ax.rowCut + 1.0) / z.L) : 1.0;
            z.bytes = fx * fy * (fr + 1.0) * spec_bytes();
        }
        return CUPSS_B200_OK;
    }
    // Inverse y pass (3-D) of the dealiased copy of field f into its W2, pruned to the dealias cut-off.
    int yinv_launch(int f, const char* tag, const float2* yin, Launch& y) {
        y = Launch{};
        y.kind = Launch::AXIS_PLAIN; y.L = sy; y.dir = +1;
        snprintf(y.name, sizeof y.name, "yinv_%s", tag);
        CKR(make_y_axis(y.ax, false));
        y.ax.in = yin; y.ax.out = fields[f].W2;
        y.bytes = 2.0 * spec_bytes();
        if (prune) {
            short cx, cy, cz;
            cutoffs(fields[f].aliasOrder, &cx, &cy, &cz);
            const int C = axis_tile_cols(sy);
            y.ax.pruneOn = 1; y.ax.pruneCutX = cx; y.ax.pruneCutY = 32767;   // batch is z here: never pruned
            y.ax.rowCut = cy < sy ? cy : -1;
            const double fx = (double)std::min(ncol, (cx / C + 1) * C) / ncol;
            y.bytes = fx * spec_bytes() * (1.0 + std::min(1.0, (2.0 * cy + 1.0) / sy));
        }
        return CUPSS_B200_OK;
    }
    // W2 of field f recomputed from its spectrum (after a Fourier-space callback changed S): field::dealias + toReal's
    // strided-axis part, outside the fused k stage.  Single GPU.
    int refresh_w2(int f) {
        Field& F = fields[f];
        if (!F.W2) return CUPSS_B200_OK;
        std::vector<Launch> v;
        Launch z;
        CKR(lastinv_launch(f, "cb", z));
        float2* w1 = F.W2;
        if (dim == 3) CKR(get_scratch(0, &w1));
        z.ax.out = w1;
        v.push_back(z);
        if (dim == 3) {
            Launch y;
            CKR(yinv_launch(f, "cb", w1, y));
            v.push_back(y);
        }
        return run_list(v);
    }

    // z-chunked software pipeline of the x passes (main stream) and the forward y passes (lane 1, higher priority): the y pass
    // of chunk c runs next to the x pass of chunk c+1.  The x pass is bound by issue slots and the y pass by memory, so the two
    // fill each other's idle resource, and a chunk of the x-pass output (zChunk planes x 1.1 MB) is still in the 126 MB L2 when
    // the y pass reads it: that read (4 B per point) never reaches HBM.
    void emit_xy_pipeline(std::vector<Launch>& out, const std::vector<Launch>& xs, const std::vector<Launch>& ys, const char* tag) {
        const int nchunk = zl / zChunk, gid = nextGroup++, pid = nextPipe++;
        const double frac = 1.0 / nchunk;
        double gbytes = 0;
        for (const Launch& x : xs) gbytes += x.bytes;
        for (const Launch& y : ys) gbytes += y.bytes - spec_bytes();   // the y pass reads what the x pass just left in L2
        bool first = true;
        for (int c = 0; c < nchunk; ++c) {
            const long long z0 = (long long)c * zChunk;
            const int ev = nextEv++;
            for (size_t i = 0; i < xs.size(); ++i) {
                Launch x = xs[i];
                for (int q = 0; q < x.xa.nIn; ++q) x.xa.in[q] += z0 * sy * pitch;
                for (int q = 0; q < x.xa.nOut; ++q) x.xa.out[q] += z0 * sy * pitch;
                x.xa.nlines = (long long)zChunk * sy;
                x.bytes *= frac;
                x.lane = 0;
                if (i + 1 == xs.size()) x.recEv = ev;
                x.group = gid; x.pipe = pid;
                if (first) { snprintf(x.groupName, sizeof x.groupName, "xyfwd_%s", tag); x.groupBytes = gbytes; first = false; }
                out.push_back(x);
            }
            for (size_t i = 0; i < ys.size(); ++i) {
                Launch y = ys[i];
                y.ax.in += z0 * y.ax.ain.bs;
                y.ax.out += z0 * y.ax.aout.bs;
                y.ax.nbatch = zChunk;
                y.bytes *= frac;
                y.lane = zChunkLanes > 1 ? 1 : 0;   // one lane: plain alternation x(c), y(c), x(c+1), ... (L2 hand-off only)
                if (i == 0) y.waitEv = ev;
                y.group = gid; y.pipe = pid;
                out.push_back(y);
            }
        }
    }

    // Column chunk [t0, t0 + nt) of a strided-axis launch.
    Launch column_chunk(const Launch& l, int t0, int nt) const {
        Launch c = l;
        const int C = axis_tile_cols(l.L);
        const double frac = (double)(std::min(ncol, (t0 + nt) * C) - std::min(ncol, t0 * C)) / ncol;
        c.ax.ctBase = t0; c.ax.ncolTiles = nt;
        c.bytes *= frac; c.commBytes *= frac;
        return c;
    }
    static bool chunk_live(const Launch& l, int t0) {   // does an inverse launch produce anything from column tile t0 on?
        return !l.ax.pruneOn || t0 * axis_tile_cols(l.L) <= l.ax.pruneCutX;
    }
    // Slab-partitioned sweep as a column-chunked pipeline.  The kx column tiles are split into xChunks chunks; per chunk c
    //   lane 0 (main):  forward y pass of every product group, rows pushed into the owners' receive slots      YF(c)
    //   lane 1:         barrier(c) -> [last-axis forward passes of further groups] -> k stage (pushes its fused
    //                   inverse z part) -> [masked inverse z passes of further dealiased fields, pushed]            KS(c)
    //   lane 2:         barrier(c) -> inverse y passes                                                              YI(c)
    // YF(c+1) -- bound by the NVLink stores -- runs next to KS(c) and YI(c-1), which are bound by local HBM; the barriers wait
    // on per-chunk flags next to the kernels of the other lanes instead of in front of them.  Hazards: a receive slot chunk is
    // rewritten in the next step only after the peers have read it -- forward slots: YF(c, n+1) follows this rank's last
    // barrier of step n on lane 2, which every peer only signals after ALL its k-stage chunks of step n; inverse slots:
    // KS(c, n+1) follows barrier(c, n+1) of lane 1, which a peer only signals from step n+1, i.e. after its YI(., n).
    void emit_exchange_pipeline(std::vector<Launch>& out, const std::vector<Launch>& yf, const std::vector<Launch>& lf,
                                const std::vector<Launch>& kk, const std::vector<Launch>& li, const std::vector<Launch>& yi) {
        const int nct = kk[0].ax.ncolTiles;
        const int nch = std::min(xChunks, nct);
        // tiles per chunk: the remainder goes to the LAST chunks (beyond the dealias cut-off: nothing to send back there)
        std::vector<int> start(nch + 1, 0);
        for (int c = 0; c < nch; ++c) start[c + 1] = start[c] + nct / nch + (c >= nch - nct % nch ? 1 : 0);
        const size_t first = out.size();
        const int pid = nextPipe++;
        for (int c = 0; c < nch; ++c) {
            const int t0 = start[c], nt = start[c + 1] - start[c];
            const int evF = nextEv++, evK = nextEv++;
            for (size_t i = 0; i < yf.size(); ++i) {
                Launch l = column_chunk(yf[i], t0, nt);
                l.lane = 0;
                if (i + 1 == yf.size()) l.recEv = evF;
                out.push_back(l);
            }
            add_barrier(out, "xbar_fwd");
            out.back().lane = 1; out.back().waitEv = evF;
            for (const Launch& z : lf) { Launch l = column_chunk(z, t0, nt); l.lane = 1; out.push_back(l); }
            { Launch l = column_chunk(kk[0], t0, nt); l.lane = 1; out.push_back(l); }
            for (const Launch& z : li) if (chunk_live(z, t0)) { Launch l = column_chunk(z, t0, nt); l.lane = 1; out.push_back(l); }
            out.back().recEv = evK;
            bool anyInv = false;
            for (const Launch& y : yi) anyInv = anyInv || chunk_live(y, t0);
            if (!anyInv && c + 1 < nch) continue;
            add_barrier(out, "xbar_inv");
            out.back().lane = 2; out.back().waitEv = evK;
            for (const Launch& y : yi) if (chunk_live(y, t0)) { Launch l = column_chunk(y, t0, nt); l.lane = 2; out.push_back(l); }
        }
        for (size_t i = first; i < out.size(); ++i) out[i].pipe = pid;
    }

    // Independent launches of one sweep that do not fill the GPU on their own (2-D grids: a strided-axis pass has ncol / 16
    // column tiles and no batch dimension -- 65 clusters at 2048^2, 22 when pruned) go out side by side on the lanes and are
    // joined before the next launch.  Lanes 1, 2 first: a lane forks from everything queued on the main stream so far.
    void emit_concurrent(std::vector<Launch>& out, std::vector<Launch> v) {
        if (v.empty()) return;
        if (v.size() == 1 || dim == 3 || nranks > 1 || !laneFanout) { for (Launch& l : v) out.push_back(l); return; }
        const int gid = nextGroup++, pid = nextPipe++;
        double bytes = 0;
        for (const Launch& l : v) bytes += l.bytes;
        for (size_t i = 0; i < v.size(); ++i) {
            Launch& l = v[i];
            l.lane = (int)((i + 1) % kMaxLanes);
            l.group = gid; l.pipe = pid;
            if (i == 0) { snprintf(l.groupName, sizeof l.groupName, "%s", l.name); l.groupBytes = bytes; }
            out.push_back(l);
        }
    }

    int build_stage(bool dyn, std::vector<Launch>& out) {
        std::vector<int> outs;
        for (size_t f = 0; f < fields.size(); ++f) if (fields[f].dynamic == dyn) outs.push_back((int)f);
        if (outs.empty()) return CUPSS_B200_OK;
        const char* tag = dyn ? "dyn" : "con";

        // ---- product-term groups (terms of one field whose prefactor vectors are proportional share one forward transform)
        std::vector<Group> groups;
        for (int f : outs) {
            for (size_t ti = 0; ti < fields[f].terms.size(); ++ti) {
                const Term& t = fields[f].terms[ti];
                if (t.product.size() == 1) continue;
                bool allZero = true;
                for (const Pres& p : t.pres) if (p.pre != 0.0f) allZero = false;
                if (allZero) continue;   // RHS "0": contributes nothing
                if (t.product.size() > (size_t)XP_MAX_FAC) return fail(CUPSS_B200_ERR_ARG, "product of %zu fields exceeds %d", t.product.size(), XP_MAX_FAC);
                bool placed = false;
                for (Group& g : groups) {
                    float lam;
                    if (g.field == f && proportional(g.pres, t.pres, &lam)) { g.monos.push_back({lam, t.product}); placed = true; break; }
                }
                if (!placed) groups.push_back({f, (int)ti, t.pres, {{1.0f, t.product}}});
            }
        }

        size_t sc = 2;   // scratch 0,1 are reserved for upload/download
        std::vector<const float2*> groupSpec(groups.size(), nullptr);   // input of the last-axis forward pass per group

        // ---- x passes (greedy split under the kernel's descriptor limits)
        std::vector<Launch> xs;
        size_t g0 = 0;
        while (g0 < groups.size()) {
            Launch x{};
            x.kind = Launch::XPASS; x.mode = X_HOT;
            snprintf(x.name, sizeof x.name, "x_%s", tag);
            std::vector<int> ins;
            size_t g1 = g0;
            int nMono = 0;
            const size_t maxIn = (size_t)xpass_max_inputs(sx);
            while (g1 < groups.size()) {
                std::vector<int> trial = ins;
                int monos = nMono;
                for (const Mono& m : groups[g1].monos) {
                    ++monos;
                    for (int fid : m.fac) {
                        bool seen = false;
                        for (int q : trial) if (q == fid) seen = true;
                        if (!seen) trial.push_back(fid);
                    }
                }
                if (g1 > g0 && (trial.size() > maxIn || monos > XP_MAX_MONO || g1 - g0 >= (size_t)XP_MAX_OUT)) break;
                // Long lines: every stashed input costs a line of shared memory (fewer CTAs per SM), so a group joins a launch
                // only when it shares an input with it -- Model H: phi^3 keeps the single-input kernel, vx*iqxphi + vy*iqyphi
                // gets a launch of its own.  The results do not depend on the grouping.
                if (g1 > g0 && xstash1_supported(sx) && trial.size() == ins.size() + group_inputs(groups[g1]).size()) break;
                if (trial.size() > maxIn || monos > XP_MAX_MONO) return fail(CUPSS_B200_ERR_ARG, "a single term group needs too many fields/monomials for sx = %d (at most %zu fields)", sx, maxIn);
                ins = trial; nMono = monos; ++g1;
            }
            x.xa.nIn = (int)ins.size();
            if (ins.empty()) {   // pure constants: still need one (dummy) input line; use the field's own W2
                ins.push_back(groups[g0].field);
                x.xa.nIn = 1;
                Field& F0 = fields[groups[g0].field];
                if (!F0.W2) {
                    CK(cudaMalloc(&F0.W2, specElems * sizeof(float2)));
                    CK(cudaMemsetAsync(F0.W2, 0, specElems * sizeof(float2), stream));
                }
            }
            double inFrac = 0.0;
            for (size_t i = 0; i < ins.size(); ++i) {
                Field& F = fields[ins[i]];
                if (!F.W2) return fail(CUPSS_B200_ERR_STATE, "internal: field %s has no dealiased buffer", F.name.c_str());
                x.xa.in[i] = F.W2;
                short cx, cy, cz;
                cutoffs(F.aliasOrder, &cx, &cy, &cz);
                x.xa.kmax[i] = (prune && !F.w2FullBand) ? (cx < ncol - 1 ? cx : ncol - 1) : ncol - 1;
                inFrac += (double)(x.xa.kmax[i] + 1) / ncol;
            }
            x.xa.nOut = (int)(g1 - g0);
            int mi = 0;
            for (size_t g = g0; g < g1; ++g) {
                float2* w3;
                CKR(get_scratch(sc++, &w3));
                x.xa.out[g - g0] = w3;
                groupSpec[g] = w3;
                for (const Mono& m : groups[g].monos) {
                    XMono xm{};
                    xm.coef = m.coef; xm.out = (signed char)(g - g0); xm.nfac = (signed char)m.fac.size();
                    for (size_t q = 0; q < m.fac.size(); ++q)
                        for (size_t i = 0; i < ins.size(); ++i) if (ins[i] == m.fac[q]) xm.fac[q] = (signed char)i;
                    x.xa.mono[mi++] = xm;
                }
            }
            x.xa.nMono = mi;
            x.xa.pitch = pitch; x.xa.nlines = (long long)zl * sy;
            x.xa.norm = 1.0f / ((float)sx * (float)sy * (float)sz);
            CKR(get_twiddle(sx, &x.xa.tw));
            CKR(get_twiddle_x3(&x.xa.tw3));
            x.bytes = (inFrac + x.xa.nOut) * spec_bytes();
            xs.push_back(x);
            g0 = g1;
        }
        // Single GPU, 3-D: the x passes and the forward y passes both work plane by plane, so they run as a z-chunked software
        // pipeline on two streams (emit_xy_pipeline); otherwise the x passes go out whole, right here.
        const bool xyPipe = dim == 3 && nranks == 1 && zChunk > 0 && zl >= 2 * zChunk && zl % zChunk == 0 && !groups.empty();
        if (!xyPipe) emit_concurrent(out, xs);
        std::vector<Launch> ys;
        // Slab-partitioned run with the fused push exchange: the exchange side of the sweep is collected per kind and emitted as
        // a column-chunked pipeline over three lanes (emit_exchange_pipeline) instead of one launch after the other.
        const bool xPipe = dim == 3 && nranks > 1 && useP2P && xChunks > 1 && !groups.empty();
        std::vector<Launch> pipeYF, pipeLF, pipeK, pipeLI, pipeYI;

        // ---- forward y passes (3-D) and slab exchange
        if (dim == 3) {
            for (size_t g = 0; g < groups.size(); ++g) {
                Launch y{};
                y.kind = Launch::AXIS_PLAIN; y.L = sy; y.dir = -1;
                snprintf(y.name, sizeof y.name, "yfwd_%s", tag);
                CKR(make_y_axis(y.ax, true));
                float2* w4;
                CKR(get_scratch(sc++, &w4));
                y.ax.in = groupSpec[g]; y.ax.out = w4;
                y.bytes = 2.0 * spec_bytes();
                if (nranks > 1 && useP2P) {
                    // fused exchange: the y pass stores each ky row straight into the owning peer's receive slot over NVLink
                    const int slot = arenaNext++;
                    set_push(y.ax, slot, true);
                    y.commBytes = (double)zl * kyl * pitch * 8.0 * (nranks - 1);
                    snprintf(y.name, sizeof y.name, "yfwd_push_%s", tag);
                    if (xPipe) pipeYF.push_back(y);
                    else { out.push_back(y); add_barrier(out, "xbar_fwd"); }
                    groupSpec[g] = arena_slot_ptr(rank, slot);
                    continue;
                }
                if (xyPipe) ys.push_back(y); else out.push_back(y);
                groupSpec[g] = w4;
                if (nranks > 1) {
                    float2* r;
                    CKR(get_scratch(sc++, &r));
                    CKR(add_a2a(out, "a2a_fwd", w4, r));
                    groupSpec[g] = r;
                }
            }
        }
        if (xyPipe) emit_xy_pipeline(out, xs, ys, tag);

        // ---- k stage
        Launch k{};
        k.kind = Launch::AXIS_KSTAGE;
        k.joinBefore = xyPipe;
        snprintf(k.name, sizeof k.name, "kstage_%s", tag);
        CKR(make_last_axis(k.ax, &k.L));
        KStageD& ks = k.ks;
        ks.dt = dt; ks.sdt = 1.0f / std::sqrt(dt); ks.noiseBase = dt / (dx * dy * dz);
        ks.stepqx = 2.0f * 3.1415926535f / (dx * (float)sx);   // PI as in inc/cupss/defines.h:54
        ks.stepqy = 2.0f * 3.1415926535f / (dy * (float)sy);
        ks.stepqz = 2.0f * 3.1415926535f / (dz * (float)sz);
        ks.sx = sx; ks.sy = sy; ks.sz = sz;
        ks.stepCounter = stepCounter;
        ks.seed = 0;
        philox_round_keys(0, ks.philoxKey);
        ks.noiseField = -1;
        ks.whiteSelf = std::sqrt((float)sx * (float)sy * (float)sz);
        ks.whitePair = std::sqrt(0.5f * ((float)sx * (float)sy * (float)sz));
        ks.hasFwd = groups.empty() ? 0 : 1;
        if (ks.hasFwd) k.ax.in = groupSpec[0];

        std::map<int, int> srcOfField;
        auto srcField = [&](int f) -> int {
            auto it = srcOfField.find(f);
            if (it != srcOfField.end()) return it->second;
            if (ks.nsrc >= KS_MAX_SRC) return -100;
            ks.src[ks.nsrc] = fields[f].S;
            srcOfField[f] = ks.nsrc;
            return ks.nsrc++;
        };
        // groups beyond the first get their own last-axis forward pass into a spectrum that the k stage reads pointwise
        std::vector<int> srcOfGroup(groups.size(), -1);
        std::vector<Launch> lastFwds;
        for (size_t g = 1; g < groups.size(); ++g) {
            Launch z{};
            z.kind = Launch::AXIS_PLAIN; z.dir = -1;
            snprintf(z.name, sizeof z.name, "lastfwd_%s", tag);
            CKR(make_last_axis(z.ax, &z.L));
            float2* that;
            CKR(get_scratch(sc++, &that));
            z.ax.in = groupSpec[g]; z.ax.out = that;
            z.bytes = 2.0 * spec_bytes();
            if (xPipe) pipeLF.push_back(z); else lastFwds.push_back(z);
            if (ks.nsrc >= KS_MAX_SRC) return fail(CUPSS_B200_ERR_ARG, "too many k-stage sources");
            ks.src[ks.nsrc] = that;
            srcOfGroup[g] = ks.nsrc++;
        }
        emit_concurrent(out, lastFwds);

        int nterm = 0, npres = 0;
        int invField = -1;
        std::vector<int> extraInv;
        for (int f : outs) {
            Field& F = fields[f];
            if (ks.nout >= KS_MAX_OUT) return fail(CUPSS_B200_ERR_ARG, "more than %d fields in one sweep", KS_MAX_OUT);
            OutD& od = ks.out[ks.nout];
            od = OutD{};
            od.dynamic = F.dynamic; od.fieldId = (signed char)f;
            // The old value of the field itself is a source only where the evaluator uses it: a dynamic field (Euler update) or a
            // constraint field without a right-hand side term (it keeps its value).  A constraint field WITH terms is assigned
            // from them (kstage.cuh), so registering its spectrum would only cost loads and inflate the byte model.
            size_t emitted = 0;
            for (size_t ti = 0; ti < F.terms.size(); ++ti) {
                if (F.terms[ti].product.size() == 1) { ++emitted; continue; }
                for (size_t g = 0; g < groups.size(); ++g) if (groups[g].field == f && groups[g].firstTerm == (int)ti) ++emitted;
            }
            int self = -1;
            if (F.dynamic || emitted == 0) {
                self = srcField(f);
                if (self < 0) return fail(CUPSS_B200_ERR_ARG, "too many k-stage sources");
            }
            od.selfSrc = (signed char)self;
            od.dst = (signed char)ks.nout;
            ks.dst[ks.nout] = F.S;
            od.termOff = (short)nterm;
            for (size_t ti = 0; ti < F.terms.size(); ++ti) {
                const Term& t = F.terms[ti];
                int src = -100;
                const std::vector<Pres>* pv = &t.pres;
                if (t.product.size() == 1) {
                    src = srcField(t.product[0]);
                    if (src < 0) return fail(CUPSS_B200_ERR_ARG, "too many k-stage sources");
                } else {
                    bool found = false;
                    for (size_t g = 0; g < groups.size(); ++g)
                        if (groups[g].field == f && groups[g].firstTerm == (int)ti) { src = g == 0 ? -1 : srcOfGroup[g]; pv = &groups[g].pres; found = true; }
                    if (!found) continue;   // merged into an earlier group, or identically zero
                }
                if (nterm >= KS_MAX_TERM || npres + (int)pv->size() > KS_MAX_PRES) return fail(CUPSS_B200_ERR_ARG, "too many terms/prefactors in one sweep");
                TermD td{};
                fold_pres(*pv, ks, npres, td);
                td.src = (signed char)src;
                ks.term[nterm++] = td;
                od.nterm++;
            }
            od.impOff = (short)npres; od.nimp = (signed char)F.implicit.size();
            if (npres + (int)F.implicit.size() > KS_MAX_PRES) return fail(CUPSS_B200_ERR_ARG, "too many prefactors in one sweep");
            for (const Pres& p : F.implicit) {
                PresD d{};
                d.pre = p.pre; d.q2n = (signed char)p.q2n; d.invq = (signed char)p.invq;
                ks.pres[npres++] = d;
            }
            od.noisy = F.noisy;
            if (F.noisy) {
                od.noise.pre = F.noise.pre; od.noise.q2n = (signed char)F.noise.q2n; od.noise.invq = (signed char)F.noise.invq;
                od.noiseAmp0 = std::sqrt(ks.noiseBase * F.noise.pre);
                ks.seed = F.seed;
                philox_round_keys(F.seed, ks.philoxKey);
                if (ks.noiseField < 0) ks.noiseField = f;
            }
            cutoffs(F.aliasOrder, &od.cutx, &od.cuty, &od.cutz);
            if (F.needsAlias) {
                if (invField < 0) { invField = f; od.inv = 1; } else extraInv.push_back(f);
            }
            ks.nout++;
        }
        ks.hasInv = invField >= 0 ? 1 : 0;
        ks.usesInvq = 0;
        for (int i = 0; i < npres; ++i) if (ks.pres[i].invq) ks.usesInvq = 1;
        for (int o = 0; o < ks.nout; ++o) if (ks.out[o].noisy && ks.out[o].noise.invq) ks.usesInvq = 1;
        // lean evaluator when the sweep is a single noise-free dynamic field whose prefactors depend on q^2 only
        ks.fastKind = KS_GENERIC;
        // (a noisy field: the lean evaluator compiled at run time for its signature, if its noise amplitude needs q^2 only)
        const bool leanNoise = ks.nout == 1 && ks.out[0].noisy;
        if (ks.nout == 1 && ks.out[0].dynamic && (!leanNoise || (ks.out[0].noise.invq == 0 && FftLevelsN(k.L) > 1 && !getenv("CUPSS_B200_NO_LEAN_NOISE"))) &&
            ks.nsrc == 1 && extraInv.empty() && nterm <= 1 && ks.out[0].nimp <= 4 && !getenv("CUPSS_B200_GENERIC_KSTAGE")) {
            bool ok = true;
            for (int i = 0; i < npres; ++i)
                if (ks.pres[i].iqx || ks.pres[i].iqy || ks.pres[i].iqz || ks.pres[i].invq || ks.pres[i].q2n < 0 || ks.pres[i].q2n > 3) ok = false;
            if (nterm == 1 && (ks.term[0].mulI || ks.term[0].src > 0 || ks.term[0].npres > 3)) ok = false;
            if (ok) {
                ScalarQ2D& q = ks.sq2;
                q = ScalarQ2D{};
                q.hasTerm = (signed char)nterm;
                if (nterm == 1) {
                    q.ntp = ks.term[0].npres;
                    q.termFused = ks.term[0].src < 0 ? 1 : 0;
                    for (int i = 0; i < q.ntp; ++i) { q.tpre[i] = (double)ks.pres[ks.term[0].presOff + i].pre; q.tn[i] = ks.pres[ks.term[0].presOff + i].q2n; }
                }
                q.nimp = ks.out[0].nimp;
                for (int i = 0; i < q.nimp; ++i) { q.ipre[i] = (double)ks.pres[ks.out[0].impOff + i].pre; q.in[i] = ks.pres[ks.out[0].impOff + i].q2n; }
                ks.fastKind = KS_SCALAR_Q2;
                if (leanNoise) {
                    const char* je = getenv("CUPSS_B200_JIT");
                    std::string why = "CUPSS_B200_JIT=0";
                    if (!(je && je[0] == '0')) k.jitFn = jit_kstage_function(ks, k.L, &why, sq2_signature(q) | SQ2_SIG_NOISE);
                    if (!k.jitFn) {   // no NVRTC: the generic evaluator (below) takes the sweep
                        if (getenv("CUPSS_B200_VERBOSE")) fprintf(stderr, "cupss_b200: lean noisy k stage not compiled (%s)\n", why.c_str());
                        ks.fastKind = KS_GENERIC;
                    }
                }
            }
        }
        if (ks.fastKind == KS_GENERIC) {
            // plan-specialised kernel for grids where the interpreter's overhead matters (or on request)
            const char* je = getenv("CUPSS_B200_JIT");
            const bool force = je && je[0] == '1', off = je && je[0] == '0';
            const double points = (double)sx * sy * sz;
            if (!off && (force || points >= (double)(1 << 18)) && FftLevelsN(k.L) > 1) {
                std::string why;
                k.jitFn = jit_kstage_function(ks, k.L, &why);
                if (!k.jitFn && (force || getenv("CUPSS_B200_VERBOSE"))) fprintf(stderr, "cupss_b200: k stage not specialised (%s); using the interpreter\n", why.c_str());
            }
        }
        float2* invOut = nullptr;
        double invLive = 1.0;   // share of the fused inverse's output that is not pruned away (written / pushed at all)
        if (ks.hasInv) {
            if (dim == 3) CKR(get_scratch(sc++, &invOut)); else invOut = fields[invField].W2;
        }
        k.ax.out = invOut ? invOut : fields[outs[0]].S;   // only its addressing is used when there is no fused inverse
        k.bytes = (double)(ks.hasFwd + ks.nsrc + ks.nout + ks.hasInv) * spec_bytes();
        if (ks.hasInv && prune) {
            short cx, cy, cz;
            cutoffs(fields[invField].aliasOrder, &cx, &cy, &cz);
            k.ax.pruneOn = 1; k.ax.pruneCutX = cx; k.ax.pruneCutY = cy;
            const int C = axis_tile_cols(k.L);
            const double fx = (double)std::min(ncol, (cx / C + 1) * C) / ncol;
            const double fy = dim == 3 ? std::min(1.0, (2.0 * cy + 1.0) / sy) : 1.0;
            k.bytes = (double)(ks.hasFwd + ks.nsrc + ks.nout) * spec_bytes() + fx * fy * spec_bytes();
            invLive = fx * fy;
        }
        const bool pushInv = nranks > 1 && useP2P && dim == 3;
        std::vector<std::pair<int, float2*>> w1s;   // (field, input of its inverse y pass)
        std::vector<int> pushed;                    // same order: 1 if that input already sits in the local arena slot
        if (ks.hasInv && pushInv) {
            const int slot = arenaNext++;
            set_push(k.ax, slot, false);
            k.commBytes = invLive * (double)zl * kyl * pitch * 8.0 * (nranks - 1);   // average over the ranks (exact per rank with the cyclic ky distribution)
            snprintf(k.name, sizeof k.name, "kstage_push_%s", tag);
            if (xPipe) pipeK.push_back(k);
            else { out.push_back(k); add_barrier(out, "xbar_inv"); }
            w1s.push_back({invField, arena_slot_ptr(rank, slot)});
            pushed.push_back(1);
        } else {
            if (xPipe) pipeK.push_back(k); else out.push_back(k);
            if (invField >= 0 && dim == 3) { w1s.push_back({invField, invOut}); pushed.push_back(0); }
        }

        // ---- remaining inverse transforms of dealiased fields
        std::vector<Launch> lastInvs;
        for (int f : extraInv) {
            Launch z;
            CKR(lastinv_launch(f, tag, z));
            float2* w1 = fields[f].W2;
            if (dim == 3) CKR(get_scratch(sc++, &w1));
            z.ax.out = w1;
            if (pushInv) {
                const int slot = arenaNext++;
                set_push(z.ax, slot, false);
                {
                    double live = 1.0;
                    if (z.ax.pruneOn) {
                        const int C = axis_tile_cols(z.L);
                        live = (double)std::min(ncol, (z.ax.pruneCutX / C + 1) * C) / ncol * std::min(1.0, (2.0 * z.ax.pruneCutY + 1.0) / sy);
                    }
                    z.commBytes = live * (double)zl * kyl * pitch * 8.0 * (nranks - 1);
                }
                snprintf(z.name, sizeof z.name, "lastinv_push_%s", tag);
                if (xPipe) pipeLI.push_back(z);
                else { out.push_back(z); add_barrier(out, "xbar_inv"); }
                w1s.push_back({f, arena_slot_ptr(rank, slot)});
                pushed.push_back(1);
                continue;
            }
            lastInvs.push_back(z);
            if (dim == 3) { w1s.push_back({f, w1}); pushed.push_back(0); }
        }
        emit_concurrent(out, lastInvs);
        for (size_t wi = 0; wi < w1s.size(); ++wi) {
            auto& pr = w1s[wi];
            const float2* yin = pr.second;
            if (nranks > 1 && !pushed[wi]) {
                float2* r;
                CKR(get_scratch(sc++, &r));
                CKR(add_a2a(out, "a2a_inv", pr.second, r));
                yin = r;
            }
            Launch y;
            CKR(yinv_launch(pr.first, tag, yin, y));
            if (xPipe) pipeYI.push_back(y); else out.push_back(y);
        }
        if (xPipe) emit_exchange_pipeline(out, pipeYF, pipeLF, pipeK, pipeLI, pipeYI);
        return CUPSS_B200_OK;
    }

    int drop_graph() {
        if (graphExec) { cudaGraphExecDestroy(graphExec); graphExec = nullptr; }
        return CUPSS_B200_OK;
    }

    int finalize() {
        if (nranks > 1 && dim != 3) return fail(CUPSS_B200_ERR_ARG, "slab partitioning needs a 3-D grid");
        drop_graph();
        // aliasing flags (term::prepareDevice, src/term_init.cpp:118-126)
        for (Field& F : fields) { F.needsAlias = false; F.aliasOrder = 1; }
        for (Field& F : fields)
            for (Term& t : F.terms)
                if (t.product.size() != 1)
                    for (int g : t.product) {
                        fields[g].needsAlias = true;
                        if (fields[g].aliasOrder < (int)t.product.size()) fields[g].aliasOrder = (int)t.product.size();
                    }
        for (Field& F : fields) {
            if (!F.S) {
                CK(cudaMalloc(&F.S, specElems * sizeof(float2)));
                CK(cudaMemsetAsync(F.S, 0, specElems * sizeof(float2), stream));
            }
            if (F.needsAlias && !F.W2) {
                CK(cudaMalloc(&F.W2, specElems * sizeof(float2)));
                CK(cudaMemsetAsync(F.W2, 0, specElems * sizeof(float2), stream));   // real_dealiased starts at zero
            }
            if (F.needsAlias) {
                short cx, cy, cz;
                cutoffs(F.aliasOrder, &cx, &cy, &cz);
                const int now[3] = {prune ? cx : -2, prune ? cy : -2, prune ? cz : -2};
                if (F.w2cut[0] != -1 && (now[0] != F.w2cut[0] || now[1] != F.w2cut[1] || now[2] != F.w2cut[2]))
                    CK(cudaMemsetAsync(F.W2, 0, specElems * sizeof(float2), stream));   // pruned regions must read as zero
                F.w2cut[0] = now[0]; F.w2cut[1] = now[1]; F.w2cut[2] = now[2];
            }
        }
        if (nranks > 1 && useP2P) {
            // dry run to count the receive slots, then (collectively) size the arena and build for real
            std::vector<Launch> dry;
            arenaNext = 0; nextPt = 0;
            CKR(build_stage(false, dry));
            CKR(build_stage(true, dry));
            CKR(ensure_arena((size_t)arenaNext));
        }
        step.clear();
        arenaNext = 0; nextPt = 0; nextEv = 0; nextGroup = 0; nextPipe = 0;
        CKR(build_stage(false, step));
        stageSplit = step.size();
        CKR(build_stage(true, step));
        for (size_t i = 1; i < step.size(); ++i)   // the launch that follows a multi-lane pipeline waits for all of its lanes
            if (step[i - 1].pipe >= 0 && step[i].pipe != step[i - 1].pipe) step[i].joinBefore = true;
        if (nextPt > XH_EPOCH / CUPSS_MAX_PEERS) return fail(CUPSS_B200_ERR_ARG, "too many exchange points (%d)", nextPt);
        Launch b{};
        b.kind = Launch::BUMP;
        snprintf(b.name, sizeof b.name, "bump");
        step.push_back(b);
        CK(cudaStreamSynchronize(stream));
        finalized = true;
        return CUPSS_B200_OK;
    }

    int capture_graph() {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int r = run_list(step);
        cudaError_t e = cudaStreamEndCapture(stream, &g);
        if (r != CUPSS_B200_OK) { if (g) cudaGraphDestroy(g); return r; }
        CK(e);
        CK(cudaGraphInstantiate(&graphExec, g, 0));
        CK(cudaGraphDestroy(g));
        return CUPSS_B200_OK;
    }

    // One sweep class of one step, eagerly (user callbacks run between the sweeps; field::setRHS, src/field.cpp:68-86).
    int do_stage(int which) {
        if (!finalized) return fail(CUPSS_B200_ERR_STATE, "step before finalize");
        const size_t b = which == 0 ? 0 : stageSplit, e = which == 0 ? stageSplit : step.size();
        return run_range(step, b, e);
    }
    // Real view of a field for a callback: which = 0 the field itself, 1 its dealiased copy (what products read).
    // Partitioned plans: the view is this rank's z-slab [zl][sy][sx] (collective for which == 0: the full transform exchanges slabs).
    int view_begin(int f, int which, float2** dev) {
        Field& F = fields[f];
        const size_t n = (size_t)sx * sy * zl;
        if (!realBuf) CK(cudaMalloc(&realBuf, n * sizeof(float)));
        if (!viewBuf) CK(cudaMalloc(&viewBuf, n * sizeof(float2)));
        if (which == 0) {
            if (!F.S) return fail(CUPSS_B200_ERR_STATE, "field %s has no device data", F.name.c_str());
            CKR(inverse_full(F.S));
        } else {
            if (!F.W2) return fail(CUPSS_B200_ERR_STATE, "field %s has no dealiased copy", F.name.c_str());
            Launch x{};
            x.kind = Launch::XPASS; x.mode = X_C2R_ONLY;
            // the fresh dealiased copy is band-limited; column tiles beyond the cut-off are never rewritten by the pruned inverse
            // passes and may still hold what an earlier callback commit left there
            short cx, cy, cz;
            cutoffs(F.aliasOrder, &cx, &cy, &cz);
            x.xa.nIn = 1; x.xa.in[0] = F.W2; x.xa.kmax[0] = prune ? (cx < ncol - 1 ? cx : ncol - 1) : ncol - 1;
            x.xa.pitch = pitch; x.xa.nlines = (long long)zl * sy;
            x.xa.norm = 1.0f / ((float)sx * (float)sy * (float)sz);
            x.xa.realOut = realBuf;
            CKR(get_twiddle(sx, &x.xa.tw));
            CKR(run_launch(x));
        }
        CK(launch_real_expand(realBuf, viewBuf, n, stream));
        CK(cudaStreamSynchronize(stream));
        *dev = viewBuf;
        return CUPSS_B200_OK;
    }
    int view_commit(int f, int which) {
        Field& F = fields[f];
        const size_t n = (size_t)sx * sy * zl;
        if (!viewBuf || !realBuf) return fail(CUPSS_B200_ERR_STATE, "view_commit without view_begin");
        CK(cudaDeviceSynchronize());   // the callback's kernels run on the legacy default stream
        CK(launch_real_compress(viewBuf, realBuf, n, stream));
        if (which == 0) {
            CKR(forward_full(F.S));
        } else {
            Launch x{};
            x.kind = Launch::XPASS; x.mode = X_R2C_ONLY;
            x.xa.nOut = 1; x.xa.out[0] = F.W2;
            x.xa.pitch = pitch; x.xa.nlines = (long long)zl * sy;
            x.xa.norm = (float)sy * (float)sz;   // real = W2-convention / (sx*sy*sz) and R2C(C2R(.)) = sx * (.)
            x.xa.realIn = realBuf;
            CKR(get_twiddle(sx, &x.xa.tw));
            CKR(run_launch(x));
            // What the callback left in the dealiased copy is not band-limited in x any more, and the reference feeds it to
            // computeProduct unchanged (src/field.cpp:75-83, src/term.cpp:85-92): from now on the hot x pass loads the full
            // band of this field instead of stopping at the dealias cut-off.
            if (!F.w2FullBand) {
                F.w2FullBand = true;
                for (Launch& l : step)
                    if (l.kind == Launch::XPASS)
                        for (int i = 0; i < l.xa.nIn; ++i)
                            if (l.xa.in[i] == F.W2) l.xa.kmax[i] = ncol - 1;
                drop_graph();
            }
        }
        return CUPSS_B200_OK;
    }

    // Fourier view of a field for callbackFourier (src/field.cpp:48-57): the full float2[sz][sy][sx] spectrum of the
    // reference, rebuilt from the Hermitian half.  The commit keeps the Hermitian part of what the callback left --
    // exactly what survives the reference's toReal -> normalize (real part) -> toComp that follows -- and redoes the
    // dealiased copy, which the fused k stage had produced from the pre-callback spectrum.
    int comp_view_begin(int f, float2** dev) {
        if (nranks != 1) return fail(CUPSS_B200_ERR_ARG, "user callbacks are single-GPU only");
        Field& F = fields[f];
        if (!F.S) return fail(CUPSS_B200_ERR_STATE, "field %s has no device data", F.name.c_str());
        if (!viewBuf) CK(cudaMalloc(&viewBuf, (size_t)sx * sy * zl * sizeof(float2)));
        CK(launch_spectrum_expand(F.S, viewBuf, sx, sy, sz, pitch, sy, 0, sz, 1, stream));
        CK(cudaStreamSynchronize(stream));
        *dev = viewBuf;
        return CUPSS_B200_OK;
    }
    int comp_view_commit(int f) {
        Field& F = fields[f];
        if (!viewBuf) return fail(CUPSS_B200_ERR_STATE, "comp_view_commit without comp_view_begin");
        CK(cudaDeviceSynchronize());   // the callback's kernels run on the legacy default stream
        CK(launch_spectrum_compress(viewBuf, F.S, sx, sy, sz, pitch, stream));
        return refresh_w2(f);
    }

    int do_steps(int n) {
        if (!finalized) return fail(CUPSS_B200_ERR_STATE, "step before finalize");
        if (useGraph && !graphExec) {
            // run one step eagerly first so every kernel's attributes are set outside the capture
            CKR(run_list(step));
            --n;
            CKR(capture_graph());
        }
        for (int i = 0; i < n; ++i) {
            if (useGraph) CK(cudaGraphLaunch(graphExec, stream));
            else CKR(run_list(step));
        }
        return CUPSS_B200_OK;
    }
};

// ---------------------------------------------------------------- C ABI
extern "C" {

const char* cupss_b200_last_error(void) { return g_err; }

int cupss_b200_create(cupss_b200_plan** out, int sx, int sy, int sz, float dx, float dy, float dz, float dt) {
    if (!out) return fail(CUPSS_B200_ERR_ARG, "null out pointer");
    if (sx < 2 || sy < 1 || sz < 1) return fail(CUPSS_B200_ERR_ARG, "bad grid %dx%dx%d", sx, sy, sz);
    if (!fft_size_supported(sx) || !fft_size_supported(sy) || !fft_size_supported(sz) || sx > 8192)
        return fail(CUPSS_B200_ERR_ARG, "grid %dx%dx%d: every axis must be a power of two <= 8192", sx, sy, sz);
    if (sy == 1 && sz > 1) return fail(CUPSS_B200_ERR_ARG, "sy == 1 with sz > 1 is not a valid cuPSS grid");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CUPSS_B200_ERR_CUDA, "no CUDA device available (%s): the B200 engine has no CPU fallback", cudaGetErrorString(e));
    cupss_b200_plan* p = new cupss_b200_plan();
    p->sx = sx; p->sy = sy; p->sz = sz; p->dx = dx; p->dy = dy; p->dz = dz; p->dt = dt;
    p->dim = sz > 1 ? 3 : (sy > 1 ? 2 : 1);
    p->zl = sz; p->kyl = sy;
    p->ncol = sx / 2 + 1;
    p->pitch = (p->ncol + 15) / 16 * 16;
    p->specElems = (size_t)p->pitch * sy * sz;
    if (p->specElems >= (1ull << 31)) { delete p; return fail(CUPSS_B200_ERR_ARG, "grid %dx%dx%d: the kernels index one array with 32 bits (< 2^31 complex elements)", sx, sy, sz); }
    const char* ng = getenv("CUPSS_B200_NO_GRAPH");
    p->useGraph = !(ng && ng[0] == '1');
    const char* np_ = getenv("CUPSS_B200_NO_PRUNE");
    p->prune = !(np_ && np_[0] == '1');
    if (const char* zc = getenv("CUPSS_B200_ZCHUNK")) p->zChunk = atoi(zc);
    if (const char* zl = getenv("CUPSS_B200_ZCHUNK_LANES")) p->zChunkLanes = atoi(zl);
    if (const char* xc = getenv("CUPSS_B200_XCHUNKS")) p->xChunks = std::max(1, atoi(xc));
    if (getenv("CUPSS_B200_NO_FANOUT")) p->laneFanout = false;
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&p->ev0));
    CK(cudaEventCreate(&p->ev1));
    CK(cudaMalloc(&p->stepCounter, sizeof(unsigned int)));
    CK(cudaMemsetAsync(p->stepCounter, 0, sizeof(unsigned int), p->stream));
    *out = p;
    return CUPSS_B200_OK;
}

void cupss_b200_destroy(cupss_b200_plan* p) {
    if (!p) return;
    if (p->stream) cudaStreamSynchronize(p->stream);
    p->drop_graph();
    for (Field& F : p->fields) { if (F.S) cudaFree(F.S); if (F.W2) cudaFree(F.W2); }
    for (float2* s : p->scratch) if (s) cudaFree(s);
    for (auto& kv : p->twiddles) cudaFree(kv.second);
    if (p->realBuf) cudaFree(p->realBuf);
    if (p->viewBuf) cudaFree(p->viewBuf);
    if (p->stepCounter) cudaFree(p->stepCounter);
    for (int d = 0; d < p->nranks; ++d)
        if (d != p->rank && p->peerArena[d]) cudaIpcCloseMemHandle(p->peerArena[d]);
    if (p->arena) cudaFree(p->arena);
    if (p->hostErr) cudaFreeHost(p->hostErr);
    if (p->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    for (cudaEvent_t e : p->evPool) cudaEventDestroy(e);
    for (int ln = 1; ln < cupss_b200_plan::kMaxLanes; ++ln) {
        if (p->forkEv[ln]) cudaEventDestroy(p->forkEv[ln]);
        if (p->joinEv[ln]) cudaEventDestroy(p->joinEv[ln]);
        if (p->laneStream[ln]) cudaStreamDestroy(p->laneStream[ln]);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int cupss_b200_nccl_unique_id(void* id128) {
    CKR(load_nccl());
    NK(g_nccl.GetUniqueId(id128));
    return CUPSS_B200_OK;
}

int cupss_b200_set_partition(cupss_b200_plan* p, int rank, int nranks, const void* id) {
    if (!p || nranks < 1 || rank < 0 || rank >= nranks) return fail(CUPSS_B200_ERR_ARG, "bad rank %d/%d", rank, nranks);
    if (!p->fields.empty() && p->fields[0].S) return fail(CUPSS_B200_ERR_STATE, "set_partition must precede finalize/upload");
    if (nranks == 1) return CUPSS_B200_OK;
    if (p->dim != 3) return fail(CUPSS_B200_ERR_ARG, "slab partitioning needs a 3-D grid");
    if (p->sz % nranks || p->sy % nranks || (nranks & (nranks - 1))) return fail(CUPSS_B200_ERR_ARG, "sz and sy must be divisible by a power-of-two rank count");
    CKR(load_nccl());
    Id128 uid;
    memcpy(uid.b, id, 128);
    NK(g_nccl.CommInitRank(&p->comm, nranks, uid, rank));
    p->rank = rank; p->nranks = nranks;
    if (nranks > CUPSS_MAX_PEERS) return fail(CUPSS_B200_ERR_ARG, "at most %d ranks", CUPSS_MAX_PEERS);
    const char* na = getenv("CUPSS_B200_NCCL_A2A");
    p->useP2P = !(na && na[0] == '1');
    const char* kb = getenv("CUPSS_B200_KY_BLOCK");
    p->kyCyclic = !(kb && kb[0] == '1');
    p->zl = p->sz / nranks; p->kyl = p->sy / nranks;
    p->specElems = (size_t)p->pitch * p->kyl * p->sz;   // == pitch * sy * zl
    if (!p->hostErr) {
        CK(cudaHostAlloc(reinterpret_cast<void**>(&p->hostErr), sizeof(int), cudaHostAllocMapped));
        *p->hostErr = 0;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->hostErrDev), p->hostErr, 0));
    }
    if (const char* to = getenv("CUPSS_B200_BARRIER_TIMEOUT_S")) p->barrierTimeoutNs = (unsigned long long)(atof(to) * 1e9);
    return CUPSS_B200_OK;
}

// A cross-GPU barrier that timed out has trapped: every later CUDA call on this context fails with a generic launch error.
// Entry points that run or consume steps pass their status through here so that the caller sees the actual cause, and a
// time-out is reported even when the failing call itself was not the one that noticed.
static int comm_guard(cupss_b200_plan* p, int rc) {
    if (p && p->hostErr && *reinterpret_cast<volatile int*>(p->hostErr))
        return fail(CUPSS_B200_ERR_COMM, "cross-GPU barrier timed out after %.0f s (a peer rank stopped or fell behind; CUPSS_B200_BARRIER_TIMEOUT_S): "
                                         "the step was aborted before it could read stale receive slots", (double)p->barrierTimeoutNs * 1e-9);
    return rc;
}

int cupss_b200_add_field(cupss_b200_plan* p, const char* name, int dynamic) {
    if (!p || !name) { fail(CUPSS_B200_ERR_ARG, "null argument"); return -CUPSS_B200_ERR_ARG; }
    Field F;
    F.name = name; F.dynamic = dynamic != 0;
    p->fields.push_back(F);
    p->finalized = false;
    return (int)p->fields.size() - 1;
}

static int check_field(cupss_b200_plan* p, int f) {
    if (!p || f < 0 || f >= (int)p->fields.size()) return fail(CUPSS_B200_ERR_ARG, "bad field id %d", f);
    return CUPSS_B200_OK;
}
static Pres to_pres(const cupss_b200_pres& q) { return Pres{q.preFactor, q.q2n, q.iqx, q.iqy, q.iqz, q.invq}; }

int cupss_b200_set_implicit(cupss_b200_plan* p, int f, const cupss_b200_pres* pres, int n) {
    CKR(check_field(p, f));
    p->fields[f].implicit.clear();
    for (int i = 0; i < n; ++i) p->fields[f].implicit.push_back(to_pres(pres[i]));
    p->finalized = false;
    return CUPSS_B200_OK;
}
int cupss_b200_clear_terms(cupss_b200_plan* p, int f) {
    CKR(check_field(p, f));
    p->fields[f].terms.clear();
    p->finalized = false;
    return CUPSS_B200_OK;
}
int cupss_b200_add_term(cupss_b200_plan* p, int f, const cupss_b200_pres* pres, int n, const int* product, int m) {
    CKR(check_field(p, f));
    if (n < 1) return fail(CUPSS_B200_ERR_ARG, "a term needs at least one prefactor");
    Term t;
    for (int i = 0; i < n; ++i) t.pres.push_back(to_pres(pres[i]));
    for (int i = 0; i < m; ++i) {
        CKR(check_field(p, product[i]));
        t.product.push_back(product[i]);
    }
    p->fields[f].terms.push_back(t);
    p->finalized = false;
    return CUPSS_B200_OK;
}
int cupss_b200_set_noise(cupss_b200_plan* p, int f, const cupss_b200_pres* amp, unsigned long long seed) {
    CKR(check_field(p, f));
    p->fields[f].noisy = amp != nullptr;
    if (amp) p->fields[f].noise = to_pres(*amp);
    p->fields[f].seed = seed;
    p->finalized = false;
    return CUPSS_B200_OK;
}
int cupss_b200_set_dealias_rule(cupss_b200_plan* p, int rule) {
    if (!p || (rule != CUPSS_B200_DEALIAS_GPU_RULE && rule != CUPSS_B200_DEALIAS_CPU_RULE)) return fail(CUPSS_B200_ERR_ARG, "bad dealias rule");
    p->dealiasRule = rule;
    p->finalized = false;
    return CUPSS_B200_OK;
}

int cupss_b200_finalize(cupss_b200_plan* p) {
    if (!p) return fail(CUPSS_B200_ERR_ARG, "null plan");
    return p->finalize();
}

static int ensure_real_buf(cupss_b200_plan* p) {
    if (!p->realBuf) CK(cudaMalloc(&p->realBuf, (size_t)p->sx * p->sy * p->zl * sizeof(float)));
    return CUPSS_B200_OK;
}

static int ensure_view_buf(cupss_b200_plan* p) {
    if (!p->viewBuf) CK(cudaMalloc(&p->viewBuf, (size_t)p->sx * p->sy * p->zl * sizeof(float2)));
    return CUPSS_B200_OK;
}

// Host arrays keep the reference's float2 layout (value in .x); the float2 <-> float conversion and the reconstruction of
// the full spectrum from the Hermitian half run on the device, so the host only sees two plain copies.
int cupss_b200_upload_real(cupss_b200_plan* p, int f, const float* host) {
    CKR(check_field(p, f));
    if (!host) return fail(CUPSS_B200_ERR_ARG, "null host pointer");
    Field& F = p->fields[f];
    if (!F.S) {
        CK(cudaMalloc(&F.S, p->specElems * sizeof(float2)));
        CK(cudaMemsetAsync(F.S, 0, p->specElems * sizeof(float2), p->stream));
    }
    CKR(ensure_real_buf(p));
    CKR(ensure_view_buf(p));
    const size_t n = (size_t)p->sx * p->sy * p->zl;
    CK(cudaMemcpyAsync(p->viewBuf, host, n * sizeof(float2), cudaMemcpyHostToDevice, p->stream));
    CK(launch_real_compress(p->viewBuf, p->realBuf, n, p->stream));
    CKR(p->forward_full(F.S));
    CK(cudaStreamSynchronize(p->stream));
    return CUPSS_B200_OK;
}

static int download_real_impl(cupss_b200_plan* p, int f, float* host) {
    CKR(check_field(p, f));
    Field& F = p->fields[f];
    if (!F.S || !host) return fail(CUPSS_B200_ERR_STATE, "field %s has no device data", F.name.c_str());
    CKR(ensure_real_buf(p));
    CKR(ensure_view_buf(p));
    CKR(p->inverse_full(F.S));
    const size_t n = (size_t)p->sx * p->sy * p->zl;
    CK(launch_real_expand(p->realBuf, p->viewBuf, n, p->stream));
    CK(cudaMemcpyAsync(host, p->viewBuf, n * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return CUPSS_B200_OK;
}
int cupss_b200_download_real(cupss_b200_plan* p, int f, float* host) {
    if (!p) return fail(CUPSS_B200_ERR_ARG, "null plan");
    return comm_guard(p, download_real_impl(p, f, host));
}

// Full spectrum float2[.][sy][sx] of the reference's comp_array (src/evolver.cpp:364-368 copies it next to the real array).
// Partitioned plans (collective: every rank calls it): the ky-slabs of all ranks are all-gathered and each rank expands the
// kz planes of its own z-slab, so `host` receives [zl][sy][sx] -- the same slab convention as upload_real / download_real.
static int download_comp_impl(cupss_b200_plan* p, int f, float* host) {
    CKR(check_field(p, f));
    Field& F = p->fields[f];
    if (!F.S || !host) return fail(CUPSS_B200_ERR_STATE, "field %s has no device data", F.name.c_str());
    CKR(ensure_view_buf(p));
    const size_t n = (size_t)p->sx * p->sy * p->zl;
    float2* all = nullptr;
    const float2* half = F.S;
    if (p->nranks > 1) {
        CK(cudaMalloc(&all, (size_t)p->nranks * p->specElems * sizeof(float2)));
        NK(g_nccl.AllGather(F.S, all, p->specElems * 2, /*ncclFloat*/ 7, p->comm, p->stream));
        half = all;
    }
    CK(launch_spectrum_expand(half, p->viewBuf, p->sx, p->sy, p->sz, p->pitch, p->kyl, p->rank * p->zl, p->zl, p->kyCyclic ? p->nranks : 1, p->stream));
    CK(cudaMemcpyAsync(host, p->viewBuf, n * sizeof(float2), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    if (all) CK(cudaFree(all));
    return CUPSS_B200_OK;
}
int cupss_b200_download_comp(cupss_b200_plan* p, int f, float* host) {
    if (!p) return fail(CUPSS_B200_ERR_ARG, "null plan");
    return comm_guard(p, download_comp_impl(p, f, host));
}

int cupss_b200_step(cupss_b200_plan* p, int nsteps) {
    if (!p) return fail(CUPSS_B200_ERR_ARG, "null plan");
    CKR(comm_guard(p, CUPSS_B200_OK));   // a time-out of an earlier (asynchronous) step
    return comm_guard(p, p->do_steps(nsteps));
}
int cupss_b200_jit_selftest(char* log, int loglen) {
    // Compiles (does not load) the plan-specialised k stage of a synthetic Model-H-like sweep with NVRTC: keeps the kernel
    // headers compilable at run time.  Works without a GPU.  Returns 0 on success, ERR_STATE if NVRTC is not installed.
    jit_load();
    if (!g_jit.nvrtcOk) { if (log && loglen > 0) snprintf(log, loglen, "NVRTC or the kernel sources are not available"); return fail(CUPSS_B200_ERR_STATE, "NVRTC not available"); }
    KStageD ks{};
    ks.nsrc = 4; ks.nout = 3; ks.hasFwd = 1; ks.hasInv = 1;
    ks.out[0] = OutD{}; ks.out[0].termOff = 0; ks.out[0].nterm = 2; ks.out[0].impOff = 4; ks.out[0].nimp = 2; ks.out[0].dynamic = 1; ks.out[0].selfSrc = 0; ks.out[0].dst = 0; ks.out[0].inv = 1; ks.out[0].noisy = 1;
    ks.out[1] = OutD{}; ks.out[1].termOff = 2; ks.out[1].nterm = 1; ks.out[1].impOff = 6; ks.out[1].nimp = 1; ks.out[1].selfSrc = 1; ks.out[1].dst = 1;
    ks.out[2] = OutD{}; ks.out[2].termOff = 3; ks.out[2].nterm = 1; ks.out[2].impOff = 7; ks.out[2].nimp = 0; ks.out[2].selfSrc = 2; ks.out[2].dst = 2;
    ks.term[0] = TermD{0, 1, -1, 0}; ks.term[1] = TermD{1, 1, 3, 1}; ks.term[2] = TermD{2, 1, 0, 1}; ks.term[3] = TermD{3, 1, 1, 0};
    ks.pres[0].q2n = 1; ks.pres[1].iqx = 1; ks.pres[2].iqy = 1; ks.pres[2].invq = 1; ks.pres[3].iqx = 2; ks.pres[3].iqz = 1;
    ks.pres[4].q2n = 1; ks.pres[5].q2n = 2; ks.pres[6].q2n = 1;
    for (int L : {512, 2048, 64}) {
        std::vector<char> cubin;
        std::string why;
        if (!jit_compile(jit_source(ks, L), &cubin, &why)) {
            if (log && loglen > 0) snprintf(log, loglen, "L=%d: %s", L, why.c_str());
            return fail(CUPSS_B200_ERR_CUDA, "plan-specialised k stage does not compile for L=%d", L);
        }
    }
    // the lean evaluator of a noisy field (KPZ: one fused term with a constant prefactor, q^2 on the left-hand side)
    for (int L : {512, 1024, 32}) {
        std::vector<char> cubin;
        std::string why;
        if (!jit_compile(jit_source_lean(L, sq2_sig(1, 0, 0, 0, 1, 1, 0, 0, 0, 1) | SQ2_SIG_NOISE), &cubin, &why)) {
            if (log && loglen > 0) snprintf(log, loglen, "lean noisy k stage, L=%d: %s", L, why.c_str());
            return fail(CUPSS_B200_ERR_CUDA, "lean noisy k stage does not compile for L=%d", L);
        }
    }
    if (log && loglen > 0) snprintf(log, loglen, "ok");
    return CUPSS_B200_OK;
}

int cupss_b200_step_stage(cupss_b200_plan* p, int stage) {
    if (!p || stage < 0 || stage > 1) return fail(CUPSS_B200_ERR_ARG, "bad stage");
    CKR(comm_guard(p, CUPSS_B200_OK));
    return comm_guard(p, p->do_stage(stage));
}
int cupss_b200_real_view_begin(cupss_b200_plan* p, int f, int which, void** dev_float2) {
    CKR(check_field(p, f));
    if (!dev_float2 || which < 0 || which > 1) return fail(CUPSS_B200_ERR_ARG, "bad view request");
    float2* d = nullptr;
    CKR(p->view_begin(f, which, &d));
    *dev_float2 = d;
    return CUPSS_B200_OK;
}
int cupss_b200_real_view_commit(cupss_b200_plan* p, int f, int which) {
    CKR(check_field(p, f));
    return p->view_commit(f, which);
}
int cupss_b200_comp_view_begin(cupss_b200_plan* p, int f, void** dev_float2) {
    CKR(check_field(p, f));
    if (!dev_float2) return fail(CUPSS_B200_ERR_ARG, "null pointer");
    return p->comp_view_begin(f, reinterpret_cast<float2**>(dev_float2));
}
int cupss_b200_comp_view_commit(cupss_b200_plan* p, int f) {
    CKR(check_field(p, f));
    return p->comp_view_commit(f);
}
static int sync_impl(cupss_b200_plan* p) {
    CK(cudaStreamSynchronize(p->stream));
    return CUPSS_B200_OK;
}
int cupss_b200_sync(cupss_b200_plan* p) {
    if (!p) return fail(CUPSS_B200_ERR_ARG, "null plan");
    return comm_guard(p, sync_impl(p));
}

int cupss_b200_field_alias(cupss_b200_plan* p, int f, int* needs, int* order) {
    CKR(check_field(p, f));
    if (needs) *needs = p->fields[f].needsAlias;
    if (order) *order = p->fields[f].aliasOrder;
    return CUPSS_B200_OK;
}

int cupss_b200_time_steps(cupss_b200_plan* p, int nsteps, float* ms) {
    if (!p || !ms) return fail(CUPSS_B200_ERR_ARG, "null argument");
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaEventRecord(p->ev0, p->stream));
    CKR(p->do_steps(nsteps));
    CK(cudaEventRecord(p->ev1, p->stream));
    if (cudaEventSynchronize(p->ev1) != cudaSuccess) return comm_guard(p, fail(CUPSS_B200_ERR_CUDA, "the timed steps failed: %s", cudaGetErrorString(cudaGetLastError())));
    CK(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return comm_guard(p, CUPSS_B200_OK);
}

int cupss_b200_profile_step(cupss_b200_plan* p, int nmax, char* names, float* ms, double* bytes, int* n) {
    if (!p || !p->finalized) return fail(CUPSS_B200_ERR_STATE, "profile before finalize");
    // units: single launches, or whole pipelines (consecutive launches of one group, run on their lanes and timed fork to join)
    std::vector<std::pair<size_t, size_t>> units;
    for (size_t i = 0; i < p->step.size();) {
        size_t j = i + 1;
        if (p->step[i].group >= 0) while (j < p->step.size() && p->step[j].group == p->step[i].group) ++j;
        units.push_back({i, j});
        i = j;
    }
    const int cnt = (int)units.size();
    std::vector<cudaEvent_t> ev(cnt + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaEventRecord(ev[0], p->stream));
    for (int i = 0; i < cnt; ++i) {
        CKR(p->run_range(p->step, units[i].first, units[i].second));
        CK(cudaEventRecord(ev[i + 1], p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < cnt && i < nmax; ++i) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
        const Launch& l0 = p->step[units[i].first];
        if (ms) ms[i] = t;
        if (bytes) bytes[i] = l0.group >= 0 ? l0.groupBytes : l0.bytes;
        if (names) snprintf(names + 64 * i, 64, "%s", l0.group >= 0 ? l0.groupName : l0.name);
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (n) *n = cnt < nmax ? cnt : nmax;
    return CUPSS_B200_OK;
}

int cupss_b200_launches_per_step(cupss_b200_plan* p) {
    if (!p) return 0;
    int c = 0;
    for (auto& l : p->step) if (l.kind != Launch::A2A) ++c;
    return c;
}
double cupss_b200_bytes_per_step(cupss_b200_plan* p) {
    double b = 0;
    if (p) for (auto& l : p->step) if (l.kind != Launch::A2A) b += l.group >= 0 ? l.groupBytes : l.bytes;   // groupBytes: first launch of a pipeline only
    return b;
}
double cupss_b200_comm_bytes_per_step(cupss_b200_plan* p) {
    double b = 0;
    if (p) for (auto& l : p->step) b += l.kind == Launch::A2A ? l.bytes : l.commBytes;
    return b;
}
void* cupss_b200_device_spectrum(cupss_b200_plan* p, int f) {
    if (!p || f < 0 || f >= (int)p->fields.size()) return nullptr;
    return p->fields[f].S;
}

}  // extern "C"
