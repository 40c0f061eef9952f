// mccba_capi.cu -- the thin C-ABI shim declared in include/mccba.h: handle, host-side problem layout, stream /
// CUDA-graph orchestration and the NCCL hook.  All arithmetic lives in mccba_kernels.cuh / mccba_math.cuh.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <map>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/mccba.h"
#include "mccba_kernels.cuh"
#include "mccba_omni.cuh"
#include "mccba_stereo.cuh"
#include "mccba_ds.cuh"

using namespace mccba;

// ---------------------------------------------------------------------------------------------------------
// NCCL, resolved at run time (no link-time dependency; a single-GPU run never touches it)
// ---------------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
        }
    }
    return api;
}
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;
constexpr int kNcclMin = 3;
constexpr int kNcclInt8 = 0, kNcclInt32 = 2;
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------
struct mccba_handle_s {
    mccba_options opts;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // bulk observation upload, so that table uploads and memsets do not queue behind it
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr;
    // pinned staging arena for the layout tables: their uploads must not block the host behind the observation copy
    // peer-memory exchange (N > 1): own window + the windows of the peers opened through CUDA IPC
    double* p2p_win = nullptr;
    void* p2p_open[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long* p2p_epoch = nullptr;
    int64_t p2p_stride = 0;
    bool p2p_ok = false;
    unsigned long long p2p_budget_ns = 30000000000ull;   // spin budget of one exchange (MCCBA_P2P_TIMEOUT_MS), then MCCBA_ERR_NCCL
    char* pin_buf = nullptr;
    size_t pin_cap = 0, pin_used = 0, pin_want = 0;
    ncclComm_t comm = nullptr;
    int num_sms = 148;
    // cameras
    int n_cam = 0;
    std::vector<CamParams> cams;
    bool have_cams = false, have_obs = false, have_params = false;
    // problem (host copies of what is needed after set_observations)
    int n_frame = 0, n_edge = 0;
    int64_t n_pts = 0;
    std::vector<int> int_of_edge;     // reference edge index -> internal edge
    std::vector<int> edge_cam_h;      // reference order
    std::vector<int64_t> edge_n_h;    // corners per reference edge
    Problem P;
    std::vector<void*> allocs;        // device allocations owned by the problem
    std::vector<size_t> alloc_bytes;  // their sizes
    std::vector<std::pair<size_t, void*>> pool;   // allocations of the previous problem, reused when the size matches
    CamParams* d_cams = nullptr;
    int cur = 0;                      // host mirror of DevState::cur between calls
    int ar_len = 0;
    int k2_occ = 2;                   // minimum resident CTAs per SM requested from the Schur kernel (register cap)
    int obs_cap = 0;                  // floats per plane per TMA stage of the residual kernel (0: no staging)
    int prec = MCCBA_PRECISION_AUTO;  // requested precision policy of the residual / Jacobian pass (mccba_set_precision)
    int graph_prec = -1;              // effective policy the iteration graph was captured with
    double* d_ext_part = nullptr;     // AUTO: 64 partial minima of the angular extent
    int f32_grid = 0, f32_smem = 0;
    int k1_grid = 0, k1_smem = 0, k5_blocked = 2, iter_kernels = 5, dag_grid = 0;   // k5_blocked: 2 tile DAG, 3 block cyclic reduction
    int band_nw = 0;                  // 6 (block bandwidth + 1) of the reduced system (agreed over the ranks)
    cudaGraphExec_t graph = nullptr;
    int* h_done = nullptr;            // pinned
    DevState* h_state = nullptr;      // pinned
    double* h_small = nullptr;        // pinned scratch (64 doubles)
    double* d_small = nullptr;
    double prof_ms[6] = {0, 0, 0, 0, 0, 0};
    int profile = 0;
    double* x_saved = nullptr;        // device snapshot of the parameters (owned by the problem)
    // single-camera Mei calibration problem (mccba_omni_*)
    OmniProblem O;
    std::vector<void*> omni_allocs;
    bool omni_have = false, omni_have_params = false;
    cudaGraphExec_t omni_graph = nullptr;
    bool have_saved = false;
    // omnidir stereo bundle adjustment (mccba_stereo_*)
    StereoProblem S;
    std::vector<void*> stereo_allocs;
    bool stereo_have = false, stereo_have_params = false;
    cudaGraphExec_t stereo_graph = nullptr;
    // double-sided board calibration (mccba_ds_*): lives on top of the rig problem (cameras + observations)
    DsProblem D;
    std::vector<void*> ds_allocs;
    bool ds_have = false, ds_have_params = false;
};

namespace {

int fail(mccba_handle h, int code, const char* fmt, ...)
{
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

#define CUDA_TRY(h, expr)                                                                                  \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(h, MCCBA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int dev_alloc(mccba_handle h, T** p, size_t count, bool zero = false)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    for (size_t i = 0; i < h->pool.size(); ++i)   // same-shape problems (outlier loop, benchmark) never touch cudaMalloc
        if (h->pool[i].first == bytes) {
            q = h->pool[i].second;
            h->pool.erase(h->pool.begin() + (long)i);
            break;
        }
    if (!q) CUDA_TRY(h, cudaMalloc(&q, bytes));
    if (zero) CUDA_TRY(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    h->allocs.push_back(q);
    h->alloc_bytes.push_back(bytes);
    *p = (T*)q;
    return MCCBA_OK;
}
template <typename T>
int dev_upload(mccba_handle h, const T** p, const std::vector<T>& v)
{
    T* q = nullptr;
    int rc = dev_alloc(h, &q, v.size());
    if (rc) return rc;
    if (!v.empty()) {
        const size_t bytes = v.size() * sizeof(T), padded = (bytes + 255) & ~(size_t)255;
        h->pin_want += padded;
        const void* src = v.data();
        if (h->pin_used + padded <= h->pin_cap) {   // staged: the copy is asynchronous for the host
            memcpy(h->pin_buf + h->pin_used, v.data(), bytes);
            src = h->pin_buf + h->pin_used;
            h->pin_used += padded;
        }
        CUDA_TRY(h, cudaMemcpyAsync(q, src, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    *p = q;
    return MCCBA_OK;
}

// release the problem; keep_pool: park its device allocations for the next set_observations instead of freeing them
void free_problem(mccba_handle h, bool keep_pool = false)
{
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    for (size_t i = 0; i < h->allocs.size(); ++i) {
        if (keep_pool) h->pool.push_back({h->alloc_bytes[i], h->allocs[i]});
        else cudaFree(h->allocs[i]);
    }
    h->allocs.clear();
    h->alloc_bytes.clear();
    h->have_obs = false;
    h->have_params = false;
    h->have_saved = false;
    h->x_saved = nullptr;
}
void drain_pool(mccba_handle h)
{
    for (auto& pr : h->pool) cudaFree(pr.second);
    h->pool.clear();
}

int64_t n_param(mccba_handle h) { return 6 * (int64_t)(h->n_cam + h->n_frame - 1); }

void launch_resid(mccba_handle h, cudaStream_t s, int forced)
{
    if (h->P.prec == MCCBA_PRECISION_MIXED) { resid_jac_accum_f32_kernel<true><<<h->f32_grid, kF32Threads, h->f32_smem, s>>>(h->P, forced); return; }
    if (h->P.prec == MCCBA_PRECISION_FAST32) { resid_jac_accum_f32_kernel<false><<<h->f32_grid, kF32Threads, h->f32_smem, s>>>(h->P, forced); return; }
    if (h->obs_cap > 0) resid_jac_accum_kernel<true><<<h->k1_grid, kK1Threads, h->k1_smem, s>>>(h->P, forced, h->obs_cap);
    else resid_jac_accum_kernel<false><<<h->k1_grid, kK1Threads, h->k1_smem, s>>>(h->P, forced, 0);
}

void launch_schur(mccba_handle h, cudaStream_t s, int sel, double lambda)
{
    const Problem& P = h->P;
    const int grid = (P.n_warps * 32 + kK2Threads - 1) / kK2Threads;
    switch (h->k2_occ) {
        case 2: frame_schur_kernel<2><<<grid, kK2Threads, kK2SmemBytes, s>>>(P, sel, lambda); break;
        case 3: frame_schur_kernel<3><<<grid, kK2Threads, kK2SmemBytes, s>>>(P, sel, lambda); break;
        case 4: frame_schur_kernel<4><<<grid, kK2Threads, kK2SmemBytes, s>>>(P, sel, lambda); break;
        default: frame_schur_kernel<1><<<grid, kK2Threads, kK2SmemBytes, s>>>(P, sel, lambda); break;
    }
}

// banded reduced system (mode 3): block cyclic reduction with super-blocks of B = 6 m rows, m = block bandwidth;
// nw = 6 (m + 1) is the width of the packed band
template <int B>
static cudaError_t launch_bcr_b(const double* A, int n, double* xout, int* fail, const int* go, const Problem& P, int fused, int packed, cudaStream_t s)
{
    const size_t smem = bcr_smem_bytes(n, B);
    cudaError_t e = cudaFuncSetAttribute(chol_bcr_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    chol_bcr_kernel<B><<<1, BcrCfg<B>::kThreads, smem, s>>>(A, n, xout, fail, go, P, fused, packed);
    return cudaGetLastError();
}
static cudaError_t launch_band(int nw, const double* A, int n, double* xout, int* fail, const int* go, const Problem& P, int fused, int packed, cudaStream_t s)
{
    switch (nw) {
        case 12: return launch_bcr_b<6>(A, n, xout, fail, go, P, fused, packed, s);
        case 18: return launch_bcr_b<12>(A, n, xout, fail, go, P, fused, packed, s);
        case 24: return launch_bcr_b<18>(A, n, xout, fail, go, P, fused, packed, s);
        case 30: return launch_bcr_b<24>(A, n, xout, fail, go, P, fused, packed, s);
        default: return cudaErrorInvalidValue;
    }
}
static bool band_fits(int nw, int n) { return nw >= 12 && nw <= 30 && nw % 6 == 0 && bcr_smem_bytes(n, nw - 6) <= 220 * 1024; }

// enqueue one iteration: [memset S] K2 K3a [allreduce] K5 K4 K1
int enqueue_iteration(mccba_handle h, bool timed)
{
    Problem& P = h->P;
    cudaStream_t s = h->stream;
    cudaEvent_t ev[7];
    if (timed)
        for (auto& e : ev) cudaEventCreate(&e);
    if (timed) cudaEventRecord(ev[0], s);
    // reduce_records rewrites every block that has a source and nothing else writes into its buffer (the collective
    // is out of place, neither solver factors in place), so blocks without sources stay zero from allocation
    launch_schur(h, s, -1, 0.0);
    if (timed) cudaEventRecord(ev[1], s);
    reduce_records_kernel<<<P.n_dest, kK3Threads, 0, s>>>(P, 0);
    if (timed) cudaEventRecord(ev[2], s);
    static const bool no_exchange = getenv("MCCBA_NO_EXCHANGE") != nullptr;   // diagnosis only: wrong results for N > 1
    if (h->opts.nranks > 1 && no_exchange) {
        CUDA_TRY(h, cudaMemcpyAsync(P.ar, P.ar_part, sizeof(double) * (size_t)h->ar_len, cudaMemcpyDeviceToDevice, s));
    } else if (h->opts.nranks > 1 && h->p2p_ok) {
        const int grid = std::max(1, std::min(h->num_sms, (h->ar_len + kP2pThreads - 1) / kP2pThreads));   // one element per thread
        p2p_exchange_kernel<<<grid, kP2pThreads, 0, s>>>(P, (int64_t)h->ar_len, h->p2p_budget_ns);
    } else if (h->opts.nranks > 1) {
        ncclResult_t r = nccl().AllReduce(P.ar_part, P.ar, (size_t)h->ar_len, kNcclFloat64, kNcclSum, h->comm, s);
        if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    }
    if (timed) cudaEventRecord(ev[3], s);
    const bool band = h->k5_blocked == 3 && P.ns > 0;
    if (!band) decide_kernel<<<1, 32, 0, s>>>(P);   // banded mode: the single-CTA solver does the loop control itself
    if (band) {
        CUDA_TRY(h, launch_band(h->band_nw, P.ar, P.ns, P.dc, &P.st->chol_fail, &P.st->go, P, 2, 1, s));
    } else if (h->k5_blocked == 2 && P.ns > 0) {
        CholDag D{P.ar, P.dag_buf, P.ns, P.dc, &P.st->go, &P.st->chol_fail, nullptr};   // sentinel fill: reduce_records
        // cooperative launch: the CTAs of the DAG spin on each other, so the runtime must place all of them at once
        // (or fail) -- a second context or stream sharing the GPU can then delay the solve but never deadlock it
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->dag_grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CUDA_TRY(h, cudaLaunchKernelEx(&cfg, chol_dag_kernel, D, P, P.ns <= 512 ? 1 : 0));
    }
    if (!((h->k5_blocked == 3 && P.ns > 0) || (h->k5_blocked == 2 && P.ns > 0 && P.ns <= 512)))   // otherwise fused into the solve
        camera_update_kernel<<<1, kK5Threads, 0, s>>>(P);
    if (timed) cudaEventRecord(ev[4], s);
    frame_update_kernel<<<P.n_k4_blocks, kK4Threads, kK4SmemBytes, s>>>(P);
    if (timed) cudaEventRecord(ev[5], s);
    launch_resid(h, s, 0);
    if (timed) {
        cudaEventRecord(ev[6], s);
        cudaEventSynchronize(ev[6]);
        for (int i = 0; i < 6; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            h->prof_ms[i] += ms;
        }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}

// AUTO precision policy: MIXED iff every live edge sees its board under an angular extent >= kAutoMinExtent at the composed
// poses just written to Problem::erec (profiles/r2_precision_vs_board.txt); else FP64.  One small kernel + an 0.5 KB readback.
constexpr double kAutoMinExtent = 0.15;
int resolve_precision(mccba_handle h)
{
    if (h->prec != MCCBA_PRECISION_AUTO) return MCCBA_OK;
    Problem& P = h->P;
    double part[64];
    min_angular_extent_kernel<<<64, 256, 0, h->stream>>>(P, h->d_ext_part);
    CUDA_TRY(h, cudaMemcpyAsync(part, h->d_ext_part, sizeof(part), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double mn = 1e300;
    for (double v : part) mn = std::min(mn, v);
    P.prec = (mn >= kAutoMinExtent) ? MCCBA_PRECISION_MIXED : MCCBA_PRECISION_FP64;   // NaN poses: FP64
    if (h->graph && h->graph_prec != P.prec) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    return MCCBA_OK;
}

int launch_forced_eval(mccba_handle h)
{
    Problem& P = h->P;
    vertex_prep_kernel<<<(P.n_vertex + 127) / 128, 128, 0, h->stream>>>(P, -1);
    edge_pose_kernel<<<(P.n_slots + 127) / 128, 128, 0, h->stream>>>(P, -1);
    int rcp = resolve_precision(h);
    if (rcp) return rcp;
    launch_resid(h, h->stream, 1);
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}

// make DevState::cur on the device agree with the host mirror (parameters live in x[cur])
int sync_state_cur(mccba_handle h)
{
    init_state_kernel<<<1, 1, 0, h->stream>>>(h->P.st, 0, 1, 0, 0.0, 0.0, 1.0, 1.0, h->cur);
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}

}  // namespace

extern "C" {

int mccba_default_options(mccba_options* o)
{
    if (!o) return MCCBA_ERR_ARG;
    memset(o, 0, sizeof(*o));
    o->device = 0; o->rank = 0; o->nranks = 1; o->use_graph = 1; o->verbose = 0;
    return MCCBA_OK;
}

int mccba_default_solve_opts(mccba_solve_opts* o)
{
    if (!o) return MCCBA_ERR_ARG;
    o->mode = MCCBA_MODE_REFERENCE_GN;
    o->crit_type = MCCBA_CRIT_COUNT;  // TermCriteria(COUNT, 20, 1e-7), multicalib.hpp:140
    o->max_count = 20;
    o->epsilon = 1e-7;
    o->lambda0 = 1e-3; o->lambda_up = 10.0; o->lambda_down = 1.0 / 3.0;
    return MCCBA_OK;
}

int mccba_set_precision(mccba_handle h, int policy)
{
    if (!h) return MCCBA_ERR_ARG;
    if (policy < MCCBA_PRECISION_FP64 || policy > MCCBA_PRECISION_AUTO) return fail(h, MCCBA_ERR_ARG, "set_precision: unknown policy %d", policy);
    if (policy != h->prec && h->have_obs) {   // the observation layout depends on the policy: the problem has to be set again
        cudaSetDevice(h->opts.device);
        cudaStreamSynchronize(h->stream);
        free_problem(h);
    }
    h->prec = policy;
    return MCCBA_OK;
}

int mccba_get_precision(mccba_handle h) { return h ? h->prec : -1; }
int mccba_effective_precision(mccba_handle h)
{
    if (!h) return -1;
    if (h->prec != MCCBA_PRECISION_AUTO) return h->prec;
    return h->have_obs ? h->P.prec : MCCBA_PRECISION_MIXED;
}

int mccba_nccl_unique_id(unsigned char out[128])
{
    if (!nccl().ok) return MCCBA_ERR_NCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != 0) return MCCBA_ERR_NCCL;
    memcpy(out, id.internal, 128);
    return MCCBA_OK;
}

int mccba_create(const mccba_options* opts, mccba_handle* out)
{
    if (!opts || !out) return MCCBA_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MCCBA_ERR_CUDA;  // no CPU fallback
    if (opts->device < 0 || opts->device >= ndev || opts->nranks < 1 || opts->rank < 0 || opts->rank >= opts->nranks)
        return MCCBA_ERR_ARG;
    mccba_handle h = new mccba_handle_s();
    h->opts = *opts;
    memset(&h->P, 0, sizeof(h->P));
    memset(&h->O, 0, sizeof(h->O));
    memset(&h->S, 0, sizeof(h->S));
    if (cudaSetDevice(opts->device) != cudaSuccess) { delete h; return MCCBA_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, opts->device);
    h->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MCCBA_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return MCCBA_ERR_CUDA; }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming);
    cudaMallocHost((void**)&h->h_done, 64 * sizeof(int));
    cudaMallocHost((void**)&h->h_state, sizeof(DevState));
    cudaMallocHost((void**)&h->h_small, 64 * sizeof(double));
    cudaMalloc((void**)&h->d_small, 64 * sizeof(double));
    if (const char* occ = getenv("MCCBA_K2_OCC")) h->k2_occ = atoi(occ);
    if (const char* pr = getenv("MCCBA_PRECISION")) {   // auto | fp64 | mixed | fast32 (or 3 | 0 | 1 | 2)
        const char c0 = pr[0], c1 = pr[0] ? pr[1] : 0;
        h->prec = (c0 == '0' || c1 == 'p' || c1 == 'P') ? MCCBA_PRECISION_FP64
                  : (c0 == '2' || ((c0 == 'f' || c0 == 'F') && (c1 == 'a' || c1 == 'A'))) ? MCCBA_PRECISION_FAST32
                  : (c0 == '1' || c0 == 'm' || c0 == 'M') ? MCCBA_PRECISION_MIXED : MCCBA_PRECISION_AUTO;
    }
    const char* prof = getenv("MCCBA_PROFILE");
    h->profile = prof && prof[0] == '1';
    if (opts->nranks > 1) {
        if (!nccl().ok) { *out = h; return fail(h, MCCBA_ERR_NCCL, "libnccl.so.2 could not be loaded"); }
        ncclUniqueId id;
        memcpy(id.internal, opts->nccl_id, 128);
        ncclResult_t r = nccl().CommInitRank(&h->comm, opts->nranks, id, opts->rank);
        if (r != 0) { *out = h; return fail(h, MCCBA_ERR_NCCL, "ncclCommInitRank failed: %d", r); }
    }
    *out = h;
    return MCCBA_OK;
}

static void p2p_teardown(mccba_handle h);
int mccba_destroy(mccba_handle h)
{
    if (!h) return MCCBA_OK;
    cudaSetDevice(h->opts.device);
    cudaStreamSynchronize(h->stream);
    free_problem(h);
    drain_pool(h);
    if (h->omni_graph) cudaGraphExecDestroy(h->omni_graph);
    for (void* q : h->omni_allocs) cudaFree(q);
    if (h->stereo_graph) cudaGraphExecDestroy(h->stereo_graph);
    for (void* q : h->stereo_allocs) cudaFree(q);
    for (void* q : h->ds_allocs) cudaFree(q);
    if (h->d_cams) cudaFree(h->d_cams);
    p2p_teardown(h);
    if (h->comm) nccl().CommDestroy(h->comm);
    cudaFreeHost(h->h_done);
    cudaFreeHost(h->h_state);
    cudaFreeHost(h->h_small);
    cudaFree(h->d_small);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->pin_buf) cudaFreeHost(h->pin_buf);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    delete h;
    return MCCBA_OK;
}

const char* mccba_last_error(mccba_handle h) { return h ? h->err.c_str() : "null handle"; }

int mccba_set_cameras(mccba_handle h, int n_cam, const int* model, const double* K5, const double* dist8,
                      const int* ndist, const double* xi)
{
    if (!h) return MCCBA_ERR_ARG;
    if (n_cam < 1 || !model || !K5 || !dist8 || !ndist || !xi) return fail(h, MCCBA_ERR_ARG, "set_cameras: null or empty input");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    std::vector<CamParams> cams((size_t)n_cam);
    for (int c = 0; c < n_cam; ++c) {
        CamParams& p = cams[c];
        memset(&p, 0, sizeof(p));
        if (model[c] != MCCBA_PINHOLE && model[c] != MCCBA_OMNIDIRECTIONAL)
            return fail(h, MCCBA_ERR_ARG, "camera %d: unknown model %d", c, model[c]);
        const int nd = ndist[c];
        if (model[c] == MCCBA_OMNIDIRECTIONAL && nd != 4)
            return fail(h, MCCBA_ERR_ARG, "camera %d: the Mei model needs exactly 4 distortion coefficients (src/omnidir.cpp:92)", c);
        if (model[c] == MCCBA_PINHOLE && nd != 0 && nd != 4 && nd != 5 && nd != 8)
            return fail(h, MCCBA_ERR_ARG, "camera %d: %d distortion coefficients unsupported (0, 4, 5 or 8)", c, nd);
        p.model = model[c];
        p.fx = K5[5 * c]; p.fy = K5[5 * c + 1]; p.cx = K5[5 * c + 2]; p.cy = K5[5 * c + 3]; p.skew = K5[5 * c + 4];
        p.xi = xi[c];
        double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < nd; ++i) k[i] = dist8[8 * c + i];
        p.k1 = k[0]; p.k2 = k[1]; p.p1 = k[2]; p.p2 = k[3]; p.k3 = k[4]; p.k4 = k[5]; p.k5 = k[6]; p.k6 = k[7];
        p.rational = (p.model == MCCBA_PINHOLE) && (k[5] != 0 || k[6] != 0 || k[7] != 0);
    }
    if (h->have_obs && n_cam != h->n_cam) free_problem(h);
    if (h->d_cams) { cudaFree(h->d_cams); h->d_cams = nullptr; }
    CUDA_TRY(h, cudaMalloc((void**)&h->d_cams, sizeof(CamParams) * (size_t)n_cam));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_cams, cams.data(), sizeof(CamParams) * (size_t)n_cam, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->cams.swap(cams);
    h->n_cam = n_cam;
    h->have_cams = true;
    if (h->have_obs) h->P.cams = h->d_cams;
    return MCCBA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// peer-memory windows for the per-iteration exchange (N > 1).  Collective: every rank calls it from the same
// set_observations.  Any failure on any rank (no IPC, no peer access, more than 8 ranks, MCCBA_P2P=0) makes ALL
// ranks fall back to ncclAllReduce -- the decision itself is all-reduced.
// ---------------------------------------------------------------------------------------------------------
static void p2p_teardown(mccba_handle h)
{
    for (int r = 0; r < 8; ++r) {
        if (h->p2p_open[r]) cudaIpcCloseMemHandle(h->p2p_open[r]);
        h->p2p_open[r] = nullptr;
        h->p2p_peer[r] = nullptr;
    }
    if (h->p2p_win) cudaFree(h->p2p_win);
    if (h->p2p_epoch) cudaFree(h->p2p_epoch);
    h->p2p_win = nullptr;
    h->p2p_epoch = nullptr;
    h->p2p_ok = false;
    h->p2p_stride = 0;
    cudaGetLastError();
}

// all-rank barrier on the library's own communicator (tiny all-reduce + host wait)
static int p2p_barrier(mccba_handle h)
{
    if (!h->comm || !nccl().AllReduce) return MCCBA_OK;
    ncclResult_t r = nccl().AllReduce(h->d_small, h->d_small, 1, kNcclInt32, kNcclMin, h->comm, h->stream);
    if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce (barrier) failed: %d", r);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

static int p2p_setup(mccba_handle h)
{
    const int n = h->opts.nranks, me = h->opts.rank;
    const int64_t stride = (2 * (int64_t)h->ar_len + 31) & ~(int64_t)31;   // LL protocol: two 8-byte words per element
    if (h->p2p_stride == stride) return MCCBA_OK;   // same reduced-system size as the previous problem: nothing to do
    if (h->p2p_win) {   // a peer may still have the old window mapped: nobody frees before everybody is here
        int rcb = p2p_barrier(h);
        if (rcb) return rcb;
    }
    p2p_teardown(h);
    h->p2p_stride = stride;                         // remembered even on fallback, so the decision is taken once per size
    const char* env = getenv("MCCBA_P2P");
    if (const char* tmo = getenv("MCCBA_P2P_TIMEOUT_MS")) {
        const long long ms = atoll(tmo);
        if (ms > 0) h->p2p_budget_ns = (unsigned long long)ms * 1000000ull;
    }
    struct Msg { cudaIpcMemHandle_t hdl; int ok; int pad[15]; };
    static_assert(sizeof(Msg) == 128, "message layout");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    // default when every rank can map every window; MCCBA_P2P=0 (on every rank) forces ncclAllReduce
    mine.ok = (n <= 8 && nccl().AllGather && !(env && env[0] == '0')) ? 1 : 0;
    const size_t words = (size_t)kP2pFlagWords + 2 * (size_t)n * (size_t)stride;
    if (mine.ok && cudaMalloc((void**)&h->p2p_win, words * sizeof(double)) != cudaSuccess) { mine.ok = 0; h->p2p_win = nullptr; cudaGetLastError(); }
    if (mine.ok && cudaMemsetAsync(h->p2p_win, 0, words * sizeof(double), h->stream) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    if (mine.ok && cudaIpcGetMemHandle(&mine.hdl, h->p2p_win) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    if (!nccl().AllGather) { p2p_teardown(h); h->p2p_stride = stride; return MCCBA_OK; }   // same on every rank
    // From here on every rank must reach both collectives whatever happens locally: a local failure only clears `ok`.
    char* d_msg = nullptr;
    if (cudaMalloc((void**)&d_msg, sizeof(Msg) * (size_t)(n + 1) + 256) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, MCCBA_ERR_CUDA, "p2p_setup: cudaMalloc of the %d-byte handshake buffer failed", (int)(sizeof(Msg) * (size_t)(n + 1) + 256));
    }
    std::vector<Msg> all((size_t)n);
    bool local_ok = cudaMemcpyAsync(d_msg, &mine, sizeof(Msg), cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
    ncclResult_t r = nccl().AllGather(d_msg, d_msg + sizeof(Msg), sizeof(Msg), kNcclInt8, h->comm, h->stream);
    if (r != 0) { cudaFree(d_msg); return fail(h, MCCBA_ERR_NCCL, "ncclAllGather failed: %d", r); }
    local_ok = local_ok && cudaMemcpyAsync(all.data(), d_msg + sizeof(Msg), sizeof(Msg) * (size_t)n, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
    local_ok = local_ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    int ok = local_ok ? 1 : 0;
    for (int q = 0; q < n && ok; ++q) ok = ok && all[(size_t)q].ok;
    if (ok) {
        for (int q = 0; q < n && ok; ++q) {
            if (q == me) { h->p2p_peer[q] = h->p2p_win; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[(size_t)q].hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
            h->p2p_open[q] = ptr;
            h->p2p_peer[q] = (double*)ptr;
        }
        if (ok && cudaMalloc((void**)&h->p2p_epoch, 256) != cudaSuccess) { ok = 0; h->p2p_epoch = nullptr; cudaGetLastError(); }
        if (ok && cudaMemsetAsync(h->p2p_epoch, 0, 256, h->stream) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
    // second round: did every rank manage to open every window?  (also the barrier behind the window memsets)
    int* d_ok = (int*)(d_msg + sizeof(Msg) * (size_t)(n + 1));
    if (cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        cudaMemsetAsync(d_ok, 0, sizeof(int), h->stream);   // the collective still runs, with "not ok"
    }
    r = nccl().AllReduce(d_ok, d_ok, 1, kNcclInt32, kNcclMin, h->comm, h->stream);
    if (r != 0) { cudaFree(d_msg); return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %d", r); }
    int all_ok = 0;
    if (cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) { cudaGetLastError(); all_ok = 0; }
    cudaFree(d_msg);
    if (!all_ok || !ok) { p2p_teardown(h); h->p2p_stride = stride; }
    else h->p2p_ok = true;
    if (h->opts.verbose) fprintf(stderr, "[mccba] rank %d: reduced-system exchange over %s\n", me, h->p2p_ok ? "NVLink peer memory" : "ncclAllReduce");
    return MCCBA_OK;
}

int mccba_set_observations(mccba_handle h, int n_frame, int n_edge, const int* edge_cam, const int* edge_pv,
                           const int64_t* edge_off, const float* obj_xyz, const float* img_uv)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_cams) return fail(h, MCCBA_ERR_STATE, "set_observations before set_cameras");
    if (n_frame < 1 || n_edge < 1 || !edge_cam || !edge_pv || !edge_off || !obj_xyz || !img_uv)
        return fail(h, MCCBA_ERR_ARG, "set_observations: null or empty input");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    h->ds_have = h->ds_have_params = false;     // a double-sided problem is tied to the edge order of the observations it was set on
    const char* timing_env = getenv("MCCBA_TIMING");
    const bool timing = timing_env != nullptr, timing_sync = timing && timing_env[0] != '2';   // "2": host laps only, no stream sync
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        if (timing_sync) cudaStreamSynchronize(h->stream);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[set_observations] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
        t_start = now;
    };
    const int nC = h->n_cam;
    if (edge_off[0] != 0) return fail(h, MCCBA_ERR_ARG, "edge_off[0] must be 0");
    const int64_t M = edge_off[n_edge];
    if (M <= 0 || M >= (int64_t)2000000000) return fail(h, MCCBA_ERR_ARG, "corner count %lld out of range", (long long)M);
    // the previous problem's buffers go to the pool; the big host->device copies start right away and overlap with
    // the host-side layout work below
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    free_problem(h, true);
    if (h->pin_want > h->pin_cap) {   // the previous call did not fit: grow the staging arena (nothing is in flight here)
        if (h->pin_buf) cudaFreeHost(h->pin_buf);
        h->pin_buf = nullptr;
        h->pin_cap = 0;
        const size_t want = h->pin_want + h->pin_want / 4 + 4096;
        if (cudaHostAlloc((void**)&h->pin_buf, want, cudaHostAllocDefault) == cudaSuccess) h->pin_cap = want;
        else cudaGetLastError();   // no arena: plain pageable uploads
    }
    h->pin_used = 0;
    h->pin_want = 0;
    float *d_obj = nullptr, *d_img = nullptr;
    {
        int rc0;
        if ((rc0 = dev_alloc(h, &d_obj, 3 * (size_t)M))) return rc0;
        if ((rc0 = dev_alloc(h, &d_img, 2 * (size_t)M))) return rc0;
    }
    CUDA_TRY(h, cudaMemcpyAsync(d_obj, obj_xyz, sizeof(float) * 3 * (size_t)M, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_img, img_uv, sizeof(float) * 2 * (size_t)M, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_copy, h->copy_stream));
    struct CopyGuard {   // an early error return must not leave the engine reading the caller's buffers
        cudaStream_t s;
        ~CopyGuard() { cudaStreamSynchronize(s); }
    } copy_guard{h->copy_stream};
    // per-frame view lists (CSR by counting sort), each sorted by camera
    std::vector<int> voff((size_t)n_frame + 1, 0);
    for (int e = 0; e < n_edge; ++e) {
        const int c = edge_cam[e], f = edge_pv[e] - nC;
        if (c < 0 || c >= nC) return fail(h, MCCBA_ERR_ARG, "edge %d: cameraVertex %d out of range", e, c);
        if (f < 0 || f >= n_frame) return fail(h, MCCBA_ERR_ARG, "edge %d: photoVertex %d out of range [%d,%d)", e, edge_pv[e], nC, nC + n_frame);
        if (edge_off[e + 1] < edge_off[e]) return fail(h, MCCBA_ERR_ARG, "edge_off not monotone at edge %d", e);
        voff[f + 1]++;
    }
    for (int f = 0; f < n_frame; ++f) {
        if (voff[f + 1] == 0) return fail(h, MCCBA_ERR_ARG, "photo vertex %d has no observation", nC + f);
        voff[f + 1] += voff[f];
    }
    std::vector<std::pair<int, int>> vlist((size_t)n_edge);   // (camera, edge) per frame, CSR
    {
        std::vector<int> fill(voff.begin(), voff.end() - 1);
        for (int e = 0; e < n_edge; ++e) vlist[fill[edge_pv[e] - nC]++] = {edge_cam[e], e};
    }
    std::map<std::vector<int>, std::vector<int>> groups;  // camera set -> frames (ascending)
    {
        // frames hashed by their (sorted) camera list: a flat open-addressing table keyed by the list packed into 64
        // bits (up to 4 views), with the previous frame's group tried first; one std::map insertion per distinct set
        struct Slot { uint64_t key; std::vector<int>* grp; };
        std::vector<Slot> table(4096, Slot{0, nullptr});
        size_t used = 0;
        auto probe = [&](uint64_t k) -> Slot& {
            size_t m = table.size() - 1, i = (size_t)((k * 0x9E3779B97F4A7C15ull) >> 20) & m;
            while (table[i].grp && table[i].key != k) i = (i + 1) & m;
            return table[i];
        };
        std::vector<int> key;
        uint64_t last_k = 0;
        std::vector<int>* last_grp = nullptr;
        for (int f = 0; f < n_frame; ++f) {
            std::pair<int, int>* b = vlist.data() + voff[f];
            const int nv = voff[f + 1] - voff[f];
            if (nv == 2) { if (b[1] < b[0]) std::swap(b[0], b[1]); }
            else if (nv > 2) std::sort(b, b + nv);
            for (int i = 1; i < nv; ++i)
                if (b[i].first == b[i - 1].first)
                    return fail(h, MCCBA_ERR_ARG, "photo vertex %d is observed twice by camera %d", nC + f, b[i].first);
            if (nv <= 4 && nC < 65535) {
                uint64_t k = 0;
                for (int i = 0; i < nv; ++i) k = (k << 16) | (uint64_t)(b[i].first + 1);
                if (k != last_k || !last_grp) {
                    Slot* sl = &probe(k);
                    if (!sl->grp) {
                        if (2 * (used + 1) > table.size()) {   // grow and rehash
                            std::vector<Slot> old;
                            old.swap(table);
                            table.assign(old.size() * 4, Slot{0, nullptr});
                            for (const Slot& o : old)
                                if (o.grp) probe(o.key) = o;
                            sl = &probe(k);
                        }
                        key.clear();
                        for (int i = 0; i < nv; ++i) key.push_back(b[i].first);
                        sl->key = k;
                        sl->grp = &groups[key];
                        sl->grp->reserve(64);
                        ++used;
                    }
                    last_k = k;
                    last_grp = sl->grp;
                }
                last_grp->push_back(f);
            } else {
                key.clear();
                for (int i = 0; i < nv; ++i) key.push_back(b[i].first);
                groups[key].push_back(f);
            }
        }
    }
    lap("validate + group frames");
    Problem& P = h->P;
    memset(&P, 0, sizeof(P));
    P.n_cam = nC; P.n_frame = n_frame; P.n_vertex = nC + n_frame; P.ns = 6 * (nC - 1);
    P.n_param = 6 * (int64_t)(nC + n_frame - 1);
    P.cams = h->d_cams;

    std::vector<int> slot_frame, warp_group, group_V, group_cam0, group_cams, group_ebase, group_stride, group_slot0, warp_rec;
    std::vector<int> wmeta;   // 8 ints per warp, see Problem::wmeta
    std::vector<int> e_cam, e_frame, e_off;
    std::vector<int64_t> e_src;
    h->int_of_edge.assign((size_t)n_edge, -1);
    std::map<std::pair<int, int>, std::vector<int>> dest_blocks;  // (A,B) -> record offsets
    std::map<int, std::vector<int>> dest_g;
    int64_t rec_total = 0;
    {
        size_t n_int = 0, n_sl = 0;
        for (auto& kv : groups) {
            const size_t st = (kv.second.size() + 31) / 32 * 32;
            n_int += st * kv.first.size();
            n_sl += st;
        }
        if (n_int >= (size_t)2000000000) return fail(h, MCCBA_ERR_ARG, "too many edges");
        e_cam.resize(n_int); e_frame.resize(n_int); e_src.resize(n_int); e_off.resize(n_int + 1);
        slot_frame.resize(n_sl);
    }
    e_off[0] = 0;
    int gi = 0;
    size_t ne = 0, nsl = 0;
    for (auto& kv : groups) {
        const std::vector<int>& cams = kv.first;
        const std::vector<int>& frames = kv.second;
        const int V = (int)cams.size(), nf = (int)frames.size();
        const int nw = (nf + 31) / 32, stride = nw * 32;
        group_V.push_back(V);
        group_cam0.push_back((int)group_cams.size());
        for (int c : cams) group_cams.push_back(c);
        group_ebase.push_back((int)ne);
        group_stride.push_back(stride);
        group_slot0.push_back((int)nsl);
        for (int ls = 0; ls < stride; ++ls) slot_frame[nsl + ls] = ls < nf ? frames[ls] : -1;
        nsl += stride;
        for (int v = 0; v < V; ++v) {
            const int cam = cams[v];
            int off = e_off[ne];
            for (int ls = 0; ls < nf; ++ls, ++ne) {
                const int frame = frames[ls];
                const int oe = vlist[(size_t)voff[frame] + v].second;
                const int64_t src = edge_off[oe];
                h->int_of_edge[oe] = (int)ne;
                e_cam[ne] = cam; e_frame[ne] = frame; e_src[ne] = src;
                off += (int)(edge_off[oe + 1] - src);
                e_off[ne + 1] = off;
            }
            for (int ls = nf; ls < stride; ++ls, ++ne) { e_cam[ne] = cam; e_frame[ne] = -1; e_src[ne] = 0; e_off[ne + 1] = off; }
        }
        std::vector<int> act;
        for (int c : cams)
            if (c != 0) act.push_back(c);
        const int Va = (int)act.size();
        const int rec_len = 2 + 42 * Va + 36 * (Va * (Va - 1) / 2);
        for (int w = 0; w < nw; ++w) {
            warp_group.push_back(gi);
            if (rec_total + rec_len >= (int64_t)2000000000) return fail(h, MCCBA_ERR_ARG, "record buffer too large");
            const int ro = (int)rec_total;
            warp_rec.push_back(ro);
            const int meta[8] = {V, group_cam0.back(), group_ebase.back(), stride, 32 * w, ro, gi, 0};
            wmeta.insert(wmeta.end(), meta, meta + 8);
            for (int i = 0; i < Va; ++i) {
                dest_blocks[{act[i] - 1, act[i] - 1}].push_back(ro + 2 + 36 * i);
                dest_g[act[i] - 1].push_back(ro + 2 + 36 * Va + 6 * i);
            }
            int pi = 0;
            for (int i = 0; i < Va; ++i)
                for (int j = i + 1; j < Va; ++j) dest_blocks[{act[i] - 1, act[j] - 1}].push_back(ro + 2 + 42 * Va + 36 * pi++);
            rec_total += rec_len;
        }
        ++gi;
    }
    {   // block bandwidth of the reduced system under the camera numbering (over all ranks)
        int m = 0;
        for (auto& kv : dest_blocks) m = std::max(m, std::abs(kv.first.first - kv.first.second));
        if (h->opts.nranks > 1) {
            CUDA_TRY(h, cudaMemcpyAsync(h->d_small, &m, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            ncclResult_t r = nccl().AllReduce(h->d_small, h->d_small, 1, kNcclInt32, 2 /* ncclMax */, h->comm, h->stream);
            if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %d", r);
            CUDA_TRY(h, cudaMemcpyAsync(&m, h->d_small, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        }
        h->band_nw = 6 * (std::max(m, 1) + 1);   // at least one off-diagonal block: the solver wants super-blocks of >= 6 rows
    }
    std::vector<int> dest_info, dest_src0, dest_src;
    dest_src0.push_back(0);
    for (auto& kv : dest_blocks) {
        dest_info.insert(dest_info.end(), {0, kv.first.first, kv.first.second, 0});
        dest_src.insert(dest_src.end(), kv.second.begin(), kv.second.end());
        dest_src0.push_back((int)dest_src.size());
    }
    for (auto& kv : dest_g) {
        dest_info.insert(dest_info.end(), {1, kv.first, 0, 0});
        dest_src.insert(dest_src.end(), kv.second.begin(), kv.second.end());
        dest_src0.push_back((int)dest_src.size());
    }
    dest_info.insert(dest_info.end(), {2, 0, 0, 0});
    dest_src0.push_back((int)dest_src.size());

    lap("build layout tables");
    P.n_edge_int = (int)e_cam.size();
    P.n_slots = (int)slot_frame.size();
    P.n_warps = P.n_slots / 32;
    P.n_dest = (int)dest_info.size() / 4;
    P.n_k4_blocks = (P.n_slots + kK4Threads - 1) / kK4Threads;
    CUDA_TRY(h, cudaFuncSetAttribute(frame_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kK4SmemBytes));
    CUDA_TRY(h, cudaFuncSetAttribute(frame_schur_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    CUDA_TRY(h, cudaFuncSetAttribute(frame_schur_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    CUDA_TRY(h, cudaFuncSetAttribute(frame_schur_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    CUDA_TRY(h, cudaFuncSetAttribute(frame_schur_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2SmemBytes));
    h->n_frame = n_frame; h->n_edge = n_edge; h->n_pts = M;
    h->edge_cam_h.assign(edge_cam, edge_cam + n_edge);
    h->edge_n_h.resize((size_t)n_edge);
    for (int e = 0; e < n_edge; ++e) h->edge_n_h[e] = edge_off[e + 1] - edge_off[e];
    {   // reduced-system solver: 3 = block cyclic reduction when the camera graph is block-banded (bandwidth <= 4 and the
        // super-blocks fit one CTA's shared memory), 2 = one-launch tile DAG otherwise; MCCBA_CHOL=2 forces the latter
        const char* sel = getenv("MCCBA_CHOL");
        int mode = band_fits(h->band_nw, P.ns) ? 3 : 2;
        if (sel && sel[0] == '2') mode = 2;
        h->k5_blocked = mode;
    }
    P.band_nw = (h->k5_blocked == 3 && P.ns > 0) ? h->band_nw : 0;
    P.ar_goff = P.band_nw > 0 ? (int64_t)P.ns * P.band_nw : (int64_t)P.ns * P.ns;
    if (P.ar_goff + P.ns + 4 >= (int64_t)2000000000) return fail(h, MCCBA_ERR_ARG, "reduced camera system too large (%d x %d)", P.ns, P.ns);
    h->ar_len = (int)P.ar_goff + P.ns + 4;
    if (h->opts.nranks > 1) {
        int rc_p2p = p2p_setup(h);
        if (rc_p2p) return rc_p2p;
    }
    for (int q = 0; q < 8; ++q) P.p2p_peer[q] = h->p2p_peer[q];
    P.p2p_n = h->p2p_ok ? h->opts.nranks : 0;
    P.p2p_rank = h->opts.rank;
    P.p2p_stride = h->p2p_stride;
    P.p2p_epoch = h->p2p_epoch;

    int rc;
#define UP(field, vec) if ((rc = dev_upload(h, &P.field, vec))) return rc
    UP(e_off, e_off); UP(e_cam, e_cam); UP(e_frame, e_frame);
    UP(slot_frame, slot_frame); UP(warp_group, warp_group); UP(group_V, group_V); UP(group_cam0, group_cam0);
    UP(group_cams, group_cams); UP(group_ebase, group_ebase); UP(group_stride, group_stride); UP(group_slot0, group_slot0);
    {
        const int* d_wmeta = nullptr;
        if ((rc = dev_upload(h, &d_wmeta, wmeta))) return rc;
        P.wmeta = reinterpret_cast<const int4*>(d_wmeta);
    }
    UP(warp_rec, warp_rec); UP(dest_info, dest_info); UP(dest_src0, dest_src0); UP(dest_src, dest_src);
#undef UP
    const int64_t* d_esrc = nullptr;
    if ((rc = dev_upload(h, &d_esrc, e_src))) return rc;
    lap("upload tables");
    // work buffers
    if ((rc = dev_alloc(h, &P.st, 1, true))) return rc;
    for (int b = 0; b < 2; ++b) {
        if ((rc = dev_alloc(h, &P.x[b], (size_t)P.n_param, true))) return rc;
        if ((rc = dev_alloc(h, &P.vR[b], 9 * (size_t)P.n_vertex, true))) return rc;
        if ((rc = dev_alloc(h, &P.blocks[b], (size_t)kBlk * P.n_edge_int, true))) return rc;
    }
    {   // vertex 0 is the gauge: identity rotation in both rotation buffers (never rewritten by the update kernels)
        static const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int b = 0; b < 2; ++b)
            CUDA_TRY(h, cudaMemcpyAsync(P.vR[b], I9, sizeof(I9), cudaMemcpyHostToDevice, h->stream));
    }
    if ((rc = dev_alloc(h, &P.frameL, 27 * (size_t)P.n_slots, true))) return rc;
    if ((rc = dev_alloc(h, &P.edgeY, 36 * (size_t)P.n_edge_int, true))) return rc;
    if ((rc = dev_alloc(h, &P.records, (size_t)rec_total, true))) return rc;
    if ((rc = dev_alloc(h, &P.warp_scal, 2 * (size_t)P.n_warps, true))) return rc;
    if ((rc = dev_alloc(h, &P.ar, (size_t)h->ar_len, true))) return rc;
    P.ar_part = P.ar;
    if (h->opts.nranks > 1 && (rc = dev_alloc(h, &P.ar_part, (size_t)h->ar_len, true))) return rc;
    if ((rc = dev_alloc(h, &h->x_saved, (size_t)P.n_param, true))) return rc;
    if ((rc = dev_alloc(h, &P.dc, (size_t)std::max(P.ns, 1), true))) return rc;
    if ((rc = dev_alloc(h, &P.norm_part, 2 * (size_t)P.n_k4_blocks, true))) return rc;
    const bool autop = h->prec == MCCBA_PRECISION_AUTO;
    P.prec = autop ? MCCBA_PRECISION_MIXED : h->prec;    // AUTO: resolved per evaluation (resolve_precision), both layouts are built
    P.edge_extent = nullptr;
    if ((rc = dev_alloc(h, &P.erec, (size_t)P.n_edge_int, true))) return rc;
    {
        std::vector<EdgeMeta> meta((size_t)P.n_edge_int);
        for (int e = 0; e < P.n_edge_int; ++e) {
            const bool live = e_frame[e] >= 0;
            meta[e].cam = e_cam[e]; meta[e].begin = live ? e_off[e] : 0; meta[e].end = live ? e_off[e + 1] : 0; meta[e].pad = 0;
        }
        if ((rc = dev_upload(h, &P.emeta, meta))) return rc;
    }
    if ((rc = dev_alloc(h, &P.err_sq, (size_t)P.n_edge_int, true))) return rc;
    if ((rc = dev_alloc(h, &P.err_nrm, (size_t)P.n_edge_int, true))) return rc;
    lap("allocate work buffers");
    // launch geometry
    {   // shared memory of the residual kernel: fixed part + two TMA stages sized for the largest 32-edge chunk,
        // provided two CTAs still fit on an SM (otherwise no staging: direct global loads)
        int max_chunk = 0;
        for (int ch = 0; ch < P.n_edge_int / kEdgesPerBlock; ++ch) {
            const int f0 = e_off[(size_t)ch * kEdgesPerBlock] & ~3, f1 = e_off[(size_t)(ch + 1) * kEdgesPerBlock];
            max_chunk = std::max(max_chunk, (f1 - f0 + 3) & ~3);
        }
        const size_t off_cam = (sizeof(K1Shared) + 15) & ~(size_t)15;
        const size_t fixed = ((off_cam + (size_t)nC * sizeof(CamParams) + 127) & ~(size_t)127);
        const size_t budget = 110 * 1024;   // per CTA, two CTAs per SM
        int cap = (max_chunk + 31) & ~31;
        const char* nostage = getenv("MCCBA_NO_TMA");
        if (nostage && nostage[0] == '1') cap = 0;
        while (cap > 0 && fixed + (size_t)cap * 40 > budget) cap -= 256;   // large edges: partial staging is not worth it
        if (cap < max_chunk) cap = (fixed + (size_t)max_chunk * 40 <= budget) ? ((max_chunk + 31) & ~31) : 0;
        h->obs_cap = cap;
        h->k1_smem = (int)(fixed + (size_t)cap * 40);
    }
    {   // the tile DAG spins on hand-overs between its CTAs: every one of them has to be resident at once
        const int ntc = chol_col_tiles(P.ns), ntr = chol_row_tiles(P.ns);
        int grid = 0;
        for (int j = 0; j < ntc; ++j) grid += ntr - j;
        if (h->k5_blocked == 2 && P.ns > 0) {
            int per_sm_dag = 0;
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_dag, chol_dag_kernel, 256, 0));
            if (per_sm_dag * h->num_sms < grid)
                return fail(h, MCCBA_ERR_ARG, "reduced camera system of %d unknowns needs %d co-resident CTAs, the device holds %d (dense camera graphs: up to ~250 cameras)", P.ns, grid, per_sm_dag * h->num_sms);
        }
        h->dag_grid = grid;
        h->iter_kernels = 6 + (P.ns > 0 ? (h->k5_blocked == 3 ? -1 : (P.ns <= 512 ? 0 : 1)) : 0);
        if ((rc = dev_alloc(h, &P.dag_buf, chol_dag_words(P.ns) + 8))) return rc;
        P.dag_words = (h->k5_blocked == 2 && P.ns > 0) ? (int64_t)chol_dag_words(P.ns) : 0;
    }
    int per_sm = 1;
    if (h->obs_cap > 0) {
        if (h->k1_smem > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(resid_jac_accum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->k1_smem));
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resid_jac_accum_kernel<true>, kK1Threads, h->k1_smem));
    } else {
        if (h->k1_smem > 48 * 1024)
            CUDA_TRY(h, cudaFuncSetAttribute(resid_jac_accum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->k1_smem));
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resid_jac_accum_kernel<false>, kK1Threads, h->k1_smem));
    }
    per_sm = std::max(per_sm, 1);
    h->k1_grid = std::max(1, std::min(P.n_edge_int / kEdgesPerBlock, h->num_sms * per_sm));
    lap("launch geometry");
    // last: everything above ran while the observation upload was in flight on its own stream
    if (autop) {
        float* ext = nullptr;
        if ((rc = dev_alloc(h, &ext, (size_t)P.n_edge_int))) return rc;
        if ((rc = dev_alloc(h, &h->d_ext_part, 64, true))) return rc;
        P.edge_extent = ext;
    }
    if (P.prec) {
        // packed pair layout of the single-precision pass (Problem::obs2): per tile 4 quarters x kp steps x 5 planes x 32
        // lanes x 2 floats, kp = ceil(max corners per edge of the tile / 8)
        const int n_tiles = P.n_edge_int / 32;
        std::vector<int64_t> tile_off((size_t)n_tiles + 1);
        std::vector<int> tile_kp((size_t)n_tiles);
        int64_t tot = 0;
        for (int t = 0; t < n_tiles; ++t) {
            int mx = 0;
            for (int e = 32 * t; e < 32 * t + 32; ++e) mx = std::max(mx, e_off[(size_t)e + 1] - e_off[(size_t)e]);
            const int kp = std::max(1, (mx + 7) / 8);    // a tile of empty edges still gets one (all-padding) step: the stream never issues an empty copy
            tile_kp[(size_t)t] = kp;
            tile_off[(size_t)t] = tot;
            tot += (int64_t)4 * kp * 5 * 32;
        }
        tile_off[(size_t)n_tiles] = tot;
        if ((rc = dev_upload(h, &P.tile_off, tile_off))) return rc;
        if ((rc = dev_upload(h, &P.tile_kp, tile_kp))) return rc;
        float2* obs2 = nullptr;
        if ((rc = dev_alloc(h, &obs2, (size_t)std::max<int64_t>(tot, 1)))) return rc;
        P.obs2 = obs2;
        int* nonplanar = nullptr;
        if ((rc = dev_alloc(h, &nonplanar, 1, true))) return rc;      // zeroed on the stream the gather kernel runs on
        P.obs_nonplanar = nonplanar;
        h->f32_smem = (int)f32_smem_bytes(nC);
        int per_sm_f = 1;
        if (P.prec == MCCBA_PRECISION_MIXED) {
            if (h->f32_smem > 48 * 1024)
                CUDA_TRY(h, cudaFuncSetAttribute(resid_jac_accum_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->f32_smem));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, resid_jac_accum_f32_kernel<true>, kF32Threads, h->f32_smem));
        } else {
            if (h->f32_smem > 48 * 1024)
                CUDA_TRY(h, cudaFuncSetAttribute(resid_jac_accum_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->f32_smem));
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, resid_jac_accum_f32_kernel<false>, kF32Threads, h->f32_smem));
        }
        h->f32_grid = std::max(1, std::min(n_tiles, h->num_sms * std::max(per_sm_f, 1)));
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
        gather_obs_packed_kernel<<<std::min(n_tiles, h->num_sms * 16), 256, 0, h->stream>>>(n_tiles, P.tile_off, P.tile_kp, P.e_off, d_esrc,
                                                                                            d_obj, d_img, obs2, nonplanar);
        if (autop) edge_extent_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(P.n_edge_int, P.e_off, d_esrc, d_obj, const_cast<float*>(P.edge_extent));
    }
    if (!P.prec || autop) {
    // observation planes: one allocation, each plane 256-byte aligned
    const size_t plane = ((size_t)M + 63) / 64 * 64;
    float* planes = nullptr;
    if ((rc = dev_alloc(h, &planes, plane * 5))) return rc;
    P.ox = planes; P.oy = planes + plane; P.oz = planes + 2 * plane; P.iu = planes + 3 * plane; P.iv = planes + 4 * plane;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    gather_obs_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(P.n_edge_int, P.e_off, d_esrc, d_obj, d_img, planes,
                                                            planes + plane, planes + 2 * plane, planes + 3 * plane,
                                                            planes + 4 * plane);
    }
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // also: the caller's buffers are no longer read
    drain_pool(h);   // whatever the new problem did not reuse
    lap("H2D observations + gather");
    h->cur = 0;
    h->have_obs = true;
    h->have_params = false;
    return MCCBA_OK;
}

int mccba_set_parameters(mccba_handle h, int64_t n, const double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_obs) return fail(h, MCCBA_ERR_STATE, "set_parameters before set_observations");
    if (!params || n != n_param(h)) return fail(h, MCCBA_ERR_ARG, "set_parameters: expected %lld doubles, got %lld", (long long)n_param(h), (long long)n);
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->P.x[h->cur], params, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_params = true;
    return MCCBA_OK;
}

int mccba_get_parameters(mccba_handle h, int64_t n, double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "get_parameters before set_parameters");
    if (!params || n != n_param(h)) return fail(h, MCCBA_ERR_ARG, "get_parameters: expected %lld doubles", (long long)n_param(h));
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(params, h->P.x[h->cur], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

int mccba_save_parameters(mccba_handle h)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "save_parameters before set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->x_saved, h->P.x[h->cur], sizeof(double) * (size_t)n_param(h), cudaMemcpyDeviceToDevice, h->stream));
    h->have_saved = true;
    return MCCBA_OK;
}

int mccba_restore_parameters(mccba_handle h)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_saved) return fail(h, MCCBA_ERR_STATE, "restore_parameters before save_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->P.x[h->cur], h->x_saved, sizeof(double) * (size_t)n_param(h), cudaMemcpyDeviceToDevice, h->stream));
    return MCCBA_OK;
}

int mccba_eval(mccba_handle h, double* cost, double* edge_H6, double* edge_g6, double* edge_cost)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "eval before set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    int rc;
    if ((rc = sync_state_cur(h))) return rc;
    if ((rc = launch_forced_eval(h))) return rc;
    const Problem& P = h->P;
    const size_t E = (size_t)P.n_edge_int;
    const bool full = edge_H6 || edge_g6;
    std::vector<double> buf(full ? kBlk * E : E);
    if (full) {
        CUDA_TRY(h, cudaMemcpyAsync(buf.data(), P.blocks[h->cur], sizeof(double) * buf.size(), cudaMemcpyDeviceToHost, h->stream));
    } else {   // cost only: row 27 of every tile (tile-major records, see tile_idx)
        CUDA_TRY(h, cudaMemcpy2DAsync(buf.data(), 32 * sizeof(double), P.blocks[h->cur] + 27 * 32, kBlk * 32 * sizeof(double),
                                      32 * sizeof(double), E / 32, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double total = 0;
    for (int e = 0; e < h->n_edge; ++e) {
        const size_t ie = (size_t)h->int_of_edge[e];
        const double c = full ? buf[tile_idx(kBlk, ie, 27)] : buf[ie];
        total += c;
        if (edge_cost) edge_cost[e] = c;
        if (edge_H6)
            for (int k = 0; k < 21; ++k) edge_H6[21 * (size_t)e + k] = buf[tile_idx(kBlk, ie, k)];
        if (edge_g6)
            for (int k = 0; k < 6; ++k) edge_g6[6 * (size_t)e + k] = buf[tile_idx(kBlk, ie, 21 + k)];
    }
    if (cost) *cost = total;
    return MCCBA_OK;
}

int mccba_reduced_system(mccba_handle h, double lambda, double* S, double* gs)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "reduced_system before set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    int rc;
    if ((rc = sync_state_cur(h))) return rc;
    if ((rc = launch_forced_eval(h))) return rc;
    Problem& P = h->P;
    CUDA_TRY(h, cudaMemsetAsync(P.ar, 0, sizeof(double) * (size_t)h->ar_len, h->stream));
    launch_schur(h, h->stream, h->cur, lambda);
    reduce_records_kernel<<<P.n_dest, kK3Threads, 0, h->stream>>>(P, 1);
    CUDA_TRY(h, cudaGetLastError());
    if (h->opts.nranks > 1) {
        ncclResult_t r = nccl().AllReduce(P.ar_part, P.ar, (size_t)h->ar_len, kNcclFloat64, kNcclSum, h->comm, h->stream);
        if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %d", r);
    }
    const size_t ns = (size_t)P.ns;
    if (P.band_nw > 0 && S && ns) {   // packed band -> dense symmetric matrix for the caller
        const size_t NW = (size_t)P.band_nw, w = NW - 1;
        std::vector<double> band(ns * NW);
        CUDA_TRY(h, cudaMemcpyAsync(band.data(), P.ar, sizeof(double) * ns * NW, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        std::fill(S, S + ns * ns, 0.0);
        for (size_t r = 0; r < ns; ++r)
            for (size_t k = 0; k < NW; ++k) {
                if (r + k < w) continue;
                const size_t c = r + k - w;
                S[r * ns + c] = band[r * NW + k];
                S[c * ns + r] = band[r * NW + k];
            }
    } else if (S && ns) {
        CUDA_TRY(h, cudaMemcpyAsync(S, P.ar, sizeof(double) * ns * ns, cudaMemcpyDeviceToHost, h->stream));
    }
    if (gs && ns) CUDA_TRY(h, cudaMemcpyAsync(gs, P.ar + P.ar_goff, sizeof(double) * ns, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

int mccba_solve(mccba_handle h, const mccba_solve_opts* o, mccba_report* rep)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!o) return fail(h, MCCBA_ERR_ARG, "solve: null options");
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "solve before set_parameters");
    if (o->mode != MCCBA_MODE_REFERENCE_GN && o->mode != MCCBA_MODE_LM) return fail(h, MCCBA_ERR_ARG, "solve: unknown mode %d", o->mode);
    if (o->crit_type < 1 || o->crit_type > 3) return fail(h, MCCBA_ERR_ARG, "solve: criteria type %d (1=COUNT, 2=EPS, 3=both)", o->crit_type);
    if (o->max_count < 0) return fail(h, MCCBA_ERR_ARG, "solve: negative max_count");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    Problem& P = h->P;
    cudaStream_t s = h->stream;
    int rc;
    int kernels = 0;
    if (h->p2p_ok) {
        // re-agree the exchange epoch: after an error on one rank (or unequal launch counts) the per-rank counters may
        // differ; the maximum is newer than every flag word any rank has written so far
        ncclResult_t r = nccl().AllReduce(h->p2p_epoch, h->p2p_epoch, 1, 5 /* ncclUint64 */, 2 /* ncclMax */, h->comm, s);
        if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce (epoch) failed: %d", r);
    }
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    init_state_kernel<<<1, 1, 0, s>>>(P.st, o->mode, o->crit_type, o->max_count, o->epsilon, o->lambda0, o->lambda_up,
                                      o->lambda_down, h->cur);
    CUDA_TRY(h, cudaMemsetAsync(P.norm_part, 0, sizeof(double) * 2 * (size_t)P.n_k4_blocks, s));
    if ((rc = launch_forced_eval(h))) return rc;
    kernels += 4;
    // one iteration as a CUDA graph (captured once per problem)
    const bool use_graph = h->opts.use_graph && !h->profile;
    if (use_graph && !h->graph) {
        cudaGraph_t g = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        rc = enqueue_iteration(h, false);
        cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&h->graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        h->graph_prec = P.prec;
    }
    for (double& v : h->prof_ms) v = 0;
    const bool has_count = (o->crit_type & MCCBA_CRIT_COUNT) != 0;
    int64_t max_launches = has_count ? (o->mode == MCCBA_MODE_REFERENCE_GN ? (int64_t)o->max_count + 1 : 2 * (int64_t)o->max_count + 4)
                                     : 200000;
    const bool exact = has_count && o->mode == MCCBA_MODE_REFERENCE_GN && !(o->crit_type & MCCBA_CRIT_EPS);
    const int chunk = 8;
    int64_t launched = 0;
    int slot = 0;
    cudaEvent_t evs[2];
    cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming);
    bool pending[2] = {false, false};
    bool stop = false;
    while (launched < max_launches && !stop) {
        const int n = (int)std::min<int64_t>(exact ? max_launches : chunk, max_launches - launched);
        for (int i = 0; i < n; ++i) {
            if (use_graph) CUDA_TRY(h, cudaGraphLaunch(h->graph, s));
            else if ((rc = enqueue_iteration(h, h->profile != 0))) return rc;
        }
        launched += n;
        if (exact) break;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_done + slot, &P.st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaEventRecord(evs[slot], s));
        pending[slot] = true;
        const int prev = slot ^ 1;
        if (pending[prev]) {  // look one chunk behind so the device never waits for the host
            CUDA_TRY(h, cudaEventSynchronize(evs[prev]));
            if (h->h_done[prev]) stop = true;
            pending[prev] = false;
        }
        slot ^= 1;
    }
    cudaEventDestroy(evs[0]);
    cudaEventDestroy(evs[1]);
    kernels += (int)launched * h->iter_kernels;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_state, P.st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    CUDA_TRY(h, cudaGetLastError());
    const DevState& st = *h->h_state;
    h->cur = st.cur;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (h->profile && st.launches > 0)
        for (double& v : h->prof_ms) v /= (double)launched;
    if (rep) {
        rep->iterations = st.iter; rep->accepted = st.n_accept; rep->rejected = st.n_reject;
        rep->status = st.status; rep->graph_launches = (int)launched; rep->kernel_launches = kernels;
        rep->change = st.change; rep->cost = st.cost_cur; rep->lambda = st.lambda; rep->device_ms = ms;
    }
    if (st.status == 5)
        return fail(h, MCCBA_ERR_NCCL, "solve: the peer-memory exchange timed out at iteration %d (a peer rank is gone or out of step)", st.iter);
    if (st.status != 0)
        return fail(h, MCCBA_ERR_NUMERIC, "solve: numeric failure at iteration %d (non-finite cost or a block that is not positive definite)", st.iter);
    if (!st.done)
        return fail(h, MCCBA_ERR_NUMERIC, "solve: launch budget exhausted before the termination test fired (iter %d)", st.iter);
    return MCCBA_OK;
}

int mccba_exchange_stats(mccba_handle h, double out[4])
{
    if (!h || !out) return MCCBA_ERR_ARG;
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (!h->p2p_ok || !h->p2p_epoch) return MCCBA_OK;
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    unsigned long long w[4] = {0, 0, 0, 0};
    CUDA_TRY(h, cudaMemcpyAsync(w, h->p2p_epoch + kP2pStatsWord, sizeof(w), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->p2p_epoch + kP2pStatsWord, 0, sizeof(w), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (w[2]) { out[0] = 1e-3 * (double)w[0] / (double)w[2]; out[1] = 1e-3 * (double)w[1] / (double)w[2]; }
    out[2] = (double)w[2];
    out[3] = 1e-3 * (double)w[3];
    return MCCBA_OK;
}

int mccba_allreduce_sum(mccba_handle h, double* buf, int n)
{
    if (!h || !buf || n < 0 || n > 64) return MCCBA_ERR_ARG;
    if (h->opts.nranks == 1 || n == 0) return MCCBA_OK;
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    memcpy(h->h_small, buf, sizeof(double) * (size_t)n);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_small, h->h_small, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = nccl().AllReduce(h->d_small, h->d_small, (size_t)n, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %d", r);
    CUDA_TRY(h, cudaMemcpyAsync(h->h_small, h->d_small, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    memcpy(buf, h->h_small, sizeof(double) * (size_t)n);
    return MCCBA_OK;
}

int mccba_reproj_error(mccba_handle h, mccba_error_stats* stats, double* per_edge_mean)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "reproj_error before set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    int rc;
    if ((rc = sync_state_cur(h))) return rc;
    Problem& P = h->P;
    vertex_prep_kernel<<<(P.n_vertex + 127) / 128, 128, 0, h->stream>>>(P, -1);
    if (P.prec) reproj_error_packed_kernel<<<h->num_sms * 8, kF32Threads, 0, h->stream>>>(P);
    else reproj_error_kernel<<<h->num_sms * 4, kK1Threads, 0, h->stream>>>(P);
    CUDA_TRY(h, cudaGetLastError());
    const size_t E = (size_t)P.n_edge_int;
    std::vector<double> sq(E), nr(E);
    CUDA_TRY(h, cudaMemcpyAsync(sq.data(), P.err_sq, sizeof(double) * E, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(nr.data(), P.err_nrm, sizeof(double) * E, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double acc[4] = {0, 0, 0, 0};  // sum_norm, sum_sq, reference point count, points
    for (int e = 0; e < h->n_edge; ++e) {
        const size_t ie = (size_t)h->int_of_edge[e];
        const int64_t n = h->edge_n_h[e];
        if (per_edge_mean) per_edge_mean[e] = n ? nr[ie] / (double)n : 0.0;
        acc[0] += nr[ie];
        acc[1] += sq[ie];
        // src/multicalib.cpp:983: error.total() is 2N for PINHOLE (N x 2 single channel) and N for OMNIDIRECTIONAL
        acc[2] += (double)(h->cams[h->edge_cam_h[e]].model == MCCBA_PINHOLE ? 2 * n : n);
        acc[3] += (double)n;
    }
    const double local_norm = acc[0], local_sq = acc[1], local_n = acc[3];
    if ((rc = mccba_allreduce_sum(h, acc, 4))) return rc;
    if (stats) {
        stats->mean_reproj_error = acc[2] > 0 ? acc[0] / acc[2] : 0.0;
        stats->rms = acc[3] > 0 ? sqrt(acc[1] / acc[3]) : 0.0;
        stats->sum_norm = local_norm; stats->sum_sq = local_sq; stats->n_points = (int64_t)local_n;
    }
    return MCCBA_OK;
}

int mccba_debug_solve_dense(mccba_handle h, int n, const double* S, const double* g, double* x, int blocked)
{
    if (!h || n < 1 || !S || !g || !x) return MCCBA_ERR_ARG;
    if (blocked != 2 && blocked != 3) return fail(h, MCCBA_ERR_ARG, "debug_solve_dense: solver %d (2 = tile DAG, 3 = block cyclic reduction)", blocked);
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    double *dA = nullptr, *dx = nullptr;
    int* dfail = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&dA, sizeof(double) * (size_t)(n + 1) * n));
    CUDA_TRY(h, cudaMalloc((void**)&dx, sizeof(double) * (size_t)n));
    CUDA_TRY(h, cudaMalloc((void**)&dfail, sizeof(int)));
    CUDA_TRY(h, cudaMemsetAsync(dfail, 0, sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dA, S, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dA + (size_t)n * n, g, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    auto release = [&]() { cudaFree(dA); cudaFree(dx); cudaFree(dfail); };
    double* dflags = nullptr;
    Problem none;
    memset(&none, 0, sizeof(none));
    if (blocked == 3) {   // block cyclic reduction: the bandwidth in 6 x 6 blocks is measured on the host copy
        int m = 1;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j)
                if (S[(size_t)i * n + j] != 0.0) m = std::max(m, i / 6 - j / 6);
        const int nw = 6 * (m + 1);
        if (!band_fits(nw, n)) { release(); return fail(h, MCCBA_ERR_ARG, "block bandwidth %d is too wide for the banded solver", m); }
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        CUDA_TRY(h, launch_band(nw, dA, n, dx, dfail, nullptr, none, 0, 0, h->stream));
    } else {
        const int ntc = chol_col_tiles(n), ntr = chol_row_tiles(n);
        int grid = 0;
        for (int j = 0; j < ntc; ++j) grid += ntr - j;
        int per_sm_dag = 0;   // every CTA of the DAG must be resident at once (see chol_dag_tile)
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_dag, chol_dag_kernel, 256, 0));
        if (per_sm_dag * h->num_sms < grid) {
            release();
            return fail(h, MCCBA_ERR_ARG, "n = %d needs %d co-resident CTAs for the tile DAG, the device holds %d", n, grid, per_sm_dag * h->num_sms);
        }
        CUDA_TRY(h, cudaMalloc((void**)&dflags, sizeof(double) * chol_dag_words(n)));
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        CUDA_TRY(h, cudaMemsetAsync(dflags, 0xFF, sizeof(double) * chol_dag_words(n), h->stream));
        CholDag D{dA, dflags, n, dx, nullptr, dfail, nullptr};
        chol_dag_kernel<<<grid, 256, 0, h->stream>>>(D, none, 0);
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    int f = 0;
    CUDA_TRY(h, cudaMemcpyAsync(x, dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(&f, dfail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->prof_ms[0] = ms;
    release();
    if (dflags) cudaFree(dflags);
    if (!f)
        for (int i = 0; i < n; ++i)
            if (!std::isfinite(x[i])) f = 1;     // the tile DAG reports a bad pivot as non-finite entries of the solution
    return f ? fail(h, MCCBA_ERR_NUMERIC, "matrix is not positive definite") : MCCBA_OK;
}

int mccba_exchange_mode(mccba_handle h)
{
    if (!h || h->opts.nranks <= 1) return 0;
    return h->p2p_ok ? 2 : 1;
}

int mccba_last_kernel_ms(mccba_handle h, double out[6])
{
    if (!h || !out) return MCCBA_ERR_ARG;
    for (int i = 0; i < 6; ++i) out[i] = h->prof_ms[i];
    return MCCBA_OK;
}

int mccba_time_eval(mccba_handle h, int reps, double* avg_ms)
{
    if (!h || reps < 1 || !avg_ms) return MCCBA_ERR_ARG;
    if (!h->have_params) return fail(h, MCCBA_ERR_STATE, "time_eval before set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    int rc;
    if ((rc = sync_state_cur(h))) return rc;
    if ((rc = launch_forced_eval(h))) return rc;  // warm-up, also computes the rotations
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < reps; ++i)
        launch_resid(h, h->stream, 1);
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    *avg_ms = (double)ms / reps;
    return MCCBA_OK;
}

}  // extern "C"

// ---- single-camera Mei calibration loop (cv::omnidir::calibrate, src/omnidir.cpp:1119-1147) -------------------
namespace {
template <typename T>
int omni_alloc(mccba_handle h, T** p, size_t count)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(h, cudaMalloc(&q, bytes));
    CUDA_TRY(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    h->omni_allocs.push_back(q);
    *p = (T*)q;
    return MCCBA_OK;
}
// sum over the ranks in place (no-op for a single rank); inside the captured iteration graph as well
int omni_allreduce(mccba_handle h, double* buf, int n)
{
    if (h->opts.nranks <= 1) return MCCBA_OK;
    ncclResult_t r = nccl().AllReduce(buf, buf, (size_t)n, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (r != 0) return fail(h, MCCBA_ERR_NCCL, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    return MCCBA_OK;
}
int omni_enqueue_iteration(mccba_handle h)
{
    OmniProblem& O = h->O;
    cudaStream_t s = h->stream;
    omni_frame_kernel<<<(O.n_frame + kOmniWarps - 1) / kOmniWarps, kOmniThreads, 0, s>>>(O, 0);
    omni_reduce_kernel<<<kOmniRec, 256, 0, s>>>(O, 0);
    int rc;
    if ((rc = omni_allreduce(h, O.tot, kOmniRec))) return rc;     // frames shard over the ranks: 78 doubles
    omni_solve_kernel<<<1, 32, 0, s>>>(O);                         // identical on every rank
    omni_update_kernel<<<O.n_blocks_upd, 128, 0, s>>>(O);
    if (h->opts.nranks > 1) {
        omni_decide_kernel<<<1, 256, 0, s>>>(O, 1);
        if ((rc = omni_allreduce(h, O.norm_tot, 2))) return rc;
        omni_decide_kernel<<<1, 256, 0, s>>>(O, 2);
    } else {
        omni_decide_kernel<<<1, 256, 0, s>>>(O, 0);
    }
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}
}  // namespace

extern "C" {

int mccba_omni_set_observations(mccba_handle h, int n_frame, const int64_t* frame_off, const float* obj_xyz, const float* img_uv)
{
    if (!h) return MCCBA_ERR_ARG;
    if (n_frame < 1 || !frame_off || !obj_xyz || !img_uv || frame_off[0] != 0) return fail(h, MCCBA_ERR_ARG, "omni_set_observations: bad input");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    const int64_t M = frame_off[n_frame];
    if (M <= 0 || M >= (int64_t)2000000000) return fail(h, MCCBA_ERR_ARG, "omni_set_observations: corner count out of range");
    for (int f = 0; f < n_frame; ++f)
        if (frame_off[f + 1] <= frame_off[f]) return fail(h, MCCBA_ERR_ARG, "frame %d has no observation", f);
    if (h->omni_graph) { cudaGraphExecDestroy(h->omni_graph); h->omni_graph = nullptr; }
    for (void* q : h->omni_allocs) cudaFree(q);
    h->omni_allocs.clear();
    OmniProblem& O = h->O;
    memset(&O, 0, sizeof(O));
    O.n_frame = n_frame; O.n_pts = M; O.n_blocks_upd = (n_frame + 127) / 128;
    std::vector<int> off((size_t)n_frame + 1);
    std::vector<int64_t> src((size_t)n_frame);
    for (int f = 0; f <= n_frame; ++f) off[f] = (int)frame_off[f];
    for (int f = 0; f < n_frame; ++f) src[f] = frame_off[f];
    int* d_off = nullptr; int64_t* d_src = nullptr; float* planes = nullptr;
    int rc;
    if ((rc = omni_alloc(h, &d_off, off.size()))) return rc;
    if ((rc = omni_alloc(h, &d_src, src.size()))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_src, src.data(), src.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    const size_t plane = ((size_t)M + 63) / 64 * 64;
    if ((rc = omni_alloc(h, &planes, plane * 5))) return rc;
    float *d_obj = nullptr, *d_img = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d_obj, sizeof(float) * 3 * (size_t)M));
    CUDA_TRY(h, cudaMalloc((void**)&d_img, sizeof(float) * 2 * (size_t)M));
    CUDA_TRY(h, cudaMemcpyAsync(d_obj, obj_xyz, sizeof(float) * 3 * (size_t)M, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_img, img_uv, sizeof(float) * 2 * (size_t)M, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_copy, h->copy_stream));
    struct CopyGuard {   // an early error return must not leave the engine reading the caller's buffers
        cudaStream_t s;
        ~CopyGuard() { cudaStreamSynchronize(s); }
    } copy_guard{h->copy_stream};
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));   // the gather reads what the copy stream uploads
    gather_obs_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(n_frame, d_off, d_src, d_obj, d_img, planes, planes + plane,
                                                            planes + 2 * plane, planes + 3 * plane, planes + 4 * plane);
    CUDA_TRY(h, cudaGetLastError());
    O.ox = planes; O.oy = planes + plane; O.oz = planes + 2 * plane; O.iu = planes + 3 * plane; O.iv = planes + 4 * plane;
    O.f_off = d_off;
    if ((rc = omni_alloc(h, &O.param, 6 * (size_t)n_frame + 10))) return rc;
    if ((rc = omni_alloc(h, &O.rec, (size_t)kOmniRec * n_frame))) return rc;
    if ((rc = omni_alloc(h, &O.save, (size_t)kOmniSave * n_frame))) return rc;
    if ((rc = omni_alloc(h, &O.tot, kOmniRec))) return rc;
    if ((rc = omni_alloc(h, &O.norm_part, 2 * (size_t)O.n_blocks_upd))) return rc;
    if ((rc = omni_alloc(h, &O.norm_tot, 2))) return rc;
    O.count_intr = h->opts.rank == 0 ? 1 : 0;
    if ((rc = omni_alloc(h, &O.st, 1))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_obj);
    cudaFree(d_img);
    h->omni_have = true;
    h->omni_have_params = false;
    return MCCBA_OK;
}

int mccba_omni_set_parameters(mccba_handle h, int64_t n, const double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->omni_have) return fail(h, MCCBA_ERR_STATE, "omni_set_parameters before omni_set_observations");
    if (!params || n != 6 * (int64_t)h->O.n_frame + 10) return fail(h, MCCBA_ERR_ARG, "omni_set_parameters: expected %lld doubles", (long long)(6 * (int64_t)h->O.n_frame + 10));
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->O.param, params, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->omni_have_params = true;
    return MCCBA_OK;
}

int mccba_omni_get_parameters(mccba_handle h, int64_t n, double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->omni_have_params) return fail(h, MCCBA_ERR_STATE, "omni_get_parameters before omni_set_parameters");
    if (!params || n != 6 * (int64_t)h->O.n_frame + 10) return fail(h, MCCBA_ERR_ARG, "omni_get_parameters: wrong size");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(params, h->O.param, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

int mccba_omni_solve(mccba_handle h, int flags, int crit_type, int max_count, double epsilon, mccba_report* rep)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->omni_have_params) return fail(h, MCCBA_ERR_STATE, "omni_solve before omni_set_parameters");
    if (crit_type < 1 || crit_type > 3 || max_count < 0) return fail(h, MCCBA_ERR_ARG, "omni_solve: bad criteria");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    OmniProblem& O = h->O;
    cudaStream_t s = h->stream;
    int rc;
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    omni_init_state_kernel<<<1, 1, 0, s>>>(O.st, flags, crit_type, max_count, epsilon);
    if (h->opts.use_graph && !h->omni_graph) {
        cudaGraph_t g = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        rc = omni_enqueue_iteration(h);
        cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&h->omni_graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
    }
    const bool exact = crit_type == 1;
    const int64_t max_launches = (crit_type & 1) ? max_count : 200000;
    int64_t launched = 0;
    int slot = 0;
    cudaEvent_t evs[2];
    cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming);
    bool pending[2] = {false, false}, stop = false;
    while (launched < max_launches && !stop) {
        const int nl = (int)std::min<int64_t>(exact ? max_launches : 8, max_launches - launched);
        for (int i = 0; i < nl; ++i) {
            if (h->opts.use_graph) CUDA_TRY(h, cudaGraphLaunch(h->omni_graph, s));
            else if ((rc = omni_enqueue_iteration(h))) return rc;
        }
        launched += nl;
        if (exact) break;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_done + slot, &O.st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaEventRecord(evs[slot], s));
        pending[slot] = true;
        const int prev = slot ^ 1;
        if (pending[prev]) {
            CUDA_TRY(h, cudaEventSynchronize(evs[prev]));
            if (h->h_done[prev]) stop = true;
            pending[prev] = false;
        }
        slot ^= 1;
    }
    cudaEventDestroy(evs[0]);
    cudaEventDestroy(evs[1]);
    // final cost at the returned parameters (estimateUncertainties' rms, src/omnidir.cpp:1794-1802)
    omni_frame_kernel<<<(O.n_frame + kOmniWarps - 1) / kOmniWarps, kOmniThreads, 0, s>>>(O, 1);
    omni_reduce_kernel<<<kOmniRec, 256, 0, s>>>(O, 1);
    if ((rc = omni_allreduce(h, O.tot, kOmniRec))) return rc;
    OmniState hs;
    double cost = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&hs, O.st, sizeof(OmniState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&cost, O.tot + 77, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    CUDA_TRY(h, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (rep) {
        memset(rep, 0, sizeof(*rep));
        rep->iterations = hs.iter; rep->accepted = hs.iter; rep->status = hs.status;
        rep->graph_launches = (int)launched; rep->kernel_launches = (int)launched * (h->opts.nranks > 1 ? 6 : 5) + 3;
        rep->change = hs.change; rep->cost = cost; rep->lambda = hs.epsilon; rep->device_ms = ms;
    }
    if (hs.status) return fail(h, MCCBA_ERR_NUMERIC, "omni_solve: numeric failure at iteration %d", hs.iter);
    if (!hs.done) return fail(h, MCCBA_ERR_NUMERIC, "omni_solve: launch budget exhausted (iter %d)", hs.iter);
    return MCCBA_OK;
}

int mccba_omni_gram(mccba_handle h, double* gram /* n_frame x 17 x 17 */, double* cost)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->omni_have_params) return fail(h, MCCBA_ERR_STATE, "omni_gram before omni_set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    OmniProblem O = h->O;
    double* d = nullptr;
    const size_t cnt = (size_t)O.n_frame * 289;
    if (gram) { CUDA_TRY(h, cudaMalloc((void**)&d, sizeof(double) * cnt)); O.dump = d; }
    omni_frame_kernel<<<(O.n_frame + kOmniWarps - 1) / kOmniWarps, kOmniThreads, 0, h->stream>>>(O, 1);
    omni_reduce_kernel<<<kOmniRec, 256, 0, h->stream>>>(O, 1);
    {
        int rc2 = omni_allreduce(h, O.tot, kOmniRec);      // cost: the whole job's
        if (rc2) { if (d) cudaFree(d); return rc2; }
    }
    if (gram) CUDA_TRY(h, cudaMemcpyAsync(gram, d, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
    if (cost) CUDA_TRY(h, cudaMemcpyAsync(cost, O.tot + 77, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    if (d) cudaFree(d);
    return MCCBA_OK;
}

}  // extern "C"

// ---- omnidir stereo bundle adjustment (cv::omnidir::stereoCalibrate's loop + estimateUncertaintiesStereo) ------------
namespace {
template <typename T>
int stereo_alloc(mccba_handle h, T** p, size_t count)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(h, cudaMalloc(&q, bytes));
    CUDA_TRY(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    h->stereo_allocs.push_back(q);
    *p = (T*)q;
    return MCCBA_OK;
}
int stereo_enqueue_iteration(mccba_handle h)
{
    StereoProblem& S = h->S;
    cudaStream_t s = h->stream;
    stereo_frame_kernel<<<(S.n_frame + kStWarps - 1) / kStWarps, kStThreads, kStFrameSmem, s>>>(S, 0);
    stereo_reduce_kernel<<<kStRec, 256, 0, s>>>(S, 0);
    stereo_solve_kernel<<<1, 32, 0, s>>>(S);
    stereo_update_kernel<<<S.n_blocks_upd, 128, 0, s>>>(S);
    stereo_decide_kernel<<<1, 256, 0, s>>>(S);
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}
}  // namespace

extern "C" {

int mccba_stereo_set_observations(mccba_handle h, int n_frame, const int64_t* frame_off, const float* obj_xyz, const float* img1_uv,
                                  const float* img2_uv)
{
    if (!h) return MCCBA_ERR_ARG;
    if (n_frame < 1 || !frame_off || !obj_xyz || !img1_uv || !img2_uv || frame_off[0] != 0) return fail(h, MCCBA_ERR_ARG, "stereo_set_observations: bad input");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    const int64_t M = frame_off[n_frame];
    if (M <= 0 || M >= (int64_t)1000000000) return fail(h, MCCBA_ERR_ARG, "stereo_set_observations: corner count out of range");
    for (int f = 0; f < n_frame; ++f)
        if (frame_off[f + 1] <= frame_off[f]) return fail(h, MCCBA_ERR_ARG, "frame %d has no observation", f);
    if (h->stereo_graph) { cudaGraphExecDestroy(h->stereo_graph); h->stereo_graph = nullptr; }
    CUDA_TRY(h, cudaFuncSetAttribute(stereo_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStFrameSmem));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (void* q : h->stereo_allocs) cudaFree(q);
    h->stereo_allocs.clear();
    StereoProblem& S = h->S;
    memset(&S, 0, sizeof(S));
    S.n_frame = n_frame; S.n_pts = M; S.n_blocks_upd = (n_frame + 127) / 128;
    std::vector<int> off((size_t)n_frame + 1);
    for (int f = 0; f <= n_frame; ++f) off[f] = (int)frame_off[f];
    // SoA planes (x y z | u1 v1 | u2 v2), split on the host: the problem sizes of this path are small
    std::vector<float> planes(7 * (size_t)M);
    for (int64_t i = 0; i < M; ++i) {
        planes[i] = obj_xyz[3 * i]; planes[M + i] = obj_xyz[3 * i + 1]; planes[2 * M + i] = obj_xyz[3 * i + 2];
        planes[3 * M + i] = img1_uv[2 * i]; planes[4 * M + i] = img1_uv[2 * i + 1];
        planes[5 * M + i] = img2_uv[2 * i]; planes[6 * M + i] = img2_uv[2 * i + 1];
    }
    int* d_off = nullptr; float* d_pl = nullptr;
    int rc;
    if ((rc = stereo_alloc(h, &d_off, off.size()))) return rc;
    if ((rc = stereo_alloc(h, &d_pl, planes.size()))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_pl, planes.data(), planes.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    S.f_off = d_off;
    S.ox = d_pl; S.oy = d_pl + M; S.oz = d_pl + 2 * M; S.u1 = d_pl + 3 * M; S.v1 = d_pl + 4 * M; S.u2 = d_pl + 5 * M; S.v2 = d_pl + 6 * M;
    const size_t np = 6 * ((size_t)n_frame + 1) + 20;
    if ((rc = stereo_alloc(h, &S.param, np))) return rc;
    if ((rc = stereo_alloc(h, &S.rec, (size_t)kStRec * n_frame))) return rc;
    if ((rc = stereo_alloc(h, &S.save, (size_t)kStSave * n_frame))) return rc;
    if ((rc = stereo_alloc(h, &S.tot, kStRec))) return rc;
    if ((rc = stereo_alloc(h, &S.norm_part, 2 * (size_t)S.n_blocks_upd))) return rc;
    if ((rc = stereo_alloc(h, &S.sinv, kStNS * kStNS))) return rc;
    if ((rc = stereo_alloc(h, &S.diag, np))) return rc;
    if ((rc = stereo_alloc(h, &S.st, 1))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->stereo_have = true;
    h->stereo_have_params = false;
    return MCCBA_OK;
}

int mccba_stereo_set_parameters(mccba_handle h, int64_t n, const double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->stereo_have) return fail(h, MCCBA_ERR_STATE, "stereo_set_parameters before stereo_set_observations");
    const int64_t np = 6 * ((int64_t)h->S.n_frame + 1) + 20;
    if (!params || n != np) return fail(h, MCCBA_ERR_ARG, "stereo_set_parameters: expected %lld doubles", (long long)np);
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->S.param, params, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->stereo_have_params = true;
    return MCCBA_OK;
}

int mccba_stereo_get_parameters(mccba_handle h, int64_t n, double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->stereo_have_params) return fail(h, MCCBA_ERR_STATE, "stereo_get_parameters before stereo_set_parameters");
    const int64_t np = 6 * ((int64_t)h->S.n_frame + 1) + 20;
    if (!params || n != np) return fail(h, MCCBA_ERR_ARG, "stereo_get_parameters: wrong size");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(params, h->S.param, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

int mccba_stereo_solve(mccba_handle h, int flags, int crit_type, int max_count, double epsilon, mccba_report* rep)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->stereo_have_params) return fail(h, MCCBA_ERR_STATE, "stereo_solve before stereo_set_parameters");
    if (crit_type < 1 || crit_type > 3 || max_count < 0) return fail(h, MCCBA_ERR_ARG, "stereo_solve: bad criteria");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    StereoProblem& S = h->S;
    cudaStream_t s = h->stream;
    int rc;
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    stereo_init_state_kernel<<<1, 1, 0, s>>>(S.st, flags, crit_type, max_count, epsilon);
    if (h->opts.use_graph && !h->stereo_graph) {
        cudaGraph_t g = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        rc = stereo_enqueue_iteration(h);
        cudaError_t ce = cudaStreamEndCapture(s, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&h->stereo_graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) return fail(h, MCCBA_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
    }
    const bool exact = crit_type == 1;
    const int64_t max_launches = (crit_type & 1) ? max_count : 200000;
    int64_t launched = 0;
    int slot = 0;
    cudaEvent_t evs[2];
    cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming);
    bool pending[2] = {false, false}, stop = false;
    while (launched < max_launches && !stop) {
        const int nl = (int)std::min<int64_t>(exact ? max_launches : 8, max_launches - launched);
        for (int i = 0; i < nl; ++i) {
            if (h->opts.use_graph) CUDA_TRY(h, cudaGraphLaunch(h->stereo_graph, s));
            else if ((rc = stereo_enqueue_iteration(h))) return rc;
        }
        launched += nl;
        if (exact) break;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_done + slot, &S.st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaEventRecord(evs[slot], s));
        pending[slot] = true;
        const int prev = slot ^ 1;
        if (pending[prev]) {
            CUDA_TRY(h, cudaEventSynchronize(evs[prev]));
            if (h->h_done[prev]) stop = true;
            pending[prev] = false;
        }
        slot ^= 1;
    }
    cudaEventDestroy(evs[0]);
    cudaEventDestroy(evs[1]);
    // final cost at the returned parameters (estimateUncertaintiesStereo's rms, :1879-1888)
    stereo_frame_kernel<<<(S.n_frame + kStWarps - 1) / kStWarps, kStThreads, kStFrameSmem, s>>>(S, 1);
    stereo_reduce_kernel<<<kStRec, 256, 0, s>>>(S, 1);
    StereoState hs;
    double cost = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&hs, S.st, sizeof(StereoState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&cost, S.tot + kStSTri + 2 * kStNS + 2, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    CUDA_TRY(h, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (rep) {
        memset(rep, 0, sizeof(*rep));
        rep->iterations = hs.iter; rep->accepted = hs.iter; rep->status = hs.status;
        rep->graph_launches = (int)launched; rep->kernel_launches = (int)launched * (h->opts.nranks > 1 ? 6 : 5) + 3;
        rep->change = hs.change; rep->cost = cost; rep->lambda = hs.epsilon; rep->device_ms = ms;
    }
    if (hs.status) return fail(h, MCCBA_ERR_NUMERIC, "stereo_solve: numeric failure at iteration %d", hs.iter);
    if (!hs.done) return fail(h, MCCBA_ERR_NUMERIC, "stereo_solve: launch budget exhausted (iter %d)", hs.iter);
    return MCCBA_OK;
}

int mccba_stereo_uncertainties(mccba_handle h, int flags, double* errors, double std_error[2], double* rms)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->stereo_have_params) return fail(h, MCCBA_ERR_STATE, "stereo_uncertainties before stereo_set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    StereoProblem& S = h->S;
    cudaStream_t s = h->stream;
    const size_t np = 6 * ((size_t)S.n_frame + 1) + 20;
    CUDA_TRY(h, cudaMemsetAsync(&S.st->status, 0, sizeof(int), s));
    stereo_frame_kernel<<<(S.n_frame + kStWarps - 1) / kStWarps, kStThreads, kStFrameSmem, s>>>(S, 1);
    stereo_reduce_kernel<<<kStRec, 256, 0, s>>>(S, 1);
    stereo_cov_shared_kernel<<<1, 32, 0, s>>>(S, flags);
    stereo_cov_frame_kernel<<<(S.n_frame + 127) / 128, 128, 0, s>>>(S);
    CUDA_TRY(h, cudaGetLastError());
    std::vector<double> diag(np);
    double mom[7];
    int status = 0;
    CUDA_TRY(h, cudaMemcpyAsync(diag.data(), S.diag, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(mom, S.tot + kStSTri + 2 * kStNS, sizeof(double) * 7, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&status, &S.st->status, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    if (status) return fail(h, MCCBA_ERR_NUMERIC, "stereo_uncertainties: the normal matrix is singular");
    // moments: mom[2] = sum |e|^2, mom[3..6] = sum ex, ey, ex^2, ey^2 over the N = 2 * corners residual points
    const double N = 2.0 * (double)S.n_pts;
    const double mx = mom[3] / N, my = mom[4] / N;
    const double vx = mom[5] / N - mx * mx, vy = mom[6] / N - my * my;
    if (std_error) { std_error[0] = sqrt(vx) * sqrt(N / (N - 1.0)); std_error[1] = sqrt(vy) * sqrt(N / (N - 1.0)); }   // :1862-1863
    const double mall = (mom[3] + mom[4]) / (2 * N), vall = (mom[5] + mom[6]) / (2 * N) - mall * mall;
    const double sdev = sqrt(vall) * sqrt(2.0 * N / (2.0 * N - 1.0));                                                    // :1865-1868
    if (errors)
        for (size_t i = 0; i < np; ++i) errors[i] = 3.0 * sdev * sqrt(diag[i]);                                          // :1875
    if (rms) *rms = sqrt(mom[2] / N);
    return MCCBA_OK;
}

}  // extern "C"

// ---- double-sided board calibration (cv::multicalib::DoubleSideCalibration's optimisation, src/doubleSide.cpp) ---------
namespace {
template <typename T>
int ds_alloc(mccba_handle h, T** p, size_t count)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(h, cudaMalloc(&q, bytes));
    CUDA_TRY(h, cudaMemsetAsync(q, 0, bytes, h->stream));
    h->ds_allocs.push_back(q);
    *p = (T*)q;
    return MCCBA_OK;
}
void ds_free(mccba_handle h)
{
    for (void* q : h->ds_allocs) cudaFree(q);
    h->ds_allocs.clear();
    h->ds_have = h->ds_have_params = false;
}
// AUTO precision policy (resolve_precision) at the composed poses of the double-sided problem
int ds_resolve_precision(mccba_handle h)
{
    if (h->prec != MCCBA_PRECISION_AUTO) return MCCBA_OK;
    ds_pose_kernel<<<(h->P.n_slots + 127) / 128, 128, 0, h->stream>>>(h->P, h->D, 1);
    return resolve_precision(h);
}
// one pass at the current parameters: composed poses, per-edge blocks (the rig path's residual / Jacobian kernel), Schur records
int ds_enqueue_eval(mccba_handle h, int forced)
{
    Problem& P = h->P;
    DsProblem& D = h->D;
    cudaStream_t s = h->stream;
    const int nb = (P.n_slots + 127) / 128;
    ds_pose_kernel<<<nb, 128, 0, s>>>(P, D, forced);
    launch_resid(h, s, 1);                       // evaluates Problem::erec, writes blocks[st->cur]
    ds_schur_kernel<<<nb, 128, 0, s>>>(P, D, forced);
    ds_reduce_kernel<<<kDsRec, 256, 0, s>>>(P, D, forced);
    CUDA_TRY(h, cudaGetLastError());
    return MCCBA_OK;
}
}  // namespace

extern "C" {

int mccba_ds_set_problem(mccba_handle h, const unsigned char* edge_back, const double* cam_pose)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->have_obs) return fail(h, MCCBA_ERR_STATE, "ds_set_problem before set_observations");
    if (!edge_back || !cam_pose) return fail(h, MCCBA_ERR_ARG, "ds_set_problem: null argument");
    if (h->opts.nranks > 1) return fail(h, MCCBA_ERR_ARG, "ds_set_problem: the double-sided path runs on one rank");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    ds_free(h);
    Problem& P = h->P;
    DsProblem& D = h->D;
    memset(&D, 0, sizeof(D));
    std::vector<unsigned char> back((size_t)P.n_edge_int, 0);
    int n_back = 0;
    for (int e = 0; e < h->n_edge; ++e) {
        if (edge_back[e] > 1) return fail(h, MCCBA_ERR_ARG, "ds_set_problem: edge_back[%d] = %d (0 = front, 1 = back)", e, (int)edge_back[e]);
        back[(size_t)h->int_of_edge[e]] = edge_back[e];
        n_back += edge_back[e];
    }
    if (n_back == 0) return fail(h, MCCBA_ERR_ARG, "ds_set_problem: no edge sees the back pattern (the front<->back transform is not observable)");
    std::vector<double> R((size_t)9 * h->n_cam), t((size_t)3 * h->n_cam);
    for (int c = 0; c < h->n_cam; ++c) {
        for (int i = 0; i < 6; ++i)
            if (!std::isfinite(cam_pose[6 * c + i])) return fail(h, MCCBA_ERR_ARG, "ds_set_problem: camera %d pose is not finite", c);
        rodrigues(cam_pose + 6 * c, R.data() + 9 * c);
        for (int i = 0; i < 3; ++i) t[(size_t)3 * c + i] = cam_pose[6 * c + 3 + i];
    }
    unsigned char* d_back = nullptr; double *d_R = nullptr, *d_t = nullptr;
    int rc;
    if ((rc = ds_alloc(h, &d_back, back.size()))) return rc;
    if ((rc = ds_alloc(h, &d_R, R.size()))) return rc;
    if ((rc = ds_alloc(h, &d_t, t.size()))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(d_back, back.data(), back.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_R, R.data(), R.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(d_t, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    D.back = d_back; D.cam_R = d_R; D.cam_t = d_t;
    D.n_blocks = (P.n_slots + 127) / 128;
    if ((rc = ds_alloc(h, &D.par, 6 + 6 * (size_t)h->n_frame))) return rc;
    if ((rc = ds_alloc(h, &D.rec, (size_t)kDsRec * (size_t)P.n_slots))) return rc;
    if ((rc = ds_alloc(h, &D.save, (size_t)kDsSave * (size_t)P.n_slots))) return rc;
    if ((rc = ds_alloc(h, &D.tot, kDsRec))) return rc;
    if ((rc = ds_alloc(h, &D.norm_part, 2 * (size_t)D.n_blocks))) return rc;
    if ((rc = ds_alloc(h, &D.st, 1))) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->ds_have = true;
    return MCCBA_OK;
}

int mccba_ds_set_parameters(mccba_handle h, int64_t n, const double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->ds_have) return fail(h, MCCBA_ERR_STATE, "ds_set_parameters before ds_set_problem");
    if (!params || n != 6 + 6 * (int64_t)h->n_frame) return fail(h, MCCBA_ERR_ARG, "ds_set_parameters: expected %lld values", (long long)(6 + 6 * (int64_t)h->n_frame));
    for (int64_t i = 0; i < n; ++i)
        if (!std::isfinite(params[i])) return fail(h, MCCBA_ERR_ARG, "ds_set_parameters: parameter %lld is not finite", (long long)i);
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(h->D.par, params, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->ds_have_params = true;
    return MCCBA_OK;
}

int mccba_ds_get_parameters(mccba_handle h, int64_t n, double* params)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->ds_have_params) return fail(h, MCCBA_ERR_STATE, "ds_get_parameters before ds_set_parameters");
    if (!params || n != 6 + 6 * (int64_t)h->n_frame) return fail(h, MCCBA_ERR_ARG, "ds_get_parameters: expected %lld values", (long long)(6 + 6 * (int64_t)h->n_frame));
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    CUDA_TRY(h, cudaMemcpyAsync(params, h->D.par, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MCCBA_OK;
}

int mccba_ds_normal(mccba_handle h, double* S36, double* g6, double* cost)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->ds_have_params) return fail(h, MCCBA_ERR_STATE, "ds_normal before ds_set_parameters");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    int rc;
    if ((rc = sync_state_cur(h))) return rc;
    if ((rc = ds_resolve_precision(h))) return rc;
    if ((rc = ds_enqueue_eval(h, 1))) return rc;
    double tot[kDsRec];
    CUDA_TRY(h, cudaMemcpyAsync(tot, h->D.tot, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (S36) memcpy(S36, tot, 36 * sizeof(double));
    if (g6) memcpy(g6, tot + 36, 6 * sizeof(double));
    if (cost) *cost = tot[42];
    if (tot[43] != 0.0) return fail(h, MCCBA_ERR_NUMERIC, "ds_normal: a frame block is not positive definite");
    return MCCBA_OK;
}

int mccba_ds_solve(mccba_handle h, int crit_type, int max_count, double epsilon, mccba_report* rep)
{
    if (!h) return MCCBA_ERR_ARG;
    if (!h->ds_have_params) return fail(h, MCCBA_ERR_STATE, "ds_solve before ds_set_parameters");
    if (crit_type < 1 || crit_type > 3 || max_count < 0) return fail(h, MCCBA_ERR_ARG, "ds_solve: bad criteria");
    CUDA_TRY(h, cudaSetDevice(h->opts.device));
    Problem& P = h->P;
    DsProblem& D = h->D;
    cudaStream_t s = h->stream;
    int rc;
    CUDA_TRY(h, cudaEventRecord(h->ev0, s));
    if ((rc = sync_state_cur(h))) return rc;
    if ((rc = ds_resolve_precision(h))) return rc;
    ds_init_state_kernel<<<1, 1, 0, s>>>(D.st, crit_type, max_count, epsilon);
    const int64_t max_launches = (crit_type & 1) ? max_count : 200000;
    int64_t launched = 0;
    bool stop = false;
    while (launched < max_launches && !stop) {
        const int nl = (int)std::min<int64_t>(8, max_launches - launched);
        for (int i = 0; i < nl; ++i) {
            if ((rc = ds_enqueue_eval(h, 0))) return rc;
            ds_solve_kernel<<<1, 32, 0, s>>>(P, D);
            ds_update_kernel<<<D.n_blocks, 128, 0, s>>>(P, D);
            ds_decide_kernel<<<1, 256, 0, s>>>(D);
        }
        launched += nl;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_done, &D.st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(h, cudaStreamSynchronize(s));
        if (h->h_done[0]) stop = true;
    }
    // cost at the returned parameters
    if ((rc = ds_enqueue_eval(h, 1))) return rc;
    DsState hs;
    double cost = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&hs, D.st, sizeof(DsState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(&cost, D.tot + 42, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaEventRecord(h->ev1, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    CUDA_TRY(h, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (rep) {
        memset(rep, 0, sizeof(*rep));
        rep->iterations = hs.iter; rep->accepted = hs.iter; rep->status = hs.status;
        rep->graph_launches = 0; rep->kernel_launches = (int)launched * 7 + 6;
        rep->change = hs.change; rep->cost = cost; rep->device_ms = ms;
    }
    if (hs.status) return fail(h, MCCBA_ERR_NUMERIC, "ds_solve: numeric failure at iteration %d (non-finite cost or a block that is not positive definite)", hs.iter);
    if (!hs.done) return fail(h, MCCBA_ERR_NUMERIC, "ds_solve: launch budget exhausted (iter %d)", hs.iter);
    return MCCBA_OK;
}

}  // extern "C"
