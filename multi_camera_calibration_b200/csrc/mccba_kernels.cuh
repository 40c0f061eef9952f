// mccba_kernels.cuh -- sm_100a kernels of the calibration bundle adjustment (included once by mccba_capi.cu).
//
// Kernel map (reference code each one replaces; all paths relative to /root/reference):
//   gather_obs_kernel        layout change only: AoS CV_32F points -> SoA planes in processing order
//   vertex_prep_kernel       cv::Rodrigues per vertex (src/multicalib.cpp:1023-1024 inside compose_motion)
//   resid_jac_accum_kernel   computePhotoCameraJacobian + scatter + J^T J / J^T E, per edge
//                            (src/multicalib.cpp:611-678, 688-689, 717-824; src/omnidir.cpp:141-244), FP64 policy
//   resid_jac_accum_f32_kernel   the same pass on the packed observation layout: two corners per lane in f32x2 registers,
//                            fp64 residual under the default MIXED policy (mccba_f32x2.cuh)
//   gather_obs_packed_kernel, reproj_error_packed_kernel   layout change / computeProjectError for that layout
//   reproj_error_kernel      computeProjectError (src/multicalib.cpp:912-984)
//   frame_schur_kernel       (new math) eliminates the per-frame pattern-pose blocks; replaces the P x P solve
//   reduce_records_kernel    deterministic, atomic-free sum of the warp records into the reduced camera system
//   p2p_exchange_kernel      (no counterpart) sum of the packed reduced system over the ranks through NVLink peer memory
//   decide_kernel            optimizeExtrinsics' loop control (src/multicalib.cpp:473-507) + accept/reject (LM)
//   chol_bcr_kernel          block cyclic reduction of the block-banded reduced system (replaces Eigen CG, :565-592)
//   chol_dag_kernel          one-launch tiled Cholesky of a dense reduced system (same)
//   camera_update_body       camera part of the update (:491-504), fused into the tail of the two solvers
//   frame_update_kernel      back-substitution + parameter update (:491-504)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mccba_math.cuh"
#include "mccba_f32x2.cuh"
#include "mccba_dense.cuh"
#include "mccba_bcr.cuh"

namespace mccba {

// Y = L^-1 W records as float for the non-fp64 policies: measured slower (frame_update 30 -> 38 us: the conversions cost more
// than the halved tile traffic saves, the tiles come out of L2 anyway), so off.
constexpr bool kYFloat = false;
constexpr int kLanesPerEdge = 8;
constexpr int kK1Threads = 256;
constexpr int kEdgesPerBlock = kK1Threads / kLanesPerEdge;  // 32
constexpr int kK2Threads = 128;
constexpr int kK4Threads = 128;
constexpr int kK5Threads = 512;
constexpr int kMaxViews = 64;
constexpr unsigned kFull = 0xffffffffu;

enum Phase { kPhaseFirst = 0, kPhaseDecide = 1, kPhaseRebuild = 2 };

// Loop state, resident in global memory; written by decide_kernel / camera_update_kernel only (single thread).
struct DevState {
    int mode, crit_type, max_count;
    int iter, done, phase, cur, status;
    int n_accept, n_reject, launches, solved;  // solved: the last iteration produced a step (update/eval may run)
    int go, chol_fail;                         // go: decide_kernel asks for a solve in this iteration
    double eps, lambda, lambda_up, lambda_down, lambda_spec;
    double cost_cur, cost_trial, change, alpha;
    double cam_step2, cam_param2;  // camera part of |step|^2 and |params|^2 of the current trial
};

// Per-edge record consumed by the residual kernel: composed pose of the edge at the point being evaluated + its
// camera and corner range.  128 bytes, fetched 32 at a time by one TMA bulk copy.
struct EdgeRec {          // dynamic part, rewritten for every evaluated point: 96 bytes
    double R3[9];
    double T3[3];
};
struct EdgeMeta {         // static part, written once by set_observations: 16 bytes
    int cam, begin, end, pad;
};
static_assert(sizeof(EdgeRec) == 96 && sizeof(EdgeMeta) == 16, "edge record layout");

struct Problem {
    int n_cam, n_frame, n_vertex, ns;
    int n_edge_int;   // internal edges (padded per group to 32-frame warps), multiple of 32
    int n_slots;      // frame slots = 32 * n_warps
    int n_warps;      // frame_schur warps
    int n_dest;       // reduce_records destinations
    int n_k4_blocks;
    int64_t n_param;  // 6 * (n_vertex - 1)
    // observations, SoA planes in internal edge order
    const float *ox, *oy, *oz, *iu, *iv;
    const int* e_off;    // n_edge_int + 1 corner offsets
    const int* e_cam;    // camera of internal edge
    const int* e_frame;  // frame (0-based) of internal edge, -1 for padding
    const CamParams* cams;
    // frame slots / groups
    const int* slot_frame;   // n_slots, -1 = padding
    const int* warp_group;   // n_warps
    const int* group_V;      // per group: number of views
    const int* group_cam0;   // per group: offset into group_cams
    const int* group_cams;   // concatenated sorted camera lists
    const int* group_ebase;  // per group: first internal edge
    const int* group_stride; // per group: slots (multiple of 32); edge(v, ls) = ebase + v*stride + ls
    const int* group_slot0;  // per group: first slot
    const int* warp_rec;     // per warp: offset of its record (in doubles)
    const int4* wmeta;       // per warp, 2 x int4: {V, first index into group_cams, ebase, stride}, {first local slot, record
                             // offset, group, 0} -- everything a frame warp needs in ONE load instead of a three-level chain
    // reduce destinations
    const int* dest_info;    // n_dest x 4: kind (0 block, 1 g, 2 scalars), A (cam-1), B (cam-1), unused
    const int* dest_src0;    // n_dest + 1
    const int* dest_src;     // record offsets (doubles)
    // state and work buffers
    DevState* st;
    double* x[2];        // parameter vectors (reference layout)
    double* vR[2];       // n_vertex x 9 rotation matrices
    double* blocks[2];   // kBlk x n_edge_int (SoA)
    double* frameL;      // 27 x n_slots (21 packed factor + 6 z)
    double* edgeY;       // 36 x n_edge_int: Y = L^-1 W per edge, tile-major; stored as FLOAT when prec != 0 (Y only shapes the step of
                         // the pattern poses, never the fixed point, and it is the largest record the two per-frame kernels exchange)
    double* records;     // warp records
    const float* edge_extent;   // n_edge_int: diameter of the object points of every edge (AUTO precision policy), else null
    double* warp_scal;   // 2 x n_warps: [cost | bad] of every Schur warp, contiguous so that the scalar reduction is coalesced
    double* ar;          // allreduce buffer: S (ns x ns) | g (ns) | cost, step2, param2, bad
    double* dc;          // ns: camera step in tangent coordinates
    // peer-memory exchange of the packed reduced system (N > 1, see p2p_exchange_kernel): every rank's window is
    // [header | parity 0: n ranks x p2p_stride doubles | parity 1: ...] (a slot holds 2 words per element, LL protocol),
    // mapped into every process by CUDA IPC
    double* p2p_peer[8];
    unsigned long long* p2p_epoch;   // local: number of completed exchanges
    int p2p_n, p2p_rank;
    int64_t p2p_stride;
    double* ar_part;     // rank-local packed buffer written by reduce_records (== ar on a single rank)
    int band_nw;         // 0: ar = [S (ns x ns) | g | 4 scalars]; NW > 0: ar = [band (ns x NW, band[r][c - r + NW - 1], lower
                         // triangle only) | g | 4 scalars] -- what the banded solver reads and all the collective has to move
    int64_t ar_goff;     // offset of g in ar (ns * ns or ns * band_nw); the scalars follow at ar_goff + ns
    int64_t dag_words;   // chol_dag_words(ns) when the tile DAG is in use, else 0
    double* dag_buf;     // tile-DAG output (chol_dag_words doubles), filled with the all-ones sentinel before every factorisation
    double* norm_part;   // 2 x n_k4_blocks: per-block |step|^2, |trial|^2 of the frames
    // packed single-precision pass (precision policies 1 and 2, mccba_f32x2.cuh): observations as corner PAIRS in the order
    // the lanes consume them -- tile (32 edges) -> quarter (8 edges) -> step k -> plane (x y z u v) -> lane -> 2 floats, so
    // a warp reads five contiguous 256-byte lines per step
    int prec;                // 0: fp64 per-corner arithmetic (SoA planes); 1: packed f32x2 Jacobian + fp64 residual; 2: all f32x2
    const float2* obs2;
    const int* obs_nonplanar; // 1 if any object point of the packed layout has z != 0 (set by gather_obs_packed_kernel); flat boards skip the z terms
    const int64_t* tile_off; // per tile: first float2 of its block
    const int* tile_kp;      // per tile: steps per quarter = ceil(max corners per edge / 8)
    EdgeRec* erec;       // n_edge_int composed poses of the point the residual kernel evaluates next
    const EdgeMeta* emeta; // n_edge_int camera / corner range of every internal edge
    double* err_sq;      // n_edge_int
    double* err_nrm;     // n_edge_int
};

// Tile-major record layout: the K values of 32 consecutive edges (frame slots) form one contiguous tile [k][32], so
// everything a warp of 32 frames touches per view is a single K x 256-byte block (one TMA/bulk-copy sized unit)
// instead of K segments spread over K planes.
__host__ __device__ __forceinline__ int64_t tile_idx(int K, int64_t e, int k) { return ((e >> 5) * K + k) * 32 + (e & 31); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// --------------------------------------------------------------------------------------------------------
// observations: AoS host layout -> SoA planes in internal edge order.  One warp per internal edge.
// --------------------------------------------------------------------------------------------------------
__global__ void gather_obs_kernel(int n_edge_int, const int* __restrict__ e_off, const int64_t* __restrict__ e_src,
                                  const float* __restrict__ obj, const float* __restrict__ img, float* ox, float* oy,
                                  float* oz, float* iu, float* iv)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < n_edge_int; e += nw) {
        const int b = e_off[e], n = e_off[e + 1] - b;
        const int64_t s = e_src[e];
        for (int i = lane; i < n; i += 32) {
            ox[b + i] = obj[3 * (s + i)];
            oy[b + i] = obj[3 * (s + i) + 1];
            oz[b + i] = obj[3 * (s + i) + 2];
            iu[b + i] = img[2 * (s + i)];
            iv[b + i] = img[2 * (s + i) + 1];
        }
    }
}

// AUTO precision policy: diameter of the object points of every edge (twice the largest distance from their centroid) ...
__global__ void edge_extent_kernel(int n_edge_int, const int* __restrict__ e_off, const int64_t* __restrict__ e_src,
                                   const float* __restrict__ obj, float* __restrict__ extent)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < n_edge_int; e += nw) {
        const int n = e_off[e + 1] - e_off[e];
        const int64_t s = e_src[e];
        float cx = 0, cy = 0, cz = 0;
        for (int i = lane; i < n; i += 32) { cx += obj[3 * (s + i)]; cy += obj[3 * (s + i) + 1]; cz += obj[3 * (s + i) + 2]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(kFull, cx, o); cy += __shfl_xor_sync(kFull, cy, o); cz += __shfl_xor_sync(kFull, cz, o);
        }
        const float inv = n > 0 ? 1.0f / (float)n : 0.0f;
        cx *= inv; cy *= inv; cz *= inv;
        float d2 = 0;
        for (int i = lane; i < n; i += 32) {
            const float dx = obj[3 * (s + i)] - cx, dy = obj[3 * (s + i) + 1] - cy, dz = obj[3 * (s + i) + 2] - cz;
            d2 = fmaxf(d2, dx * dx + dy * dy + dz * dz);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(kFull, d2, o));
        if (lane == 0) extent[e] = 2.0f * sqrtf(d2);
    }
}
// ... and the smallest angular extent  diameter / distance  over the live edges at the composed poses in Problem::erec
// (one partial minimum per block; the host takes the minimum of the partials)
__global__ void __launch_bounds__(256) min_angular_extent_kernel(Problem P, double* __restrict__ part)
{
    __shared__ double sm[8];
    double v = 1e300;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < P.n_edge_int; e += gridDim.x * blockDim.x) {
        const EdgeMeta em = P.emeta[e];
        if (em.end > em.begin) {
            const double* T = P.erec[e].T3;
            v = fmin(v, (double)P.edge_extent[e] / sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) v = fmin(v, sm[w]);
        part[blockIdx.x] = fmin(v, sm[0]);
    }
}

// --------------------------------------------------------------------------------------------------------
// vertex prep: R_v = exp([om_v]x) for every vertex of parameter buffer `which` (-1: the state's current buffer)
// --------------------------------------------------------------------------------------------------------
__global__ void vertex_prep_kernel(Problem P, int which)
{
    if (which < 0) which = P.st->cur;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.n_vertex) return;
    double R[9];
    if (v == 0) {
        R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
    } else {
        const double* p = P.x[which] + 6 * (int64_t)(v - 1);
        const double om[3] = {p[0], p[1], p[2]};
        rodrigues(om, R);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) P.vR[which][9 * (int64_t)v + i] = R[i];
}

// composed poses of the edges of one frame slot for parameter buffer `which` (R_p, t_p given): one record per view
__device__ __forceinline__ void write_edge_records(const Problem& P, int which, int slot, int frame, const double* Rp,
                                                   const double* tp)
{
    const int g = P.warp_group[slot >> 5];
    const int V = P.group_V[g];
    const int* gc = P.group_cams + P.group_cam0[g];
    const int ls = slot - P.group_slot0[g];
    const int64_t ebase = P.group_ebase[g], stride = P.group_stride[g];
    for (int v = 0; v < V; ++v) {
        const int c = gc[v];
        EdgeRec r;
        if (frame >= 0) {
            double Rc[9], tc[3];
            if (c == 0) {
                Rc[0] = Rc[4] = Rc[8] = 1; Rc[1] = Rc[2] = Rc[3] = Rc[5] = Rc[6] = Rc[7] = 0;
                tc[0] = tc[1] = tc[2] = 0;
            } else {
#pragma unroll
                for (int i = 0; i < 9; ++i) Rc[i] = P.vR[which][9 * c + i];
#pragma unroll
                for (int i = 0; i < 3; ++i) tc[i] = P.x[which][6 * (c - 1) + 3 + i];
            }
            compose_pose(Rc, tc, Rp, tp, r.R3, r.T3);
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) r.R3[i] = 0;
            r.T3[0] = r.T3[1] = 0;
            r.T3[2] = 1;      // padding slot: a benign pose (the packed pass evaluates its zero-weight corners)
        }
        P.erec[ebase + v * stride + ls] = r;
    }
}

// edge records for the forced evaluation of buffer `which` (-1: the state's current buffer); one thread per frame slot
__global__ void edge_pose_kernel(Problem P, int which)
{
    if (which < 0) which = P.st->cur;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.n_slots) return;
    const int frame = P.slot_frame[slot];
    double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tp[3] = {0, 0, 0};
    if (frame >= 0) {
        const int64_t pv = P.n_cam + frame;
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = P.vR[which][9 * pv + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) tp[i] = P.x[which][6 * (pv - 1) + 3 + i];
    }
    write_edge_records(P, which, slot, frame, Rp, tp);
}

// --------------------------------------------------------------------------------------------------------
// K1: residual + Jacobian + per-edge normal-equation block.  8 lanes per edge, 32 edges per CTA iteration,
// persistent grid-stride over edge chunks.  Intrinsics and camera poses staged in shared memory once per CTA.
// Output: blocks[which][k * n_edge_int + e], k < 28, written coalesced through a shared-memory transpose.
// --------------------------------------------------------------------------------------------------------
struct K1Shared {
    double stage[kBlk][kEdgesPerBlock + 1];
    EdgeRec erec[2][kEdgesPerBlock];   // TMA destination: composed poses of the two stages
    EdgeMeta emeta[2][kEdgesPerBlock]; // TMA destination: camera / corner range
    unsigned long long bar[2];         // mbarriers of the two stages
    int first[2];                      // first float index (16-byte aligned) held by each stage
};

// ---- TMA (1-D bulk copy) + mbarrier primitives, sm_90+/sm_100a PTX ---------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on the mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk copy (TMA engine): make generic-proxy writes to shared memory visible to the async proxy,
// issue, and wait until the source has been read (the CTA may then reuse or release the buffer)
__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int kModel, bool kRational>
__device__ __forceinline__ void edge_corner_loop(const float* __restrict__ ox, const float* __restrict__ oy,
                                                 const float* __restrict__ oz, const float* __restrict__ iu,
                                                 const float* __restrict__ iv, const CamParams& cam, const double* R3,
                                                 const double* T3, int begin, int end, int sub, double* acc)
{
    for (int i = begin + sub; i < end; i += kLanesPerEdge) {
        // The pose and the intrinsics live in shared memory; the barrier keeps the compiler from hoisting those ~25
        // loads out of the loop into registers it does not have (it spilled the accumulators to local memory instead:
        // 6 LDL/STL per corner).  Re-reading them per corner costs broadcast LDS only.
        asm volatile("" ::: "memory");
        corner_accumulate<kModel, kRational>(cam, R3, T3, ox[i], oy[i], oz[i], iu[i], iv[i], acc);
    }
}

// forced: 0 = loop launch (evaluate the trial buffer iff the state says a step was produced), 1 = forced on `cur`.
// Shared memory: K1Shared | camera table | 2 stages x 5 planes x obs_cap floats.
// Everything the kernel reads per chunk -- 32 edge records (composed pose, camera, corner range) and the five
// observation planes of those edges -- is fetched by the TMA engine (cp.async.bulk + mbarrier) one chunk ahead, so
// the arithmetic never waits on a dependent global load.  kStaged == false (a chunk does not fit a stage): the edge
// records are still staged, the observations are read straight from global memory.
template <bool kStaged>
__global__ void __launch_bounds__(kK1Threads, 2) resid_jac_accum_kernel(Problem P, int forced, int obs_cap)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    K1Shared* sh = reinterpret_cast<K1Shared*>(smem_raw);
    const size_t off_cam = (sizeof(K1Shared) + 15) & ~(size_t)15;
    CamParams* s_cam = reinterpret_cast<CamParams*>(smem_raw + off_cam);
    const size_t off_obs = (off_cam + (size_t)P.n_cam * sizeof(CamParams) + 127) & ~(size_t)127;
    float* s_obs = reinterpret_cast<float*>(smem_raw + off_obs);   // [2][5][obs_cap]

    const DevState* st = P.st;
    int which;
    if (forced) which = st->cur;
    else {
        if (st->done || !st->solved) return;
        which = 1 - st->cur;
    }
    double* __restrict__ out = P.blocks[which];

    for (int c = threadIdx.x; c < P.n_cam; c += blockDim.x) s_cam[c] = P.cams[c];
    if (threadIdx.x == 0) {
        mbar_init(&sh->bar[0], 1);
        mbar_init(&sh->bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int sub = threadIdx.x & (kLanesPerEdge - 1);
    const int eb = threadIdx.x / kLanesPerEdge;
    const int n_chunks = P.n_edge_int / kEdgesPerBlock;
    const float* planes[5] = {P.ox, P.oy, P.oz, P.iu, P.iv};

    // producer: one thread programs the TMA engine for chunk `chunk` into stage `stg`
    auto issue = [&](int chunk, int stg) {
        unsigned bytes = (unsigned)(kEdgesPerBlock * (sizeof(EdgeRec) + sizeof(EdgeMeta)));
        int f0 = 0, nfl = 0;
        if (kStaged) {
            f0 = P.e_off[chunk * kEdgesPerBlock] & ~3;
            nfl = (P.e_off[(chunk + 1) * kEdgesPerBlock] - f0 + 3) & ~3;
            bytes += 5u * (unsigned)nfl * 4u;
        }
        sh->first[stg] = f0;
        mbar_expect_tx(&sh->bar[stg], bytes);
        tma_load_1d(&sh->erec[stg][0], P.erec + (size_t)chunk * kEdgesPerBlock, (unsigned)(kEdgesPerBlock * sizeof(EdgeRec)), &sh->bar[stg]);
        tma_load_1d(&sh->emeta[stg][0], P.emeta + (size_t)chunk * kEdgesPerBlock, (unsigned)(kEdgesPerBlock * sizeof(EdgeMeta)), &sh->bar[stg]);
        if (kStaged && nfl > 0) {
#pragma unroll
            for (int pl = 0; pl < 5; ++pl)
                tma_load_1d(s_obs + ((size_t)stg * 5 + pl) * obs_cap, planes[pl] + f0, (unsigned)nfl * 4u, &sh->bar[stg]);
        }
    };

    if (threadIdx.x == 0 && blockIdx.x < n_chunks) issue(blockIdx.x, 0);
    int it = 0;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int stg = it & 1;
        if (threadIdx.x == 0 && chunk + (int)gridDim.x < n_chunks) issue(chunk + gridDim.x, stg ^ 1);
        double acc[kBlk];
#pragma unroll
        for (int k = 0; k < kBlk; ++k) acc[k] = 0.0;
        mbar_wait(&sh->bar[stg], (unsigned)((it >> 1) & 1));
        {
            const EdgeRec& er = sh->erec[stg][eb];
            const EdgeMeta em = sh->emeta[stg][eb];
            const int b = em.begin, n = em.end;
            if (n > b) {
                const float *ox, *oy, *oz, *iu, *iv;
                if (kStaged) {
                    const float* base = s_obs + (size_t)stg * 5 * obs_cap - sh->first[stg];
                    ox = base; oy = base + obs_cap; oz = base + 2 * obs_cap; iu = base + 3 * obs_cap; iv = base + 4 * obs_cap;
                } else {
                    ox = P.ox; oy = P.oy; oz = P.oz; iu = P.iu; iv = P.iv;
                }
                const CamParams& cam = s_cam[em.cam];
                if (cam.model == kPinhole) {
                    if (cam.rational) edge_corner_loop<kPinhole, true>(ox, oy, oz, iu, iv, cam, er.R3, er.T3, b, n, sub, acc);
                    else edge_corner_loop<kPinhole, false>(ox, oy, oz, iu, iv, cam, er.R3, er.T3, b, n, sub, acc);
                } else {
                    edge_corner_loop<kOmnidir, false>(ox, oy, oz, iu, iv, cam, er.R3, er.T3, b, n, sub, acc);
                }
            }
        }
        // transposed reduction over the 8 lanes of the edge: 28 -> 14 -> 7 -> 3(+1) values per lane, 25 adds
        int base = 0;
        {
            const bool up = (sub & 4) != 0;
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                const double a = acc[i], b2 = acc[i + 14];
                const double send = up ? a : b2, keep = up ? b2 : a;
                acc[i] = keep + __shfl_xor_sync(kFull, send, 4);
            }
            base += up ? 14 : 0;
        }
        {
            const bool up = (sub & 2) != 0;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const double a = acc[i], b2 = acc[i + 7];
                const double send = up ? a : b2, keep = up ? b2 : a;
                acc[i] = keep + __shfl_xor_sync(kFull, send, 2);
            }
            base += up ? 7 : 0;
        }
        {
            const bool up = (sub & 1) != 0;
            const double last = acc[6] + __shfl_xor_sync(kFull, acc[6], 1);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double a = acc[i], b2 = acc[i + 3];
                const double send = up ? a : b2, keep = up ? b2 : a;
                acc[i] = keep + __shfl_xor_sync(kFull, send, 1);
            }
            if (!up) sh->stage[base + 6][eb] = last;
            base += up ? 3 : 0;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) sh->stage[base + i][eb] = acc[i];
        __syncthreads();   // also: every thread is done reading stage `stg` (edge records and observations)
        for (int t = threadIdx.x; t < kBlk * kEdgesPerBlock; t += kK1Threads) {
            const int k = t / kEdgesPerBlock, col = t % kEdgesPerBlock;
            out[(int64_t)chunk * (kBlk * kEdgesPerBlock) + t] = sh->stage[k][col];   // tile-major: one contiguous 7 KB tile per chunk
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------------------
// K1, packed single precision (precision policy 1).  Unit of work = one quarter tile = 8 edges = one warp:
// lane = 4 * (edge within the quarter) + q, and in step k the lane evaluates corners 8k + 2q and 8k + 2q + 1 of its
// edge as ONE f32x2 pair (mccba_f32x2.cuh).  The observations are stored in exactly that order (Problem::obs2), so a
// step is one contiguous 1280-byte block that the TMA engine drops into the warp's private ring several steps ahead --
// no block-wide barrier, no atomics; a warp never waits for another.  At the end of the edge the two halves and the four lanes are
// summed in float (transposed butterfly, 21 shuffles), promoted to double once, and stored tile-major like the fp64
// kernel's output (what frame_schur_kernel reads).  kErr: residual-only variant for computeProjectError in double.
// --------------------------------------------------------------------------------------------------------
constexpr int kF32Threads = 128;

__device__ __forceinline__ float2 ldg_f2(const float2* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}


// The stream of observation steps a warp consumes, across tile boundaries.  A step is one contiguous 1280-byte block
// (5 planes x 32 lanes x 2 floats), so ONE lane programs the TMA engine (cp.async.bulk + mbarrier) to drop it into the
// warp's private ring in shared memory kObsStages chunks ahead of its use; when the current tile runs out the next tile
// of this CTA (tile + gridDim.x) takes over without a bubble -- its offset and step count were fetched one tile earlier.
// (First version: register double buffer filled by LDG.  The compiler rotated the buffer at the END of a step, the
// prefetch distance collapsed to < 1 step and a third of all stall samples sat on that move waiting for DRAM --
// profiles/r2_k1_f32_v1.txt.)
// One copy moves a CHUNK of up to kObsChunk consecutive steps (they are contiguous in the packed layout) into one ring stage:
// the producer code (address arithmetic, expect-tx, the bulk copy: ~45 instructions that the whole warp sits through
// while lane 0 executes them) and the mbarrier wait run once per chunk instead of once per step (profiles/r2_experiments_late.md).
#ifndef MCCBA_OBS_STAGES
#define MCCBA_OBS_STAGES 2
#endif
#ifndef MCCBA_OBS_CHUNK
#define MCCBA_OBS_CHUNK 4
#endif
#ifndef MCCBA_F32_MINBLOCKS
#define MCCBA_F32_MINBLOCKS 3
#endif
#ifndef MCCBA_F32_RELOAD
#define MCCBA_F32_RELOAD 1
#endif
#ifndef MCCBA_F32_POSE_REGS
#define MCCBA_F32_POSE_REGS 0
#endif
constexpr int kObsStages = MCCBA_OBS_STAGES;   // power of two: stage and phase are a mask and a shift
constexpr int kObsChunk = MCCBA_OBS_CHUNK;     // steps per copy
constexpr int kObsStepBytes = 5 * 32 * 8;
constexpr int kObsStageF2 = kObsChunk * 5 * 32;   // float2 elements of one ring stage
struct ObsStream {
    const float2* base;     // current tile: first float2 of this warp's block
    int k, kp, tile;        // next step to load, steps of the current tile, current tile
    int kp_next;            // prefetched for tile + stride
    int64_t off_next;
    int issued;             // chunks issued so far (stage = issued % kObsStages)
};
__device__ __forceinline__ void obs_stream_open(ObsStream& S, const Problem& P, int tile, int stride, int n_tiles, int wq)
{
    S.tile = tile; S.k = 0; S.issued = 0;
    S.kp = tile < n_tiles ? P.tile_kp[tile] : 0;
    S.base = P.obs2 + (tile < n_tiles ? P.tile_off[tile] : 0) + (size_t)wq * S.kp * 5 * 32;
    const int nt = tile + stride;
    S.kp_next = nt < n_tiles ? P.tile_kp[nt] : 0;
    S.off_next = nt < n_tiles ? P.tile_off[nt] : 0;
}
// called by ONE lane of the warp: issue the copy of the next chunk of the stream (chunks do not span tiles: the consumer
// cuts a tile into the same min(kObsChunk, kp - k) pieces) into its ring stage
__device__ __forceinline__ void obs_stream_issue(ObsStream& S, const Problem& P, int stride, int n_tiles, int wq, float2* ring,
                                                 unsigned long long* bars)
{
    if (S.tile >= n_tiles) return;
    const int stg = S.issued & (kObsStages - 1);
    const int nst = min(kObsChunk, S.kp - S.k);
    const unsigned bytes = (unsigned)nst * kObsStepBytes;
    mbar_expect_tx(&bars[stg], bytes);
    tma_load_1d(ring + (size_t)stg * kObsStageF2, S.base + (size_t)S.k * 5 * 32, bytes, &bars[stg]);
    ++S.issued;
    S.k += nst;
    if (S.k == S.kp) {    // on to the next tile of this CTA; fetch the layout of the one after it
        S.tile += stride; S.k = 0; S.kp = S.kp_next;
        S.base = P.obs2 + S.off_next + (size_t)wq * S.kp * 5 * 32;
        const int nt = S.tile + stride;
        S.kp_next = nt < n_tiles ? P.tile_kp[nt] : 0;
        S.off_next = nt < n_tiles ? P.tile_off[nt] : 0;
    }
}

// Per-edge pose of the packed pass in shared memory, one record per edge of the warp's quarter tile: the double pose
// (for the MIXED policy's residual) and the float pose (for the f32x2 Jacobian; broadcast operands, see CamF2).  36-word
// stride: the 8 edges of a warp land in distinct banks.
struct alignas(16) PackedPose {
    double Rd[9], Td[3];
    float Rf[9], Tf[3];
};
__host__ __device__ inline size_t f32_smem_bytes(int n_cam)
{
    return (size_t)4 * kObsStages * kObsChunk * kObsStepBytes + 4 * kObsStages * 8 + 32 * sizeof(PackedPose) + (sizeof(CamF2) + sizeof(CamResid)) * (size_t)n_cam;
}

// One step of one lane: two corners.  kMixed: the residual pair comes from a double projection (corner_residual) and the
// cost is summed in double (g[6]); Jacobian, J^T J and J^T e in f32x2 (acc[0..26]).  Else everything in f32x2 (acc[0..27]).
template <int kModel, bool kRational, bool kMixed, bool kMasked, bool kPlanar>
__device__ __forceinline__ void packed_step(const float2 (&cur)[5], int k, int n, int q, const CamF2& cam,
                                            const CamResid& camd, const PackedPose& pose, f2* acc, double* g)
{
    // kMasked: some lane of the warp runs out of corners in this step (k >= k_full, warp-uniform: the last step of an edge);
    // the full steps carry no weights and no selects at all
    const int c0 = 8 * k + 2 * q;
    const f2 w = kMasked ? f2_make(c0 < n ? 1.0f : 0.0f, c0 + 1 < n ? 1.0f : 0.0f) : f2_dup(1.0f);
    f2 e0 = f2_dup(0.0f), e1 = f2_dup(0.0f);
    if (kMixed) {
        double ea[2], eb[2];
        corner_residual<kModel, kRational, kPlanar>(camd, pose.Rd, pose.Td, cur[0].x, cur[1].x, cur[2].x, cur[3].x, cur[4].x, ea);
        corner_residual<kModel, kRational, kPlanar>(camd, pose.Rd, pose.Td, cur[0].y, cur[1].y, cur[2].y, cur[3].y, cur[4].y, eb);
        e0 = f2_make((float)ea[0], (float)eb[0]);
        e1 = f2_make((float)ea[1], (float)eb[1]);
        // the cost the accept / reject test compares is summed in double from the exact residuals
        const double ca = fma(ea[0], ea[0], ea[1] * ea[1]), cb = fma(eb[0], eb[0], eb[1] * eb[1]);
        if (kMasked) g[6] += (c0 < n ? ca : 0.0) + (c0 + 1 < n ? cb : 0.0);
        else g[6] += ca + cb;
    }
    corner_pair_accumulate<kModel, kRational, kMixed, kPlanar>(cam, pose.Rf, pose.Tf, f2_make(cur[0].x, cur[0].y), f2_make(cur[1].x, cur[1].y),
                                                      f2_make(cur[2].x, cur[2].y), f2_make(cur[3].x, cur[3].y),
                                                      f2_make(cur[4].x, cur[4].y), w, kMasked, acc, e0, e1);
}

template <int kModel, bool kRational, bool kExactE, bool kPlanar>
__device__ __forceinline__ void packed_edge_loop(ObsStream& S, int& consumed, float2* ring, unsigned long long* bars, const Problem& P,
                                                 int stride, int n_tiles, int wq, int lane, int kp, int n, int q, int k_full,
                                                 const CamF2& cam, const CamResid& camd, const PackedPose& pose, f2* acc, double* cost)
{
    for (int k = 0; k < kp;) {
        const int nst = min(kObsChunk, kp - k);
        const int stg = consumed & (kObsStages - 1);
        mbar_wait(&bars[stg], (unsigned)((consumed / kObsStages) & 1));
        const float2* src = ring + (size_t)stg * kObsStageF2 + lane;
        for (int j = 0; j < nst; ++j, ++k) {
            float2 cur[5];
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) cur[pl] = src[(j * 5 + pl) * 32];
            if (j == nst - 1) {                     // every lane has taken the last pairs out of the stage: refill it
                ++consumed;
                __syncwarp();
                if (lane == 0) obs_stream_issue(S, P, stride, n_tiles, wq, ring, bars);
            }
#if MCCBA_F32_RELOAD
            asm volatile("" ::: "memory");     // pose and intrinsics are re-read from shared memory per step, not kept in ~90 registers
#endif
            if (k < k_full) packed_step<kModel, kRational, kExactE, false, kPlanar>(cur, k, n, q, cam, camd, pose, acc, cost);
            else packed_step<kModel, kRational, kExactE, true, kPlanar>(cur, k, n, q, cam, camd, pose, acc, cost);
        }
    }
}

// forced: as resid_jac_accum_kernel.  Shared memory: f32_smem_bytes(n_cam).
template <bool kExactE>
__global__ void __launch_bounds__(kF32Threads, MCCBA_F32_MINBLOCKS) resid_jac_accum_f32_kernel(Problem P, int forced)
{
    extern __shared__ __align__(128) unsigned char f32_smem[];
    float2* s_ring = reinterpret_cast<float2*>(f32_smem);                                   // [4 warps][kObsStages][kObsChunk][5][32]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_ring + 4 * kObsStages * kObsStageF2);   // [4][kObsStages]
    PackedPose* s_pose = reinterpret_cast<PackedPose*>(s_bar + 4 * kObsStages);            // [32]
    CamF2* s_cam = reinterpret_cast<CamF2*>(s_pose + 32);
    CamResid* s_camd = reinterpret_cast<CamResid*>(s_cam + P.n_cam);
    const DevState* st = P.st;
    int which;
    if (forced) which = st->cur;
    else {
        if (st->done || !st->solved) return;
        which = 1 - st->cur;
    }
    double* __restrict__ out = P.blocks[which];
    const bool planar = *P.obs_nonplanar == 0;
    for (int c = threadIdx.x; c < P.n_cam; c += blockDim.x) {
        const CamParams cp = P.cams[c];
        s_camd[c] = make_cam_resid(cp);
        s_cam[c] = make_cam_f2(cp);
    }
    const int lane = threadIdx.x & 31, q = lane & 3, el = lane >> 2, wq = threadIdx.x >> 5;
    float2* ring = s_ring + (size_t)wq * kObsStages * kObsStageF2;
    unsigned long long* bars = s_bar + wq * kObsStages;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kObsStages; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one CTA = 4 warps = one tile per pass; the loop runs on the tile index, which is uniform over the CTA, so the warp
    // collectives below sit in provably convergent control flow (plain SHFL / REDUX, no WARPSYNC wrappers)
    const int n_tiles = P.n_edge_int >> 5, stride = gridDim.x;
    if ((int)blockIdx.x >= n_tiles) return;
    ObsStream S;        // producer state: meaningful in lane 0 only
    obs_stream_open(S, P, blockIdx.x, stride, n_tiles, wq);
    if (lane == 0) {
        for (int d = 0; d < kObsStages; ++d) obs_stream_issue(S, P, stride, n_tiles, wq, ring, bars);
    }
    int consumed = 0;
    // per-edge record of the first tile: the 4 lanes of an edge share the load (lane q takes doubles q, q + 4, q + 8)
    EdgeMeta em = P.emeta[(blockIdx.x << 5) + (wq << 3) + el];
    double rq[3];
    {
        const double* r = reinterpret_cast<const double*>(P.erec + ((blockIdx.x << 5) + (wq << 3) + el));
#pragma unroll
        for (int i = 0; i < 3; ++i) rq[i] = r[q + 4 * i];
    }
    int kp_cur = P.tile_kp[blockIdx.x];
    for (int tile = blockIdx.x; tile < n_tiles; tile += stride) {
        const int n = em.end - em.begin, cam_idx = em.cam;
        const int kp = kp_cur;
        const int k_full = __reduce_min_sync(kFull, n) >> 3;
        PackedPose& pose = s_pose[(wq << 3) + el];
        __syncwarp();                       // the previous tile's reads of this warp's pose records are over
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int j = q + 4 * i;
            if (j < 9) { pose.Rd[j] = rq[i]; pose.Rf[j] = (float)rq[i]; }
            else { pose.Td[j - 9] = rq[i]; pose.Tf[j - 9] = (float)rq[i]; }
        }
        __syncwarp();
        {   // the next tile's records travel while this one is evaluated
            const int nt = tile + stride;
            if (nt < n_tiles) {
                const int e2 = (nt << 5) + (wq << 3) + el;
                kp_cur = P.tile_kp[nt];
                em = P.emeta[e2];
                const double* r = reinterpret_cast<const double*>(P.erec + e2);
#pragma unroll
                for (int i = 0; i < 3; ++i) rq[i] = r[q + 4 * i];
            }
        }
        f2 acc[kBlk];
#pragma unroll
        for (int k = 0; k < kBlk; ++k) acc[k] = f2_dup(0.0f);
        double cost[7] = {0, 0, 0, 0, 0, 0, 0};      // MIXED: [6] = sum |e|^2 in double
#if MCCBA_F32_POSE_REGS
        PackedPose pose_r;      // the edge's pose in registers for the whole edge (36 registers, no reloads)
#pragma unroll
        for (int i = 0; i < 9; ++i) { pose_r.Rd[i] = pose.Rd[i]; pose_r.Rf[i] = pose.Rf[i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { pose_r.Td[i] = pose.Td[i]; pose_r.Tf[i] = pose.Tf[i]; }
#define MCCBA_POSE pose_r
#else
#define MCCBA_POSE pose
#endif
        const CamF2& cam = s_cam[cam_idx];      // the 32 edges of a tile are one (group, view) run: one camera per warp
        const CamResid& camd = s_camd[cam_idx];
#define MCCBA_EDGE_LOOP(MODEL, RATIONAL, PLANAR)                                                                                   \
    packed_edge_loop<MODEL, RATIONAL, kExactE, PLANAR>(S, consumed, ring, bars, P, stride, n_tiles, wq, lane, kp, n, q, k_full, cam, \
                                                       camd, MCCBA_POSE, acc, cost)
        if (planar) {       // flat boards (every object z is 0): the z column of the pose drops out, bit-identically
            if (cam.model == kPinhole) {
                if (cam.rational) MCCBA_EDGE_LOOP(kPinhole, true, true);
                else MCCBA_EDGE_LOOP(kPinhole, false, true);
            } else MCCBA_EDGE_LOOP(kOmnidir, false, true);
        } else {
            if (cam.model == kPinhole) {
                if (cam.rational) MCCBA_EDGE_LOOP(kPinhole, true, false);
                else MCCBA_EDGE_LOOP(kPinhole, false, false);
            } else MCCBA_EDGE_LOOP(kOmnidir, false, false);
        }
#undef MCCBA_EDGE_LOOP
        // halves, then the 4 lanes of the edge: 28 -> 14 -> 7 values per lane (transposed butterfly)
        float v[kBlk];
#pragma unroll
        for (int k = 0; k < kBlk; ++k) v[k] = f2_lo(acc[k]) + f2_hi(acc[k]);
        int kbase = 0;
        {
            const bool up = (q & 2) != 0;
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                const float send = up ? v[i] : v[i + 14], keep = up ? v[i + 14] : v[i];
                v[i] = keep + __shfl_xor_sync(kFull, send, 2);
            }
            kbase += up ? 14 : 0;
        }
        {
            const bool up = (q & 1) != 0;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const float send = up ? v[i] : v[i + 7], keep = up ? v[i + 7] : v[i];
                v[i] = keep + __shfl_xor_sync(kFull, send, 1);
            }
            kbase += up ? 7 : 0;
        }
        double* o = out + ((size_t)tile * kBlk + kbase) * 32 + (wq << 3) + el;
#pragma unroll
        for (int i = 0; i < 7; ++i) o[(size_t)i * 32] = (double)v[i];
        if (kExactE) {   // cost (element 27: the last of lane q == 3) from the double sum
            double c = cost[6];
            c += __shfl_xor_sync(kFull, c, 2);
            c += __shfl_xor_sync(kFull, c, 1);
            if (q == 3) o[(size_t)6 * 32] = c;
        }
    }
}

// computeProjectError on the packed layout, double arithmetic (src/multicalib.cpp:969-983)
__global__ void __launch_bounds__(kF32Threads) reproj_error_packed_kernel(Problem P)
{
    const int which = P.st->cur;
    const double* __restrict__ x = P.x[which];
    const double* __restrict__ vR = P.vR[which];
    const int lane = threadIdx.x & 31, q = lane & 3, el = lane >> 2;
    const int n_units = P.n_edge_int >> 3;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_units; u += warps_total) {
        const int tile = u >> 2, wq = u & 3;
        const int edge = (u << 3) + el;
        const int frame = P.e_frame[edge];
        const EdgeMeta em = P.emeta[edge];
        const int n = em.end - em.begin;
        const int kp = P.tile_kp[tile];
        double sq = 0, nrm = 0;
        if (frame >= 0) {
            const int c = em.cam;
            const int64_t pv = P.n_cam + frame;
            double Rc[9], tc[3], Rp[9], tp[3], R3[9], T3[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) { Rc[i] = vR[9 * c + i]; Rp[i] = vR[9 * pv + i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) { tc[i] = c == 0 ? 0.0 : x[6 * (c - 1) + 3 + i]; tp[i] = x[6 * (pv - 1) + 3 + i]; }
            compose_pose(Rc, tc, Rp, tp, R3, T3);
            const CamParams cam = P.cams[c];
            const float2* base = P.obs2 + P.tile_off[tile] + (size_t)wq * kp * 5 * 32 + lane;
            for (int k = 0; k < kp; ++k) {
                const int c0 = 8 * k + 2 * q;
                if (c0 >= n) break;
                float2 o[5];
#pragma unroll
                for (int pl = 0; pl < 5; ++pl) o[pl] = __ldg(base + ((size_t)k * 5 + pl) * 32);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c0 + h >= n) break;
                    const float ox = h ? o[0].y : o[0].x, oy = h ? o[1].y : o[1].x, oz = h ? o[2].y : o[2].x;
                    const float iu = h ? o[3].y : o[3].x, iv = h ? o[4].y : o[4].x;
                    if (cam.model == kPinhole) {
                        if (cam.rational) corner_error<kPinhole, true>(cam, R3, T3, ox, oy, oz, iu, iv, &sq, &nrm);
                        else corner_error<kPinhole, false>(cam, R3, T3, ox, oy, oz, iu, iv, &sq, &nrm);
                    } else {
                        corner_error<kOmnidir, false>(cam, R3, T3, ox, oy, oz, iu, iv, &sq, &nrm);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 2; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(kFull, sq, o);
            nrm += __shfl_xor_sync(kFull, nrm, o);
        }
        if (q == 0) { P.err_sq[edge] = sq; P.err_nrm[edge] = nrm; }
    }
}

// observations: AoS host layout -> packed pair layout of the single-precision pass.  One thread per (step, lane) slot.
__global__ void gather_obs_packed_kernel(int n_tiles, const int64_t* __restrict__ tile_off, const int* __restrict__ tile_kp,
                                         const int* __restrict__ e_off, const int64_t* __restrict__ e_src,
                                         const float* __restrict__ obj, const float* __restrict__ img, float2* __restrict__ obs2,
                                         int* __restrict__ nonplanar)
{
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int kp = tile_kp[tile];
        float2* dst = obs2 + tile_off[tile];
        for (int idx = threadIdx.x; idx < 4 * kp * 32; idx += blockDim.x) {
            const int lane = idx & 31, wk = idx >> 5, wq = wk / kp, k = wk - wq * kp;
            const int edge = tile * 32 + wq * 8 + (lane >> 2);
            const int n = e_off[edge + 1] - e_off[edge];
            const int64_t src = e_src[edge];
            const int c0 = 8 * k + 2 * (lane & 3);
            float a[5] = {0, 0, 0, 0, 0}, b[5] = {0, 0, 0, 0, 0};
            if (c0 < n) {
                a[0] = obj[3 * (src + c0)]; a[1] = obj[3 * (src + c0) + 1]; a[2] = obj[3 * (src + c0) + 2];
                a[3] = img[2 * (src + c0)]; a[4] = img[2 * (src + c0) + 1];
            }
            if (c0 + 1 < n) {
                b[0] = obj[3 * (src + c0 + 1)]; b[1] = obj[3 * (src + c0 + 1) + 1]; b[2] = obj[3 * (src + c0 + 1) + 2];
                b[3] = img[2 * (src + c0 + 1)]; b[4] = img[2 * (src + c0 + 1) + 1];
            }
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) dst[((size_t)wk * 5 + pl) * 32 + lane] = make_float2(a[pl], b[pl]);
            if (a[2] != 0.0f || b[2] != 0.0f) *nonplanar = 1;   // benign race: every writer stores the same value
        }
    }
}

// Residual-only variant for computeProjectError: per edge sum |e|^2 and sum |e| at the current parameters.
template <int kModel, bool kRational>
__device__ __forceinline__ void edge_error_loop(const Problem& P, const CamParams& cam, const double* R3,
                                                const double* T3, int begin, int end, int sub, double* sq, double* nrm)
{
    for (int i = begin + sub; i < end; i += kLanesPerEdge)
        corner_error<kModel, kRational>(cam, R3, T3, __ldg(P.ox + i), __ldg(P.oy + i), __ldg(P.oz + i), __ldg(P.iu + i),
                                        __ldg(P.iv + i), sq, nrm);
}

__global__ void __launch_bounds__(kK1Threads, 2) reproj_error_kernel(Problem P)
{
    const int which = P.st->cur;
    const double* __restrict__ x = P.x[which];
    const double* __restrict__ vR = P.vR[which];
    const int sub = threadIdx.x & (kLanesPerEdge - 1);
    const int n_groups = (gridDim.x * blockDim.x) / kLanesPerEdge;
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) / kLanesPerEdge; e < P.n_edge_int; e += n_groups) {
        double sq = 0, nrm = 0;
        const int frame = P.e_frame[e];
        if (frame >= 0) {
            const int c = P.e_cam[e];
            const int64_t pv = P.n_cam + frame;
            double Rc[9], tc[3], Rp[9], tp[3], R3[9], T3[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) { Rc[i] = vR[9 * c + i]; Rp[i] = vR[9 * pv + i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) { tc[i] = c == 0 ? 0.0 : x[6 * (c - 1) + 3 + i]; tp[i] = x[6 * (pv - 1) + 3 + i]; }
            compose_pose(Rc, tc, Rp, tp, R3, T3);
            const CamParams cam = P.cams[c];
            const int b = P.e_off[e], n = P.e_off[e + 1];
            if (cam.model == kPinhole) {
                if (cam.rational) edge_error_loop<kPinhole, true>(P, cam, R3, T3, b, n, sub, &sq, &nrm);
                else edge_error_loop<kPinhole, false>(P, cam, R3, T3, b, n, sub, &sq, &nrm);
            } else {
                edge_error_loop<kOmnidir, false>(P, cam, R3, T3, b, n, sub, &sq, &nrm);
            }
        }
#pragma unroll
        for (int o = kLanesPerEdge / 2; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(kFull, sq, o);
            nrm += __shfl_xor_sync(kFull, nrm, o);
        }
        if (sub == 0) { P.err_sq[e] = sq; P.err_nrm[e] = nrm; }
    }
}

// --------------------------------------------------------------------------------------------------------
// K2: per-frame Schur elimination.  One thread per frame slot; the 32 slots of a warp belong to one group
// (frames seen by the same camera set), so their contributions to the reduced camera system land in the same
// 6x6 blocks and are summed with warp shuffles into ONE record per warp -- no atomics.
// Record layout: [cost, bad] | Va x 36 diagonal blocks | Va x 6 gradient | Va(Va-1)/2 x 36 off-diagonal blocks,
// Va = number of non-gauge views of the group.
// --------------------------------------------------------------------------------------------------------
// Transposed butterfly reduction of N (power of two <= 32) per-lane values over the warp: each halving step trades
// half of the values with the partner lane, so N values cost N - 1 + (5 - log2 N) shuffles instead of 5 N.
// On return v[0] holds the warp total of element  lane >> (5 - log2 N)  (the same in every lane of that group).
template <int N>
__device__ __forceinline__ void warp_tr_reduce(double (&v)[N], int lane)
{
    int o = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double send = up ? v[i] : v[i + n / 2], keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
    }
#pragma unroll
    for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(kFull, v[0], o);
}

// warp sum of a 6 x 6 block held per lane -> dst[0..36)
__device__ __forceinline__ void reduce_store36(const double* v, double* dst, int lane)
{
    double a[32], b[4];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = v[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = v[32 + k];
    warp_tr_reduce<32>(a, lane);
    warp_tr_reduce<4>(b, lane);
    dst[lane] = a[0];
    if ((lane & 7) == 0) dst[32 + (lane >> 3)] = b[0];
}
// warp sum of the LOWER triangle of a symmetric 6 x 6 block held per lane (v[i * 6 + j], i >= j) -> dst[i * 6 + j], i >= j;
// the upper entries of dst are not written (reduce_records mirrors the diagonal blocks): 21 values travel instead of 36
__device__ __forceinline__ void reduce_store21(const double* v, double* dst, int lane)
{
    double a[16], b[4], c;
#pragma unroll
    for (int i = 0, t = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j, ++t) {
            if (t < 16) a[t] = v[i * 6 + j];
            else if (t < 20) b[t - 16] = v[i * 6 + j];
            else c = v[i * 6 + j];
        }
    warp_tr_reduce<16>(a, lane);
    warp_tr_reduce<4>(b, lane);
    c = warp_sum(c);
    auto at = [](int t) { const int i = (t >= 1) + (t >= 3) + (t >= 6) + (t >= 10) + (t >= 15); return i * 6 + (t - i * (i + 1) / 2); };
    if ((lane & 1) == 0) dst[at(lane >> 1)] = a[0];
    if ((lane & 7) == 0) dst[at(16 + (lane >> 3))] = b[0];
    if (lane == 0) dst[35] = c;
}
// warp sum of 6 values -> dst[0..6)
__device__ __forceinline__ void reduce_store6(const double* v, double* dst, int lane)
{
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = k < 6 ? v[k] : 0.0;
    warp_tr_reduce<8>(a, lane);
    if ((lane & 3) == 0 && (lane >> 2) < 6) dst[lane >> 2] = a[0];
}

// sel: -1 = decide from the state (loop), 0/1 = explicit buffer with explicit lambda (diagnostics)
constexpr int kK2WarpDoubles = 2 * 36 * 32;
constexpr int kK2SmemBytes = (kK2Threads / 32) * kK2WarpDoubles * 8 + (kK2Threads / 32) * 8;
template <int kMinBlocks>
__global__ void __launch_bounds__(kK2Threads, kMinBlocks) frame_schur_kernel(Problem P, int sel_arg, double lambda_arg)
{
    const DevState* st = P.st;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    if (warp >= P.n_warps) return;
    // first-level loads that do not depend on the loop state are requested together with it (the head of this kernel was
    // a chain of five dependent round trips: 18 % of the stall samples, profiles/r2_ncu_iteration_kernels.txt)
    const int4 m0 = P.wmeta[2 * warp], m1 = P.wmeta[2 * warp + 1];
    const int slot = warp * 32 + lane;
    const int frame = P.slot_frame[slot];
    int sel;
    double lambda;
    if (sel_arg >= 0) { sel = sel_arg; lambda = lambda_arg; }
    else {
        const int done = st->done, phase = st->phase, cur = st->cur;
        const double lam_spec = st->lambda_spec, lam = st->lambda;
        if (done) return;
        if (phase == kPhaseDecide) { sel = 1 - cur; lambda = lam_spec; }
        else { sel = cur; lambda = lam; }
    }
    extern __shared__ __align__(128) unsigned char k2_smem[];
    const int V = m0.x;
    const int* gc = P.group_cams + m0.y;
    const int ls = m1.x + lane;
    const int ebase = m0.z, stride = m0.w;
    const double* __restrict__ blk = P.blocks[sel];
    // per warp: two 36 x 32 tiles.  A tile first receives the 28 x 32 edge-block tile of view 0 / 1 from the TMA
    // engine (tile-major records: one contiguous 7 KB block), later the lanes overwrite their own column with Y.
    double* stage = reinterpret_cast<double*>(k2_smem) + wic * kK2WarpDoubles;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(k2_smem + (kK2Threads / 32) * kK2WarpDoubles * 8) + wic;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (V > 1 ? 2u : 1u) * kBlk * 256u);
        tma_load_1d(stage, blk + (int64_t)((ebase + ls) >> 5) * kBlk * 32, kBlk * 256u, bar);
        if (V > 1) tma_load_1d(stage + 36 * 32, blk + (int64_t)((ebase + stride + ls) >> 5) * kBlk * 32, kBlk * 256u, bar);
    }
    __syncwarp();
    const int c0 = gc[0], c1 = V > 1 ? gc[1] : 0;      // camera ids of the two staged views, in one round
    const bool active = frame >= 0;
    const double* __restrict__ x = P.x[sel];
    const double* __restrict__ vR = P.vR[sel];
    double* rec = P.records + m1.y;

    double tp[3] = {0, 0, 0};
    if (active) {
        const int64_t pv = P.n_cam + frame;
#pragma unroll
        for (int i = 0; i < 3; ++i) tp[i] = x[6 * (pv - 1) + 3 + i];
    }
    double U[21], z[6], cost = 0;
#pragma unroll
    for (int i = 0; i < 21; ++i) U[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) z[i] = 0;
    int Va = 0;
    mbar_wait(bar, 0);
    // pass 1: pattern-pose block
    for (int v = 0; v < V; ++v) {
        const int c = v == 0 ? c0 : (v == 1 ? c1 : gc[v]);
        if (c != 0) ++Va;
        if (!active) continue;
        const int64_t e = ebase + (int64_t)v * stride + ls;
        double t[kBlk], H[36], Rc[9];
        if (v < 2) {
#pragma unroll
            for (int k = 0; k < kBlk; ++k) t[k] = stage[(v * 36 + k) * 32 + lane];
        } else {
#pragma unroll
            for (int k = 0; k < kBlk; ++k) t[k] = blk[tile_idx(kBlk, e, k)];
        }
        cost += t[27];
        unpack_sym6(t, H);
        if (c != 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) Rc[i] = vR[9 * c + i];
        }
        lift_frame(H, t + 21, Rc, c == 0, U, z);
    }
    int bad = 0;
    if (active) {
#pragma unroll
        for (int i = 0; i < 6; ++i) U[tri6(i, i)] *= (1.0 + lambda);
        if (!chol6_packed(U)) bad = 1;
        chol6_forward(U, z, 1);
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) U[tri6(i, i)] = 1.0;
    }
#pragma unroll
    for (int k = 0; k < 21; ++k) P.frameL[tile_idx(27, slot, k)] = U[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) P.frameL[tile_idx(27, slot, 21 + k)] = z[k];
    {
        const double c2 = warp_sum(cost), b2 = warp_sum((double)bad);
        if (lane == 0) { rec[0] = c2; rec[1] = b2; P.warp_scal[warp] = c2; P.warp_scal[P.n_warps + warp] = b2; }
    }
    // pass 2: camera blocks of each non-gauge view
    int ai = 0;
    for (int v = 0; v < V; ++v) {
        const int c = v == 0 ? c0 : (v == 1 ? c1 : gc[v]);
        if (c == 0) continue;
        const int64_t e = ebase + (int64_t)v * stride + ls;
        double D[36], gd[6];
#pragma unroll
        for (int i = 0; i < 36; ++i) D[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) gd[i] = 0;
        if (active) {
            double t[kBlk], H[36], Rc[9], s[3], Y[36], gcv[6];
            if (v < 2) {
#pragma unroll
                for (int k = 0; k < kBlk; ++k) t[k] = stage[(v * 36 + k) * 32 + lane];
            } else {
#pragma unroll
                for (int k = 0; k < kBlk; ++k) t[k] = blk[tile_idx(kBlk, e, k)];
            }
            unpack_sym6(t, H);
#pragma unroll
            for (int i = 0; i < 9; ++i) Rc[i] = vR[9 * c + i];
            mat3_vec(Rc, tp, s);
            lift_camera(H, t + 21, Rc, s, D, gcv, Y);
#pragma unroll
            for (int j = 0; j < 6; ++j) chol6_forward(U, Y + j, 6);  // Y = L^-1 W, column by column
#pragma unroll
            for (int k = 0; k < 36; ++k) {
                if (kYFloat && P.prec) reinterpret_cast<float*>(P.edgeY)[tile_idx(36, e, k)] = (float)Y[k];
                else P.edgeY[tile_idx(36, e, k)] = Y[k];
            }
            if (v < 2) {   // keep Y for the off-diagonal pass (own column only: no other lane reads it)
#pragma unroll
                for (int k = 0; k < 36; ++k) stage[(v * 36 + k) * 32 + lane] = Y[k];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) D[i * 6 + i] *= (1.0 + lambda);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {      // the camera block is symmetric: lower triangle only
                    double acc = 0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc += Y[k * 6 + i] * Y[k * 6 + j];
                    D[i * 6 + j] -= acc;
                }
                double acc = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) acc += Y[k * 6 + i] * z[k];
                gd[i] = gcv[i] - acc;
            }
        }
        reduce_store21(D, rec + 2 + 36 * ai, lane);
        reduce_store6(gd, rec + 2 + 36 * Va + 6 * ai, lane);
        ++ai;
    }
    // pass 3: off-diagonal blocks  -Y_a^T Y_b  for non-gauge views a < b
    int pi = 0;
    for (int va = 0; va < V; ++va) {
        if (gc[va] == 0) continue;
        for (int vb = va + 1; vb < V; ++vb) {
            if (gc[vb] == 0) continue;
            double D[36];
#pragma unroll
            for (int i = 0; i < 36; ++i) D[i] = 0;
            if (active) {
                const int64_t ea = ebase + (int64_t)va * stride + ls, ebx = ebase + (int64_t)vb * stride + ls;
                double Ya[36], Yb[36];
                if (vb < 2) {   // va = 0, vb = 1: both still in shared memory
#pragma unroll
                    for (int k = 0; k < 36; ++k) { Ya[k] = stage[k * 32 + lane]; Yb[k] = stage[(36 + k) * 32 + lane]; }
                } else {
#pragma unroll
                    for (int k = 0; k < 36; ++k) {
                        if (kYFloat && P.prec) { Ya[k] = reinterpret_cast<const float*>(P.edgeY)[tile_idx(36, ea, k)]; Yb[k] = reinterpret_cast<const float*>(P.edgeY)[tile_idx(36, ebx, k)]; }
                        else { Ya[k] = P.edgeY[tile_idx(36, ea, k)]; Yb[k] = P.edgeY[tile_idx(36, ebx, k)]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        double acc = 0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc += Ya[k * 6 + i] * Yb[k * 6 + j];
                        D[i * 6 + j] = -acc;
                    }
            }
            reduce_store36(D, rec + 2 + 42 * Va + 36 * pi, lane);
            ++pi;
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// K3a: sum the warp records into the reduced system.  One CTA (8 warps) per destination: the source list is cut
// into 8 contiguous slices (a fixed partition: the result is bit-stable run to run), every warp sums its slice in
// order with the record offsets fetched 32 at a time and 8 gathers in flight, and warp 0 adds the 8 partial sums in
// order.  The kernel is pure L2 latency (a few hundred dependent gathers per destination), hence the width.
// ar must be zeroed beforehand (blocks without sources stay zero).
// --------------------------------------------------------------------------------------------------------
constexpr int kK3Threads = 256;
__global__ void __launch_bounds__(kK3Threads) reduce_records_kernel(Problem P, int forced)
{
    __shared__ double s_part[kK3Threads / 32][40];
    const int dest = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the static tables are requested together with the loop state (this kernel is a chain of dependent round trips)
    const int kind = P.dest_info[4 * dest], A = P.dest_info[4 * dest + 1], B = P.dest_info[4 * dest + 2];
    const int s0 = P.dest_src0[dest], s1 = P.dest_src0[dest + 1];
    const int done = P.st->done;
    if (!forced && done) return;
    const int ns = P.ns;
    double* S = P.ar_part;
    double* gs = P.ar_part + P.ar_goff;
    double* sc = gs + ns;
    if (P.dag_words > 0) {
        // the solve that follows wants its output buffer filled with the all-ones sentinel (mccba_dense.cuh); this
        // kernel is pure gather latency, so its CTAs write the 1.2 MB on the side instead of a separate memset node
        const int64_t per = (P.dag_words / 2 + gridDim.x - 1) / gridDim.x;
        const int64_t b2 = (int64_t)blockIdx.x * per, e2 = min(b2 + per, P.dag_words / 2);
        double2* d2 = reinterpret_cast<double2*>(P.dag_buf);
        const double sent = __longlong_as_double(-1LL);
        for (int64_t i = b2 + threadIdx.x; i < e2; i += kK3Threads) d2[i] = make_double2(sent, sent);
        if ((P.dag_words & 1) && blockIdx.x == 0 && threadIdx.x == 0) P.dag_buf[P.dag_words - 1] = sent;
    }
    if (kind == 2) {
        // scalars: cost and bad from every warp record; frame step / param norms from frame_update's partials
        double c = 0, b = 0, n0 = 0, n1 = 0;
        for (int w = threadIdx.x; w < P.n_warps; w += kK3Threads) { c += P.warp_scal[w]; b += P.warp_scal[P.n_warps + w]; }
        for (int k = threadIdx.x; k < P.n_k4_blocks; k += kK3Threads) { n0 += P.norm_part[k]; n1 += P.norm_part[P.n_k4_blocks + k]; }
        c = warp_sum(c); b = warp_sum(b); n0 = warp_sum(n0); n1 = warp_sum(n1);
        if (lane == 0) { s_part[warp][0] = c; s_part[warp][1] = n0; s_part[warp][2] = n1; s_part[warp][3] = b; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < kK3Threads / 32; ++w) t += s_part[w][threadIdx.x];
            sc[threadIdx.x] = t;
        }
        return;
    }
    const int per = (s1 - s0 + kK3Threads / 32 - 1) / (kK3Threads / 32);
    const int b0 = min(s0 + warp * per, s1), e0 = min(b0 + per, s1);
    const bool hi = kind == 0 && lane < 4;        // elements 32..35 of a 6 x 6 block
    const bool lo = kind == 0 || lane < 6;        // kind 1: 6 gradient entries
    double a0 = 0, a1 = 0;
    for (int s = b0; s < e0; s += 32) {
        const int cnt = min(32, e0 - s);
        const int myoff = lane < cnt ? P.dest_src[s + lane] : 0;
        for (int q0 = 0; q0 < cnt; q0 += 8) {
            double v[8], u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int off = __shfl_sync(kFull, myoff, (q0 + q) & 31);
                const bool on = q0 + q < cnt;
                v[q] = (on && lo) ? P.records[off + lane] : 0.0;
                u[q] = (on && hi) ? P.records[off + 32 + lane] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) { a0 += v[q]; a1 += u[q]; }
        }
    }
    s_part[warp][lane] = a0;
    if (lane < 4) s_part[warp][32 + lane] = a1;
    __syncthreads();
    if (warp != 0) return;
    a0 = 0; a1 = 0;
#pragma unroll
    for (int w = 0; w < kK3Threads / 32; ++w) { a0 += s_part[w][lane]; if (lane < 4) a1 += s_part[w][32 + lane]; }
    if (kind == 0 && P.band_nw > 0) {
        // packed band, lower triangle: element (row, col), row >= col, at band[row][col - row + w]
        const int NW = P.band_nw, w = NW - 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int k = half ? 32 + lane : lane;
            if (half && lane >= 4) break;
            const int i = k / 6, j = k % 6;
            int row = 6 * A + i, col = 6 * B + j;
            if (row < col) { const int t = row; row = col; col = t; }   // (A, B) with A < B holds the transpose
            if (A != B || i >= j) S[(int64_t)row * NW + (col - row + w)] = half ? a1 : a0;
        }
    } else if (kind == 0) {
        // diagonal blocks (A == B) arrive as their lower triangle only (reduce_store21) and are mirrored here
        {
            const int i = lane / 6, j = lane % 6;
            if (A != B || i >= j) {
                S[(int64_t)(6 * A + i) * ns + 6 * B + j] = a0;
                S[(int64_t)(6 * B + j) * ns + 6 * A + i] = a0;
            }
        }
        if (lane < 4) {
            const int k = 32 + lane, i = k / 6, j = k % 6;
            if (A != B || i >= j) {
                S[(int64_t)(6 * A + i) * ns + 6 * B + j] = a1;
                S[(int64_t)(6 * B + j) * ns + 6 * A + i] = a1;
            }
        }
    } else if (lane < 6) {
        gs[6 * A + lane] = a0;
    }
}

// loop control (single warp): accept/reject of the trial point, damping, termination; sets st->go
// --------------------------------------------------------------------------------------------------------
// C1 (N > 1): the per-iteration exchange of the packed buffer [S | g | scalars] over NVLink peer memory.
// Every rank stores its partial buffer into slot[rank] of EVERY rank's window (IPC-mapped peer memory) and adds the n
// slots of its own window in rank order -- every rank performs the identical additions, so the reduced system and
// therefore the camera parameters are bit-identical on all ranks.
// Windows are double-buffered by epoch parity: a rank can only be one exchange ahead of its slowest peer.
// --------------------------------------------------------------------------------------------------------
constexpr int kP2pFlagWords = 32;   // header of a window (unused by the LL protocol, keeps the slots 256-byte aligned)
constexpr int kP2pThreads = 256;
constexpr int kP2pStatsWord = 8;    // p2p_epoch[8..12): sum of store ns, sum of wait ns, launches, max wait ns
// Low-latency protocol (the one NCCL calls LL): every double travels as two 8-byte words {32 data bits | 32-bit epoch},
// so the data carry their own arrival flag -- no fence, no separate flag, no grid-wide step.  A thread stores its
// elements into slot[rank] of every window, then spins on the n slots of its own window until both words of an
// element carry this launch's epoch, and adds them in rank order.  8-byte stores are single-copy atomic, which is all
// the protocol needs; slots are double-buffered by epoch parity (a rank can be at most one exchange ahead of a peer).
__device__ __forceinline__ void p2p_ll_store(unsigned long long* dst, double x, unsigned flag)
{
    const unsigned long long v = (unsigned long long)__double_as_longlong(x);
    const unsigned long long w0 = (v & 0xffffffffull) | ((unsigned long long)flag << 32);
    const unsigned long long w1 = (v >> 32) | ((unsigned long long)flag << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
}
// Bounded spin (a peer that died, returned early from mccba_solve or launched a different number of graphs must not
// hang the GPU): the wall clock is sampled every 1024 polls; on expiry the caller gets *timed_out = 1 and zeros.
__device__ __forceinline__ unsigned long long p2p_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Sum of element i over the n rank slots of the own window, in rank order.  All slots that have not arrived yet are
// polled in ONE batch of independent loads per round (a round costs one memory round trip, not n of them: at 8 ranks the
// slot-after-slot version spent 11-15 us per exchange just on serialised polling, profiles/r2_bench_n8.json).
__device__ __forceinline__ double p2p_ll_sum(const unsigned long long* win, int64_t stride, int n, int64_t i, unsigned flag,
                                             unsigned long long budget_ns, int* timed_out)
{
    unsigned long long w0[8], w1[8], t0 = 0;
    unsigned pending = (1u << n) - 1u, polls = 0;
    for (;;) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((pending >> r) & 1u)
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[r]), "=l"(w1[r]) : "l"(win + (int64_t)r * stride + 2 * i) : "memory");
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (((pending >> r) & 1u) && (unsigned)(w0[r] >> 32) == flag && (unsigned)(w1[r] >> 32) == flag) pending &= ~(1u << r);
        if (!pending) break;
        if ((++polls & 255u) == 0) {
            const unsigned long long now = p2p_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > budget_ns) { *timed_out = 1; return 0.0; }
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
        if (r < n) {
            const double v = __longlong_as_double((long long)((w0[r] & 0xffffffffull) | (w1[r] << 32)));
            acc = r == 0 ? v : acc + v;
        }
    return acc;
}
__global__ void __launch_bounds__(kP2pThreads) p2p_exchange_kernel(Problem P, int64_t len, unsigned long long budget_ns)
{
    // A finished loop exchanges nothing (every rank takes the same decisions from the same sums, so all of them stop at the
    // same launch; decide_body still advances the epoch with every launch).  After a time-out (status 5) this also keeps the
    // remaining launches of the chunk from spinning through their budgets again.
    if (P.st->done) return;
    const unsigned long long e = *P.p2p_epoch + 1;
    const unsigned flag = (unsigned)e;
    const int n = P.p2p_n, me = P.p2p_rank;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    const int64_t par = (int64_t)((e & 1) * n) * P.p2p_stride;          // slots of this parity (in doubles)
    // %globaltimer stamps of one thread (mccba_exchange_stats): time to post the stores, then time until the last peer's
    // words have arrived = arrival skew of the ranks + one NVLink trip
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_in = 0, t_sent = 0;
    if (stamp) t_in = p2p_now_ns();
    for (int r = 0; r < n; ++r) {
        const int peer = (me + r) % n;      // spread the traffic: own window first, then the next rank ...
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.p2p_peer[peer] + kP2pFlagWords + par + (int64_t)me * P.p2p_stride);
        for (int64_t i = tid; i < len; i += step) p2p_ll_store(dst + 2 * i, P.ar_part[i], flag);
    }
    if (stamp) t_sent = p2p_now_ns();
    const unsigned long long* win = reinterpret_cast<const unsigned long long*>(P.p2p_peer[me] + kP2pFlagWords + par);
    int timed_out = 0;
    for (int64_t i = tid; i < len; i += step) {
        const double acc = p2p_ll_sum(win, P.p2p_stride, n, i, flag, budget_ns, &timed_out);
        P.ar[i] = timed_out ? 0.0 : acc;
        if (timed_out) break;
    }
    if (stamp) {
        const unsigned long long t_done = p2p_now_ns();
        unsigned long long* stats = P.p2p_epoch + kP2pStatsWord;
        stats[0] += t_sent - t_in;
        stats[1] += t_done - t_sent;
        stats[2] += 1;
        if (t_done - t_sent > stats[3]) stats[3] = t_done - t_sent;
    }
    if (timed_out) {   // sticky: the loop control of this launch (decide_body) sees done and does nothing else
        P.st->status = 5;
        P.st->done = 1;
    }
}

// loop control of one launch (single thread): accept/reject of the trial point, damping, termination
__device__ __forceinline__ void decide_body(const Problem& P)
{
    DevState* st = P.st;
    if (P.p2p_n > 1) *P.p2p_epoch += 1;   // the exchange of this launch is complete (both kernels ran before this one)
    const int ns = P.ns;
    const double* sc = P.ar + P.ar_goff + ns;
    int go = 0;
    if (!st->done) {
        st->launches += 1;
        st->solved = 0;
        const double cost_sel = sc[0];
        const bool numeric_ok = (sc[3] == 0.0) && isfinite(cost_sel);
        if (!numeric_ok && !(st->phase == kPhaseDecide && st->mode == 1)) {
            // a non-SPD pattern-pose block or a non-finite cost at an ACCEPTED point is fatal; at an LM
            // trial point it is just a rejected step
            st->status = 4;
            st->done = 1;
        }
        if (!st->done) {
            if (st->phase == kPhaseDecide) {
                st->iter += 1;
                st->cost_trial = cost_sel;
                const bool accept = (st->mode == 0) || (numeric_ok && cost_sel < st->cost_cur);
                if (accept) {
                    st->cur = 1 - st->cur;
                    st->cost_cur = cost_sel;
                    st->change = sqrt(sc[1] + st->cam_step2) / sqrt(sc[2] + st->cam_param2);
                    st->lambda = st->lambda_spec;
                    st->n_accept += 1;
                    go = 1;
                } else {
                    st->lambda = fmin(st->lambda * st->lambda_up, 1e15);
                    st->n_reject += 1;
                    st->phase = kPhaseRebuild;
                }
            } else {
                if (st->phase == kPhaseFirst) st->cost_cur = cost_sel;
                go = 1;
            }
            const int t = st->crit_type;
            const bool stop = (t == 1 && st->iter >= st->max_count) || (t == 2 && st->change <= st->eps) ||
                              (t == 3 && (st->change <= st->eps || st->iter >= st->max_count));
            if (stop) { st->done = 1; go = 0; }
        }
    }
    st->go = go;
    st->chol_fail = 0;
}
__global__ void decide_kernel(Problem P)
{
    if (threadIdx.x == 0) decide_body(P);
}

// the whole factorisation + both substitutions in one launch (mccba_dense.cuh, tile DAG)
__device__ __forceinline__ void camera_update_body(const Problem& P, const double* dc, int fail);

// fused != 0: the CTA of tile (0, 0) -- the last one to finish, the backward sweep ends there -- also runs the camera
// update (one kernel boundary less per iteration).  It reads the solution from the sentinel buffer, where every word
// validates itself, and treats a non-finite entry as a failed factorisation (a bad pivot turns its column into NaN).
__global__ void __launch_bounds__(256) chol_dag_kernel(CholDag D, Problem P, int fused)
{
    if (D.go && !*D.go) return;
    chol_dag_tile(D);
    if (!fused || blockIdx.x != 0) return;
    __shared__ double s_dx[512];
    __shared__ int s_nonfinite;
    const int n = D.n;
    if (threadIdx.x == 0) s_nonfinite = 0;
    __syncthreads();   // also: warp 0 is back from the backward sweep
    const double* xs = D.L + (size_t)(n + 1) * n + (size_t)chol_row_tiles(n) * chol_col_tiles(n) * kCT + n;
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = dag_poll(xs + i);
        s_dx[i] = v;
        if (!isfinite(v)) bad = 1;
    }
    if (bad) s_nonfinite = 1;
    __syncthreads();
    const int fail = s_nonfinite || *reinterpret_cast<volatile int*>(D.fail);
    camera_update_body(P, s_dx, fail);
}

// camera step (tangent -> additive Rodrigues step, scaled), trial camera parameters and rotations, end-of-iteration
// state.  Called by one whole CTA (any multiple of 32 threads); dc = solution of the reduced system.
__device__ __forceinline__ void camera_update_body(const Problem& P, const double* dc, int fail)
{
    DevState* st = P.st;
    // tangent step -> additive Rodrigues step, scaled, trial parameters
    const int cur = st->cur, tr = 1 - cur;
    const double alpha = st->mode == 0 ? pow(0.95, (double)st->iter + 1.0) : 1.0;
    double step2 = 0, par2 = 0;
    for (int c = 1 + threadIdx.x; c < P.n_cam; c += blockDim.x) {
        const double* p = P.x[cur] + 6 * (c - 1);
        const double om[3] = {p[0], p[1], p[2]};
        const double* d = dc + 6 * (c - 1);
        double dom[3], R[9];
        left_jacobian_inv_apply(om, d, dom);
        double q[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double s0 = alpha * dom[i], s1 = alpha * d[3 + i];
            q[i] = p[i] + s0; q[3 + i] = p[3 + i] + s1;
            step2 += s0 * s0 + s1 * s1;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { P.x[tr][6 * (c - 1) + i] = q[i]; par2 += q[i] * q[i]; }
        rodrigues(q, R);
#pragma unroll
        for (int i = 0; i < 9; ++i) P.vR[tr][9 * c + i] = R[i];
    }
    // block reduction of the camera norms (fixed order: warp shuffle, then warp 0 over the warp sums)
    __shared__ double s_red[2][32];
    step2 = warp_sum(step2); par2 = warp_sum(par2);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = step2; s_red[1][threadIdx.x >> 5] = par2; }
    __syncthreads();
    if (threadIdx.x < 32) {
        double a = threadIdx.x < (blockDim.x >> 5) ? s_red[0][threadIdx.x] : 0.0;
        double b = threadIdx.x < (blockDim.x >> 5) ? s_red[1][threadIdx.x] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
        if (threadIdx.x == 0) {
            st->cam_step2 = a;
            st->cam_param2 = b;
            st->alpha = alpha;
            if (fail) { st->status = 4; st->done = 1; }
            else {
                st->solved = 1;
                st->phase = kPhaseDecide;
                st->lambda_spec = st->mode == 1 ? fmax(st->lambda * st->lambda_down, 1e-15) : 0.0;
            }
        }
    }
}

// Banded reduced system (block cyclic reduction, mccba_bcr.cuh): one CTA stages the packed band as super-blocks, solves,
// and -- in the iteration (fused != 0) -- runs the camera update; fused == 2: the loop control (decide_body) runs here
// too, so the whole serial part of an iteration is ONE launch.  Dynamic shared memory: bcr_smem_bytes(n, B).
template <int B>
__global__ void __launch_bounds__(BcrCfg<B>::kThreads) chol_bcr_kernel(const double* A, int n, double* xout, int* fail_out,
                                                                       const int* go, Problem P, int fused, int packed)
{
    extern __shared__ __align__(16) unsigned char bcr_smem[];
    const int Nb = bcr_blocks(n, B);
    double* Dg = reinterpret_cast<double*>(bcr_smem);
    double* Lo = Dg + (size_t)Nb * B * B;
    double* Tmp = Lo + (size_t)Nb * B * B;
    double* rhs = Tmp + (size_t)((Nb + 1) / 2) * B * B;
    // the loop control (one thread, a chain of dependent global loads) runs while the other warps stage the system: the
    // packed buffer is final before this kernel starts whatever the decision will be
    if (fused == 2 && threadIdx.x == 0) decide_body(P);
    if (!(fused == 2 && threadIdx.x < 32)) bcr_stage<B>(A, n, packed, Dg, Lo, rhs, Nb, fused == 2 ? 32 : 0);
    __syncthreads();
    if (go && !*reinterpret_cast<const volatile int*>(go)) return;
    int fail = bcr_solve_cta<B>(Dg, Lo, Tmp, rhs, Nb);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const double v = rhs[idx];
        xout[idx] = v;
        if (!isfinite(v)) fail = 1;
    }
    fail = __syncthreads_or(fail);
    if (fused) camera_update_body(P, rhs, fail);
    else if (threadIdx.x == 0 && fail) *fail_out = 1;
}

// camera update for the cases where it is not fused into the solve (tile DAG with n_s > 512, or no camera unknowns at
// all): the solution is already in P.dc.  Single CTA.
__global__ void __launch_bounds__(kK5Threads) camera_update_kernel(Problem P)
{
    DevState* st = P.st;
    if (!st->go) return;
    camera_update_body(P, P.dc, P.ns > 0 ? st->chol_fail : 0);
}

// --------------------------------------------------------------------------------------------------------
// K4: back-substitution of the pattern-pose steps + trial parameters + rotations of the trial point.
// One thread per frame slot.  d_p = L^-T (z - sum_v Y_v dc_v).
// --------------------------------------------------------------------------------------------------------
// Every record the thread reads is tile-major, so lane-indexed loads of one value are one contiguous 256-byte row per
// warp: the factor (27 values) and the Y blocks (36 per view) are read straight from global memory, ~100 independent
// coalesced loads per thread in flight, and Y is consumed as it arrives (6 accumulators instead of 72 registers).  Shared
// memory only stages the outgoing composed edge poses (two 3 KB bulk stores per warp), which leaves the register file
// as the only limit on residency: kK4MinBlocks CTAs per SM instead of the 2 the 25 KB-per-warp TMA staging allowed
// (measured on config #5: 33.4 us with the staged tiles, 30.9 / 29.2 / 33.0 us at 4 / 3 / 5 CTAs per SM = 128 / 168 / 96 registers).
#ifndef MCCBA_K4_MINBLOCKS
#define MCCBA_K4_MINBLOCKS 3
#endif

constexpr int kK4MinBlocks = MCCBA_K4_MINBLOCKS;
constexpr int kK4WarpBytes = 2 * 32 * (int)sizeof(EdgeRec);
constexpr int kK4SmemBytes = (kK4Threads / 32) * kK4WarpBytes;
__global__ void __launch_bounds__(kK4Threads, kK4MinBlocks) frame_update_kernel(Problem P)
{
    extern __shared__ __align__(128) unsigned char k4_smem[];
    const DevState* st = P.st;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool in = slot < P.n_slots;         // n_slots is a multiple of 32: whole warps
    // The kernel is a chain of dependent loads (profiles/r2_ncu_iteration_kernels.txt: 69 % of the stall samples are
    // long-scoreboard): everything that does not depend on the loop state is requested together with it ...
    int frame = -1;
    int4 m0 = make_int4(0, 0, 0, 0), m1 = m0;
    if (in) {
        frame = P.slot_frame[slot];
        m0 = P.wmeta[2 * (slot >> 5)];
        m1 = P.wmeta[2 * (slot >> 5) + 1];
    }
    const int done = st->done, solved = st->solved, cur = st->cur, tr = 1 - cur;
    const double alpha = st->alpha;
    if (done || !solved) return;
    EdgeRec* stage = reinterpret_cast<EdgeRec*>(k4_smem + wic * kK4WarpBytes);
    double step2 = 0, par2 = 0;
    if (in) {
        const int warp = slot >> 5;
        const int V = m0.x;
        const int* gc = P.group_cams + m0.y;
        const int ls = m1.x + lane;
        const int64_t ebase = m0.z, stride = m0.w;
        const double* __restrict__ fl = P.frameL + (int64_t)warp * 27 * 32 + lane;
        const double* __restrict__ dc = P.dc;
        // ... and the camera ids of the first two views come in one round.  (Pulling the second view's Y tile into L2 with
        // prefetch.global.L2 while the first is consumed was measured slower: 30.5 vs 29.0 us.)
        const int c0 = gc[0], c1 = V > 1 ? gc[1] : 0;
        if (frame >= 0) {
            const int64_t pv = P.n_cam + frame;
            double U[21], r[6], pold[6];
#pragma unroll
            for (int k = 0; k < 21; ++k) U[k] = fl[k * 32];
#pragma unroll
            for (int k = 0; k < 6; ++k) r[k] = fl[(21 + k) * 32];
#pragma unroll
            for (int k = 0; k < 6; ++k) pold[k] = P.x[cur][6 * (pv - 1) + k];
            for (int v = 0; v < V; ++v) {
                const int c = v == 0 ? c0 : (v == 1 ? c1 : gc[v]);
                if (c == 0) continue;
                const double* __restrict__ y = P.edgeY + ((ebase + v * stride + ls) >> 5) * 36 * 32 + lane;   // ls & 31 == lane
                double d[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) d[k] = dc[6 * (c - 1) + k];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double acc = 0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc += y[(i * 6 + k) * 32] * d[k];
                    r[i] -= acc;
                }
            }
            chol6_backward(U, r);
            const double om[3] = {pold[0], pold[1], pold[2]};
            double dom[3], q[6], R[9];
            left_jacobian_inv_apply(om, r, dom);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double s0 = alpha * dom[i], s1 = alpha * r[3 + i];
                q[i] = pold[i] + s0; q[3 + i] = pold[3 + i] + s1;
                step2 += s0 * s0 + s1 * s1;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) { P.x[tr][6 * (pv - 1) + i] = q[i]; par2 += q[i] * q[i]; }
            rodrigues(q, R);
#pragma unroll
            for (int i = 0; i < 9; ++i) P.vR[tr][9 * pv + i] = R[i];
            // composed poses the residual kernel evaluates next: the first two views go out through shared memory as
            // one 3 KB bulk store per view (32 consecutive edge records), the rest straight from registers
            for (int v = 0; v < V; ++v) {
                const int c = v == 0 ? c0 : (v == 1 ? c1 : gc[v]);
                double Rc[9], tc[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) Rc[k] = c != 0 ? P.vR[tr][9 * c + k] : ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0);
#pragma unroll
                for (int k = 0; k < 3; ++k) tc[k] = c != 0 ? P.x[tr][6 * (c - 1) + 3 + k] : 0.0;
                if (v < 2) compose_pose(Rc, tc, R, q + 3, stage[v * 32 + lane].R3, stage[v * 32 + lane].T3);
                else {
                    EdgeRec er;
                    compose_pose(Rc, tc, R, q + 3, er.R3, er.T3);
                    P.erec[ebase + v * stride + ls] = er;
                }
            }
        } else {
            EdgeRec z;
#pragma unroll
            for (int i = 0; i < 9; ++i) z.R3[i] = 0;
            z.T3[0] = z.T3[1] = 0;
            z.T3[2] = 1;      // padding slot: a benign pose (the packed pass evaluates its zero-weight corners)
            for (int v = 0; v < V; ++v) {
                if (v < 2) stage[v * 32 + lane] = z;
                else P.erec[ebase + v * stride + ls] = z;
            }
        }
        tma_store_fence();
        __syncwarp();
        if (lane == 0) {
            tma_store_1d(P.erec + ebase + ls, stage, 32u * (unsigned)sizeof(EdgeRec));
            if (V > 1) tma_store_1d(P.erec + ebase + stride + ls, stage + 32, 32u * (unsigned)sizeof(EdgeRec));
            tma_store_commit_wait();
        }
    }
    __shared__ double s_red[2][kK4Threads / 32];
    step2 = warp_sum(step2); par2 = warp_sum(par2);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = step2; s_red[1][threadIdx.x >> 5] = par2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
#pragma unroll
        for (int w = 0; w < kK4Threads / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        P.norm_part[blockIdx.x] = a;
        P.norm_part[P.n_k4_blocks + blockIdx.x] = b;
    }
}

// initialise the loop state for a solve (single thread)
__global__ void init_state_kernel(DevState* st, int mode, int crit_type, int max_count, double eps, double lambda0,
                                  double up, double down, int cur)
{
    st->mode = mode; st->crit_type = crit_type; st->max_count = max_count;
    st->iter = 0; st->done = 0; st->phase = kPhaseFirst; st->cur = cur; st->status = 0;
    st->n_accept = 0; st->n_reject = 0; st->launches = 0; st->solved = 0; st->go = 0; st->chol_fail = 0;
    st->eps = eps; st->lambda = mode == 1 ? lambda0 : 0.0; st->lambda_up = up; st->lambda_down = down;
    st->lambda_spec = st->lambda;
    st->cost_cur = 0; st->cost_trial = 0; st->change = 1.0; st->alpha = 1.0;
    st->cam_step2 = 0; st->cam_param2 = 0;
}

}  // namespace mccba
