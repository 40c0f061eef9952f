// mccba_ds.cuh -- double-sided board calibration on the device.
//
// Replaces the optimisation of cv::multicalib::DoubleSideCalibration (src/doubleSide.cpp): cameras FIXED at known poses,
// unknowns = the front<->back transform D of the board (6, shared by every edge that sees the back pattern) + one pose
// per frame (6 each); parameter vector [D | frame poses] as buildParas builds it (src/doubleSide.cpp:233-261); per-edge
// chain of computePhotoCameraJacobian (:288-429): front pose = camera o frame, and for a back-pattern edge a second
// compose with D innermost,  X_cam = R_c (R_p (R_D X + t_D) + t_p) + t_c ;  assembly of computeJacobianExtrinsic
// (:434-581: D in columns 0..5, frame f in columns 6(f+1)..); loop, step schedule G = 0.95^(iter+1) x and termination
// inherited from MultiCameraCalibration::optimizeExtrinsics (src/multicalib.cpp:462-514).  The reference marks this
// class "not working" (README.md:22); what is reproduced is its mathematics.
//
// Formulation (left-perturbation tangents, as in the rig path): with the per-edge block (H', g') the residual kernel
// accumulates for the COMPOSED pose (R3', T3') of the edge,
//   front edge:  (R3', T3') = (R3, T3) = camera o frame;   frame block  = lift_frame(H', g', R_c)
//   back edge:   R3' = R3 R_D,  T3' = R3 t_D + T3,  d = R3 t_D:
//       frame:   (psi', tau') = [[I, 0], [-[d]x, I]] (psi3, tau3)   -> the lever transform of lift_camera with s = d, then lift_frame
//       D:       (psi', tau') = blockdiag(R3, R3) (psi_D, tau_D)    -> lift_frame(H', g', R3)
//       cross:   W = blockdiag(R_c)^T L^T H' blockdiag(R3)
// The frame blocks are eliminated (Schur complement), the 6 x 6 system of D is solved, the steps are mapped back to
// additive Rodrigues increments with J_l^-1 and added like the reference adds them.
//
//   ds_pose_kernel     thread per frame slot: composed (primed) edge poses -> Problem::erec, which the residual /
//                      Jacobian kernels of the rig path evaluate unchanged (resid_jac_accum_*_kernel, forced mode)
//   ds_schur_kernel    thread per frame slot: lifts, 6 x 6 Cholesky, Y = L^-1 W, Schur record (SoA, no atomics)
//   ds_reduce_kernel   fixed-order sums of the record entries over the slots
//   ds_solve_kernel    loop control + 6 x 6 solve + update of D (one thread)
//   ds_update_kernel   back-substitution and update of the frame poses, norm partials
//   ds_decide_kernel   change = |G| / |params| (src/multicalib.cpp:502-504), iteration count, termination
#pragma once
#include "mccba_kernels.cuh"

namespace mccba {

constexpr int kDsRec = 44;    // S 36 | g 6 | cost | bad
constexpr int kDsSave = 63;   // U 21 | z 6 | Y 36

struct DsState {
    int crit_type, max_count, iter, done, status;
    double eps, change, alpha, cost;
    double dD[6];               // tangent step of D (unscaled)
    double step2_D, par2_D;     // |G_D|^2, |D_new|^2
};

struct DsProblem {
    const unsigned char* back;  // n_edge_int: 1 = the edge sees the back pattern
    const double* cam_R;        // 9 n_cam: fixed camera rotations
    const double* cam_t;        // 3 n_cam
    double* par;                // 6 + 6 n_frame: [D | frame poses], each [rvec | tvec]
    double* rec;                // kDsRec x n_slots (SoA)
    double* save;               // kDsSave x n_slots (SoA)
    double* tot;                // kDsRec
    double* norm_part;          // 2 x n_blocks
    int n_blocks;               // blocks of 128 frame slots
    DsState* st;
};

// out += blockdiag(Rl, Rl)^T A blockdiag(Rr, Rr)  (6 x 6, row-major)
__device__ __forceinline__ void ds_congruence_acc(const double* A, const double* Rl, const double* Rr, double* out)
{
    double T[36];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                T[i * 6 + 3 * b + j] = A[i * 6 + 3 * b] * Rr[j] + A[i * 6 + 3 * b + 1] * Rr[3 + j] + A[i * 6 + 3 * b + 2] * Rr[6 + j];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
                out[(3 * a + i) * 6 + j] += Rl[i] * T[(3 * a) * 6 + j] + Rl[3 + i] * T[(3 * a + 1) * 6 + j] + Rl[6 + i] * T[(3 * a + 2) * 6 + j];
}

struct DsSlot {
    int frame, V, ls;
    const int* gc;
    int64_t ebase, stride;
};
__device__ __forceinline__ DsSlot ds_slot(const Problem& P, int slot)
{
    const int warp = slot >> 5, lane = slot & 31;
    const int4 m0 = P.wmeta[2 * warp], m1 = P.wmeta[2 * warp + 1];
    DsSlot s;
    s.frame = P.slot_frame[slot];
    s.V = m0.x;
    s.gc = P.group_cams + m0.y;
    s.ls = m1.x + lane;
    s.ebase = m0.z;
    s.stride = m0.w;
    return s;
}

__global__ void __launch_bounds__(128) ds_pose_kernel(Problem P, DsProblem D, int forced)
{
    if (!forced && D.st->done) return;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.n_slots) return;
    const DsSlot s = ds_slot(P, slot);
    double Rd[9], Rp[9];
    rodrigues(D.par, Rd);
    const double td[3] = {D.par[3], D.par[4], D.par[5]};
    const double* q = D.par + 6 + 6 * (int64_t)(s.frame >= 0 ? s.frame : 0);
    rodrigues(q, Rp);
    for (int v = 0; v < s.V; ++v) {
        const int64_t e = s.ebase + v * s.stride + s.ls;
        EdgeRec er;
        if (s.frame < 0) {      // padding slot: a benign pose (the packed pass evaluates its zero-weight corners)
#pragma unroll
            for (int i = 0; i < 9; ++i) er.R3[i] = 0;
            er.T3[0] = er.T3[1] = 0; er.T3[2] = 1;
        } else {
            const int c = s.gc[v];
            compose_pose(D.cam_R + 9 * c, D.cam_t + 3 * c, Rp, q + 3, er.R3, er.T3);
            if (D.back[e]) {
                double d[3], R2[9];
                mat3_vec(er.R3, td, d);
                mat3_mul(er.R3, Rd, R2);
#pragma unroll
                for (int i = 0; i < 3; ++i) er.T3[i] += d[i];
#pragma unroll
                for (int i = 0; i < 9; ++i) er.R3[i] = R2[i];
            }
        }
        P.erec[e] = er;
    }
}

__global__ void __launch_bounds__(128) ds_schur_kernel(Problem P, DsProblem D, int forced)
{
    if (!forced && D.st->done) return;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.n_slots) return;
    const DsSlot s = ds_slot(P, slot);
    const int64_t ns = P.n_slots;
    const double* __restrict__ blk = P.blocks[P.st->cur];
    double U[21], z[6], Hdd[36], gd[6], W[36], cost = 0;
#pragma unroll
    for (int i = 0; i < 21; ++i) U[i] = 0;
#pragma unroll
    for (int i = 0; i < 36; ++i) { Hdd[i] = 0; W[i] = 0; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { z[i] = 0; gd[i] = 0; }
    int bad = 0;
    if (s.frame >= 0) {
        const double* q = D.par + 6 + 6 * (int64_t)s.frame;
        double Rp[9];
        rodrigues(q, Rp);
        const double td[3] = {D.par[3], D.par[4], D.par[5]};
        for (int v = 0; v < s.V; ++v) {
            const int64_t e = s.ebase + v * s.stride + s.ls;
            const int c = s.gc[v];
            double t[kBlk], H[36];
#pragma unroll
            for (int k = 0; k < kBlk; ++k) t[k] = blk[tile_idx(kBlk, e, k)];
            cost += t[27];
            unpack_sym6(t, H);
            double Rc[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Rc[i] = D.cam_R[9 * c + i];
            if (!D.back[e]) {
                lift_frame(H, t + 21, Rc, 0, U, z);
            } else {
                double R3[9], d[3], H3[36], g3[6], Wc[36], Ud[21], N[36];
                mat3_mul(Rc, Rp, R3);
                mat3_vec(R3, td, d);
                // frame side: lever transform (lift_camera with s = d), then the rotation into frame coordinates
                lift_camera(H, t + 21, Rc, d, H3, g3, Wc);
                lift_frame(H3, g3, Rc, 0, U, z);
                // D side
#pragma unroll
                for (int i = 0; i < 21; ++i) Ud[i] = 0;
                lift_frame(H, t + 21, R3, 0, Ud, gd);
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = i; j < 6; ++j) { Hdd[i * 6 + j] += Ud[tri6(i, j)]; if (j > i) Hdd[j * 6 + i] += Ud[tri6(i, j)]; }
                // cross block: N = L^T H' (rows psi get d x rows tau), W += blockdiag(Rc)^T N blockdiag(R3)
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double bot[3] = {H[18 + j], H[24 + j], H[30 + j]};
                    double cx[3];
                    cross3(d, bot, cx);
                    N[j] = H[j] + cx[0]; N[6 + j] = H[6 + j] + cx[1]; N[12 + j] = H[12 + j] + cx[2];
                    N[18 + j] = bot[0]; N[24 + j] = bot[1]; N[30 + j] = bot[2];
                }
                ds_congruence_acc(N, Rc, R3, W);
            }
        }
        if (!chol6_packed(U)) bad = 1;
        chol6_forward(U, z, 1);
#pragma unroll
        for (int j = 0; j < 6; ++j) chol6_forward(U, W + j, 6);      // Y = L^-1 W, column by column
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) U[tri6(i, i)] = 1.0;
    }
    // Schur record of the frame: S = H_DD - Y^T Y, g = g_D - Y^T z
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc += W[k * 6 + i] * W[k * 6 + j];
            D.rec[(int64_t)(i * 6 + j) * ns + slot] = Hdd[i * 6 + j] - acc;
        }
        double acc = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) acc += W[k * 6 + i] * z[k];
        D.rec[(int64_t)(36 + i) * ns + slot] = gd[i] - acc;
    }
    D.rec[(int64_t)42 * ns + slot] = cost;
    D.rec[(int64_t)43 * ns + slot] = (double)bad;
#pragma unroll
    for (int k = 0; k < 21; ++k) D.save[(int64_t)k * ns + slot] = U[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) D.save[(int64_t)(21 + k) * ns + slot] = z[k];
#pragma unroll
    for (int k = 0; k < 36; ++k) D.save[(int64_t)(27 + k) * ns + slot] = W[k];
}

// one block per record entry: fixed-order sum over the frame slots
__global__ void __launch_bounds__(256) ds_reduce_kernel(Problem P, DsProblem D, int forced)
{
    if (!forced && D.st->done) return;
    __shared__ double sm[8];
    const int v = blockIdx.x;
    const double* src = D.rec + (int64_t)v * P.n_slots;
    double s = 0.0;
    for (int f = threadIdx.x; f < P.n_slots; f += blockDim.x) s += src[f];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        D.tot[v] = t;
    }
}

__global__ void ds_solve_kernel(Problem P, DsProblem D)
{
    if (threadIdx.x != 0) return;
    DsState* st = D.st;
    if (st->done) return;
    const double* tot = D.tot;
    const double cost = tot[42];
    if (tot[43] != 0.0 || !isfinite(cost)) { st->status = 4; st->done = 1; return; }
    st->cost = cost;
    double U[21], x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) U[tri6(i, j)] = tot[i * 6 + j];
        x[i] = tot[36 + i];
    }
    if (!chol6_packed(U)) { st->status = 4; st->done = 1; return; }
    chol6_forward(U, x, 1);
    chol6_backward(U, x);
    const double alpha = pow(0.95, (double)st->iter + 1.0);      // src/multicalib.cpp:482-483
    st->alpha = alpha;
    const double om[3] = {D.par[0], D.par[1], D.par[2]};
    double dom[3], s2 = 0, p2 = 0;
    left_jacobian_inv_apply(om, x, dom);
#pragma unroll
    for (int i = 0; i < 6; ++i) st->dD[i] = x[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double g0 = alpha * dom[i], g1 = alpha * x[3 + i];
        const double n0 = D.par[i] + g0, n1 = D.par[3 + i] + g1;
        D.par[i] = n0; D.par[3 + i] = n1;
        s2 += g0 * g0 + g1 * g1; p2 += n0 * n0 + n1 * n1;
    }
    st->step2_D = s2; st->par2_D = p2;
}

__global__ void __launch_bounds__(128) ds_update_kernel(Problem P, DsProblem D)
{
    const DsState* st = D.st;
    if (st->done) return;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ns = P.n_slots;
    double step2 = 0, par2 = 0;
    if (slot < P.n_slots) {
        const int frame = P.slot_frame[slot];
        if (frame >= 0) {
            const double alpha = st->alpha;
            double U[21], r[6];
#pragma unroll
            for (int k = 0; k < 21; ++k) U[k] = D.save[(int64_t)k * ns + slot];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double acc = D.save[(int64_t)(21 + i) * ns + slot];
#pragma unroll
                for (int j = 0; j < 6; ++j) acc -= D.save[(int64_t)(27 + i * 6 + j) * ns + slot] * st->dD[j];
                r[i] = acc;
            }
            chol6_backward(U, r);
            double* q = D.par + 6 + 6 * (int64_t)frame;
            const double om[3] = {q[0], q[1], q[2]};
            double dom[3];
            left_jacobian_inv_apply(om, r, dom);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double g0 = alpha * dom[i], g1 = alpha * r[3 + i];
                const double n0 = q[i] + g0, n1 = q[3 + i] + g1;
                q[i] = n0; q[3 + i] = n1;
                step2 += g0 * g0 + g1 * g1; par2 += n0 * n0 + n1 * n1;
            }
        }
    }
    __shared__ double sr[2][4];
    step2 = warp_sum(step2); par2 = warp_sum(par2);
    if ((threadIdx.x & 31) == 0) { sr[0][threadIdx.x >> 5] = step2; sr[1][threadIdx.x >> 5] = par2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        D.norm_part[blockIdx.x] = sr[0][0] + sr[0][1] + sr[0][2] + sr[0][3];
        D.norm_part[D.n_blocks + blockIdx.x] = sr[1][0] + sr[1][1] + sr[1][2] + sr[1][3];
    }
}

__global__ void __launch_bounds__(256) ds_decide_kernel(DsProblem D)
{
    DsState* st = D.st;
    if (st->done) return;
    __shared__ double sm[2][8];
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < D.n_blocks; k += blockDim.x) { a += D.norm_part[k]; b += D.norm_part[D.n_blocks + k]; }
    a = warp_sum(a); b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g2 = st->step2_D, p2 = st->par2_D;
        for (int w = 0; w < 8; ++w) { g2 += sm[0][w]; p2 += sm[1][w]; }
        st->change = sqrt(g2) / sqrt(p2);        // the norm of the parameters AFTER the update (src/multicalib.cpp:501-504)
        st->iter += 1;
        const int t = st->crit_type;
        if ((t == 1 && st->iter >= st->max_count) || (t == 2 && st->change <= st->eps) ||
            (t == 3 && (st->change <= st->eps || st->iter >= st->max_count)))
            st->done = 1;
    }
}

__global__ void ds_init_state_kernel(DsState* st, int crit_type, int max_count, double eps)
{
    st->crit_type = crit_type; st->max_count = max_count; st->iter = 0; st->status = 0;
    st->eps = eps; st->change = 1.0; st->alpha = 0; st->cost = 0; st->step2_D = 0; st->par2_D = 0;
    for (int i = 0; i < 6; ++i) st->dD[i] = 0;
    // loop-top test of src/multicalib.cpp:475-477 with iter = 0, change = 1
    st->done = (crit_type == 1 && 0 >= max_count) || (crit_type == 3 && (1.0 <= eps || 0 >= max_count)) || (crit_type == 2 && 1.0 <= eps) ? 1 : 0;
}

}  // namespace mccba
