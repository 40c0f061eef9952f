// TEST INFRASTRUCTURE: host-side unit-test harness for the product's device arithmetic.
// Compiles multi_camera_calibration_b200/csrc/mccba_math.cuh with g++ (the functions are __host__ __device__) and
// re-enacts the kernels' data flow sequentially, so the tangent-space formulation can be checked against the oracle
// on a machine without a GPU.  It is NOT reachable from the product API and is not a fallback path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../multi_camera_calibration_b200/csrc/mccba_math.cuh"
#include "../../multi_camera_calibration_b200/csrc/mccba_f32x2.cuh"

using namespace mccba;

static CamParams make_cam(int model, const double* K5, const double* d8, int nd, double xi)
{
    CamParams p;
    memset(&p, 0, sizeof(p));
    p.model = model; p.fx = K5[0]; p.fy = K5[1]; p.cx = K5[2]; p.cy = K5[3]; p.skew = K5[4]; p.xi = xi;
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nd; ++i) k[i] = d8[i];
    p.k1 = k[0]; p.k2 = k[1]; p.p1 = k[2]; p.p2 = k[3]; p.k3 = k[4]; p.k4 = k[5]; p.k5 = k[6]; p.k6 = k[7];
    p.rational = model == kPinhole && (k[5] != 0 || k[6] != 0 || k[7] != 0);
    return p;
}

static void edge_block(const CamParams& cam, const double* Rc, const double* tc, const double* Rp, const double* tp,
                       int64_t b, int64_t e, const float* obj, const float* img, double* acc)
{
    double R3[9], T3[3];
    compose_pose(Rc, tc, Rp, tp, R3, T3);
    for (int k = 0; k < kBlk; ++k) acc[k] = 0;
    for (int64_t i = b; i < e; ++i) {
        if (cam.model == kPinhole) {
            if (cam.rational) corner_accumulate<kPinhole, true>(cam, R3, T3, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1], acc);
            else corner_accumulate<kPinhole, false>(cam, R3, T3, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1], acc);
        } else corner_accumulate<kOmnidir, false>(cam, R3, T3, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1], acc);
    }
}

// The packed single-precision pass re-enacted in the device kernel's own order: 4 lanes per edge, lane q takes corners
// 8k + 2q, 8k + 2q + 1 in step k as one f32x2 pair; halves, then lanes (q, q^2), then (q, q^1) are added in float; the
// 28 sums are promoted to double (resid_jac_accum_f32_kernel).  exact_e: MIXED policy (residual from the double
// projection); otherwise the all-float32 variant.  kPlanar: the flat-board variant of both projections (object z = 0).
template <bool kExactE, bool kPlanar>
static void edge_block_f32(const CamParams& cam, const double* Rc, const double* tc, const double* Rp, const double* tp,
                           int64_t b, int64_t e, const float* obj, const float* img, double* acc)
{
    double R3d[9], T3d[3];
    compose_pose(Rc, tc, Rp, tp, R3d, T3d);
    float R3[9], T3[3];
    for (int i = 0; i < 9; ++i) R3[i] = (float)R3d[i];
    for (int i = 0; i < 3; ++i) T3[i] = (float)T3d[i];
    const CamF2 c2 = make_cam_f2(cam);
    const int n = (int)(e - b), kp = (n + 7) / 8;
    float v[4][kBlk];
    for (int q = 0; q < 4; ++q) {
        f2 a[kBlk];
        for (int k = 0; k < kBlk; ++k) a[k] = f2_dup(0.0f);
        for (int k = 0; k < kp; ++k) {
            const int c0 = 8 * k + 2 * q;
            float o[5][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
            float ex[2][2] = {{0, 0}, {0, 0}};
            for (int h = 0; h < 2; ++h)
                if (c0 + h < n) {
                    const int64_t i = b + c0 + h;
                    o[0][h] = obj[3 * i]; o[1][h] = obj[3 * i + 1]; o[2][h] = obj[3 * i + 2]; o[3][h] = img[2 * i]; o[4][h] = img[2 * i + 1];
                }
            if (kExactE)
                for (int h = 0; h < 2; ++h) {
                    double ed[2];
                    const CamResid camr = make_cam_resid(cam);
                    if (cam.model == kPinhole) {
                        if (cam.rational) corner_residual<kPinhole, true, kPlanar>(camr, R3d, T3d, o[0][h], o[1][h], o[2][h], o[3][h], o[4][h], ed);
                        else corner_residual<kPinhole, false, kPlanar>(camr, R3d, T3d, o[0][h], o[1][h], o[2][h], o[3][h], o[4][h], ed);
                    } else corner_residual<kOmnidir, false, kPlanar>(camr, R3d, T3d, o[0][h], o[1][h], o[2][h], o[3][h], o[4][h], ed);
                    ex[0][h] = (float)ed[0]; ex[1][h] = (float)ed[1];
                }
            const f2 w = f2_make(c0 < n ? 1.0f : 0.0f, c0 + 1 < n ? 1.0f : 0.0f);
            const bool masked = 8 * k + 8 > n;
            const f2 X0 = f2_make(o[0][0], o[0][1]), X1 = f2_make(o[1][0], o[1][1]), X2 = f2_make(o[2][0], o[2][1]);
            const f2 U = f2_make(o[3][0], o[3][1]), V = f2_make(o[4][0], o[4][1]);
            const f2 e0 = f2_make(ex[0][0], ex[0][1]), e1 = f2_make(ex[1][0], ex[1][1]);
            if (cam.model == kPinhole) {
                if (cam.rational) corner_pair_accumulate<kPinhole, true, kExactE, kPlanar>(c2, R3, T3, X0, X1, X2, U, V, w, masked, a, e0, e1);
                else corner_pair_accumulate<kPinhole, false, kExactE, kPlanar>(c2, R3, T3, X0, X1, X2, U, V, w, masked, a, e0, e1);
            } else corner_pair_accumulate<kOmnidir, false, kExactE, kPlanar>(c2, R3, T3, X0, X1, X2, U, V, w, masked, a, e0, e1);
        }
        for (int k = 0; k < kBlk; ++k) v[q][k] = f2_lo(a[k]) + f2_hi(a[k]);
    }
    for (int k = 0; k < kBlk; ++k) acc[k] = (double)((v[0][k] + v[2][k]) + (v[1][k] + v[3][k]));
}

extern "C" {

// per-edge blocks (E x 28) at params, plus the full normal-equation solve in tangent coordinates:
// S (ns x ns), gs (ns), step (6(nV-1), additive rvec/tvec step, unscaled).  Returns 0, or 1 if not SPD.
int hm_rig_step(int n_cam, int n_frame, int n_edge, const int* edge_cam, const int* edge_pv, const int64_t* edge_off,
                const float* obj, const float* img, const int* cam_model, const double* K5, const double* d8,
                const int* nd, const double* xi, const double* params, double lambda, double* blocks, double* S,
                double* gs, double* step, int policy)
{
    const int nV = n_cam + n_frame, ns = 6 * (n_cam - 1);
    std::vector<CamParams> cams;
    for (int c = 0; c < n_cam; ++c) cams.push_back(make_cam(cam_model[c], K5 + 5 * c, d8 + 8 * c, nd[c], xi[c]));
    std::vector<double> vR(9 * (size_t)nV), vt(3 * (size_t)nV, 0.0);
    for (int v = 0; v < nV; ++v) {
        double om[3] = {0, 0, 0};
        if (v > 0) {
            for (int i = 0; i < 3; ++i) { om[i] = params[6 * (v - 1) + i]; vt[3 * v + i] = params[6 * (v - 1) + 3 + i]; }
        }
        rodrigues(om, &vR[9 * v]);
    }
    for (int e = 0; e < n_edge; ++e) {
        const CamParams& cm = cams[edge_cam[e]];
        const double *Rc = &vR[9 * edge_cam[e]], *tc = &vt[3 * edge_cam[e]], *Rp = &vR[9 * edge_pv[e]], *tp = &vt[3 * edge_pv[e]];
        double* blk = blocks + (size_t)kBlk * e;
        if (policy == 1) edge_block_f32<true, false>(cm, Rc, tc, Rp, tp, edge_off[e], edge_off[e + 1], obj, img, blk);        // MIXED
        else if (policy == 2) edge_block_f32<false, false>(cm, Rc, tc, Rp, tp, edge_off[e], edge_off[e + 1], obj, img, blk);  // all float32
        else if (policy == 3) edge_block_f32<true, true>(cm, Rc, tc, Rp, tp, edge_off[e], edge_off[e + 1], obj, img, blk);    // MIXED, flat-board path
        else if (policy == 4) edge_block_f32<false, true>(cm, Rc, tc, Rp, tp, edge_off[e], edge_off[e + 1], obj, img, blk);   // all float32, flat-board path
        else edge_block(cm, Rc, tc, Rp, tp, edge_off[e], edge_off[e + 1], obj, img, blk);
    }
    for (int i = 0; i < ns * ns; ++i) S[i] = 0;
    for (int i = 0; i < ns; ++i) gs[i] = 0;
    std::vector<double> U(21 * (size_t)n_frame), Z(6 * (size_t)n_frame), Y(36 * (size_t)n_edge, 0.0);
    int bad = 0;
    for (int f = 0; f < n_frame; ++f) {
        const int pv = n_cam + f;
        double* u = &U[21 * f];
        double* z = &Z[6 * f];
        for (int i = 0; i < 21; ++i) u[i] = 0;
        for (int i = 0; i < 6; ++i) z[i] = 0;
        std::vector<int> es;
        for (int e = 0; e < n_edge; ++e)
            if (edge_pv[e] == pv) es.push_back(e);
        for (int e : es) {
            double H[36];
            unpack_sym6(blocks + (size_t)kBlk * e, H);
            lift_frame(H, blocks + (size_t)kBlk * e + 21, &vR[9 * edge_cam[e]], edge_cam[e] == 0, u, z);
        }
        for (int i = 0; i < 6; ++i) u[tri6(i, i)] *= (1.0 + lambda);
        if (!chol6_packed(u)) bad = 1;
        chol6_forward(u, z, 1);
        for (int e : es) {
            const int c = edge_cam[e];
            if (c == 0) continue;
            double H[36], D[36], gcv[6], s[3];
            unpack_sym6(blocks + (size_t)kBlk * e, H);
            mat3_vec(&vR[9 * c], &vt[3 * pv], s);
            double* y = &Y[36 * (size_t)e];
            lift_camera(H, blocks + (size_t)kBlk * e + 21, &vR[9 * c], s, D, gcv, y);
            for (int j = 0; j < 6; ++j) chol6_forward(u, y + j, 6);
            for (int i = 0; i < 6; ++i) D[i * 6 + i] *= (1.0 + lambda);
            for (int i = 0; i < 6; ++i) {
                for (int j = 0; j < 6; ++j) {
                    double a = 0;
                    for (int k = 0; k < 6; ++k) a += y[k * 6 + i] * y[k * 6 + j];
                    S[(6 * (c - 1) + i) * ns + 6 * (c - 1) + j] += D[i * 6 + j] - a;
                }
                double a = 0;
                for (int k = 0; k < 6; ++k) a += y[k * 6 + i] * z[k];
                gs[6 * (c - 1) + i] += gcv[i] - a;
            }
        }
        for (size_t a = 0; a < es.size(); ++a)
            for (size_t b = 0; b < es.size(); ++b) {
                const int ca = edge_cam[es[a]], cb = edge_cam[es[b]];
                if (a == b || ca == 0 || cb == 0) continue;
                const double *ya = &Y[36 * (size_t)es[a]], *yb = &Y[36 * (size_t)es[b]];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) {
                        double acc = 0;
                        for (int k = 0; k < 6; ++k) acc += ya[k * 6 + i] * yb[k * 6 + j];
                        S[(6 * (ca - 1) + i) * ns + 6 * (cb - 1) + j] -= acc;
                    }
            }
    }
    // dense Cholesky solve of S dc = gs
    std::vector<double> L(S, S + (size_t)ns * ns), dc(gs, gs + ns);
    for (int j = 0; j < ns; ++j) {
        double d = L[j * ns + j];
        for (int k = 0; k < j; ++k) d -= L[j * ns + k] * L[j * ns + k];
        if (!(d > 0)) { bad = 1; d = 1; }
        d = sqrt(d);
        L[j * ns + j] = d;
        for (int i = j + 1; i < ns; ++i) {
            double s = L[i * ns + j];
            for (int k = 0; k < j; ++k) s -= L[i * ns + k] * L[j * ns + k];
            L[i * ns + j] = s / d;
        }
    }
    for (int i = 0; i < ns; ++i) {
        double s = dc[i];
        for (int k = 0; k < i; ++k) s -= L[i * ns + k] * dc[k];
        dc[i] = s / L[i * ns + i];
    }
    for (int i = ns - 1; i >= 0; --i) {
        double s = dc[i];
        for (int k = i + 1; k < ns; ++k) s -= L[k * ns + i] * dc[k];
        dc[i] = s / L[i * ns + i];
    }
    for (int c = 1; c < n_cam; ++c) {
        double dom[3];
        left_jacobian_inv_apply(params + 6 * (c - 1), &dc[6 * (c - 1)], dom);
        for (int i = 0; i < 3; ++i) { step[6 * (c - 1) + i] = dom[i]; step[6 * (c - 1) + 3 + i] = dc[6 * (c - 1) + 3 + i]; }
    }
    for (int f = 0; f < n_frame; ++f) {
        const int pv = n_cam + f;
        double r[6];
        for (int i = 0; i < 6; ++i) r[i] = Z[6 * f + i];
        for (int e = 0; e < n_edge; ++e) {
            if (edge_pv[e] != pv || edge_cam[e] == 0) continue;
            const double* y = &Y[36 * (size_t)e];
            for (int i = 0; i < 6; ++i)
                for (int k = 0; k < 6; ++k) r[i] -= y[i * 6 + k] * dc[6 * (edge_cam[e] - 1) + k];
        }
        chol6_backward(&U[21 * f], r);
        double dom[3];
        left_jacobian_inv_apply(params + 6 * (pv - 1), r, dom);
        for (int i = 0; i < 3; ++i) { step[6 * (pv - 1) + i] = dom[i]; step[6 * (pv - 1) + 3 + i] = r[3 + i]; }
    }
    return bad;
}

// Mei projection of n points with the reference's 2n x 16 Jacobian layout (dom dT df ds dc dxi dkp), built from the
// product's tangent-space pieces: d/dom = (Q x a) J_l(om).
void hm_omni_project(int n, const double* obj, const double* om, const double* T, const double* K5, double xi,
                     const double* D4, double* proj, double* jac16)
{
    double d8[8] = {D4[0], D4[1], D4[2], D4[3], 0, 0, 0, 0};
    CamParams cam = make_cam(kOmnidir, K5, d8, 4, xi);
    double R[9], Jl[9];
    rodrigues(om, R);
    left_jacobian(om, Jl);
    for (int i = 0; i < n; ++i) {
        double Q[3], Xc[3], A[6], Jin[20];
        mat3_vec(R, obj + 3 * i, Q);
        for (int k = 0; k < 3; ++k) Xc[k] = Q[k] + T[k];
        omnidir_point_full(cam, Xc, proj + 2 * i, A, Jin);
        for (int r = 0; r < 2; ++r) {
            double jphi[3];
            cross3(Q, A + 3 * r, jphi);
            double* row = jac16 + (size_t)(2 * i + r) * 16;
            for (int k = 0; k < 3; ++k) row[k] = jphi[0] * Jl[k] + jphi[1] * Jl[3 + k] + jphi[2] * Jl[6 + k];
            for (int k = 0; k < 3; ++k) row[3 + k] = A[3 * r + k];
            for (int k = 0; k < 10; ++k) row[6 + k] = Jin[10 * r + k];
        }
    }
}

void hm_rodrigues(const double* om, double* R) { rodrigues(om, R); }
void hm_left_jacobian_inv_apply(const double* om, const double* psi, double* out) { left_jacobian_inv_apply(om, psi, out); }
}
