/*
 * mccba.h -- C ABI of the B200-native calibration bundle-adjustment core ("multi-camera calibration BA").
 *
 * This is the drop-in boundary for the hot path of yulong314/multi_camera_calibration: the iteration loop of
 * cv::multicalib::MultiCameraCalibration::optimizeExtrinsics and everything it calls.  The reference has no
 * FFI; its seam is C++ virtual dispatch on MultiCameraCalibration (include/opencv2/ccalib/multicalib.hpp:
 * 138-191).  Each entry point below names the reference member(s) it replaces; the host-side C++17 class in
 * include/mccba_host.hpp mirrors the reference class on top of this ABI, and INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++ / torch / OpenCV types.
 *  - every function returns an int status (MCCBA_OK == 0) and never throws; mccba_last_error(handle) returns a
 *    message owned by the handle (valid until the next call on that handle).
 *  - host buffers are caller-owned and COPIED by the set_* calls; device memory is owned by the handle.
 *  - one handle = one CUDA device + one stream.  A handle is not thread-safe; distinct handles are independent.
 *  - there is no CPU fallback: without a CUDA device mccba_create fails with MCCBA_ERR_CUDA.
 *
 * Indexing contract (bit-exact with the reference):
 *  - vertices 0..nC-1 are cameras, vertices nC..nC+F-1 are pattern poses ("photo vertices") in first-seen order;
 *  - edge e connects cameraVertex edge_cam[e] and photoVertex edge_pv[e] (multicalib.hpp:86-103);
 *  - the parameter vector has 6*(nC+F-1) doubles: [rvec | tvec] of vertex v at 6*(v-1); vertex 0 (camera 0) is
 *    the gauge (src/multicalib.cpp:422-440, 636-640);
 *  - X_cam = R_cam (R_photo X + t_photo) + t_cam (call order at src/multicalib.cpp:734);
 *  - corner i of edge e is row pair (2i, 2i+1) of the reference's Jacobian block starting at 2*edge_off[e]
 *    (pointsLocation, src/multicalib.cpp:597-603).
 *
 * Multi-GPU: frames shard across ranks.  Each rank creates its own handle (rank, nranks, shared ncclUniqueId),
 * passes ALL cameras and only ITS frames/edges (photo vertices renumbered nC..nC+F_local-1), and calls
 * mccba_solve collectively.  One exchange per iteration sums the reduced camera system (a kernel over NVLink peer
 * memory when every rank can map every other rank's window, else ncclAllReduce: mccba_exchange_mode); every rank solves
 * it redundantly.  Camera parameters are replicated, frame parameters stay on their rank.
 */
#ifndef MCCBA_H_
#define MCCBA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCCBA_VERSION 1

/* status codes */
#define MCCBA_OK 0
#define MCCBA_ERR_ARG 1        /* bad argument / inconsistent sizes / unsupported model (CV_Assert in the reference) */
#define MCCBA_ERR_CUDA 2       /* CUDA runtime error or no device */
#define MCCBA_ERR_STATE 3      /* call order violated (e.g. solve before set_observations) */
#define MCCBA_ERR_NUMERIC 4    /* non-finite cost or reduced system not positive definite */
#define MCCBA_ERR_NCCL 5       /* NCCL unavailable or collective failed */

/* camera models: values of MultiCameraCalibration::PINHOLE / OMNIDIRECTIONAL (multicalib.hpp:76-80) */
#define MCCBA_PINHOLE 0
#define MCCBA_OMNIDIRECTIONAL 1

/* cv::TermCriteria type bits as decoded at src/multicalib.cpp:475-477 */
#define MCCBA_CRIT_COUNT 1
#define MCCBA_CRIT_EPS 2

/* solver modes */
#define MCCBA_MODE_REFERENCE_GN 0 /* step-scaled Gauss-Newton, G = 0.95^(iter+1) x (src/multicalib.cpp:482-504) */
#define MCCBA_MODE_LM 1           /* Levenberg-Marquardt: damping, accept/reject on device (north star) */

/* precision policy of the residual / Jacobian pass (mccba_set_precision).  Everything summed over many observations
 * (per-edge blocks onwards: Schur complement, reduced solve, update, cost test, the reported RMS) is fp64 in both. */
#define MCCBA_PRECISION_FP64 0  /* per-corner projection, Jacobian and accumulation in double */
#define MCCBA_PRECISION_MIXED 1 /* residual (projection) and cost in double; Jacobian, J^T J and J^T e products in
                                   packed float32 (two corners per lane), per-edge sums promoted to double.  Final parameters
                                   agree with FP64 to ~1e-8 relative; the worst of the 600 378 parameters of config #5 -- the
                                   tilt of a board that faces a camera squarely -- to 1e-6 (tests/test_precision_gpu.py,
                                   tests/test_full_size_gpu.py).  The error grows with the conditioning of the rig: boards
                                   that subtend less than ~8 degrees (extent / distance < 0.15) exceed 1e-6 -- select
                                   MCCBA_PRECISION_FP64 for those (profiles/r2_precision_vs_board.txt) */
#define MCCBA_PRECISION_FAST32 2 /* everything per corner in packed float32, like the reference, which evaluates the
                                   projection through float32 (src/multicalib.cpp:742-749, 789-792): RMS agrees to 1e-9,
                                   but the tilt of boards that face a camera squarely moves by ~2e-6 -- outside the 1e-6
                                   parity gate, hence opt-in */
#define MCCBA_PRECISION_AUTO 3   /* default: MIXED where it is safe, FP64 otherwise, decided for every solve / evaluation from the
                                   geometry at the current parameters: MIXED iff every image sees its board under an angular
                                   extent (diameter of the object points / distance to the camera) of at least 0.15; below that
                                   one view no longer pins the tilt of the board and the float32 Jacobian products of MIXED would
                                   show above 1e-6 (profiles/r2_precision_vs_board.txt).  Both observation layouts are kept on
                                   the device; mccba_effective_precision tells what ran last. */

typedef struct mccba_handle_s *mccba_handle;

typedef struct {
    int device;                 /* CUDA device ordinal */
    int rank, nranks;           /* frame-sharding rank / world size (1 = single GPU) */
    unsigned char nccl_id[128]; /* ncclUniqueId bytes, identical on all ranks; ignored when nranks == 1 */
    int use_graph;              /* 1: capture one LM iteration in a CUDA graph (default); 0: plain stream launches */
    int verbose;
} mccba_options;

typedef struct {
    int mode;           /* MCCBA_MODE_* */
    int crit_type;      /* MCCBA_CRIT_COUNT | MCCBA_CRIT_EPS, same decoding as the reference */
    int max_count;      /* criteria.maxCount */
    double epsilon;     /* criteria.epsilon, compared with change = |step| / |params| */
    double lambda0;     /* LM: initial damping (H + lambda diag H), tangent-space diag; default 1e-3 */
    double lambda_up;   /* LM: factor on reject (default 10) */
    double lambda_down; /* LM: factor on accept (default 1/3) */
} mccba_solve_opts;

typedef struct {
    int iterations;     /* trial evaluations performed (== the reference's iter counter at loop exit) */
    int accepted, rejected;
    int status;         /* MCCBA_OK or MCCBA_ERR_NUMERIC */
    int graph_launches; /* iteration graphs (or stream sequences) enqueued */
    int kernel_launches;/* CUDA kernels of this library launched by the call */
    double change;      /* last |step| / |params| */
    double cost;        /* sum of squared residuals at the returned parameters (all ranks) */
    double lambda;      /* final damping */
    double device_ms;   /* CUDA-event time of the whole call on the handle's stream */
} mccba_report;

typedef struct {
    double mean_reproj_error; /* the reference's meanReprojectError: sum ||e|| / totalNPoints, where PINHOLE edges
                                 count 2N points and OMNIDIRECTIONAL edges N (src/multicalib.cpp:983, 989) */
    double rms;               /* fp64 RMS sqrt(sum(ex^2+ey^2)/Npoints) (src/omnidir.cpp:1794-1802) */
    double sum_norm, sum_sq;  /* local (this rank's edges) sums; the two ratios above are global when nranks>1 */
    int64_t n_points;
} mccba_error_stats;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
/* fills defaults (device 0, single rank, graph on) */
int mccba_default_options(mccba_options *opts);
int mccba_default_solve_opts(mccba_solve_opts *opts);
/* rank 0 obtains an ncclUniqueId to broadcast to the other ranks (any transport); MCCBA_ERR_NCCL if libnccl absent */
int mccba_nccl_unique_id(unsigned char out[128]);
/* replaces the MultiCameraCalibration constructor's numerical state (src/multicalib.cpp:75-104) */
int mccba_create(const mccba_options *opts, mccba_handle *out);
int mccba_destroy(mccba_handle h);
const char *mccba_last_error(mccba_handle h);
/* Precision policy of the residual / Jacobian pass (MCCBA_PRECISION_*; default AUTO, or the environment variable
 * MCCBA_PRECISION=auto|fp64|mixed|fast32 read at mccba_create).  The observation layout on the device depends on it: changing the
 * policy discards the current problem (set_observations / set_parameters have to be called again).
 * Replaces the float32 conversions of src/multicalib.cpp:742-749. */
int mccba_set_precision(mccba_handle h, int policy);
int mccba_get_precision(mccba_handle h);
/* the policy the last solve / evaluation actually ran (MCCBA_PRECISION_FP64, _MIXED or _FAST32): differs from
 * mccba_get_precision only under MCCBA_PRECISION_AUTO */
int mccba_effective_precision(mccba_handle h);

/* ---- problem ----------------------------------------------------------------------------------------------- */
/* Replaces _cameraMatrix / _distortCoeffs / _xi (multicalib.hpp:211-213).  K5 = fx fy cx cy skew per camera
 * (skew is used by the Mei model only: cv::projectPoints ignores K(0,1)); dist8 = k1 k2 p1 p2 k3 k4 k5 k6 per
 * camera, ndist in {0,4,5,8} for PINHOLE and exactly 4 for OMNIDIRECTIONAL (src/omnidir.cpp:92). */
int mccba_set_cameras(mccba_handle h, int n_cam, const int *model, const double *K5, const double *dist8,
                      const int *ndist, const double *xi);
/* Replaces _edgeList + _objectPointsForEachCamera / _imagePointsForEachCamera (multicalib.hpp:207-210).
 * obj_xyz / img_uv are the reference's CV_32F points concatenated edge by edge (AoS, 3 / 2 floats per corner);
 * the library re-lays them out as SoA planes in HBM.  edge_off has n_edge+1 entries (in corners). */
int mccba_set_observations(mccba_handle h, int n_frame, int n_edge, const int *edge_cam, const int *edge_pv,
                           const int64_t *edge_off, const float *obj_xyz, const float *img_uv);
/* Replaces buildParas / paras2vertex (src/multicalib.cpp:422-459): n = 6*(nC+F-1) doubles. */
int mccba_set_parameters(mccba_handle h, int64_t n, const double *params);
int mccba_get_parameters(mccba_handle h, int64_t n, double *params);

/* Device-side snapshot / restore of the parameter vector (no host copy): lets a benchmark or an outlier loop rerun
 * the optimisation from the same starting point with everything resident in HBM. */
int mccba_save_parameters(mccba_handle h);
int mccba_restore_parameters(mccba_handle h);

/* ---- hot path ---------------------------------------------------------------------------------------------- */
/* One residual+Jacobian evaluation at the current parameters: the per-edge work of computeJacobianExtrinsic +
 * computePhotoCameraJacobian (src/multicalib.cpp:611-678, 717-824) with the per-edge normal-equation blocks
 * accumulated instead of a dense J.  Outputs (each optional, host buffers, reference edge order):
 *   cost      sum of squared residuals over this rank's edges
 *   edge_H6   n_edge x 21: upper triangle (row-major) of sum_i J_i^T J_i, J_i = d(u,v)/d(phi3, tau3) the 2x6
 *             Jacobian wrt a LEFT perturbation of the composed pose (R3 <- exp(phi3) R3, T3 <- T3 + tau3)
 *   edge_g6   n_edge x 6: sum_i J_i^T e_i,  e = observed - projected
 *   edge_cost n_edge: sum_i |e_i|^2
 * This is the "residual+Jacobian evals/sec" unit of the benchmark. */
int mccba_eval(mccba_handle h, double *cost, double *edge_H6, double *edge_g6, double *edge_cost);
/* Reduced camera system at the current parameters for damping lambda (Schur complement of the frame blocks, in
 * tangent coordinates): S is n_s x n_s row-major, gs n_s, n_s = 6*(nC-1).  Summed over ranks. Test/diagnostic. */
int mccba_reduced_system(mccba_handle h, double lambda, double *S, double *gs);
/* The whole optimisation loop of optimizeExtrinsics (src/multicalib.cpp:473-507) on the device: no host round
 * trip per iteration.  Parameters are updated in place (read back with mccba_get_parameters). */
int mccba_solve(mccba_handle h, const mccba_solve_opts *opts, mccba_report *report);
/* computeProjectError (src/multicalib.cpp:895-1006) at the current parameters, in fp64.
 * per_edge_mean (optional, n_edge) = edge.reprojecterror. */
int mccba_reproj_error(mccba_handle h, mccba_error_stats *stats, double *per_edge_mean);

/* ---- single-camera Mei calibration loop ------------------------------------------------------------------------
 * The optimisation part of cv::omnidir::calibrate (src/omnidir.cpp:1119-1147 with internal::computeJacobian :851-935,
 * flags2idx :2031-2076, fillFixed :2138-2153): per-frame poses + 10 intrinsics, parameter vector
 * [om_i, T_i] * n, fx, fy, s, cx, cy, xi, k1, k2, p1, p2 (encodeParameters, :1541-1568), step
 * G = (1 - 0.99^(iter+1)) (JTJ_sub + eps 11^T)^-1 JTE_sub with eps = 0.01 * 0.9^(iter/10), fixed parameters selected
 * by the omnidir::CALIB_FIX_* bits of `flags`.  The closed-form initialisation (initializeCalibration) is out of
 * scope: the caller supplies the starting parameters.
 * Several ranks (mccba_options.nranks > 1): the frames shard over the ranks -- every rank passes ITS frames and the
 * parameter vector [its poses | the 10 intrinsics]; the 78 sums of the reduced intrinsic system and the two norms of
 * the `change` criterion are summed with ncclAllReduce inside the iteration graph, every rank solves the 10-wide block
 * redundantly (bit-identical intrinsics), report->cost is the whole job's. */
int mccba_omni_set_observations(mccba_handle h, int n_frame, const int64_t *frame_off, const float *obj_xyz,
                                const float *img_uv);
int mccba_omni_set_parameters(mccba_handle h, int64_t n, const double *params);   /* n = 6 n_frame + 10 */
int mccba_omni_get_parameters(mccba_handle h, int64_t n, double *params);
/* report->cost = sum of squared residuals at the returned parameters; rms = sqrt(cost / n_points) (:1794-1802) */
int mccba_omni_solve(mccba_handle h, int flags, int crit_type, int max_count, double epsilon, mccba_report *report);
/* test hook: per-frame 17x17 Gram matrices sum [J(16) | e]^T [J(16) | e] at the current parameters (Jacobian columns
 * in the order of src/omnidir.cpp:65-73) and the total cost */
int mccba_omni_gram(mccba_handle h, double *gram, double *cost);

/* ---- utilities --------------------------------------------------------------------------------------------- */
/* sum a small host array of doubles over all ranks (NCCL); identity when nranks == 1 */
int mccba_allreduce_sum(mccba_handle h, double *buf, int n);
/* With the environment variable MCCBA_PROFILE=1 mccba_solve launches every kernel separately (no graph) with a
 * CUDA event between each; this returns the average ms per iteration of
 * out[0] frame_schur, out[1] reduce_records, out[2] allreduce, out[3] decide_solve, out[4] frame_update,
 * out[5] resid_jac_accum (diagnostic only -- the events serialise the stream). */
int mccba_last_kernel_ms(mccba_handle h, double out[6]);
/* How the packed reduced system [S | g | scalars] is summed over the ranks in every iteration (replaces nothing in the
 * reference, which is single-process): 0 = single rank, 1 = ncclAllReduce, 2 = low-latency exchange over NVLink peer
 * memory: every rank stores {data, epoch} words into every rank's CUDA-IPC mapped window and sums the slots in rank order
 * (p2p_exchange_kernel; bit-identical result on every rank).  Mode 2 is chosen collectively in mccba_set_observations
 * when all ranks can map all windows (<= 8 ranks on one NVLink domain); MCCBA_P2P=0 in the environment of every rank
 * forces mode 1. */
int mccba_exchange_mode(mccba_handle h);
/* Timing of the peer-memory exchange since the last call (device %globaltimer stamps of one thread per launch):
 * out[0] mean microseconds to post the stores into the peers' windows, out[1] mean microseconds from there until the
 * last peer's words have arrived (arrival skew of the ranks + one NVLink trip), out[2] launches counted, out[3] the
 * largest wait in microseconds.  Zeros in exchange modes 0 and 1.  Resets the counters. */
int mccba_exchange_stats(mccba_handle h, double out[4]);

/* Test hook: solve the SPD system S x = g (n x n row-major, lower triangle read) on the device with the loop's own
 * solvers.  blocked = 3: block cyclic reduction of a block-banded system (mccba_bcr.cuh) -- the bandwidth is measured
 * from S in 6 x 6 blocks and must be <= 4 blocks (half bandwidth <= 29); this is what the loop uses when the camera
 * graph is banded.  blocked = 2: one-launch tile-DAG Cholesky (the loop's choice for dense camera graphs).  Other values
 * are rejected with MCCBA_ERR_ARG.  Replaces the Eigen conjugate-gradient solve of src/multicalib.cpp:565-592 on the
 * Schur-reduced system.
 * The kernel time in ms is left in mccba_last_kernel_ms()[0]. */
int mccba_debug_solve_dense(mccba_handle h, int n, const double *S, const double *g, double *x, int blocked);
/* time `reps` back-to-back launches of the residual+Jacobian kernel at the current parameters with CUDA events on
 * the handle's stream; returns average ms per launch (benchmark helper, no other side effects) */
int mccba_time_eval(mccba_handle h, int reps, double *avg_ms);

/* ---- omnidir stereo bundle adjustment --------------------------------------------------
 * The optimisation loop of cv::omnidir::stereoCalibrate (src/omnidir.cpp:1268-1296 with computeJacobianStereo :937-1020,
 * flags2idxStereo :2078-2136) and estimateUncertaintiesStereo (:1804-1889), without the closed-form initialisation.
 * Parameter vector (encodeParametersStereo :1570-1620): [om, T of camera 2 relative to camera 1 | om_i, T_i of the n
 * frames in camera 1 | fx fy s cx cy xi k1 k2 p1 p2 of camera 1 | the same of camera 2] = 6 (n + 1) + 20 doubles.
 * Every frame is seen by both cameras with the same object points (frame_off[n_frame + 1] corner offsets). */
int mccba_stereo_set_observations(mccba_handle h, int n_frame, const int64_t *frame_off, const float *obj_xyz,
                                  const float *img1_uv, const float *img2_uv);
int mccba_stereo_set_parameters(mccba_handle h, int64_t n, const double *params);
int mccba_stereo_get_parameters(mccba_handle h, int64_t n, double *params);
/* flags: cv::omnidir::CALIB_FIX_* bits (applied to both cameras); criteria decoded as at :1271-1274 */
int mccba_stereo_solve(mccba_handle h, int flags, int crit_type, int max_count, double epsilon, mccba_report *rep);
/* errors[6 (n + 1) + 20] = 3 s sqrt(diag((J^T J)^-1)) (0 at fixed parameters), std_error = (sigma_x, sigma_y), rms */
int mccba_stereo_uncertainties(mccba_handle h, int flags, double *errors, double std_error[2], double *rms);

/* ---- double-sided board calibration --------------------------------------------------------
 * The optimisation of cv::multicalib::DoubleSideCalibration (src/doubleSide.cpp): the cameras are FIXED at known poses;
 * unknown are the front<->back transform D of the board and one pose per frame.  An edge that sees the back pattern is
 * composed with D innermost, X_cam = R_c (R_p (R_D X + t_D) + t_p) + t_c (computePhotoCameraJacobian, :288-429;
 * the same chain as the back-pattern branch of src/mymulticalib.cpp:468-614); parameter vector
 * [D | frame 0 | frame 1 | ...], 6 values [rvec | tvec] each (buildParas, :233-261); step schedule and termination as
 * optimizeExtrinsics (src/multicalib.cpp:462-514).  The problem sits on top of mccba_set_cameras +
 * mccba_set_observations (edges, corners, intrinsics; the camera parameters of the rig path are not used).
 *   edge_back[e] (reference edge order): 1 if edge e sees the back pattern (edge::patternSide == BACK_PATTERN)
 *   cam_pose: n_cam x 6 [rvec | tvec], world -> camera (DoubleSideCalibration::camerasPose, loadCameraPose :276-287)
 * MCCBA_ERR_ARG if no edge sees the back pattern (D would not be observable).  Single rank. */
int mccba_ds_set_problem(mccba_handle h, const unsigned char *edge_back, const double *cam_pose);
int mccba_ds_set_parameters(mccba_handle h, int64_t n, const double *params);   /* n = 6 + 6 n_frame */
int mccba_ds_get_parameters(mccba_handle h, int64_t n, double *params);
/* report->cost = sum of squared residuals at the returned parameters */
int mccba_ds_solve(mccba_handle h, int crit_type, int max_count, double epsilon, mccba_report *report);
/* test hook: the 6 x 6 system of D after elimination of the frame poses (tangent coordinates) and the cost, at the
 * current parameters */
int mccba_ds_normal(mccba_handle h, double *S36, double *g6, double *cost);

#ifdef __cplusplus
}
#endif
#endif /* MCCBA_H_ */
