// Device-side problem view shared by the kernels (kb_kernels.cu) and the host layer (kb_host.cpp).
// Data layout in HBM: structure-of-arrays observations in the reference's term order,
// views (one camera image of one synced set) as contiguous term ranges, per-view 16x16 Gram blocks, per-set
// Schur blocks, one dense reduced camera system.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/kalibr_b200/lm_state_machine.h"

namespace kb {

enum : int { PINHOLE_RADTAN = 0, PINHOLE_EQUI = 1, OMNI_RADTAN = 2, EUCM_NONE = 3, DS_NONE = 4, PINHOLE_FOV = 5, OMNI_NONE = 6, NUM_MODELS = 7 };  // = kb_camera_model
__host__ __device__ constexpr int model_P(int m) { return (m == OMNI_RADTAN || m == OMNI_NONE) ? 5 : (m == EUCM_NONE || m == DS_NONE) ? 6 : 4; }
__host__ __device__ constexpr int model_D(int m) { return (m == EUCM_NONE || m == DS_NONE || m == OMNI_NONE) ? 0 : m == PINHOLE_FOV ? 1 : 4; }

constexpr int MAX_CAMS = 32;
constexpr int CAM_PARAM_STRIDE = 10;
constexpr int POSE_STRIDE = 7;
constexpr int GRAM_DIM = 16;                       // local columns: xi(6) | proj(P) | dist(D) | pad | e (col 15)
constexpr int GRAM_SIZE = GRAM_DIM * GRAM_DIM;     // row-major; only the tiles (0,0), (0,1), (1,1) are written
constexpr int E_COL = 15;
constexpr int GRAM_TILES = 192;                    // the three stored 8x8 tiles (0,0), (0,1), (1,1) in mma C-fragment order
constexpr int SETPREP_STRIDE = 48;                 // per set: C^-1 (9), -C^-1 t (3), P_v (36)
// per-view block written by the fused kernel: Gram tiles (0,0) and (0,1), each row-major 8x8
//   tile (0,0): G[0:8][0:8]  (pose x pose, pose x first two intrinsics)      tile (0,1): G[0:8][8:16]  (pose x remaining intrinsics, pose x e)
constexpr int VB_STRIDE = 128;

// Control block of the device-resident Levenberg-Marquardt loop (kb_optimize): the optimiser / trust-region state machine of
// include/kalibr_b200/lm_state_machine.h (shared with the host mirror) and the flags that make an enqueued kernel a no-op when
// the iteration it belongs to does not need it.  In the call-by-call API (kb_evaluate_error, kb_build_system, ...) the flags stay
// neutral (done = 0, need_build = 1, skip_eval = 0, revert = 0) and the scalar parameters travel as kernel arguments.
using LmCtrl = kalibr_b200::LmState;

// Peer exchange over NVLink (multi-rank, one process per GPU, kb_attach_peers): every rank maps the exchange buffer of every
// other rank (CUDA IPC) and the producing kernel of each exchange step stores its contribution straight into the slot it owns
// in EVERY rank's buffer, then raises a per-source epoch flag there; the consuming kernel waits for the flags and sums the
// slots in rank order.  No collective library call sits between producer and consumer, and the sum order is fixed.
//   exchange A: this rank's reduced camera system (schur_finalize_kernel -> px_reduce_system_kernel)
//   exchange B: (rho denominator, max|dx|, pos-def) of a solve (rho_stage2_kernel -> px_combine_solve_kernel)
//   exchange C: cost of an evaluation (gram_cost_kernel -> px_combine_cost_kernel)
// Slots are double-buffered by epoch parity: a rank can run at most one exchange ahead of the slowest one.
constexpr int PX_MAX_RANKS = 8;
struct PeerXchg {
  int enabled, n_ranks, rank, na2;  // na2 = n_aug^2 rounded up to a multiple of 2
  double* base[PX_MAX_RANKS];       // base[r]: exchange buffer of rank r as mapped into this process (base[rank] = own)
};
// layout of an exchange buffer, in doubles
__host__ __device__ inline size_t px_off_a(const PeerXchg& x, int parity, int src) { return ((size_t)parity * x.n_ranks + src) * x.na2; }
__host__ __device__ inline size_t px_off_b(const PeerXchg& x, int parity, int src) { return (size_t)2 * x.n_ranks * x.na2 + ((size_t)parity * x.n_ranks + src) * 4; }
__host__ __device__ inline size_t px_off_c(const PeerXchg& x, int parity, int src) { return (size_t)2 * x.n_ranks * x.na2 + 8 * x.n_ranks + ((size_t)parity * x.n_ranks + src) * 2; }
// 64-bit words after the slots: flags [3][n_ranks] (written by the peers), then local: epoch[3], block counter, error flag
__host__ __device__ inline size_t px_off_flags(const PeerXchg& x) { return (size_t)2 * x.n_ranks * x.na2 + 12 * x.n_ranks; }
__host__ __device__ inline size_t px_doubles(const PeerXchg& x) { return px_off_flags(x) + 3 * x.n_ranks + 8; }

struct DevProblem {
  // ---- structure (immutable after kb_create) ----
  int n_cams;
  int n_sets;        // local synced sets
  int n_views;       // local views
  int n_target;
  long long n_terms; // local terms
  int n_c;           // reduced camera system dimension (global)
  int n_aug;         // n_c + 1 (rhs folded in as last row/col)
  int cam_model[MAX_CAMS];
  int cam_P[MAX_CAMS];
  int cam_D[MAX_CAMS];
  int intr_off[MAX_CAMS];  // offset of [proj|dist] of camera k in the reduced system
  int base_off[MAX_CAMS];  // offset of baseline j (q then t) in the reduced system
  const double* y_u;
  const double* y_v;
  const uint16_t* corner;
  const double* target;    // [n_target][3]
  const int* view_set;     // [n_views] local set index
  const int* view_cam;     // [n_views]
  const int* view_begin;   // [n_views+1] local term offsets
  const int* set_view;     // [n_sets][n_cams] view index or -1
  const int* lin_off;      // [n_cams][LIN_OFF_STRIDE] column offsets of the CCS J^T layout per camera
  const long long* view_jbase;  // [n_views] offset of the view's first value in the CCS J^T value array
  const int* col_desc;     // [n_cams][col_desc_stride] per CCS J^T column of camera k: kind << 16 | j << 8 | sub
                           //   kind 0 = set pose (sub 0..5), 1 = baseline j (sub 0..5), 2 = intrinsics (sub = local column - 6)
  int col_desc_stride;
  // ---- state ----
  double* cam_params;   // [n_cams][10]
  double* baselines;    // [n_cams-1][7]
  double* set_poses;    // [n_sets][7]
  // ---- per-linearisation constants (prep kernel) ----
  double* camT;         // [n_cams][12]  R (row-major 9), t (3): T_cam(k)_cam(0)
  double* camPi;        // [n_cams][36]  product of boxTimes(B_{k-1}) ... boxTimes(B_0)
  double* camA;         // [n_cams][n_cams][36]  A_{j,k} at [k][j]: d(xi_k)/d(baseline j) = X_{k,j} [M_q(t_j) | M_t]
  double* baseBt;       // [n_cams-1][36]  boxTimes(B_j)
  double* baseM;        // [n_cams-1][36]  [M_q(t_j) | M_t]
  // ---- outputs ----
  double* e;            // [2*n_terms]   -(y - y_hat), the reference's _e
  double* view_cost;    // [n_views]
  double* set_prep;     // [n_sets][48]
  double* VB;           // [n_views][128] view blocks (two Gram tiles)
  double* gram_partial; // [n_slices][192]
  double* sumG;         // [n_cams][256] full symmetric
  double* V;            // [n_sets][36]
  double* bv;           // [n_sets][6]
  double* W;            // [n_sets][n_c][6]
  double* Lv;           // [n_sets][36]   inverse of the Cholesky factor of V + damping (row-major lower)
  double* yv;           // [n_sets][6]    L^-1 b_v
  double* U;            // [n_aug*n_aug]  camera block with rhs b_c in the last row/col (this rank's partial)
  double* Sred;         // [n_aug*n_aug]  reduced system (after all-reduce), then its Cholesky factor
  double* dxc;          // [n_c]
  double* dx;           // [jcols] in design-variable order (poses of other ranks stay 0)
  unsigned int* n_invalid;  // terms whose projection bailed out (Q6)
  double* rho_partial;      // [2 * max(64, back-substitution blocks)] stage-1 partials of the rho denominator / max|dx| reduction (per handle)
  unsigned int* tickets;    // [4] zero-initialised "last block" ticket counters ([0] finalize_gram, [1] backsub)
  LmCtrl* ctrl;             // control block (device)
  PeerXchg px;              // peer exchange (enabled after kb_attach_peers)
  // ---- weighting of the terms (kb_set_inv_r / kb_set_m_estimator) ----
  // weighted == 0 (default): invR = I and NoMEstimator, the kernels run their unweighted instantiations.  Otherwise every term's
  // rows and residual are multiplied by sqrt(w) sqrtInvR^T (BE/include/aslam/backend/implementation/ErrorTerm.hpp:97-109, 170-192)
  // with w = policy(e^T invR e) (BE/src/MEstimatorPolicies.cpp), and the cost is sum w e^T invR e (BE/src/ErrorTerm.cpp:19-24).
  int weighted;
  int mest_kind;       // kb_m_estimator
  int mest_rows;       // 1: w also scales e() and the Jacobian rows (useMEstimator = true); 0: only the cost
  double mest_param;   // Huber: k; Cauchy, Geman-McClure: sigma^2; Blake-Zisserman: epsilon
  double sT[4];        // sqrtInvR^T, row-major
};

constexpr int LIN_OFF_STRIDE = 4 + MAX_CAMS;  // pose_q, pose_t, proj, dist, baseline j (q; t = +3)

struct StreamCtx {
  cudaStream_t stream;
  long long* launches;  // incremented per kernel launch
  // optional: side streams + events so that the per-model launches of a mixed rig run concurrently (fork / join on `stream`)
  cudaStream_t* side = nullptr;    // [NUM_MODELS]
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t* ev_join = nullptr;  // [NUM_MODELS]
};

// launchers (kb_kernels.cu); every one returns the cudaGetLastError() of its launches
cudaError_t launch_prep(const DevProblem& p, StreamCtx& s);
cudaError_t launch_evaluate(const DevProblem& p, const int* view_list, const int* model_begin, double* cost_out, StreamCtx& s);
// reprojection statistics per camera from raw residuals in e_raw (= -(y - y_hat) per term, as evaluate writes them): pass 0 sums
// (n, sum e_u, sum e_v) into acc[cam][4], pass 1 (after the sums are global) the squared deviations from the mean into acc[cam][4 + {0,1}]
cudaError_t launch_reproj_stats(const DevProblem& p, const double* e_raw, const int* cam_view_list, const int* cam_view_begin, int pass,
                                double* acc /*[n_cams][8]*/, StreamCtx& s);
int la_grid_warps();
// with_set_prep / with_cam_prep: the per-set and per-camera constants are (re)computed by one launch in front of the fused kernel
cudaError_t launch_linearise_assemble(const DevProblem& p, const int4* vmeta /*(view,set,begin,end) in view-list order*/, const int4* slices,
                                      const int* slice_model_begin, bool write_e, bool with_set_prep, bool with_cam_prep, StreamCtx& s);
// lm_mode 1 (single-rank device-resident loop): the kernel also runs the loop boundary (accept / reject, next iteration's decisions)
cudaError_t launch_finalize_gram(const DevProblem& p, const int* cam_slice_range /*[n_cams][n_ranges][2]*/, int n_ranges, double* cost_out,
                                 bool exchange_cost /* also the producer of peer exchange C */, int lm_mode, double* trace, int* pos_def, StreamCtx& s);
cudaError_t launch_set_reduce(const DevProblem& p, StreamCtx& s);
cudaError_t launch_linearise_materialise(const DevProblem& p, const int4* vmeta, const int4* slices, const int* slice_model_begin,
                                         const int* bfrag_pairs /*[NUM_MODELS]*/, unsigned int* counters /*[NUM_MODELS], device*/, double* jt_values,
                                         StreamCtx& s);
// damping argument: KB_DAMPING_FROM_CTRL (NaN) means "read it from the control block" (device-resident loop).  A real damping may be
// NEGATIVE: the BlockCholesky semantic un-augments with lambda instead of lambda^2 (Q2), so after a rejected step the diagonal carries
// lambda_old^2 - lambda_old + lambda_new^2 < 0 and the solve must see exactly that (the reference then reports "not positive definite").
// lambda arguments: a negative value means the control block's (lambda itself is never negative)
#define KB_DAMPING_FROM_CTRL (__builtin_nan(""))
cudaError_t launch_schur(const DevProblem& p, double damping, double* partials, int n_partials, int* pos_def_flag, StreamCtx& s);
cudaError_t launch_schur_finalize(const DevProblem& p, double damping, const double* partials, int n_partials, bool add_camera_block, StreamCtx& s);
// from_peers: the kernel waits for exchange A and sums the ranks' slots itself (no launch_px_reduce_system in front)
cudaError_t launch_reduced_solve(const DevProblem& p, double damping, int* pos_def_flag, bool from_peers, StreamCtx& s);
int backsub_blocks(const DevProblem& p);
cudaError_t launch_backsub(const DevProblem& p, const int* set_col_q, const int* set_col_t, const int* cam_cols, double lambda, int include_shared,
                           double* out2, const int* pos_def_for_exchange, const int* pos_def, int lm_mode, StreamCtx& s);
cudaError_t launch_rho_denominator(const DevProblem& p, double lambda, const int* set_col_q, const int* set_col_t, const int* cam_cols,
                                   int include_shared, double* out2 /* [0]=sum, [1]=max|dx| */,
                                   const int* pos_def_for_exchange /* non-null: also the producer of peer exchange B */, StreamCtx& s);
cudaError_t launch_pack_rank_scalars(double* pk /*[n_ranks][4], device*/, int rank, int n_ranks, const double* rho_max, const int* pos_def, StreamCtx& s);
cudaError_t launch_apply_update(const DevProblem& p, const int* set_col_q, const int* set_col_t, const int* cam_cols, double* backup_cam,
                                double* backup_base, double* backup_sets, StreamCtx& s);
// device-resident LM loop: single-thread control kernels and the conditional revert
cudaError_t launch_lm_post_solve(const DevProblem& p, const int* pos_def_flag, const double* rho_max, const double* rank_slots, int n_ranks, StreamCtx& s);
cudaError_t launch_lm_boundary(const DevProblem& p, double* trace, int* pos_def, StreamCtx& s);
cudaError_t launch_lm_revert(const DevProblem& p, const double* backup_cam, const double* backup_base, const double* backup_sets, StreamCtx& s);
cudaError_t launch_lm_finish(const DevProblem& p, StreamCtx& s);
// eigen-decomposition of the reduced system in p.Sred (G, V: scratch of (n_c + 1) * n_c doubles each, V_tmp: n_c * n_c):
// Householder + QL, then a Jacobi polish from the QL vectors.  status: int[2] = {Jacobi sweeps, QL failed}
// V_warm (may be null): vectors of a nearby system to start the Jacobi iteration from instead of running QL; V_keep (may be null): gets V_out
cudaError_t launch_marginal_eig(const DevProblem& p, double* G, double* V, double* sv_out /*[n_c]*/, double* V_out /*[n_c][n_c]*/, double* V_tmp,
                                int* status, const double* V_warm, double* V_keep, StreamCtx& s);
// truncated-SVD solve of the reduced system in p.Sred (undamped Schur complement): optional column scaling from diag_h (global diagonal
// of the camera block), eigen-decomposition (QL; plus the Jacobi polish when the system is not scaled), rank cut, x_r into p.dxc;
// result = {rank, tolerance, gap}
cudaError_t launch_camera_diag(const DevProblem& p, double* out /*[n_c]*/, StreamCtx& s);
cudaError_t launch_svd_solve(const DevProblem& p, const double* diag_h, double norm_tol, int column_scaling, double eps_svd, double svd_tol,
                             double* g /*[n_c]*/, double* G, double* V, double* sv, double* V_out, double* V_tmp, int* status, double* result /*[4]*/,
                             const double* V_warm, double* V_keep, StreamCtx& s);
// peer exchange consumers
cudaError_t launch_px_reduce_system(const DevProblem& p, StreamCtx& s);
cudaError_t launch_px_combine_solve(const DevProblem& p, double* rho_max /*[2]*/, int* pos_def_flag, int lm_mode, StreamCtx& s);
cudaError_t launch_px_combine_cost(const DevProblem& p, double* cost, int lm_mode, double* trace, int* pos_def, StreamCtx& s);
// initial-guess stage (kb_init.cu): PnP per view (view_mask: null = every view of the list), best view per set, target pose guesses
cudaError_t launch_estimate_transformations(const DevProblem& p, const int* view_list, const int* model_begin, const unsigned char* view_mask,
                                            const int* resolution /*[n_cams][2] device, or null*/, double* T_out /*[n_views][7]*/, int* ok_out /*[n_views]*/,
                                            StreamCtx& s);
// initializeIntrinsics: pinhole family -> out2 = (median focal guess, number of guesses); omni family -> out3 = (gamma, success, error)
cudaError_t launch_focal_guesses(const DevProblem& p, const int* cam_views, int n_views, int rows, int cols, double* fg /*[n_views][pairs]*/, double* out2,
                                 StreamCtx& s);
cudaError_t launch_omni_candidates(const DevProblem& p, const int* cam_views, int n_views, int rows, int cols, int ru, int rv,
                                   double* cand /*[n_views * rows][2]*/, double* out3, StreamCtx& s);
cudaError_t launch_best_view_mask(const DevProblem& p, unsigned char* mask /*[n_views]*/, StreamCtx& s);
cudaError_t launch_set_pose_guess(const DevProblem& p, const double* T_views, const int* ok_views, double* set_poses_out, int* set_ok, StreamCtx& s);
// su / sv (device, n floats each; 8-byte aligned offsets) -> du / dv (doubles); offsets must be even for the vectorised path
cudaError_t launch_widen_observations(const float* su, const float* sv, double* du, double* dv, long long n, cudaStream_t stream, long long* launches);
int schur_num_partials(const DevProblem& p);
size_t schur_partial_stride(const DevProblem& p);

}  // namespace kb
