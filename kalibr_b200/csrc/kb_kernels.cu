// Hand-written sm_100a kernels of the batch-calibration hot path.
//
//   prep / set_prep   per-linearisation constants: camera chain (T_cam_k_cam_0, boxTimes products, baseline adjoints), per-set
//                     inverse pose and P_v
//   evaluate          K5: residual-only pass, cost (non-speculative mode)   ≙ BE/src/LinearSystemSolver.cpp:12-23, 81-92
//   linearise_assemble  K1+K2 fused: per term chain + projection + 3 Jacobians, per view Gram block
//                     G = [J_xi|J_proj|J_dist|e]^T [..] on the FP64 tensor pipe (DMMA m8n8k4)
//                                                                       ≙ ReprojectionError.hpp:63-77 + JacobianContainer.cpp:103-167
//   linearise_materialise  K1 with J written out in the reference's CCS J^T layout (J = A B on the tensor pipe)
//                                                                       ≙ CompressedColumnJacobianTransposeBuilder.hpp:59-100
//   set_reduce        per set: V_v, b_v, W_v from the views' Gram tiles (register-resident DMMA chain)
//                                                                       ≙ sparse_block_matrix.hpp:121-143 (block += J1^T J2)
//   finalize_gram / camera_block / gram_cost  U, b_c, cost
//   pose_factor / schur / schur_finalize  K3a: S = U - sum_v W_v (V_v + d I)^-1 W_v^T on DMMA  ≙ BE/src/sparse_matrix_functions.cpp:8-60
//   reduced_solve     K3b: dense Cholesky of the reduced system            ≙ linear_solver_cholmod.h:70-112
//   backsub           K3c: dx_v = (V_v + d I)^-1 (b_v - W_v^T dx_c)         ≙ sparse_matrix_functions.cpp:64-83
//   rho / apply_update                                                   ≙ LevenbergMarquardtTrustRegionPolicy.cpp:107-113, Optimizer2.cpp:290-318
//   lm_boundary (inside finalize_gram / px_combine_cost) / lm_after_solve (inside backsub's last block / px_combine_solve) / lm_revert / lm_finish
//                     device-resident LM loop: the transitions of include/kalibr_b200/lm_state_machine.h
//                                                                       ≙ Optimizer2.cpp:215-266, LevenbergMarquardtTrustRegionPolicy.cpp:50-113
//   px_*              NVLink peer exchange between ranks (producers are fused into schur_finalize / backsub / finalize_gram)
#include <algorithm>
#include <cstring>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>

#include "kb_device.cuh"
#include "kb_models.cuh"

namespace kb {

// =========================================================================================================
// small fixed-size algebra (registers / local arrays)
// =========================================================================================================
// sm_kinematics quat2r, scalar-last (Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:77-101); R row-major
__device__ __forceinline__ void quat2r(const double* __restrict__ q, double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = x * x - y * y - z * z + w * w;
  R[1] = x * y * 2.0 + z * w * 2.0;
  R[2] = x * z * 2.0 - y * w * 2.0;
  R[3] = x * y * 2.0 - z * w * 2.0;
  R[4] = -x * x + y * y - z * z + w * w;
  R[5] = x * w * 2.0 + y * z * 2.0;
  R[6] = x * z * 2.0 + y * w * 2.0;
  R[7] = x * w * (-2.0) + y * z * 2.0;
  R[8] = -x * x - y * y + z * z + w * w;
}

// general 3x3 inverse (the reference inverts the whole 4x4 with Eigen's general inverse:
// BX/src/TransformationExpressionNode.cpp:81,88; for [C t; 0 1] that is [C^-1, -C^-1 t])
__device__ __forceinline__ void inv3(const double A[9], double B[9]) {
  const double c00 = A[4] * A[8] - A[5] * A[7];
  const double c01 = A[5] * A[6] - A[3] * A[8];
  const double c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  B[0] = c00 * id;
  B[1] = (A[2] * A[7] - A[1] * A[8]) * id;
  B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = c01 * id;
  B[4] = (A[0] * A[8] - A[2] * A[6]) * id;
  B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = c02 * id;
  B[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// boxTimes(T) = [C, -t^x C; 0, C]  (Schweizer-Messer/sm_kinematics/src/transformations.cpp:132-142); 6x6 row-major
__device__ __forceinline__ void box_times(const double C[9], const double t[3], double o[36]) {
#pragma unroll
  for (int i = 0; i < 36; ++i) o[i] = 0.0;
  // -t^x = [0, t2, -t1; -t2, 0, t0; t1, -t0, 0]
  const double m[9] = {0.0, t[2], -t[1], -t[2], 0.0, t[0], t[1], -t[0], 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      o[i * 6 + j] = C[i * 3 + j];
      o[(i + 3) * 6 + j + 3] = C[i * 3 + j];
      o[i * 6 + j + 3] = m[i * 3 + 0] * C[0 * 3 + j] + m[i * 3 + 1] * C[1 * 3 + j] + m[i * 3 + 2] * C[2 * 3 + j];
    }
}

// [M_q(t) | M_t] = [[-t^x, I], [I, 0]]  (BX/src/TransformationBasic.cpp:48-67); 6x6 row-major, cols 0..2 = q, 3..5 = t
__device__ __forceinline__ void pose_jac(const double t[3], double o[36]) {
#pragma unroll
  for (int i = 0; i < 36; ++i) o[i] = 0.0;
  o[0 * 6 + 1] = t[2];  o[0 * 6 + 2] = -t[1];
  o[1 * 6 + 0] = -t[2]; o[1 * 6 + 2] = t[0];
  o[2 * 6 + 0] = t[1];  o[2 * 6 + 1] = -t[0];
  o[0 * 6 + 3] = 1.0; o[1 * 6 + 4] = 1.0; o[2 * 6 + 5] = 1.0;
  o[3 * 6 + 0] = 1.0; o[4 * 6 + 1] = 1.0; o[5 * 6 + 2] = 1.0;
}

__device__ __forceinline__ void mul6(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ Cc) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += A[i * 6 + k] * B[k * 6 + j];
      Cc[i * 6 + j] = s;
    }
}

// T_cam_w = T_cam_k_cam_0 * inverse(T_v): R_cw (row-major), t_cw.  BX/src/TransformationExpressionNode.cpp:54-59, 86-90
__device__ __forceinline__ void view_transform(const double* __restrict__ pose7, const double* __restrict__ camT, double Rcw[9], double tcw[3]) {
  double C[9], Ci[9];
  quat2r(pose7, C);
  inv3(C, Ci);
  const double tx = pose7[4], ty = pose7[5], tz = pose7[6];
  const double ti[3] = {-(Ci[0] * tx + Ci[1] * ty + Ci[2] * tz), -(Ci[3] * tx + Ci[4] * ty + Ci[5] * tz), -(Ci[6] * tx + Ci[7] * ty + Ci[8] * tz)};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Rcw[i * 3 + j] = camT[i * 3 + 0] * Ci[0 * 3 + j] + camT[i * 3 + 1] * Ci[1 * 3 + j] + camT[i * 3 + 2] * Ci[2 * 3 + j];
    tcw[i] = camT[i * 3 + 0] * ti[0] + camT[i * 3 + 1] * ti[1] + camT[i * 3 + 2] * ti[2] + camT[9 + i];
  }
}

// P_v = -boxTimes(inverse(T_v)) [M_q(t_v) | M_t]: d(xi at the Inverse node)/d(q_v, t_v)
// BX/src/TransformationExpressionNode.cpp:98-101 chained into TransformationBasic.cpp:48-67
__device__ __forceinline__ void set_pose_jac(const double* __restrict__ pose7, double Pv[36]) {
  double C[9], Ci[9];
  quat2r(pose7, C);
  inv3(C, Ci);
  const double t[3] = {pose7[4], pose7[5], pose7[6]};
  const double ti[3] = {-(Ci[0] * t[0] + Ci[1] * t[1] + Ci[2] * t[2]), -(Ci[3] * t[0] + Ci[4] * t[1] + Ci[5] * t[2]), -(Ci[6] * t[0] + Ci[7] * t[1] + Ci[8] * t[2])};
  double bt[36], M[36];
  box_times(Ci, ti, bt);
  pose_jac(t, M);
  mul6(bt, M, Pv);
#pragma unroll
  for (int i = 0; i < 36; ++i) Pv[i] = -Pv[i];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 6x6 Cholesky (row-major lower) ; returns false if not positive definite
__device__ __forceinline__ bool chol6(double A[36]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * 6 + k] * A[j * 6 + k];
    if (!(d > 0.0)) ok = false;
    d = sqrt(d);
    A[j * 6 + j] = d;
    const double id = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = s * id;
    }
  }
  return ok;
}
__device__ __forceinline__ void fwd6(const double* __restrict__ L, double x[6]) {  // L z = x
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * x[k];
    x[i] = s / L[i * 6 + i];
  }
}
__device__ __forceinline__ void bwd6(const double* __restrict__ L, double x[6]) {  // L^T z = x
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
}

// FP64 tensor-core MMA, D(8x8) += A(8x4) B(4x8).  Fragment ownership (PTX ISA, mma.m8n8k4 .f64):
//   a: row = lane/4, col = lane%4 ; b: row = lane%4, col = lane/4 ; c0,c1: row = lane/4, col = 2*(lane%4) + {0,1}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- programmatic dependent launch -------------------------------------------------------------------------------------------
// Every kernel of an LM iteration starts with pdl_enter(): wait until the preceding kernel of the stream has completed and its
// writes are visible (griddepcontrol.wait; a no-op for a kernel launched without the attribute), then let the kernel that follows
// start launching.  With launch_pdl() on the host side the CTAs of kernel N + 1 are scheduled while kernel N drains instead of
// after it; results are unchanged because nothing is read before the wait.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_launch_dependents();
}
static bool pdl_enabled() {
  static const bool on = getenv("KB_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- peer exchange primitives (kb_device.cuh: PeerXchg) -------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long* px_words(const PeerXchg& x, int r) {
  return reinterpret_cast<unsigned long long*>(x.base[r] + px_off_flags(x));
}
// words: [which * n_ranks + src] flags | 3 n_ranks + which: local epoch | 3 n_ranks + 3: block counter | 3 n_ranks + 4: error
__device__ __forceinline__ unsigned long long px_next_epoch(const PeerXchg& x, int which) { return px_words(x, x.rank)[3 * x.n_ranks + which] + 1ull; }
__device__ __forceinline__ unsigned long long px_cur_epoch(const PeerXchg& x, int which) { return px_words(x, x.rank)[3 * x.n_ranks + which]; }
// after this rank's stores of exchange `which`, epoch e, are fenced: publish
__device__ __forceinline__ void px_signal(const PeerXchg& x, int which, unsigned long long e) {
  px_words(x, x.rank)[3 * x.n_ranks + which] = e;
  __threadfence_system();
  for (int r = 0; r < x.n_ranks; ++r) st_release_sys(px_words(x, r) + which * x.n_ranks + x.rank, e);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
constexpr unsigned long long PX_TIMEOUT_NS = 5000000000ull;  // wall clock (globaltimer), independent of the SM clock
// wait until rank src has published epoch e of exchange `which`; gives up after 5 s (a peer that died must not hang the GPU):
// the error word of the exchange buffer is raised and the device-resident loop is told to stop (ctrl->done), so that nothing
// downstream iterates on the stale slots; the host turns the error word into KB_ERR_NCCL.
__device__ __forceinline__ void px_wait(const PeerXchg& x, int which, int src, unsigned long long e, LmCtrl* ctrl) {
  const unsigned long long* f = px_words(x, x.rank) + which * x.n_ranks + src;
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(f) < e) {
    if (global_timer_ns() - t0 > PX_TIMEOUT_NS) {
      px_words(x, x.rank)[3 * x.n_ranks + 4] = 1ull;
      ctrl->done = 1;
      break;
    }
    __nanosleep(64);
  }
}

// =========================================================================================================
// prep: per camera k, the constants shared by all of its views.
// =========================================================================================================
__device__ __forceinline__ void camera_prep(const DevProblem& p, int k) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
  for (int l = 0; l < k; ++l) {  // T_k = B_{k-1} ... B_0   (CalibrationTools.hpp:405-408)
    const double* b = p.baselines + l * POSE_STRIDE;
    double Rl[9], Rn[9], tn[3];
    quat2r(b, Rl);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = Rl[i * 3 + 0] * R[0 * 3 + j] + Rl[i * 3 + 1] * R[1 * 3 + j] + Rl[i * 3 + 2] * R[2 * 3 + j];
      tn[i] = Rl[i * 3 + 0] * t[0] + Rl[i * 3 + 1] * t[1] + Rl[i * 3 + 2] * t[2] + b[4 + i];
    }
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    for (int i = 0; i < 3; ++i) t[i] = tn[i];
  }
  double* o = p.camT + k * 12;
  for (int i = 0; i < 9; ++i) o[i] = R[i];
  for (int i = 0; i < 3; ++i) o[9 + i] = t[i];
  // chain rule down the Multiply nodes (BX/src/TransformationExpressionNode.cpp:68-72): outermost baseline first
  double X[36];
  for (int i = 0; i < 36; ++i) X[i] = (i % 7 == 0) ? 1.0 : 0.0;
  for (int l = k - 1; l >= 0; --l) {
    const double* b = p.baselines + l * POSE_STRIDE;
    double Rl[9], M[36], A[36], bt[36], Xn[36];
    quat2r(b, Rl);
    pose_jac(b + 4, M);
    mul6(X, M, A);
    double* Ao = p.camA + ((size_t)k * p.n_cams + l) * 36;
    for (int i = 0; i < 36; ++i) Ao[i] = A[i];
    box_times(Rl, b + 4, bt);
    mul6(X, bt, Xn);
    for (int i = 0; i < 36; ++i) X[i] = Xn[i];
  }
  double* Po = p.camPi + k * 36;
  for (int i = 0; i < 36; ++i) Po[i] = X[i];
  if (k + 1 < p.n_cams) {  // per baseline k: boxTimes(B_k) and [M_q(t_k) | M_t] for the set reduction's recurrence
    const double* b = p.baselines + k * POSE_STRIDE;
    double Rl[9], M[36], bt[36];
    quat2r(b, Rl);
    pose_jac(b + 4, M);
    box_times(Rl, b + 4, bt);
    for (int i = 0; i < 36; ++i) {
      p.baseBt[k * 36 + i] = bt[i];
      p.baseM[k * 36 + i] = M[i];
    }
  }
}
__global__ void prep_kernel(DevProblem p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_cams || p.ctrl->done || p.ctrl->skip_eval) return;
  camera_prep(p, k);
}

// per synced set: inverse pose (C^-1, -C^-1 t) and P_v, shared by the views of every camera of the set
// with_cam_prep: one extra block at the end of the grid computes the per-camera constants (prep_kernel's work) in the same launch
constexpr int SET_PREP_THREADS = 64;
__global__ void __launch_bounds__(SET_PREP_THREADS) set_prep_kernel(DevProblem p, int with_cam_prep) {
  pdl_enter();
  // thread per set; the 48 values of a set go through shared memory so that the block writes its rows as one coalesced run
  __shared__ double stage[SET_PREP_THREADS][SETPREP_STRIDE + 1];
  if (p.ctrl->done || p.ctrl->skip_eval) return;
  if (with_cam_prep && blockIdx.x == gridDim.x - 1) {
    if (threadIdx.x < p.n_cams) camera_prep(p, threadIdx.x);
    return;
  }
  const int s0 = blockIdx.x * SET_PREP_THREADS;
  const int s = s0 + threadIdx.x;
  if (s < p.n_sets) {
    const double* pose = p.set_poses + (size_t)s * POSE_STRIDE;
    double C[9], Ci[9];
    quat2r(pose, C);
    inv3(C, Ci);
    const double t[3] = {pose[4], pose[5], pose[6]};
    const double ti[3] = {-(Ci[0] * t[0] + Ci[1] * t[1] + Ci[2] * t[2]), -(Ci[3] * t[0] + Ci[4] * t[1] + Ci[5] * t[2]), -(Ci[6] * t[0] + Ci[7] * t[1] + Ci[8] * t[2])};
    double* o = stage[threadIdx.x];
    for (int i = 0; i < 9; ++i) o[i] = Ci[i];
    for (int i = 0; i < 3; ++i) o[9 + i] = ti[i];
    double bt[36], M[36], Pv[36];
    box_times(Ci, ti, bt);
    pose_jac(t, M);
    mul6(bt, M, Pv);
    for (int i = 0; i < 36; ++i) o[12 + i] = -Pv[i];
  }
  __syncthreads();
  const int n_here = min(SET_PREP_THREADS, p.n_sets - s0);
  double* out = p.set_prep + (size_t)s0 * SETPREP_STRIDE;
  for (int i = threadIdx.x; i < n_here * SETPREP_STRIDE; i += SET_PREP_THREADS) out[i] = stage[i / SETPREP_STRIDE][i % SETPREP_STRIDE];
}

// T_cam_w from the per-set inverse pose and the per-camera chain
__device__ __forceinline__ void view_transform_prepped(const double* __restrict__ sp, const double* __restrict__ camT, double Rcw[9], double tcw[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Rcw[i * 3 + j] = camT[i * 3 + 0] * sp[0 * 3 + j] + camT[i * 3 + 1] * sp[1 * 3 + j] + camT[i * 3 + 2] * sp[2 * 3 + j];
    tcw[i] = camT[i * 3 + 0] * sp[9] + camT[i * 3 + 1] * sp[10] + camT[i * 3 + 2] * sp[11] + camT[9 + i];
  }
}

// =========================================================================================================
// evaluate (residual only)
// =========================================================================================================
constexpr int EVAL_THREADS = 256;

// M-estimator weight of a term from its raw squared error e^T invR e (BE/src/MEstimatorPolicies.cpp:16-21 none, :61-75 Huber,
// :41-57 Cauchy, :23-39 Geman-McClure, :101-103 Blake-Zisserman with epsilon precomputed on the host as in :121-124)
__device__ __forceinline__ double mest_weight(int kind, double prm, double s) {
  switch (kind) {
    case 1: return s < prm * prm ? 1.0 : prm / sqrt(s);
    case 2: return 1.0 / (1.0 + s / prm);
    case 3: { const double se = prm + s; return prm / (se * se); }
    case 4: { const double ex = exp(-s); return ex / (ex + prm); }
  }
  return 1.0;
}
// The weighting of one term: t = sqrtInvR^T e, raw = t^T t = e^T invR e, w = policy(raw).  Returns the 2x2 matrix
// A = sw sqrtInvR^T (row-major a, b, c, d) that turns the raw rows into the weighted ones (sw = sqrt(w) when the M-estimator
// also scales rows, else 1), replaces (e0, e1) by A e and gives the term's cost contribution w * raw.
__device__ __forceinline__ double term_weighting(const DevProblem& p, double& e0, double& e1, double& a, double& b, double& c, double& d) {
  const double t0 = p.sT[0] * e0 + p.sT[1] * e1, t1 = p.sT[2] * e0 + p.sT[3] * e1;
  const double raw = t0 * t0 + t1 * t1;
  const double w = mest_weight(p.mest_kind, p.mest_param, raw);
  const double sw = p.mest_rows ? sqrt(w) : 1.0;
  a = sw * p.sT[0]; b = sw * p.sT[1]; c = sw * p.sT[2]; d = sw * p.sT[3];
  e0 = sw * t0;
  e1 = sw * t1;
  return w * raw;
}

template <int MODEL, bool WEIGHTED>
__global__ void __launch_bounds__(EVAL_THREADS) evaluate_kernel(DevProblem p, const int* __restrict__ view_list, int n_list) {
  extern __shared__ double smem_target[];
  for (int i = threadIdx.x; i < p.n_target * 3; i += blockDim.x) smem_target[i] = p.target[i];
  __syncthreads();
  using Cam = Camera<MODEL, false>;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int vi = warp; vi < n_list; vi += n_warps) {
    const int view = view_list[vi];
    const int set = p.view_set[view], cam = p.view_cam[view];
    const int b = p.view_begin[view], e = p.view_begin[view + 1];
    double Rcw[9], tcw[3];
    view_transform(p.set_poses + (size_t)set * POSE_STRIDE, p.camT + cam * 12, Rcw, tcw);
    double prm[CAM_PARAM_STRIDE];
#pragma unroll
    for (int i = 0; i < CAM_PARAM_STRIDE; ++i) prm[i] = p.cam_params[cam * CAM_PARAM_STRIDE + i];
    double cost = 0.0;
    for (int i = b + lane; i < e; i += 32) {
      const double* pt = smem_target + 3 * p.corner[i];
      const double pc[3] = {Rcw[0] * pt[0] + Rcw[1] * pt[1] + Rcw[2] * pt[2] + tcw[0], Rcw[3] * pt[0] + Rcw[4] * pt[1] + Rcw[5] * pt[2] + tcw[1],
                            Rcw[6] * pt[0] + Rcw[7] * pt[1] + Rcw[8] * pt[2] + tcw[2]};
      Linearisation<Cam::P, Cam::D> L;
      Cam::eval(prm, pc, L);
      double e0 = p.y_u[i] - L.y[0], e1 = p.y_v[i] - L.y[1];
      if (!L.valid) {
        e0 = 0.0;
        e1 = 0.0;
        atomicAdd(p.n_invalid, 1u);
      }
      if (WEIGHTED) {
        double wa, wb, wc, wd;
        cost += term_weighting(p, e0, e1, wa, wb, wc, wd);
      } else {
        cost += e0 * e0 + e1 * e1;
      }
      reinterpret_cast<double2*>(p.e)[i] = make_double2(-e0, -e1);
    }
    cost = warp_sum(cost);
    if (lane == 0) p.view_cost[view] = cost;
  }
}

// per camera: sums of the raw residuals (pass 0) and of their squared deviations from the mean (pass 1) over the camera's views, in
// a fixed order (thread-strided, then the block tree).  acc[cam] = {n, sum e_u, sum e_v, -, ssd_u, ssd_v, -, -}; between the passes
// the host makes the first three global (multi-rank).  ≙ K2/include/kalibr2/CameraCalibrator.hpp:368-405 (two-pass mean / variance)
__global__ void __launch_bounds__(1024) reproj_stats_kernel(DevProblem p, const double* __restrict__ e_raw, const int* __restrict__ cam_view_list,
                                                            const int* __restrict__ cam_view_begin, int pass, double* __restrict__ acc) {
  __shared__ double sh[3][32];
  const int cam = blockIdx.x;
  double* a = acc + cam * 8;
  double mu = 0.0, mv = 0.0;
  if (pass == 1 && a[0] > 0.0) {
    mu = a[1] / a[0];
    mv = a[2] / a[0];
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int vi = cam_view_begin[cam]; vi < cam_view_begin[cam + 1]; ++vi) {
    const int view = cam_view_list[vi];
    const int b = p.view_begin[view], e = p.view_begin[view + 1];
    for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
      const double2 r = reinterpret_cast<const double2*>(e_raw)[i];
      const double eu = -r.x, ev = -r.y;  // y - y_hat
      if (pass == 0) {
        s0 += 1.0;
        s1 += eu;
        s2 += ev;
      } else {
        s1 += (eu - mu) * (eu - mu);
        s2 += (ev - mv) * (ev - mv);
      }
    }
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; sh[2][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s0 = warp_sum(sh[0][threadIdx.x]); s1 = warp_sum(sh[1][threadIdx.x]); s2 = warp_sum(sh[2][threadIdx.x]);
    if (threadIdx.x == 0) {
      if (pass == 0) { a[0] = s0; a[1] = s1; a[2] = s2; a[3] = 0.0; }
      else { a[4] = s1; a[5] = s2; a[6] = 0.0; a[7] = 0.0; }
    }
  }
}

// deterministic sum of n doubles -> out[0] (fixed thread/tree order)
__global__ void __launch_bounds__(1024) sum_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s;
  }
}

// =========================================================================================================
// linearise: rows of the local design matrix of one term.
//   a_r = [ J_xi(r, 0..5) | -Ji(r, 0..P-1) | -Jd(r, 0..D-1) | 0.. | e_r ]   (16 columns, e in column 15)
//   J_xi = -Jp * boxMinus(p_c) = -[Jp | Jp [p_c]x]   (BX/src/HomogeneousExpressionNode.cpp:77-82, transformations.cpp:45-53)
// =========================================================================================================
// The rows go straight to the warp's staging buffer in shared memory, transposed: column c of the u-row of the term of
// lane l at xt[c * XT_LD + l], of its v-row at xt[c * XT_LD + 32 + l] (conflict-free stores, and no 32 values to keep
// live in registers).  Returns the projection's validity flag; e0, e1 = the residuals.
constexpr int XT_LD = 68;  // 64 rows (32 terms x 2) + 4: conflict-free for the row stores and the DMMA operand loads
constexpr int XT_WARP_DOUBLES = GRAM_DIM * XT_LD;

// CHECK: *nonfinite is set when any staged Jacobian value of the term is NaN or Inf (value * 0 is then NaN) - the materialising
// kernel needs to know, because its J = A B product on the tensor pipe would spread such a value over the structural zeros of B.
template <int MODEL, bool WEIGHTED, bool CHECK = false>
__device__ __forceinline__ bool term_rows(const DevProblem& p, const double* __restrict__ prm, const double Rcw[9], const double tcw[3],
                                          const double* __restrict__ pt, double yu, double yv, double* __restrict__ xt_lane, double& e0, double& e1,
                                          bool* nonfinite = nullptr) {
  using Cam = Camera<MODEL, true, true>;  // negated Jacobians: e = y - y_hat
  constexpr int P = Cam::P, D = Cam::D;
  const double pc[3] = {Rcw[0] * pt[0] + Rcw[1] * pt[1] + Rcw[2] * pt[2] + tcw[0], Rcw[3] * pt[0] + Rcw[4] * pt[1] + Rcw[5] * pt[2] + tcw[1],
                        Rcw[6] * pt[0] + Rcw[7] * pt[1] + Rcw[8] * pt[2] + tcw[2]};
  Linearisation<P, D> L;
  Cam::eval(prm, pc, L);
  e0 = yu - L.y[0];
  e1 = yv - L.y[1];
  if constexpr (!WEIGHTED) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double* __restrict__ x = xt_lane + 32 * r;
      const double j0 = L.Jp[r][0], j1 = L.Jp[r][1], j2 = L.Jp[r][2];  // = -Jp
      x[0 * XT_LD] = j0;
      x[1 * XT_LD] = j1;
      x[2 * XT_LD] = j2;
      x[3 * XT_LD] = j1 * pc[2] - j2 * pc[1];
      x[4 * XT_LD] = j2 * pc[0] - j0 * pc[2];
      x[5 * XT_LD] = j0 * pc[1] - j1 * pc[0];
#pragma unroll
      for (int c = 0; c < P; ++c) x[(6 + c) * XT_LD] = L.Ji[r][c];
#pragma unroll
      for (int c = 0; c < D; ++c) x[(6 + P + c) * XT_LD] = L.Jd[r][c];
#pragma unroll
      for (int c = 6 + P + D; c < E_COL; ++c) x[c * XT_LD] = 0.0;
      x[E_COL * XT_LD] = r == 0 ? e0 : e1;
    }
    if constexpr (CHECK) {
      double chk = 0.0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) chk = fma(L.Jp[r][c], 0.0, chk);
#pragma unroll
        for (int c = 0; c < P; ++c) chk = fma(L.Ji[r][c], 0.0, chk);
#pragma unroll
        for (int c = 0; c < D; ++c) chk = fma(L.Jd[r][c], 0.0, chk);
      }
      chk = fma(pc[0] + pc[1] + pc[2], 0.0, chk);
      *nonfinite = chk != chk;
    }
  } else {
    // weighted rows: [u-row; v-row] <- sqrt(w) sqrtInvR^T [u-row; v-row], column by column
    double wa, wb, wc, wd;
    if (L.valid) term_weighting(p, e0, e1, wa, wb, wc, wd);
    else { wa = wb = wc = wd = 0.0; }
    auto put = [&](int col, double u, double v) {
      xt_lane[col * XT_LD] = wa * u + wb * v;
      xt_lane[col * XT_LD + 32] = wc * u + wd * v;
    };
    const double u0 = L.Jp[0][0], u1 = L.Jp[0][1], u2 = L.Jp[0][2], v0 = L.Jp[1][0], v1 = L.Jp[1][1], v2 = L.Jp[1][2];
    put(0, u0, v0);
    put(1, u1, v1);
    put(2, u2, v2);
    put(3, u1 * pc[2] - u2 * pc[1], v1 * pc[2] - v2 * pc[1]);
    put(4, u2 * pc[0] - u0 * pc[2], v2 * pc[0] - v0 * pc[2]);
    put(5, u0 * pc[1] - u1 * pc[0], v0 * pc[1] - v1 * pc[0]);
#pragma unroll
    for (int c = 0; c < P; ++c) put(6 + c, L.Ji[0][c], L.Ji[1][c]);
#pragma unroll
    for (int c = 0; c < D; ++c) put(6 + P + c, L.Jd[0][c], L.Jd[1][c]);
#pragma unroll
    for (int c = 6 + P + D; c < E_COL; ++c) {
      xt_lane[c * XT_LD] = 0.0;
      xt_lane[c * XT_LD + 32] = 0.0;
    }
    xt_lane[E_COL * XT_LD] = e0;       // already A e
    xt_lane[E_COL * XT_LD + 32] = e1;
    if constexpr (CHECK) {
      double chk = 0.0;
#pragma unroll
      for (int c = 0; c < E_COL; ++c) chk = fma(xt_lane[c * XT_LD], 0.0, fma(xt_lane[c * XT_LD + 32], 0.0, chk));
      *nonfinite = chk != chk;
    }
  }
  return L.valid;
}
// zero weight (inactive lane of a partial chunk, or a projection that bailed out: SURVEY.md Q6): the rows become zero.
// Only reached when some lane of the warp needs it, so the common case spends nothing on it.
__device__ __forceinline__ void zero_rows(double* __restrict__ xt_lane) {
#pragma unroll
  for (int c = 0; c < GRAM_DIM; ++c) {
    xt_lane[c * XT_LD] = 0.0;
    xt_lane[c * XT_LD + 32] = 0.0;
  }
}

// ---- fused linearise + assemble on DMMA ----------------------------------------------------------------------
// One warp walks a slice of the camera-sorted view list (a slice never crosses a camera).  Per view: 32-term chunks are
// linearised (lane per term), staged transposed in shared memory (u-rows of the 32 terms, then their v-rows) and reduced
// into the 16x16 Gram block G with mma.sync.m8n8k4.f64 (two accumulator sets: u-rows / v-rows); the next chunk's
// observations - or the next view's first chunk - are prefetched before the DMMA phase.  Per view the kernel writes only
// the two Gram tiles the set reduction needs (G[0:8][0:16]: pose x pose, pose x intrinsics, pose x e) straight from the
// accumulator registers, and adds all three tiles to the slice's per-camera Gram sum; the M = Pi_k P_v products that turn
// G into V_v / W_v / b_v live in set_reduce_kernel, where the FP64 units are idle.
#ifndef KB_LA_WARPS
#define KB_LA_WARPS 8
#endif
#ifndef KB_LA_VARIANT
#define KB_LA_VARIANT 0  // ablation builds only (tools/build_variants.sh): 1 = no DMMA, 2 = no projection math
#endif
constexpr int LA_WARPS = KB_LA_WARPS;
constexpr int LA_THREADS = LA_WARPS * 32;
constexpr int LA_SP_DOUBLES = 16;  // per view: C^-1 (9), -C^-1 t (3) of the set, padded
// per warp: XT | slice Gram sum | set constants x2 | [R_cw | t_cw] of the view (12) + the camera's parameters (10, padded to 12)
constexpr int LA_WARP_DOUBLES = XT_WARP_DOUBLES + GRAM_TILES + 2 * LA_SP_DOUBLES + 24;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// one k-step pair (4 u-rows, 4 v-rows) of the Gram reduction: six DMMAs on three accumulator chains
__device__ __forceinline__ void gram_kstep(double (&c00)[2], double (&c01)[2], double (&c11)[2], double x0, double x1, double z0, double z1) {
  dmma(c00[0], c00[1], x0, x0);
  dmma(c01[0], c01[1], x0, x1);
  dmma(c11[0], c11[1], x1, x1);
  dmma(c00[0], c00[1], z0, z0);
  dmma(c01[0], c01[1], z0, z1);
  dmma(c11[0], c11[1], z1, z1);
}

template <int MODEL, bool WRITE_E, bool WEIGHTED>
__global__ void __launch_bounds__(LA_THREADS, 2) linearise_assemble_kernel(DevProblem p, const int4* __restrict__ vmeta,
                                                                            const int4* __restrict__ slices, int slice_lo, int slice_hi) {
  extern __shared__ __align__(16) double smem[];
  // target corners, one array per coordinate (a warp's 32 consecutive corners then cost 2 wavefronts per coordinate, not 6)
  const int tpad = (p.n_target + 1) & ~1;
  double* s_tx = smem;
  double* s_ty = s_tx + tpad;
  double* s_tz = s_ty + tpad;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  double* XT = smem + 3 * tpad + wib * LA_WARP_DOUBLES;       // [16 cols][68]
  double* sG = XT + XT_WARP_DOUBLES;                           // [3][64] slice sum of the three tiles
  double* sSP = sG + GRAM_TILES;                               // [2][16] per-set constants of the current / next view (cp.async)
  double* sC = sSP + 2 * LA_SP_DOUBLES;                        // [R_cw | t_cw] of the current view
  double* prm = sC + 12;                                       // parameters of the slice's camera
  for (int i = threadIdx.x; i < p.n_target; i += blockDim.x) {  // constant data: staged while the preceding kernel drains
    s_tx[i] = p.target[3 * i];
    s_ty[i] = p.target[3 * i + 1];
    s_tz[i] = p.target[3 * i + 2];
  }
  pdl_enter();
  if (p.ctrl->done || p.ctrl->skip_eval) return;
  __syncthreads();

  const int warp = blockIdx.x * LA_WARPS + wib;
  const int n_warps = gridDim.x * LA_WARPS;
  const int arow = lane >> 2, acol = lane & 3;
  // element of [R_cw | t_cw] this lane computes per view (lanes 0..11): row tr of camT times column tj of the set's C^-1 / -C^-1 t
  const int tl = lane < 12 ? lane : 0;
  const int tr = tl < 9 ? tl / 3 : tl - 9;
  const int sp_base = tl < 9 ? tl % 3 : 9, sp_stride = tl < 9 ? 3 : 1;
  for (int sl = slice_lo + warp; sl < slice_hi; sl += n_warps) {
    const int4 S = slices[sl];
    const int cam = S.z;
    __syncwarp();
    if (lane < CAM_PARAM_STRIDE) prm[lane] = p.cam_params[cam * CAM_PARAM_STRIDE + lane];
    for (int o = lane; o < GRAM_TILES; o += 32) sG[o] = 0.0;
    const double ct0 = __ldg(p.camT + cam * 12 + tr * 3), ct1 = __ldg(p.camT + cam * 12 + tr * 3 + 1), ct2 = __ldg(p.camT + cam * 12 + tr * 3 + 2);
    const double ct3 = tl < 9 ? 0.0 : __ldg(p.camT + cam * 12 + 9 + tr);
    // pipeline: metadata two views ahead (registers), per-set constants one view ahead (cp.async into shared memory),
    // observations one chunk ahead (registers), across view boundaries too
    int4 m_cur = vmeta[S.x];                                   // (view, set, begin, end)
    int4 m_nxt = (S.x + 1 < S.y) ? vmeta[S.x + 1] : m_cur;
    if (lane < 6) cp_async_16(sSP + 2 * lane, p.set_prep + (size_t)m_cur.y * SETPREP_STRIDE + 2 * lane);
    cp_async_commit();
    int i = 0;
    bool active = false, have_pf = false;
    double yu = 0.0, yv = 0.0;
    int cid = 0;
    for (int vi = S.x; vi < S.y; ++vi) {
      const int buf = (vi - S.x) & 1;
      const int view = m_cur.x;
      const int b = m_cur.z, e = m_cur.w;
      if (!have_pf) {  // first view of the slice, or the previous view was empty
        i = b + lane;
        active = i < e;
        const int ii = active ? i : b;
        if (b < e) { yu = p.y_u[ii]; yv = p.y_v[ii]; cid = p.corner[ii]; }
      }
      have_pf = false;
      const int4 m_nn = (vi + 2 < S.y) ? vmeta[vi + 2] : m_nxt;
      cp_async_wait_all();
      __syncwarp();  // the set constants have landed; every lane is done with the previous view's transform
      {
        const double* sp = sSP + buf * LA_SP_DOUBLES + sp_base;
        const double v = ct0 * sp[0] + ct1 * sp[sp_stride] + ct2 * sp[2 * sp_stride] + ct3;
        if (lane < 12) sC[lane] = v;
      }
      __syncwarp();
      if (vi + 1 < S.y && lane < 6) cp_async_16(sSP + (buf ^ 1) * LA_SP_DOUBLES + 2 * lane, p.set_prep + (size_t)m_nxt.y * SETPREP_STRIDE + 2 * lane);
      cp_async_commit();
      double c00[2] = {0.0, 0.0}, c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};  // three independent DMMA chains saturate the pipe
      for (int base = b; base < e; base += 32) {
        const double cyu = yu, cyv = yv;
        const int ccid = cid;
        const bool cactive = active;
        const int ci = i;
        if (base + 32 < e) {  // next chunk of this view
          i = base + 32 + lane;
          active = i < e;
          const int ii = active ? i : b;
          yu = p.y_u[ii]; yv = p.y_v[ii]; cid = p.corner[ii];
        } else if (vi + 1 < S.y && m_nxt.z < m_nxt.w) {  // first chunk of the next view
          i = m_nxt.z + lane;
          active = i < m_nxt.w;
          const int ii = active ? i : m_nxt.z;
          yu = p.y_u[ii]; yv = p.y_v[ii]; cid = p.corner[ii];
          have_pf = true;
        }
        const int n = min(32, e - base);
        double e0, e1;
#if KB_LA_VARIANT == 2
        const bool valid = true;
        e0 = cyu + s_tx[ccid]; e1 = cyv;
#else
        const double pt[3] = {s_tx[ccid], s_ty[ccid], s_tz[ccid]};
        const bool valid = term_rows<MODEL, WEIGHTED>(p, prm, sC, sC + 9, pt, cyu, cyv, XT + lane, e0, e1);
#endif
        if (cactive && !valid) atomicAdd(p.n_invalid, 1u);
        const bool keep = cactive && valid;
        const int kq = (n + 3) >> 2;  // DMMA k-steps per row half: a partial last chunk only spends DMMAs on rows that exist
        if (__any_sync(0xffffffffu, !keep && lane < 4 * kq)) {
          if (!keep) {
            zero_rows(XT + lane);
            e0 = 0.0;
            e1 = 0.0;
          }
        }
        if (WRITE_E && cactive) reinterpret_cast<double2*>(p.e)[ci] = make_double2(neg_int(e0), neg_int(e1));
        __syncwarp();
        const double* xa = XT + arow * XT_LD + acol;
#if KB_LA_VARIANT != 1
        if (n == 32) {  // full chunk: eight k-steps, straight-line
#pragma unroll
          for (int s = 0; s < 8; ++s) gram_kstep(c00, c01, c11, xa[4 * s], xa[8 * XT_LD + 4 * s], xa[32 + 4 * s], xa[8 * XT_LD + 32 + 4 * s]);
        } else {
#pragma unroll 1
          for (int s = 0; s < kq; ++s) gram_kstep(c00, c01, c11, xa[4 * s], xa[8 * XT_LD + 4 * s], xa[32 + 4 * s], xa[8 * XT_LD + 32 + 4 * s]);
        }
#endif
        __syncwarp();
      }
      // ---- per view: tiles (0,0) and (0,1) row-major (the C fragment order is row-major 8x8), slice sum of all three ----
      {
        double2* vb = reinterpret_cast<double2*>(p.VB + (size_t)view * VB_STRIDE);
        vb[lane] = make_double2(c00[0], c00[1]);
        vb[32 + lane] = make_double2(c01[0], c01[1]);
        double2* g2 = reinterpret_cast<double2*>(sG);
        double2 t0 = g2[lane], t1 = g2[32 + lane], t2 = g2[64 + lane];
        t0.x += c00[0]; t0.y += c00[1]; t1.x += c01[0]; t1.y += c01[1]; t2.x += c11[0]; t2.y += c11[1];
        g2[lane] = t0; g2[32 + lane] = t1; g2[64 + lane] = t2;
      }
      m_cur = m_nxt;
      m_nxt = m_nn;
    }
    cp_async_wait_all();
    double* out = p.gram_partial + (size_t)sl * GRAM_TILES;
    for (int o = lane; o < GRAM_TILES; o += 32) out[o] = sG[o];
    __syncwarp();
  }
}

// ---- materialising linearise: e and J in the CCS J^T layout of the reference -----------------------------------
// Per term two columns of J^T (= rows of J), each W_k = 6 + 6k + P + D values ordered by design-variable block index, so
// the Jacobian of one view is a dense row-major [2 * terms][W_k] array: the product of the local rows
// a = [J_xi | -J_i | -J_d | . | e] (16 columns, lane per term) with a 16 x W_k matrix B that holds Pi_k P_v under the set-pose
// columns, A_{j,k} under the baseline columns and unit vectors under the intrinsics columns (col_desc says which).
// That product runs on the FP64 tensor pipe: the rows of a 32-term chunk are staged transposed in shared memory (same
// layout as the fused kernel), B lives in shared memory in DMMA fragment order (only the k-steps that are structurally
// non-zero per 8-column tile), and every 8x8 output tile goes from the accumulator registers straight to HBM.
// Work is handed out in sub-slices of the camera-sorted view list through an atomic counter (scheduling only: every
// value is computed the same way wherever it runs).
constexpr int LM_WARPS = 4;
constexpr int LM_THREADS = LM_WARPS * 32;
constexpr int LM_SUB = 4;  // sub-slices per slice of the fused kernel's slice table

__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream2(double* p, double v0, double v1) { __stcs(reinterpret_cast<double2*>(p), make_double2(v0, v1)); }

// A non-finite Jacobian entry must stay in the columns the reference's chain rule puts it in (J_pose = J_xi A, the intrinsics
// columns are copies): the tensor-pipe product J = A B would multiply it with the structural zeros of B.  Rare (an equidistant
// corner exactly on the optical axis): the whole 32-term chunk is then written by this scalar routine instead -
// out of line and with by-value arguments only, so that the hot path's register allocation does not see it.
__device__ __noinline__ void exact_chunk(const double* __restrict__ XT, const int* __restrict__ desc, const double* __restrict__ sM,
                                         const double* __restrict__ camA_k, double* __restrict__ dst_chunk, int W, int rows_valid, int lane) {
  for (int idx = lane; idx < rows_valid * W; idx += 32) {
    const int row = idx / W, col = idx - row * W;
    const int pos = 32 * (row & 1) + (row >> 1);  // staged position of (term row / 2, residual row % 2)
    const int dsc = desc[col];
    const int kind = dsc >> 16, j = (dsc >> 8) & 0xff, sub = dsc & 0xff;
    double v = 0.0;
    if (kind == 2) {
      v = XT[(6 + sub) * XT_LD + pos];
    } else if (kind == 0 || kind == 1) {
      for (int k = 0; k < 6; ++k) {
        const double bk = kind == 0 ? sM[k * 6 + sub] : __ldg(camA_k + (size_t)j * 36 + k * 6 + sub);
        v = fma(XT[k * XT_LD + pos], bk, v);
      }
    }
    dst_chunk[(long long)row * W + col] = v;
  }
}

template <int MODEL, bool WEIGHTED>
__global__ void __launch_bounds__(LM_THREADS, 4) linearise_materialise_kernel(DevProblem p, const int4* __restrict__ vmeta, const int4* __restrict__ slices,
                                                                               int slice_lo, int slice_hi, double* __restrict__ jt, int bfrag_pairs,
                                                                               unsigned int* __restrict__ work_counter) {
  extern __shared__ __align__(16) double smem[];
  double* s_target = smem;
  const int target_doubles = (p.n_target * 3 + 1) & ~1;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp_doubles = XT_WARP_DOUBLES + bfrag_pairs * 32 + 36 + SETPREP_STRIDE;
  double* XT = smem + target_doubles + (size_t)wib * warp_doubles;  // [16 cols][68]
  double* Bf = XT + XT_WARP_DOUBLES;                                  // [pairs][32] B fragments
  double* sM = Bf + bfrag_pairs * 32;                                 // Pi_k P_v
  double* sSP = sM + 36;                                              // per-set constants
  for (int i = threadIdx.x; i < p.n_target * 3; i += blockDim.x) s_target[i] = p.target[i];
  __syncthreads();
  constexpr int P = model_P(MODEL), D = model_D(MODEL);
  const int arow = lane >> 2, acol = lane & 3;
  const int n_work = (slice_hi - slice_lo) * LM_SUB;
  for (;;) {
    int w = 0;
    if (lane == 0) w = (int)atomicAdd(work_counter, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= n_work) break;
    const int4 S = slices[slice_lo + w / LM_SUB];
    const int part = w % LM_SUB, nv = S.y - S.x;
    const int v_lo = S.x + (int)((long long)nv * part / LM_SUB), v_hi = S.x + (int)((long long)nv * (part + 1) / LM_SUB);
    if (v_lo >= v_hi) continue;
    const int cam = S.z;
    double prm[CAM_PARAM_STRIDE];
#pragma unroll
    for (int i = 0; i < CAM_PARAM_STRIDE; ++i) prm[i] = p.cam_params[cam * CAM_PARAM_STRIDE + i];
    double camT[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) camT[i] = p.camT[cam * 12 + i];
    const int W = 6 + 6 * cam + P + D;
    const int ntiles = (W + 7) >> 3;
    const bool w_even = (W & 1) == 0;
    const int* __restrict__ desc = p.col_desc + (size_t)cam * p.col_desc_stride;
    for (int vi = v_lo; vi < v_hi; ++vi) {
      const int4 m = vmeta[vi];  // (view, set, begin, end)
      const int b = m.z, e = m.w;
      // first chunk's observations: in flight while the per-view constants are set up
      int i = b + lane;
      bool active = i < e;
      int ii = active ? i : b;
      double yu = 0.0, yv = 0.0;
      int cid = 0;
      if (b < e) { yu = p.y_u[ii]; yv = p.y_v[ii]; cid = p.corner[ii]; }
      const double* sp_g = p.set_prep + (size_t)m.y * SETPREP_STRIDE;
      sSP[lane] = sp_g[lane];
      if (lane < SETPREP_STRIDE - 32) sSP[32 + lane] = sp_g[32 + lane];
      __syncwarp();
      double Rcw[9], tcw[3];
      view_transform_prepped(sSP, camT, Rcw, tcw);
      for (int o = lane; o < 36; o += 32) {  // M = Pi_k P_v
        const int r = o / 6, c = o % 6;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) s += __ldg(p.camPi + cam * 36 + r * 6 + a) * sSP[12 + a * 6 + c];
        sM[o] = s;
      }
      __syncwarp();
      // B fragments (b: row k = lane % 4, col n = lane / 4) of the structurally non-zero (tile, k-step) pairs
      unsigned long long kmask = 0ull;
      {
        int np = 0;
        for (int nt = 0; nt < ntiles; ++nt) {
          const int c = 8 * nt + arow;
          const int d = (c < W) ? desc[c] : (3 << 16);
          const int kind = d >> 16, j = (d >> 8) & 0xff, sub = d & 0xff;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int kr = 4 * ks + acol;
            bool need = false;
            double val = 0.0;
            if (kind == 0) {
              need = ks < 2;
              if (kr < 6) val = sM[kr * 6 + sub];
            } else if (kind == 1) {
              need = ks < 2;
              if (kr < 6) val = __ldg(p.camA + ((size_t)cam * p.n_cams + j) * 36 + kr * 6 + sub);
            } else if (kind == 2) {
              need = ks == ((6 + sub) >> 2);
              val = (kr == 6 + sub) ? 1.0 : 0.0;
            }
            if (__any_sync(0xffffffffu, need)) {
              Bf[np * 32 + lane] = val;
              kmask |= 1ull << (4 * nt + ks);
              ++np;
            }
          }
        }
      }
      __syncwarp();
      double* __restrict__ dst_view = jt + p.view_jbase[m.x];
      for (int base = b; base < e; base += 32) {
        const double cyu = yu, cyv = yv;
        const int ccid = cid;
        const bool cactive = active;
        const int ci = i;
        i = base + 32 + lane;
        active = i < e;
        ii = active ? i : b;
        if (base + 32 < e) { yu = p.y_u[ii]; yv = p.y_v[ii]; cid = p.corner[ii]; }
        bool exact_rows = false;  // warp-uniform: some term of the chunk carries a NaN / Inf in its Jacobian
        {
          double e0, e1;
          bool nonfinite = false;
          const bool valid = term_rows<MODEL, WEIGHTED, true>(p, prm, Rcw, tcw, s_target + 3 * ccid, cyu, cyv, XT + lane, e0, e1, &nonfinite);
          if (cactive && !valid) atomicAdd(p.n_invalid, 1u);
          if (__any_sync(0xffffffffu, cactive && !valid)) {
            if (cactive && !valid) {
              zero_rows(XT + lane);
              e0 = 0.0;
              e1 = 0.0;
              nonfinite = false;
            }
          }
          exact_rows = __any_sync(0xffffffffu, cactive && nonfinite);
          if (cactive) reinterpret_cast<double2*>(p.e)[ci] = make_double2(-e0, -e1);
        }
        __syncwarp();
        const int rows_valid = 2 * min(32, e - base);
        const int mtiles = (rows_valid + 7) >> 3;
        double* __restrict__ dst_chunk = dst_view + (long long)(base - b) * 2 * W;
        if (exact_rows) {
          exact_chunk(XT, desc, sM, p.camA + (size_t)cam * p.n_cams * 36, dst_chunk, W, rows_valid, lane);
          __syncwarp();
          continue;
        }
        for (int mt = 0; mt < mtiles; ++mt) {
          double a[4];
#pragma unroll
          // fragment row arow = (term 4 * mt + arow % 4, residual arow / 4): conflict-free loads from the u-rows | v-rows staging
          for (int ks = 0; ks < 4; ++ks) a[ks] = XT[(4 * ks + acol) * XT_LD + 32 * (arow >> 2) + 4 * mt + (arow & 3)];
          const int row = 8 * mt + 2 * (arow & 3) + (arow >> 2);
          const bool row_ok = row < rows_valid;
          double* __restrict__ drow = dst_chunk + (long long)row * W + 2 * acol;
          const double* bf = Bf + lane;
          unsigned long long km = kmask;
          for (int nt = 0; nt < ntiles; ++nt, km >>= 4) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (km & (1ull << ks)) {
                dmma(c0, c1, a[ks], *bf);
                bf += 32;
              }
            const int col = 8 * nt + 2 * acol;
            if (row_ok) {
              if (w_even) {
                if (col < W) st_stream2(drow + 8 * nt, c0, c1);
              } else {
                if (col < W) st_stream(drow + 8 * nt, c0);
                if (col + 1 < W) st_stream(drow + 8 * nt + 1, c1);
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// =========================================================================================================
// set reduction: one warp per synced set.  From the two Gram tiles the fused kernel left per view (G_xx = G[0:6][0:6],
// G_cx = G[6:6+PD][0:6], G_xe = G[0:6][15]) and M = Pi_k P_v:
//   Y_k = G_xx M,  V_v = sum_k M^T Y_k,  b_v = -sum_k M^T G_xe,  W_v rows: intrinsics of camera k = G_cx M,
//   baseline j = sum_{k>j} A_{j,k}^T Y_k.                                  ≙ SparseBlockMatrix::block(r,c) += J1^T J2
// =========================================================================================================
constexpr int SR_WARPS = 8;

// Layout changes between DMMA fragments, by warp shuffles.  X is an 8x8 matrix in C-fragment layout: lane l holds
// X[l/4][2(l%4)] and X[l/4][2(l%4)+1].  The B fragment of X for k-step ks (b: row k = l%4, col n = l/4) is X[4ks + l%4][l/4];
// the A fragment of X^T for k-step ks (a: row m = l/4, col k = l%4) is the very same element.
__device__ __forceinline__ double c_to_b(double c0, double c1, int ks, int arow, int acol) {
  const int src = 4 * (4 * ks + acol) + (arow >> 1);
  const double v0 = __shfl_sync(0xffffffffu, c0, src), v1 = __shfl_sync(0xffffffffu, c1, src);
  return (arow & 1) ? v1 : v0;
}

// Everything is a chain of 8x8 FP64 tensor-core products held in registers - no shared memory, so the kernel is bounded
// only by the tile loads.  Cameras are walked from the last to the first so that the baseline rows follow from a running
//   Z_j := sum_{k>j} X_{k,j}^T Y_k = Y_{j+1} + boxTimes(B_{j+1})^T Z_{j+1},   W_base(j) = M_j^T Z_j
// (W_base(j) = sum_{k>j} A_{j,k}^T Y_k with A_{j,k} = X_{k,j} M_j, X_{k,j} = X_{k,j+1} boxTimes(B_{j+1}): O(C) instead of O(C^2)).
// Per camera k with its Gram tiles G = [G00 | G01] (rows: pose 0..5, first two intrinsics; G01 column 7 = e):
//   M  = Pi_k P_v                                  (2 DMMA)
//   R0 = G00 M       rows 0..5 = Y_k, rows 6,7 = W rows of the first two intrinsics          (2 DMMA)
//   R1 = G01^T M     rows 0..PD-3 = W rows of the remaining intrinsics, row 7 = -b_k^T       (2 DMMA)
//   V += M^T Y_k                                                                              (2 DMMA)
__device__ void camera_block_element(const DevProblem& p, int idx);
// Blocks [0, n_set_blocks) reduce the sets; the remaining blocks of the grid compute the camera block U (one element per thread).
__global__ void __launch_bounds__(SR_WARPS * 32) set_reduce_kernel(DevProblem p, int n_set_blocks) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int arow = lane >> 2, acol = lane & 3;
  const int C = p.n_cams;
  if (p.ctrl->done || !p.ctrl->need_build) return;
  if ((int)blockIdx.x >= n_set_blocks) {
    camera_block_element(p, ((int)blockIdx.x - n_set_blocks) * SR_WARPS * 32 + threadIdx.x);
    return;
  }
  // operand element of a 6x6 row-major matrix, zero padded to 8x8:  direct[ks] = X[arow][4ks + acol],  transposed[ks] = X[4ks + acol][arow]
  const bool in0 = arow < 6, in1 = arow < 6 && acol < 2;  // k-step 0: k = acol < 4 ; k-step 1: k = 4 + acol < 6
  const int d0 = arow * 6 + acol, d1 = arow * 6 + 4 + acol;
  const int t0 = acol * 6 + arow, t1 = (4 + acol) * 6 + arow;
  for (int set = blockIdx.x * SR_WARPS + wib; set < p.n_sets; set += n_set_blocks * SR_WARPS) {
    double* __restrict__ Wout = p.W + (size_t)set * p.n_c * 6;
    // lane k fetches the view of camera k (n_cams <= 32), so that the tile loads below have no dependent address chain
    int my_view = -1;
    if (lane < C) {
      const int w = p.set_view[(size_t)set * C + lane];
      if (w >= 0 && p.view_begin[w + 1] > p.view_begin[w]) my_view = w;
    }
    const double* pv = p.set_prep + (size_t)set * SETPREP_STRIDE + 12;
    const double pb0 = in0 ? pv[t0] : 0.0, pb1 = in1 ? pv[t1] : 0.0;  // B fragments of P_v
    double v0 = 0.0, v1 = 0.0;    // V_v accumulator (C layout)
    double bs0 = 0.0, bs1 = 0.0;  // running R1 (row 7 = -b_v^T)
    double z0 = 0.0, z1 = 0.0;    // Z (C layout)
    double zb0 = 0.0, zb1 = 0.0;  // its B fragments
    // A fragments of [G00 ; G01^T] of the camera about to be processed (zeros when it has no view in this set)
    double g00a = 0.0, g00b = 0.0, g01a = 0.0, g01b = 0.0;
    {
      const int view = __shfl_sync(0xffffffffu, my_view, C - 1);
      if (view >= 0) {
        const double* t = p.VB + (size_t)view * VB_STRIDE;
        g00a = t[arow * 8 + acol]; g00b = t[arow * 8 + 4 + acol];
        g01a = t[64 + acol * 8 + arow]; g01b = t[64 + (4 + acol) * 8 + arow];
      }
    }
    for (int k = C - 1; k >= 0; --k) {
      const double a00 = g00a, a01 = g00b, a10 = g01a, a11 = g01b;
      if (k > 0) {  // next camera's tiles: in flight during this camera's products
        const int view = __shfl_sync(0xffffffffu, my_view, k - 1);
        g00a = g00b = g01a = g01b = 0.0;
        if (view >= 0) {
          const double* t = p.VB + (size_t)view * VB_STRIDE;
          g00a = t[arow * 8 + acol]; g00b = t[arow * 8 + 4 + acol];
          g01a = t[64 + acol * 8 + arow]; g01b = t[64 + (4 + acol) * 8 + arow];
        }
      }
      const int PD = p.cam_P[k] + p.cam_D[k];
      // M = Pi_k P_v
      const double* Pi = p.camPi + k * 36;
      double m0 = 0.0, m1 = 0.0;
      dmma(m0, m1, in0 ? __ldg(Pi + d0) : 0.0, pb0);
      dmma(m0, m1, in1 ? __ldg(Pi + d1) : 0.0, pb1);
      const double mb0 = c_to_b(m0, m1, 0, arow, acol), mb1 = c_to_b(m0, m1, 1, arow, acol);  // B fragments of M = A fragments of M^T
      double r00 = 0.0, r01 = 0.0, r10 = 0.0, r11 = 0.0;
      dmma(r00, r01, a00, mb0);
      dmma(r10, r11, a10, mb0);
      dmma(r00, r01, a01, mb1);
      dmma(r10, r11, a11, mb1);
      // W rows of the intrinsics of camera k (columns 0..5 of R: acol < 3)
      if (acol < 3) {
        double* w = Wout + (size_t)p.intr_off[k] * 6 + 2 * acol;
        if (arow >= 6 && arow - 6 < PD) *reinterpret_cast<double2*>(w + (arow - 6) * 6) = make_double2(r00, r01);
        if (arow < 7 && arow + 2 < PD) *reinterpret_cast<double2*>(w + (arow + 2) * 6) = make_double2(r10, r11);
      }
      bs0 += r10;
      bs1 += r11;
      // Y_k = rows 0..5 of R0
      const double y0 = in0 ? r00 : 0.0, y1 = in0 ? r01 : 0.0;
      const double yb0 = c_to_b(y0, y1, 0, arow, acol), yb1 = c_to_b(y0, y1, 1, arow, acol);
      dmma(v0, v1, mb0, yb0);
      dmma(v0, v1, mb1, yb1);
      if (k >= 1) {  // baseline j = k - 1
        const int j = k - 1;
        double n0 = y0, n1 = y1;
        if (k < C - 1) {  // Z_j = Y_k + boxTimes(B_k)^T Z_k
          const double* Bt = p.baseBt + k * 36;
          dmma(n0, n1, in0 ? __ldg(Bt + t0) : 0.0, zb0);
          dmma(n0, n1, in1 ? __ldg(Bt + t1) : 0.0, zb1);
        }
        z0 = n0;
        z1 = n1;
        zb0 = c_to_b(z0, z1, 0, arow, acol);
        zb1 = c_to_b(z0, z1, 1, arow, acol);
        const double* Mj = p.baseM + j * 36;
        double w0 = 0.0, w1 = 0.0;  // W_base(j) = M_j^T Z_j
        dmma(w0, w1, in0 ? __ldg(Mj + t0) : 0.0, zb0);
        dmma(w0, w1, in1 ? __ldg(Mj + t1) : 0.0, zb1);
        if (in0 && acol < 3) *reinterpret_cast<double2*>(Wout + ((size_t)p.base_off[j] + arow) * 6 + 2 * acol) = make_double2(w0, w1);
      }
    }
    if (in0 && acol < 3) *reinterpret_cast<double2*>(p.V + (size_t)set * 36 + arow * 6 + 2 * acol) = make_double2(v0, v1);
    if (arow == 7 && acol < 3) *reinterpret_cast<double2*>(p.bv + (size_t)set * 6 + 2 * acol) = make_double2(-bs0, -bs1);
  }
}

// ---- per-camera Gram sums from the slice partials (fixed order) -------------------------------------------------
constexpr int FG_GROUPS = 5;
// cam_slice_range: [n_cams][n_ranges][2] slice ranges of camera k (one per chunk of the streamed slice table)
// end of an iteration of the device-resident loop: accept / reject of the evaluated step (skipped after a failed solve), trace, loop
// condition, and - unless the loop has ended - the trust-region decisions of the NEXT iteration (build or not, lambda, damping)
__device__ __forceinline__ void lm_boundary(LmCtrl* c, double* __restrict__ trace, int* __restrict__ pos_def, bool evaluated) {
  if (evaluated) {
    const int it = c->iterations;
    kalibr_b200::lm_after_eval(c, c->cost_new);
    trace[3 * it] = c->J;
    trace[3 * it + 1] = c->deltaX;
    trace[3 * it + 2] = c->lambda;
  }
  if (!c->done) {
    kalibr_b200::lm_before_solve(c);
    pos_def[0] = 1;
  }
}

// The last block to finish (ticket counter) adds the cameras' e^T e entries in camera order: the cost at the linearisation point
// (-> cost_out, and into every rank's slot as the producer of peer exchange C).  lm_mode 1 (single-rank device loop): it then also
// runs the loop boundary, so that an iteration ends with this launch.
__global__ void __launch_bounds__(GRAM_TILES * FG_GROUPS) finalize_gram_kernel(DevProblem p, const int* __restrict__ cam_slice_range, int n_ranges,
                                                                               double* __restrict__ cost_out, int exchange, int lm_mode,
                                                                               double* __restrict__ trace, int* __restrict__ pos_def) {
  pdl_enter();
  __shared__ double sh[FG_GROUPS][GRAM_TILES];
  __shared__ int s_last;
  const int k = blockIdx.x, t = threadIdx.x % GRAM_TILES, g = threadIdx.x / GRAM_TILES;
  if (p.ctrl->done) return;
  if (p.ctrl->skip_eval) {  // failed solve: nothing was evaluated, but the next iteration still needs its decisions
    if (lm_mode == 1 && blockIdx.x == 0 && threadIdx.x == 0) lm_boundary(p.ctrl, trace, pos_def, false);
    return;
  }
  double s = 0.0;
  for (int r = 0; r < n_ranges; ++r) {
    const int lo = cam_slice_range[(k * n_ranges + r) * 2], hi = cam_slice_range[(k * n_ranges + r) * 2 + 1];
    const int per = (hi - lo + FG_GROUPS - 1) / FG_GROUPS;
    const int a = lo + g * per, b = min(hi, a + per);
#pragma unroll 16
    for (int sl = a; sl < b; ++sl) s += p.gram_partial[(size_t)sl * GRAM_TILES + t];
  }
  sh[g][t] = s;
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int q = 1; q < FG_GROUPS; ++q) s += sh[q][t];
    // tile t/64: 0 -> (0,0), 1 -> (0,1), 2 -> (1,1); within a tile the mma C layout: lane = (t%64)/2, element = t%2
    const int tile = t >> 6, lane = (t & 63) >> 1, el = t & 1;
    const int r = (tile == 2 ? 8 : 0) + (lane >> 2), c = (tile >= 1 ? 8 : 0) + 2 * (lane & 3) + el;
    double* G = p.sumG + (size_t)k * GRAM_SIZE;
    G[r * GRAM_DIM + c] = s;
    if (tile == 1) G[c * GRAM_DIM + r] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // this block's Gram sum is visible before its ticket
    const unsigned int ticket = atomicAdd(p.tickets, 1u);
    s_last = ticket == gridDim.x - 1;
    if (s_last) {
      p.tickets[0] = 0u;
      __threadfence();
      double cost = 0.0;
      for (int q = 0; q < p.n_cams; ++q) cost += __ldcg(p.sumG + (size_t)q * GRAM_SIZE + E_COL * GRAM_DIM + E_COL);
      cost_out[0] = cost;
      if (p.px.enabled && exchange) {  // exchange C: this rank's cost into every rank's slot
        const unsigned long long e = px_next_epoch(p.px, 2);
        for (int r = 0; r < p.px.n_ranks; ++r) p.px.base[r][px_off_c(p.px, (int)(e & 1), p.px.rank)] = cost;
        __threadfence_system();
        px_signal(p.px, 2, e);
      }
      if (lm_mode == 1) {
        p.ctrl->cost_new = cost;
        lm_boundary(p.ctrl, trace, pos_def, true);
      }
    }
  }
}

// ---- camera block U (augmented with b_c as last row/col and the linearisation-point cost in the corner) ----------
__device__ __forceinline__ double sumg_sym(const double* __restrict__ G, int r, int c) {
  if (r >= 8 && c < 8) return G[c * GRAM_DIM + r];
  if ((r < 8) == (c < 8) && r > c) return G[c * GRAM_DIM + r];  // diagonal tiles: use the upper triangle for exact symmetry
  return G[r * GRAM_DIM + c];
}

// descriptor of one reduced-system index: which camera-side variable it belongs to
struct RedIndex {
  int kind;  // 0 = intrinsics of camera `id` (local column 6 + sub), 1 = baseline `id` (sub in 0..5), 2 = the augmented rhs row
  int id;
  int sub;
};
__device__ __forceinline__ RedIndex red_index(const DevProblem& p, int i) {
  RedIndex r;
  if (i == p.n_c) { r.kind = 2; r.id = 0; r.sub = 0; return r; }
  if (p.n_cams > 1 && i >= p.base_off[0]) {
    r.kind = 1; r.id = (i - p.base_off[0]) / 6; r.sub = (i - p.base_off[0]) % 6; return r;
  }
  int k = 0;
  while (k + 1 < p.n_cams && i >= p.intr_off[k + 1]) ++k;
  r.kind = 0; r.id = k; r.sub = i - p.intr_off[k];
  return r;
}
// "local Jacobian column" of reduced index ri as seen from camera k: a 16-vector c such that J_red = J_local * c, either a
// signed unit vector (intrinsics: column 6 + sub; the augmented rhs row: -e) or a dense vector over the six pose columns
// (baseline: column `sub` of A_{id,k}).  Returns 0 when camera k does not depend on ri, 1 for a unit vector, 2 for a dense one.
__device__ __forceinline__ int red_column(const DevProblem& p, const RedIndex& ri, int k, int& unit, double& sign, double dense[6]) {
  if (ri.kind == 0) {
    if (ri.id != k) return 0;
    unit = 6 + ri.sub;
    sign = 1.0;
    return 1;
  }
  if (ri.kind == 1) {
    if (ri.id >= k) return 0;
    const double* A = p.camA + ((size_t)k * p.n_cams + ri.id) * 36;
#pragma unroll
    for (int a = 0; a < 6; ++a) dense[a] = A[a * 6 + ri.sub];
    return 2;
  }
  unit = E_COL;  // rhs = -J^T e
  sign = -1.0;
  return 1;
}

__device__ void camera_block_element(const DevProblem& p, int idx) {
  const int n = p.n_aug;
  if (idx >= n * n) return;
  const int i = idx / n, j = idx % n;
  if (i > j) return;
  const RedIndex ri = red_index(p, i), rj = red_index(p, j);
  double acc = 0.0;
  for (int k = 0; k < p.n_cams; ++k) {
    int ui = 0, uj = 0;
    double si = 1.0, sj = 1.0, di[6], dj[6];
    const int ti = red_column(p, ri, k, ui, si, di);
    if (!ti) continue;
    const int tj = red_column(p, rj, k, uj, sj, dj);
    if (!tj) continue;
    const double* G = p.sumG + (size_t)k * 256;
    double s = 0.0;
    if (ti == 1 && tj == 1) {
      s = si * sj * sumg_sym(G, ui, uj);
    } else if (ti == 1) {
#pragma unroll
      for (int b = 0; b < 6; ++b) s += sumg_sym(G, ui, b) * dj[b];
      s *= si;
    } else if (tj == 1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) s += di[a] * sumg_sym(G, a, uj);
      s *= sj;
    } else {
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) t += sumg_sym(G, a, b) * dj[b];
        s += di[a] * t;
      }
    }
    acc += s;
  }
  p.U[(size_t)i * n + j] = acc;
  p.U[(size_t)j * n + i] = acc;
}

// exchange C consumer: total cost, summed in rank order
// lm_mode 1 (device loop over the peer exchange): the loop boundary runs here, on the combined cost
__global__ void px_combine_cost_kernel(DevProblem p, double* __restrict__ out, int lm_mode, double* __restrict__ trace, int* __restrict__ pos_def) {
  pdl_enter();
  if (p.ctrl->done) return;
  if (p.ctrl->skip_eval) {
    if (lm_mode == 1) lm_boundary(p.ctrl, trace, pos_def, false);
    return;
  }
  const unsigned long long e = px_cur_epoch(p.px, 2);
  double s = 0.0;
  for (int r = 0; r < p.px.n_ranks; ++r) {
    px_wait(p.px, 2, r, e, p.ctrl);
    s += __ldcg(p.px.base[p.px.rank] + px_off_c(p.px, (int)(e & 1), r));
  }
  out[0] = s;
  if (lm_mode == 1 && !p.ctrl->done) {
    p.ctrl->cost_new = s;
    lm_boundary(p.ctrl, trace, pos_def, true);
  }
}
// exchange A consumer: the reduced camera system = the ranks' partials summed in rank order, spread over many CTAs
__global__ void __launch_bounds__(256) px_reduce_system_kernel(DevProblem p) {
  if (p.ctrl->done) return;
  const unsigned long long e = px_cur_epoch(p.px, 0);
  if (threadIdx.x < p.px.n_ranks) px_wait(p.px, 0, threadIdx.x, e, p.ctrl);
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n_aug * p.n_aug) return;
  const double* slots = p.px.base[p.px.rank] + px_off_a(p.px, (int)(e & 1), 0) + idx;
  double v = 0.0;
  for (int r = 0; r < p.px.n_ranks; ++r) v += __ldcg(slots + (size_t)r * p.px.na2);
  p.Sred[idx] = v;
}
// exchange B consumer: rho denominator (sum), max|dx| (max), pos-def (min) over the ranks, in rank order
__global__ void px_combine_solve_kernel(DevProblem p, double* __restrict__ rho_max, int* __restrict__ pos_def, int lm_mode) {
  pdl_enter();
  if (p.ctrl->done) return;
  const unsigned long long e = px_cur_epoch(p.px, 1);
  double rho = 0.0, mx = 0.0;
  int pd = 1;
  for (int r = 0; r < p.px.n_ranks; ++r) {
    px_wait(p.px, 1, r, e, p.ctrl);
    const double* slot = p.px.base[p.px.rank] + px_off_b(p.px, (int)(e & 1), r);
    rho += __ldcg(slot);
    mx = fmax(mx, __ldcg(slot + 1));
    if (__ldcg(slot + 2) < 0.5) pd = 0;
  }
  rho_max[0] = rho;
  rho_max[1] = mx;
  pos_def[0] = pd;
  if (lm_mode == 1 && !p.ctrl->done) kalibr_b200::lm_after_solve(p.ctrl, rho, mx, pd);
}

// =========================================================================================================
// Schur complement on the FP64 tensor pipe:  partial = sum_{v in CTA slice} Z_v Z_v^T,  Z_v = [W_v ; b_v^T] L_v^-T,
// (V_v + d I) = L_v L_v^T.  Every CTA first inverts the 6x6 factors of its own slice of sets (thread per set, -> p.Lv, which the
// back substitution reads too), then turns the rows of [W_v ; b_v^T] into Z rows with a 6x6 triangular product and accumulates
// Z Z^T with DMMA, SC_SETS sets per step.
// =========================================================================================================
// inverse Cholesky factor of V_v + d I of one set -> p.Lv (row-major lower); clears the flag when the block is not positive definite
__device__ __forceinline__ void pose_factor(const DevProblem& p, int set, double damping, int* __restrict__ pos_def_flag) {
  double L[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) L[i] = p.V[(size_t)set * 36 + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) L[i * 6 + i] += damping;
  if (!chol6(L)) *pos_def_flag = 0;
  // Li = L^-1 (lower triangular), column by column
  double Li[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) Li[i] = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    Li[c * 6 + c] = 1.0 / L[c * 6 + c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = c; k < r; ++k) s -= L[r * 6 + k] * Li[k * 6 + c];
      Li[r * 6 + c] = s / L[r * 6 + r];
    }
  }
#pragma unroll
  for (int i = 0; i < 36; ++i) p.Lv[(size_t)set * 36 + i] = Li[i];
}

constexpr int SC_SETS = 4;              // sets per step: K = 24 = 6 k-steps of the m8n8k4 DMMA
constexpr int SC_K = SC_SETS * 6;
constexpr int SC_LD = 36;               // ld % 16 == 4 keeps the operand loads conflict-free

// Per step of SC_SETS sets: the raw rows [W_v ; b_v^T] and the inverse factors L_v^-1 travel with cp.async straight into a Z
// buffer two steps ahead; one step ahead every thread turns the rows it copied into Z rows in place (z = L^-1 w, a 6x6
// triangular product) while the DMMA phase of the current step runs.
// BLOCKED: a warp owns up to SC_UNITS rectangles of up to 4 x 4 tile pairs (SchurUnits, filled by the launcher: the upper triangle of
// the tile grid cut into 4 x 4 macro blocks, largest first onto the least loaded warp), so the four row operands and four column
// operands of a k-step feed up to sixteen DMMAs - the operand loads of the scattered assignment (two per DMMA) kept the LSU as busy as
// the tensor pipe.  !BLOCKED (small systems): MAX_PAIRS scattered tile pairs per warp.
constexpr int SC_UNITS = 2;
constexpr int SC_MAX_SLOTS = 64;  // warps x units x gridDim.y
struct SchurUnits {
  unsigned char ti0[SC_MAX_SLOTS], nr[SC_MAX_SLOTS], tj0[SC_MAX_SLOTS], nc[SC_MAX_SLOTS];  // nr = 0: empty slot; ti0 == tj0: diagonal block (c >= r only)
};

template <int WARPS, int MAX_PAIRS, bool BLOCKED>
__global__ void __launch_bounds__(WARPS * 32, 1) schur_kernel(DevProblem p, double* __restrict__ partials, int sets_per_cta, double damping_arg,
                                                              int* __restrict__ pos_def_flag, const SchurUnits units) {
  pdl_enter();
  extern __shared__ __align__(16) double smem[];
  const int n = p.n_aug;
  const int nt = (n + 7) >> 3;
  const int n_pad = nt * 8;
  double* Zbuf = smem;                         // [3][n_pad][SC_LD]: reduced on / being transformed / in flight
  double* sLi = smem + 3 * n_pad * SC_LD;      // [3][SC_SETS][36]
  if (p.ctrl->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int arow = lane >> 2, acol = lane & 3;
  const int npairs = nt * (nt + 1) / 2;
  // !BLOCKED: tile pairs of this warp (upper triangle, row-major enumeration)
  int ti[BLOCKED ? 1 : MAX_PAIRS], tj[BLOCKED ? 1 : MAX_PAIRS], offA[BLOCKED ? 1 : MAX_PAIRS], offB[BLOCKED ? 1 : MAX_PAIRS];
  double acc[BLOCKED ? 1 : MAX_PAIRS][2];
  // BLOCKED: the warp's rectangles
  int u_ti0[SC_UNITS], u_nr[SC_UNITS], u_tj0[SC_UNITS], u_nc[SC_UNITS];
  double bacc[BLOCKED ? SC_UNITS : 1][4][4][2];
  if constexpr (BLOCKED) {
#pragma unroll
    for (int u = 0; u < SC_UNITS; ++u) {
      const int slot = (blockIdx.y * WARPS + warp) * SC_UNITS + u;
      u_ti0[u] = units.ti0[slot]; u_nr[u] = units.nr[slot]; u_tj0[u] = units.tj0[slot]; u_nc[u] = units.nc[slot];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) bacc[u][r][c][0] = bacc[u][r][c][1] = 0.0;
    }
  } else {
#pragma unroll
    for (int q = 0; q < MAX_PAIRS; ++q) {
      int pair = warp + q * WARPS + blockIdx.y * (WARPS * MAX_PAIRS);  // gridDim.y splits the tile pairs of large systems
      int i = 0;
      if (pair < npairs) {
        int rem = pair;
        while (rem >= nt - i) { rem -= nt - i; ++i; }
        ti[q] = i;
        tj[q] = i + rem;
      } else {
        ti[q] = -1;
        tj[q] = -1;
      }
      offA[q] = (8 * max(ti[q], 0) + arow) * SC_LD + acol;
      offB[q] = (8 * max(tj[q], 0) + arow) * SC_LD + acol;
      acc[q][0] = 0.0;
      acc[q][1] = 0.0;
    }
  }
  for (int i = tid; i < 3 * n_pad * SC_LD; i += blockDim.x) Zbuf[i] = 0.0;
  const int s_lo = blockIdx.x * sets_per_cta, s_hi = min(p.n_sets, s_lo + sets_per_cta);
  {
    // the factors of this slice (the CTAs of a gridDim.y split compute the same values: identical, benign double stores)
    const double damping = damping_arg == damping_arg ? damping_arg : p.ctrl->damping;  // NaN: the device loop's own (a real damping can be negative: Q2 residual)
    for (int set = s_lo + tid; set < s_hi; set += blockDim.x) pose_factor(p, set, damping, pos_def_flag);
  }
  __syncthreads();
  // the rows this thread copies and transforms in every step: o = tid + i * blockDim  ->  (set of the step, row); computed once
  // (the integer divisions by the runtime n were a fifth of the instructions of a step)
  constexpr int SC_MAX_ROWS = (SC_SETS * 256 + WARPS * 32 - 1) / (WARPS * 32);  // n <= 256
  int row_which[SC_MAX_ROWS], row_r[SC_MAX_ROWS];
#pragma unroll
  for (int i = 0; i < SC_MAX_ROWS; ++i) {
    const int o = tid + i * WARPS * 32;
    row_which[i] = o < SC_SETS * n ? o / n : -1;
    row_r[i] = o - max(row_which[i], 0) * n;
  }
  // raw rows of the sets [s0, s0 + SC_SETS) -> Zs, their inverse factors -> Ls  (asynchronous; rows of sets past the end are zeroed)
  auto fetch = [&](double* Zs, double* Ls, int s0) {
#pragma unroll
    for (int i = 0; i < SC_MAX_ROWS; ++i) {
      const int which = row_which[i], r = row_r[i];
      if (which < 0) continue;
      const int set = s0 + which;
      double* dst = Zs + r * SC_LD + which * 6;
      if (set < s_hi) {
        const double* src = (r < p.n_c) ? (p.W + ((size_t)set * p.n_c + r) * 6) : (p.bv + (size_t)set * 6);
        cp_async_16(dst, src);
        cp_async_16(dst + 2, src + 2);
        cp_async_16(dst + 4, src + 4);
      } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[c] = 0.0;
      }
    }
    if (tid < SC_SETS * 18) {
      const int which = tid / 18, q = tid - which * 18;
      if (s0 + which < s_hi) cp_async_16(Ls + which * 36 + 2 * q, p.Lv + (size_t)(s0 + which) * 36 + 2 * q);
    }
    cp_async_commit();
  };
  // in place: z = Linv * w for the rows this thread fetched (same index mapping as fetch)
  auto transform = [&](double* Zs, const double* Ls, int s0) {
#pragma unroll
    for (int i = 0; i < SC_MAX_ROWS; ++i) {
      const int which = row_which[i], r = row_r[i];
      if (which < 0) continue;
      const int set = s0 + which;
      if (set >= s_hi) continue;
      double* zr = Zs + r * SC_LD + which * 6;
      const double* Li = Ls + which * 36;
      double w[6], z[6];
      *reinterpret_cast<double2*>(w) = *reinterpret_cast<const double2*>(zr);
      *reinterpret_cast<double2*>(w + 2) = *reinterpret_cast<const double2*>(zr + 2);
      *reinterpret_cast<double2*>(w + 4) = *reinterpret_cast<const double2*>(zr + 4);
#pragma unroll
      for (int i2 = 0; i2 < 6; ++i2) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k <= i2; ++k) t += Li[i2 * 6 + k] * w[k];
        z[i2] = t;
      }
      *reinterpret_cast<double2*>(zr) = make_double2(z[0], z[1]);
      *reinterpret_cast<double2*>(zr + 2) = make_double2(z[2], z[3]);
      *reinterpret_cast<double2*>(zr + 4) = make_double2(z[4], z[5]);
      if (r == p.n_c && blockIdx.y == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) p.yv[(size_t)set * 6 + c] = z[c];
      }
    }
  };
  // Three buffers, ONE barrier per step: while the tensor pipe reduces step k, the rows of step k + 1 (landed during step k - 1) are
  // turned into Z rows and the raw rows of step k + 2 are in flight.  Half of the warps do the triangular products first and the
  // DMMAs second, the other half the other way round, so the two kinds of FP64 work overlap on the shared pipe.
  auto zb = [&](int k) { return Zbuf + (size_t)(k % 3) * n_pad * SC_LD; };
  auto lb = [&](int k) { return sLi + (k % 3) * SC_SETS * 36; };
  const int n_steps = (s_hi - s_lo + SC_SETS - 1) / SC_SETS;
  if (n_steps > 0) {
    fetch(zb(0), lb(0), s_lo);
    if (n_steps > 1) fetch(zb(1), lb(1), s_lo + SC_SETS);
    cp_async_wait_all();
    __syncthreads();
    transform(zb(0), lb(0), s_lo);
    __syncthreads();
  }
  auto reduce = [&](const double* Zs) {
    if constexpr (BLOCKED) {
#pragma unroll
      for (int u = 0; u < SC_UNITS; ++u) {
        if (u_nr[u] == 0) continue;
        const double* za = Zs + (8 * u_ti0[u] + arow) * SC_LD + acol;
        const double* zb = Zs + (8 * u_tj0[u] + arow) * SC_LD + acol;
        const bool diag = u_ti0[u] == u_tj0[u];
        if (u_nr[u] == 4 && u_nc[u] == 4 && !diag) {  // full rectangle: 8 operand loads, 16 DMMAs per k-step
#pragma unroll
          for (int kk = 0; kk < SC_K / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              a[r] = za[r * 8 * SC_LD + 4 * kk];
              b[r] = zb[r * 8 * SC_LD + 4 * kk];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) dmma(bacc[u][r][c][0], bacc[u][r][c][1], a[r], b[c]);
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < SC_K / 4; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              a[r] = r < u_nr[u] ? za[r * 8 * SC_LD + 4 * kk] : 0.0;
              b[r] = r < u_nc[u] ? zb[r * 8 * SC_LD + 4 * kk] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (r < u_nr[u] && c < u_nc[u] && (!diag || c >= r)) dmma(bacc[u][r][c][0], bacc[u][r][c][1], a[r], b[c]);
          }
        }
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < SC_K / 4; ++kk) {
#pragma unroll
        for (int q = 0; q < MAX_PAIRS; ++q) {
          if (ti[q] >= 0) {  // (an unpredicated loop with dummy pairs in the empty slots was measured: no faster)
            const double a = Zs[offA[q] + 4 * kk];  // operand offsets computed once: the k-step is an immediate of the load
            const double b = Zs[offB[q] + 4 * kk];
            dmma(acc[q][0], acc[q][1], a, b);
          }
        }
      }
    }
  };
  for (int k = 0; k < n_steps; ++k) {
    if (k + 2 < n_steps) fetch(zb(k + 2), lb(k + 2), s_lo + (k + 2) * SC_SETS);  // lands during this step and the next
    if (warp & 1) {
      reduce(zb(k));
      if (k + 1 < n_steps) transform(zb(k + 1), lb(k + 1), s_lo + (k + 1) * SC_SETS);
    } else {
      if (k + 1 < n_steps) transform(zb(k + 1), lb(k + 1), s_lo + (k + 1) * SC_SETS);
      reduce(zb(k));
    }
    cp_async_wait_all();  // this thread's copies of step k + 2
    __syncthreads();      // step k + 1 is transformed, step k + 2 has landed (the L factors are shared), step k's buffer is free
  }
  double* out = partials + (size_t)blockIdx.x * n_pad * n_pad;
  if constexpr (BLOCKED) {
#pragma unroll
    for (int u = 0; u < SC_UNITS; ++u) {
      const bool diag = u_ti0[u] == u_tj0[u];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (r < u_nr[u] && c < u_nc[u] && (!diag || c >= r)) {
            double* o = out + (size_t)(8 * (u_ti0[u] + r) + arow) * n_pad + 8 * (u_tj0[u] + c) + 2 * acol;
            *reinterpret_cast<double2*>(o) = make_double2(bacc[u][r][c][0], bacc[u][r][c][1]);
          }
    }
  } else {
#pragma unroll
    for (int q = 0; q < MAX_PAIRS; ++q) {
      if (ti[q] >= 0) {
        double* o = out + (size_t)(8 * ti[q] + arow) * n_pad + 8 * tj[q] + 2 * acol;
        *reinterpret_cast<double2*>(o) = make_double2(acc[q][0], acc[q][1]);
      }
    }
  }
}

// Sred = U - sum_partials (fixed order), symmetric, undamped; this rank's contribution to the all-reduce.
constexpr int SF_LANES = 16;
__global__ void __launch_bounds__(256) schur_finalize_kernel(DevProblem p, const double* __restrict__ partials, int n_partials) {
  pdl_enter();
  const int n = p.n_aug;
  const int n_pad = ((n + 7) >> 3) * 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  // SF_LANES lanes per element, each a contiguous share of the partials (all of its loads in flight at once); combined in a fixed order
  const int idx = t / SF_LANES, seg = t % SF_LANES;
  const int i = idx / n, j = idx - i * n;
  const bool live = idx < n * n && i <= j && !p.ctrl->done;
  // within a diagonal tile only the mma's own (i,j) entry is used for i<=j, so the result is exactly symmetric
  double s = 0.0;
  if (live) {
    const int per = (n_partials + SF_LANES - 1) / SF_LANES;
    const int c0 = seg * per, c1 = min(n_partials, c0 + per);
    const double* src = partials + (size_t)i * n_pad + j;
#pragma unroll 10
    for (int c = c0; c < c1; ++c) s += src[(size_t)c * n_pad * n_pad];
  }
#pragma unroll
  for (int o = SF_LANES / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (!p.px.enabled) {
    if (live && seg == 0) {
      const double v = p.U[(size_t)i * n + j] - s;
      p.Sred[(size_t)i * n + j] = v;
      p.Sred[(size_t)j * n + i] = v;
    }
    return;
  }
  // exchange A: the partial goes straight into this rank's slot of EVERY rank's exchange buffer (NVLink stores); the last
  // block to finish publishes the epoch.  reduced_solve_kernel sums the slots.
  if (p.ctrl->done) return;
  const unsigned long long e = px_next_epoch(p.px, 0);
  if (live && seg == 0) {
    const double v = p.U[(size_t)i * n + j] - s;
    const size_t off = px_off_a(p.px, (int)(e & 1), p.px.rank);
    for (int r = 0; r < p.px.n_ranks; ++r) {
      double* dst = p.px.base[r] + off;
      dst[(size_t)i * n + j] = v;
      dst[(size_t)j * n + i] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* words = px_words(p.px, p.px.rank);
    __threadfence_system();  // cumulative: orders the block's stores (observed through the barrier) before the counter update
    const unsigned long long prev = atomicAdd(words + 3 * p.px.n_ranks + 3, 1ull);
    if (prev == gridDim.x - 1) {
      words[3 * p.px.n_ranks + 3] = 0ull;
      px_signal(p.px, 0, e);
    }
  }
}

// =========================================================================================================
// reduced solve: one CTA.  Left-looking Cholesky of the augmented reduced system held packed (lower) in shared
// memory; the augmented last row comes out as y = L^-1 b, then L^T x = y.
// =========================================================================================================
constexpr int RS_THREADS = 512;  // 16 warps: one tile of the panel update per warp for n <= 128, and 16 loads in flight per thread while loading
constexpr int RS_NB = 8;  // panel width = DMMA tile
__device__ __forceinline__ int tri(int i, int k) { return i * (i + 1) / 2 + k; }

// Cholesky of the W x W diagonal block at (j0, j0) of the packed matrix, fully unrolled so that the block lives in
// registers; leaves the factor in place and the reciprocals of its diagonal in rd[j0..j0+W).  Returns false if a pivot is
// not positive.
template <int W>
__device__ __forceinline__ bool diag_block(double* __restrict__ Lp, int j0, double* __restrict__ rd_all) {
  double L[W][W], rd[W];
  bool ok = true;
#pragma unroll
  for (int r = 0; r < W; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) L[r][c] = Lp[(j0 + r) * (j0 + r + 1) / 2 + j0 + c];
#pragma unroll
  for (int c = 0; c < W; ++c) {
    double d = L[c][c];
#pragma unroll
    for (int k = 0; k < c; ++k) d -= L[c][k] * L[c][k];
    if (!(d > 0.0)) ok = false;
    d = sqrt(d);
    L[c][c] = d;
    rd[c] = 1.0 / d;
#pragma unroll
    for (int r = c + 1; r < W; ++r) {
      double v = L[r][c];
#pragma unroll
      for (int k = 0; k < c; ++k) v -= L[r][k] * L[c][k];
      L[r][c] = v * rd[c];
    }
  }
#pragma unroll
  for (int r = 0; r < W; ++r) {
    rd_all[j0 + r] = rd[r];
#pragma unroll
    for (int c = 0; c <= r; ++c) Lp[(j0 + r) * (j0 + r + 1) / 2 + j0 + c] = L[r][c];
  }
  return ok;
}
// The same factorisation by ONE WARP (call with all 32 lanes of a warp; lanes 0..7 own the rows): right-looking, so that the
// dependent chain per column is one reciprocal square root, one product, one shuffle and one FMA instead of a serial sweep over the
// block by a single thread (FP64 results take ~20 cycles to come back: the serial version spends ~2.5 us per panel on this GPU).
// Rows / columns past w are padded with the identity.  Returns false if a pivot is not positive (same in every lane).
__device__ __forceinline__ bool diag_block_warp(double* __restrict__ Lp, int j0, int w, double* __restrict__ rd_all, int lane) {
  const int r = lane & 7;  // lanes 8..31 shadow lanes 0..7 (no divergence around the shuffles); only lanes 0..7 store
  __syncwarp();            // converged from here on: a diverged warp would take the slow path of every shuffle below
  double a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = (r < w && c <= r) ? Lp[(j0 + r) * (j0 + r + 1) / 2 + j0 + c] : (c == r ? 1.0 : 0.0);
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double d = __shfl_sync(0xffffffffu, a[c], c);  // the pivot, from the lane that owns row c
    if (c < w && !(d > 0.0)) ok = false;
    const double rdc = rsqrt(d);
    if (r == c) a[c] = d * rdc;          // sqrt(d)
    else if (r > c) a[c] *= rdc;         // L[r][c]
    if (lane == c && c < w) rd_all[j0 + c] = rdc;
    // trailing update of this lane's row: a[j] -= L[r][c] L[j][c] for c < j <= r
#pragma unroll
    for (int j = c + 1; j < 8; ++j) {
      const double ljc = __shfl_sync(0xffffffffu, a[c], j);
      if (j <= r) a[j] = fma(-a[c], ljc, a[j]);
    }
  }
  if (lane < w) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c <= lane) Lp[(j0 + lane) * (j0 + lane + 1) / 2 + j0 + c] = a[c];
  }
  return ok;
}

// rows below the diagonal block: x Ldd^T = a, a forward substitution per row (thread per row) with the block in registers
template <int W>
__device__ __forceinline__ void panel_rows(double* __restrict__ Lp, int j0, int n, const double* __restrict__ rd_all, int tid, int nthreads) {
  if (j0 + W + tid >= n) return;
  double L[W][W], rd[W];
#pragma unroll
  for (int r = 0; r < W; ++r) {
    rd[r] = rd_all[j0 + r];
#pragma unroll
    for (int c = 0; c < r; ++c) L[r][c] = Lp[(j0 + r) * (j0 + r + 1) / 2 + j0 + c];
  }
  for (int row = j0 + W + tid; row < n; row += nthreads) {
    double x[W];
    double* base = Lp + row * (row + 1) / 2 + j0;
#pragma unroll
    for (int c = 0; c < W; ++c) x[c] = base[c];
#pragma unroll
    for (int c = 0; c < W; ++c) {
      double v = x[c];
#pragma unroll
      for (int k = 0; k < c; ++k) v -= x[k] * L[c][k];
      x[c] = v * rd[c];
    }
#pragma unroll
    for (int c = 0; c < W; ++c) base[c] = x[c];
  }
}

// Blocked left-looking Cholesky of the augmented reduced system (n_aug = n_c + 1 rows; the last row carries the rhs and
// comes out as y = L^-1 b), packed lower triangle in shared memory.  Per 8-column panel:
//   1. all warps: A[rows >= j0][panel] -= L[rows][0:j0] L[panel][0:j0]^T with DMMA m8n8k4 (operands straight from the packed rows),
//   2. thread 0: Cholesky of the 8x8 diagonal block (registers),
//   3. thread per row: L[row][panel] = A[row][panel] Ldd^-T by forward substitution.
// Then L^T x = y panel by panel from the last: one thread solves the 8x8 triangle, all threads push the panel's x into the
// rows above (two barriers per panel instead of one per unknown).
__global__ void __launch_bounds__(RS_THREADS, 1) reduced_solve_kernel(DevProblem p, double damping_arg, int* __restrict__ pos_def_flag, int from_peers,
                                                                      long long* __restrict__ dbg /* KB_RS_TRACE: cycles per phase, or null */) {
  pdl_enter();
  __shared__ long long s_dbg[8];
  if (dbg && threadIdx.x < 8) s_dbg[threadIdx.x] = 0;
  long long t_mark = clock64();
  auto lap = [&](int slot) {  // accumulated in shared memory: the trace must not add memory round trips to what it measures
    if (dbg && threadIdx.x == 0) {
      const long long t = clock64();
      s_dbg[slot] += t - t_mark;
      t_mark = t;
    }
  };
  extern __shared__ __align__(16) double Lp[];  // packed lower triangle, n_rows rows (zero padded), then rd[n_rows], x[n_rows]
  __shared__ int s_ok;
  const int n = p.n_aug, nc = p.n_c;
  if (p.ctrl->done) return;
  const double damping = damping_arg == damping_arg ? damping_arg : p.ctrl->damping;  // NaN: the device loop's own (a real damping can be negative: Q2 residual)
  const int n_rows = ((n + RS_NB - 1) / RS_NB) * RS_NB + RS_NB;
  double* s_rd = Lp + n_rows * (n_rows + 1) / 2;
  double* s_x = s_rd + n_rows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < n_rows * (n_rows + 1) / 2; idx += RS_THREADS) Lp[idx] = 0.0;
  __syncthreads();
  // from_peers: the reduced system is the sum of the ranks' partials in the peer-exchange slots (exchange A): wait for the epoch
  // flags, then sum the slots in rank order while loading - no separate reduction launch
  const double* slots = nullptr;
  if (from_peers) {
    const unsigned long long e = px_cur_epoch(p.px, 0);
    if (tid < p.px.n_ranks) px_wait(p.px, 0, tid, e, p.ctrl);
    __syncthreads();
    slots = p.px.base[p.px.rank] + px_off_a(p.px, (int)(e & 1), 0);
  }
  // lower triangle only, RS_LOAD_UNROLL independent loads in flight per thread (a dependent load per iteration exposed the full
  // L2 latency ~190 times at n = 219: a third of the kernel)
  {
    constexpr int U = 8;
    const int n_tri = n * (n + 1) / 2;
    for (int base = tid; base < n_tri; base += RS_THREADS * U) {
      int ii[U], kk[U];
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = base + u * RS_THREADS;
        // row of packed index t: largest i with i (i + 1) / 2 <= t
        int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > t) --i;
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        ii[u] = t < n_tri ? i : -1;
        kk[u] = t - i * (i + 1) / 2;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u] = 0.0;
        if (ii[u] < 0) continue;
        const int idx = ii[u] * n + kk[u];
        if (from_peers) {
          for (int r = 0; r < p.px.n_ranks; ++r) v[u] += __ldcg(slots + (size_t)r * p.px.na2 + idx);
        } else {
          v[u] = p.Sred[idx];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ii[u] < 0) continue;
        Lp[base + u * RS_THREADS] = (ii[u] == kk[u] && ii[u] < nc) ? v[u] + damping : v[u];
      }
    }
  }
  if (tid == 0) s_ok = 1;
  __syncthreads();
  lap(0);
  const int arow = lane >> 2, acol = lane & 3;
  for (int j0 = 0; j0 < nc; j0 += RS_NB) {
    const int w = min(RS_NB, nc - j0);
    // ---- 1. panel update on the tensor pipe ----
    if (j0 > 0) {
      const int first_tile = j0 / RS_NB, n_tiles = (n + RS_NB - 1) / RS_NB;
      const double* Bbase = Lp + tri(j0 + arow, 0);  // B fragment: L[j0 + lane/4][k0 + lane%4]
      for (int t = first_tile + warp; t < n_tiles; t += RS_THREADS / 32) {
        const double* Abase = Lp + tri(RS_NB * t + arow, 0);
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0;
        int k0 = 0;
        for (; k0 + 16 <= j0; k0 += 16) {  // four independent accumulator chains: the tensor pipe's result latency is what this loop waits for
          dmma(c0, c1, Abase[k0 + acol], Bbase[k0 + acol]);
          dmma(e0, e1, Abase[k0 + 4 + acol], Bbase[k0 + 4 + acol]);
          dmma(f0, f1, Abase[k0 + 8 + acol], Bbase[k0 + 8 + acol]);
          dmma(g0, g1, Abase[k0 + 12 + acol], Bbase[k0 + 12 + acol]);
        }
        if (k0 < j0) {  // j0 is a multiple of 8
          dmma(c0, c1, Abase[k0 + acol], Bbase[k0 + acol]);
          dmma(e0, e1, Abase[k0 + 4 + acol], Bbase[k0 + 4 + acol]);
        }
        c0 = (c0 + e0) + (f0 + g0);
        c1 = (c1 + e1) + (f1 + g1);
        const int row = RS_NB * t + arow, col = j0 + 2 * acol;
        if (row < n) {
          if (col <= row && col < j0 + w) Lp[tri(row, col)] -= c0;
          if (col + 1 <= row && col + 1 < j0 + w) Lp[tri(row, col + 1)] -= c1;
        }
      }
      __syncthreads();
    }
    lap(1);
    // ---- 2. diagonal block (one warp, a row per lane) ----
    if (warp == 0) {
      const bool ok = diag_block_warp(Lp, j0, w, s_rd, lane);
      if (!ok && lane == 0) s_ok = 0;
    }
    __syncthreads();
    lap(2);
    // ---- 3. rows below the block ----
    switch (w) {
      case 8: panel_rows<8>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 7: panel_rows<7>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 6: panel_rows<6>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 5: panel_rows<5>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 4: panel_rows<4>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 3: panel_rows<3>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      case 2: panel_rows<2>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
      default: panel_rows<1>(Lp, j0, n, s_rd, tid, RS_THREADS); break;
    }
    __syncthreads();
    lap(3);
  }
  if (tid == 0 && !s_ok) *pos_def_flag = 0;
  // ---- back substitution L^T x = y, y = row nc of the factor; panels from the last to the first ----
  for (int i = tid; i < nc; i += RS_THREADS) s_x[i] = Lp[tri(nc, i)];
  __syncthreads();
  const int last = ((nc - 1) / RS_NB) * RS_NB;
  for (int j0 = last; j0 >= 0; j0 -= RS_NB) {
    const int w = min(RS_NB, nc - j0);
    if (warp == 0) {  // the w x w triangle of the panel: lane c owns x_c and column c of the block; per step one product, one shuffle, one FMA
      __syncwarp();
      const int c = lane & 7;
      double y = c < w ? s_x[j0 + c] : 0.0;
      const double rdc = c < w ? s_rd[j0 + c] : 0.0;
      double col[8];  // L[j0 + k][j0 + c] for k > c
#pragma unroll
      for (int k = 0; k < 8; ++k) col[k] = (k > c && k < w) ? Lp[tri(j0 + k, j0 + c)] : 0.0;
#pragma unroll
      for (int k = 7; k >= 0; --k) {
        const double xk = __shfl_sync(0xffffffffu, y * rdc, k);  // x_k, final once every later unknown has been pushed into y_k
        if (c == k) y = xk;
        else if (c < k) y = fma(-col[k], xk, y);
      }
      if (lane < w) s_x[j0 + lane] = y;
    }
    __syncthreads();
    for (int k = tid; k < j0; k += RS_THREADS) {  // y[k] -= sum_c L[j0 + c][k] x[j0 + c]
      double v = s_x[k];
      for (int c = 0; c < w; ++c) v -= Lp[tri(j0 + c, k)] * s_x[j0 + c];
      s_x[k] = v;
    }
    __syncthreads();
  }
  for (int i = tid; i < nc; i += RS_THREADS) p.dxc[i] = s_x[i];
  lap(4);
  if (dbg && tid == 0)
    for (int i = 0; i < 5; ++i) dbg[i] = s_dbg[i];
}

// =========================================================================================================
// Symmetric eigen-decomposition of the reduced camera system (the calibration block's information matrix with the set poses
// marginalised out).                 ≙ analyzeSVD (Eigen::JacobiSVD of Omega), aslam_incremental_calibration/.../linalg.cpp:409-425
// Two kernels, one CTA each:
//   sym_eig_kernel       Householder tridiagonalisation + implicit-shift QL.  Absolutely accurate (errors ~ eps |S|): all the
//                        column-SCALED solve of every Gauss-Newton iteration needs (its tolerance is 1e-6 n sv_0).
//   jacobi_polish_kernel one-sided (Hestenes) Jacobi sweeps on G = S V started from the QL vectors: restores the RELATIVE accuracy
//                        of the small singular values of the unscaled, graded system (they span 12 decades; the marginal analysis
//                        sums their logarithms).  From the QL start it needs 1-2 rotating sweeps instead of 17-21 from the identity.
// =========================================================================================================
constexpr int EIG_THREADS = 1024;
constexpr int EIG_MAX_N = 256;
__device__ __forceinline__ double eig_block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return warp_sum(red[threadIdx.x & 31]);  // EIG_THREADS / 32 == 32 partials: one butterfly, identical in every warp
}

// status_out[0]: 0 = converged, 1 = a QL iteration did not converge.  sv_out: |eigenvalues| sorted descending; V_out: [n][n] row-major,
// column k = eigenvector of sv_out[k].  A_glob / Z_glob: scratch of n * (n | 1) doubles each, used when the matrices do not fit in
// shared memory (n > 114).
__global__ void __launch_bounds__(EIG_THREADS, 1) sym_eig_kernel(DevProblem p, double* __restrict__ A_glob, double* __restrict__ Z_glob,
                                                                  double* __restrict__ sv_out, double* __restrict__ V_out, int* __restrict__ status_out,
                                                                  int use_smem) {
  extern __shared__ __align__(16) double eig_smem[];
  __shared__ double red[EIG_THREADS / 32];
  __shared__ double v[EIG_MAX_N], w[EIG_MAX_N], d[EIG_MAX_N], e[EIG_MAX_N];
  __shared__ double cs[2][EIG_MAX_N], sn[2][EIG_MAX_N];
  __shared__ int s_m[2], s_lo[2], s_has[2], s_done[2], s_fail[2];  // per buffer: a round only reads what the previous round wrote
  __shared__ int s_perm[EIG_MAX_N];
  const int n = p.n_c, na = p.n_aug;
  const int ld = n | 1;  // odd leading dimension: conflict-free rows and columns
  double* __restrict__ A = use_smem ? eig_smem : A_glob;
  double* __restrict__ Z = use_smem ? eig_smem + (size_t)n * ld : Z_glob;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = EIG_THREADS / 32;
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int r = idx / n, c = idx - r * n;
    A[r * ld + c] = p.Sred[(size_t)r * na + c];
    Z[r * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  if (tid < EIG_MAX_N) e[tid] = 0.0;
  __syncthreads();
  const long long t_start = clock64();
  // ---- Householder tridiagonalisation: H_k zeroes A[k+2.., k]; Z <- Z H_k ----
  for (int k = 0; k + 2 < n; ++k) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < n; i += EIG_THREADS) part += A[i * ld + k] * A[i * ld + k];
    const double norm2 = eig_block_sum(part, red);
    const double x0 = A[(k + 1) * ld + k];
    if (norm2 == 0.0) continue;  // uniform: the column is already reduced (e[k] stays 0 = A[k+1][k])
    const double norm = sqrt(norm2);
    const double alpha = x0 >= 0.0 ? -norm : norm;
    const double vn2 = 2.0 * norm * (norm + fabs(x0));  // |x - alpha e_1|^2 without a second reduction
    const double ivn = rsqrt(vn2);
    for (int i = k + 1 + tid; i < n; i += EIG_THREADS) v[i] = (A[i * ld + k] - (i == k + 1 ? alpha : 0.0)) * ivn;
    __syncthreads();
    // p = A_sub v (a warp per row) and, in the same phase, Z <- Z (I - 2 v v^T) (a warp per row of Z)
    part = 0.0;
    for (int j = k + 1 + warp; j < n; j += NW) {
      double t = 0.0;
      for (int i = k + 1 + lane; i < n; i += 32) t += A[j * ld + i] * v[i];
      t = warp_sum(t);
      if (lane == 0) {
        w[j] = t;
        part += v[j] * t;
      }
    }
    for (int r = warp; r < n; r += NW) {
      double t = 0.0;
      for (int i = k + 1 + lane; i < n; i += 32) t += Z[r * ld + i] * v[i];
      t = 2.0 * warp_sum(t);
      for (int i = k + 1 + lane; i < n; i += 32) Z[r * ld + i] -= t * v[i];
    }
    const double K = eig_block_sum(part, red);
    for (int j = k + 1 + tid; j < n; j += EIG_THREADS) w[j] = 2.0 * (w[j] - K * v[j]);
    if (tid == 0) e[k] = alpha;
    __syncthreads();
    const int m = n - k - 1;
    for (int idx = tid; idx < m * m; idx += EIG_THREADS) {
      const int j = k + 1 + idx / m, i = k + 1 + idx - (idx / m) * m;
      A[j * ld + i] -= v[j] * w[i] + w[j] * v[i];
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += EIG_THREADS) d[i] = A[i * ld + i];
  if (tid == 0) {
    if (n >= 2) e[n - 2] = A[(n - 1) * ld + n - 2];
    s_fail[0] = s_fail[1] = 0;
    s_done[0] = s_done[1] = 0;
    s_has[0] = s_has[1] = 0;
  }
  // transpose Z in place: from now on Z[i * ld + k] = component k of basis vector i, so that the rotation of the vector pair (i, i + 1)
  // is a coalesced, conflict-free sweep over k
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int r = idx / n, c = idx - r * n;
    if (r < c) {
      const double t = Z[r * ld + c];
      Z[r * ld + c] = Z[c * ld + r];
      Z[c * ld + r] = t;
    }
  }
  __syncthreads();
  // ---- implicit-shift QL on (d, e).  Thread 0 walks the rotation chain of one QL step (a dependent chain of one square root and one
  // reciprocal square root per rotation); meanwhile the other warps apply the PREVIOUS step's rotations to the vectors, so the
  // application is off the critical path. ----
  int l = 0, iter = 0;  // thread 0 only
  int buf = 0, have_prev = 0, failed = 0;
  int n_steps = 0, n_rot = 0;  // thread 0: diagnostics
  const long long t_tridiag = clock64();
  for (;;) {
    if (tid == 0) {
      s_has[buf] = 0;
      s_done[buf] = 0;
      s_fail[buf] = 0;
      while (l < n) {
        int m;
        for (m = l; m + 1 < n; ++m)
          if (fabs(e[m]) <= 2.220446049250313e-16 * (fabs(d[m]) + fabs(d[m + 1]))) break;
        if (m == l) {
          ++l;
          iter = 0;
          continue;
        }
        if (iter++ == 60) {
          s_fail[buf] = 1;
          break;
        }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        const double r = sqrt(g * g + 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? r : -r));
        double sr = 1.0, c = 1.0, pp = 0.0;
        double ei = e[m - 1], di = d[m - 1];
        int i;
        bool split = false;
        for (i = m - 1; i >= l; --i) {
          const double en = i > l ? e[i - 1] : 0.0, dn = i > l ? d[i - 1] : 0.0;  // operands of the next rotation, off the chain
          // critical path per rotation: rr -> rsqrt -> c -> r2 -> g -> rr (one reciprocal square root and a handful of dependent
          // FP64 operations); the square root itself, r = rr / sqrt(rr), is off the chain
          const double f = sr * ei, b = c * ei;
          const double rr = fma(g, g, f * f);
          if (rr == 0.0) {
            e[i + 1] = 0.0;
            d[i + 1] -= pp;
            e[m] = 0.0;
            split = true;
            break;
          }
          const double ir = rsqrt(rr);
          const double gd = d[i + 1] - pp, b2 = b + b;
          e[i + 1] = rr * ir;
          sr = f * ir;
          c = g * ir;
          const double r2 = fma(c, b2, (di - gd) * sr);  // (d_i - g) s + 2 c b with c and s entering side by side
          pp = sr * r2;
          d[i + 1] = gd + pp;
          g = fma(c, r2, -b);
          cs[buf][i] = c;
          sn[buf][i] = sr;
          ei = en;
          di = dn;
        }
        ++n_steps;
        n_rot += m - 1 - i;
        s_m[buf] = m;
        s_lo[buf] = i + 1;  // rotations exist for the indices m - 1 .. i + 1
        s_has[buf] = (i + 1 <= m - 1) ? 1 : 0;
        if (!split) {
          d[l] -= pp;
          e[l] = g;
          e[m] = 0.0;
        }
        break;
      }
      if (l >= n) s_done[buf] = 1;
    } else if (have_prev && tid >= 32 && tid - 32 < n) {
      const int k = tid - 32, pb = buf ^ 1;
      const int m = s_m[pb], lo = s_lo[pb];
      double zi1 = Z[m * ld + k];
      double zi = lo <= m - 1 ? Z[(m - 1) * ld + k] : 0.0;
      for (int i = m - 1; i >= lo; --i) {
        const double zn = i > lo ? Z[(i - 1) * ld + k] : 0.0;  // next rotation's operand, requested before this one's store
        const double c = cs[pb][i], sr = sn[pb][i];
        Z[(i + 1) * ld + k] = sr * zi + c * zi1;
        zi1 = c * zi - sr * zi1;
        zi = zn;
      }
      Z[lo * ld + k] = zi1;
    }
    __syncthreads();
    if (s_fail[buf]) {
      failed = 1;
      break;
    }
    have_prev = s_has[buf];
    const int done = s_done[buf];
    buf ^= 1;
    if (done && !have_prev) break;
  }
  if (tid == 0) {
    status_out[0] = failed;
    status_out[1] = (int)((t_tridiag - t_start) >> 10);  // diagnostics (KB_SVD_TRACE): kilo-cycles of the two phases, QL steps, rotations
    status_out[2] = (int)((clock64() - t_tridiag) >> 10);
    status_out[3] = n_steps;
    status_out[4] = n_rot;
  }
  // ---- sorted output: |lambda| descending (ties by index) ----
  if (tid < n) {
    const double me = fabs(d[tid]);
    int pos = 0;
    for (int j = 0; j < n; ++j) {
      const double o = fabs(d[j]);
      pos += (o > me) || (o == me && j < tid);
    }
    s_perm[pos] = tid;
    sv_out[pos] = me;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int r = idx / n, k = idx - r * n;
    V_out[idx] = Z[s_perm[k] * ld + r];
  }
}

// One-sided (Hestenes) Jacobi on G = S V0, started from the eigenvectors V0 of sym_eig_kernel (row-major [n][n], column k = k-th vector;
// V0 == nullptr starts from the identity).  One CTA; the columns of G and V live in shared memory when they fit (n <= 119), else in
// the global scratch buffers (column-major, L1/L2-resident).  A sweep is n - 1 round-robin steps of n / 2 disjoint column pairs; a
// half-warp takes a pair: three dot products, then the plane rotation that makes the two columns orthogonal, applied to G and V.
// Converged when a whole sweep rotates nothing; the column norms are the singular values, the columns of V the singular vectors;
// both are written sorted by descending singular value.
__global__ void __launch_bounds__(EIG_THREADS, 1) jacobi_polish_kernel(DevProblem p, const double* __restrict__ V0, double* __restrict__ G_glob,
                                                                        double* __restrict__ V_glob, double* __restrict__ sv_out, double* __restrict__ V_out,
                                                                        int* __restrict__ sweeps_out, int use_smem) {
  extern __shared__ __align__(16) double eig_smem[];
  __shared__ double s_sigma[EIG_MAX_N];
  __shared__ int s_perm[EIG_MAX_N];
  const int n = p.n_c, na = p.n_aug;
  const int np = (n + 1) & ~1;  // padded to an even number of columns (the extra one is zero and never rotates)
  double* __restrict__ G = use_smem ? eig_smem : G_glob;
  double* __restrict__ V = use_smem ? eig_smem + (size_t)np * n : V_glob;
  // two columns count as orthogonal below n * eps relative to their norms (the usual one-sided Jacobi criterion); anything tighter
  // only adds sweeps that chase rounding noise
  const double ortho_tol = 2.220446049250313e-16 * (double)(n > 8 ? n : 8);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = EIG_THREADS / 32;
  const int half = lane >> 4, hl = lane & 15;
  for (int idx = tid; idx < np * n; idx += EIG_THREADS) {
    const int c = idx / n, r = idx - c * n;
    double g = 0.0, vv = (c == r) ? 1.0 : 0.0;
    if (c < n) {
      if (V0) {
        vv = V0[(size_t)r * n + c];
        for (int j = 0; j < n; ++j) g = fma(p.Sred[(size_t)r * na + j], V0[(size_t)j * n + c], g);
      } else {
        g = p.Sred[(size_t)r * na + c];
      }
    } else {
      vv = 0.0;
    }
    G[idx] = g;
    V[idx] = vv;
  }
  __syncthreads();
  int sweep = 0;
  for (; sweep < 40; ++sweep) {
    int rotated = 0;
    for (int step = 0; step < np - 1; ++step) {
      for (int k0 = 2 * warp; k0 < np / 2; k0 += 2 * n_warps) {
        const int k = k0 + half;
        const bool valid = k < np / 2;
        // round-robin pairing: player np - 1 stays, the others rotate
        int a = 0, b = 0;
        if (valid) {
          if (k == 0) {
            a = np - 1;
            b = step;
          } else {  // (step + k) mod (np - 1), (step - k) mod (np - 1) without integer division: step, k < np - 1
            a = step + k;
            if (a >= np - 1) a -= np - 1;
            b = step - k;
            if (b < 0) b += np - 1;
          }
        }
        double* ga = G + (size_t)a * n;
        double* gb = G + (size_t)b * n;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        if (valid)
          for (int r = hl; r < n; r += 16) {
            const double x = ga[r], y = gb[r];
            alpha += x * x;
            beta += y * y;
            gamma += x * y;
          }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {  // sums inside the half-warp (xor offsets below 16 stay in the half)
          alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
          beta += __shfl_xor_sync(0xffffffffu, beta, o);
          gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
        }
        if (valid && gamma * gamma > ortho_tol * ortho_tol * (alpha * beta) && gamma != 0.0) {
          // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (beta - alpha) / (2 gamma), written with one square root and one
          // division: t = sign(d) g2 / (|d| + sqrt(d^2 + g2^2)), d = beta - alpha, g2 = 2 gamma
          const double dd = beta - alpha, g2 = 2.0 * gamma;
          const double t = copysign(1.0, dd) * g2 / (fabs(dd) + sqrt(dd * dd + g2 * g2));
          const double c = rsqrt(1.0 + t * t), sn = c * t;
          double* va = V + (size_t)a * n;
          double* vb = V + (size_t)b * n;
          for (int r = hl; r < n; r += 16) {
            const double x = ga[r], y = gb[r];
            ga[r] = c * x - sn * y;
            gb[r] = sn * x + c * y;
            const double u = va[r], ww = vb[r];
            va[r] = c * u - sn * ww;
            vb[r] = sn * u + c * ww;
          }
          rotated = 1;
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) break;
  }
  if (tid == 0) sweeps_out[0] = sweep;
  for (int c = warp; c < n; c += n_warps) {
    double s = 0.0;
    for (int r = lane; r < n; r += 32) s += G[(size_t)c * n + r] * G[(size_t)c * n + r];
    s = warp_sum(s);
    if (lane == 0) s_sigma[c] = sqrt(s);
  }
  __syncthreads();
  if (tid < n) {  // rank sort, descending, ties by index
    const double me = s_sigma[tid];
    int pos = 0;
    for (int j = 0; j < n; ++j) pos += (s_sigma[j] > me) || (s_sigma[j] == me && j < tid);
    s_perm[pos] = tid;
    sv_out[pos] = me;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += EIG_THREADS) {
    const int r = idx / n, k = idx - r * n;
    V_out[idx] = V[(size_t)s_perm[k] * n + r];
  }
}

// ---- truncated-SVD solve of the reduced system (the incremental estimator's linear solver) ----------------------------------------
// ≙ aslam::calibration::LinearSolver::solve (IC/src/core/LinearSolver.cpp:299-463, IC = aslam_incremental_calibration/
// incremental_calibration): the pose columns are eliminated by QR, the calibration columns are solved through the SVD of
// Omega = A_r^T A_r - (A_r^T Q)(A_r^T Q)^T — this path's undamped Schur-reduced camera system — cut at the numerical rank.
// diag of the camera block of H = squared column norms of the calibration columns of J
__global__ void __launch_bounds__(256) camera_diag_kernel(DevProblem p, double* __restrict__ out) {
  for (int i = threadIdx.x; i < p.n_c; i += blockDim.x) out[i] = p.U[(size_t)i * p.n_aug + i];
}
// columnScalingMatrix (IC/src/algorithms/linalg.cpp:128-152): g_i = 1 / |column i| (0 below sqrt(rows * eps)); S <- G S G, b <- G b
__global__ void __launch_bounds__(256) svd_scale_kernel(DevProblem p, const double* __restrict__ diag_h, double norm_tol, int enable, double* __restrict__ g) {
  const int n = p.n_c, na = p.n_aug;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double norm = sqrt(diag_h[i]);
    g[i] = enable ? (norm < norm_tol ? 0.0 : 1.0 / norm) : 1.0;
  }
  __syncthreads();
  if (!enable) return;
  for (int idx = threadIdx.x; idx < na * na; idx += blockDim.x) {
    const int r = idx / na, c = idx - r * na;
    p.Sred[idx] *= (r < n ? g[r] : 1.0) * (c < n ? g[c] : 1.0);
  }
}
// rankTol / estimateNumericalRank / svGap / solveSVD (linalg.cpp:244-282, 426-443) on the singular values (descending) and vectors
// of the scaled reduced system, then the un-scaling x_r = G x_r' (LinearSolver.cpp:443-453).  result = {rank, tolerance, gap}
__global__ void __launch_bounds__(256) svd_truncated_solve_kernel(DevProblem p, const double* __restrict__ sv, const double* __restrict__ V,
                                                                  const double* __restrict__ g, double eps_svd, double svd_tol,
                                                                  double* __restrict__ result) {
  __shared__ double s_coef[256];
  __shared__ int s_rank;
  const int n = p.n_c, na = p.n_aug;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double tol = svd_tol != -1.0 ? svd_tol : sv[0] * eps_svd * (double)n;
  if (tid == 0) {
    int rank = n;
    for (int i = n - 1; i > 0; --i) {
      if (sv[i] > tol) break;
      --rank;
    }
    s_rank = rank;
    result[0] = (double)rank;
    result[1] = tol;
    result[2] = rank < n ? sv[rank - 1] / sv[rank] : __longlong_as_double(0x7ff0000000000000ll);
  }
  __syncthreads();
  const int rank = s_rank;
  const double* b = p.Sred + (size_t)n * na;  // right-hand side: last row of the augmented system
  for (int k = warp; k < n; k += 8) {
    double d = 0.0;
    if (k < rank)
      for (int r = lane; r < n; r += 32) d += V[(size_t)r * n + k] * b[r];
    d = warp_sum(d);
    if (lane == 0) s_coef[k] = k < rank ? d / sv[k] : 0.0;
  }
  __syncthreads();
  for (int r = tid; r < n; r += blockDim.x) {
    double x = 0.0;
    for (int k = 0; k < rank; ++k) x += V[(size_t)r * n + k] * s_coef[k];
    p.dxc[r] = g[r] * x;
  }
}

constexpr int RHO_BLOCKS = 64;
__device__ __forceinline__ void block_sum_max(double& s, double& m, double* sh_s, double* sh_m) {
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_m[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? sh_s[threadIdx.x] : 0.0;
    m = threadIdx.x < nw ? sh_m[threadIdx.x] : 0.0;
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
}
// Second stage of the solve scalars, run by ONE block of 256 threads (the last block of backsub_kernel to finish): out[0] = the sum of the
// blocks' partials of dx^T (lambda dx + rhs), out[1] = max |dx|, in a fixed order; producer of peer exchange B; lm_mode 1: thread 0
// runs the after-solve transition of the device-resident loop (single rank).
constexpr int RHO2_THREADS = 256;
__device__ __forceinline__ void solve_scalars_stage2(LmCtrl* __restrict__ ctrl, const PeerXchg& px, const int* __restrict__ pos_def,
                                                     const double* __restrict__ partial, int n, double* __restrict__ out, int lm_mode) {
  __shared__ double sh_s[32], sh_m[32];
  double s = 0.0, m = 0.0;
  // fixed assignment of partials to threads, fixed order inside a thread; four independent 16-byte loads in flight per thread
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * RHO2_THREADS) {
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * RHO2_THREADS;
      v[u] = i < n ? __ldcg(reinterpret_cast<const double2*>(partial) + i) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += v[u].x;
      m = fmax(m, v[u].y);
    }
  }
  block_sum_max(s, m, sh_s, sh_m);
  if (threadIdx.x == 0) {
    out[0] = s;
    out[1] = m;
    if (lm_mode == 1) kalibr_b200::lm_after_solve(ctrl, s, m, pos_def[0]);
    if (px.enabled) {  // exchange B: (rho partial, max|dx|, pos-def) into every rank's slot
      const unsigned long long e = px_next_epoch(px, 1);
      const double pd = (double)pos_def[0];
      for (int r = 0; r < px.n_ranks; ++r) {
        double* slot = px.base[r] + px_off_b(px, (int)(e & 1), px.rank);
        slot[0] = s;
        slot[1] = m;
        slot[2] = pd;
      }
      __threadfence_system();
      px_signal(px, 1, e);
    }
  }
}

// =========================================================================================================
// back substitution for the poses + scatter of dx into design-variable order
// =========================================================================================================
// partial (may be null): [gridDim.x][2] = per block (sum over its sets of dx_v^T (lambda dx_v + b_v), max |dx_v|); block 0 adds the
// camera-side part (dx_c^T b_c of this rank, + lambda |dx_c|^2 when include_shared) - stage 1 of the rho denominator for free
constexpr int BS_SETS = 1;  // sets per warp (4 was measured: slower, 0.046 -> 0.052 ms at 20 000 sets)
__global__ void __launch_bounds__(256) backsub_kernel(DevProblem p, const int* __restrict__ set_col_q, const int* __restrict__ set_col_t,
                                                      const int* __restrict__ cam_cols, double lambda_arg, int include_shared,
                                                      double* __restrict__ partial, double* __restrict__ out2, PeerXchg px,
                                                      const int* __restrict__ pos_def, int lm_mode) {
  pdl_enter();
  __shared__ double sh_s[8], sh_m[8];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p.ctrl->done) return;
  const double lambda = lambda_arg >= 0.0 ? lambda_arg : p.ctrl->lambda;
  double ps = 0.0, pm = 0.0;  // this warp's contribution (lane 0)
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < p.n_c; i += blockDim.x) {
      const double d = p.dxc[i];
      p.dx[cam_cols[i]] = d;
      ps += d * p.U[(size_t)i * p.n_aug + p.n_c];  // this rank's partial b_c
      if (include_shared) ps += lambda * d * d;
      pm = fmax(pm, fabs(d));
    }
    ps = warp_sum(ps);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pm = fmax(pm, __shfl_xor_sync(0xffffffffu, pm, o));
  }
  // BS_SETS consecutive sets per warp: fewer, longer-lived blocks (each block ends with a fence + ticket)
  for (int j = 0; j < BS_SETS; ++j) {
  const int set = gw * BS_SETS + j;
  if (set >= p.n_sets) break;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const double* W = p.W + (size_t)set * p.n_c * 6;
  for (int i = lane; i < p.n_c; i += 32) {
    const double x = p.dxc[i];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] += W[i * 6 + c] * x;
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
    double r[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) r[c] = p.bv[(size_t)set * 6 + c] - acc[c];
    // (V + dI)^-1 r = Linv^T (Linv r)
    const double* Li = p.Lv + (size_t)set * 36;
    double t[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k <= i; ++k) s += Li[i * 6 + k] * r[k];
      t[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = i; k < 6; ++k) s += Li[k * 6 + i] * t[k];
      r[i] = s;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.dx[set_col_q[set] + c] = r[c];
      p.dx[set_col_t[set] + c] = r[3 + c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      ps += r[c] * (lambda * r[c] + p.bv[(size_t)set * 6 + c]);
      pm = fmax(pm, fabs(r[c]));
    }
  }
    }
  if (!partial) return;
  if (lane == 0) { sh_s[wib] = ps; sh_m[wib] = pm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, m = 0.0;
    for (int w = 0; w < 8; ++w) { s += sh_s[w]; m = fmax(m, sh_m[w]); }  // fixed order
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = m;
    // the last block to arrive adds the partials up (fixed order: the result does not depend on which block that is)
    __threadfence();
    const unsigned int ticket = atomicAdd(p.tickets + 1, 1u);
    s_last = ticket == gridDim.x - 1;
    if (s_last) p.tickets[1] = 0u;
  }
  __syncthreads();
  if (!s_last || !out2) return;
  __threadfence();
  solve_scalars_stage2(p.ctrl, px, pos_def, partial, (int)gridDim.x, out2, lm_mode);
}

// out[0] = sum_local dx (lambda dx + rhs) [+ lambda |dx_c|^2 once], out[1] = max |dx| over local poses and the shared block.
// Two stages with fixed order: per-block partials, then one block.
__global__ void __launch_bounds__(256) rho_stage1_kernel(DevProblem p, double lambda_arg, const int* __restrict__ set_col_q, const int* __restrict__ set_col_t,
                                                           int include_shared, double* __restrict__ partial /*[RHO_BLOCKS][2]*/) {
  __shared__ double sh_s[32], sh_m[32];
  if (p.ctrl->done) return;
  const double lambda = lambda_arg >= 0.0 ? lambda_arg : p.ctrl->lambda;
  double s = 0.0, m = 0.0;
  for (int set = blockIdx.x * blockDim.x + threadIdx.x; set < p.n_sets; set += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double d = p.dx[(c < 3 ? set_col_q[set] : set_col_t[set] - 3) + c];
      s += d * (lambda * d + p.bv[(size_t)set * 6 + c]);
      m = fmax(m, fabs(d));
    }
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < p.n_c; i += blockDim.x) {
      const double d = p.dxc[i];
      s += d * p.U[(size_t)i * p.n_aug + p.n_c];  // this rank's partial b_c
      if (include_shared) s += lambda * d * d;
      m = fmax(m, fabs(d));
    }
  }
  block_sum_max(s, m, sh_s, sh_m);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s; partial[2 * blockIdx.x + 1] = m; }
}
// second stage on its own (the rho query outside a solve: kb_lm_rho_denominator)
__global__ void __launch_bounds__(RHO2_THREADS) rho_stage2_kernel(LmCtrl* __restrict__ ctrl, PeerXchg px, const int* __restrict__ pos_def,
                                                                  const double* __restrict__ partial, int n, double* __restrict__ out, int lm_mode) {
  if (ctrl->done) return;
  solve_scalars_stage2(ctrl, px, pos_def, partial, n, out, lm_mode);
}
// Multi-rank: every rank drops (rho partial, max|dx|, pos-def flag) into its own slot of a zeroed [n_ranks][4] array; ONE
// sum all-reduce then hands every rank all slots, and the host combines them in rank order (sum / max / min).
__global__ void pack_rank_scalars_kernel(double* __restrict__ pk, int rank, int n_ranks, const double* __restrict__ rho_max, const int* __restrict__ pos_def) {
  const int i = threadIdx.x;
  if (i >= 4 * n_ranks) return;
  double v = 0.0;
  if ((i >> 2) == rank) v = (i & 3) == 0 ? rho_max[0] : (i & 3) == 1 ? rho_max[1] : (i & 3) == 2 ? (double)pos_def[0] : 0.0;
  pk[i] = v;
}

// sm::kinematics::updateQuat (quaternion_algebra.cpp:200-219, 302-317)
__device__ __forceinline__ void update_quat(double* q, const double* dq) {
  const double theta = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
  double na;
  if (theta < 1.220703125e-4 /* eps^(1/4) = 2^-13 */) na = 0.5 + (theta * theta) * (1.0 / 48.0);
  else na = sin(theta * 0.5) / theta;
  const double d0 = dq[0] * na, d1 = dq[1] * na, d2 = dq[2] * na, ca = cos(theta * 0.5);
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  q[0] = q0 * ca + d0 * q3 - d1 * q2 + d2 * q1;
  q[1] = q1 * ca + d0 * q2 + d1 * q3 - d2 * q0;
  q[2] = q2 * ca - d0 * q1 + d1 * q0 + d2 * q3;
  q[3] = q3 * ca - d0 * q0 - d1 * q1 - d2 * q2;
}

// backup + update of every design variable (Optimizer2.cpp:290-307; RotationQuaternion.cpp:22-36; EuclideanPoint.cpp:23-32;
// DesignVariableAdapter.hpp:42-55 over Projection::update / Distortion::update, all of which are parameter += delta)
__global__ void __launch_bounds__(256) apply_update_kernel(DevProblem p, const int* __restrict__ set_col_q, const int* __restrict__ set_col_t,
                                                           double* __restrict__ backup_cam, double* __restrict__ backup_base, double* __restrict__ backup_sets) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (p.ctrl->done || p.ctrl->skip_eval) return;
  // device-resident loop: a rejected step is undone lazily, here (the backup IS the state to start from) - outside the loop the flag is 0
  const bool restore = p.ctrl->revert != 0;
  if (idx < p.n_sets && blockIdx.x != gridDim.x - 1) {
    double* pose = p.set_poses + (size_t)idx * POSE_STRIDE;
    double* bk = backup_sets + (size_t)idx * POSE_STRIDE;
    const int cq = set_col_q[idx], ct = set_col_t[idx];
    double q[POSE_STRIDE], dq[3], dt[3];
#pragma unroll
    for (int i = 0; i < POSE_STRIDE; ++i) q[i] = restore ? bk[i] : pose[i];  // all loads first: nothing below aliases them
#pragma unroll
    for (int c = 0; c < 3; ++c) { dq[c] = p.dx[cq + c]; dt[c] = p.dx[ct + c]; }
    if (!restore) {
#pragma unroll
      for (int i = 0; i < POSE_STRIDE; ++i) bk[i] = q[i];
    }
    update_quat(q, dq);
#pragma unroll
    for (int c = 0; c < 3; ++c) q[4 + c] += dt[c];
#pragma unroll
    for (int i = 0; i < POSE_STRIDE; ++i) pose[i] = q[i];
  }
  if (blockIdx.x == gridDim.x - 1) {  // the camera-side design variables: a block of their own, a thread per parameter / per baseline
    for (int i = threadIdx.x; i < p.n_cams * CAM_PARAM_STRIDE; i += blockDim.x) {
      const int k = i / CAM_PARAM_STRIDE, c = i - k * CAM_PARAM_STRIDE;
      const double v = restore ? backup_cam[i] : p.cam_params[i];
      backup_cam[i] = v;
      p.cam_params[i] = c < p.cam_P[k] + p.cam_D[k] ? v + p.dxc[p.intr_off[k] + c] : v;
    }
    for (int j = threadIdx.x; j < p.n_cams - 1; j += blockDim.x) {
      double* b = p.baselines + j * POSE_STRIDE;
      double q[POSE_STRIDE], dq[3], dt[3];
#pragma unroll
      for (int i = 0; i < POSE_STRIDE; ++i) q[i] = restore ? backup_base[j * POSE_STRIDE + i] : b[i];
      const int off = p.base_off[j];
#pragma unroll
      for (int c = 0; c < 3; ++c) { dq[c] = p.dxc[off + c]; dt[c] = p.dxc[off + 3 + c]; }
#pragma unroll
      for (int i = 0; i < POSE_STRIDE; ++i) backup_base[j * POSE_STRIDE + i] = q[i];
      update_quat(q, dq);
#pragma unroll
      for (int c = 0; c < 3; ++c) q[4 + c] += dt[c];
#pragma unroll
      for (int i = 0; i < POSE_STRIDE; ++i) b[i] = q[i];
    }
  }
}

// =========================================================================================================
// device-resident Levenberg-Marquardt loop: three single-thread control kernels per iteration carry the scalar logic of
// Optimizer2::optimize (BE/src/Optimizer2.cpp:215-266) and LevenbergMarquardtTrustRegionPolicy::solveSystemImplementation
// (BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:50-113), so that no host round trip sits inside an iteration.
// =========================================================================================================
// after the solve: combine the ranks' (rho, max|dx|, pos-def) slots, the lambda^2 / lambda residual (Q2), failed solves
__global__ void lm_post_solve_kernel(LmCtrl* c, const int* pos_def, const double* rho_max, const double* rank_slots, int n_ranks) {
  if (c->done) return;
  double rho = rho_max[0], mx = rho_max[1];
  int pd = pos_def[0];
  if (n_ranks > 1) {
    rho = 0.0;
    mx = 0.0;
    pd = 1;
    for (int r = 0; r < n_ranks; ++r) {
      rho += rank_slots[4 * r];
      mx = fmax(mx, rank_slots[4 * r + 1]);
      if (rank_slots[4 * r + 2] < 0.5) pd = 0;
    }
  }
  kalibr_b200::lm_after_solve(c, rho, mx, pd);
}

// NCCL path only (the single-rank and peer-exchange paths run the boundary inside finalize_gram / px_combine_cost): after the
// all-reduce of the cost
__global__ void lm_boundary_kernel(LmCtrl* c, double* trace, int* pos_def) {
  if (c->done) return;
  lm_boundary(c, trace, pos_def, !c->skip_eval);
}

// restore the backup when the step was rejected (idempotent: a finished loop may run it again with the same flag)
__global__ void __launch_bounds__(256) lm_revert_kernel(DevProblem p, const double* __restrict__ bk_cam, const double* __restrict__ bk_base,
                                                        const double* __restrict__ bk_sets) {
  if (!p.ctrl->revert) return;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < p.n_sets * POSE_STRIDE) p.set_poses[idx] = bk_sets[idx];
  if (idx < p.n_cams * CAM_PARAM_STRIDE) p.cam_params[idx] = bk_cam[idx];
  if (idx < (p.n_cams - 1) * POSE_STRIDE) p.baselines[idx] = bk_base[idx];
}

// back to the neutral flags of the call-by-call API
__global__ void lm_finish_kernel(LmCtrl* c) {
  c->done = 0;
  c->need_build = 1;
  c->skip_eval = 0;
  c->revert = 0;
}

// =========================================================================================================
// launchers
// =========================================================================================================
// Per-model launches of a mixed rig: the first model with work stays on the main stream, the others fork onto side streams and
// join back, so that launches which each fill only part of the GPU overlap.  `launch(m, ctx)` launches model m on ctx.stream.
template <typename F>
static cudaError_t for_each_model_concurrently(const int* begin /*[NUM_MODELS+1]*/, StreamCtx& s, F launch) {
  int n_models = 0;
  for (int m = 0; m < NUM_MODELS; ++m) n_models += begin[m + 1] > begin[m];
  const bool fork = n_models > 1 && s.side && s.ev_fork && s.ev_join;
  cudaError_t e;
  if (fork && (e = cudaEventRecord(s.ev_fork, s.stream)) != cudaSuccess) return e;
  bool first = true;
  for (int m = 0; m < NUM_MODELS; ++m) {
    if (begin[m + 1] <= begin[m]) continue;
    if (!fork || first) {
      if ((e = launch(m, s)) != cudaSuccess) return e;
    } else {
      StreamCtx side = s;
      side.stream = s.side[m];
      if ((e = cudaStreamWaitEvent(side.stream, s.ev_fork, 0)) != cudaSuccess) return e;
      if ((e = launch(m, side)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(s.ev_join[m], side.stream)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(s.stream, s.ev_join[m], 0)) != cudaSuccess) return e;
    }
    first = false;
  }
  return cudaSuccess;
}

static_assert(NUM_MODELS == 7, "the per-model switch statements below list models 0..6");
// launch-time caches are kept per device, so that several handles on different GPUs can live in one process
constexpr int MAX_DEVICES = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & (MAX_DEVICES - 1);
}
static int sm_count() {
  static std::atomic<int> count[MAX_DEVICES] = {};
  const int dev = cur_device();
  int c = count[dev].load(std::memory_order_relaxed);
  if (!c) {
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
    if (c <= 0) c = 148;
    count[dev].store(c, std::memory_order_relaxed);
  }
  return c;
}
#define KB_LAUNCHED(s) (++*(s).launches)

// cudaFuncAttributeMaxDynamicSharedMemorySize is raised once per (kernel, device); the cache is guarded so that handles driven
// from different threads can launch concurrently (the attribute itself is only ever raised)
static std::mutex g_attr_mutex;
template <typename K>
static cudaError_t ensure_dynamic_smem(K kernel, size_t smem, size_t (&cache)[MAX_DEVICES]) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  size_t& have = cache[cur_device()];
  if (smem > have) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    have = smem;
  }
  return cudaSuccess;
}

cudaError_t launch_prep(const DevProblem& p, StreamCtx& s) {
  prep_kernel<<<1, 32, 0, s.stream>>>(p);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_evaluate_model(const DevProblem& p, const int* list, int n, StreamCtx& s) {
  if (n <= 0) return cudaSuccess;
  const int warps_per_cta = EVAL_THREADS / 32;
  const int grid = min((n + warps_per_cta - 1) / warps_per_cta, sm_count() * 8);
  if (p.weighted) evaluate_kernel<MODEL, true><<<grid, EVAL_THREADS, sizeof(double) * p.n_target * 3, s.stream>>>(p, list, n);
  else evaluate_kernel<MODEL, false><<<grid, EVAL_THREADS, sizeof(double) * p.n_target * 3, s.stream>>>(p, list, n);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_evaluate(const DevProblem& p, const int* view_list, const int* mb, double* cost_out, StreamCtx& s) {
  cudaError_t e = for_each_model_concurrently(mb, s, [&](int m, StreamCtx& c) -> cudaError_t {
    switch (m) {
      case 0: return launch_evaluate_model<0>(p, view_list + mb[0], mb[1] - mb[0], c);
      case 1: return launch_evaluate_model<1>(p, view_list + mb[1], mb[2] - mb[1], c);
      case 2: return launch_evaluate_model<2>(p, view_list + mb[2], mb[3] - mb[2], c);
      case 3: return launch_evaluate_model<3>(p, view_list + mb[3], mb[4] - mb[3], c);
      case 4: return launch_evaluate_model<4>(p, view_list + mb[4], mb[5] - mb[4], c);
      case 5: return launch_evaluate_model<5>(p, view_list + mb[5], mb[6] - mb[5], c);
      case 6: return launch_evaluate_model<6>(p, view_list + mb[6], mb[7] - mb[6], c);
    }
    return cudaSuccess;
  });
  if (e != cudaSuccess) return e;
  sum_kernel<<<1, 1024, 0, s.stream>>>(p.view_cost, p.n_views, cost_out);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_reproj_stats(const DevProblem& p, const double* e_raw, const int* cam_view_list, const int* cam_view_begin, int pass, double* acc,
                                StreamCtx& s) {
  reproj_stats_kernel<<<p.n_cams, 1024, 0, s.stream>>>(p, e_raw, cam_view_list, cam_view_begin, pass, acc);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

int la_grid_warps() { return sm_count() * 2 * LA_WARPS; }

template <int MODEL, bool WRITE_E, bool WEIGHTED>
static cudaError_t launch_la_model(const DevProblem& p, const int4* vmeta, const int4* slices, int lo, int hi, StreamCtx& s) {
  if (hi <= lo) return cudaSuccess;
  static size_t attr_smem_dev[MAX_DEVICES] = {};
  const size_t smem = sizeof(double) * (3 * ((p.n_target + 1) & ~1) + LA_WARPS * LA_WARP_DOUBLES);
  if (cudaError_t e = ensure_dynamic_smem(linearise_assemble_kernel<MODEL, WRITE_E, WEIGHTED>, smem, attr_smem_dev); e != cudaSuccess) return e;
  const int grid = min((hi - lo + LA_WARPS - 1) / LA_WARPS, sm_count() * 2);
  if (cudaError_t e = launch_pdl(linearise_assemble_kernel<MODEL, WRITE_E, WEIGHTED>, grid, LA_THREADS, smem, s.stream, p, vmeta, slices, lo, hi); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

// slice_model_begin[m] .. [m+1]: slices of camera model m
cudaError_t launch_linearise_assemble(const DevProblem& p, const int4* vmeta, const int4* slices, const int* smb, bool write_e, bool with_set_prep,
                                      bool with_cam_prep, StreamCtx& s) {
  cudaError_t e;
  if ((with_set_prep && p.n_sets > 0) || with_cam_prep) {  // per-set constants and (extra block) per-camera constants in one launch
    if ((e = launch_pdl(set_prep_kernel, (with_set_prep ? (p.n_sets + SET_PREP_THREADS - 1) / SET_PREP_THREADS : 0) + (with_cam_prep ? 1 : 0), SET_PREP_THREADS, 0, s.stream, p,
                        with_cam_prep ? 1 : 0)) != cudaSuccess) return e;
    KB_LAUNCHED(s);
  }
  e = for_each_model_concurrently(smb, s, [&](int m, StreamCtx& c) -> cudaError_t {
    switch (m) {
#define KB_LA(M)                                                                                                                      \
  case M:                                                                                                                             \
    if (p.weighted)                                                                                                                   \
      return write_e ? launch_la_model<M, true, true>(p, vmeta, slices, smb[M], smb[M + 1], c)                                       \
                     : launch_la_model<M, false, true>(p, vmeta, slices, smb[M], smb[M + 1], c);                                     \
    return write_e ? launch_la_model<M, true, false>(p, vmeta, slices, smb[M], smb[M + 1], c)                                        \
                   : launch_la_model<M, false, false>(p, vmeta, slices, smb[M], smb[M + 1], c);
      KB_LA(0) KB_LA(1) KB_LA(2) KB_LA(3) KB_LA(4) KB_LA(5) KB_LA(6)
#undef KB_LA
    }
    return cudaSuccess;
  });
  return e != cudaSuccess ? e : cudaGetLastError();
}

// per-camera Gram sums + cost of the linearisation point (-> cost_out[0]); lm_mode 1: + the loop boundary of the device-resident loop
cudaError_t launch_finalize_gram(const DevProblem& p, const int* cam_slice_range, int n_ranges, double* cost_out, bool exchange_cost, int lm_mode,
                                 double* trace, int* pos_def, StreamCtx& s) {
  if (cudaError_t e = launch_pdl(finalize_gram_kernel, p.n_cams, GRAM_TILES * FG_GROUPS, 0, s.stream, p, cam_slice_range, n_ranges, cost_out, exchange_cost ? 1 : 0, lm_mode,
                                 trace, pos_def);
      e != cudaSuccess)
    return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

template <int MODEL, bool WEIGHTED>
static cudaError_t launch_lm_model(const DevProblem& p, const int4* vmeta, const int4* slices, int lo, int hi, double* jt, int bfrag_pairs,
                                   unsigned int* counter, StreamCtx& s) {
  if (hi <= lo) return cudaSuccess;
  const size_t smem = sizeof(double) * (((p.n_target * 3 + 1) & ~1) + (size_t)LM_WARPS * (XT_WARP_DOUBLES + bfrag_pairs * 32 + 36 + SETPREP_STRIDE));
  static size_t attr_smem_dev[MAX_DEVICES] = {};
  if (cudaError_t e = ensure_dynamic_smem(linearise_materialise_kernel<MODEL, WEIGHTED>, smem, attr_smem_dev); e != cudaSuccess) return e;
  const int ctas_per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / (smem + 1024)));
  const int grid = min(((hi - lo) * LM_SUB + LM_WARPS - 1) / LM_WARPS, sm_count() * ctas_per_sm);
  linearise_materialise_kernel<MODEL, WEIGHTED><<<grid, LM_THREADS, smem, s.stream>>>(p, vmeta, slices, lo, hi, jt, bfrag_pairs, counter);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

// slice_model_begin[m] .. [m+1]: slices of camera model m; bfrag_pairs[m]: B-fragment slots a camera of model m needs at most;
// counters: KB_NUM_MODELS work counters (zeroed here)
cudaError_t launch_linearise_materialise(const DevProblem& p, const int4* vmeta, const int4* slices, const int* smb, const int* bfrag_pairs,
                                         unsigned int* counters, double* jt, StreamCtx& s) {
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * NUM_MODELS, s.stream);
  if (e != cudaSuccess) return e;
  set_prep_kernel<<<(p.n_sets + SET_PREP_THREADS - 1) / SET_PREP_THREADS + 1, SET_PREP_THREADS, 0, s.stream>>>(p, 1);  // per-set and (last block) per-camera constants
  KB_LAUNCHED(s);
  e = for_each_model_concurrently(smb, s, [&](int m, StreamCtx& c) -> cudaError_t {
    switch (m) {
#define KB_LM(M)                                                                                                          \
  case M:                                                                                                                 \
    return p.weighted ? launch_lm_model<M, true>(p, vmeta, slices, smb[M], smb[M + 1], jt, bfrag_pairs[M], counters + M, c) \
                      : launch_lm_model<M, false>(p, vmeta, slices, smb[M], smb[M + 1], jt, bfrag_pairs[M], counters + M, c);
      KB_LM(0) KB_LM(1) KB_LM(2) KB_LM(3) KB_LM(4) KB_LM(5) KB_LM(6)
#undef KB_LM
    }
    return cudaSuccess;
  });
  return e != cudaSuccess ? e : cudaGetLastError();
}

// V_v, b_v, W_v from the view blocks (first blocks of the grid); U, b_c from the per-camera Gram sums (remaining blocks): one launch
cudaError_t launch_set_reduce(const DevProblem& p, StreamCtx& s) {
  const int set_blocks = p.n_sets > 0 ? min((p.n_sets + SR_WARPS - 1) / SR_WARPS, sm_count() * 8) : 0;
  const int n2 = p.n_aug * p.n_aug;
  if (cudaError_t e = launch_pdl(set_reduce_kernel, set_blocks + (n2 + SR_WARPS * 32 - 1) / (SR_WARPS * 32), SR_WARPS * 32, 0, s.stream, p, set_blocks); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

static int schur_sets_per_cta(const DevProblem& p) {
  const int ctas = sm_count();
  int per = (p.n_sets + ctas - 1) / ctas;
  per = ((per + SC_SETS - 1) / SC_SETS) * SC_SETS;
  return per < SC_SETS ? SC_SETS : per;
}
int schur_num_partials(const DevProblem& p) {
  const int per = schur_sets_per_cta(p);
  const int n = (p.n_sets + per - 1) / per;
  return n < 1 ? 1 : n;
}
size_t schur_partial_stride(const DevProblem& p) {
  const size_t n_pad = ((p.n_aug + 7) >> 3) * 8;
  return n_pad * n_pad;
}

template <int WARPS, int MAX_PAIRS, bool BLOCKED>
static cudaError_t launch_schur_t(const DevProblem& p, double damping, double* partials, int n_partials, int* flag, StreamCtx& s) {
  const int n_pad = ((p.n_aug + 7) >> 3) * 8;
  const size_t smem = sizeof(double) * (3 * (size_t)n_pad * SC_LD + 3 * SC_SETS * 36);
  static size_t attr_smem_dev[MAX_DEVICES] = {};
  if (cudaError_t e = ensure_dynamic_smem(schur_kernel<WARPS, MAX_PAIRS, BLOCKED>, smem, attr_smem_dev); e != cudaSuccess) return e;
  const int nt = (p.n_aug + 7) >> 3;
  SchurUnits units;
  std::memset(&units, 0, sizeof(units));
  int gy;
  if (BLOCKED) {
    // the upper triangle of the nt x nt tile grid in 4 x 4 macro blocks, largest first onto the least loaded warp with a free slot
    // (computed once per system size and thread)
    thread_local int cached_nt = -1, cached_gy = 0;
    thread_local SchurUnits cached;
    if (cached_nt != nt) {
      struct Blk { int ti0, nr, tj0, nc, pairs; };
      std::vector<Blk> blocks;
      const int ng = (nt + 3) / 4;
      for (int g = 0; g < ng; ++g)
        for (int h2 = g; h2 < ng; ++h2) {
          const int nr = std::min(4, nt - 4 * g), nc = std::min(4, nt - 4 * h2);
          blocks.push_back({4 * g, nr, 4 * h2, nc, g == h2 ? nr * (nr + 1) / 2 : nr * nc});
        }
      std::stable_sort(blocks.begin(), blocks.end(), [](const Blk& a, const Blk& b) { return a.pairs > b.pairs; });
      const int gy_new = ((int)blocks.size() + WARPS * SC_UNITS - 1) / (WARPS * SC_UNITS);
      if (gy_new * WARPS * SC_UNITS > SC_MAX_SLOTS) return cudaErrorInvalidValue;
      const int nw = gy_new * WARPS;
      std::vector<int> load(nw, 0), used(nw, 0);
      SchurUnits u;
      std::memset(&u, 0, sizeof(u));
      for (const Blk& b : blocks) {
        int best = -1;
        for (int w = 0; w < nw; ++w)
          if (used[w] < SC_UNITS && (best < 0 || load[w] < load[best])) best = w;
        const int slot = best * SC_UNITS + used[best]++;
        load[best] += b.pairs;
        u.ti0[slot] = (unsigned char)b.ti0; u.nr[slot] = (unsigned char)b.nr; u.tj0[slot] = (unsigned char)b.tj0; u.nc[slot] = (unsigned char)b.nc;
      }
      cached = u;
      cached_gy = gy_new;
      cached_nt = nt;
    }
    units = cached;
    gy = cached_gy;
  } else {
    const int npairs = nt * (nt + 1) / 2;
    gy = (npairs + WARPS * MAX_PAIRS - 1) / (WARPS * MAX_PAIRS);
  }
  if (cudaError_t e = launch_pdl(schur_kernel<WARPS, MAX_PAIRS, BLOCKED>, dim3(n_partials, gy), WARPS * 32, smem, s.stream, p, partials, schur_sets_per_cta(p), damping, flag,
                                 units);
      e != cudaSuccess)
    return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_schur(const DevProblem& p, double damping, double* partials, int n_partials, int* flag, StreamCtx& s) {
  // every CTA writes all of its tile pairs (zeros when it has no sets), so the partials need no clearing; the 6x6 pose factors are
  // computed by the CTAs themselves
  const int nt = (p.n_aug + 7) >> 3;
  if (nt <= 6) return launch_schur_t<8, 3, false>(p, damping, partials, n_partials, flag, s);
  // measured (profiles/r02_schur_variants.md): up to 14 tiles a side the scattered assignment on 16 warps wins; beyond that the tile pairs
  // no longer fit one CTA and the macro-block assignment halves the number of CTAs that redo the fetch + transform of a slice
  if (nt <= 14) return launch_schur_t<16, 7, false>(p, damping, partials, n_partials, flag, s);
  return launch_schur_t<8, 1, true>(p, damping, partials, n_partials, flag, s);  // 4 x 4 macro blocks of tile pairs per warp; large systems split over gridDim.y
}

cudaError_t launch_schur_finalize(const DevProblem& p, double /*damping*/, const double* partials, int n_partials, bool, StreamCtx& s) {
  const int n2 = p.n_aug * p.n_aug;
  if (cudaError_t e = launch_pdl(schur_finalize_kernel, (SF_LANES * n2 + 255) / 256, 256, 0, s.stream, p, partials, n_partials); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_reduced_solve(const DevProblem& p, double damping, int* pos_def_flag, bool from_peers, StreamCtx& s) {
  const size_t n_rows = ((p.n_aug + RS_NB - 1) / RS_NB) * RS_NB + RS_NB;
  const size_t smem = sizeof(double) * (n_rows * (n_rows + 1) / 2 + 2 * n_rows);
  static size_t attr_smem_dev[MAX_DEVICES] = {};
  if (cudaError_t e = ensure_dynamic_smem(reduced_solve_kernel, smem, attr_smem_dev); e != cudaSuccess) return e;
  static long long* dbg = nullptr;
  static const bool trace = getenv("KB_RS_TRACE") != nullptr;
  if (trace && !dbg) {
    cudaMallocManaged(&dbg, 8 * sizeof(long long));
    cudaMemset(dbg, 0, 8 * sizeof(long long));
  }
  if (cudaError_t e = launch_pdl(reduced_solve_kernel, 1, RS_THREADS, smem, s.stream, p, damping, pos_def_flag, from_peers && p.px.enabled ? 1 : 0, trace ? dbg : nullptr);
      e != cudaSuccess)
    return e;
  KB_LAUNCHED(s);
  if (trace) {  // experiments only: cycles per phase of this launch (load, panel update, diagonal block, panel rows, back substitution)
    cudaStreamSynchronize(s.stream);
    std::fprintf(stderr, "[kb trace] reduced_solve n = %d: load %lld, update %lld, diag %lld, rows %lld, backsub %lld cycles\n", p.n_aug, dbg[0], dbg[1], dbg[2],
                 dbg[3], dbg[4]);
    cudaMemset(dbg, 0, 8 * sizeof(long long));
  }
  return cudaGetLastError();
}

int backsub_blocks(const DevProblem& p) { return ((p.n_sets > 0 ? p.n_sets : 1) + 8 * BS_SETS - 1) / (8 * BS_SETS); }
// back substitution; its blocks also leave the stage-1 partials of the rho denominator / max|dx| in p.rho_partial.
// out2 (optional) = (dx^T (lambda dx + rhs), max|dx|) of this rank, written by the last block; it is also the producer of peer
// exchange B when pos_def_for_exchange is given; lm_mode 1: + the after-solve transition of the device-resident loop (single rank)
cudaError_t launch_backsub(const DevProblem& p, const int* set_col_q, const int* set_col_t, const int* cam_cols, double lambda, int include_shared,
                           double* out2, const int* pos_def_for_exchange, const int* pos_def, int lm_mode, StreamCtx& s) {
  PeerXchg px = p.px;
  if (!pos_def_for_exchange) px.enabled = 0;
  if (cudaError_t e = launch_pdl(backsub_kernel, backsub_blocks(p), 256, 0, s.stream, p, set_col_q, set_col_t, cam_cols, lambda, include_shared,
                                 out2 ? p.rho_partial : nullptr, out2, px, pos_def_for_exchange ? pos_def_for_exchange : pos_def, lm_mode);
      e != cudaSuccess)
    return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_rho_denominator(const DevProblem& p, double lambda, const int* set_col_q, const int* set_col_t, const int*, int include_shared,
                                   double* out2, const int* pos_def_for_exchange, StreamCtx& s) {
  double* partial = p.rho_partial;  // per handle (kb_create): two handles on one GPU never share scratch
  const int blocks = max(1, min(RHO_BLOCKS, (p.n_sets + 255) / 256));
  rho_stage1_kernel<<<blocks, 256, 0, s.stream>>>(p, lambda, set_col_q, set_col_t, include_shared, partial);
  KB_LAUNCHED(s);
  PeerXchg px = p.px;
  if (!pos_def_for_exchange) px.enabled = 0;  // a rho query outside a solve is not an exchange step
  rho_stage2_kernel<<<1, RHO2_THREADS, 0, s.stream>>>(p.ctrl, px, pos_def_for_exchange, partial, blocks, out2, 0);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_pack_rank_scalars(double* pk, int rank, int n_ranks, const double* rho_max, const int* pos_def, StreamCtx& s) {
  pack_rank_scalars_kernel<<<1, 128, 0, s.stream>>>(pk, rank, n_ranks, rho_max, pos_def);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

cudaError_t launch_px_reduce_system(const DevProblem& p, StreamCtx& s) {
  px_reduce_system_kernel<<<(p.n_aug * p.n_aug + 255) / 256, 256, 0, s.stream>>>(p);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_px_combine_solve(const DevProblem& p, double* rho_max, int* pos_def_flag, int lm_mode, StreamCtx& s) {
  if (cudaError_t e = launch_pdl(px_combine_solve_kernel, 1, 1, 0, s.stream, p, rho_max, pos_def_flag, lm_mode); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_px_combine_cost(const DevProblem& p, double* cost, int lm_mode, double* trace, int* pos_def, StreamCtx& s) {
  if (cudaError_t e = launch_pdl(px_combine_cost_kernel, 1, 1, 0, s.stream, p, cost, lm_mode, trace, pos_def); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
// Eigen-decomposition of the reduced system in p.Sred: QL (always), then the Jacobi polish when `polish` (unscaled, graded systems).
// status[0] = Jacobi sweeps of the polish (0 without it), status[1] = 1 when the QL iteration did not converge.
// V_warm (may be null): eigenvectors of a NEARBY system (the previous solve / analysis of this handle).  The reduced system changes
// little from one Gauss-Newton iteration or batch to the next, so the one-sided Jacobi iteration started from them converges in 2-3
// sweeps and replaces the whole QL stage (whose rotation chain is serial); V_keep receives the new vectors for the next call.
static cudaError_t launch_sym_eig(const DevProblem& p, double* G, double* V, double* sv_out, double* V_out, double* V_tmp, int* status, bool polish,
                                  const double* V_warm, double* V_keep, StreamCtx& s) {
  const size_t n = (size_t)p.n_c, ld = n | 1;
  if (V_warm) {
    polish = true;
    cudaError_t e = cudaMemsetAsync(status + 1, 0, 5 * sizeof(int), s.stream);  // no QL stage: its status / diagnostics read 0
    if (e != cudaSuccess) return e;
  } else {
    const size_t bytes = sizeof(double) * 2 * n * ld;
    const size_t smem = bytes <= 200 * 1024 ? bytes : 0;
    static size_t attr_smem_dev[MAX_DEVICES] = {};
    if (cudaError_t e = ensure_dynamic_smem(sym_eig_kernel, smem, attr_smem_dev); e != cudaSuccess) return e;
    sym_eig_kernel<<<1, EIG_THREADS, smem, s.stream>>>(p, G, V, sv_out, polish ? V_tmp : V_out, status + 1, smem ? 1 : 0);
    KB_LAUNCHED(s);
  }
  if (polish) {
    const size_t np = (n + 1) & ~(size_t)1;
    const size_t bytes = sizeof(double) * 2 * np * n;
    const size_t smem = bytes <= 220 * 1024 ? bytes : 0;
    static size_t attr_smem_dev[MAX_DEVICES] = {};
    if (cudaError_t e = ensure_dynamic_smem(jacobi_polish_kernel, smem, attr_smem_dev); e != cudaSuccess) return e;
    jacobi_polish_kernel<<<1, EIG_THREADS, smem, s.stream>>>(p, V_warm ? V_warm : V_tmp, G, V, sv_out, V_out, status, smem ? 1 : 0);
    KB_LAUNCHED(s);
  } else {
    cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), s.stream);
    if (e != cudaSuccess) return e;
  }
  if (V_keep) {
    cudaError_t e = cudaMemcpyAsync(V_keep, V_out, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, s.stream);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}
cudaError_t launch_marginal_eig(const DevProblem& p, double* G, double* V, double* sv_out, double* V_out, double* V_tmp, int* status, const double* V_warm,
                                double* V_keep, StreamCtx& s) {
  return launch_sym_eig(p, G, V, sv_out, V_out, V_tmp, status, true, V_warm, V_keep, s);
}
cudaError_t launch_camera_diag(const DevProblem& p, double* out, StreamCtx& s) {
  camera_diag_kernel<<<1, 256, 0, s.stream>>>(p, out);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_svd_solve(const DevProblem& p, const double* diag_h, double norm_tol, int column_scaling, double eps_svd, double svd_tol, double* g,
                             double* G, double* V, double* sv, double* V_out, double* V_tmp, int* sweeps, double* result, const double* V_warm,
                             double* V_keep, StreamCtx& s) {
  svd_scale_kernel<<<1, 256, 0, s.stream>>>(p, diag_h, norm_tol, column_scaling, g);
  KB_LAUNCHED(s);
  // the column-scaled system is well graded: absolute accuracy suffices (tolerance 1e-6 n sv_0); the unscaled one gets the polish
  if (cudaError_t e = launch_sym_eig(p, G, V, sv, V_out, V_tmp, sweeps, column_scaling == 0, V_warm, V_keep, s); e != cudaSuccess) return e;
  svd_truncated_solve_kernel<<<1, 256, 0, s.stream>>>(p, sv, V_out, g, eps_svd, svd_tol, result);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_lm_post_solve(const DevProblem& p, const int* pos_def_flag, const double* rho_max, const double* rank_slots, int n_ranks, StreamCtx& s) {
  lm_post_solve_kernel<<<1, 1, 0, s.stream>>>(p.ctrl, pos_def_flag, rho_max, rank_slots, n_ranks);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_lm_boundary(const DevProblem& p, double* trace, int* pos_def, StreamCtx& s) {
  lm_boundary_kernel<<<1, 1, 0, s.stream>>>(p.ctrl, trace, pos_def);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_lm_revert(const DevProblem& p, const double* backup_cam, const double* backup_base, const double* backup_sets, StreamCtx& s) {
  const int n = max(max(p.n_sets * POSE_STRIDE, p.n_cams * CAM_PARAM_STRIDE), 1);
  lm_revert_kernel<<<(n + 255) / 256, 256, 0, s.stream>>>(p, backup_cam, backup_base, backup_sets);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}
cudaError_t launch_lm_finish(const DevProblem& p, StreamCtx& s) {
  lm_finish_kernel<<<1, 1, 0, s.stream>>>(p.ctrl);
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

// measurements that arrived in single precision (the detector's type): exact widening into the FP64 arrays the kernels read
__global__ void __launch_bounds__(256) widen_observations_kernel(const float* __restrict__ su, const float* __restrict__ sv, double* __restrict__ du,
                                                                 double* __restrict__ dv, long long n) {
  const long long i2 = 2ll * (blockIdx.x * (long long)blockDim.x + threadIdx.x);  // two terms per thread: 8-byte loads, 16-byte stores
  if (i2 + 1 < n) {
    const float2 a = *reinterpret_cast<const float2*>(su + i2), b = *reinterpret_cast<const float2*>(sv + i2);
    *reinterpret_cast<double2*>(du + i2) = make_double2((double)a.x, (double)a.y);
    *reinterpret_cast<double2*>(dv + i2) = make_double2((double)b.x, (double)b.y);
  } else if (i2 < n) {
    du[i2] = (double)su[i2];
    dv[i2] = (double)sv[i2];
  }
}
cudaError_t launch_widen_observations(const float* su, const float* sv, double* du, double* dv, long long n, cudaStream_t stream, long long* launches) {
  if (n <= 0) return cudaSuccess;
  const long long threads = (n + 1) / 2;
  widen_observations_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(su, sv, du, dv, n);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_apply_update(const DevProblem& p, const int* set_col_q, const int* set_col_t, const int*, double* backup_cam, double* backup_base,
                                double* backup_sets, StreamCtx& s) {
  const int n = p.n_sets > 0 ? p.n_sets : 1;
  if (cudaError_t e = launch_pdl(apply_update_kernel, (n + 255) / 256 + 1, 256, 0, s.stream, p, set_col_q, set_col_t, backup_cam, backup_base, backup_sets); e != cudaSuccess) return e;
  KB_LAUNCHED(s);
  return cudaGetLastError();
}

}  // namespace kb
