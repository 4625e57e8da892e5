// C ABI (include/kalibr_b200.h) and host-side logic of the B200 batch-calibration hot path: design-variable
// ordering of the three kalibr2 drivers, sharding of synced sets over ranks, device buffer ownership, the call
// sequence of LinearSystemSolver (evaluate -> build -> setConditioner -> solve -> update/revert), NCCL plumbing for
// the reduced system, and the export of the Hessian block pattern / CCS Jacobian structure for parity checks.
// There is no CPU compute path in here: every number comes from the kernels in kb_kernels.cu.
#include <atomic>
#include <thread>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/kalibr_b200.h"
#include "../../include/kalibr_b200/optimizer.hpp"
#include "kb_device.cuh"

using namespace kb;
using kalibr_b200::backend::KbError;

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency; single-GPU use never touches it)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::mutex mu;
  bool load(std::string& err) {
    std::lock_guard<std::mutex> lock(mu);
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy || !GetErrorString) { err = "libnccl is missing symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3;
thread_local std::string g_create_error;  // kb_last_error(NULL) reports the calling thread's last kb_create failure

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0, cap = 0;  // logical size, allocated elements
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; cap = 0; }
  // `count` elements with undefined contents; an allocation that is already large enough is kept (handles are re-sized in place)
  cudaError_t alloc(size_t count) {
    if (p && count <= cap) { n = count; return cudaSuccess; }
    release();
    const size_t c = count ? count : 1;
    cudaError_t e = cudaMalloc((void**)&p, sizeof(T) * c);
    if (e == cudaSuccess) { n = count; cap = c; }
    return e;
  }
  // grow to `count` elements KEEPING the first n (capacity doubling): the append-only arrays of a live handle
  cudaError_t grow(size_t count, cudaStream_t s) {
    if (p && count <= cap) { n = count; return cudaSuccess; }
    const size_t c = std::max<size_t>(std::max(count, 2 * cap), 1);
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, sizeof(T) * c);
    if (e != cudaSuccess) return e;
    if (p && n) {
      e = cudaMemcpyAsync(q, p, sizeof(T) * n, cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q;
    cap = c;
    n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s);
  }
};
}  // namespace

constexpr int KB_STREAM_CHUNKS = 4;

struct kb_handle {
  std::string error;
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  // ---- host-side structure ----
  int driver_order = 0, n_cams = 0, n_sets_global = 0, set_lo = 0, set_hi = 0, n_ranks = 1, rank = 0;
  int64_t n_terms_global = 0, n_terms_local = 0;
  std::vector<int> cam_model;
  std::vector<int> dv_dim, dv_col;  // active design variables in insertion order
  std::vector<int> dv_proj, dv_dist, dv_base_q, dv_base_t;
  int dv_first_set = 0;             // block index of q of set 0 (q_v = first + 2v, t_v = first + 2v + 1)
  int64_t jcols = 0;
  std::vector<int> h_view_set, h_view_cam, h_view_begin;  // local
  std::vector<long long> h_view_jbase;
  int64_t jac_nnz = 0;
  std::vector<int> h_cam_cols;
  // ---- device ----
  DevProblem d;
  DevBuf<double> y_u, y_v, target, cam_params, baselines, set_poses, camT, camPi, camA, baseBt, baseM, e, view_cost, set_prep, VB, gram_partial, sumG, V, bv, W, Lv, yv, U, Sred, dxc, dx;
  DevBuf<double> init_cam, init_base, init_sets, bk_cam, bk_base, bk_sets, partials, scalars, jt, gather, rho_partial, sv_cam, sv_base, sv_sets;
  size_t saved_sets = 0;
  bool has_saved = false;
  DevBuf<uint16_t> corner;
  DevBuf<int> view_set, view_cam, view_begin, set_view, lin_off, view_list, cam_view_list, cam_view_begin, set_col_q, set_col_t, cam_cols, posdef;
  DevBuf<long long> view_jbase;
  DevBuf<int4> slices, vmeta;
  DevBuf<int> cam_slice_range;
  int slice_model_begin[KB_NUM_MODELS + 1] = {};
  // second slice table for the streamed evaluate: the views are cut into KB_STREAM_CHUNKS contiguous term ranges, slices are
  // chunk-major and never cross a chunk, so the fused kernel can start on a chunk as soon as its observations have landed
  DevBuf<int4> st_slices;
  DevBuf<int> st_cam_slice_range;                       // [n_cams][KB_STREAM_CHUNKS][2]
  int st_chunk_model_begin[KB_STREAM_CHUNKS][KB_NUM_MODELS + 1] = {};
  int64_t st_chunk_term[KB_STREAM_CHUNKS + 1] = {};
  cudaStream_t copy_stream = nullptr;
  cudaStream_t model_stream[KB_NUM_MODELS] = {};  // per-model launches of a mixed rig run concurrently (fork / join)
  cudaEvent_t ev_fork = nullptr, ev_join[KB_NUM_MODELS] = {};
  cudaEvent_t ev_chunk[KB_STREAM_CHUNKS] = {}, ev_main = nullptr;
  // double-buffered observations: kb_prefetch_observations fills the back buffers on the copy stream while the front ones
  // are in use, kb_commit_observations swaps them
  DevBuf<double> y_u_back, y_v_back;
  DevBuf<float> stage_u, stage_v;  // single-precision measurements land here and are widened on the device (the *_f32 entry points)
  double *front_u = nullptr, *front_v = nullptr, *back_u = nullptr, *back_v = nullptr;
  cudaEvent_t ev_prefetch = nullptr, ev_front_free = nullptr;
  bool prefetch_pending = false, front_free_recorded = false;
  DevBuf<unsigned int> n_invalid, lm_counters, tickets;
  DevBuf<int> col_desc;
  int lm_bfrag_pairs[KB_NUM_MODELS] = {};
  int model_begin[KB_NUM_MODELS + 1] = {};
  int n_partials = 0;
  double* h_scalars = nullptr;  // pinned [8 + 4 * MAX_CAMS]: scalars, then the per-rank slots of the packed all-reduce
  DevBuf<double> rank_slots;    // [n_ranks][4]
  DevBuf<LmCtrl> ctrl;          // control block of the device-resident LM loop (neutral flags outside kb_optimize)
  DevBuf<double> eig_G, eig_V, eig_sv, eig_Vout, eig_Vtmp;  // marginal analysis scratch / results
  // eigenvectors of the last decomposition per system kind (0: unscaled marginal system, 1: column-scaled system of the solver): the
  // warm start of the next one.  Every 32nd decomposition goes through QL again, which also renews the orthogonality of the vectors.
  DevBuf<double> eig_warm[2];
  int eig_warm_age[2] = {-1, -1};  // -1: no vectors yet
  DevBuf<int> eig_sweeps;
  DevBuf<double> trace_dev;
  LmCtrl* h_ctrl = nullptr;     // pinned
  // peer exchange over NVLink (kb_attach_peers): own buffer + the peers' buffers opened through CUDA IPC
  DevBuf<double> px_buf;
  void* px_opened[PX_MAX_RANKS] = {};
  bool px_on = false;
  cudaGraphExec_t lm_graph = nullptr;  // one LM iteration, captured once per handle (single rank, stage timing off)
  bool lm_warm = false;                // a plain-launched iteration has run (first-launch attribute / allocation calls are done)
  const double* lm_graph_trace = nullptr;
  long long lm_graph_kernels = 0;      // kernels one replay of the graph launches
  int* h_posdef = nullptr;      // pinned
  // ---- solver state ----
  double lambda = 0.0;          // _diagonalConditioner (constant)
  double rho_lambda = -1.0;     // lambda the cached dx^T(lambda dx + rhs) of the last solve was computed with
  double diag_residual = 0.0;   // what the lambda^2 / lambda asymmetry leaves on diag(H) since the last build (Q2)
  int semantic = 0;
  bool built = false, solved = false, has_backup = false, presharded = false;
  bool speculative = true;       // kb_evaluate_error linearises too, so that a build at the same state is free
  unsigned timing_mask = ~0u;    // stages that kb_enable_stage_timing selected
  bool defer_sync = false;       // inside kb_iterate: the entry points enqueue only, one synchronisation at the end
  long long state_version = 0;   // bumped whenever design variables or observations change
  long long la_version = -1;     // state the view blocks / Gram sums were computed at
  int la_rows = 0;               // DevProblem::mest_rows they were computed with
  double inv_r[4] = {1.0, 0.0, 0.0, 1.0}, sqrt_inv_r[4] = {1.0, 0.0, 0.0, 1.0};  // row-major
  DevBuf<double> stats_e, stats_acc;  // reprojection statistics scratch
  DevBuf<double> pnp_T;               // initial-guess stage: T_target_camera per view
  DevBuf<int> pnp_ok, pnp_res, pnp_set_ok;
  DevBuf<unsigned char> pnp_mask;
  kb_svd_solve_result last_svd = {0, -1, -1, 0.0, 0.0};  // of the last kb_solve_system_svd (rank -1: none yet)
  DevBuf<double> svd_diag, svd_g, svd_result;  // truncated-SVD solve: diag of the camera block, column scales, {rank, tol, gap}
  DevBuf<double> init_scratch, init_out;  // initializeIntrinsics: candidates / guesses, results
  std::vector<double> trace;
  // ---- multi-GPU ----
  NcclComm comm = nullptr;
  // ---- timing ----
  bool timing = false;
  static constexpr int STAGE_RING = 64;          // stage measurements that may wait for a synchronisation (kb_iterate without wait)
  cudaEvent_t ev[2 * KB_NUM_STAGES * STAGE_RING] = {};
  long long stage_head[KB_NUM_STAGES] = {}, stage_tail[KB_NUM_STAGES] = {};  // measurements recorded / collected
  double stage_ms[KB_NUM_STAGES] = {};
  double stage_total[KB_NUM_STAGES] = {};
  long long stage_calls[KB_NUM_STAGES] = {};
};

namespace {

#define KB_CUDA(h, call)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      (h)->error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
      return KB_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

kb_status fail(kb_handle* h, kb_status code, const std::string& msg) {
  if (h) h->error = msg; else g_create_error = msg;
  return code;
}

void shard_range(int n_sets, int n_ranks, int rank, int& lo, int& hi) {
  const int base = n_sets / n_ranks, rem = n_sets % n_ranks;
  lo = rank * base + std::min(rank, rem);
  hi = lo + base + (rank < rem ? 1 : 0);
}

StreamCtx ctx(kb_handle* h) {
  StreamCtx c{h->stream, &h->launches};
  c.side = h->model_stream;
  c.ev_fork = h->ev_fork;
  c.ev_join = h->ev_join;
  return c;
}

struct StageTimer {
  kb_handle* h;
  int stage;
  StageTimer(kb_handle* h_, int s) : h(h_), stage(s) { if (on()) cudaEventRecord(h->ev[slot()], h->stream); }
  bool on() const { return h->timing && ((h->timing_mask >> stage) & 1u); }
  ~StageTimer() {
    if (on()) {
      cudaEventRecord(h->ev[slot() + 1], h->stream);
      ++h->stage_head[stage];
    }
  }
  int slot() const { return 2 * (stage * kb_handle::STAGE_RING + (int)(h->stage_head[stage] % kb_handle::STAGE_RING)); }
};
// call after a stream synchronisation: folds every finished stage measurement into the totals
void collect_stages(kb_handle* h) {
  if (!h->timing) return;
  for (int s = 0; s < KB_NUM_STAGES; ++s) {
    for (long long i = std::max(h->stage_tail[s], h->stage_head[s] - kb_handle::STAGE_RING); i < h->stage_head[s]; ++i) {
      const int slot = 2 * (s * kb_handle::STAGE_RING + (int)(i % kb_handle::STAGE_RING));
      float ms = 0;
      if (cudaEventElapsedTime(&ms, h->ev[slot], h->ev[slot + 1]) == cudaSuccess) {
        h->stage_ms[s] = ms;
        h->stage_total[s] += ms;
        h->stage_calls[s] += 1;
      }
    }
    h->stage_tail[s] = h->stage_head[s];
  }
}

kb_status nccl_allreduce(kb_handle* h, void* buf, size_t count, int dtype, int op) {
  if (h->n_ranks <= 1) return KB_OK;
  int r = g_nccl.AllReduce(buf, buf, count, dtype, op, h->comm, h->stream);
  if (r != 0) return fail(h, KB_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  return KB_OK;
}

// design-variable layout of the three drivers
void build_dv_layout(kb_handle* h) {
  h->dv_dim.clear();
  h->dv_proj.assign(h->n_cams, -1);
  h->dv_dist.assign(h->n_cams, -1);
  h->dv_base_q.assign(std::max(h->n_cams - 1, 0), -1);
  h->dv_base_t.assign(std::max(h->n_cams - 1, 0), -1);
  auto intr = [&](int k) {
    h->dv_proj[k] = (int)h->dv_dim.size();
    h->dv_dim.push_back(model_P(h->cam_model[k]));
    h->dv_dist[k] = (int)h->dv_dim.size();
    h->dv_dim.push_back(model_D(h->cam_model[k]));
  };
  auto base = [&]() {
    for (int j = 0; j + 1 < h->n_cams; ++j) {
      h->dv_base_q[j] = (int)h->dv_dim.size();
      h->dv_dim.push_back(3);
      h->dv_base_t[j] = (int)h->dv_dim.size();
      h->dv_dim.push_back(3);
    }
  };
  auto sets = [&]() {
    h->dv_first_set = (int)h->dv_dim.size();
    for (int v = 0; v < h->n_sets_global; ++v) { h->dv_dim.push_back(3); h->dv_dim.push_back(3); }
  };
  if (h->driver_order == KB_ORDER_SINGLE) { intr(0); sets(); }
  else if (h->driver_order == KB_ORDER_STEREO) { base(); sets(); intr(0); intr(1); }
  else if (h->driver_order == KB_ORDER_BATCH) { sets(); base(); for (int k = 0; k < h->n_cams; ++k) intr(k); }
  else { for (int k = 0; k < h->n_cams; ++k) intr(k); base(); sets(); }
  h->dv_col.resize(h->dv_dim.size());
  int c = 0;
  for (size_t i = 0; i < h->dv_dim.size(); ++i) { h->dv_col[i] = c; c += h->dv_dim[i]; }
  h->jcols = c;
}

// camera-side design variables a view of camera k touches
std::vector<int> camside_dvs(const kb_handle* h, int k) {
  std::vector<int> v = {h->dv_proj[k], h->dv_dist[k]};
  for (int j = 0; j < k; ++j) { v.push_back(h->dv_base_q[j]); v.push_back(h->dv_base_t[j]); }
  return v;
}

}  // namespace

namespace {
struct CreateTrace {  // KB_CREATE_TRACE=1: wall-clock checkpoints of kb_create / kb_destroy on stderr
  bool on = getenv("KB_CREATE_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[kb trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count());
    t0 = t;
  }
};
}  // namespace

// =========================================================================================================
extern "C" {

const char* kb_last_error(const kb_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

kb_status kb_nccl_unique_id(char out[128]) {
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, KB_ERR_NCCL, err);
  NcclUniqueId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(nullptr, KB_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  std::memcpy(out, id.internal, 128);
  return KB_OK;
}

void kb_destroy(kb_handle* h) {
  if (!h) return;
  CreateTrace trace;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) g_nccl.CommDestroy(h->comm);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_chunk) if (e) cudaEventDestroy(e);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (auto& e : h->ev_join) if (e) cudaEventDestroy(e);
  for (auto& st : h->model_stream) if (st) cudaStreamDestroy(st);
  if (h->ev_prefetch) cudaEventDestroy(h->ev_prefetch);
  if (h->ev_front_free) cudaEventDestroy(h->ev_front_free);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h_scalars) cudaFreeHost(h->h_scalars);
  if (h->h_posdef) cudaFreeHost(h->h_posdef);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  for (auto& q : h->px_opened) if (q) cudaIpcCloseMemHandle(q);
  if (h->lm_graph) cudaGraphExecDestroy(h->lm_graph);
  if (h->stream) cudaStreamDestroy(h->stream);
  trace.mark("destroy: streams, events");
  delete h;
  trace.mark("destroy: device buffers");
}

// Everything DERIVED from the host-side master copy of the structure (cameras, local synced sets, views as term ranges): the
// design-variable layout, the reduced-system layout, the CCS J^T descriptors, the view lists / slice tables, the work buffers (kept
// when their capacity suffices) and the device-side problem view.  Called by kb_create and again after kb_append_set /
// kb_remove_last_set.  Observations, target points and the state arrays are NOT touched here.
static kb_status build_tables(kb_handle* h) {
  cudaStream_t s = h->stream;
  const int n_local_sets = h->set_hi - h->set_lo;
  build_dv_layout(h);
  std::vector<int>& vs = h->h_view_set;
  std::vector<int>& vc = h->h_view_cam;
  std::vector<int>& vb = h->h_view_begin;
  const int n_views = (int)vs.size();
  std::vector<int> set_view((size_t)n_local_sets * h->n_cams, -1);
  for (int w = 0; w < n_views; ++w) set_view[(size_t)vs[w] * h->n_cams + vc[w]] = w;
  std::fill(std::begin(h->lm_bfrag_pairs), std::end(h->lm_bfrag_pairs), 0);
  // ---- reduced-system layout: [cam0 proj|dist, cam1 ..., | baseline 0 q,t, ...] ----
  DevProblem& D = h->d;
  D.n_cams = h->n_cams;
  D.n_sets = n_local_sets;
  D.n_views = n_views;
  D.n_terms = h->n_terms_local;
  int off = 0;
  for (int k = 0; k < h->n_cams; ++k) {
    D.cam_model[k] = h->cam_model[k];
    D.cam_P[k] = model_P(h->cam_model[k]);
    D.cam_D[k] = model_D(h->cam_model[k]);
    D.intr_off[k] = off;
    off += D.cam_P[k] + D.cam_D[k];
  }
  for (int j = 0; j + 1 < h->n_cams; ++j) { D.base_off[j] = off; off += 6; }
  D.n_c = off;
  D.n_aug = off + 1;
  if (D.n_aug > 224) return fail(h, KB_ERR_INVALID_ARGUMENT, "reduced camera system larger than 223 unknowns is not supported");
  h->h_cam_cols.assign(D.n_c, 0);
  for (int k = 0; k < h->n_cams; ++k) {
    for (int c = 0; c < D.cam_P[k]; ++c) h->h_cam_cols[D.intr_off[k] + c] = h->dv_col[h->dv_proj[k]] + c;
    for (int c = 0; c < D.cam_D[k]; ++c) h->h_cam_cols[D.intr_off[k] + D.cam_P[k] + c] = h->dv_col[h->dv_dist[k]] + c;
  }
  for (int j = 0; j + 1 < h->n_cams; ++j)
    for (int c = 0; c < 3; ++c) {
      h->h_cam_cols[D.base_off[j] + c] = h->dv_col[h->dv_base_q[j]] + c;
      h->h_cam_cols[D.base_off[j] + 3 + c] = h->dv_col[h->dv_base_t[j]] + c;
    }
  std::vector<int> set_col_q(n_local_sets), set_col_t(n_local_sets);
  for (int lv = 0; lv < n_local_sets; ++lv) {
    set_col_q[lv] = h->dv_col[h->dv_first_set + 2 * (h->set_lo + lv)];
    set_col_t[lv] = h->dv_col[h->dv_first_set + 2 * (h->set_lo + lv) + 1];
  }
  // ---- CCS J^T slot offsets per camera: design variables of a term sorted by block index ----
  std::vector<int> lin_off((size_t)h->n_cams * LIN_OFF_STRIDE, 0);
  h->h_view_jbase.assign(n_views, 0);
  {
    const int pose_q_block = h->dv_first_set;  // relative order against the camera-side blocks is the same for every set
    for (int k = 0; k < h->n_cams; ++k) {
      struct Seg { int block, dim, slot; };
      std::vector<Seg> segs;
      segs.push_back({pose_q_block, 3, 0});
      segs.push_back({pose_q_block + 1, 3, 1});
      segs.push_back({h->dv_proj[k], D.cam_P[k], 2});
      segs.push_back({h->dv_dist[k], D.cam_D[k], 3});
      for (int j = 0; j < k; ++j) segs.push_back({h->dv_base_q[j], 6, 4 + j});  // q,t adjacent in every order
      std::sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.block < b.block; });
      int o = 0;
      for (auto& s : segs) { lin_off[(size_t)k * LIN_OFF_STRIDE + s.slot] = o; o += s.dim; }
    }
    // per-column descriptors of the CCS J^T layout and the DMMA B-fragment slots the materialising kernel needs per model
    const int desc_stride = 6 * h->n_cams + KB_CAM_PARAM_STRIDE;
    std::vector<int> col_desc((size_t)h->n_cams * desc_stride, 3 << 16);
    for (int k = 0; k < h->n_cams; ++k) {
      const int* off = &lin_off[(size_t)k * LIN_OFF_STRIDE];
      int* cd = &col_desc[(size_t)k * desc_stride];
      for (int c = 0; c < 3; ++c) { cd[off[0] + c] = (0 << 16) | c; cd[off[1] + c] = (0 << 16) | (3 + c); }
      for (int c = 0; c < D.cam_P[k]; ++c) cd[off[2] + c] = (2 << 16) | c;
      for (int c = 0; c < D.cam_D[k]; ++c) cd[off[3] + c] = (2 << 16) | (D.cam_P[k] + c);
      for (int j = 0; j < k; ++j)
        for (int c = 0; c < 6; ++c) cd[off[4 + j] + c] = (1 << 16) | (j << 8) | c;
      const int Wk = 6 + 6 * k + D.cam_P[k] + D.cam_D[k];
      int pairs = 0;
      for (int n0 = 0; n0 < Wk; n0 += 8) {
        unsigned need = 0;
        for (int c = n0; c < std::min(Wk, n0 + 8); ++c) {
          const int kind = cd[c] >> 16, sub = cd[c] & 0xff;
          need |= kind == 2 ? 1u << ((6 + sub) >> 2) : 3u;
        }
        for (int ks = 0; ks < 4; ++ks) pairs += (need >> ks) & 1;
      }
      h->lm_bfrag_pairs[h->cam_model[k]] = std::max(h->lm_bfrag_pairs[h->cam_model[k]], pairs);
    }
    KB_CUDA(h, h->col_desc.upload(col_desc, h->stream));
    D.col_desc_stride = desc_stride;
    long long jb = 0;
    for (int w = 0; w < n_views; ++w) {
      h->h_view_jbase[w] = jb;
      const int k = vc[w];
      jb += (long long)(vb[w + 1] - vb[w]) * 2 * (6 + 6 * k + D.cam_P[k] + D.cam_D[k]);
    }
    h->jac_nnz = jb;
  }
  // ---- view lists: by model (kernel specialisation) and by camera (Gram sums) ----
  std::vector<int> view_list;
  for (int m = 0; m < KB_NUM_MODELS; ++m) {
    h->model_begin[m] = (int)view_list.size();
    for (int k = 0; k < h->n_cams; ++k)
      if (h->cam_model[k] == m)
        for (int w = 0; w < n_views; ++w)
          if (vc[w] == k) view_list.push_back(w);
  }
  h->model_begin[KB_NUM_MODELS] = (int)view_list.size();
  std::vector<int> cam_view_list, cam_view_begin(h->n_cams + 1, 0);
  for (int k = 0; k < h->n_cams; ++k) {
    cam_view_begin[k] = (int)cam_view_list.size();
    for (int w = 0; w < n_views; ++w)
      if (vc[w] == k) cam_view_list.push_back(w);
  }
  cam_view_begin[h->n_cams] = (int)cam_view_list.size();
  // ---- slices of the model/camera-sorted view list: one warp of the fused kernel per slice, never crossing a camera ----
  std::vector<int4> slices;
  std::vector<int> cam_slice_range(2 * (size_t)h->n_cams, 0);
  {
    const int per = std::max(1, (n_views + la_grid_warps() - 1) / la_grid_warps());
    int pos = 0;  // position in view_list
    for (int m = 0; m < KB_NUM_MODELS; ++m) {
      h->slice_model_begin[m] = (int)slices.size();
      for (int k = 0; k < h->n_cams; ++k) {
        if (h->cam_model[k] != m) continue;
        const int nk = cam_view_begin[k + 1] - cam_view_begin[k];
        cam_slice_range[2 * k] = (int)slices.size();
        for (int a = 0; a < nk; a += per) slices.push_back(make_int4(pos + a, pos + std::min(nk, a + per), k, 0));
        cam_slice_range[2 * k + 1] = (int)slices.size();
        pos += nk;
      }
    }
    h->slice_model_begin[KB_NUM_MODELS] = (int)slices.size();
  }
  // streamed table: chunk c = local views [cv[c], cv[c+1]) = terms [st_chunk_term[c], st_chunk_term[c+1])
  std::vector<int4> st_slices;
  std::vector<int> st_cam_slice_range((size_t)h->n_cams * KB_STREAM_CHUNKS * 2, 0);
  {
    int cv[KB_STREAM_CHUNKS + 1];
    for (int c = 0; c <= KB_STREAM_CHUNKS; ++c) {
      cv[c] = (int)((long long)n_views * c / KB_STREAM_CHUNKS);
      h->st_chunk_term[c] = vb[cv[c]];
    }
    // position of every camera's first view in the model/camera-sorted list
    std::vector<int> cam_pos(h->n_cams, 0);
    {
      int pos = 0;
      for (int m = 0; m < KB_NUM_MODELS; ++m)
        for (int k = 0; k < h->n_cams; ++k)
          if (h->cam_model[k] == m) { cam_pos[k] = pos; pos += cam_view_begin[k + 1] - cam_view_begin[k]; }
    }
    for (int c = 0; c < KB_STREAM_CHUNKS; ++c) {
      const int per = std::max(1, (cv[c + 1] - cv[c] + la_grid_warps() - 1) / la_grid_warps());
      for (int m = 0; m < KB_NUM_MODELS; ++m) {
        h->st_chunk_model_begin[c][m] = (int)st_slices.size();
        for (int k = 0; k < h->n_cams; ++k) {
          if (h->cam_model[k] != m) continue;
          // views of camera k inside the chunk: a contiguous run of its (ascending) view list
          const int* lst = cam_view_list.data() + cam_view_begin[k];
          const int nk = cam_view_begin[k + 1] - cam_view_begin[k];
          const int lo = (int)(std::lower_bound(lst, lst + nk, cv[c]) - lst), hi = (int)(std::lower_bound(lst, lst + nk, cv[c + 1]) - lst);
          st_cam_slice_range[((size_t)k * KB_STREAM_CHUNKS + c) * 2] = (int)st_slices.size();
          for (int a = lo; a < hi; a += per) st_slices.push_back(make_int4(cam_pos[k] + a, cam_pos[k] + std::min(hi, a + per), k, 0));
          st_cam_slice_range[((size_t)k * KB_STREAM_CHUNKS + c) * 2 + 1] = (int)st_slices.size();
        }
      }
      h->st_chunk_model_begin[c][KB_NUM_MODELS] = (int)st_slices.size();
    }
  }

  KB_CUDA(h, h->view_set.upload(vs, s));
  KB_CUDA(h, h->view_cam.upload(vc, s));
  KB_CUDA(h, h->view_begin.upload(vb, s));
  KB_CUDA(h, h->set_view.upload(set_view, s));
  KB_CUDA(h, h->lin_off.upload(lin_off, s));
  KB_CUDA(h, h->view_jbase.upload(h->h_view_jbase, s));
  KB_CUDA(h, h->view_list.upload(view_list, s));
  KB_CUDA(h, h->cam_view_list.upload(cam_view_list, s));
  KB_CUDA(h, h->cam_view_begin.upload(cam_view_begin, s));
  KB_CUDA(h, h->slices.upload(slices, s));
  {
    std::vector<int4> vmeta(view_list.size());
    for (size_t i = 0; i < view_list.size(); ++i) {
      const int w = view_list[i];
      vmeta[i] = make_int4(w, vs[w], vb[w], vb[w + 1]);
    }
    KB_CUDA(h, h->vmeta.upload(vmeta, s));
  }
  KB_CUDA(h, h->cam_slice_range.upload(cam_slice_range, s));
  KB_CUDA(h, h->st_slices.upload(st_slices, s));
  KB_CUDA(h, h->st_cam_slice_range.upload(st_cam_slice_range, s));
  KB_CUDA(h, h->set_col_q.upload(set_col_q, s));
  KB_CUDA(h, h->set_col_t.upload(set_col_t, s));
  KB_CUDA(h, h->cam_cols.upload(h->h_cam_cols, s));
  const size_t C = h->n_cams, S = n_local_sets, NA = D.n_aug;

  KB_CUDA(h, h->camT.alloc(C * 12));
  KB_CUDA(h, h->camPi.alloc(C * 36));
  KB_CUDA(h, h->camA.alloc(C * C * 36));
  KB_CUDA(h, h->baseBt.alloc(C * 36));
  KB_CUDA(h, h->baseM.alloc(C * 36));
  KB_CUDA(h, h->e.alloc(2 * (size_t)h->n_terms_local));
  KB_CUDA(h, h->view_cost.alloc(n_views));
  KB_CUDA(h, h->set_prep.alloc(S * SETPREP_STRIDE));
  KB_CUDA(h, h->VB.alloc((size_t)n_views * VB_STRIDE));
  KB_CUDA(h, h->gram_partial.alloc(std::max(slices.size(), st_slices.size()) * GRAM_TILES));
  KB_CUDA(h, h->sumG.alloc(C * GRAM_SIZE));
  KB_CUDA(h, h->V.alloc(S * 36));
  KB_CUDA(h, h->bv.alloc(S * 6));
  KB_CUDA(h, h->W.alloc(S * D.n_c * 6));
  KB_CUDA(h, h->Lv.alloc(S * 36));
  KB_CUDA(h, h->yv.alloc(S * 6));
  KB_CUDA(h, h->U.alloc(NA * NA));
  KB_CUDA(h, h->Sred.alloc(NA * NA));
  KB_CUDA(h, h->dxc.alloc(D.n_c));
  KB_CUDA(h, h->dx.alloc((size_t)h->jcols));
  KB_CUDA(h, h->scalars.alloc(8));
  KB_CUDA(h, h->rho_partial.alloc(2 * std::max<size_t>(64, (size_t)n_local_sets / 8 + 2)));
  D.ctrl = h->ctrl.p;
  KB_CUDA(h, h->posdef.alloc(4));
  {  // posdef[2] = 1 for good: the flag is re-armed by a device-to-device copy (steps may be enqueued ahead of the host)
    static const int one = 1;
    KB_CUDA(h, cudaMemcpyAsync(h->posdef.p + 2, &one, sizeof(int), cudaMemcpyHostToDevice, s));
  }
  KB_CUDA(h, h->lm_counters.alloc(KB_NUM_MODELS));
  KB_CUDA(h, cudaMemsetAsync(h->dx.p, 0, sizeof(double) * std::max<size_t>(1, (size_t)h->jcols), s));
  KB_CUDA(h, cudaMemsetAsync(h->VB.p, 0, sizeof(double) * std::max<size_t>(1, (size_t)n_views * VB_STRIDE), s));
  KB_CUDA(h, cudaMemsetAsync(h->e.p, 0, sizeof(double) * std::max<size_t>(1, 2 * (size_t)h->n_terms_local), s));
  KB_CUDA(h, cudaMemsetAsync(h->camA.p, 0, sizeof(double) * C * C * 36, s));
  D.y_u = h->y_u.p; D.y_v = h->y_v.p; D.corner = h->corner.p; D.target = h->target.p;
  h->front_u = h->y_u.p;
  h->front_v = h->y_v.p;
  h->back_u = h->back_v = nullptr;  // the double buffer (kb_prefetch_observations) is re-created at the new size on demand
  D.view_set = h->view_set.p; D.view_cam = h->view_cam.p; D.view_begin = h->view_begin.p; D.set_view = h->set_view.p;
  D.lin_off = h->lin_off.p; D.view_jbase = h->view_jbase.p; D.col_desc = h->col_desc.p;
  D.cam_params = h->cam_params.p; D.baselines = h->baselines.p; D.set_poses = h->set_poses.p;
  D.camT = h->camT.p; D.camPi = h->camPi.p; D.camA = h->camA.p; D.baseBt = h->baseBt.p; D.baseM = h->baseM.p;
  D.e = h->e.p; D.view_cost = h->view_cost.p; D.set_prep = h->set_prep.p; D.VB = h->VB.p; D.gram_partial = h->gram_partial.p; D.sumG = h->sumG.p;
  D.V = h->V.p; D.bv = h->bv.p; D.W = h->W.p; D.Lv = h->Lv.p; D.yv = h->yv.p;
  D.U = h->U.p; D.Sred = h->Sred.p; D.dxc = h->dxc.p; D.dx = h->dx.p; D.n_invalid = h->n_invalid.p; D.rho_partial = h->rho_partial.p; D.tickets = h->tickets.p;
  h->n_partials = schur_num_partials(D);
  KB_CUDA(h, h->partials.alloc(schur_partial_stride(D) * h->n_partials));
  return KB_OK;
}

kb_status kb_create(const kb_problem_desc* d, kb_handle** out) {
  CreateTrace trace;
  if (!d || !out) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (d->n_cams < 1 || d->n_cams > MAX_CAMS) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "n_cams out of range (1..32)");
  if (d->driver_order < 0 || d->driver_order > KB_ORDER_BATCH) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "unknown driver order");
  if (d->driver_order == KB_ORDER_SINGLE && d->n_cams != 1) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "single-camera order needs one camera");
  if (d->driver_order == KB_ORDER_STEREO && d->n_cams != 2) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "stereo order needs two cameras");
  if (d->n_target_points < 1 || d->n_target_points > 65535) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "n_target_points out of range");
  if (d->n_sets < 0 || d->n_views < 0 || d->n_terms < 0) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "negative size");
  if (d->n_ranks < 1 || d->rank < 0 || d->rank >= d->n_ranks) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "bad rank / n_ranks");
  for (int k = 0; k < d->n_cams; ++k)
    if (d->cam_model[k] < 0 || d->cam_model[k] >= KB_NUM_MODELS) return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "unknown camera model");

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
    return fail(nullptr, KB_ERR_NO_DEVICE, "no CUDA device: the B200 hot path has no CPU fallback");
  if (d->device < 0 || d->device >= n_dev) return fail(nullptr, KB_ERR_NO_DEVICE, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaSetDevice(d->device) != cudaSuccess || cudaGetDeviceProperties(&prop, d->device) != cudaSuccess)
    return fail(nullptr, KB_ERR_NO_DEVICE, "cannot select CUDA device");
  if (prop.major != 10) return fail(nullptr, KB_ERR_NO_DEVICE, "kernels are built for sm_100a only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));

  std::unique_ptr<kb_handle, void (*)(kb_handle*)> hp(new kb_handle(), kb_destroy);
  kb_handle* h = hp.get();
  std::memset(&h->d, 0, sizeof(h->d));
  h->device = d->device;
  h->driver_order = d->driver_order;
  h->n_cams = d->n_cams;
  const bool presharded = d->n_sets_total > 0;
  if (presharded && (d->set_offset < 0 || d->set_offset + d->n_sets > d->n_sets_total))
    return fail(nullptr, KB_ERR_INVALID_ARGUMENT, "pre-sharded set range outside [0, n_sets_total)");
  h->n_sets_global = presharded ? d->n_sets_total : d->n_sets;
  h->n_ranks = d->n_ranks;
  h->rank = d->rank;
  h->n_terms_global = presharded ? std::max<int64_t>(d->n_terms_total, d->n_terms) : d->n_terms;
  h->presharded = presharded;
  h->cam_model.assign(d->cam_model, d->cam_model + d->n_cams);
  if (presharded) { h->set_lo = d->set_offset; h->set_hi = d->set_offset + d->n_sets; }
  else shard_range(d->n_sets, d->n_ranks, d->rank, h->set_lo, h->set_hi);
  const int in_set_shift = presharded ? d->set_offset : 0;  // input view_set is local when pre-sharded
  auto cfail = [&](kb_status c, const std::string& m) { g_create_error = m; return c; };
#define KB_CCUDA(call)                                                                                        \
  do {                                                                                                        \
    cudaError_t _e = (call);                                                                                  \
    if (_e != cudaSuccess) return cfail(KB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));     \
  } while (0)
  KB_CCUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (d->n_ranks > MAX_CAMS) return cfail(KB_ERR_INVALID_ARGUMENT, "more than 32 ranks");
  KB_CCUDA(cudaMallocHost((void**)&h->h_scalars, sizeof(double) * (8 + 4 * MAX_CAMS)));
  KB_CCUDA(cudaMallocHost((void**)&h->h_posdef, sizeof(int) * 2));
  KB_CCUDA(cudaMallocHost((void**)&h->h_ctrl, sizeof(LmCtrl)));

  // every stream / event of the handle up front (re-used for the whole life of the handle, also across kb_append_set)
  KB_CCUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& e : h->ev_chunk) KB_CCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  KB_CCUDA(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
  KB_CCUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  for (auto& e : h->ev_join) KB_CCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& st : h->model_stream) KB_CCUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  KB_CCUDA(cudaEventCreateWithFlags(&h->ev_prefetch, cudaEventDisableTiming));
  KB_CCUDA(cudaEventCreateWithFlags(&h->ev_front_free, cudaEventDisableTiming));
  trace.mark("device, streams, events");
  // ---- local views / terms ----
  const int n_local_sets = h->set_hi - h->set_lo;
  std::vector<double> yu, yv;  // only filled when the local terms are not one contiguous run of the caller's arrays
  std::vector<uint16_t> corner;
  std::vector<int>& vs = h->h_view_set;
  std::vector<int>& vc = h->h_view_cam;
  std::vector<int>& vb = h->h_view_begin;
  vb.push_back(0);
  std::vector<char> seen((size_t)n_local_sets * d->n_cams, 0);
  std::vector<int64_t> src_begin;  // per local view: where its terms start in the caller's arrays
  bool contiguous = true;          // the local views' term ranges follow each other in the caller's arrays
  int64_t n_local_terms = 0;
  for (int w = 0; w < d->n_views; ++w) {
    const int k = d->view_cam[w];
    if (d->view_set[w] < 0 || d->view_set[w] >= d->n_sets || k < 0 || k >= d->n_cams) return cfail(KB_ERR_INVALID_ARGUMENT, "view index out of range");
    const int v = d->view_set[w] + in_set_shift;
    const int64_t b = d->view_begin[w], e = d->view_begin[w + 1];
    if (b < 0 || e < b || e > d->n_terms) return cfail(KB_ERR_INVALID_ARGUMENT, "view_begin is not a monotone partition of the terms");
    if (v < h->set_lo || v >= h->set_hi) continue;
    const int lv = v - h->set_lo;
    if (seen[(size_t)lv * d->n_cams + k]) return cfail(KB_ERR_INVALID_ARGUMENT, "two views for the same (set, camera)");
    seen[(size_t)lv * d->n_cams + k] = 1;
    vs.push_back(lv);
    vc.push_back(k);
    if (!src_begin.empty() && b != src_begin.back() + (n_local_terms - vb[vb.size() - 2])) contiguous = false;
    src_begin.push_back(b);
    n_local_terms += e - b;
    if (n_local_terms > (int64_t)0x7fffffff) return cfail(KB_ERR_INVALID_ARGUMENT, "more than 2^31 terms on one rank");
    vb.push_back((int)n_local_terms);
  }
  h->n_terms_local = n_local_terms;
  // the observations: validated and narrowed (corner ids) / gathered (measurements, only when needed) by a few host threads
  corner.resize((size_t)n_local_terms);
  if (!contiguous) {
    yu.resize((size_t)n_local_terms);
    yv.resize((size_t)n_local_terms);
  }
  {
    const int n_lv = (int)src_begin.size();
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)8, n_local_terms / 200000 + 1}));
    std::atomic<int> bad{0};
    auto work = [&](int t) {
      const int lo = (int)((int64_t)n_lv * t / n_thr), hi = (int)((int64_t)n_lv * (t + 1) / n_thr);
      for (int j = lo; j < hi; ++j) {
        const int64_t sb = src_begin[j];
        const int o = vb[j], cnt = vb[j + 1] - vb[j];
        const int32_t* cid = d->corner_id + sb;
        uint16_t* dst = corner.data() + o;
        int flag = 0;
        for (int i = 0; i < cnt; ++i) {
          flag |= (cid[i] < 0) | (cid[i] >= d->n_target_points);
          dst[i] = (uint16_t)cid[i];
        }
        if (flag) bad.store(1);
        if (!contiguous && cnt > 0) {
          std::memcpy(yu.data() + o, d->y_u + sb, sizeof(double) * (size_t)cnt);
          std::memcpy(yv.data() + o, d->y_v + sb, sizeof(double) * (size_t)cnt);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (bad.load()) return cfail(KB_ERR_INVALID_ARGUMENT, "corner_id out of range");
  }
  const double* src_u = contiguous ? (src_begin.empty() ? nullptr : d->y_u + src_begin[0]) : yu.data();
  const double* src_v = contiguous ? (src_begin.empty() ? nullptr : d->y_v + src_begin[0]) : yv.data();
  trace.mark("host copy of the structure");
  // ---- upload: observations, target, state ----
  cudaStream_t s = h->stream;
  h->d.n_target = d->n_target_points;
  std::vector<double> target(d->target_points, d->target_points + (size_t)3 * d->n_target_points);
  std::vector<double> cam(d->cam_params, d->cam_params + (size_t)KB_CAM_PARAM_STRIDE * d->n_cams);
  std::vector<double> base;
  if (d->n_cams > 1) base.assign(d->baselines, d->baselines + (size_t)KB_POSE_STRIDE * (d->n_cams - 1));
  std::vector<double> sets;
  if (n_local_sets > 0) {
    const double* sp = d->set_poses + (presharded ? 0 : (size_t)KB_POSE_STRIDE * h->set_lo);
    sets.assign(sp, sp + (size_t)KB_POSE_STRIDE * n_local_sets);
  }
  KB_CCUDA(h->y_u.alloc((size_t)n_local_terms));
  KB_CCUDA(h->y_v.alloc((size_t)n_local_terms));
  // the measurements go straight from the caller's (pageable) arrays when the local terms are one contiguous run of them; large
  // ones on two helper threads with a stream each (a pageable copy occupies its calling thread with the staging), while this thread
  // goes on with the allocations and tables below.  The guard joins them on every path out of this function, before the handle dies.
  struct UploadThreads {
    std::thread t[2];
    cudaError_t err[2] = {cudaSuccess, cudaSuccess};
    void join() { for (auto& x : t) if (x.joinable()) x.join(); }
    ~UploadThreads() { join(); }
  } up;
  if (n_local_terms > (int64_t)1 << 20) {
    const size_t bytes = sizeof(double) * (size_t)n_local_terms;
    const int dev = h->device;
    double *du = h->y_u.p, *dv = h->y_v.p;
    cudaStream_t s0 = h->copy_stream, s1 = h->model_stream[0];
    cudaError_t* er = up.err;
    up.t[0] = std::thread([=] {
      cudaError_t e = cudaSetDevice(dev);
      if (e == cudaSuccess) e = cudaMemcpyAsync(du, src_u, bytes, cudaMemcpyHostToDevice, s0);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
      er[0] = e;
    });
    up.t[1] = std::thread([=] {
      cudaError_t e = cudaSetDevice(dev);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dv, src_v, bytes, cudaMemcpyHostToDevice, s1);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s1);
      er[1] = e;
    });
  } else if (n_local_terms > 0) {
    KB_CCUDA(cudaMemcpyAsync(h->y_u.p, src_u, sizeof(double) * (size_t)n_local_terms, cudaMemcpyHostToDevice, s));
    KB_CCUDA(cudaMemcpyAsync(h->y_v.p, src_v, sizeof(double) * (size_t)n_local_terms, cudaMemcpyHostToDevice, s));
  }
  KB_CCUDA(h->corner.upload(corner, s));
  KB_CCUDA(h->target.upload(target, s));
  KB_CCUDA(h->cam_params.upload(cam, s));
  KB_CCUDA(h->baselines.upload(base, s));
  KB_CCUDA(h->set_poses.upload(sets, s));
  KB_CCUDA(h->init_cam.upload(cam, s));
  KB_CCUDA(h->init_base.upload(base, s));
  KB_CCUDA(h->init_sets.upload(sets, s));
  KB_CCUDA(h->bk_cam.upload(cam, s));
  KB_CCUDA(h->bk_base.upload(base, s));
  KB_CCUDA(h->bk_sets.upload(sets, s));
  trace.mark("uploads");
  // ---- one-time device objects ----
  KB_CCUDA(h->ctrl.alloc(1));
  KB_CCUDA(h->n_invalid.alloc(1));
  KB_CCUDA(cudaMemsetAsync(h->n_invalid.p, 0, sizeof(unsigned int), s));
  KB_CCUDA(h->tickets.alloc(4));
  KB_CCUDA(cudaMemsetAsync(h->tickets.p, 0, 4 * sizeof(unsigned int), s));
  KB_CCUDA(h->rank_slots.alloc(4 * (size_t)d->n_ranks));
  {
    LmCtrl neutral;
    std::memset(&neutral, 0, sizeof(neutral));
    neutral.need_build = 1;
    KB_CCUDA(cudaMemcpyAsync(h->ctrl.p, &neutral, sizeof(neutral), cudaMemcpyHostToDevice, s));  // pageable source: staged at the call
  }
  // ---- derived tables, work buffers, device problem view ----
  h->d.sT[0] = h->d.sT[3] = 1.0;  // invR = I, NoMEstimator (CalibrationTools.hpp:105-108; BE/src/ErrorTerm.cpp:8-12)
  {
    const kb_status st = build_tables(h);
    if (st != KB_OK) return cfail(st, h->error);
  }
  {
    DevProblem& D = h->d;
    const size_t NA = D.n_aug;
    D.px.enabled = 0;
    D.px.n_ranks = d->n_ranks;
    D.px.rank = d->rank;
    D.px.na2 = (int)((NA * NA + 1) & ~(size_t)1);
    if (d->n_ranks > 1 && d->n_ranks <= PX_MAX_RANKS) {
      KB_CCUDA(h->px_buf.alloc(px_doubles(D.px)));
      KB_CCUDA(cudaMemsetAsync(h->px_buf.p, 0, sizeof(double) * px_doubles(D.px), s));
      D.px.base[d->rank] = h->px_buf.p;
    }
  }
  trace.mark("allocations, memsets");

  if (d->n_ranks > 1) {
    std::string err;
    if (!g_nccl.load(err)) return cfail(KB_ERR_NCCL, err);
    if (!d->nccl_id) return cfail(KB_ERR_INVALID_ARGUMENT, "nccl_id is required when n_ranks > 1");
    NcclUniqueId id;
    std::memcpy(id.internal, d->nccl_id, 128);
    int r = g_nccl.CommInitRank(&h->comm, d->n_ranks, id, d->rank);
    if (r != 0) return cfail(KB_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  }
  up.join();
  KB_CCUDA(up.err[0]);
  KB_CCUDA(up.err[1]);
  KB_CCUDA(cudaStreamSynchronize(s));
  trace.mark("nccl, synchronise");
#undef KB_CCUDA
  *out = hp.release();
  return KB_OK;
}

int64_t kb_jrows(const kb_handle* h) { return 2 * h->n_terms_global; }
int64_t kb_local_jrows(const kb_handle* h) { return 2 * h->n_terms_local; }
int64_t kb_jcols(const kb_handle* h) { return h->jcols; }
int32_t kb_num_design_variables(const kb_handle* h) { return (int32_t)h->dv_dim.size(); }
kb_status kb_get_dv_layout(const kb_handle* h, int32_t* column_base, int32_t* dims) {
  for (size_t i = 0; i < h->dv_dim.size(); ++i) { column_base[i] = h->dv_col[i]; dims[i] = h->dv_dim[i]; }
  return KB_OK;
}
int64_t kb_kernel_launches(const kb_handle* h) { return h->launches; }
void* kb_cuda_stream(kb_handle* h) { return (void*)h->stream; }
kb_status kb_enable_stage_timing(kb_handle* h, int32_t on) {
  if (on) {  // the event ring is created on first use
    KB_CUDA(h, cudaSetDevice(h->device));
    for (auto& e : h->ev)
      if (!e) KB_CUDA(h, cudaEventCreate(&e));
  }
  h->timing = on != 0;
  h->timing_mask = on > 1 ? (unsigned)on >> 1 : ~0u;  // on > 1: bit (s + 1) selects stage s; 1: every stage
  if (on)
    for (int i = 0; i < KB_NUM_STAGES; ++i) { h->stage_total[i] = 0.0; h->stage_calls[i] = 0; h->stage_tail[i] = h->stage_head[i]; }
  return KB_OK;
}
kb_status kb_get_stage_totals(kb_handle* h, double* total_ms, int64_t* calls) {
  for (int i = 0; i < KB_NUM_STAGES; ++i) { total_ms[i] = h->stage_total[i]; calls[i] = h->stage_calls[i]; }
  return KB_OK;
}
kb_status kb_get_stage_ms(kb_handle* h, double* ms) {
  for (int i = 0; i < KB_NUM_STAGES; ++i) ms[i] = h->stage_ms[i];
  return KB_OK;
}
kb_status kb_set_solver_semantic(kb_handle* h, int32_t semantic) {
  if (semantic != 0 && semantic != 1) return fail(h, KB_ERR_INVALID_ARGUMENT, "semantic must be 0 (block_cholesky) or 1 (sparse_cholesky)");
  h->semantic = semantic;
  return KB_OK;
}

// fused linearise+assemble at the current state: view blocks, per-camera Gram sums, cost (-> scalars[slot]) and, if asked, e()
static kb_status run_linearise_assemble(kb_handle* h, bool write_e, int cost_slot) {
  StreamCtx c = ctx(h);
  {
    StageTimer t(h, 1);
    KB_CUDA(h, launch_linearise_assemble(h->d, h->vmeta.p, h->slices.p, h->slice_model_begin, write_e, true, true, c));
  }
  KB_CUDA(h, launch_finalize_gram(h->d, h->cam_slice_range.p, 1, h->scalars.p + cost_slot, write_e /* the evaluate-time call */, 0, nullptr, nullptr, c));
  h->la_version = h->state_version;
  h->la_rows = h->d.mest_rows;
  return KB_OK;
}

static kb_status finish_evaluate(kb_handle* h, double* out_cost) {
  KB_CUDA(h, cudaMemcpyAsync(h->h_scalars, h->scalars.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (h->defer_sync) return KB_OK;
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  collect_stages(h);
  if (out_cost) *out_cost = h->h_scalars[0];
  return KB_OK;
}

// useMEstimator of the reference only gates whether the policy weight scales e() and the Jacobian rows; the cost is weighted by
// the policy regardless (BE/src/ErrorTerm.cpp:19-24, BE/include/aslam/backend/implementation/ErrorTerm.hpp:97-109, 170-192)
static void set_rows_mode(kb_handle* h, int32_t use_m_estimator) { h->d.mest_rows = (h->d.mest_kind != 0 && use_m_estimator) ? 1 : 0; }

kb_status kb_evaluate_error(kb_handle* h, int32_t use_m_estimator, double* out_cost) {
  KB_CUDA(h, cudaSetDevice(h->device));
  set_rows_mode(h, use_m_estimator);
  // with a policy installed but useMEstimator off, the Gram block's e-column holds sum e^T invR e, not the weighted cost:
  // that combination runs the residual-only kernel, which accumulates w e^T invR e directly
  const bool speculative = h->speculative && !(h->d.mest_kind != 0 && !use_m_estimator);
  StreamCtx c = ctx(h);
  {
    StageTimer t(h, 0);
    if (speculative) {
      // The LM loop evaluates the cost of a trial state and, when the step is accepted, linearises at that same state
      // (Optimizer2.cpp:237 then LevenbergMarquardtTrustRegionPolicy.cpp:72): do both now, the Gram column of e gives the cost.
      kb_status st = run_linearise_assemble(h, true, 0);
      if (st != KB_OK) return st;
    } else {
      KB_CUDA(h, launch_prep(h->d, c));
      KB_CUDA(h, launch_evaluate(h->d, h->view_list.p, h->model_begin, h->scalars.p, c));
    }
    if (h->px_on && speculative) {  // gram_cost_kernel has already put this rank's cost into every rank's slot
      KB_CUDA(h, launch_px_combine_cost(h->d, h->scalars.p, 0, nullptr, nullptr, c));
    } else {
      kb_status st = nccl_allreduce(h, h->scalars.p, 1, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
    }
  }
  return finish_evaluate(h, out_cost);
}

// kb_set_observations + kb_evaluate_error with the upload overlapped: the observations travel in KB_STREAM_CHUNKS pieces on a
// copy stream and the fused kernel starts on each piece as soon as it has landed.  T = double, or float for measurements in the
// detector's own precision (half the bytes over PCIe; widened exactly on the device, on the copy stream).
}  // extern "C"
template <typename T>
static kb_status upload_chunk(kb_handle* h, const T* y_u, const T* y_v, int64_t lo, int64_t n, double* dst_u, double* dst_v, cudaStream_t cs);
template <>
kb_status upload_chunk<double>(kb_handle* h, const double* y_u, const double* y_v, int64_t lo, int64_t n, double* dst_u, double* dst_v, cudaStream_t cs) {
  KB_CUDA(h, cudaMemcpyAsync(dst_u + lo, y_u + lo, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
  KB_CUDA(h, cudaMemcpyAsync(dst_v + lo, y_v + lo, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
  return KB_OK;
}
template <>
kb_status upload_chunk<float>(kb_handle* h, const float* y_u, const float* y_v, int64_t lo, int64_t n, double* dst_u, double* dst_v, cudaStream_t cs) {
  KB_CUDA(h, cudaMemcpyAsync(h->stage_u.p + lo, y_u + lo, sizeof(float) * n, cudaMemcpyHostToDevice, cs));
  KB_CUDA(h, cudaMemcpyAsync(h->stage_v.p + lo, y_v + lo, sizeof(float) * n, cudaMemcpyHostToDevice, cs));
  KB_CUDA(h, launch_widen_observations(h->stage_u.p + lo, h->stage_v.p + lo, dst_u + lo, dst_v + lo, n, cs, &h->launches));
  return KB_OK;
}
template <typename T>
static kb_status ensure_stage(kb_handle*) { return KB_OK; }
template <>
kb_status ensure_stage<float>(kb_handle* h) {
  const size_t n = (size_t)std::max<int64_t>(h->n_terms_local, 1);
  if (h->stage_u.n < n) KB_CUDA(h, h->stage_u.alloc(n));
  if (h->stage_v.n < n) KB_CUDA(h, h->stage_v.alloc(n));
  return KB_OK;
}

template <typename T>
static kb_status set_observations_impl(kb_handle* h, const T* y_u, const T* y_v) {
  if (h->n_ranks != 1 && !h->presharded)
    return fail(h, KB_ERR_STATE, "kb_set_observations needs a single rank or a pre-sharded problem (terms are re-packed per rank otherwise)");
  if (!y_u || !y_v) return fail(h, KB_ERR_INVALID_ARGUMENT, "null observation array");
  KB_CUDA(h, cudaSetDevice(h->device));
  kb_status st = ensure_stage<T>(h);
  if (st != KB_OK) return st;
  if ((st = upload_chunk<T>(h, y_u, y_v, 0, h->n_terms_local, h->front_u, h->front_v, h->stream)) != KB_OK) return st;
  ++h->state_version;
  return KB_OK;
}

template <typename T>
static kb_status evaluate_error_streamed_impl(kb_handle* h, const T* y_u, const T* y_v, int32_t use_m_estimator, double* out_cost) {
  if (h->n_ranks != 1 && !h->presharded)
    return fail(h, KB_ERR_STATE, "kb_evaluate_error_streamed needs a single rank or a pre-sharded problem (terms are re-packed per rank otherwise)");
  if (!y_u || !y_v) return fail(h, KB_ERR_INVALID_ARGUMENT, "null observation array");
  if (!h->speculative || (h->d.mest_kind != 0 && !use_m_estimator)) {
    kb_status st = set_observations_impl<T>(h, y_u, y_v);
    return st != KB_OK ? st : kb_evaluate_error(h, use_m_estimator, out_cost);
  }
  KB_CUDA(h, cudaSetDevice(h->device));
  kb_status st = ensure_stage<T>(h);
  if (st != KB_OK) return st;
  set_rows_mode(h, use_m_estimator);
  StreamCtx c = ctx(h);
  ++h->state_version;
  // the copies may only overwrite the observations once everything queued so far has read them
  KB_CUDA(h, cudaEventRecord(h->ev_main, h->stream));
  KB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_main, 0));
  for (int k = 0; k < KB_STREAM_CHUNKS; ++k) {
    const int64_t lo = h->st_chunk_term[k], n = h->st_chunk_term[k + 1] - lo;
    if (n > 0 && (st = upload_chunk<T>(h, y_u, y_v, lo, n, h->front_u, h->front_v, h->copy_stream)) != KB_OK) return st;
    KB_CUDA(h, cudaEventRecord(h->ev_chunk[k], h->copy_stream));
  }
  {
    StageTimer t(h, 0);
    for (int k = 0; k < KB_STREAM_CHUNKS; ++k) {
      KB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[k], 0));
      KB_CUDA(h, launch_linearise_assemble(h->d, h->vmeta.p, h->st_slices.p, h->st_chunk_model_begin[k], true, k == 0, k == 0, c));
    }
    KB_CUDA(h, launch_finalize_gram(h->d, h->st_cam_slice_range.p, KB_STREAM_CHUNKS, h->scalars.p, true, 0, nullptr, nullptr, c));
    h->la_version = h->state_version;
    h->la_rows = h->d.mest_rows;
    if (h->px_on) {
      KB_CUDA(h, launch_px_combine_cost(h->d, h->scalars.p, 0, nullptr, nullptr, c));
    } else {
      kb_status st2 = nccl_allreduce(h, h->scalars.p, 1, kNcclFloat64, kNcclSum);
      if (st2 != KB_OK) return st2;
    }
  }
  return finish_evaluate(h, out_cost);
}
extern "C" {
kb_status kb_evaluate_error_streamed(kb_handle* h, const double* y_u, const double* y_v, int32_t use_m_estimator, double* out_cost) {
  return evaluate_error_streamed_impl<double>(h, y_u, y_v, use_m_estimator, out_cost);
}
kb_status kb_evaluate_error_streamed_f32(kb_handle* h, const float* y_u, const float* y_v, int32_t use_m_estimator, double* out_cost) {
  return evaluate_error_streamed_impl<float>(h, y_u, y_v, use_m_estimator, out_cost);
}

kb_status kb_build_system(kb_handle* h, int32_t use_m_estimator) {
  KB_CUDA(h, cudaSetDevice(h->device));
  set_rows_mode(h, use_m_estimator);
  StreamCtx c = ctx(h);
  if (h->la_version != h->state_version || h->la_rows != h->d.mest_rows) {
    kb_status st = run_linearise_assemble(h, false, 6);
    if (st != KB_OK) return st;
  }
  {
    StageTimer t(h, 2);
    KB_CUDA(h, launch_set_reduce(h->d, c));
  }
  h->diag_residual = 0.0;  // H.clear(false): BlockCholeskyLinearSystemSolver.cpp:64
  h->built = true;
  h->solved = false;
  return KB_OK;  // stage times are collected at the next synchronising call
}

// ---- weighting of the terms: inverse measurement covariance and M-estimator policy --------------------------------------
namespace {
// regularised lower incomplete gamma function P(a, x): power series below a + 1, continued fraction (modified Lentz) above
double gamma_p(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  const double pre = std::exp(a * std::log(x) - x - std::lgamma(a));
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term;
    for (int n = 1; n < 100000 && std::fabs(term) > 1e-18 * std::fabs(sum); ++n) {
      term *= x / (a + n);
      sum += term;
    }
    return pre * sum;
  }
  const double fpmin = 1e-290;
  double b = x + 1.0 - a, c = 1.0 / fpmin, d = 1.0 / b, f = d;
  for (int i = 1; i < 100000; ++i) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < fpmin) d = fpmin;
    c = b + an / c;
    if (std::fabs(c) < fpmin) c = fpmin;
    d = 1.0 / d;
    const double delta = d * c;
    f *= delta;
    if (std::fabs(delta - 1.0) < 1e-16) break;
  }
  return 1.0 - pre * f;
}
// chi-squared quantile: the x with P(df / 2, x / 2) = p.  The reference takes it from boost::math::quantile
// (BE/src/MEstimatorPolicies.cpp:116-119); here a bracketed Newton iteration from the Wilson-Hilferty guess.
double chi2_quantile(double p, double df) {
  const double a = 0.5 * df;
  // Wilson-Hilferty: x ~ df (1 - 2/(9 df) + z sqrt(2/(9 df)))^3, z the normal quantile (Acklam-free: bisection on erfc)
  double zlo = -40.0, zhi = 40.0;
  for (int i = 0; i < 200; ++i) {
    const double zm = 0.5 * (zlo + zhi);
    if (0.5 * std::erfc(-zm / std::sqrt(2.0)) < p) zlo = zm; else zhi = zm;
  }
  const double z = 0.5 * (zlo + zhi), k = 2.0 / (9.0 * df);
  double x = df * std::pow(std::max(1.0 - k + z * std::sqrt(k), 1e-3), 3.0);
  double lo = 0.0, hi = std::max(2.0 * x, 4.0 * df + 16.0);
  while (gamma_p(a, 0.5 * hi) < p) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    const double F = gamma_p(a, 0.5 * x) - p;
    if (F < 0.0) lo = x; else hi = x;
    const double pdf = 0.5 * std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a));
    double nx = (pdf > 0.0) ? x - F / pdf : 0.5 * (lo + hi);
    if (!(nx > lo && nx < hi)) nx = 0.5 * (lo + hi);
    if (std::fabs(nx - x) <= 1e-16 * std::fabs(x)) { x = nx; break; }
    x = nx;
  }
  return x;
}
void weighting_changed(kb_handle* h) {
  const double* S = h->sqrt_inv_r;
  h->d.sT[0] = S[0]; h->d.sT[1] = S[2]; h->d.sT[2] = S[1]; h->d.sT[3] = S[3];
  const bool identity = S[0] == 1.0 && S[1] == 0.0 && S[2] == 0.0 && S[3] == 1.0;
  h->d.weighted = (!identity || h->d.mest_kind != 0) ? 1 : 0;
  ++h->state_version;  // every cached linearisation is stale
  h->built = h->solved = false;
  if (h->lm_graph) { cudaGraphExecDestroy(h->lm_graph); h->lm_graph = nullptr; }  // captured with the old kernel arguments
}
}  // namespace

kb_status kb_set_inv_r(kb_handle* h, const double inv_r[4]) {
  if (!inv_r) return fail(h, KB_ERR_INVALID_ARGUMENT, "null invR");
  const double a00 = inv_r[0], a01 = inv_r[1], a10 = inv_r[2], a11 = inv_r[3];
  if (!(std::fabs(a01 - a10) <= 1e-12 * (std::fabs(a00) + std::fabs(a11))) || !(a00 > 0.0) || !(a11 > 0.0) || !(a00 * a11 - a01 * a10 > 0.0))
    return fail(h, KB_ERR_INVALID_ARGUMENT, "invR must be symmetric positive definite");
  // sm::eigen::computeMatrixSqrt (Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40): S = P^T L sqrt(D) of the
  // pivoted LDL^T (the larger diagonal entry is eliminated first, the first one on ties; the lower triangle is read)
  const bool swap = std::fabs(a11) > std::fabs(a00);
  const double d0 = swap ? a11 : a00, dd = swap ? a00 : a11;
  const double l10 = a10 / d0;
  const double d1 = dd - l10 * (d0 * l10);
  const double s0 = std::sqrt(d0), s1 = std::sqrt(d1);
  double* S = h->sqrt_inv_r;
  if (!swap) { S[0] = s0; S[1] = 0.0; S[2] = l10 * s0; S[3] = s1; }
  else { S[0] = l10 * s0; S[1] = s1; S[2] = s0; S[3] = 0.0; }
  for (int i = 0; i < 4; ++i) h->inv_r[i] = inv_r[i];
  weighting_changed(h);
  return KB_OK;
}
kb_status kb_get_sqrt_inv_r(const kb_handle* h, double out[4]) {
  for (int i = 0; i < 4; ++i) out[i] = h->sqrt_inv_r[i];
  return KB_OK;
}

kb_status kb_set_m_estimator(kb_handle* h, int32_t kind, double p0, double p1, double p2) {
  double prm = 0.0;
  switch (kind) {
    case KB_MEST_NONE: break;
    case KB_MEST_HUBER:
    case KB_MEST_CAUCHY:
    case KB_MEST_GEMAN_MCCLURE:
      if (!(p0 > 0.0)) return fail(h, KB_ERR_INVALID_ARGUMENT, "the M-estimator parameter must be positive");
      prm = p0;
      break;
    case KB_MEST_BLAKE_ZISSERMAN: {  // BlakeZissermanMEstimator(df, pCut, wCut): MEstimatorPolicies.cpp:80-86, 116-124
      const double df = p0, p_cut = p1, w_cut = p2;
      if (!(df >= 1.0) || !(p_cut > 0.0 && p_cut < 1.0) || !(w_cut > 0.0 && w_cut < 1.0))
        return fail(h, KB_ERR_INVALID_ARGUMENT, "Blake-Zisserman needs df >= 1, 0 < pCut < 1, 0 < wCut < 1");
      prm = (1.0 - w_cut) / w_cut * std::exp(-chi2_quantile(p_cut, std::floor(df)));
      break;
    }
    default: return fail(h, KB_ERR_INVALID_ARGUMENT, "unknown M-estimator kind");
  }
  h->d.mest_kind = kind;
  h->d.mest_param = prm;
  h->d.mest_rows = kind != KB_MEST_NONE ? 1 : 0;
  weighting_changed(h);
  return KB_OK;
}
double kb_m_estimator_parameter(const kb_handle* h) { return h->d.mest_param; }

// ---- reprojection statistics ≙ CameraCalibrator::PrintReprojectionErrorStatistics (K2/include/kalibr2/CameraCalibrator.hpp:368-405)
kb_status kb_reprojection_statistics(kb_handle* h, double* out) {
  if (!out) return fail(h, KB_ERR_INVALID_ARGUMENT, "null output");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  if (h->stats_e.n < (size_t)2 * std::max<int64_t>(h->n_terms_local, 1)) KB_CUDA(h, h->stats_e.alloc((size_t)2 * std::max<int64_t>(h->n_terms_local, 1)));
  if (h->stats_acc.n < (size_t)8 * h->n_cams) KB_CUDA(h, h->stats_acc.alloc((size_t)8 * h->n_cams));
  // raw residuals y - y_hat at the current state (getMeasurement() - getPredictedMeasurement(), CameraCalibrator.hpp:267-286):
  // the residual-only kernel, unweighted, into a scratch vector so that e() keeps what the last evaluation wrote
  DevProblem q = h->d;
  q.weighted = 0;
  q.e = h->stats_e.p;
  KB_CUDA(h, launch_prep(q, c));
  KB_CUDA(h, launch_evaluate(q, h->view_list.p, h->model_begin, h->scalars.p + 7, c));
  KB_CUDA(h, launch_reproj_stats(q, h->stats_e.p, h->cam_view_list.p, h->cam_view_begin.p, 0, h->stats_acc.p, c));
  if (h->n_ranks > 1) {  // make (n, sum e_u, sum e_v) global before the second pass
    kb_status st = nccl_allreduce(h, h->stats_acc.p, (size_t)8 * h->n_cams, kNcclFloat64, kNcclSum);
    if (st != KB_OK) return st;
  }
  KB_CUDA(h, launch_reproj_stats(q, h->stats_e.p, h->cam_view_list.p, h->cam_view_begin.p, 1, h->stats_acc.p, c));
  std::vector<double> acc((size_t)8 * h->n_cams, 0.0), acc1;
  KB_CUDA(h, cudaMemcpyAsync(acc.data(), h->stats_acc.p, sizeof(double) * acc.size(), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->n_ranks > 1) {  // the squared deviations: sum over ranks (slots 0..3 were already global: keep one copy)
    acc1 = acc;
    for (int k = 0; k < h->n_cams; ++k) for (int i = 0; i < 4; ++i) acc1[8 * k + i] = 0.0;
    KB_CUDA(h, cudaMemcpyAsync(h->stats_acc.p, acc1.data(), sizeof(double) * acc1.size(), cudaMemcpyHostToDevice, h->stream));
    kb_status st = nccl_allreduce(h, h->stats_acc.p, (size_t)8 * h->n_cams, kNcclFloat64, kNcclSum);
    if (st != KB_OK) return st;
    KB_CUDA(h, cudaMemcpyAsync(acc1.data(), h->stats_acc.p, sizeof(double) * acc1.size(), cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < h->n_cams; ++k) for (int i = 4; i < 8; ++i) acc[8 * k + i] = acc1[8 * k + i];
  }
  for (int k = 0; k < h->n_cams; ++k) {
    const double* a = &acc[8 * k];
    double* o = out + (size_t)k * KB_REPROJ_STAT_STRIDE;
    const double n = a[0];
    o[0] = n;
    o[1] = n > 0 ? a[1] / n : 0.0;
    o[2] = n > 0 ? a[2] / n : 0.0;
    o[3] = n > 1 ? std::sqrt(a[4] / (n - 1.0)) : 0.0;  // sample standard deviation; zero below two samples (:384-397)
    o[4] = n > 1 ? std::sqrt(a[5] / (n - 1.0)) : 0.0;
    o[5] = n > 0 ? std::sqrt(a[1] * a[1] + a[2] * a[2]) / std::sqrt(n) : 0.0;  // "RMSE" as printed: |sum of errors| / sqrt(n) (:404)
  }
  return KB_OK;
}

// ---- initial-guess stage -----------------------------------------------------------------------------
namespace {
kb_status pnp_prepare(kb_handle* h, const int32_t* resolution, const int** res_dev) {
  // a view is one image of the target: it cannot hold more corners than the target has (the kernels stage n_target corners per view)
  for (size_t w = 0; w + 1 < h->h_view_begin.size(); ++w)
    if (h->h_view_begin[w + 1] - h->h_view_begin[w] > h->d.n_target)
      return fail(h, KB_ERR_INVALID_ARGUMENT, "a view lists more corners than the target has points");
  const size_t V = (size_t)std::max(h->d.n_views, 1);
  if (h->pnp_T.n < V * POSE_STRIDE) KB_CUDA(h, h->pnp_T.alloc(V * POSE_STRIDE));
  if (h->pnp_ok.n < V) KB_CUDA(h, h->pnp_ok.alloc(V));
  *res_dev = nullptr;
  if (resolution) {
    std::vector<int> r(resolution, resolution + 2 * h->n_cams);
    KB_CUDA(h, h->pnp_res.upload(r, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));  // r is a stack-lifetime staging buffer
    *res_dev = h->pnp_res.p;
  }
  return KB_OK;
}
void host_quat2r(const double* q, double R[9]) {  // sm quat2r, row-major
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = x * x - y * y - z * z + w * w; R[1] = 2 * x * y + 2 * z * w; R[2] = 2 * x * z - 2 * y * w;
  R[3] = 2 * x * y - 2 * z * w; R[4] = -x * x + y * y - z * z + w * w; R[5] = 2 * x * w + 2 * y * z;
  R[6] = 2 * x * z + 2 * y * w; R[7] = -2 * x * w + 2 * y * z; R[8] = -x * x - y * y + z * z + w * w;
}
void host_r2quat(const double R[9], double q[4]) {  // sm r2quat (quaternion_algebra.cpp:16-75)
  const double c1 = R[0], c2 = R[3], c3 = R[6], c4 = R[1], c5 = R[4], c6 = R[7], c7 = R[2], c8 = R[5], c9 = R[8];
  const double dc[4] = {std::fabs(1.0 + c1 - c5 - c9), std::fabs(1.0 - c1 + c5 - c9), std::fabs(1.0 - c1 - c5 + c9), std::fabs(1.0 + c1 + c5 + c9)};
  int m = 0;
  for (int i = 1; i < 4; ++i) if (dc[i] > dc[m]) m = i;
  double c;
  if (m == 0) { q[0] = 0.5 * std::sqrt(dc[0]); c = 0.25 / q[0]; q[1] = c * (c4 + c2); q[2] = c * (c7 + c3); q[3] = c * (c8 - c6); }
  else if (m == 1) { q[1] = 0.5 * std::sqrt(dc[1]); c = 0.25 / q[1]; q[0] = c * (c4 + c2); q[2] = c * (c6 + c8); q[3] = c * (c3 - c7); }
  else if (m == 2) { q[2] = 0.5 * std::sqrt(dc[2]); c = 0.25 / q[2]; q[0] = c * (c3 + c7); q[1] = c * (c6 + c8); q[3] = c * (c4 - c2); }
  else { q[3] = 0.5 * std::sqrt(dc[3]); c = 0.25 / q[3]; q[0] = c * (c8 - c6); q[1] = c * (c3 - c7); q[2] = c * (c4 - c2); }
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
}
double upper_median(std::vector<double> v) {  // kalibr2::math::median (K2/src/BasicMathUtils.cpp:10-17)
  std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
  return v[v.size() / 2];
}
}  // namespace

kb_status kb_estimate_transformations(kb_handle* h, const int32_t* resolution, double* T_t_c, int32_t* ok) {
  KB_CUDA(h, cudaSetDevice(h->device));
  const int* res_dev = nullptr;
  kb_status st = pnp_prepare(h, resolution, &res_dev);
  if (st != KB_OK) return st;
  StreamCtx c = ctx(h);
  KB_CUDA(h, launch_estimate_transformations(h->d, h->view_list.p, h->model_begin, nullptr, res_dev, h->pnp_T.p, h->pnp_ok.p, c));
  const size_t V = (size_t)h->d.n_views;
  if (T_t_c && V) KB_CUDA(h, cudaMemcpyAsync(T_t_c, h->pnp_T.p, sizeof(double) * V * POSE_STRIDE, cudaMemcpyDeviceToHost, h->stream));
  if (ok && V) KB_CUDA(h, cudaMemcpyAsync(ok, h->pnp_ok.p, sizeof(int32_t) * V, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

kb_status kb_initialize_set_poses(kb_handle* h, const int32_t* resolution, int32_t* n_failed) {
  KB_CUDA(h, cudaSetDevice(h->device));
  const int* res_dev = nullptr;
  kb_status st = pnp_prepare(h, resolution, &res_dev);
  if (st != KB_OK) return st;
  const size_t V = (size_t)std::max(h->d.n_views, 1), S = (size_t)std::max(h->d.n_sets, 1);
  if (h->pnp_mask.n < V) KB_CUDA(h, h->pnp_mask.alloc(V));
  if (h->pnp_set_ok.n < S) KB_CUDA(h, h->pnp_set_ok.alloc(S));
  StreamCtx c = ctx(h);
  KB_CUDA(h, cudaMemsetAsync(h->pnp_mask.p, 0, V, h->stream));
  KB_CUDA(h, launch_best_view_mask(h->d, h->pnp_mask.p, c));
  KB_CUDA(h, launch_estimate_transformations(h->d, h->view_list.p, h->model_begin, h->pnp_mask.p, res_dev, h->pnp_T.p, h->pnp_ok.p, c));
  KB_CUDA(h, launch_set_pose_guess(h->d, h->pnp_T.p, h->pnp_ok.p, h->set_poses.p, h->pnp_set_ok.p, c));
  // the guesses are the state AND what kb_reset_state returns to
  if (h->set_poses.n) KB_CUDA(h, cudaMemcpyAsync(h->init_sets.p, h->set_poses.p, sizeof(double) * h->set_poses.n, cudaMemcpyDeviceToDevice, h->stream));
  std::vector<int> sok((size_t)h->d.n_sets, 0);
  if (!sok.empty()) KB_CUDA(h, cudaMemcpyAsync(sok.data(), h->pnp_set_ok.p, sizeof(int) * sok.size(), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  int failed = 0;
  for (int v : sok) failed += v ? 0 : 1;
  if (n_failed) *n_failed = failed;
  ++h->state_version;
  h->built = h->solved = h->has_backup = false;
  return KB_OK;
}

kb_status kb_estimate_stereo_baseline(kb_handle* h, const int32_t* resolution, int32_t cam_l, int32_t cam_h, double* baseline, int32_t* n_used) {
  if (h->n_ranks != 1) return fail(h, KB_ERR_STATE, "kb_estimate_stereo_baseline needs a single rank (the median runs over all sets)");
  if (cam_l < 0 || cam_h < 0 || cam_l >= h->n_cams || cam_h >= h->n_cams || cam_l == cam_h || !baseline)
    return fail(h, KB_ERR_INVALID_ARGUMENT, "bad camera pair");
  const size_t V = (size_t)h->d.n_views;
  std::vector<double> T(V * POSE_STRIDE);
  std::vector<int32_t> ok(V);
  kb_status st = kb_estimate_transformations(h, resolution, T.data(), ok.data());
  if (st != KB_OK) return st;
  std::vector<int> vl((size_t)h->n_sets_global, -1), vh((size_t)h->n_sets_global, -1);
  for (size_t w = 0; w < V; ++w) {
    if (h->h_view_cam[w] == cam_l) vl[h->h_view_set[w]] = (int)w;
    if (h->h_view_cam[w] == cam_h) vh[h->h_view_set[w]] = (int)w;
  }
  std::vector<double> tr[3], rv[3];
  for (int s = 0; s < h->n_sets_global; ++s) {
    if (vl[s] < 0 || vh[s] < 0 || !ok[vl[s]] || !ok[vh[s]]) continue;  // CalibrationTools.hpp:198-214
    const double* L = &T[(size_t)vl[s] * POSE_STRIDE];
    const double* H = &T[(size_t)vh[s] * POSE_STRIDE];
    double Rl[9], Rh[9], C[9], t[3];
    host_quat2r(L, Rl);
    host_quat2r(H, Rh);
    // T_H^-1 T_L: C = Rh^T Rl, t = Rh^T (t_L - t_H)
    const double d[3] = {L[4] - H[4], L[5] - H[5], L[6] - H[6]};
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) C[i * 3 + j] = Rh[0 * 3 + i] * Rl[0 * 3 + j] + Rh[1 * 3 + i] * Rl[1 * 3 + j] + Rh[2 * 3 + i] * Rl[2 * 3 + j];
      t[i] = Rh[0 * 3 + i] * d[0] + Rh[1 * 3 + i] * d[1] + Rh[2 * 3 + i] * d[2];
    }
    // RotationVector::rotationMatrixToParameters (Schweizer-Messer/sm_kinematics/src/RotationVector.cpp:58-80)
    const double trc = std::max(-1.0, std::min((C[0] + C[4] + C[8] - 1.0) * 0.5, 1.0));
    const double a = std::acos(trc);
    double pr[3] = {0.0, 0.0, 0.0};
    if (std::fabs(a) >= 1e-14) {
      const double px = C[7] - C[5], py = C[2] - C[6], pz = C[3] - C[1];
      const double n2 = std::sqrt(px * px + py * py + pz * pz);
      if (std::fabs(n2) >= 1e-14) { const double sc = -a / n2; pr[0] = sc * px; pr[1] = sc * py; pr[2] = sc * pz; }
    }
    for (int i = 0; i < 3; ++i) { tr[i].push_back(t[i]); rv[i].push_back(pr[i]); }
  }
  if (n_used) *n_used = (int32_t)tr[0].size();
  if (tr[0].empty()) return fail(h, KB_ERR_STATE, "no synced set was seen by both cameras");  // the reference's median throws
  double mt[3], mr[3];
  for (int i = 0; i < 3; ++i) { mt[i] = upper_median(tr[i]); mr[i] = upper_median(rv[i]); }
  // RotationVector::parametersToRotationMatrix (RotationVector.cpp:10-55)
  double C[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double angle = std::sqrt(mr[0] * mr[0] + mr[1] * mr[1] + mr[2] * mr[2]);
  if (angle >= 1e-14) {
    const double ax = mr[0] / angle, ay = mr[1] / angle, az = mr[2] / angle, sa = std::sin(angle), ca = std::cos(angle);
    const double m[9] = {ax * ax + ca * (1 - ax * ax), ax * ay - ca * ax * ay + sa * az, ax * az - ca * ax * az - sa * ay,
                         ax * ay - ca * ax * ay - sa * az, ay * ay + ca * (1 - ay * ay), ay * az - ca * ay * az + sa * ax,
                         ax * az - ca * ax * az + sa * ay, ay * az - ca * ay * az - sa * ax, az * az + ca * (1 - az * az)};
    for (int i = 0; i < 9; ++i) C[i] = m[i];
  }
  host_r2quat(C, baseline);
  for (int i = 0; i < 3; ++i) baseline[4 + i] = mt[i];
  return KB_OK;
}

kb_status kb_initialize_intrinsics(kb_handle* h, int32_t cam, int32_t target_rows, int32_t target_cols, const int32_t* resolution,
                                   double fallback_focal_length, double* params, int32_t* success) {
  if (h->n_ranks != 1) return fail(h, KB_ERR_STATE, "kb_initialize_intrinsics needs a single rank (it looks at every view of the camera)");
  if (cam < 0 || cam >= h->n_cams || !resolution) return fail(h, KB_ERR_INVALID_ARGUMENT, "bad camera index or null resolution");
  if (target_rows < 1 || target_cols < 1 || target_rows * target_cols != h->d.n_target || target_cols > 32 || target_rows > 1024)
    return fail(h, KB_ERR_INVALID_ARGUMENT, "target_rows x target_cols must equal the number of target points (cols <= 32)");
  for (size_t w = 0; w + 1 < h->h_view_begin.size(); ++w)
    if (h->h_view_begin[w + 1] - h->h_view_begin[w] > h->d.n_target)
      return fail(h, KB_ERR_INVALID_ARGUMENT, "a view lists more corners than the target has points");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  const int model = h->cam_model[cam];
  const int ru = resolution[2 * cam], rv = resolution[2 * cam + 1];
  const double cu = (ru - 1.0) / 2.0, cv = (rv - 1.0) / 2.0;  // "initialize the image center at the center of the image"
  int n_views = 0, first = 0;
  for (size_t w = 0; w < h->h_view_cam.size(); ++w) n_views += h->h_view_cam[w] == cam;
  for (int k = 0; k < cam; ++k)
    for (size_t w = 0; w < h->h_view_cam.size(); ++w) first += h->h_view_cam[w] == k;  // offset into the camera-major view list
  const int* cam_views = h->cam_view_list.p + first;
  if (h->init_out.n < 4) KB_CUDA(h, h->init_out.alloc(4));
  double res4[4] = {0, 0, 0, 0};
  double prm[KB_CAM_PARAM_STRIDE] = {0};
  bool ok = false, write = false;
  const bool pinhole = model == KB_PINHOLE_RADTAN || model == KB_PINHOLE_EQUI || model == KB_PINHOLE_FOV;
  if (pinhole) {
    const int rr = std::min(target_rows, target_cols), n_pairs = rr * (rr - 1) / 2;
    const size_t need = (size_t)std::max(n_views, 1) * std::max(n_pairs, 1);
    if (h->init_scratch.n < need) KB_CUDA(h, h->init_scratch.alloc(need));
    KB_CUDA(h, launch_focal_guesses(h->d, cam_views, n_views, target_rows, target_cols, h->init_scratch.p, h->init_out.p, c));
    KB_CUDA(h, cudaMemcpyAsync(res4, h->init_out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    double f0 = res4[0];
    ok = res4[1] > 0.0;
    if (!ok && fallback_focal_length > 0.0) { f0 = fallback_focal_length; ok = true; }  // PinholeProjection.hpp:781-791: returns true
    if (ok) { prm[0] = prm[1] = f0; prm[2] = cu; prm[3] = cv; write = true; }
  } else {
    const size_t need = (size_t)2 * std::max(n_views, 1) * target_rows;
    if (h->init_scratch.n < need) KB_CUDA(h, h->init_scratch.alloc(need));
    KB_CUDA(h, launch_omni_candidates(h->d, cam_views, n_views, target_rows, target_cols, ru, rv, h->init_scratch.p, h->init_out.p, c));
    KB_CUDA(h, cudaMemcpyAsync(res4, h->init_out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    double gamma0 = res4[0];
    ok = res4[1] > 0.0;
    const bool have_gamma = ok || fallback_focal_length > 0.0;
    if (!ok && fallback_focal_length > 0.0) gamma0 = fallback_focal_length;  // OmniProjection.hpp:826-841: parameters set, still returns false
    if (model == KB_OMNI_RADTAN || model == KB_OMNI_NONE) {
      if (have_gamma) { prm[0] = 1.0; prm[1] = prm[2] = gamma0; prm[3] = cu; prm[4] = cv; write = true; }
    } else if (ok) {  // EUCM / double sphere adopt the omni result only on success (ExtendedUnifiedProjection.hpp:741-757, DoubleSphereProjection.hpp:792-808)
      if (model == KB_EUCM_NONE) { prm[0] = 0.5; prm[1] = 1.0; }
      else { prm[0] = 0.0; prm[1] = 0.5; }
      prm[2] = prm[3] = 0.5 * gamma0; prm[4] = cu; prm[5] = cv;
      write = true;
    }
  }
  if (write) {  // the reference mutates the geometry in place; the guess also becomes what kb_reset_state returns to
    const size_t off = (size_t)cam * CAM_PARAM_STRIDE;
    KB_CUDA(h, cudaMemcpyAsync(h->cam_params.p + off, prm, sizeof(prm), cudaMemcpyHostToDevice, h->stream));
    KB_CUDA(h, cudaMemcpyAsync(h->init_cam.p + off, prm, sizeof(prm), cudaMemcpyHostToDevice, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    ++h->state_version;
    h->built = h->solved = h->has_backup = false;
    if (params) std::memcpy(params, prm, sizeof(prm));
  } else if (params) {
    std::vector<double> cur((size_t)h->n_cams * CAM_PARAM_STRIDE);
    KB_CUDA(h, cudaMemcpyAsync(cur.data(), h->cam_params.p, sizeof(double) * cur.size(), cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    std::memcpy(params, &cur[(size_t)cam * CAM_PARAM_STRIDE], sizeof(prm));
  }
  if (success) *success = ok ? 1 : 0;
  return KB_OK;
}

// ---- marginal analysis of the calibration block ----------------------------------------------------------------------
void kb_default_marginal_options(kb_marginal_options* o) {
  o->eps_svd = 2.220446049250313e-16;  // std::numeric_limits<double>::epsilon(): LinearSolverOptions.cpp:33
  o->svd_tol = -1.0;
}

// warm start of the eigen-decomposition of system kind `kind`: the vectors of the previous one, or null (first call, every 32nd call,
// KB_EIG_NO_WARM set).  Allocates the keep buffer.
static const double* eig_warm_start(kb_handle* h, int kind) {
  const size_t n = (size_t)h->d.n_c;
  if (h->eig_warm[kind].n != n * n) {
    if (h->eig_warm[kind].alloc(n * n) != cudaSuccess) return nullptr;
    h->eig_warm_age[kind] = -1;
  }
  static const bool disabled = getenv("KB_EIG_NO_WARM") != nullptr;
  // measured inside the estimator loop (profiles/r02_estimator_timing.md): from the previous vectors the Jacobi iteration needs 4-8
  // sweeps (scaled system) / 10-13 (unscaled), which beats QL + polish while the columns live in shared memory (0.13 ms per sweep at
  // n = 106) and loses once they spill to global memory (2.3 ms per sweep at n = 218): warm starts only for systems that fit
  const size_t np = (n + 1) & ~(size_t)1;
  const bool fits_smem = sizeof(double) * 2 * np * n <= 220 * 1024;
  const bool warm = !disabled && fits_smem && h->eig_warm_age[kind] >= 0 && h->eig_warm_age[kind] < 31;
  h->eig_warm_age[kind] = warm ? h->eig_warm_age[kind] + 1 : 0;
  return warm ? h->eig_warm[kind].p : nullptr;
}

static kb_status analyze_marginal_impl(kb_handle* h, const kb_marginal_options* o, kb_marginal_result* out, double* singular_values, double* V,
                                       int32_t* columns, bool rebuild) {
  if (!o || !out || !singular_values) return fail(h, KB_ERR_INVALID_ARGUMENT, "null argument");
  if (!rebuild && !h->built) return fail(h, KB_ERR_STATE, "kb_analyze_marginal_last_build called before any kb_build_system");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  const int n = h->d.n_c;
  if (n > 255) return fail(h, KB_ERR_INVALID_ARGUMENT, "calibration block too large");
  if (h->eig_G.n != (size_t)(n + 1) * n) {
    KB_CUDA(h, h->eig_G.alloc((size_t)(n + 1) * n));
    KB_CUDA(h, h->eig_V.alloc((size_t)(n + 1) * n));
    KB_CUDA(h, h->eig_sv.alloc(n));
    KB_CUDA(h, h->eig_Vout.alloc((size_t)n * n));
    KB_CUDA(h, h->eig_Vtmp.alloc((size_t)n * n));
    KB_CUDA(h, h->eig_sweeps.alloc(8));
  }
  // the undamped normal equations at the current state (or of the last build), set poses eliminated: exactly the analyzeMarginal matrix
  kb_status st = rebuild ? kb_build_system(h, 1) : KB_OK;
  if (st != KB_OK) return st;
  KB_CUDA(h, cudaMemcpyAsync(h->posdef.p, h->posdef.p + 2, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  KB_CUDA(h, launch_schur(h->d, 0.0, h->partials.p, h->n_partials, h->posdef.p, c));
  KB_CUDA(h, launch_schur_finalize(h->d, 0.0, h->partials.p, h->n_partials, true, c));
  if (h->px_on) {
    KB_CUDA(h, launch_px_reduce_system(h->d, c));
  } else if ((st = nccl_allreduce(h, h->Sred.p, (size_t)h->d.n_aug * h->d.n_aug, kNcclFloat64, kNcclSum)) != KB_OK) {
    return st;
  }
  if (h->n_ranks > 1 && (st = nccl_allreduce(h, h->posdef.p, 1, kNcclInt32, kNcclMin)) != KB_OK) return st;
  {
    StageTimer t(h, 4);  // reported as "reduced_solve": the dense stage of this entry point
    const double* warm = eig_warm_start(h, 0);
    KB_CUDA(h, launch_marginal_eig(h->d, h->eig_G.p, h->eig_V.p, h->eig_sv.p, h->eig_Vout.p, h->eig_Vtmp.p, h->eig_sweeps.p, warm, h->eig_warm[0].p, c));
  }
  KB_CUDA(h, cudaMemcpyAsync(h->h_posdef, h->posdef.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaMemcpyAsync(singular_values, h->eig_sv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (V) KB_CUDA(h, cudaMemcpyAsync(V, h->eig_Vout.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
  int sweeps2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  KB_CUDA(h, cudaMemcpyAsync(sweeps2, h->eig_sweeps.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  collect_stages(h);
  const int sweeps = sweeps2[1] ? 99 : sweeps2[0];
  if (getenv("KB_SVD_TRACE")) std::fprintf(stderr, "[kb trace] marginal analysis: n = %d, Jacobi polish sweeps = %d, QL failed = %d, tridiag %d kcycles, QL %d kcycles, %d QL steps, %d rotations\n", n, sweeps2[0], sweeps2[1], sweeps2[2], sweeps2[3], sweeps2[4], sweeps2[5]);
  h->solved = false;  // the pose factors now belong to the undamped system
  if (!h->h_posdef[0])
    return fail(h, KB_ERR_NUMERICAL, "a set pose is not constrained by its observations (pose block not positive definite): the marginal is undefined");
  if (sweeps >= 40) return fail(h, KB_ERR_NUMERICAL, "the eigen-decomposition of the marginal analysis did not converge");
  if (columns)
    for (int i = 0; i < n; ++i) columns[i] = h->h_cam_cols[i];
  // IC/src/algorithms/linalg.cpp:244-282, IC/src/core/LinearSolver.cpp:196-200
  out->n = n;
  out->tolerance = o->svd_tol != -1.0 ? o->svd_tol : singular_values[0] * o->eps_svd * n;
  int rank = n;
  for (int i = n - 1; i > 0; --i) {
    if (singular_values[i] > out->tolerance) break;
    --rank;
  }
  out->rank = rank;
  out->rank_deficiency = n - rank;
  out->sv_gap = rank < n ? singular_values[rank - 1] / singular_values[rank] : INFINITY;
  double lg = 0.0;
  for (int i = 0; i < rank; ++i) lg += std::log(singular_values[i]);
  out->sv_log2_sum = lg / std::log(2.0);
  return KB_OK;
}

// ---- peer exchange over NVLink ------------------------------------------------------------------------------------
kb_status kb_peer_exchange_handle(kb_handle* h, char out[64]) {
  if (!h->px_buf.p) return fail(h, KB_ERR_STATE, "no peer-exchange buffer (n_ranks must be 2..8)");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t ipc;
  KB_CUDA(h, cudaIpcGetMemHandle(&ipc, h->px_buf.p));
  static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(out, &ipc, 64);
  return KB_OK;
}

kb_status kb_attach_peers(kb_handle* h, const char* handles /*[n_ranks][64], in rank order*/) {
  if (!h->px_buf.p) return fail(h, KB_ERR_STATE, "no peer-exchange buffer (n_ranks must be 2..8)");
  if (h->px_on) return fail(h, KB_ERR_STATE, "peers are already attached");
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < h->n_ranks; ++r) {
    if (r == h->rank) continue;
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    KB_CUDA(h, cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->px_opened[r] = ptr;
    h->d.px.base[r] = (double*)ptr;
  }
  h->d.px.enabled = 1;
  h->px_on = true;
  if (h->lm_graph) { cudaGraphExecDestroy(h->lm_graph); h->lm_graph = nullptr; }
  return KB_OK;
}

kb_status kb_set_speculative_linearise(kb_handle* h, int32_t on) {
  h->speculative = on != 0;
  return KB_OK;
}

kb_status kb_set_constant_conditioner(kb_handle* h, double lambda) {
  h->lambda = lambda;
  return KB_OK;
}

// Everything after the camera-side solution sits in dxc: back-substitution of the set poses, dx^T (lambda dx + rhs) and max|dx|
// (combined over the ranks), the optional copy of dx to the host.  The pose factors Lv must belong to the same damping.
static kb_status solve_post_sync(kb_handle* h);
static kb_status solve_finish(kb_handle* h, double* dx, int32_t gather_dx) {
  StreamCtx c = ctx(h);
  {
    StageTimer t(h, 5);
    // the back substitution leaves the per-block partials of dx^T (lambda dx + rhs) and max|dx| behind
    // ... and its last block adds them up, so that getLmRho / applyStateUpdate need no further pass over dx
    KB_CUDA(h, launch_backsub(h->d, h->set_col_q.p, h->set_col_t.p, h->cam_cols.p, h->lambda, h->rank == 0 ? 1 : 0, h->scalars.p + 2, h->posdef.p, h->posdef.p, 0, c));
  }
  {
    if (h->px_on) {
      KB_CUDA(h, launch_px_combine_solve(h->d, h->scalars.p + 2, h->posdef.p, 0, c));
    } else if (h->n_ranks > 1) {  // one packed all-reduce instead of three (min / sum / max)
      KB_CUDA(h, launch_pack_rank_scalars(h->rank_slots.p, h->rank, h->n_ranks, h->scalars.p + 2, h->posdef.p, c));
      kb_status st = nccl_allreduce(h, h->rank_slots.p, 4 * (size_t)h->n_ranks, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
      KB_CUDA(h, cudaMemcpyAsync(h->h_scalars + 8, h->rank_slots.p, 4 * sizeof(double) * h->n_ranks, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  KB_CUDA(h, cudaMemcpyAsync(h->h_posdef, h->posdef.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaMemcpyAsync(h->h_scalars + 2, h->scalars.p + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (h->px_on)  // the exchange's time-out flag travels with the scalars
    KB_CUDA(h, cudaMemcpyAsync(h->h_scalars + 7, h->px_buf.p + px_off_flags(h->d.px) + 3 * (size_t)h->n_ranks + 4, 8, cudaMemcpyDeviceToHost, h->stream));
  if (dx) {
    if (gather_dx && h->n_ranks > 1) {
      // poses are disjoint across ranks, the shared block is identical: sum with the shared block kept on rank 0 only
      KB_CUDA(h, h->gather.n == (size_t)h->jcols ? cudaSuccess : h->gather.alloc((size_t)h->jcols));
      KB_CUDA(h, cudaMemcpyAsync(h->gather.p, h->dx.p, sizeof(double) * h->jcols, cudaMemcpyDeviceToDevice, h->stream));
      if (h->rank != 0) {
        std::vector<int> cols = h->h_cam_cols;
        std::sort(cols.begin(), cols.end());
        // camera-side columns form at most a few contiguous runs
        size_t i = 0;
        while (i < cols.size()) {
          size_t j = i;
          while (j + 1 < cols.size() && cols[j + 1] == cols[j] + 1) ++j;
          KB_CUDA(h, cudaMemsetAsync(h->gather.p + cols[i], 0, sizeof(double) * (j - i + 1), h->stream));
          i = j + 1;
        }
      }
      kb_status st = nccl_allreduce(h, h->gather.p, (size_t)h->jcols, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
      KB_CUDA(h, cudaMemcpyAsync(dx, h->gather.p, sizeof(double) * h->jcols, cudaMemcpyDeviceToHost, h->stream));
    } else {
      KB_CUDA(h, cudaMemcpyAsync(dx, h->dx.p, sizeof(double) * h->jcols, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  if (h->defer_sync) return KB_OK;
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return solve_post_sync(h);
}

static kb_status solve_post_sync(kb_handle* h) {
  collect_stages(h);
  if (h->n_ranks > 1 && !h->px_on) {  // combine the ranks' slots in rank order: identical on every rank
    double rho = 0.0, mx = 0.0;
    int pd = 1;
    for (int r = 0; r < h->n_ranks; ++r) {
      rho += h->h_scalars[8 + 4 * r];
      mx = std::max(mx, h->h_scalars[8 + 4 * r + 1]);
      if (h->h_scalars[8 + 4 * r + 2] < 0.5) pd = 0;
    }
    h->h_scalars[2] = rho;
    h->h_scalars[3] = mx;
    h->h_posdef[0] = pd;
  }
  if (h->px_on) {
    unsigned long long err = 0;
    std::memcpy(&err, h->h_scalars + 7, 8);
    if (err) return fail(h, KB_ERR_NCCL, "peer exchange timed out: a rank did not reach the exchange step");
  }
  return KB_OK;
}

kb_status kb_solve_system(kb_handle* h, double* dx, int32_t gather_dx, int32_t* pos_def) {
  if (!h->built) return fail(h, KB_ERR_STATE, "kb_solve_system called before kb_build_system");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  // diag(H) += lambda^2 on top of whatever earlier solves left there (BlockCholeskyLinearSystemSolver.cpp:77-86)
  const double damping = h->diag_residual + h->lambda * h->lambda;
  KB_CUDA(h, cudaMemcpyAsync(h->posdef.p, h->posdef.p + 2, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  {
    StageTimer t(h, 3);
    KB_CUDA(h, launch_schur(h->d, damping, h->partials.p, h->n_partials, h->posdef.p, c));
    KB_CUDA(h, launch_schur_finalize(h->d, damping, h->partials.p, h->n_partials, true, c));
    if (!h->px_on) {  // (peer exchange: the partials have gone straight into every rank's buffer; the solve kernel sums them in rank order)
      kb_status st = nccl_allreduce(h, h->Sred.p, (size_t)h->d.n_aug * h->d.n_aug, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
    }
  }
  {
    StageTimer t(h, 4);
    KB_CUDA(h, launch_reduced_solve(h->d, damping, h->posdef.p, h->px_on, c));
  }
  {
    kb_status st = solve_finish(h, dx, gather_dx);
    if (st != KB_OK) return st;
  }
  // un-augment: BlockCholesky subtracts lambda, not lambda^2 (BlockCholeskyLinearSystemSolver.cpp:91-97)
  if (h->semantic == 0) h->diag_residual += h->lambda * h->lambda - h->lambda;
  h->solved = true;
  h->rho_lambda = h->lambda;
  if (pos_def) *pos_def = h->h_posdef[0];
  return KB_OK;
}

// One iteration's worth of the call-by-call sequence - evaluateError, buildSystem, setConstantConditioner, solveSystem,
// applyStateUpdate [, revertLastStateUpdate] - enqueued back to back with ONE host synchronisation at the end.
kb_status kb_iterate(kb_handle* h, double lambda, int32_t use_m_estimator, int32_t revert, kb_iteration_result* out) {
  h->defer_sync = true;
  kb_status st = kb_evaluate_error(h, use_m_estimator, nullptr);
  if (st == KB_OK) st = kb_build_system(h, use_m_estimator);
  if (st == KB_OK) st = kb_set_constant_conditioner(h, lambda);
  if (st == KB_OK) st = kb_solve_system(h, nullptr, 0, nullptr);
  if (st == KB_OK) st = kb_apply_state_update(h, nullptr);
  if (st == KB_OK && revert) st = kb_revert_last_state_update(h);
  h->defer_sync = false;
  if (st != KB_OK) {
    cudaStreamSynchronize(h->stream);
    return st;
  }
  return out ? kb_wait(h, out) : KB_OK;  // out == NULL: enqueue only
}

kb_status kb_wait(kb_handle* h, kb_iteration_result* out) {
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (kb_status st = solve_post_sync(h); st != KB_OK) return st;
  if (out) {
    out->cost = h->h_scalars[0];
    out->rho_denominator = h->h_scalars[2];
    out->max_abs_dx = h->h_scalars[3];
    out->pos_def = h->h_posdef[0];
  }
  return KB_OK;
}

kb_status kb_analyze_marginal(kb_handle* h, const kb_marginal_options* o, kb_marginal_result* out, double* singular_values, double* V, int32_t* columns) {
  return analyze_marginal_impl(h, o, out, singular_values, V, columns, true);
}
kb_status kb_analyze_marginal_last_build(kb_handle* h, const kb_marginal_options* o, kb_marginal_result* out, double* singular_values, double* V,
                                         int32_t* columns) {
  return analyze_marginal_impl(h, o, out, singular_values, V, columns, false);
}
kb_status kb_get_last_svd_decomposition(kb_handle* h, double* singular_values, double* V, int32_t* columns) {
  if (h->last_svd.rank < 0) return fail(h, KB_ERR_STATE, "kb_get_last_svd_decomposition before any kb_solve_system_svd");
  const int n = h->d.n_c;
  KB_CUDA(h, cudaSetDevice(h->device));
  if (singular_values) KB_CUDA(h, cudaMemcpyAsync(singular_values, h->eig_sv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (V) KB_CUDA(h, cudaMemcpyAsync(V, h->eig_Vout.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (columns)
    for (int i = 0; i < n; ++i) columns[i] = h->h_cam_cols[i];
  return KB_OK;
}
kb_status kb_get_last_svd_solve(const kb_handle* h, kb_svd_solve_result* out) {
  if (!out) return KB_ERR_INVALID_ARGUMENT;
  *out = h->last_svd;
  return KB_OK;
}

// ---- the incremental estimator's linear solver: undamped, calibration block through a truncated SVD ---------------------------
void kb_default_svd_solver_options(kb_svd_solver_options* o) {  // IC/src/core/LinearSolverOptions.cpp:30-38
  o->column_scaling = 0;
  o->eps_norm = 2.220446049250313e-16;
  o->eps_svd = 2.220446049250313e-16;
  o->svd_tol = -1.0;
}

kb_status kb_solve_system_svd(kb_handle* h, const kb_svd_solver_options* o, double* dx, int32_t gather_dx, kb_svd_solve_result* out,
                              double* singular_values) {
  if (!o) return fail(h, KB_ERR_INVALID_ARGUMENT, "null options");
  if (!h->built) return fail(h, KB_ERR_STATE, "kb_solve_system_svd called before kb_build_system");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  const int n = h->d.n_c;
  if (n > 255) return fail(h, KB_ERR_INVALID_ARGUMENT, "calibration block too large");
  if (h->eig_G.n != (size_t)(n + 1) * n) {
    KB_CUDA(h, h->eig_G.alloc((size_t)(n + 1) * n));
    KB_CUDA(h, h->eig_V.alloc((size_t)(n + 1) * n));
    KB_CUDA(h, h->eig_sv.alloc(n));
    KB_CUDA(h, h->eig_Vout.alloc((size_t)n * n));
    KB_CUDA(h, h->eig_Vtmp.alloc((size_t)n * n));
    KB_CUDA(h, h->eig_sweeps.alloc(8));
  }
  if (h->svd_diag.n != (size_t)n) {
    KB_CUDA(h, h->svd_diag.alloc(n));
    KB_CUDA(h, h->svd_g.alloc(n));
    KB_CUDA(h, h->svd_result.alloc(4));
  }
  h->lambda = 0.0;  // GaussNewtonTrustRegionPolicy: no conditioner (requiresAugmentedDiagonal() == false)
  KB_CUDA(h, cudaMemcpyAsync(h->posdef.p, h->posdef.p + 2, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  kb_status st;
  {
    StageTimer t(h, 3);
    KB_CUDA(h, launch_schur(h->d, 0.0, h->partials.p, h->n_partials, h->posdef.p, c));
    KB_CUDA(h, launch_schur_finalize(h->d, 0.0, h->partials.p, h->n_partials, true, c));
    if (h->px_on) {
      KB_CUDA(h, launch_px_reduce_system(h->d, c));
    } else if ((st = nccl_allreduce(h, h->Sred.p, (size_t)h->d.n_aug * h->d.n_aug, kNcclFloat64, kNcclSum)) != KB_OK) {
      return st;
    }
    KB_CUDA(h, launch_camera_diag(h->d, h->svd_diag.p, c));
    if ((st = nccl_allreduce(h, h->svd_diag.p, (size_t)n, kNcclFloat64, kNcclSum)) != KB_OK) return st;
  }
  {
    StageTimer t(h, 4);
    const double norm_tol = std::sqrt((double)kb_jrows(h) * o->eps_norm);  // columnScalingMatrix: sqrt(A->nrow * eps)
    const int kind = o->column_scaling ? 1 : 0;
    const double* warm = eig_warm_start(h, kind);
    KB_CUDA(h, launch_svd_solve(h->d, h->svd_diag.p, norm_tol, o->column_scaling ? 1 : 0, o->eps_svd, o->svd_tol, h->svd_g.p, h->eig_G.p, h->eig_V.p,
                                h->eig_sv.p, h->eig_Vout.p, h->eig_Vtmp.p, h->eig_sweeps.p, h->svd_result.p, warm, h->eig_warm[kind].p, c));
  }
  double res[4] = {0, 0, 0, 0};
  int sweeps2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  KB_CUDA(h, cudaMemcpyAsync(res, h->svd_result.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaMemcpyAsync(sweeps2, h->eig_sweeps.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (singular_values) KB_CUDA(h, cudaMemcpyAsync(singular_values, h->eig_sv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  st = solve_finish(h, dx, gather_dx);  // synchronises
  if (st != KB_OK) return st;
  h->solved = true;
  h->rho_lambda = 0.0;
  if (!h->h_posdef[0]) return fail(h, KB_ERR_NUMERICAL, "a set pose is not constrained by its observations (pose block not positive definite)");
  const int sweeps = sweeps2[1] ? 99 : sweeps2[0];
  if (getenv("KB_SVD_TRACE")) std::fprintf(stderr, "[kb trace] truncated-SVD solve: n = %d, Jacobi polish sweeps = %d, QL failed = %d, rank = %d, tridiag %d kcycles, QL %d kcycles, %d QL steps, %d rotations\n", n, sweeps2[0], sweeps2[1], (int)res[0], sweeps2[2], sweeps2[3], sweeps2[4], sweeps2[5]);
  if (sweeps >= 40) return fail(h, KB_ERR_NUMERICAL, "the eigen-decomposition of the truncated-SVD solve did not converge");
  h->last_svd.n = n;
  h->last_svd.rank = (int32_t)res[0];
  h->last_svd.rank_deficiency = n - h->last_svd.rank;
  h->last_svd.tolerance = res[1];
  h->last_svd.sv_gap = res[2];
  if (out) *out = h->last_svd;
  return KB_OK;
}

kb_status kb_lm_rho_denominator(kb_handle* h, double lambda, double* out) {
  if (!h->solved) return fail(h, KB_ERR_STATE, "kb_lm_rho_denominator called before kb_solve_system");
  if (lambda != h->rho_lambda) {  // the solve already computed it for its own lambda (the LM policy's case)
    KB_CUDA(h, cudaSetDevice(h->device));
    StreamCtx c = ctx(h);
    KB_CUDA(h, launch_rho_denominator(h->d, lambda, h->set_col_q.p, h->set_col_t.p, h->cam_cols.p, h->rank == 0 ? 1 : 0, h->scalars.p + 6, nullptr, c));
    kb_status st = nccl_allreduce(h, h->scalars.p + 6, 1, kNcclFloat64, kNcclSum);
    if (st != KB_OK) return st;
    KB_CUDA(h, cudaMemcpyAsync(h->h_scalars + 6, h->scalars.p + 6, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out) *out = h->h_scalars[6];
    return KB_OK;
  }
  if (out) *out = h->h_scalars[2];
  return KB_OK;
}

kb_status kb_apply_state_update(kb_handle* h, double* out_max_abs_dx) {
  if (!h->solved) return fail(h, KB_ERR_STATE, "kb_apply_state_update called before kb_solve_system");
  KB_CUDA(h, cudaSetDevice(h->device));
  StreamCtx c = ctx(h);
  {
    StageTimer t(h, 6);
    KB_CUDA(h, launch_apply_update(h->d, h->set_col_q.p, h->set_col_t.p, h->cam_cols.p, h->bk_cam.p, h->bk_base.p, h->bk_sets.p, c));
  }
  // no synchronisation needed: max|dx| came back with the solve; the next call on this stream sees the updated state
  h->has_backup = true;
  ++h->state_version;
  if (out_max_abs_dx) *out_max_abs_dx = h->h_scalars[3];
  return KB_OK;
}

kb_status kb_revert_last_state_update(kb_handle* h) {
  if (!h->has_backup) return fail(h, KB_ERR_STATE, "kb_revert_last_state_update without a previous update");
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaMemcpyAsync(h->cam_params.p, h->bk_cam.p, sizeof(double) * h->cam_params.n, cudaMemcpyDeviceToDevice, h->stream));
  if (h->baselines.n) KB_CUDA(h, cudaMemcpyAsync(h->baselines.p, h->bk_base.p, sizeof(double) * h->baselines.n, cudaMemcpyDeviceToDevice, h->stream));
  if (h->set_poses.n) KB_CUDA(h, cudaMemcpyAsync(h->set_poses.p, h->bk_sets.p, sizeof(double) * h->set_poses.n, cudaMemcpyDeviceToDevice, h->stream));
  ++h->state_version;
  return KB_OK;
}

kb_status kb_reset_state(kb_handle* h) {
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaMemcpyAsync(h->cam_params.p, h->init_cam.p, sizeof(double) * h->cam_params.n, cudaMemcpyDeviceToDevice, h->stream));
  if (h->baselines.n) KB_CUDA(h, cudaMemcpyAsync(h->baselines.p, h->init_base.p, sizeof(double) * h->baselines.n, cudaMemcpyDeviceToDevice, h->stream));
  if (h->set_poses.n) KB_CUDA(h, cudaMemcpyAsync(h->set_poses.p, h->init_sets.p, sizeof(double) * h->set_poses.n, cudaMemcpyDeviceToDevice, h->stream));
  h->built = h->solved = h->has_backup = false;
  ++h->state_version;
  h->diag_residual = 0.0;
  h->lambda = 0.0;
  return KB_OK;
}

kb_status kb_set_observations(kb_handle* h, const double* y_u, const double* y_v) { return set_observations_impl<double>(h, y_u, y_v); }
kb_status kb_set_observations_f32(kb_handle* h, const float* y_u, const float* y_v) { return set_observations_impl<float>(h, y_u, y_v); }

}  // extern "C"
// Double buffering of the measurements for callers that feed a new batch per step: the upload of the NEXT batch runs on the
// copy stream into the back buffers while the current step computes on the front buffers.
template <typename T>
static kb_status prefetch_observations_impl(kb_handle* h, const T* y_u, const T* y_v) {
  if (h->n_ranks != 1 && !h->presharded)
    return fail(h, KB_ERR_STATE, "kb_prefetch_observations needs a single rank or a pre-sharded problem (terms are re-packed per rank otherwise)");
  if (!y_u || !y_v) return fail(h, KB_ERR_INVALID_ARGUMENT, "null observation array");
  KB_CUDA(h, cudaSetDevice(h->device));
  kb_status st = ensure_stage<T>(h);
  if (st != KB_OK) return st;
  if (!h->back_u) {
    KB_CUDA(h, h->y_u_back.alloc((size_t)h->n_terms_local));
    KB_CUDA(h, h->y_v_back.alloc((size_t)h->n_terms_local));
    h->back_u = h->y_u_back.p;
    h->back_v = h->y_v_back.p;
  }
  // the back buffers were the front ones until the last commit: wait for everything that was enqueued on them
  if (h->front_free_recorded) KB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_front_free, 0));
  if ((st = upload_chunk<T>(h, y_u, y_v, 0, h->n_terms_local, h->back_u, h->back_v, h->copy_stream)) != KB_OK) return st;
  KB_CUDA(h, cudaEventRecord(h->ev_prefetch, h->copy_stream));
  h->prefetch_pending = true;
  return KB_OK;
}
extern "C" {
kb_status kb_prefetch_observations(kb_handle* h, const double* y_u, const double* y_v) { return prefetch_observations_impl<double>(h, y_u, y_v); }
kb_status kb_prefetch_observations_f32(kb_handle* h, const float* y_u, const float* y_v) { return prefetch_observations_impl<float>(h, y_u, y_v); }

kb_status kb_commit_observations(kb_handle* h) {
  if (!h->prefetch_pending) return fail(h, KB_ERR_STATE, "kb_commit_observations without a pending kb_prefetch_observations");
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_prefetch, 0));
  std::swap(h->front_u, h->back_u);
  std::swap(h->front_v, h->back_v);
  h->d.y_u = h->front_u;
  h->d.y_v = h->front_v;
  KB_CUDA(h, cudaEventRecord(h->ev_front_free, h->stream));  // everything enqueued so far read the old front buffers
  h->front_free_recorded = true;
  h->prefetch_pending = false;
  ++h->state_version;
  if (h->lm_graph) { cudaGraphExecDestroy(h->lm_graph); h->lm_graph = nullptr; }  // captured with the old buffer addresses
  return KB_OK;
}

int64_t kb_num_invalid_terms(kb_handle* h) {  // -1 on a CUDA error (message in kb_last_error)
  unsigned int v = 0;
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&v, h->n_invalid.p, sizeof(v), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    h->error = std::string("kb_num_invalid_terms: ") + cudaGetErrorString(e);
    return -1;
  }
  return (int64_t)v;
}

void kb_default_optimizer_options(kb_optimizer_options* o) {
  o->convergence_delta_x = 1e-3;
  o->convergence_delta_j = 1.0;
  o->max_iterations = 200;
  o->lm_lambda_init = 10.0;
  o->verbose = 0;
  o->device_loop = 1;
}

// One LM iteration, enqueued without any host decision: every kernel looks at the control block and is a no-op when the
// iteration does not need it (no rebuild after a rejected step, no update after a failed solve, nothing once the loop is done).
static kb_status enqueue_lm_iteration(kb_handle* h) {
  StreamCtx c = ctx(h);
  const DevProblem& D = h->d;
  // where the state-machine transitions run: inside the last kernel of the solve / of the evaluation (single rank: 10 launches per
  // iteration; peer exchange: inside the exchange consumers), or in their own one-thread kernels behind an NCCL all-reduce
  const int lm_local = h->n_ranks == 1 ? 1 : 0, lm_px = h->px_on ? 1 : 0;
  {
    StageTimer t(h, 2);
    KB_CUDA(h, launch_set_reduce(D, c));
  }
  {
    StageTimer t(h, 3);
    KB_CUDA(h, launch_schur(D, KB_DAMPING_FROM_CTRL, h->partials.p, h->n_partials, h->posdef.p, c));
    KB_CUDA(h, launch_schur_finalize(D, KB_DAMPING_FROM_CTRL, h->partials.p, h->n_partials, true, c));
    if (!h->px_on) {
      kb_status st = nccl_allreduce(h, D.Sred, (size_t)D.n_aug * D.n_aug, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
    }
  }
  {
    StageTimer t(h, 4);
    KB_CUDA(h, launch_reduced_solve(D, KB_DAMPING_FROM_CTRL, h->posdef.p, h->px_on, c));
  }
  {
    StageTimer t(h, 5);
    KB_CUDA(h, launch_backsub(D, h->set_col_q.p, h->set_col_t.p, h->cam_cols.p, -1.0, h->rank == 0 ? 1 : 0, h->scalars.p + 2, h->posdef.p, h->posdef.p, lm_local, c));
  }
  if (h->px_on) {
    KB_CUDA(h, launch_px_combine_solve(D, h->scalars.p + 2, h->posdef.p, lm_px, c));
  } else if (h->n_ranks > 1) {
    KB_CUDA(h, launch_pack_rank_scalars(h->rank_slots.p, h->rank, h->n_ranks, h->scalars.p + 2, h->posdef.p, c));
    kb_status st = nccl_allreduce(h, h->rank_slots.p, 4 * (size_t)h->n_ranks, kNcclFloat64, kNcclSum);
    if (st != KB_OK) return st;
    KB_CUDA(h, launch_lm_post_solve(D, h->posdef.p, h->scalars.p + 2, h->rank_slots.p, h->n_ranks, c));
  }
  {
    StageTimer t(h, 6);
    KB_CUDA(h, launch_apply_update(D, h->set_col_q.p, h->set_col_t.p, h->cam_cols.p, h->bk_cam.p, h->bk_base.p, h->bk_sets.p, c));
  }
  {
    StageTimer t(h, 0);
    {
      StageTimer t1(h, 1);
      KB_CUDA(h, launch_linearise_assemble(D, h->vmeta.p, h->slices.p, h->slice_model_begin, true, true, true, c));
    }
    KB_CUDA(h, launch_finalize_gram(D, h->cam_slice_range.p, 1, &h->ctrl.p->cost_new, true, lm_local, h->trace_dev.p, h->posdef.p, c));
    if (h->px_on) {
      KB_CUDA(h, launch_px_combine_cost(D, &h->ctrl.p->cost_new, lm_px, h->trace_dev.p, h->posdef.p, c));
    } else if (h->n_ranks > 1) {
      kb_status st = nccl_allreduce(h, &h->ctrl.p->cost_new, 1, kNcclFloat64, kNcclSum);
      if (st != KB_OK) return st;
      KB_CUDA(h, launch_lm_boundary(D, h->trace_dev.p, h->posdef.p, c));
    }
  }
  return KB_OK;
}

static kb_status optimize_on_device(kb_handle* h, const kb_optimizer_options* o, kb_solution* out) {
  KB_CUDA(h, cudaSetDevice(h->device));
  double J0 = 0.0;
  kb_status st = kb_evaluate_error(h, 1, &J0);  // Optimizer2.cpp:198
  if (st != KB_OK) return st;
  const int max_it = std::max(o->max_iterations, 0);
  if (h->trace_dev.n < (size_t)3 * (max_it + 1)) KB_CUDA(h, h->trace_dev.alloc((size_t)3 * (max_it + 1)));
  LmCtrl c0;
  std::memset(&c0, 0, sizeof(c0));
  kalibr_b200::lm_start(&c0, kalibr_b200::KB_POLICY_LEVENBERG_MARQUARDT, J0, o->lm_lambda_init, o->convergence_delta_x, o->convergence_delta_j, max_it,
                        h->semantic);
  // the decisions of the FIRST iteration are taken here; those of every later one by the device at the end of the iteration before
  if (!c0.done) kalibr_b200::lm_before_solve(&c0);
  KB_CUDA(h, cudaMemcpyAsync(h->posdef.p, h->posdef.p + 2, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  KB_CUDA(h, cudaMemcpyAsync(h->ctrl.p, &c0, sizeof(c0), cudaMemcpyHostToDevice, h->stream));  // pageable source: staged at the call
  *h->h_ctrl = c0;
  // iterations are enqueued two at a time; the host only reads the control block back to see whether the loop has ended.
  // The kernel arguments of an iteration never change (everything data-dependent lives in the control block), so after
  // one plain-launched batch the iteration is captured into a CUDA graph and replayed: one launch per iteration instead of ~25.
  const bool may_graph = (h->n_ranks == 1 || h->px_on) && !h->timing && !getenv("KB_NO_GRAPH");  // no NCCL call inside an iteration
  while (!h->h_ctrl->done) {
    if (may_graph && h->lm_warm && (!h->lm_graph || h->lm_graph_trace != h->trace_dev.p)) {
      if (h->lm_graph) { cudaGraphExecDestroy(h->lm_graph); h->lm_graph = nullptr; }
      cudaGraph_t graph = nullptr;
      const long long l0 = h->launches;
      KB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      st = enqueue_lm_iteration(h);
      cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      h->lm_graph_kernels = h->launches - l0;  // captured, not launched
      h->launches = l0;
      if (st != KB_OK || ce != cudaSuccess) {  // never leak the captured graph
        if (graph) cudaGraphDestroy(graph);
        if (st != KB_OK) return st;
        KB_CUDA(h, ce);
      }
      ce = cudaGraphInstantiate(&h->lm_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { h->lm_graph = nullptr; KB_CUDA(h, ce); }
      h->lm_graph_trace = h->trace_dev.p;
    }
    for (int k = 0; k < 2; ++k) {
      if (may_graph && h->lm_graph) {
        KB_CUDA(h, cudaGraphLaunch(h->lm_graph, h->stream));
        h->launches += h->lm_graph_kernels;
      } else if ((st = enqueue_lm_iteration(h)) != KB_OK) {
        return st;
      }
    }
    h->lm_warm = true;
    KB_CUDA(h, cudaMemcpyAsync(h->h_ctrl, h->ctrl.p, sizeof(LmCtrl), cudaMemcpyDeviceToHost, h->stream));
    if (h->px_on)  // the exchange's time-out flag travels with the control block (a timed-out wait also sets ctrl->done)
      KB_CUDA(h, cudaMemcpyAsync(h->h_scalars + 7, h->px_buf.p + px_off_flags(h->d.px) + 3 * (size_t)h->n_ranks + 4, 8, cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    collect_stages(h);
    if (h->px_on) {
      unsigned long long err = 0;
      std::memcpy(&err, h->h_scalars + 7, 8);
      if (err) {
        StreamCtx sc = ctx(h);
        launch_lm_finish(h->d, sc);  // back to the neutral flags: the handle reports the error instead of silently idling
        cudaStreamSynchronize(h->stream);
        return fail(h, KB_ERR_NCCL, "peer exchange timed out inside the device-resident loop: a rank did not reach the exchange step");
      }
    }
  }
  const LmCtrl& c = *h->h_ctrl;
  h->trace.assign((size_t)3 * c.iterations, 0.0);
  if (c.iterations > 0) KB_CUDA(h, cudaMemcpyAsync(h->trace.data(), h->trace_dev.p, sizeof(double) * 3 * c.iterations, cudaMemcpyDeviceToHost, h->stream));
  StreamCtx sc = ctx(h);
  KB_CUDA(h, launch_lm_revert(h->d, h->bk_cam.p, h->bk_base.p, h->bk_sets.p, sc));  // a rejected last step is undone now (restores are lazy)
  KB_CUDA(h, launch_lm_finish(h->d, sc));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  // host-side solver state after the loop
  h->lambda = c.lambda;
  h->rho_lambda = c.lambda;
  h->diag_residual = c.diag_residual;
  h->h_scalars[2] = c.rho_den;
  h->h_scalars[3] = c.max_dx;
  h->built = h->solved = h->has_backup = c.iterations + c.failed > 0;
  ++h->state_version;
  h->la_version = -1;  // the view blocks belong to the last TRIAL state: the next build linearises again
  if (out) {
    out->j_start = c.JStart;
    out->j_final = c.pJ;
    out->dx_final = c.deltaX;
    out->dj_final = c.deltaJ;
    out->iterations = c.iterations;
    out->failed_iterations = c.failed;
    out->linear_solver_failure = c.solver_failure;
  }
  return KB_OK;
}

kb_status kb_optimize(kb_handle* h, const kb_optimizer_options* o, kb_solution* out) {
  using namespace kalibr_b200::backend;
  if (o->device_loop && !o->verbose && h->speculative) return optimize_on_device(h, o, out);
  try {
    Optimizer2Options opt;
    opt.convergenceDeltaX = o->convergence_delta_x;
    opt.convergenceDeltaJ = o->convergence_delta_j;
    opt.maxIterations = o->max_iterations;
    opt.verbose = o->verbose != 0;
    opt.linearSystemSolver = std::make_shared<B200SchurLinearSystemSolver>(h, true);
    opt.trustRegionPolicy = std::make_shared<LevenbergMarquardtTrustRegionPolicy>(o->lm_lambda_init);
    Optimizer2 optimizer(opt);
    SolutionReturnValue srv = optimizer.optimize();
    h->trace = optimizer.trace();
    if (out) {
      out->j_start = srv.JStart;
      out->j_final = srv.JFinal;
      out->dx_final = srv.dXFinal;
      out->dj_final = srv.dJFinal;
      out->iterations = srv.iterations;
      out->failed_iterations = srv.failedIterations;
      out->linear_solver_failure = srv.linearSolverFailure ? 1 : 0;
    }
    return KB_OK;
  } catch (const KbError& e) {  // the status of the entry point that failed, not a blanket KB_ERR_CUDA
    if (h->error.empty()) h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return KB_ERR_INVALID_ARGUMENT;
  }
}

// Optimizer2::optimize with the Gauss-Newton policy and the estimator's linear solver: what IncrementalEstimator runs per batch
// (IC/src/core/IncrementalEstimator.cpp:66-71, 377).  Host-driven (two synchronisations per iteration).
kb_status kb_optimize_gauss_newton(kb_handle* h, const kb_optimizer_options* o, const kb_svd_solver_options* so, kb_solution* out) {
  using namespace kalibr_b200::backend;
  if (!o || !so) return fail(h, KB_ERR_INVALID_ARGUMENT, "null options");
  try {
    Optimizer2Options opt;
    opt.convergenceDeltaX = o->convergence_delta_x;
    opt.convergenceDeltaJ = o->convergence_delta_j;
    opt.maxIterations = o->max_iterations;
    opt.verbose = o->verbose != 0;
    opt.linearSystemSolver = std::make_shared<B200SvdLinearSystemSolver>(h, *so);
    opt.trustRegionPolicy = std::make_shared<GaussNewtonTrustRegionPolicy>();
    Optimizer2 optimizer(opt);
    SolutionReturnValue srv = optimizer.optimize();
    h->trace = optimizer.trace();
    if (out) {
      out->j_start = srv.JStart;
      out->j_final = srv.JFinal;
      out->dx_final = srv.dXFinal;
      out->dj_final = srv.dJFinal;
      out->iterations = srv.iterations;
      out->failed_iterations = srv.failedIterations;
      out->linear_solver_failure = srv.linearSolverFailure ? 1 : 0;
    }
    return KB_OK;
  } catch (const KbError& e) {
    if (h->error.empty()) h->error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->error = e.what();
    return KB_ERR_INVALID_ARGUMENT;
  }
}

int32_t kb_get_trace(const kb_handle* h, double* out, int32_t max_triples) {
  const int32_t n = (int32_t)(h->trace.size() / 3);
  if (out)
    for (int32_t i = 0; i < std::min(n, max_triples) * 3; ++i) out[i] = h->trace[i];
  return n;
}

// ---- read-back --------------------------------------------------------------------------------------------
kb_status kb_get_error_vector(kb_handle* h, double* e) {
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaMemcpyAsync(e, h->e.p, sizeof(double) * 2 * h->n_terms_local, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

static kb_status download(kb_handle* h, const DevBuf<double>& b, std::vector<double>& out) {
  out.resize(b.n);
  if (b.n) KB_CUDA(h, cudaMemcpyAsync(out.data(), b.p, sizeof(double) * b.n, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

kb_status kb_get_rhs(kb_handle* h, double* rhs) {
  if (!h->built) return fail(h, KB_ERR_STATE, "kb_get_rhs called before kb_build_system");
  KB_CUDA(h, cudaSetDevice(h->device));
  std::vector<double> bv, U;
  kb_status st;
  if ((st = download(h, h->bv, bv)) != KB_OK) return st;
  if (h->n_ranks > 1) {  // b_c is a per-rank partial: reduce a copy
    KB_CUDA(h, h->gather.alloc((size_t)h->d.n_aug * h->d.n_aug));
    KB_CUDA(h, cudaMemcpyAsync(h->gather.p, h->U.p, sizeof(double) * h->U.n, cudaMemcpyDeviceToDevice, h->stream));
    if ((st = nccl_allreduce(h, h->gather.p, h->U.n, kNcclFloat64, kNcclSum)) != KB_OK) return st;
    if ((st = download(h, h->gather, U)) != KB_OK) return st;
  } else if ((st = download(h, h->U, U)) != KB_OK) return st;
  std::fill(rhs, rhs + h->jcols, 0.0);
  const int n = h->d.n_aug;
  if (h->rank == 0 || h->n_ranks == 1)
    for (int i = 0; i < h->d.n_c; ++i) rhs[h->h_cam_cols[i]] = U[(size_t)i * n + h->d.n_c];
  for (int lv = 0; lv < h->d.n_sets; ++lv) {
    const int cq = h->dv_col[h->dv_first_set + 2 * (h->set_lo + lv)], ct = h->dv_col[h->dv_first_set + 2 * (h->set_lo + lv) + 1];
    for (int c = 0; c < 3; ++c) { rhs[cq + c] = bv[(size_t)lv * 6 + c]; rhs[ct + c] = bv[(size_t)lv * 6 + 3 + c]; }
  }
  if (h->n_ranks > 1) {  // pose segments are disjoint across ranks, the camera segment was kept on rank 0 only: sum = full vector
    KB_CUDA(h, h->gather.alloc((size_t)h->jcols));
    KB_CUDA(h, cudaMemcpyAsync(h->gather.p, rhs, sizeof(double) * h->jcols, cudaMemcpyHostToDevice, h->stream));
    if ((st = nccl_allreduce(h, h->gather.p, (size_t)h->jcols, kNcclFloat64, kNcclSum)) != KB_OK) return st;
    KB_CUDA(h, cudaMemcpyAsync(rhs, h->gather.p, sizeof(double) * h->jcols, cudaMemcpyDeviceToHost, h->stream));
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return KB_OK;
}

kb_status kb_linearise(kb_handle* h) {
  KB_CUDA(h, cudaSetDevice(h->device));
  if (h->jt.n != (size_t)h->jac_nnz) KB_CUDA(h, h->jt.alloc((size_t)h->jac_nnz));
  StreamCtx c = ctx(h);
  {
    StageTimer t(h, 7);
    KB_CUDA(h, launch_linearise_materialise(h->d, h->vmeta.p, h->slices.p, h->slice_model_begin, h->lm_bfrag_pairs, h->lm_counters.p, h->jt.p, c));
  }
  if (h->timing) {
    KB_CUDA(h, cudaStreamSynchronize(h->stream));
    collect_stages(h);
  }
  return KB_OK;
}
int64_t kb_jacobian_nnz(const kb_handle* h) { return h->jac_nnz; }

kb_status kb_get_jacobian_ccs(kb_handle* h, int64_t* col_ptr, int32_t* row_idx, double* values) {
  if (!col_ptr || !row_idx || !values) return fail(h, KB_ERR_INVALID_ARGUMENT, "null output array (use kb_jacobian_nnz for the size)");
  if (h->jt.n != (size_t)h->jac_nnz) return fail(h, KB_ERR_STATE, "kb_get_jacobian_ccs called before kb_linearise");
  KB_CUDA(h, cudaSetDevice(h->device));
  if (h->jac_nnz) KB_CUDA(h, cudaMemcpyAsync(values, h->jt.p, sizeof(double) * h->jac_nnz, cudaMemcpyDeviceToHost, h->stream));
  // structure: per term, design variables sorted by block index (CompressedColumnMatrix.hpp:236-304)
  int64_t o = 0;
  size_t col = 0;
  col_ptr[0] = 0;
  for (size_t w = 0; w < h->h_view_set.size(); ++w) {
    const int k = h->h_view_cam[w], gv = h->set_lo + h->h_view_set[w];
    std::vector<int> blocks = camside_dvs(h, k);
    blocks.push_back(h->dv_first_set + 2 * gv);
    blocks.push_back(h->dv_first_set + 2 * gv + 1);
    std::sort(blocks.begin(), blocks.end());
    std::vector<int32_t> rows;
    for (int b : blocks)
      for (int c = 0; c < h->dv_dim[b]; ++c) rows.push_back(h->dv_col[b] + c);
    for (int i = h->h_view_begin[w]; i < h->h_view_begin[w + 1]; ++i)
      for (int r = 0; r < 2; ++r) {
        std::memcpy(row_idx + o, rows.data(), sizeof(int32_t) * rows.size());
        o += (int64_t)rows.size();
        col_ptr[++col] = o;
      }
  }
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

kb_status kb_get_hessian_blocks(kb_handle* h, int64_t* n_blocks, int64_t* n_values, int64_t* col_ptr, int32_t* block_row, int64_t* value_ptr,
                                double* values) {
  if (h->n_ranks != 1) return fail(h, KB_ERR_STATE, "kb_get_hessian_blocks is a single-rank parity export");
  if (!h->built) return fail(h, KB_ERR_STATE, "kb_get_hessian_blocks called before kb_build_system");
  const int n_dv = (int)h->dv_dim.size();
  // ---- pattern: pairs of design variables sharing a term (JacobianContainer.cpp:112-126), plus every diagonal block once
  //      solveSystem has run (BlockCholeskyLinearSystemSolver.cpp:80, block(i,i,true)) ----
  std::vector<std::set<int>> cols(n_dv);
  std::vector<char> cam_used(h->n_cams, 0);
  for (size_t w = 0; w < h->h_view_set.size(); ++w) {
    if (h->h_view_begin[w + 1] == h->h_view_begin[w]) continue;
    const int k = h->h_view_cam[w], v = h->set_lo + h->h_view_set[w];
    const int bq = h->dv_first_set + 2 * v, bt = bq + 1;
    cols[bq].insert(bq); cols[bt].insert(bq); cols[bt].insert(bt);
    for (int c : camside_dvs(h, k)) {
      for (int pb : {bq, bt}) cols[std::max(c, pb)].insert(std::min(c, pb));
    }
    cam_used[k] = 1;
  }
  for (int k = 0; k < h->n_cams; ++k) {
    if (!cam_used[k]) continue;
    std::vector<int> cs = camside_dvs(h, k);
    for (int a : cs)
      for (int b : cs) cols[std::max(a, b)].insert(std::min(a, b));
  }
  if (h->solved)
    for (int i = 0; i < n_dv; ++i) cols[i].insert(i);
  int64_t nb = 0, nv = 0;
  for (int c = 0; c < n_dv; ++c)
    for (int r : cols[c]) { ++nb; nv += (int64_t)h->dv_dim[r] * h->dv_dim[c]; }
  if (n_blocks) *n_blocks = nb;
  if (n_values) *n_values = nv;
  if (!col_ptr || !block_row || !value_ptr || !values) return KB_OK;
  // ---- values from the device blocks ----
  KB_CUDA(h, cudaSetDevice(h->device));
  std::vector<double> V, W, U;
  kb_status st;
  if ((st = download(h, h->V, V)) != KB_OK) return st;
  if ((st = download(h, h->W, W)) != KB_OK) return st;
  if ((st = download(h, h->U, U)) != KB_OK) return st;
  const int nc = h->d.n_c, na = h->d.n_aug;
  // reduced-system offset of every camera-side block
  std::vector<int> red_off(n_dv, -1);
  for (int k = 0; k < h->n_cams; ++k) { red_off[h->dv_proj[k]] = h->d.intr_off[k]; red_off[h->dv_dist[k]] = h->d.intr_off[k] + h->d.cam_P[k]; }
  for (int j = 0; j + 1 < h->n_cams; ++j) { red_off[h->dv_base_q[j]] = h->d.base_off[j]; red_off[h->dv_base_t[j]] = h->d.base_off[j] + 3; }
  auto is_pose = [&](int b) { return b >= h->dv_first_set && b < h->dv_first_set + 2 * h->n_sets_global; };
  // H(a, b) entry accessor for design-variable blocks a, b (any order), local indices i, j
  auto entry = [&](int a, int i, int b, int j) -> double {
    const bool pa = is_pose(a), pb = is_pose(b);
    if (pa && pb) {
      const int va = (a - h->dv_first_set) / 2, vb2 = (b - h->dv_first_set) / 2;
      if (va != vb2) return 0.0;
      const int ia = ((a - h->dv_first_set) & 1) * 3 + i, ib = ((b - h->dv_first_set) & 1) * 3 + j;
      double v = V[(size_t)(va - h->set_lo) * 36 + ia * 6 + ib];
      if (a == b && i == j) v += h->diag_residual;
      return v;
    }
    if (!pa && !pb) {
      double v = U[(size_t)(red_off[a] + i) * na + red_off[b] + j];
      if (a == b && i == j) v += h->diag_residual;
      return v;
    }
    const int cam_b = pa ? b : a, cam_i = pa ? j : i, pose_b = pa ? a : b, pose_i = pa ? i : j;
    const int v = (pose_b - h->dv_first_set) / 2;
    return W[((size_t)(v - h->set_lo) * nc + red_off[cam_b] + cam_i) * 6 + ((pose_b - h->dv_first_set) & 1) * 3 + pose_i];
  };
  int64_t bi = 0, vi = 0;
  for (int c = 0; c < n_dv; ++c) {
    col_ptr[c] = bi;
    for (int r : cols[c]) {
      block_row[bi] = r;
      value_ptr[bi] = vi;
      for (int j = 0; j < h->dv_dim[c]; ++j)       // column-major like Eigen
        for (int i = 0; i < h->dv_dim[r]; ++i) values[vi++] = entry(r, i, c, j);
      ++bi;
    }
  }
  col_ptr[n_dv] = bi;
  return KB_OK;
}

// ---- a live handle grows and shrinks by synced sets (the incremental estimator's batches) -------------------------------------
namespace {
void structure_changed(kb_handle* h) {
  h->built = h->solved = h->has_backup = false;
  ++h->state_version;
  h->la_version = -1;
  h->diag_residual = 0.0;
  if (h->lm_graph) { cudaGraphExecDestroy(h->lm_graph); h->lm_graph = nullptr; }  // captured with the old sizes / addresses
  h->lm_warm = false;
}
}  // namespace

kb_status kb_append_set(kb_handle* h, int32_t n_views, const int32_t* view_cam, const int64_t* view_begin, const double* y_u, const double* y_v,
                        const int32_t* corner_id, const double* set_pose) {
  if (h->n_ranks != 1) return fail(h, KB_ERR_STATE, "kb_append_set needs a single-rank handle");
  if (h->prefetch_pending || (h->front_u && h->front_u != h->y_u.p))
    return fail(h, KB_ERR_STATE, "kb_append_set while the observation double buffer is in use (commit and stop prefetching first)");
  if (n_views < 0 || !view_begin || !set_pose || (n_views > 0 && (!view_cam || !y_u || !y_v || !corner_id)))
    return fail(h, KB_ERR_INVALID_ARGUMENT, "null argument");
  if (h->driver_order == KB_ORDER_SINGLE && n_views > 1) return fail(h, KB_ERR_INVALID_ARGUMENT, "single-camera problem: one view per set");
  const int64_t n_new = view_begin[n_views] - view_begin[0];
  std::vector<char> seen((size_t)h->n_cams, 0);
  for (int w = 0; w < n_views; ++w) {
    if (view_cam[w] < 0 || view_cam[w] >= h->n_cams || seen[(size_t)view_cam[w]]) return fail(h, KB_ERR_INVALID_ARGUMENT, "bad or repeated camera index in the new set");
    seen[(size_t)view_cam[w]] = 1;
    if (view_begin[w + 1] < view_begin[w]) return fail(h, KB_ERR_INVALID_ARGUMENT, "view_begin is not monotone");
  }
  for (int64_t i = 0; i < n_new; ++i)
    if (corner_id[i] < 0 || corner_id[i] >= h->d.n_target) return fail(h, KB_ERR_INVALID_ARGUMENT, "corner_id out of range");
  if (h->n_terms_local + n_new > (int64_t)0x7fffffff) return fail(h, KB_ERR_INVALID_ARGUMENT, "more than 2^31 terms on one rank");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  KB_CUDA(h, cudaStreamSynchronize(s));
  // observations: appended behind the existing terms (only the new ones travel)
  const size_t n_old = (size_t)h->n_terms_local;
  KB_CUDA(h, h->y_u.grow(n_old + (size_t)n_new, s));
  KB_CUDA(h, h->y_v.grow(n_old + (size_t)n_new, s));
  KB_CUDA(h, h->corner.grow(n_old + (size_t)n_new, s));
  if (n_new > 0) {
    std::vector<uint16_t> c16((size_t)n_new);
    const int64_t b0 = view_begin[0];
    for (int64_t i = 0; i < n_new; ++i) c16[(size_t)i] = (uint16_t)corner_id[b0 + i];
    KB_CUDA(h, cudaMemcpyAsync(h->y_u.p + n_old, y_u + b0, sizeof(double) * (size_t)n_new, cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaMemcpyAsync(h->y_v.p + n_old, y_v + b0, sizeof(double) * (size_t)n_new, cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaMemcpyAsync(h->corner.p + n_old, c16.data(), sizeof(uint16_t) * (size_t)n_new, cudaMemcpyHostToDevice, s));
    KB_CUDA(h, cudaStreamSynchronize(s));  // c16 is a local
  }
  // state: the new pose joins the current state, the reset point and the backup
  const size_t S_old = (size_t)(h->set_hi - h->set_lo);
  for (DevBuf<double>* b : {&h->set_poses, &h->init_sets, &h->bk_sets}) {
    KB_CUDA(h, b->grow((S_old + 1) * KB_POSE_STRIDE, s));
    KB_CUDA(h, cudaMemcpyAsync(b->p + S_old * KB_POSE_STRIDE, set_pose, sizeof(double) * KB_POSE_STRIDE, cudaMemcpyHostToDevice, s));
  }
  KB_CUDA(h, cudaStreamSynchronize(s));
  // structure
  for (int w = 0; w < n_views; ++w) {
    h->h_view_set.push_back((int)S_old);
    h->h_view_cam.push_back(view_cam[w]);
    h->h_view_begin.push_back((int)(n_old + (size_t)(view_begin[w + 1] - view_begin[0])));
  }
  h->n_terms_local += n_new;
  h->n_terms_global = h->n_terms_local;
  h->set_hi += 1;
  h->n_sets_global += 1;
  structure_changed(h);
  kb_status st = build_tables(h);
  if (st != KB_OK) return st;
  KB_CUDA(h, cudaStreamSynchronize(s));
  return KB_OK;
}

kb_status kb_remove_last_set(kb_handle* h) {
  if (h->n_ranks != 1) return fail(h, KB_ERR_STATE, "kb_remove_last_set needs a single-rank handle");
  if (h->prefetch_pending || (h->front_u && h->front_u != h->y_u.p))
    return fail(h, KB_ERR_STATE, "kb_remove_last_set while the observation double buffer is in use");
  const int S = h->set_hi - h->set_lo;
  if (S <= 0) return fail(h, KB_ERR_STATE, "no synced set to remove");
  // the views of the last set must be the trailing views (true for every set that was appended, and for the per-set term order of
  // the rig / batch drivers)
  size_t nv = h->h_view_set.size();
  size_t first = nv;
  while (first > 0 && h->h_view_set[first - 1] == S - 1) --first;
  for (size_t w = 0; w < first; ++w)
    if (h->h_view_set[w] == S - 1) return fail(h, KB_ERR_STATE, "the views of the last synced set are not the trailing views of the problem");
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  const int n_terms = h->h_view_begin[first];
  h->h_view_set.resize(first);
  h->h_view_cam.resize(first);
  h->h_view_begin.resize(first + 1);
  h->n_terms_local = n_terms;
  h->n_terms_global = n_terms;
  h->y_u.n = h->y_v.n = h->corner.n = (size_t)n_terms;  // capacity stays
  for (DevBuf<double>* b : {&h->set_poses, &h->init_sets, &h->bk_sets}) b->n = (size_t)(S - 1) * KB_POSE_STRIDE;
  h->set_hi -= 1;
  h->n_sets_global -= 1;
  structure_changed(h);
  kb_status st = build_tables(h);
  if (st != KB_OK) return st;
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

// ≙ OptimizationProblem::saveDesignVariables / restoreDesignVariables (IC/src/core/OptimizationProblem.cpp:260-272): what
// IncrementalEstimator::addBatch brackets a batch with, so that a rejected batch leaves every design variable as it was
kb_status kb_save_design_variables(kb_handle* h) {
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  KB_CUDA(h, h->sv_cam.alloc(h->cam_params.n));
  KB_CUDA(h, h->sv_base.alloc(h->baselines.n));
  KB_CUDA(h, h->sv_sets.alloc(h->set_poses.n));
  KB_CUDA(h, cudaMemcpyAsync(h->sv_cam.p, h->cam_params.p, sizeof(double) * h->cam_params.n, cudaMemcpyDeviceToDevice, s));
  if (h->baselines.n) KB_CUDA(h, cudaMemcpyAsync(h->sv_base.p, h->baselines.p, sizeof(double) * h->baselines.n, cudaMemcpyDeviceToDevice, s));
  if (h->set_poses.n) KB_CUDA(h, cudaMemcpyAsync(h->sv_sets.p, h->set_poses.p, sizeof(double) * h->set_poses.n, cudaMemcpyDeviceToDevice, s));
  h->saved_sets = h->set_poses.n;
  h->has_saved = true;
  return KB_OK;
}
kb_status kb_restore_design_variables(kb_handle* h) {
  if (!h->has_saved) return fail(h, KB_ERR_STATE, "kb_restore_design_variables without kb_save_design_variables");
  KB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  KB_CUDA(h, cudaMemcpyAsync(h->cam_params.p, h->sv_cam.p, sizeof(double) * h->cam_params.n, cudaMemcpyDeviceToDevice, s));
  if (h->baselines.n) KB_CUDA(h, cudaMemcpyAsync(h->baselines.p, h->sv_base.p, sizeof(double) * h->baselines.n, cudaMemcpyDeviceToDevice, s));
  // the sets that existed when the state was saved (a set appended since then keeps its pose; one removed since then is gone)
  const size_t n = std::min(h->saved_sets, h->set_poses.n);
  if (n) KB_CUDA(h, cudaMemcpyAsync(h->set_poses.p, h->sv_sets.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  ++h->state_version;
  h->built = h->solved = h->has_backup = false;
  return KB_OK;
}

// ---- state setters: the design variables live on the host when an unmodified Optimizer2 drives the solver ------------------
kb_status kb_set_state(kb_handle* h, const double* cam_params, const double* baselines, const double* set_poses) {
  KB_CUDA(h, cudaSetDevice(h->device));
  if (cam_params)
    KB_CUDA(h, cudaMemcpyAsync(h->cam_params.p, cam_params, sizeof(double) * h->cam_params.n, cudaMemcpyHostToDevice, h->stream));
  if (baselines && h->baselines.n)
    KB_CUDA(h, cudaMemcpyAsync(h->baselines.p, baselines, sizeof(double) * h->baselines.n, cudaMemcpyHostToDevice, h->stream));
  if (set_poses && h->set_poses.n)
    KB_CUDA(h, cudaMemcpyAsync(h->set_poses.p, set_poses + (h->presharded ? 0 : (size_t)KB_POSE_STRIDE * h->set_lo), sizeof(double) * h->set_poses.n,
                               cudaMemcpyHostToDevice, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));  // the caller's arrays may be pageable and may change right after the call
  ++h->state_version;     // any cached (speculative) linearisation belongs to the old state
  h->has_backup = false;  // the device-side backup belongs to the old state: reverting is the host's business now
  return KB_OK;
}
kb_status kb_set_camera_params(kb_handle* h, const double* cam_params) {
  if (!cam_params) return fail(h, KB_ERR_INVALID_ARGUMENT, "null camera parameters");
  return kb_set_state(h, cam_params, nullptr, nullptr);
}
kb_status kb_set_baselines(kb_handle* h, const double* baselines) {
  if (!baselines && h->n_cams > 1) return fail(h, KB_ERR_INVALID_ARGUMENT, "null baselines");
  return kb_set_state(h, nullptr, baselines, nullptr);
}
kb_status kb_set_set_poses(kb_handle* h, const double* set_poses) {
  if (!set_poses && h->set_poses.n) return fail(h, KB_ERR_INVALID_ARGUMENT, "null set poses");
  return kb_set_state(h, nullptr, nullptr, set_poses);
}
kb_status kb_set_conditioner(kb_handle* h, const double* diag) {
  if (!diag) return fail(h, KB_ERR_INVALID_ARGUMENT, "null conditioner");
  for (int64_t i = 1; i < h->jcols; ++i)
    if (diag[i] != diag[0]) return fail(h, KB_ERR_INVALID_ARGUMENT, "only a constant diagonal conditioner is supported (every trust-region policy of the reference installs a constant one)");
  return kb_set_constant_conditioner(h, h->jcols > 0 ? diag[0] : 0.0);
}

kb_status kb_get_camera_params(kb_handle* h, double* out) {
  KB_CUDA(h, cudaSetDevice(h->device));
  KB_CUDA(h, cudaMemcpyAsync(out, h->cam_params.p, sizeof(double) * h->cam_params.n, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}
kb_status kb_get_baselines(kb_handle* h, double* out) {
  KB_CUDA(h, cudaSetDevice(h->device));
  if (h->baselines.n) KB_CUDA(h, cudaMemcpyAsync(out, h->baselines.p, sizeof(double) * h->baselines.n, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}
kb_status kb_get_set_poses(kb_handle* h, double* out) {
  KB_CUDA(h, cudaSetDevice(h->device));
  if (h->set_poses.n)
    KB_CUDA(h, cudaMemcpyAsync(out + (h->presharded ? 0 : (size_t)KB_POSE_STRIDE * h->set_lo), h->set_poses.p, sizeof(double) * h->set_poses.n, cudaMemcpyDeviceToHost, h->stream));
  KB_CUDA(h, cudaStreamSynchronize(h->stream));
  return KB_OK;
}

}  // extern "C"
