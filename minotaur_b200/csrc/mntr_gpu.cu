// mntr_gpu.cu -- C ABI of libmntr_gpu.so (include/mntr_gpu.h): context, problem upload
// (CSR padding, CSC, wavefront levels), dispatch of the tighten kernels, result download.
// Host C++ only orchestrates; all bound arithmetic runs in the CUDA kernels.  There is no CPU
// fallback: every entry point fails when CUDA is unusable.
#include "../../include/mntr_gpu.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>

#include "cgraph.cuh"
#include "device_problem.cuh"
#include "kernels.h"

using namespace mntr;

// The handful of NCCL declarations this file needs, spelled out so that the module builds without NCCL's headers
// (the library itself is bound at run time with dlopen, see nccl_api()).  Values are NCCL 2.x's stable ABI (nccl.h).
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclUint64 = 5, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}

// host mirror of the single-box kernel's control block (SingleWs::ring/status/counters/bar)
constexpr int kTraceWords = 64 + 256 * 16 + 256 * 16 * 8;   // MNTR_GPU_TRACE buffer: phases, one warp's passes, barrier arrivals
struct SingleCtrl {
  unsigned sync[4]; int32_t status[8]; unsigned long long counters[2]; unsigned done; unsigned pad[15];
  // the loop's verdict, or MNTR_INFEAS_BOUNDS when the bound check of the last round's moved variables failed
  int verdict() const { return status[6] ? status[6] : ((sync[2] & mntr::kCtlFinalCross) ? 1 : 0); }
};
static_assert(sizeof(SingleCtrl) == 128, "control block layout");
static_assert(MNTR_GPU_MAX_TAPE == mntr::kMaxTape, "include/mntr_gpu.h and cgraph.cuh disagree on the longest tape");

struct mntr_gpu_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  char err[512] = {0};

  // ---- linear rows ----
  bool lin_loaded = false;
  int32_t m = 0, n = 0;
  int64_t nnz = 0, nnz_padded = 0;
  std::vector<uint8_t> h_var_type;   // host copy: integer bit of the stored columns
  std::vector<int32_t> h_perm;       // stored row -> caller's row (rows are stored in wavefront-level order)
  int lanes_per_row = 8;             // sub-warp group size of the per-round kernels
  bool no_zero_copy = false;         // MNTR_GPU_NO_ZEROCOPY=1: always stage pinned host boxes through device copies
  LinDev lin{};
  std::vector<void *> lin_allocs, cut_allocs, qrel_allocs;
  QRelDev qrel{};                    // QuadHandler relations (mntr_gpu_load_quad_relations)
  bool qrel_loaded = false;

  // ---- cgraph tapes ----
  bool nl_loaded = false;            // tapes and / or QuadraticFunction constraints are on the device
  bool tapes_loaded = false, quad_loaded = false;
  NlDev nl{};
  std::vector<void *> nl_allocs, quad_allocs;

  // ---- single-box workspace ----
  SingleWs sws{};
  double *d_lb = nullptr, *d_ub = nullptr;   // [n] staging of one box
  void *d_ctrl = nullptr;                    // SingleCtrl
  std::vector<void *> single_allocs;

  // ---- batch workspace ----
  int64_t batch_ld = 0;        // capacity in boxes (multiple of 32)
  double2 *d_boxes = nullptr;
  uint32_t *d_rowflag = nullptr, *d_varflag = nullptr;
  int32_t *d_prepared = nullptr;                // [1] BatchIo::prepared word of the batch built last (0 = root box clean)
  const void *prepared_boxes = nullptr;         // the batch boxes_from_deltas built and nobody has touched since
  int32_t prepared_n = 0;
  unsigned char *d_tstate = nullptr;
  int32_t *d_verdict = nullptr, *d_rounds = nullptr;
  long long *d_nnzb = nullptr;
  unsigned long long *d_nl_evals = nullptr;   // [1] CGraph evaluations of the last batch call
  // grow-only device scratch of the node-batch calls (root, deltas, mod counts / offsets / tuples): allocating and
  // freeing hundreds of MB per call costs tens of milliseconds on the host, unpredictably
  struct DevBuf { void *p = nullptr; size_t cap = 0; };
  DevBuf nb_buf[16];
  double *d_stage_lb = nullptr, *d_stage_ub = nullptr;
  int64_t stage_boxes = 0;

  // ---- per-round workspace (row-partitioned multi-GPU mode) ----
  RoundsWs rws{};
  int32_t *h_ctrl = nullptr;          // pinned mirror of rws.ctrl + counters
  SingleCtrl *h_single = nullptr;     // pinned, mapped: the fixpoint kernel's last block writes its control block here
  bool ctrl_clean = false;            // the device control block is zero (left so by the previous launch)
  bool force_rounds = false;          // MNTR_GPU_ROUNDS=1: per-round kernels even without a communicator
  BoundMsg *h_xhdr = nullptr;         // pinned: the ranks' message headers of the sparse bound exchange
  int32_t *h_progress = nullptr;      // pinned, mapped: {round finished, stop, verdict, changed} written by rounds_finalize_kernel
  std::vector<cudaEvent_t> round_ev;  // 4 events per timed round (rows / exchange / vars boundaries)
  // bound exchange over NVLink peer memory (CUDA IPC between the ranks' processes)
  void *p2p_region = nullptr;         // this rank's inbox + tags (exported to the peers)
  std::vector<void *> p2p_peers;      // peers' regions as mapped into this process (own entry nullptr)
  void *p2p_dev_tables = nullptr;     // device arrays: peers' inbox / tag pointers
  int64_t p2p_n = -1;                 // number of variables the region was sized for
  bool p2p_failed = false;            // IPC is unavailable: NCCL exchange
  unsigned p2p_tag = 0;               // rounds exchanged so far (tags are monotone over the calls)

  // ---- NCCL communicator (resolved at run time with dlopen: no link-time dependency) ----
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;

  mntr_gpu_stats stats{};
};


namespace {

// NCCL entry points, bound with dlopen("libnccl.so.2") on first use.  In a process that already loaded an
// NCCL (e.g. PyTorch's bundled one) the same library instance is reused.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl_api()
{
  static NcclApi api;
  if (api.handle) return api;
  // an NCCL the process has already loaded (e.g. PyTorch's bundled one) is reused; otherwise the system's is loaded
  // with RTLD_LOCAL, so that its symbols cannot shadow those of a different NCCL a later import brings along
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int k = 0; names[k] && !api.handle; ++k) api.handle = dlopen(names[k], RTLD_NOW | RTLD_NOLOAD);
  for (int k = 0; names[k] && !api.handle; ++k) api.handle = dlopen(names[k], RTLD_NOW | RTLD_LOCAL);
  if (!api.handle) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
  api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
  return api;
}

int fail(mntr_gpu_ctx *c, int code, const char *fmt, ...)
{
  if (c) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c->err, sizeof(c->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define NC(call)                                                                              \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess)                                                                   \
      return fail(ctx, MNTR_E_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(ctx, MNTR_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                  __FILE__, __LINE__);                                                        \
  } while (0)

template <class T>
int dev_upload(mntr_gpu_ctx *ctx, std::vector<void *> &owner, const T *host, size_t count, const T **out)
{
  T *d = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CU(cudaMalloc((void **)&d, bytes));
  owner.push_back(d);
  if (count) CU(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *out = d;
  return MNTR_OK;
}

void free_all(std::vector<void *> &v)
{
  for (void *p : v) cudaFree(p);
  v.clear();
}

// closes the peers' regions and frees this rank's (collective in effect: every rank does it before the region is reused)
void p2p_teardown(mntr_gpu_ctx *c)
{
  for (void *p : c->p2p_peers) if (p) cudaIpcCloseMemHandle(p);
  c->p2p_peers.clear();
  if (c->p2p_region) cudaFree(c->p2p_region);
  if (c->p2p_dev_tables) cudaFree(c->p2p_dev_tables);
  c->p2p_region = nullptr; c->p2p_dev_tables = nullptr; c->p2p_n = -1;
  c->rws.p2p = 0; c->rws.inbox = nullptr; c->rws.inbox_tag = nullptr; c->rws.peer_inbox = nullptr; c->rws.peer_tag = nullptr;
  (void)cudaGetLastError();
}

void free_stage(mntr_gpu_ctx *c)
{
  cudaFree(c->d_stage_lb); cudaFree(c->d_stage_ub);
  c->d_stage_lb = nullptr; c->d_stage_ub = nullptr; c->stage_boxes = 0;
}

// grow-only scratch slot k of at least `bytes`
int scratch(mntr_gpu_ctx *ctx, int k, size_t bytes, void **out)
{
  mntr_gpu_ctx::DevBuf &b = ctx->nb_buf[k];
  bytes = std::max<size_t>(bytes, 16);
  if (b.cap < bytes) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 4;            // headroom: batches of similar size do not reallocate
    if (cudaMalloc(&b.p, want) != cudaSuccess) {
      (void)cudaGetLastError();
      b.p = nullptr;
      return fail(ctx, MNTR_E_NOMEM, "node batch: out of device memory (%zu bytes)", want);
    }
    b.cap = want;
  }
  *out = b.p;
  return MNTR_OK;
}

void free_scratch(mntr_gpu_ctx *c)
{
  for (auto &b : c->nb_buf) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }
}

void free_batch(mntr_gpu_ctx *c)
{
  cudaFree(c->d_boxes); cudaFree(c->d_rowflag); cudaFree(c->d_varflag); cudaFree(c->d_tstate); cudaFree(c->d_verdict); cudaFree(c->d_rounds);
  cudaFree(c->d_nnzb); cudaFree(c->d_nl_evals); cudaFree(c->d_prepared);
  c->d_prepared = nullptr; c->prepared_boxes = nullptr; c->prepared_n = 0;
  c->d_nl_evals = nullptr;
  c->d_boxes = nullptr; c->d_rowflag = nullptr; c->d_verdict = nullptr; c->d_rounds = nullptr;
  c->d_nnzb = nullptr; c->d_varflag = nullptr; c->d_tstate = nullptr;
  c->batch_ld = 0;
}

int ensure_batch(mntr_gpu_ctx *ctx, int32_t n_boxes, bool need_boxes)
{
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  if (ld <= ctx->batch_ld && (!need_boxes || ctx->d_boxes)) return MNTR_OK;
  free_batch(ctx);
  const int64_t tiles = ld / 32;
  if (need_boxes)
    CU(cudaMalloc((void **)&ctx->d_boxes, sizeof(double2) * (size_t)std::max<int64_t>(1, (int64_t)ctx->n * ld)));
  CU(cudaMalloc((void **)&ctx->d_rowflag, sizeof(uint32_t) * (size_t)std::max<int64_t>(1, (int64_t)ctx->m * tiles)));
  CU(cudaMalloc((void **)&ctx->d_varflag, sizeof(uint32_t) * (size_t)std::max<int64_t>(1, (int64_t)ctx->n * tiles)));
  CU(cudaMalloc((void **)&ctx->d_tstate, (size_t)kTileStateBytes * (size_t)tiles));
  CU(cudaMalloc((void **)&ctx->d_verdict, sizeof(int32_t) * (size_t)ld));
  CU(cudaMalloc((void **)&ctx->d_rounds, sizeof(int32_t) * (size_t)ld));
  CU(cudaMalloc((void **)&ctx->d_nnzb, sizeof(long long) * (size_t)ld));
  CU(cudaMalloc((void **)&ctx->d_nl_evals, sizeof(unsigned long long)));
  CU(cudaMalloc((void **)&ctx->d_prepared, sizeof(int32_t)));
  ctx->batch_ld = ld;
  return MNTR_OK;
}

// box-major staging for host <-> engine-layout transposes: ~256 MiB per array
int ensure_stage(mntr_gpu_ctx *ctx)
{
  if (ctx->stage_boxes > 0) return MNTR_OK;
  int64_t sb = ((int64_t)256 << 20) / std::max<int64_t>(8, (int64_t)ctx->n * 8);
  sb = std::max<int64_t>(1, std::min<int64_t>(sb, 1 << 20));
  if (sb >= 32) sb = sb / 32 * 32;
  CU(cudaMalloc((void **)&ctx->d_stage_lb, sizeof(double) * (size_t)std::max<int64_t>(1, sb * ctx->n)));
  CU(cudaMalloc((void **)&ctx->d_stage_ub, sizeof(double) * (size_t)std::max<int64_t>(1, sb * ctx->n)));
  ctx->stage_boxes = sb;
  return MNTR_OK;
}

mntr_gpu_options resolve_opts(const mntr_gpu_options *o, int32_t n_boxes)
{
  mntr_gpu_options r;
  r.rounding = MNTR_ROUND_DIRECTED;
  r.order = (n_boxes == 1) ? MNTR_ORDER_JACOBI : MNTR_ORDER_REFERENCE;
  r.loop = MNTR_LOOP_FIXPOINT;
  r.max_rounds = 0;
  r.handlers = MNTR_HANDLERS_ALL;
  r.flags = 0;
  r.reserved[0] = r.reserved[1] = 0;
  if (o) {
    r = *o;
    if (r.order < 0) r.order = (n_boxes == 1) ? MNTR_ORDER_JACOBI : MNTR_ORDER_REFERENCE;
  }
  return r;
}

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

}  // namespace

extern "C" {

int mntr_gpu_abi_version(void) { return MNTR_GPU_ABI_VERSION; }

int mntr_gpu_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int64_t mntr_gpu_box_ld(int32_t n_boxes) { return ((int64_t)std::max(n_boxes, 1) + 31) / 32 * 32; }

int mntr_gpu_create(int device, mntr_gpu_ctx **out)
{
  if (!out) return MNTR_E_ARG;
  *out = nullptr;
  mntr_gpu_ctx *ctx = new (std::nothrow) mntr_gpu_ctx();
  if (!ctx) return MNTR_E_NOMEM;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    // no CPU fallback: the engine is unusable without a CUDA device
    delete ctx;
    return MNTR_E_CUDA;
  }
  ctx->device = device;
  auto bail = [&](cudaError_t err) { (void)err; mntr_gpu_destroy(ctx); return MNTR_E_CUDA; };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e);
  ctx->sm_count = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) return bail(cudaErrorNotSupported);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
  for (auto &ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e);
  *out = ctx;
  return MNTR_OK;
}

void mntr_gpu_destroy(mntr_gpu_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) { nccl_api().CommDestroy(ctx->comm); ctx->comm = nullptr; }
  if (ctx->rws.xsend) cudaFree(ctx->rws.xsend);
  if (ctx->rws.xrecv) cudaFree(ctx->rws.xrecv);
  if (ctx->h_xhdr) cudaFreeHost(ctx->h_xhdr);
  if (ctx->h_ctrl) { cudaFreeHost(ctx->h_ctrl); ctx->h_ctrl = nullptr; }
  if (ctx->h_progress) { cudaFreeHost(ctx->h_progress); ctx->h_progress = nullptr; }
  for (cudaEvent_t e : ctx->round_ev) cudaEventDestroy(e);
  ctx->round_ev.clear();
  p2p_teardown(ctx);
  if (ctx->h_single) { cudaFreeHost(ctx->h_single); ctx->h_single = nullptr; }
  free_all(ctx->lin_allocs); free_all(ctx->cut_allocs); free_all(ctx->nl_allocs); free_all(ctx->quad_allocs); free_all(ctx->single_allocs);
  free_all(ctx->qrel_allocs);
  free_batch(ctx); free_stage(ctx); free_scratch(ctx);
  for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *mntr_gpu_last_error(const mntr_gpu_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int mntr_gpu_get_stats(const mntr_gpu_ctx *ctx, mntr_gpu_stats *out)
{
  if (!ctx || !out) return MNTR_E_ARG;
  *out = ctx->stats;
  return MNTR_OK;
}

int mntr_gpu_load_linear(mntr_gpu_ctx *ctx, int32_t m, int32_t n, const int32_t *row_ptr,
                         const int32_t *col, const double *val, const double *row_lb,
                         const double *row_ub, const uint8_t *var_type, const uint8_t *row_active)
{
  if (!ctx) return MNTR_E_ARG;
  if (m < 0 || n < 0 || !row_ptr || (m > 0 && (!row_lb || !row_ub)) || (n > 0 && !var_type))
    return fail(ctx, MNTR_E_ARG, "load_linear: null or negative argument");
  if (row_ptr[0] != 0) return fail(ctx, MNTR_E_ARG, "load_linear: row_ptr[0] != 0");
  CU(cudaSetDevice(ctx->device));
  free_all(ctx->lin_allocs); free_all(ctx->cut_allocs); free_all(ctx->single_allocs); free_batch(ctx); free_stage(ctx);
  free_all(ctx->qrel_allocs); ctx->qrel_loaded = false; ctx->qrel = QRelDev{};
  free_all(ctx->nl_allocs); free_all(ctx->quad_allocs);
  ctx->nl_loaded = ctx->tapes_loaded = ctx->quad_loaded = false;
  ctx->nl = NlDev{};
  ctx->lin_loaded = false;

  // ---- host flattening: drop |a| <= 1e-9 (LinearFunction.cpp:89-95) ----
  std::vector<int32_t> cptr0(m + 1, 0), ccol;     // compact CSR in the caller's row order
  std::vector<double> cval;
  ccol.reserve((size_t)row_ptr[m]); cval.reserve((size_t)row_ptr[m]);
  for (int32_t i = 0; i < m; ++i) {
    if (row_ptr[i + 1] < row_ptr[i]) return fail(ctx, MNTR_E_ARG, "load_linear: row_ptr not monotone at row %d", i);
    int32_t prev = -1;
    for (int32_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) {
      const int32_t j = col[t];
      if (j < 0 || j >= n) return fail(ctx, MNTR_E_ARG, "load_linear: column %d out of range in row %d", j, i);
      if (j <= prev) return fail(ctx, MNTR_E_ARG, "load_linear: columns not strictly ascending in row %d", i);
      prev = j;
      if (std::fabs(val[t]) <= kCoefDrop) continue;
      ccol.push_back(j); cval.push_back(val[t]);
    }
    cptr0[i + 1] = (int32_t)ccol.size();
  }
  const int64_t nnz = (int64_t)ccol.size();
  if (nnz + 3 * (int64_t)m > (int64_t)INT32_MAX - 4) return fail(ctx, MNTR_E_UNSUPPORTED, "load_linear: more than 2^31 entries");

  // ---- wavefront levels of the reference's index-ordered in-place sweep: a row's level is one more
  //      than the highest level of an EARLIER row sharing a variable; rows of one level are pairwise
  //      variable-disjoint, so running levels in order reproduces the sequential sweep exactly ----
  std::vector<int32_t> last(n, -1), level(m, 0);
  int32_t n_levels = 0;
  for (int32_t i = 0; i < m; ++i) {
    if (row_active && !row_active[i]) { level[i] = -1; continue; }
    int32_t lev = 0;
    for (int32_t t = cptr0[i]; t < cptr0[i + 1]; ++t) lev = std::max(lev, last[ccol[t]] + 1);
    level[i] = lev;
    for (int32_t t = cptr0[i]; t < cptr0[i + 1]; ++t) last[ccol[t]] = lev;
    n_levels = std::max(n_levels, lev + 1);
  }
  // rows are STORED in (level, index) order; deleted rows go last and are never scheduled
  std::vector<int32_t> lptr(n_levels + 2, 0), perm((size_t)std::max(m, 1));
  int32_t n_sched = 0;
  for (int32_t i = 0; i < m; ++i) if (level[i] >= 0) { lptr[level[i] + 2]++; ++n_sched; }
  for (int32_t l = 0; l < n_levels; ++l) lptr[l + 2] += lptr[l + 1];
  {
    int32_t tail = n_sched;
    for (int32_t i = 0; i < m; ++i) {
      if (level[i] >= 0) perm[lptr[level[i] + 1]++] = i;
      else perm[tail++] = i;
    }
  }

  // ---- padded CSR in stored order: rows padded to 4 entries (128-bit val and col loads) ----
  std::vector<int32_t> prow(m + 1, 0), pcol;
  std::vector<double> pval;
  std::vector<int2> pinfo((size_t)std::max(m, 1));
  std::vector<double2> pbnd((size_t)std::max(m, 1));
  pcol.reserve((size_t)nnz + 3 * (size_t)m); pval.reserve((size_t)nnz + 3 * (size_t)m);
  for (int32_t q = 0; q < m; ++q) {
    const int32_t i = perm[q];
    prow[q] = (int32_t)pcol.size();
    for (int32_t t = cptr0[i]; t < cptr0[i + 1]; ++t) { pcol.push_back(ccol[t]); pval.push_back(cval[t]); }
    const int32_t cnt = cptr0[i + 1] - cptr0[i];
    while (pcol.size() % kRowPad) { pcol.push_back(pcol.empty() ? 0 : pcol.back()); pval.push_back(0.0); }
    const bool deleted = row_active && !row_active[i];
    pinfo[q] = make_int2(prow[q], deleted ? -1 : cnt);
    pbnd[q] = make_double2(row_lb[i], row_ub[i]);
  }
  prow[m] = (int32_t)pcol.size();

  // ---- CSC: var -> stored rows (Variable::cons_) ----
  std::vector<int32_t> cptr(n + 2, 0), crow((size_t)std::max<int64_t>(nnz, 1));
  for (int32_t q = 0; q < m; ++q)
    for (int32_t t = prow[q]; t < prow[q + 1]; ++t) if (pval[t] != 0.0) cptr[pcol[t] + 2]++;
  for (int32_t j = 0; j < n; ++j) cptr[j + 2] += cptr[j + 1];
  for (int32_t q = 0; q < m; ++q)
    for (int32_t t = prow[q]; t < prow[q + 1]; ++t) if (pval[t] != 0.0) crow[cptr[pcol[t] + 1]++] = q;

  // ---- upload ----
  LinDev &L = ctx->lin;
  L = LinDev{};
  L.cut_rhs = INFINITY; L.obj_ub = INFINITY;
  L.m = m; L.n = n; L.n_levels = n_levels;
  int rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, pinfo.data(), (size_t)m, &L.row_info))) return rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, pbnd.data(), (size_t)m, &L.row_bnd))) return rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, pcol.data(), pcol.size(), &L.col))) return rc;
  {
    std::vector<int32_t> pcolx(pcol);           // bit 31: the variable is integer (rounded by the reader)
    for (auto &c : pcolx) if (is_int_type(var_type[c])) c |= INT32_MIN;
    if ((rc = dev_upload(ctx, ctx->lin_allocs, pcolx.data(), pcolx.size(), &L.colx))) return rc;
  }
  ctx->h_var_type.assign(var_type, var_type + n);
  ctx->h_perm.assign(perm.begin(), perm.begin() + m);
  if ((rc = dev_upload(ctx, ctx->lin_allocs, pval.data(), pval.size(), &L.val))) return rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, var_type, (size_t)n, &L.var_type))) return rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, cptr.data(), (size_t)n + 1, &L.csc_ptr))) return rc;
  if ((rc = dev_upload(ctx, ctx->lin_allocs, crow.data(), (size_t)nnz, &L.csc_row))) return rc;
  L.csc_nnz = (int32_t)nnz;
  L.nnz_pad = (int32_t)pcol.size();
  if ((rc = dev_upload(ctx, ctx->lin_allocs, lptr.data(), (size_t)n_levels + 1, &L.level_ptr))) return rc;

  // ---- single-box workspace ----
  SingleWs &W = ctx->sws;
  auto dalloc = [&](void **p, size_t bytes) -> int {
    CU(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    ctx->single_allocs.push_back(*p);
    return MNTR_OK;
  };
  for (int k = 0; k < 2; ++k) {
    if ((rc = dalloc((void **)&W.box[k], sizeof(double2) * (size_t)n))) return rc;
    if ((rc = dalloc((void **)&W.due[k], sizeof(uint32_t) * (size_t)((m + 31) / 32 + 1)))) return rc;
    if ((rc = dalloc((void **)&W.touched[k], sizeof(uint32_t) * (size_t)((n + 31) / 32 + 1)))) return rc;
    if (k == 0 && (rc = dalloc((void **)&W.ever, sizeof(uint32_t) * (size_t)((n + 31) / 32 + 1)))) return rc;
  }
  if ((rc = dalloc((void **)&ctx->d_lb, sizeof(double) * (size_t)n))) return rc;
  if ((rc = dalloc((void **)&ctx->d_ub, sizeof(double) * (size_t)n))) return rc;
  if ((rc = dalloc(&ctx->d_ctrl, sizeof(SingleCtrl)))) return rc;
  W.sync = (unsigned *)ctx->d_ctrl;
  W.status = (int32_t *)((char *)ctx->d_ctrl + offsetof(SingleCtrl, status));
  W.counters = (unsigned long long *)((char *)ctx->d_ctrl + offsetof(SingleCtrl, counters));
  W.done = (unsigned *)((char *)ctx->d_ctrl + offsetof(SingleCtrl, done));
  if (!ctx->h_single) CU(cudaHostAlloc((void **)&ctx->h_single, sizeof(SingleCtrl), cudaHostAllocMapped));
  {
    void *mapped = nullptr;
    CU(cudaHostGetDevicePointer(&mapped, ctx->h_single, 0));
    W.result = (int32_t *)mapped;
  }
  ctx->ctrl_clean = false;
  W.trace = nullptr;
  W.stress_ns = 0u;
  if (const char *sb = getenv("MNTR_GPU_STRESS_BARRIER")) W.stress_ns = (unsigned)atoi(sb);   // test hook, see grid_barrier
  if (const char *tr = getenv("MNTR_GPU_TRACE")) {
    if (tr[0] == '1') { if ((rc = dalloc((void **)&W.trace, kTraceWords * sizeof(unsigned long long)))) return rc; }
  }

  // per-round workspace shares box / bits / list with the single-launch kernel
  RoundsWs &RW = ctx->rws;
  RW.box = W.box[0]; RW.bits = W.due[0];
  if ((rc = dalloc((void **)&RW.nlb, sizeof(double) * ((size_t)n + 1)))) return rc;
  if ((rc = dalloc((void **)&RW.nub, sizeof(double) * ((size_t)n + 1)))) return rc;
  if ((rc = dalloc((void **)&RW.ctrl, 128))) return rc;
  RW.counters = (unsigned long long *)((char *)RW.ctrl + 64);
  if ((rc = dalloc((void **)&RW.tbits, sizeof(uint32_t) * (size_t)((n + 31) / 32 + 1)))) return rc;
  if ((rc = dalloc((void **)&RW.ebits, sizeof(uint32_t) * (size_t)((n + 31) / 32 + 1)))) return rc;
  if ((rc = dalloc((void **)&RW.tlist, sizeof(int32_t) * (size_t)std::max(n, 1)))) return rc;
  if ((rc = dalloc((void **)&RW.elist, sizeof(int32_t) * (size_t)std::max(n, 1)))) return rc;
  if (!ctx->h_ctrl) CU(cudaMallocHost((void **)&ctx->h_ctrl, 128));
  if (!ctx->h_progress) CU(cudaHostAlloc((void **)&ctx->h_progress, 64, cudaHostAllocMapped));
  {
    void *mapped = nullptr;
    CU(cudaHostGetDevicePointer(&mapped, ctx->h_progress, 0));
    RW.progress = (int32_t *)mapped;
  }
  RW.rank = ctx->rank; RW.n_ranks = ctx->n_ranks;
  if (const char *fr = getenv("MNTR_GPU_ROUNDS")) ctx->force_rounds = fr[0] == '1';
  if (const char *zc = getenv("MNTR_GPU_NO_ZEROCOPY")) ctx->no_zero_copy = zc[0] == '1';

  {   // sub-warp group size of the per-round kernels from the mean row length (four entries per lane per step)
    const double mean = m > 0 ? (double)nnz / m : 0.0;
    int g = 2;
    while (g < 32 && 4 * g < mean + 0.5) g *= 2;
    ctx->lanes_per_row = g;
  }
  ctx->m = m; ctx->n = n; ctx->nnz = nnz; ctx->nnz_padded = (int64_t)pcol.size();
  CU(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
  ctx->lin_loaded = true;
  return MNTR_OK;
}

int mntr_gpu_load_cgraph(mntr_gpu_ctx *ctx, int32_t n_cons, const int32_t *tape_ptr, const uint8_t *op,
                         const int32_t *arg0, const int32_t *arg1, const double *cnst, const int32_t *child,
                         const int32_t *lin_ptr, const int32_t *lin_col, const double *lin_val, const double *c_lb,
                         const double *c_ub)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "load_cgraph: call load_linear first (m may be 0)");
  CU(cudaSetDevice(ctx->device));
  free_all(ctx->nl_allocs);
  const NlDev keep_quad = ctx->nl;        // the QuadraticFunction constraints are loaded separately and stay
  auto restore_quad = [&](NlDev &N) {
    N.n_quad = ctx->quad_loaded ? keep_quad.n_quad : 0;
    N.q_ptr = keep_quad.q_ptr; N.q_v1 = keep_quad.q_v1; N.q_v2 = keep_quad.q_v2; N.q_coef = keep_quad.q_coef;
    N.q_lin_ptr = keep_quad.q_lin_ptr; N.q_lin_col = keep_quad.q_lin_col; N.q_lin_val = keep_quad.q_lin_val;
    N.q_lb = keep_quad.q_lb; N.q_ub = keep_quad.q_ub;
  };
  ctx->tapes_loaded = false;
  ctx->nl = NlDev{};
  restore_quad(ctx->nl);
  ctx->nl_loaded = ctx->quad_loaded;
  if (n_cons == 0) return MNTR_OK;
  if (n_cons < 0 || !tape_ptr || !op || !arg0 || !arg1 || !cnst || !lin_ptr || !c_lb || !c_ub)
    return fail(ctx, MNTR_E_ARG, "load_cgraph: null or negative argument");
  const int32_t n = ctx->n;

  // ---- validate, and collect the variables each constraint reads / writes ----
  std::vector<int32_t> level(n_cons, 0), lastW(n, -1), lastR(n, -1);
  int32_t n_levels = 0, max_nodes = 0;
  for (int32_t c = 0; c < n_cons; ++c) {
    const int32_t b = tape_ptr[c], nn = tape_ptr[c + 1] - b;
    if (nn < 1 || nn > kMaxTape)
      return fail(ctx, MNTR_E_UNSUPPORTED, "load_cgraph: constraint %d has %d tape nodes (supported: 1..%d)", c, nn, kMaxTape);
    max_nodes = std::max(max_nodes, nn);
    int32_t lev = 0;
    for (int32_t i = 0; i < nn; ++i) {
      const int o = op[b + i];
      if (o > OpVar) return fail(ctx, MNTR_E_ARG, "load_cgraph: bad opcode %d in constraint %d", o, c);
      if (o == OpVar) {
        const int32_t v = arg0[b + i];
        if (v < 0 || v >= n) return fail(ctx, MNTR_E_ARG, "load_cgraph: variable %d out of range in constraint %d", v, c);
        lev = std::max(lev, std::max(lastW[v], lastR[v]) + 1);     // the tape's variables are read AND written
      } else if (o == OpNum || o == OpInt) {
      } else if (o == OpSumList) {
        if (!child || arg0[b + i] < 0 || arg1[b + i] < arg0[b + i])
          return fail(ctx, MNTR_E_ARG, "load_cgraph: bad child list in constraint %d", c);
        for (int32_t q = arg0[b + i]; q < arg1[b + i]; ++q)
          if (child[q] < 0 || child[q] >= i) return fail(ctx, MNTR_E_ARG, "load_cgraph: child after parent in constraint %d", c);
      } else {
        if (arg0[b + i] < 0 || arg0[b + i] >= i || arg1[b + i] >= i)
          return fail(ctx, MNTR_E_ARG, "load_cgraph: operand after operator in constraint %d", c);
      }
    }
    if (op[b + nn - 1] == OpVar || op[b + nn - 1] == OpNum || op[b + nn - 1] == OpInt)
      return fail(ctx, MNTR_E_ARG, "load_cgraph: constraint %d has no operator node", c);
    for (int32_t q = lin_ptr[c]; q < lin_ptr[c + 1]; ++q) {
      const int32_t v = lin_col[q];
      if (v < 0 || v >= n) return fail(ctx, MNTR_E_ARG, "load_cgraph: linear column out of range in constraint %d", c);
      lev = std::max(lev, lastW[v] + 1);                            // the linear part is only read
    }
    // wavefront level of the in-place, index-ordered sweep (NlPresHandler.cpp:1686-1806): after every
    // earlier constraint that writes something this one touches, or reads something this one writes
    level[c] = lev;
    for (int32_t i = 0; i < nn; ++i)
      if (op[b + i] == OpVar) { lastW[arg0[b + i]] = lev; lastR[arg0[b + i]] = std::max(lastR[arg0[b + i]], lev); }
    for (int32_t q = lin_ptr[c]; q < lin_ptr[c + 1]; ++q) lastR[lin_col[q]] = std::max(lastR[lin_col[q]], lev);
    n_levels = std::max(n_levels, lev + 1);
  }

  // ---- store the constraints in (level, index) order ----
  std::vector<int32_t> lptr(n_levels + 2, 0), perm(n_cons);
  for (int32_t c = 0; c < n_cons; ++c) lptr[level[c] + 2]++;
  for (int32_t l = 0; l < n_levels; ++l) lptr[l + 2] += lptr[l + 1];
  for (int32_t c = 0; c < n_cons; ++c) perm[lptr[level[c] + 1]++] = c;
  std::vector<int32_t> tp(n_cons + 1, 0), lp(n_cons + 1, 0), a0v, a1v, chv, lcv;
  std::vector<uint8_t> opv;
  std::vector<double> cnv, lvv, clb(n_cons), cub(n_cons);
  bool all_shaped = true;         // every tape is [Var,Var,Mult] or [Var,Var,Sqr,Sqr,SumList(2,3)] (the rule of stage_batch, cgraph.cuh)
  for (int32_t q = 0; q < n_cons; ++q) {
    const int32_t c = perm[q], b = tape_ptr[c], nn = tape_ptr[c + 1] - b;
    {
      const uint8_t *o = op + b; const int32_t *x0 = arg0 + b, *x1 = arg1 + b;
      int nv = 0;
      while (nv < nn && o[nv] == OpVar) ++nv;
      const bool bil = nv == 2 && nn == 3 && o[2] == OpMult && (unsigned)x0[2] < 2u && (unsigned)x1[2] < 2u && x0[2] != x1[2];
      const bool ssq = nv == 2 && nn == 5 && o[2] == OpSqr && o[3] == OpSqr && o[4] == OpSumList && (unsigned)x0[2] < 2u &&
                       (unsigned)x0[3] < 2u && x1[4] - x0[4] == 2 && child[x0[4]] == 2 && child[x0[4] + 1] == 3;
      all_shaped = all_shaped && (bil || ssq);
    }
    for (int32_t i = 0; i < nn; ++i) {
      opv.push_back(op[b + i]); cnv.push_back(cnst[b + i]);
      if (op[b + i] == OpSumList) {
        a0v.push_back((int32_t)chv.size());
        for (int32_t k = arg0[b + i]; k < arg1[b + i]; ++k) chv.push_back(child[k]);
        a1v.push_back((int32_t)chv.size());
      } else { a0v.push_back(arg0[b + i]); a1v.push_back(arg1[b + i]); }
    }
    tp[q + 1] = (int32_t)opv.size();
    for (int32_t k = lin_ptr[c]; k < lin_ptr[c + 1]; ++k) { lcv.push_back(lin_col[k]); lvv.push_back(lin_val[k]); }
    lp[q + 1] = (int32_t)lcv.size();
    clb[q] = c_lb[c]; cub[q] = c_ub[c];
  }

  NlDev &N = ctx->nl;
  N = NlDev{};
  restore_quad(N);
  N.n_cons = n_cons; N.max_nodes = max_nodes; N.n_levels = n_levels;
  N.all_shaped = all_shaped ? 1 : 0;
  if (const char *e = getenv("MNTR_GPU_NO_SHAPED_KERNEL")) if (e[0] == '1') N.all_shaped = 0;
  int rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, tp.data(), tp.size(), &N.tape_ptr))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, opv.data(), opv.size(), &N.op))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, a0v.data(), a0v.size(), &N.arg0))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, a1v.data(), a1v.size(), &N.arg1))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, cnv.data(), cnv.size(), &N.cnst))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, chv.data(), chv.size(), &N.child))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, lp.data(), lp.size(), &N.lin_ptr))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, lcv.data(), lcv.size(), &N.lin_col))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, lvv.data(), lvv.size(), &N.lin_val))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, clb.data(), clb.size(), &N.c_lb))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, cub.data(), cub.size(), &N.c_ub))) return rc;
  if ((rc = dev_upload(ctx, ctx->nl_allocs, lptr.data(), (size_t)n_levels + 1, &N.level_ptr))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->tapes_loaded = true;
  ctx->nl_loaded = true;
  return MNTR_OK;
}

int mntr_gpu_load_quad(mntr_gpu_ctx *ctx, int32_t n_quad, const int32_t *q_ptr, const int32_t *v1, const int32_t *v2,
                       const double *coef, const int32_t *lin_ptr, const int32_t *lin_col, const double *lin_val,
                       const double *q_lb, const double *q_ub)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "load_quad: call load_linear first (m may be 0)");
  CU(cudaSetDevice(ctx->device));
  free_all(ctx->quad_allocs);
  ctx->quad_loaded = false;
  ctx->nl.n_quad = 0;
  ctx->nl_loaded = ctx->tapes_loaded;
  if (n_quad == 0) return MNTR_OK;
  if (n_quad < 0 || !q_ptr || !v1 || !v2 || !coef || !lin_ptr || !q_lb || !q_ub || q_ptr[0] != 0 || lin_ptr[0] != 0)
    return fail(ctx, MNTR_E_ARG, "load_quad: null or negative argument");
  const int32_t n = ctx->n;
  std::vector<int32_t> qp(1, 0), a1, a2;
  std::vector<double> cf;
  for (int32_t q = 0; q < n_quad; ++q) {
    if (q_ptr[q + 1] < q_ptr[q] || lin_ptr[q + 1] < lin_ptr[q]) return fail(ctx, MNTR_E_ARG, "load_quad: offsets not monotone at constraint %d", q);
    long long prev = -1;
    for (int32_t t = q_ptr[q]; t < q_ptr[q + 1]; ++t) {
      if (v1[t] < 0 || v2[t] >= n || v1[t] > v2[t]) return fail(ctx, MNTR_E_ARG, "load_quad: bad variable pair in constraint %d (need 0 <= v1 <= v2 < n)", q);
      const long long key = (long long)v1[t] * n + v2[t];
      if (key <= prev) return fail(ctx, MNTR_E_ARG, "load_quad: pairs not strictly ascending in constraint %d", q);
      prev = key;
      if (!(std::fabs(coef[t]) >= 1e-8)) continue;            // QuadraticFunction::addTerm, etol_
      a1.push_back(v1[t]); a2.push_back(v2[t]); cf.push_back(coef[t]);
    }
    qp.push_back((int32_t)a1.size());
    for (int32_t k = lin_ptr[q]; k < lin_ptr[q + 1]; ++k)
      if (!lin_col || !lin_val || lin_col[k] < 0 || lin_col[k] >= n) return fail(ctx, MNTR_E_ARG, "load_quad: bad linear term in constraint %d", q);
  }
  NlDev &N = ctx->nl;
  int rc;
  const int32_t n_lin = lin_ptr[n_quad];
  if ((rc = dev_upload(ctx, ctx->quad_allocs, qp.data(), qp.size(), &N.q_ptr))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, a1.data(), a1.size(), &N.q_v1))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, a2.data(), a2.size(), &N.q_v2))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, cf.data(), cf.size(), &N.q_coef))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, lin_ptr, (size_t)n_quad + 1, &N.q_lin_ptr))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, lin_col, (size_t)n_lin, &N.q_lin_col))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, lin_val, (size_t)n_lin, &N.q_lin_val))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, q_lb, (size_t)n_quad, &N.q_lb))) return rc;
  if ((rc = dev_upload(ctx, ctx->quad_allocs, q_ub, (size_t)n_quad, &N.q_ub))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  N.n_quad = n_quad;
  ctx->quad_loaded = true;
  ctx->nl_loaded = true;
  return MNTR_OK;
}

int mntr_gpu_update_row_bounds(mntr_gpu_ctx *ctx, int32_t m, const double *row_lb, const double *row_ub)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "update_row_bounds: call load_linear first");
  if (m != ctx->m || (m > 0 && (!row_lb || !row_ub))) return fail(ctx, MNTR_E_ARG, "update_row_bounds: row count differs from the loaded problem");
  if (m == 0) return MNTR_OK;
  CU(cudaSetDevice(ctx->device));
  std::vector<double2> pbnd((size_t)m);
  for (int32_t q = 0; q < m; ++q) pbnd[(size_t)q] = make_double2(row_lb[ctx->h_perm[(size_t)q]], row_ub[ctx->h_perm[(size_t)q]]);
  CU(cudaMemcpyAsync(const_cast<double2 *>(ctx->lin.row_bnd), pbnd.data(), sizeof(double2) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MNTR_OK;
}

int mntr_gpu_set_cutoff(mntr_gpu_ctx *ctx, int32_t k, const int32_t *col, const double *val, double rhs)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "set_cutoff: call load_linear first");
  CU(cudaSetDevice(ctx->device));
  free_all(ctx->cut_allocs);
  ctx->lin.cut_cnt = 0; ctx->lin.cut_col = nullptr; ctx->lin.cut_colx = nullptr; ctx->lin.cut_val = nullptr;
  ctx->lin.cut_bnd = nullptr; ctx->lin.cut_rhs = INFINITY; ctx->lin.obj_ub = INFINITY;
  if (k <= 0) return MNTR_OK;
  if (!col || !val) return fail(ctx, MNTR_E_ARG, "set_cutoff: null argument");
  std::vector<int32_t> c; std::vector<double> v;
  int32_t prev = -1;
  for (int32_t t = 0; t < k; ++t) {
    if (col[t] < 0 || col[t] >= ctx->n) return fail(ctx, MNTR_E_ARG, "set_cutoff: column %d out of range", col[t]);
    if (col[t] <= prev) return fail(ctx, MNTR_E_ARG, "set_cutoff: columns not strictly ascending");
    prev = col[t];
    if (std::fabs(val[t]) <= kCoefDrop) continue;       // LinearFunction drops these, LinearFunction.cpp:89-95
    c.push_back(col[t]); v.push_back(val[t]);
  }
  const int32_t cnt = (int32_t)c.size();
  if (cnt == 0) return MNTR_OK;
  while (c.size() % kRowPad) { c.push_back(c.back()); v.push_back(0.0); }
  std::vector<int32_t> cx(c);
  for (auto &j : cx) if (is_int_type(ctx->h_var_type[j])) j |= INT32_MIN;
  int rc;
  if ((rc = dev_upload(ctx, ctx->cut_allocs, c.data(), c.size(), &ctx->lin.cut_col))) return rc;
  if ((rc = dev_upload(ctx, ctx->cut_allocs, cx.data(), cx.size(), &ctx->lin.cut_colx))) return rc;
  if ((rc = dev_upload(ctx, ctx->cut_allocs, v.data(), v.size(), &ctx->lin.cut_val))) return rc;
  const double2 bnd = make_double2(-INFINITY, rhs);
  if ((rc = dev_upload(ctx, ctx->cut_allocs, &bnd, 1, &ctx->lin.cut_bnd))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->lin.cut_rhs = rhs;
  ctx->lin.cut_cnt = cnt;
  return MNTR_OK;
}

int mntr_gpu_set_incumbent(mntr_gpu_ctx *ctx, double best_value)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "set_incumbent: call load_linear first");
  if (best_value != best_value) return fail(ctx, MNTR_E_ARG, "set_incumbent: NaN");
  ctx->lin.obj_ub = best_value;
  return MNTR_OK;
}

// K1 on a device-resident box; leaves verdict/rounds/counters in the control block
static int run_single_dev(mntr_gpu_ctx *ctx, double *lb_dev, double *ub_dev, const mntr_gpu_options &o)
{
  if (!ctx->ctrl_clean) CU(cudaMemsetAsync(ctx->d_ctrl, 0, sizeof(SingleCtrl), ctx->stream));
  ctx->ctrl_clean = false;          // until the launch has been seen to complete
  CU(launch_single_jacobi(ctx->lin, ctx->sws, lb_dev, ub_dev, o.rounding == MNTR_ROUND_DIRECTED, o.max_rounds, o.loop, ctx->sm_count,
                           (o.flags & MNTR_FLAG_STAGED_ROWS) != 0, ctx->stream));
  return MNTR_OK;
}

static void account_single(mntr_gpu_ctx *ctx, const SingleCtrl &ctrl)
{
  ctx->stats.nnz_updates += (int64_t)ctrl.counters[0];
  ctx->stats.rows_evaluated += (int64_t)ctrl.counters[1];
  ctx->stats.n_changes += (int64_t)ctrl.status[2];
  ctx->stats.n_infeasible += ctrl.verdict() != 0;
  ctx->stats.max_rounds = std::max(ctx->stats.max_rounds, ctrl.status[1]);
}

// Peer-memory exchange of the row-partitioned mode: every rank exports ONE region (round tags + a double-buffered inbox
// with a slot per sender, each large enough for n candidates -- a message always fits, no overflow path) with CUDA IPC;
// the handles travel through one NCCL all-gather.  Collective: every rank calls it at the same point of its first
// tighten call after comm_init / load_linear.  Any failure (IPC unavailable, peers not reachable) leaves the NCCL
// exchange in place -- consistently on all ranks, because the outcome is all-reduced.
static int p2p_setup(mntr_gpu_ctx *ctx)
{
  NcclApi &nc = nccl_api();
  const int R = ctx->n_ranks;
  const int64_t n = ctx->n;
  const int64_t stride = n + 1;
  const size_t tag_bytes = 256;
  const size_t inbox_bytes = sizeof(BoundMsg) * (size_t)(2 * R) * (size_t)stride;
  int ok = 1;
  if (const char *e = getenv("MNTR_GPU_P2P")) ok = e[0] != '0';
  if (getenv("MNTR_GPU_SPARSE_XCHG")) ok = 0;        // the NCCL exchange was asked for explicitly (capacity / dense)
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&ctx->p2p_region, tag_bytes + inbox_bytes) != cudaSuccess) { (void)cudaGetLastError(); ctx->p2p_region = nullptr; ok = 0; }
  if (ok && cudaMemsetAsync(ctx->p2p_region, 0, tag_bytes, ctx->stream) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, ctx->p2p_region) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; }
  // all-gather {ok, handle}
  // ranks that are threads of ONE process (one context per device) cannot open each other's IPC handles: they take the
  // raw pointer and enable peer access instead
  struct Card { int ok; int device; long long pid; void *ptr; cudaIpcMemHandle_t h; };
  Card card; card.ok = ok; card.device = ctx->device; card.pid = (long long)getpid(); card.ptr = ctx->p2p_region; card.h = mine;
  Card *d_cards = nullptr;
  std::vector<Card> cards((size_t)R);
  CU(cudaMalloc((void **)&d_cards, sizeof(Card) * (size_t)(R + 1)));
  CU(cudaMemcpyAsync(d_cards + R, &card, sizeof(Card), cudaMemcpyHostToDevice, ctx->stream));
  NC(nc.AllGather(d_cards + R, d_cards, sizeof(Card), ncclChar, ctx->comm, ctx->stream));
  CU(cudaMemcpyAsync(cards.data(), d_cards, sizeof(Card) * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < R; ++r) ok = ok && cards[(size_t)r].ok;
  ctx->p2p_peers.assign((size_t)R, nullptr);
  std::vector<void *> same_process((size_t)R, nullptr);
  if (ok) {
    for (int r = 0; r < R && ok; ++r) {
      if (r == ctx->rank) continue;
      const Card &c = cards[(size_t)r];
      if (c.pid == (long long)getpid()) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, c.device) != cudaSuccess || !can) { (void)cudaGetLastError(); ok = 0; continue; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(c.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
        (void)cudaGetLastError();
        same_process[(size_t)r] = c.ptr;
        continue;
      }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, c.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = 0; }
      else ctx->p2p_peers[(size_t)r] = p;
    }
  }
  // everybody must have opened everybody: agree on the outcome
  double okd = ok ? 1.0 : 0.0, *d_okd = (double *)(d_cards + R);
  CU(cudaMemcpyAsync(d_okd, &okd, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NC(nc.AllReduce(d_okd, d_okd, 1, ncclDouble, ncclMin, ctx->comm, ctx->stream));
  CU(cudaMemcpyAsync(&okd, d_okd, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_cards);
  if (okd < 0.5) {
    p2p_teardown(ctx);
    ctx->p2p_failed = true;
    return MNTR_OK;
  }
  // device tables of the peers' inbox / tag pointers
  std::vector<void *> tab((size_t)(2 * R), nullptr);
  for (int r = 0; r < R; ++r) {
    char *base = (char *)(r == ctx->rank ? ctx->p2p_region : (same_process[(size_t)r] ? same_process[(size_t)r] : ctx->p2p_peers[(size_t)r]));
    tab[(size_t)r] = base + tag_bytes;         // inbox
    tab[(size_t)(R + r)] = base;               // tags
  }
  CU(cudaMalloc(&ctx->p2p_dev_tables, sizeof(void *) * (size_t)(2 * R)));
  CU(cudaMemcpy(ctx->p2p_dev_tables, tab.data(), sizeof(void *) * (size_t)(2 * R), cudaMemcpyHostToDevice));
  RoundsWs &W = ctx->rws;
  W.p2p = 1; W.inbox_stride = stride;
  W.inbox = (BoundMsg *)((char *)ctx->p2p_region + tag_bytes);
  W.inbox_tag = (unsigned *)ctx->p2p_region;
  W.peer_inbox = (BoundMsg *const *)ctx->p2p_dev_tables;
  W.peer_tag = (unsigned *const *)((void **)ctx->p2p_dev_tables + R);
  ctx->p2p_n = n;
  ctx->p2p_tag = 0;
  return MNTR_OK;
}

// Jacobi rounds as separate launches with the bound merge between the rows phase and the variables phase: the
// row-partitioned multi-GPU path (and, with MNTR_FLAG_PER_ROUND_KERNELS, the same kernels on one GPU).
// The loop runs on the DEVICE: a one-warp kernel ends every round (verdict, stop decision, progress words in mapped
// host memory) and every kernel enqueued after the stop returns at once; the host only keeps the stream fed -- it
// enqueues a round as soon as the round before the previous one has reported, polling the progress words without ever
// synchronising the stream.  Per-round work is proportional to the CHANGES (touched lists), not to n.
static int run_rounds_dev(mntr_gpu_ctx *ctx, double *lb_dev, double *ub_dev, const mntr_gpu_options &o,
                          int32_t *verdict, int32_t *rounds, int64_t *nnz_updates, double *lb_out = nullptr, double *ub_out = nullptr)
{
  // lb_out / ub_out: where the bounds that moved are written (default: in place)
  if (!lb_out) lb_out = lb_dev;
  if (!ub_out) ub_out = ub_dev;
  const LinDev &P = ctx->lin;
  const bool directed = o.rounding == MNTR_ROUND_DIRECTED;
  NcclApi &nc = nccl_api();
  cudaStream_t s = ctx->stream;
  ctx->rws.rank = ctx->rank; ctx->rws.n_ranks = ctx->n_ranks;
  if (ctx->comm && ctx->n_ranks > 1 && !ctx->p2p_failed && ctx->p2p_n != (int64_t)ctx->n) {
    if (ctx->p2p_region) p2p_teardown(ctx);
    int rc = p2p_setup(ctx);
    if (rc) return rc;
  }
  const RoundsWs &W = ctx->rws;
  const bool p2p = ctx->comm && W.p2p;
  volatile int32_t *prog = ctx->h_progress;
  prog[0] = 0; prog[1] = 0; prog[2] = 0; prog[3] = 0;
  CU(cudaMemsetAsync(W.ctrl, 0, 128, s));
  CU(launch_rounds_init(P, W, lb_dev, ub_dev, ctx->sm_count, s));
  const int kTimed = 64;                      // rounds with per-phase device timing
  while ((int)ctx->round_ev.size() < 4 * kTimed) {
    cudaEvent_t e; CU(cudaEventCreate(&e)); ctx->round_ev.push_back(e);
  }
  int enq = 0;                                // rounds enqueued
  auto enqueue_round = [&](int round) -> int {
    cudaEvent_t *ev = round <= kTimed ? &ctx->round_ev[(size_t)(4 * (round - 1))] : nullptr;
    if (ev) CU(cudaEventRecord(ev[0], s));
    CU(launch_rounds_rows(P, W, ctx->lanes_per_row, directed, round == 1, ctx->sm_count, s));
    if (ev) CU(cudaEventRecord(ev[1], s));
    if (p2p) {
      const unsigned tag = ctx->p2p_tag + (unsigned)round;
      CU(launch_rounds_push(P, W, tag, ctx->sm_count, s));
      CU(launch_rounds_pull(P, W, tag, ctx->sm_count, s));
    }
    if (ev) CU(cudaEventRecord(ev[2], s));
    CU(launch_rounds_vars_list(P, W, ctx->sm_count, s));
    CU(launch_rounds_finalize(P, W, o.max_rounds, o.loop, s));
    if (ev) CU(cudaEventRecord(ev[3], s));
    return MNTR_OK;
  };
  int finished = 0, stop = 0;
  if (!ctx->comm || p2p) {
    // ---- device-driven loop ----
    for (unsigned spins = 0;; ++spins) {
      finished = prog[0]; stop = prog[1];
      if (stop) break;
      if (enq - finished < 2) {
        int rc = enqueue_round(++enq);
        if (rc) return rc;
        continue;
      }
      if ((spins & 1023u) == 1023u) {           // a failed launch would never report: look at the stream now and then
        const cudaError_t q = cudaStreamQuery(s);
        if (q != cudaSuccess && q != cudaErrorNotReady)
          return fail(ctx, MNTR_E_CUDA, "per-round kernels: %s", cudaGetErrorString(q));
      }
    }
    // rounds enqueued beyond the stop are no-ops on every rank alike
  } else {
    // ---- NCCL fallback (no peer memory): the host reads the message sizes and chooses sparse all-gather / dense
    //      all-reduce per round (one host round trip per round) ----
    auto dense_merge = [&]() -> int {
      NC(nc.GroupStart());
      NC(nc.AllReduce(W.nlb, W.nlb, (size_t)P.n + 1, ncclDouble, ncclMax, ctx->comm, s));
      NC(nc.AllReduce(W.nub, W.nub, (size_t)P.n, ncclDouble, ncclMin, ctx->comm, s));
      NC(nc.GroupEnd());
      return MNTR_OK;
    };
    while (!stop) {
      const int round = ++enq;
      cudaEvent_t *ev = round <= kTimed ? &ctx->round_ev[(size_t)(4 * (round - 1))] : nullptr;
      if (ev) CU(cudaEventRecord(ev[0], s));
      CU(launch_rounds_rows(P, W, ctx->lanes_per_row, directed, round == 1, ctx->sm_count, s));
      if (ev) CU(cudaEventRecord(ev[1], s));
      int cap = 0, rc2;
      if (W.xcap > 0) {
        CU(launch_rounds_compact(P, W, W.xcap, ctx->sm_count, s));
        NC(nc.AllGather(W.xsend, W.xrecv, sizeof(BoundMsg), ncclChar, ctx->comm, s));
        CU(cudaMemcpyAsync(ctx->h_xhdr, W.xrecv, sizeof(BoundMsg) * (size_t)ctx->n_ranks, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        long long maxc = 0;
        for (int r = 0; r < ctx->n_ranks; ++r) maxc = std::max(maxc, ctx->h_xhdr[r].j);
        long long c2 = 64;
        while (c2 < maxc) c2 *= 2;
        c2 = std::min<long long>(c2, W.xcap);
        const double sparse_bytes = (double)sizeof(BoundMsg) * (double)(c2 + 1) * ctx->n_ranks;
        if (maxc <= W.xcap && sparse_bytes <= 0.5 * 16.0 * (double)P.n) cap = (int)c2;
      }
      if (cap > 0) {
        NC(nc.AllGather(W.xsend, W.xrecv, sizeof(BoundMsg) * ((size_t)cap + 1), ncclChar, ctx->comm, s));
        CU(launch_rounds_apply(P, W, ctx->rank, cap, ctx->sm_count, s));
        if (ev) CU(cudaEventRecord(ev[2], s));
        CU(launch_rounds_vars_list(P, W, ctx->sm_count, s));
        ++ctx->stats.sparse_rounds;
      } else {
        CU(launch_rounds_clear_list(W, ctx->sm_count, s));
        if ((rc2 = dense_merge())) return rc2;
        if (ev) CU(cudaEventRecord(ev[2], s));
        CU(launch_rounds_vars(P, W, ctx->sm_count, s));
      }
      CU(launch_rounds_finalize(P, W, o.max_rounds, o.loop, s));
      if (ev) CU(cudaEventRecord(ev[3], s));
      CU(cudaStreamSynchronize(s));
      stop = prog[1];
    }
  }
  CU(launch_rounds_finish(P, W, lb_out, ub_out, ctx->sm_count, s));
  CU(cudaMemcpyAsync(ctx->h_ctrl, W.ctrl, 128, cudaMemcpyDeviceToHost, s));
  unsigned long long cnt[2];
  if (ctx->comm) {     // nnz-updates / rows are per rank: sum them so every rank reports the job total
    NC(nc.AllReduce(W.counters, W.counters, 2, ncclUint64, ncclSum, ctx->comm, s));
    CU(cudaMemcpyAsync(cnt, W.counters, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  }
  CU(cudaStreamSynchronize(s));
  if (!ctx->comm) memcpy(cnt, (char *)ctx->h_ctrl + 64, sizeof(cnt));
  const int round = ctx->h_ctrl[kRcRound], verd = ctx->h_ctrl[kRcVerdict];
  if (p2p) ctx->p2p_tag += (unsigned)round;
  double rows_ms = 0, comm_ms = 0, vars_ms = 0;
  static const bool trace_rounds = getenv("MNTR_GPU_TRACE_ROUNDS") != nullptr;
  for (int r = 0; r < round && r < kTimed; ++r) {
    cudaEvent_t *ev = &ctx->round_ev[(size_t)(4 * r)];
    rows_ms += elapsed(ev[0], ev[1]); comm_ms += elapsed(ev[1], ev[2]); vars_ms += elapsed(ev[2], ev[3]);
    if (trace_rounds)
      fprintf(stderr, "[mntr rounds] rank %d round %d: rows %.4f ms, merge %.4f ms, vars %.4f ms, gap to next %.4f ms\n", ctx->rank,
              r + 1, elapsed(ev[0], ev[1]), elapsed(ev[1], ev[2]), elapsed(ev[2], ev[3]),
              r + 1 < round && r + 1 < kTimed ? elapsed(ev[3], ev[4]) : 0.0);
  }
  if (verdict) *verdict = verd;
  if (rounds) *rounds = round;
  if (nnz_updates) *nnz_updates = (int64_t)cnt[0];
  ctx->stats.nnz_updates += (int64_t)cnt[0];
  ctx->stats.rows_evaluated += (int64_t)cnt[1];
  ctx->stats.n_changes += ctx->h_ctrl[kRcPairs];
  ctx->stats.n_infeasible += verd != 0;
  ctx->stats.max_rounds = std::max(ctx->stats.max_rounds, round);
  ctx->stats.rows_ms += rows_ms; ctx->stats.comm_ms += comm_ms; ctx->stats.vars_ms += vars_ms;
  if (p2p) ctx->stats.sparse_rounds += round;
  return MNTR_OK;
}

static bool use_rounds(const mntr_gpu_ctx *ctx, const mntr_gpu_options &o)
{
  return ctx->comm != nullptr || ctx->force_rounds || (o.flags & MNTR_FLAG_PER_ROUND_KERNELS);
}

static double *mapped_host_ptr(const double *p);
static int tighten_single_zero_copy(mntr_gpu_ctx *ctx, double *lb_map, double *ub_map, const mntr_gpu_options &o,
                                    int32_t *verdict, int32_t *rounds, int64_t *nnz_updates);

static int tighten_single(mntr_gpu_ctx *ctx, double *lb, double *ub, const mntr_gpu_options &o,
                          int32_t *verdict, int32_t *rounds, int64_t *nnz_updates)
{
  const size_t bytes = sizeof(double) * (size_t)ctx->n;
  int rc;
  if (!ctx->no_zero_copy) {
    double *lb_map = mapped_host_ptr(lb), *ub_map = mapped_host_ptr(ub);
    if (lb_map && ub_map && !use_rounds(ctx, o)) return tighten_single_zero_copy(ctx, lb_map, ub_map, o, verdict, rounds, nnz_updates);
    if (lb_map && ub_map) {
      // per-round kernels on a page-locked, mapped box: the box comes in with the copy engine (55 GB/s; a kernel reading
      // mapped memory gets 42), the finish kernel writes only the bounds that moved straight back to the host box
      CU(cudaMemcpyAsync(ctx->d_lb, lb, bytes, cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaMemcpyAsync(ctx->d_ub, ub, bytes, cudaMemcpyHostToDevice, ctx->stream));
      if ((rc = run_rounds_dev(ctx, ctx->d_lb, ctx->d_ub, o, verdict, rounds, nnz_updates, lb_map, ub_map))) return rc;
      ctx->stats.kernel_ms += ctx->stats.rows_ms + ctx->stats.comm_ms + ctx->stats.vars_ms;
      return MNTR_OK;
    }
  }
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_lb, lb, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_ub, ub, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (use_rounds(ctx, o)) {
    if ((rc = run_rounds_dev(ctx, ctx->d_lb, ctx->d_ub, o, verdict, rounds, nnz_updates))) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaMemcpyAsync(lb, ctx->d_lb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ub, ctx->d_ub, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_ms += elapsed(ctx->ev[0], ctx->ev[1]);
    ctx->stats.kernel_ms += ctx->stats.rows_ms + ctx->stats.comm_ms + ctx->stats.vars_ms;
    ctx->stats.d2h_ms += elapsed(ctx->ev[2], ctx->ev[3]);
    return MNTR_OK;
  }
  if ((rc = run_single_dev(ctx, ctx->d_lb, ctx->d_ub, o))) return rc;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaMemcpyAsync(lb, ctx->d_lb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ub, ctx->d_ub, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const SingleCtrl ctrl = *ctx->h_single;      // written by the kernel's last block
  ctx->ctrl_clean = true;
  if (verdict) *verdict = ctrl.verdict();
  if (rounds) *rounds = ctrl.status[1];
  if (nnz_updates) *nnz_updates = (int64_t)ctrl.counters[0];
  account_single(ctx, ctrl);
  ctx->stats.h2d_ms += elapsed(ctx->ev[0], ctx->ev[1]);
  ctx->stats.kernel_ms += elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->stats.d2h_ms += elapsed(ctx->ev[2], ctx->ev[3]);
  return MNTR_OK;
}

// Device pointer of a PINNED (page-locked, mapped) host buffer, or nullptr for pageable memory.
static double *mapped_host_ptr(const double *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return nullptr;
  return static_cast<double *>(a.devicePointer);
}

// Single box whose host buffers are pinned: the fixpoint kernel reads the incoming bounds straight from host memory
// (one coalesced pass over PCIe in its first phase) and writes back ONLY the bounds that moved, so the call is one
// launch plus the 128-byte control block -- no staging copies of the whole box in either direction.
static int tighten_single_zero_copy(mntr_gpu_ctx *ctx, double *lb_map, double *ub_map, const mntr_gpu_options &o,
                                    int32_t *verdict, int32_t *rounds, int64_t *nnz_updates)
{
  int rc;
  // device-side timing of the call (stats.kernel_ms) costs two event records on the critical path of a ~125 us call:
  // only when asked for (MNTR_GPU_TIMING=1)
  static const bool timing = [] { const char *e = getenv("MNTR_GPU_TIMING"); return e && e[0] == '1'; }();
  if (timing) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if ((rc = run_single_dev(ctx, lb_map, ub_map, o))) return rc;
  if (timing) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const SingleCtrl ctrl = *ctx->h_single;      // written by the kernel's last block
  ctx->ctrl_clean = true;
  if (verdict) *verdict = ctrl.verdict();
  if (rounds) *rounds = ctrl.status[1];
  if (nnz_updates) *nnz_updates = (int64_t)ctrl.counters[0];
  account_single(ctx, ctrl);
  if (timing) ctx->stats.kernel_ms += elapsed(ctx->ev[1], ctx->ev[2]);       // includes the PCIe reads / writes of the box
  return MNTR_OK;
}

int mntr_gpu_tighten_single_dev(mntr_gpu_ctx *ctx, double *lb_dev, double *ub_dev, const mntr_gpu_options *opts,
                                int32_t *verdict, int32_t *rounds, int64_t *nnz_updates)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "tighten_single_dev: no problem loaded");
  if (!lb_dev || !ub_dev) return fail(ctx, MNTR_E_ARG, "tighten_single_dev: null box");
  CU(cudaSetDevice(ctx->device));
  mntr_gpu_options o = resolve_opts(opts, 1);
  o.order = MNTR_ORDER_JACOBI;
  int rc;
  ctx->stats = mntr_gpu_stats{};
  if (use_rounds(ctx, o)) {
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if ((rc = run_rounds_dev(ctx, lb_dev, ub_dev, o, verdict, rounds, nnz_updates))) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.kernel_ms = elapsed(ctx->ev[0], ctx->ev[1]);
    return MNTR_OK;
  }
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if ((rc = run_single_dev(ctx, lb_dev, ub_dev, o))) return rc;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const SingleCtrl ctrl = *ctx->h_single;      // written by the kernel's last block
  ctx->ctrl_clean = true;
  if (verdict) *verdict = ctrl.verdict();
  if (rounds) *rounds = ctrl.status[1];
  if (nnz_updates) *nnz_updates = (int64_t)ctrl.counters[0];
  account_single(ctx, ctrl);
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  if (ctx->sws.trace) {     // debug: phase durations of the fixpoint kernel
    static unsigned long long t[kTraceWords];
    CU(cudaMemcpy(t, ctx->sws.trace, sizeof(t), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[mntr trace] rounds=%d kernel=%.1fus phases(us):", ctrl.status[1], ctx->stats.kernel_ms * 1e3);
    const int np = 4 + 2 * (ctrl.status[1] + 1);
    for (int k = 1; k < np && k < 64 && t[k] != 0; ++k) fprintf(stderr, " %.1f", (double)(t[k] - t[k - 1]) * 1e-3);
    fprintf(stderr, "\n[mntr probe] warp 0 (us since kernel start):");
    for (int k = 32; k < 64 && t[k] != 0; ++k) fprintf(stderr, " %.1f", (double)(t[k] - t[0]) * 1e-3);
    fprintf(stderr, "\n[mntr arrive] per round: first / median / last block at the barrier (us since kernel start):");
    for (int r = 1; r < 16; ++r) {
      std::vector<double> a;
      for (int b = 0; b < 256; ++b) if (t[64 + b * 16 + r] != 0) a.push_back((double)(t[64 + b * 16 + r] - t[0]) * 1e-3);
      if (a.empty()) break;
      std::sort(a.begin(), a.end());
      fprintf(stderr, "  r%d %.1f/%.1f/%.1f", r, a.front(), a[a.size() / 2], a.back());
    }
    fprintf(stderr, "\n[mntr blocks] per round, phase k: median / max over blocks of the latest warp (us since kernel start)\n");
    for (int r = 1; r < 16; ++r) {
      bool any = false;
      for (int k = 0; k < 8; ++k) {
        std::vector<double> a;
        for (int b = 0; b < 256; ++b) {
          const unsigned long long v = t[64 + 256 * 16 + (b * 16 + r) * 8 + k];
          if (v != 0) a.push_back((double)(v - t[0]) * 1e-3);
        }
        if (a.empty()) continue;
        std::sort(a.begin(), a.end());
        if (!any) fprintf(stderr, "  r%d:", r);
        any = true;
        fprintf(stderr, "  k%d %.1f/%.1f(n=%d)", k, a[a.size() / 2], a.back(), (int)a.size());
      }
      if (any) fprintf(stderr, "\n");
    }
    CU(cudaMemset(ctx->sws.trace, 0, sizeof(t)));
  }
  return MNTR_OK;
}

void *mntr_gpu_stream(mntr_gpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int mntr_gpu_boxes_upload(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *lb, const double *ub, void *boxes_dev)
{
  if (!ctx || !ctx->lin_loaded || n_boxes <= 0 || !lb || !ub || !boxes_dev) return fail(ctx, MNTR_E_ARG, "boxes_upload: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (ctx->prepared_boxes == boxes_dev) ctx->prepared_boxes = nullptr;      // uploaded boxes: nothing is known about them
  int rc = ensure_stage(ctx);
  if (rc) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  const int32_t n = ctx->n;
  for (int64_t b0 = 0; b0 < n_boxes; b0 += ctx->stage_boxes) {
    const int32_t nb = (int32_t)std::min<int64_t>(ctx->stage_boxes, n_boxes - b0);
    CU(cudaMemcpyAsync(ctx->d_stage_lb, lb + b0 * n, sizeof(double) * (size_t)nb * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_stage_ub, ub + b0 * n, sizeof(double) * (size_t)nb * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_boxes_pack(ctx->d_stage_lb, ctx->d_stage_ub, n, (int32_t)b0, nb, (double2 *)boxes_dev, ld, ctx->stream));
  }
  CU(launch_boxes_pad((double2 *)boxes_dev, ld, n, n_boxes, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MNTR_OK;
}

int mntr_gpu_boxes_download(mntr_gpu_ctx *ctx, int32_t n_boxes, const void *boxes_dev, double *lb, double *ub)
{
  if (!ctx || !ctx->lin_loaded || n_boxes <= 0 || !lb || !ub || !boxes_dev) return fail(ctx, MNTR_E_ARG, "boxes_download: bad argument");
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx);
  if (rc) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  const int32_t n = ctx->n;
  for (int64_t b0 = 0; b0 < n_boxes; b0 += ctx->stage_boxes) {
    const int32_t nb = (int32_t)std::min<int64_t>(ctx->stage_boxes, n_boxes - b0);
    CU(launch_boxes_unpack((const double2 *)boxes_dev, ld, n, (int32_t)b0, nb, ctx->d_stage_lb, ctx->d_stage_ub, ctx->stream));
    CU(cudaMemcpyAsync(lb + b0 * n, ctx->d_stage_lb, sizeof(double) * (size_t)nb * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ub + b0 * n, ctx->d_stage_ub, sizeof(double) * (size_t)nb * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return MNTR_OK;
}

int mntr_gpu_tighten_dev(mntr_gpu_ctx *ctx, int32_t n_boxes, void *boxes_dev, const mntr_gpu_options *opts,
                         int32_t *verdict_dev, int32_t *rounds_dev, int64_t *nnz_dev)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "tighten_dev: no problem loaded");
  if (n_boxes <= 0 || !boxes_dev || !verdict_dev || !rounds_dev || !nnz_dev)
    return fail(ctx, MNTR_E_ARG, "tighten_dev: bad argument");
  CU(cudaSetDevice(ctx->device));
  const mntr_gpu_options o = resolve_opts(opts, n_boxes > 1 ? n_boxes : 2);
  if (o.order != MNTR_ORDER_REFERENCE)
    return fail(ctx, MNTR_E_UNSUPPORTED, "tighten_dev: device-resident boxes use MNTR_ORDER_REFERENCE");
  if (o.rounding != MNTR_ROUND_DIRECTED && o.rounding != MNTR_ROUND_NEAREST) return fail(ctx, MNTR_E_ARG, "tighten_dev: bad rounding");
  if (o.loop != MNTR_LOOP_FIXPOINT && o.loop != MNTR_LOOP_SIMPLEPRESOLVE) return fail(ctx, MNTR_E_ARG, "tighten_dev: bad loop mode");
  if (o.handlers < 0 || o.handlers > 2) return fail(ctx, MNTR_E_ARG, "tighten_dev: bad handlers");
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  const int64_t tiles = ld / 32;
  // row flags are sized for the context's batch capacity
  int rc = ensure_batch(ctx, n_boxes, false);
  if (rc) return rc;
  BatchIo io;
  io.boxes = (double2 *)boxes_dev; io.ld = ld; io.n_boxes = n_boxes; io.rowflag = ctx->d_rowflag; io.varflag = ctx->d_varflag; io.tstate = ctx->d_tstate;
  io.verdict = verdict_dev; io.rounds = rounds_dev; io.nnz = (long long *)nnz_dev; io.nl_evals = ctx->d_nl_evals;
  io.prepared = (ctx->prepared_boxes == boxes_dev && ctx->prepared_n == n_boxes) ? ctx->d_prepared : nullptr;
  ctx->prepared_boxes = nullptr;                   // consumed: the boxes are tightened in place
  (void)tiles;
  CU(cudaMemsetAsync(ctx->d_nl_evals, 0, sizeof(unsigned long long), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  CU(launch_batch_reference(ctx->lin, ctx->nl_loaded ? &ctx->nl : nullptr, io, o.rounding == MNTR_ROUND_DIRECTED,
                            o.loop, o.max_rounds, o.handlers != MNTR_HANDLERS_NONLINEAR,
                            (ctx->nl_loaded && o.handlers != MNTR_HANDLERS_LINEAR) ? 1 : 0, ctx->sm_count, ctx->stream));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  unsigned long long evals = 0;
  CU(cudaMemcpyAsync(&evals, ctx->d_nl_evals, sizeof(evals), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->stats.nl_evals = (int64_t)evals;
  return MNTR_OK;
}

int mntr_gpu_tighten(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub, const mntr_gpu_options *opts,
                     int32_t *verdict, int32_t *rounds, int64_t *nnz_updates)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "tighten: no problem loaded");
  if (n_boxes <= 0 || !lb || !ub) return fail(ctx, MNTR_E_ARG, "tighten: bad argument");
  CU(cudaSetDevice(ctx->device));
  const mntr_gpu_options o = resolve_opts(opts, n_boxes);
  if (o.rounding != MNTR_ROUND_DIRECTED && o.rounding != MNTR_ROUND_NEAREST) return fail(ctx, MNTR_E_ARG, "tighten: bad rounding");
  if (o.loop != MNTR_LOOP_FIXPOINT && o.loop != MNTR_LOOP_SIMPLEPRESOLVE) return fail(ctx, MNTR_E_ARG, "tighten: bad loop mode");
  ctx->stats = mntr_gpu_stats{};
  const int32_t n = ctx->n;

  if (o.handlers < 0 || o.handlers > 2) return fail(ctx, MNTR_E_ARG, "tighten: bad handlers");
  if (o.order == MNTR_ORDER_JACOBI) {
    if (ctx->nl_loaded && o.handlers != MNTR_HANDLERS_LINEAR)
      return fail(ctx, MNTR_E_UNSUPPORTED, "tighten: CGraph tapes need MNTR_ORDER_REFERENCE");
    for (int32_t b = 0; b < n_boxes; ++b) {
      int rc = tighten_single(ctx, lb + (int64_t)b * n, ub + (int64_t)b * n, o, verdict ? verdict + b : nullptr,
                              rounds ? rounds + b : nullptr, nnz_updates ? nnz_updates + b : nullptr);
      if (rc) return rc;
    }
    return MNTR_OK;
  }
  if (o.order != MNTR_ORDER_REFERENCE) return fail(ctx, MNTR_E_ARG, "tighten: bad order");

  int rc = ensure_batch(ctx, n_boxes, true);
  if (rc) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if ((rc = mntr_gpu_boxes_upload(ctx, n_boxes, lb, ub, ctx->d_boxes))) return rc;
  BatchIo io;
  io.boxes = ctx->d_boxes; io.ld = ld; io.n_boxes = n_boxes; io.rowflag = ctx->d_rowflag; io.varflag = ctx->d_varflag; io.tstate = ctx->d_tstate;
  io.verdict = ctx->d_verdict; io.rounds = ctx->d_rounds; io.nnz = ctx->d_nnzb; io.nl_evals = ctx->d_nl_evals;
  ctx->prepared_boxes = nullptr;                   // (uploaded boxes: nothing is known about them)
  CU(cudaMemsetAsync(ctx->d_nl_evals, 0, sizeof(unsigned long long), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  CU(launch_batch_reference(ctx->lin, ctx->nl_loaded ? &ctx->nl : nullptr, io, o.rounding == MNTR_ROUND_DIRECTED,
                            o.loop, o.max_rounds, o.handlers != MNTR_HANDLERS_NONLINEAR,
                            (ctx->nl_loaded && o.handlers != MNTR_HANDLERS_LINEAR) ? 1 : 0, ctx->sm_count, ctx->stream));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if ((rc = mntr_gpu_boxes_download(ctx, n_boxes, ctx->d_boxes, lb, ub))) return rc;
  std::vector<int32_t> hv(n_boxes), hr(n_boxes);
  std::vector<long long> hn(n_boxes);
  CU(cudaMemcpyAsync(hv.data(), ctx->d_verdict, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hr.data(), ctx->d_rounds, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hn.data(), ctx->d_nnzb, sizeof(long long) * (size_t)n_boxes, cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long evals = 0;
  CU(cudaMemcpyAsync(&evals, ctx->d_nl_evals, sizeof(evals), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stats.nl_evals = (int64_t)evals;
  for (int32_t b = 0; b < n_boxes; ++b) {
    if (verdict) verdict[b] = hv[b];
    if (rounds) rounds[b] = hr[b];
    if (nnz_updates) nnz_updates[b] = hn[b];
    ctx->stats.nnz_updates += hn[b];
    ctx->stats.n_infeasible += hv[b] != 0;
    ctx->stats.max_rounds = std::max(ctx->stats.max_rounds, hr[b]);
  }
  ctx->stats.h2d_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->stats.d2h_ms = elapsed(ctx->ev[2], ctx->ev[3]);
  return MNTR_OK;
}

// ---- node batches given as branching deltas on a common root box ----
namespace {

// device copies of a root box and the boxes' delta lists
struct DeviceDeltas {
  double *rl = nullptr, *ru = nullptr, *val = nullptr;
  long long *ptr = nullptr;
  int32_t *var = nullptr;
  uint8_t *up = nullptr;
};

int check_deltas(mntr_gpu_ctx *ctx, const char *who, int32_t n_boxes, const double *root_lb, const double *root_ub,
                 const int64_t *delta_ptr, const int32_t *delta_var, const uint8_t *delta_is_upper, const double *delta_val)
{
  if (n_boxes <= 0 || !root_lb || !root_ub || !delta_ptr) return fail(ctx, MNTR_E_ARG, "%s: bad argument", who);
  const int64_t n_delta = delta_ptr[n_boxes];
  if (delta_ptr[0] != 0 || n_delta < 0 || (n_delta > 0 && (!delta_var || !delta_is_upper || !delta_val)))
    return fail(ctx, MNTR_E_ARG, "%s: bad delta lists", who);
  for (int32_t b = 0; b < n_boxes; ++b)
    if (delta_ptr[b + 1] < delta_ptr[b]) return fail(ctx, MNTR_E_ARG, "%s: delta_ptr not monotone at box %d", who, b);
  for (int64_t q = 0; q < n_delta; ++q)
    if (delta_var[q] < 0 || delta_var[q] >= ctx->n) return fail(ctx, MNTR_E_ARG, "%s: delta variable %d out of range", who, delta_var[q]);
  return MNTR_OK;
}

// uploads root + deltas (the context's grow-only scratch) and builds the boxes in the engine's layout
int boxes_from_deltas(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *root_lb,
                      const double *root_ub, const int64_t *delta_ptr, const int32_t *delta_var,
                      const uint8_t *delta_is_upper, const double *delta_val, double2 *boxes, DeviceDeltas &D)
{
  const int32_t n = ctx->n;
  const int64_t n_delta = delta_ptr[n_boxes], ld = mntr_gpu_box_ld(n_boxes);
  int rc;
  if ((rc = scratch(ctx, 0, sizeof(double) * (size_t)n, (void **)&D.rl))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(double) * (size_t)n, (void **)&D.ru))) return rc;
  if ((rc = scratch(ctx, 2, sizeof(long long) * ((size_t)n_boxes + 1), (void **)&D.ptr))) return rc;
  if ((rc = scratch(ctx, 3, sizeof(int32_t) * (size_t)n_delta, (void **)&D.var))) return rc;
  if ((rc = scratch(ctx, 4, (size_t)n_delta, (void **)&D.up))) return rc;
  if ((rc = scratch(ctx, 5, sizeof(double) * (size_t)n_delta, (void **)&D.val))) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "delta_ptr / mod_ptr are copied as long long");
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(D.rl, root_lb, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(D.ru, root_ub, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(D.ptr, delta_ptr, sizeof(long long) * ((size_t)n_boxes + 1), cudaMemcpyHostToDevice, s));
  if (n_delta > 0) {
    CU(cudaMemcpyAsync(D.var, delta_var, sizeof(int32_t) * (size_t)n_delta, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(D.up, delta_is_upper, (size_t)n_delta, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(D.val, delta_val, sizeof(double) * (size_t)n_delta, cudaMemcpyHostToDevice, s));
  }
  // the batch is PREPARED (BatchIo::prepared): the root variables that need rounding or the bound check and the deltas'
  // variables are flagged on the device -- the next batch call on these boxes skips the all-variables pass of its first sweep
  ctx->prepared_boxes = nullptr;
  const bool prep = ctx->d_varflag != nullptr && ld <= ctx->batch_ld && !getenv("MNTR_GPU_NO_PREPARED");
  if (prep) {
    CU(cudaMemsetAsync(ctx->d_varflag, 0, sizeof(uint32_t) * (size_t)n * (size_t)(ld / 32), s));
    CU(cudaMemsetAsync(ctx->d_prepared, 0, sizeof(int32_t), s));
  }
  CU(launch_boxes_from_root(D.rl, D.ru, n, n_boxes, boxes, ld, ctx->lin.var_type, prep ? ctx->d_varflag : nullptr, s));
  CU(launch_apply_deltas(D.ptr, D.var, D.up, D.val, n_boxes, boxes, ld, prep ? ctx->d_varflag : nullptr, n, s));
  if (prep) { ctx->prepared_boxes = boxes; ctx->prepared_n = n_boxes; }
  return MNTR_OK;
}

}  // namespace

int mntr_gpu_boxes_from_deltas(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *root_lb, const double *root_ub,
                               const int64_t *delta_ptr, const int32_t *delta_var, const uint8_t *delta_is_upper,
                               const double *delta_val, void *boxes_dev)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "boxes_from_deltas: no problem loaded");
  if (!boxes_dev) return fail(ctx, MNTR_E_ARG, "boxes_from_deltas: null boxes");
  int rc = check_deltas(ctx, "boxes_from_deltas", n_boxes, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_batch(ctx, n_boxes, false))) return rc;         // (the flag words of the prepared batch)
  DeviceDeltas D;
  rc = boxes_from_deltas(ctx, n_boxes, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val,
                         (double2 *)boxes_dev, D);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, MNTR_E_CUDA, "boxes_from_deltas: %s", cudaGetErrorString(e));
  return MNTR_OK;
}

// ---- root presolve row operations ----
static int upload_perm(mntr_gpu_ctx *ctx, int32_t **d_perm)
{
  int rc = scratch(ctx, 14, sizeof(int32_t) * (size_t)std::max(ctx->m, 1), (void **)d_perm);
  if (rc) return rc;
  if (ctx->m > 0) CU(cudaMemcpyAsync(*d_perm, ctx->h_perm.data(), sizeof(int32_t) * (size_t)ctx->m, cudaMemcpyHostToDevice, ctx->stream));
  return MNTR_OK;
}

int mntr_gpu_root_dup_rows(mntr_gpu_ctx *ctx, const double *r1, const double *r2, double *h1_out, double *h2_out,
                           int64_t cap, int32_t *pair_i, int32_t *pair_j, uint8_t *pair_kind, int64_t *n_pairs_out)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "root_dup_rows: no problem loaded");
  if (!r1 || !r2 || cap < 0 || !n_pairs_out || (cap > 0 && (!pair_i || !pair_j || !pair_kind)))
    return fail(ctx, MNTR_E_ARG, "root_dup_rows: bad argument");
  CU(cudaSetDevice(ctx->device));
  const int32_t m = ctx->m, n = ctx->n;
  *n_pairs_out = 0;
  if (m <= 0) return MNTR_OK;
  int rc;
  int32_t *d_perm = nullptr, *d_pi = nullptr, *d_pj = nullptr; uint8_t *d_pk = nullptr;
  double *d_r1 = nullptr, *d_r2 = nullptr, *d_h1 = nullptr, *d_h2 = nullptr;
  unsigned long long *d_cnt = nullptr;
  if ((rc = upload_perm(ctx, &d_perm))) return rc;
  if ((rc = scratch(ctx, 0, sizeof(double) * (size_t)n, (void **)&d_r1))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(double) * (size_t)n, (void **)&d_r2))) return rc;
  if ((rc = scratch(ctx, 5, sizeof(double) * (size_t)m, (void **)&d_h1))) return rc;
  if ((rc = scratch(ctx, 13, sizeof(double) * (size_t)m, (void **)&d_h2))) return rc;
  if ((rc = scratch(ctx, 3, sizeof(int32_t) * (size_t)cap, (void **)&d_pi))) return rc;
  if ((rc = scratch(ctx, 11, sizeof(int32_t) * (size_t)cap, (void **)&d_pj))) return rc;
  if ((rc = scratch(ctx, 12, (size_t)cap, (void **)&d_pk))) return rc;
  if ((rc = scratch(ctx, 6, sizeof(unsigned long long), (void **)&d_cnt))) return rc;
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_r1, r1, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_r2, r2, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
  CU(cudaEventRecord(ctx->ev[1], s));
  CU(launch_row_hash(ctx->lin, d_perm, d_r1, d_r2, d_h1, d_h2, s));
  CU(launch_dup_pairs(m, d_h1, d_h2, (long long)cap, d_pi, d_pj, d_pk, d_cnt, s));
  CU(cudaEventRecord(ctx->ev[2], s));
  unsigned long long cnt = 0;
  CU(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  if (h1_out) CU(cudaMemcpyAsync(h1_out, d_h1, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
  if (h2_out) CU(cudaMemcpyAsync(h2_out, d_h2, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  *n_pairs_out = (int64_t)cnt;
  const size_t k = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)cap);
  if (k > 0) {
    std::vector<int32_t> pi(k), pj(k); std::vector<uint8_t> pk(k);
    CU(cudaMemcpy(pi.data(), d_pi, sizeof(int32_t) * k, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pj.data(), d_pj, sizeof(int32_t) * k, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pk.data(), d_pk, k, cudaMemcpyDeviceToHost));
    std::vector<size_t> idx(k);
    for (size_t q = 0; q < k; ++q) idx[q] = q;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return pi[a] != pi[b] ? pi[a] < pi[b] : pj[a] < pj[b]; });
    for (size_t q = 0; q < k; ++q) { pair_i[q] = pi[idx[q]]; pair_j[q] = pj[idx[q]]; pair_kind[q] = pk[idx[q]]; }
  }
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  return MNTR_OK;
}

int mntr_gpu_root_redundant_rows(mntr_gpu_ctx *ctx, const double *lb, const double *ub, uint8_t *redundant, int64_t *n_redundant)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "root_redundant_rows: no problem loaded");
  if (!lb || !ub || (ctx->m > 0 && !redundant)) return fail(ctx, MNTR_E_ARG, "root_redundant_rows: bad argument");
  CU(cudaSetDevice(ctx->device));
  const int32_t m = ctx->m, n = ctx->n;
  if (n_redundant) *n_redundant = 0;
  if (m <= 0) return MNTR_OK;
  int rc;
  int32_t *d_perm = nullptr; double *d_lb = nullptr, *d_ub = nullptr; uint8_t *d_flag = nullptr; unsigned long long *d_cnt = nullptr;
  if ((rc = upload_perm(ctx, &d_perm))) return rc;
  if ((rc = scratch(ctx, 0, sizeof(double) * (size_t)n, (void **)&d_lb))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(double) * (size_t)n, (void **)&d_ub))) return rc;
  if ((rc = scratch(ctx, 12, (size_t)m, (void **)&d_flag))) return rc;
  if ((rc = scratch(ctx, 6, sizeof(unsigned long long), (void **)&d_cnt))) return rc;
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(d_lb, lb, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_ub, ub, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), s));
  CU(cudaEventRecord(ctx->ev[1], s));
  CU(launch_redundant_rows(ctx->lin, d_perm, d_lb, d_ub, d_flag, d_cnt, s));
  CU(cudaEventRecord(ctx->ev[2], s));
  unsigned long long cnt = 0;
  CU(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(redundant, d_flag, (size_t)m, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (n_redundant) *n_redundant = (int64_t)cnt;
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  return MNTR_OK;
}

// LinearHandler::coeffImp_ on the device.  The pass is sequential in row index only through the 2-term rows that the
// implications (computeImpBounds_) read: a row sees the improved version of those with a smaller index.  Rows are
// therefore scheduled in dependency levels (level = 1 + the highest level of an earlier 2-term row the row's
// implications can read), one launch per level, one thread per row; a level reads the coefficients and row bounds as
// the lower levels left them.
int mntr_gpu_root_coeff_imp(mntr_gpu_ctx *ctx, int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col,
                            const double *val, const double *row_lb, const double *row_ub, const uint8_t *var_type,
                            const double *lb, const double *ub, int64_t cap, int32_t *out_row, int32_t *out_var,
                            double *out_coef, int32_t *out_side, double *out_bnd, double *out_delta, int64_t *n_out,
                            int32_t *n_levels_out, int32_t *n_erased_out)
{
  if (!ctx) return MNTR_E_ARG;
  if (m < 0 || n < 0 || !row_ptr || !n_out || cap < 0) return fail(ctx, MNTR_E_ARG, "root_coeff_imp: bad argument");
  if (cap > 0 && (!out_row || !out_var || !out_coef || !out_side || !out_bnd)) return fail(ctx, MNTR_E_ARG, "root_coeff_imp: null output");
  *n_out = 0;
  if (n_levels_out) *n_levels_out = 0;
  if (n_erased_out) *n_erased_out = 0;
  if (m == 0) return MNTR_OK;
  if (!col || !val || !row_lb || !row_ub || !var_type || !lb || !ub) return fail(ctx, MNTR_E_ARG, "root_coeff_imp: null input");
  CU(cudaSetDevice(ctx->device));
  const int64_t nnz = row_ptr[m];
  auto is_term = [](double a) { return std::fabs(a) > 1e-9; };
  auto is_bin = [&](int32_t j) { return (var_type[j] == 0 || var_type[j] == 2) && ub[j] > lb[j] + 0.5; };
  // ---- host: CSC, 2-term rows, candidate rows and their dependency levels ----
  std::vector<int32_t> cptr((size_t)n + 2, 0), crow((size_t)std::max<int64_t>(nnz, 1)), nterms((size_t)m, 0);
  for (int32_t i = 0; i < m; ++i)
    for (int32_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) {
      if (col[t] < 0 || col[t] >= n) return fail(ctx, MNTR_E_ARG, "root_coeff_imp: column out of range in row %d", i);
      cptr[(size_t)col[t] + 2]++;
      nterms[(size_t)i] += is_term(val[t]);
    }
  for (int32_t j = 0; j < n; ++j) cptr[(size_t)j + 2] += cptr[(size_t)j + 1];
  for (int32_t i = 0; i < m; ++i)
    for (int32_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) crow[(size_t)cptr[(size_t)col[t] + 1]++] = i;
  std::vector<uint8_t> is2((size_t)m, 0);
  for (int32_t i = 0; i < m; ++i) is2[(size_t)i] = nterms[(size_t)i] == 2;
  std::vector<int32_t> level((size_t)m, -1);
  int32_t n_levels = 0;
  for (int32_t c = 0; c < m; ++c) {
    if (!(row_lb[c] <= -INFINITY || row_ub[c] >= INFINITY) || nterms[(size_t)c] < 2) continue;     // :620-627
    bool has_bin = false;
    int32_t lev = 0;
    for (int32_t t = row_ptr[c]; t < row_ptr[c + 1]; ++t) {
      const int32_t z = col[t];
      if (!is_term(val[t]) || !is_bin(z)) continue;
      has_bin = true;
      if (nterms[(size_t)c] >= 50) continue;          // no implications (:625-629)
      // 2-term rows with z and another variable of c, earlier in the pass
      for (int32_t q = cptr[(size_t)z]; q < cptr[(size_t)z + 1]; ++q) {
        const int32_t c2 = crow[(size_t)q];
        if (c2 >= c) break;                           // (rows of a variable ascend)
        if (!is2[(size_t)c2] || level[(size_t)c2] < 0) continue;      // never improved: reads as it came
        int32_t other = -1;
        for (int32_t t2 = row_ptr[c2]; t2 < row_ptr[c2 + 1]; ++t2) if (is_term(val[t2]) && col[t2] != z) other = col[t2];
        if (other < 0) continue;
        const int32_t *b = col + row_ptr[c], *e = col + row_ptr[c + 1];
        if (std::find(b, e, other) != e) lev = std::max(lev, level[(size_t)c2] + 1);
      }
    }
    if (!has_bin) continue;
    level[(size_t)c] = lev;
    n_levels = std::max(n_levels, lev + 1);
  }
  std::vector<int32_t> lptr((size_t)n_levels + 1, 0), lrows;
  for (int32_t c = 0; c < m; ++c) if (level[(size_t)c] >= 0) lptr[(size_t)level[(size_t)c] + 1]++;
  for (int32_t l = 0; l < n_levels; ++l) lptr[(size_t)l + 1] += lptr[(size_t)l];
  lrows.resize((size_t)std::max(lptr[(size_t)n_levels], 1));
  {
    std::vector<int32_t> pos(lptr.begin(), lptr.end());
    for (int32_t c = 0; c < m; ++c) if (level[(size_t)c] >= 0) lrows[(size_t)pos[(size_t)level[(size_t)c]]++] = c;
  }
  if (n_levels_out) *n_levels_out = n_levels;
  if (n_levels == 0) return MNTR_OK;
  // ---- device ----
  std::vector<void *> owned;
  struct Free { std::vector<void *> &v; ~Free() { for (void *p : v) cudaFree(p); } } guard{owned};
  CoeffProb Q{};
  Q.m = m; Q.n = n;
  int rc;
  if ((rc = dev_upload(ctx, owned, row_ptr, (size_t)m + 1, &Q.row_ptr))) return rc;
  if ((rc = dev_upload(ctx, owned, col, (size_t)nnz, &Q.col))) return rc;
  if ((rc = dev_upload(ctx, owned, val, (size_t)nnz, &Q.val0))) return rc;
  if ((rc = dev_upload(ctx, owned, row_lb, (size_t)m, &Q.rlb0))) return rc;
  if ((rc = dev_upload(ctx, owned, row_ub, (size_t)m, &Q.rub0))) return rc;
  { const double *p; if ((rc = dev_upload(ctx, owned, val, (size_t)nnz, &p))) return rc; Q.val = const_cast<double *>(p); }
  { const double *p; if ((rc = dev_upload(ctx, owned, row_lb, (size_t)m, &p))) return rc; Q.rlb = const_cast<double *>(p); }
  { const double *p; if ((rc = dev_upload(ctx, owned, row_ub, (size_t)m, &p))) return rc; Q.rub = const_cast<double *>(p); }
  if ((rc = dev_upload(ctx, owned, var_type, (size_t)n, &Q.var_type))) return rc;
  if ((rc = dev_upload(ctx, owned, lb, (size_t)n, &Q.lb))) return rc;
  if ((rc = dev_upload(ctx, owned, ub, (size_t)n, &Q.ub))) return rc;
  if ((rc = dev_upload(ctx, owned, cptr.data(), (size_t)n + 1, &Q.cptr))) return rc;
  if ((rc = dev_upload(ctx, owned, crow.data(), (size_t)nnz, &Q.crow))) return rc;
  if ((rc = dev_upload(ctx, owned, is2.data(), (size_t)m, &Q.is2))) return rc;
  const int32_t *d_rows;
  if ((rc = dev_upload(ctx, owned, lrows.data(), lrows.size(), &d_rows))) return rc;
  const size_t capz = (size_t)std::max<int64_t>(cap, 1);
  int32_t *d_orow, *d_ovar, *d_oside, *d_erased; double *d_ocoef, *d_obnd, *d_odelta; unsigned long long *d_cnt;
  auto dalloc = [&](void **p, size_t bytes) -> int { CU(cudaMalloc(p, std::max<size_t>(bytes, 16))); owned.push_back(*p); return MNTR_OK; };
  if ((rc = dalloc((void **)&d_orow, 4 * capz)) || (rc = dalloc((void **)&d_ovar, 4 * capz)) || (rc = dalloc((void **)&d_oside, 4 * capz)) ||
      (rc = dalloc((void **)&d_ocoef, 8 * capz)) || (rc = dalloc((void **)&d_obnd, 8 * capz)) || (rc = dalloc((void **)&d_odelta, 8 * capz)) ||
      (rc = dalloc((void **)&d_cnt, 16)) ||
      (rc = dalloc((void **)&d_erased, 16))) return rc;
  cudaStream_t s = ctx->stream;
  CU(cudaMemsetAsync(d_cnt, 0, 16, s));
  CU(cudaMemsetAsync(d_erased, 0, 16, s));
  CU(cudaEventRecord(ctx->ev[1], s));
  for (int32_t l = 0; l < n_levels; ++l)
    CU(launch_coeff_imp(Q, d_rows + lptr[(size_t)l], lptr[(size_t)l + 1] - lptr[(size_t)l], (long long)cap, d_orow, d_ovar, d_ocoef,
                        d_oside, d_obnd, d_odelta, d_cnt, d_erased, s));
  CU(cudaEventRecord(ctx->ev[2], s));
  unsigned long long cnt = 0; int32_t erased = 0;
  CU(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&erased, d_erased, sizeof(erased), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  const size_t k = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)cap);
  if (k) {
    std::vector<int32_t> r(k), v(k), sd(k); std::vector<double> cf(k), bd(k), dl(k);
    CU(cudaMemcpyAsync(dl.data(), d_odelta, 8 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(r.data(), d_orow, 4 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(v.data(), d_ovar, 4 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(sd.data(), d_oside, 4 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(cf.data(), d_ocoef, 8 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(bd.data(), d_obnd, 8 * k, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    std::vector<size_t> ord(k);
    for (size_t i = 0; i < k; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return r[a] < r[b]; });      // the reference's order: by row
    for (size_t i = 0; i < k; ++i) {
      out_row[i] = r[ord[i]]; out_var[i] = v[ord[i]]; out_coef[i] = cf[ord[i]]; out_side[i] = sd[ord[i]]; out_bnd[i] = bd[ord[i]];
      if (out_delta) out_delta[i] = dl[ord[i]];
    }
  }
  *n_out = (int64_t)cnt;
  if (n_erased_out) *n_erased_out = erased;
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  return MNTR_OK;
}

// ---- QuadHandler::simplePresolve (QuadHandler.cpp:1146-1201) ------------------------------------------------------
int mntr_gpu_load_quad_relations(mntr_gpu_ctx *ctx, int32_t n_sq, const int32_t *sq_x, const int32_t *sq_y, int32_t n_bil,
                                 const int32_t *b_x0, const int32_t *b_x1, const int32_t *b_y)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "load_quad_relations: call load_linear first (m may be 0)");
  CU(cudaSetDevice(ctx->device));
  free_all(ctx->qrel_allocs);
  ctx->qrel_loaded = false;
  ctx->qrel = QRelDev{};
  if (n_sq < 0 || n_bil < 0 || (n_sq > 0 && (!sq_x || !sq_y)) || (n_bil > 0 && (!b_x0 || !b_x1 || !b_y)))
    return fail(ctx, MNTR_E_ARG, "load_quad_relations: null or negative argument");
  const int32_t n = ctx->n, n_rel = n_sq + n_bil;
  if (n_rel == 0) return MNTR_OK;
  // the handler's container order: squares ascending in x, one per x (a map); products ascending in (x0, x1), x0 < x1
  for (int32_t k = 0; k < n_sq; ++k) {
    if (sq_x[k] < 0 || sq_x[k] >= n || sq_y[k] < 0 || sq_y[k] >= n) return fail(ctx, MNTR_E_ARG, "load_quad_relations: variable out of range in square %d", k);
    if (k > 0 && sq_x[k] <= sq_x[k - 1]) return fail(ctx, MNTR_E_ARG, "load_quad_relations: squares not strictly ascending in x at %d", k);
  }
  for (int32_t k = 0; k < n_bil; ++k) {
    if (b_x0[k] < 0 || b_x1[k] >= n || b_x0[k] >= b_x1[k] || b_y[k] < 0 || b_y[k] >= n)
      return fail(ctx, MNTR_E_ARG, "load_quad_relations: bad product %d (need 0 <= x0 < x1 < n)", k);
    if (k > 0 && (b_x0[k] < b_x0[k - 1] || (b_x0[k] == b_x0[k - 1] && b_x1[k] <= b_x1[k - 1])))
      return fail(ctx, MNTR_E_ARG, "load_quad_relations: products not strictly ascending in (x0, x1) at %d", k);
  }
  // wavefront levels of the sequential sweep: a relation reads and may write all of its variables
  std::vector<int32_t> last((size_t)std::max(n, 1), -1), level((size_t)n_rel);
  int32_t n_levels = 0;
  auto vars_of = [&](int32_t r, int32_t v[3]) {
    if (r < n_sq) { v[0] = sq_x[r]; v[1] = sq_y[r]; v[2] = sq_y[r]; }
    else { v[0] = b_x0[r - n_sq]; v[1] = b_x1[r - n_sq]; v[2] = b_y[r - n_sq]; }
  };
  for (int32_t r = 0; r < n_rel; ++r) {
    int32_t v[3]; vars_of(r, v);
    int32_t lev = 0;
    for (int t = 0; t < 3; ++t) lev = std::max(lev, last[(size_t)v[t]] + 1);
    level[(size_t)r] = lev;
    for (int t = 0; t < 3; ++t) last[(size_t)v[t]] = lev;
    n_levels = std::max(n_levels, lev + 1);
  }
  std::vector<int32_t> lptr((size_t)n_levels + 1, 0), a((size_t)n_rel), b((size_t)n_rel), y((size_t)n_rel);
  for (int32_t r = 0; r < n_rel; ++r) lptr[(size_t)level[(size_t)r] + 1]++;
  for (int32_t l = 0; l < n_levels; ++l) lptr[(size_t)l + 1] += lptr[(size_t)l];
  {
    std::vector<int32_t> pos(lptr.begin(), lptr.end());
    for (int32_t r = 0; r < n_rel; ++r) {
      const int32_t q = pos[(size_t)level[(size_t)r]]++;
      if (r < n_sq) { a[(size_t)q] = sq_x[r]; b[(size_t)q] = -1; y[(size_t)q] = sq_y[r]; }
      else { a[(size_t)q] = b_x0[r - n_sq]; b[(size_t)q] = b_x1[r - n_sq]; y[(size_t)q] = b_y[r - n_sq]; }
    }
  }
  QRelDev &Q = ctx->qrel;
  int rc;
  if ((rc = dev_upload(ctx, ctx->qrel_allocs, a.data(), a.size(), &Q.a))) return rc;
  if ((rc = dev_upload(ctx, ctx->qrel_allocs, b.data(), b.size(), &Q.b))) return rc;
  if ((rc = dev_upload(ctx, ctx->qrel_allocs, y.data(), y.size(), &Q.y))) return rc;
  if ((rc = dev_upload(ctx, ctx->qrel_allocs, lptr.data(), lptr.size(), &Q.level_ptr))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  Q.n_rel = n_rel; Q.n_levels = n_levels; Q.var_type = ctx->lin.var_type;
  ctx->qrel_loaded = true;
  return MNTR_OK;
}

int mntr_gpu_quad_simple_presolve(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub, int32_t rounding,
                                  int32_t *n_mods, int32_t *n_inconsistent)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "quad_simple_presolve: no problem loaded");
  if (n_boxes <= 0 || !lb || !ub) return fail(ctx, MNTR_E_ARG, "quad_simple_presolve: bad argument");
  if (rounding != MNTR_ROUND_DIRECTED && rounding != MNTR_ROUND_NEAREST) return fail(ctx, MNTR_E_ARG, "quad_simple_presolve: bad rounding");
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_batch(ctx, n_boxes, true))) return rc;
  if ((rc = mntr_gpu_boxes_upload(ctx, n_boxes, lb, ub, ctx->d_boxes))) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  cudaStream_t s = ctx->stream;
  CU(cudaMemsetAsync(ctx->d_verdict, 0, sizeof(int32_t) * (size_t)ld, s));
  CU(cudaMemsetAsync(ctx->d_rounds, 0, sizeof(int32_t) * (size_t)ld, s));
  CU(cudaEventRecord(ctx->ev[1], s));
  if (ctx->qrel_loaded) {
    QRelDev Q = ctx->qrel;
    Q.var_type = ctx->lin.var_type;
    CU(launch_quad_relations(Q, ctx->d_boxes, ld, n_boxes, rounding == MNTR_ROUND_DIRECTED, ctx->d_rounds, ctx->d_verdict, s));
  }
  CU(cudaEventRecord(ctx->ev[2], s));
  if (n_mods) CU(cudaMemcpyAsync(n_mods, ctx->d_rounds, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  if (n_inconsistent) CU(cudaMemcpyAsync(n_inconsistent, ctx->d_verdict, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if ((rc = mntr_gpu_boxes_download(ctx, n_boxes, ctx->d_boxes, lb, ub))) return rc;
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  return MNTR_OK;
}

int mntr_gpu_quad_presolve_node(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub, int32_t rounding, int32_t max_sweeps,
                                int32_t *verdict, int32_t *n_mods, int32_t *n_sweeps)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "quad_presolve_node: no problem loaded");
  if (n_boxes <= 0 || !lb || !ub || !verdict) return fail(ctx, MNTR_E_ARG, "quad_presolve_node: bad argument");
  if (rounding != MNTR_ROUND_DIRECTED && rounding != MNTR_ROUND_NEAREST) return fail(ctx, MNTR_E_ARG, "quad_presolve_node: bad rounding");
  CU(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_batch(ctx, n_boxes, true))) return rc;
  if ((rc = mntr_gpu_boxes_upload(ctx, n_boxes, lb, ub, ctx->d_boxes))) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);
  cudaStream_t s = ctx->stream;
  int32_t *d_sweeps = (int32_t *)ctx->d_nnzb;                  // [ld] long long: room for ld int32
  CU(cudaMemsetAsync(ctx->d_verdict, 0, sizeof(int32_t) * (size_t)ld, s));
  CU(cudaMemsetAsync(ctx->d_rounds, 0, sizeof(int32_t) * (size_t)ld, s));
  CU(cudaMemsetAsync(d_sweeps, 0, sizeof(int32_t) * (size_t)ld, s));
  CU(cudaEventRecord(ctx->ev[1], s));
  if (ctx->qrel_loaded) {
    QRelDev Q = ctx->qrel;
    Q.var_type = ctx->lin.var_type;
    CU(launch_quad_node(Q, ctx->d_boxes, ld, n_boxes, rounding == MNTR_ROUND_DIRECTED, max_sweeps, ctx->d_verdict, ctx->d_rounds,
                        d_sweeps, s));
  }
  CU(cudaEventRecord(ctx->ev[2], s));
  CU(cudaMemcpyAsync(verdict, ctx->d_verdict, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  if (n_mods) CU(cudaMemcpyAsync(n_mods, ctx->d_rounds, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  if (n_sweeps) CU(cudaMemcpyAsync(n_sweeps, d_sweeps, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (!ctx->qrel_loaded && n_sweeps)                       // no relations: the reference's loop still runs once (:1215-1217)
    for (int32_t b = 0; b < n_boxes; ++b) n_sweeps[b] = 1;
  if ((rc = mntr_gpu_boxes_download(ctx, n_boxes, ctx->d_boxes, lb, ub))) return rc;
  ctx->stats = mntr_gpu_stats{};
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  return MNTR_OK;
}

void *mntr_gpu_alloc_host(mntr_gpu_ctx *ctx, int64_t bytes)
{
  if (!ctx || bytes <= 0) return nullptr;
  if (cudaSetDevice(ctx->device) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  void *p = nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    (void)cudaGetLastError();
    fail(ctx, MNTR_E_NOMEM, "alloc_host: cudaHostAlloc of %lld bytes failed", (long long)bytes);
    return nullptr;
  }
  return p;
}

void mntr_gpu_free_host(mntr_gpu_ctx *ctx, void *p)
{
  if (!p) return;
  if (ctx) cudaSetDevice(ctx->device);
  if (cudaFreeHost(p) != cudaSuccess) (void)cudaGetLastError();
}

// Node batch given as branching deltas on a common root box; results as the VarBoundMod tuples to emit.
int mntr_gpu_tighten_nodes(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *root_lb, const double *root_ub,
                           const int64_t *delta_ptr, const int32_t *delta_var, const uint8_t *delta_is_upper,
                           const double *delta_val, const mntr_gpu_options *opts, int32_t *verdict, int32_t *rounds,
                           int64_t *mod_ptr, int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val, int64_t mod_cap,
                           int64_t *n_mods_out)
{
  if (!ctx) return MNTR_E_ARG;
  if (!ctx->lin_loaded) return fail(ctx, MNTR_E_STATE, "tighten_nodes: no problem loaded");
  if (!mod_ptr || mod_cap < 0 || (mod_cap > 0 && (!mod_var || !mod_is_upper || !mod_val)))
    return fail(ctx, MNTR_E_ARG, "tighten_nodes: bad argument");
  int rc = check_deltas(ctx, "tighten_nodes", n_boxes, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  mntr_gpu_options o = resolve_opts(opts, n_boxes);
  o.order = MNTR_ORDER_REFERENCE;      // a node batch always runs the reference-order kernel
  if (o.rounding != MNTR_ROUND_DIRECTED && o.rounding != MNTR_ROUND_NEAREST) return fail(ctx, MNTR_E_ARG, "tighten_nodes: bad rounding");
  if (o.loop != MNTR_LOOP_FIXPOINT && o.loop != MNTR_LOOP_SIMPLEPRESOLVE) return fail(ctx, MNTR_E_ARG, "tighten_nodes: bad loop mode");
  if (o.handlers < 0 || o.handlers > 2) return fail(ctx, MNTR_E_ARG, "tighten_nodes: bad handlers");
  ctx->stats = mntr_gpu_stats{};
  const int32_t n = ctx->n;
  if ((rc = ensure_batch(ctx, n_boxes, true))) return rc;
  const int64_t ld = mntr_gpu_box_ld(n_boxes);

  // device scratch (grow-only, kept by the context): root, deltas, mod counts / offsets / tuples
  auto done = [&](int code) { return code; };
#define CUN(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { return fail(ctx, MNTR_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
  cudaStream_t s = ctx->stream;
  CUN(cudaEventRecord(ctx->ev[0], s));
  DeviceDeltas D;
  if ((rc = boxes_from_deltas(ctx, n_boxes, root_lb, root_ub, delta_ptr, delta_var, delta_is_upper, delta_val,
                              ctx->d_boxes, D))) return rc;
  double *d_mval = nullptr;
  long long *d_cnt = nullptr, *d_mptr = nullptr, *d_xcnt = nullptr, *d_xat = nullptr;
  int32_t *d_mvar = nullptr, *d_strip = nullptr; uint8_t *d_mup = nullptr;
  const int64_t n_strips = ((int64_t)n + 255) / 256;
  if ((rc = scratch(ctx, 6, sizeof(long long) * (size_t)n_boxes, (void **)&d_cnt))) return rc;
  if ((rc = scratch(ctx, 7, sizeof(long long) * (size_t)n_boxes, (void **)&d_xcnt))) return rc;
  if ((rc = scratch(ctx, 8, sizeof(long long) * ((size_t)n_boxes + 1), (void **)&d_mptr))) return rc;
  if ((rc = scratch(ctx, 9, sizeof(long long) * (size_t)n_boxes, (void **)&d_xat))) return rc;
  if ((rc = scratch(ctx, 10, sizeof(int32_t) * (size_t)(n_strips * ld), (void **)&d_strip))) return rc;
  if ((rc = scratch(ctx, 11, sizeof(int32_t) * (size_t)mod_cap, (void **)&d_mvar))) return rc;
  if ((rc = scratch(ctx, 12, (size_t)mod_cap, (void **)&d_mup))) return rc;
  if ((rc = scratch(ctx, 13, sizeof(double) * (size_t)mod_cap, (void **)&d_mval))) return rc;
  BatchIo io;
  io.boxes = ctx->d_boxes; io.ld = ld; io.n_boxes = n_boxes; io.rowflag = ctx->d_rowflag; io.varflag = ctx->d_varflag; io.tstate = ctx->d_tstate;
  io.verdict = ctx->d_verdict; io.rounds = ctx->d_rounds; io.nnz = ctx->d_nnzb; io.nl_evals = ctx->d_nl_evals;
  io.prepared = (ctx->prepared_boxes == (const void *)ctx->d_boxes && ctx->prepared_n == n_boxes) ? ctx->d_prepared : nullptr;
  ctx->prepared_boxes = nullptr;
  CUN(cudaMemsetAsync(ctx->d_nl_evals, 0, sizeof(unsigned long long), s));
  CUN(cudaEventRecord(ctx->ev[1], s));
  CUN(launch_batch_reference(ctx->lin, ctx->nl_loaded ? &ctx->nl : nullptr, io, o.rounding == MNTR_ROUND_DIRECTED,
                             o.loop, o.max_rounds, o.handlers != MNTR_HANDLERS_NONLINEAR,
                             (ctx->nl_loaded && o.handlers != MNTR_HANDLERS_LINEAR) ? 1 : 0, ctx->sm_count, s));
  CUN(cudaEventRecord(ctx->ev[2], s));
  // mods: per box the ordered count (strip offsets stay on the device) and the extras; offsets on the host (n_boxes
  // numbers), then emit in ascending (variable, side) order
  CUN(launch_count_mods(ctx->d_boxes, D.rl, D.ru, D.ptr, D.var, D.up, D.val, ld, n, n_boxes, d_strip, d_cnt, d_xcnt, s));
  std::vector<long long> cnt((size_t)n_boxes), xcnt((size_t)n_boxes), ptr((size_t)n_boxes + 1, 0), xat((size_t)n_boxes), hz((size_t)n_boxes);
  std::vector<int32_t> hv((size_t)n_boxes), hr((size_t)n_boxes);
  CUN(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(long long) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CUN(cudaMemcpyAsync(xcnt.data(), d_xcnt, sizeof(long long) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CUN(cudaMemcpyAsync(hv.data(), ctx->d_verdict, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CUN(cudaMemcpyAsync(hr.data(), ctx->d_rounds, sizeof(int32_t) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  CUN(cudaMemcpyAsync(hz.data(), ctx->d_nnzb, sizeof(long long) * (size_t)n_boxes, cudaMemcpyDeviceToHost, s));
  unsigned long long evals = 0;
  CUN(cudaMemcpyAsync(&evals, ctx->d_nl_evals, sizeof(evals), cudaMemcpyDeviceToHost, s));
  CUN(cudaStreamSynchronize(s));
  ctx->stats.nl_evals = (int64_t)evals;
  // an infeasible box reports no mods (the node is pruned; its box is not meaningful): it gets an empty range and a
  // start that is already past any capacity
  bool any_extra = false;
  for (int32_t b = 0; b < n_boxes; ++b) {
    const bool feas = hv[(size_t)b] == MNTR_FEASIBLE;
    ptr[(size_t)b + 1] = ptr[(size_t)b] + (feas ? cnt[(size_t)b] + xcnt[(size_t)b] : 0);
    xat[(size_t)b] = feas ? ptr[(size_t)b] + cnt[(size_t)b] : ((long long)1 << 60);
    any_extra = any_extra || (feas && xcnt[(size_t)b] > 0);
  }
  const long long total = ptr[(size_t)n_boxes];
  if (n_mods_out) *n_mods_out = (int64_t)total;
  for (int32_t b = 0; b <= n_boxes; ++b) mod_ptr[b] = (int64_t)ptr[(size_t)b];
  if (verdict) memcpy(verdict, hv.data(), sizeof(int32_t) * (size_t)n_boxes);
  if (rounds) memcpy(rounds, hr.data(), sizeof(int32_t) * (size_t)n_boxes);
  if (total > 0 && total <= mod_cap) {
    std::vector<long long> start(ptr.begin(), ptr.end() - 1);
    for (int32_t b = 0; b < n_boxes; ++b) if (hv[(size_t)b] != MNTR_FEASIBLE) start[(size_t)b] = (long long)1 << 60;
    CUN(cudaMemcpyAsync(d_mptr, start.data(), sizeof(long long) * (size_t)n_boxes, cudaMemcpyHostToDevice, s));
    CUN(cudaMemcpyAsync(d_xat, xat.data(), sizeof(long long) * (size_t)n_boxes, cudaMemcpyHostToDevice, s));
    CUN(launch_emit_mods(ctx->d_boxes, D.rl, D.ru, D.ptr, D.var, D.up, D.val, ld, n, n_boxes, d_strip, d_mptr, d_xat,
                         (long long)mod_cap, d_mvar, d_mup, d_mval, s));
    CUN(cudaMemcpyAsync(mod_var, d_mvar, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
    CUN(cudaMemcpyAsync(mod_is_upper, d_mup, (size_t)total, cudaMemcpyDeviceToHost, s));
    CUN(cudaMemcpyAsync(mod_val, d_mval, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  CUN(cudaEventRecord(ctx->ev[3], s));
  CUN(cudaStreamSynchronize(s));
#undef CUN
  // the device emits a box's mods in ascending (variable, side) order; only boxes with extras (a delta that loosened
  // the root bound) need their few extras merged in
  if (any_extra && total > 0 && total <= mod_cap) {
    std::vector<size_t> idx;
    std::vector<int32_t> tv; std::vector<uint8_t> tu; std::vector<double> tx;
    for (int32_t b = 0; b < n_boxes; ++b) {
      if (hv[(size_t)b] != MNTR_FEASIBLE || xcnt[(size_t)b] == 0) continue;
      const size_t lo = (size_t)ptr[(size_t)b], hi = (size_t)ptr[(size_t)b + 1];
      idx.resize(hi - lo);
      for (size_t k = 0; k < idx.size(); ++k) idx[k] = lo + k;
      std::sort(idx.begin(), idx.end(), [&](size_t x, size_t y) {
        return mod_var[x] != mod_var[y] ? mod_var[x] < mod_var[y] : mod_is_upper[x] < mod_is_upper[y]; });
      tv.resize(idx.size()); tu.resize(idx.size()); tx.resize(idx.size());
      for (size_t k = 0; k < idx.size(); ++k) { tv[k] = mod_var[idx[k]]; tu[k] = mod_is_upper[idx[k]]; tx[k] = mod_val[idx[k]]; }
      for (size_t k = 0; k < idx.size(); ++k) { mod_var[lo + k] = tv[k]; mod_is_upper[lo + k] = tu[k]; mod_val[lo + k] = tx[k]; }
    }
  }
  ctx->stats.h2d_ms = elapsed(ctx->ev[0], ctx->ev[1]);
  ctx->stats.kernel_ms = elapsed(ctx->ev[1], ctx->ev[2]);
  ctx->stats.d2h_ms = elapsed(ctx->ev[2], ctx->ev[3]);
  for (int32_t b = 0; b < n_boxes; ++b) {
    ctx->stats.nnz_updates += hz[(size_t)b];
    ctx->stats.n_infeasible += hv[(size_t)b] != MNTR_FEASIBLE;
    ctx->stats.max_rounds = std::max(ctx->stats.max_rounds, hr[(size_t)b]);
  }
  return done(MNTR_OK);
}

int mntr_gpu_nccl_unique_id(void *id128)
{
  if (!id128) return MNTR_E_ARG;
  NcclApi &nc = nccl_api();
  if (!nc.ok) return MNTR_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (nc.GetUniqueId(&id) != ncclSuccess) return MNTR_E_NCCL;
  memcpy(id128, &id, sizeof(id));
  return MNTR_OK;
}

static void free_xchg(mntr_gpu_ctx *ctx)
{
  if (ctx->rws.xsend) cudaFree(ctx->rws.xsend);
  if (ctx->rws.xrecv) cudaFree(ctx->rws.xrecv);
  if (ctx->h_xhdr) cudaFreeHost(ctx->h_xhdr);
  ctx->h_xhdr = nullptr;
  ctx->rws.xsend = nullptr; ctx->rws.xrecv = nullptr; ctx->rws.xcap = 0; ctx->rws.n_ranks = 1;
}

int mntr_gpu_comm_init(mntr_gpu_ctx *ctx, int32_t n_ranks, int32_t rank, const void *id128)
{
  if (!ctx) return MNTR_E_ARG;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id128) return fail(ctx, MNTR_E_ARG, "comm_init: bad argument");
  NcclApi &nc = nccl_api();
  if (!nc.ok) return fail(ctx, MNTR_E_NCCL, "comm_init: libnccl.so.2 could not be loaded");
  CU(cudaSetDevice(ctx->device));
  if (ctx->comm) { nc.CommDestroy(ctx->comm); ctx->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NC(nc.CommInitRank(&ctx->comm, n_ranks, id, rank));
  ctx->n_ranks = n_ranks; ctx->rank = rank;
  // buffers of the sparse bound exchange: a rank's message holds up to n/16 changed candidates (at most 2M; together
  // the messages of all ranks then still move less than the dense all-reduce).  MNTR_GPU_SPARSE_XCHG=0 switches the
  // exchange off, =<entries> sets the capacity.
  free_xchg(ctx);
  long long cap = std::min<long long>((long long)2 << 20, std::max<long long>(65536, (long long)ctx->n / 16));
  if (const char *e = getenv("MNTR_GPU_SPARSE_XCHG")) cap = atoll(e);
  if (cap > 0 && n_ranks > 1) {
    const size_t msg = sizeof(BoundMsg) * ((size_t)cap + 1);
    CU(cudaMalloc((void **)&ctx->rws.xsend, msg));
    CU(cudaMalloc((void **)&ctx->rws.xrecv, msg * (size_t)n_ranks));
    CU(cudaMallocHost((void **)&ctx->h_xhdr, sizeof(BoundMsg) * (size_t)n_ranks));
    ctx->rws.xcap = (int32_t)cap;
  }
  ctx->rws.n_ranks = n_ranks; ctx->rws.rank = rank;
  p2p_teardown(ctx);
  ctx->p2p_failed = false;
  return MNTR_OK;
}

int mntr_gpu_comm_destroy(mntr_gpu_ctx *ctx)
{
  if (!ctx) return MNTR_E_ARG;
  CU(cudaSetDevice(ctx->device));
  if (ctx->stream) CU(cudaStreamSynchronize(ctx->stream));
  p2p_teardown(ctx);
  ctx->rws.rank = 0;
  if (ctx->comm) { nccl_api().CommDestroy(ctx->comm); ctx->comm = nullptr; }
  ctx->n_ranks = 1; ctx->rank = 0;
  free_xchg(ctx);
  return MNTR_OK;
}

}  // extern "C"
