// linear_single.cu -- K1: single-box FBBT of the linear rows, Jacobi rounds to a fixpoint
// inside ONE cooperative launch (device-side change flag and work list, no host round trips).
//
// Per round (reference lines in brackets):
//   rows phase : sub-warp group per flagged CSR row, four entries per lane -- two 128-bit value
//                loads, one 128-bit column load, four 128-bit gathers of {lb,ub}; min/max activity
//                with outward rounding [getLfBnds_ LinearHandler.cpp:1237-1258]; singleton-
//                infinity sums by finite-sum + infinity-count [getSingLfBnds_ :1261-1319];
//                warp-shuffle butterfly reduction; activity infeasibility [:994-1015];
//                implied bounds [updateLfBoundsFromLb_/Ub_ :1048-1226] merged with fp64
//                atomic max/min into the next box.  A product test skips the fp64 division for terms
//                that cannot tighten anything.
//   vars phase : integer rounding [tightenInts_ :415-490], lb>ub check [checkBounds_
//                :328-359], change detection; the rows of every changed variable
//                [changeBFlag_ :1229-1234] are flagged in a bit set with fire-and-forget atomics
//                (a warp walks a changed variable's CSC list with one entry per lane).
// A device-wide barrier (one arrive counter, acquire polling) separates the phases; the loop
// condition is evaluated on the device from the round's change flags.
#include "device_problem.cuh"
#include "kernels.h"
#include "linear_row.cuh"

namespace mntr {

namespace {

constexpr int kSingleThreads = 1024;

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// optional phase trace (MNTR_GPU_TRACE=1): thread 0 of block 0 stamps every phase boundary
#define MNTR_TRACE() do { if (W.trace != nullptr && tid == 0 && tr < 64) W.trace[tr] = globaltimer_ns(); ++tr; } while (0)

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Device-wide barrier for a cooperative (co-resident) grid: a cumulative arrive counter, thread 0
// of each block arrives with a release fence and polls with acquire loads; the trailing fence also
// invalidates this SM's L1 so the block's ordinary loads see the other blocks' writes.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned n_blocks, unsigned &target)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n_blocks;
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_gpu(bar) < target) { }
    __threadfence();
  }
  __syncthreads();
}

template <int G, class R>
__global__ void __launch_bounds__(kSingleThreads, 1)
fbbt_single_jacobi_kernel(LinDev P, SingleWs W, double *lb_io, double *ub_io, int max_rounds,
                          int loop_mode)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  unsigned bar_target = 0;
  __shared__ int s_count;
  __shared__ unsigned long long s_nnz, s_rows;
  if (threadIdx.x == 0) { s_count = 0; s_nnz = 0ull; s_rows = 0ull; }
  int tr = 0;
  MNTR_TRACE();

  // ---- phase 0: build {lb,ub} boxes; every active row is on round 1's work list
  //      (simplePresolve :1618-1622), so the list is the identity and is not materialised ----
  int infeasible0 = 0;
  for (int j = tid; j < P.n; j += nthreads) {
    double2 b = make_double2(lb_io[j], ub_io[j]);
    W.box[j] = b;
    W.nbox[j] = b;
  }
  for (int i = tid; i < P.m; i += nthreads) {
    const double2 bnd = __ldg(P.row_bnd + i);
    if (__ldg(P.row_info + i).y >= 0 && bnd.x > bnd.y + kETol) infeasible0 = 1;   // checkBounds_, rows
  }
  for (int w = tid; w < (P.m + 31) / 32; w += nthreads) W.bits[w] = 0u;
  if (infeasible0) W.status[0] = 1 /* MNTR_INFEAS_BOUNDS */;
  MNTR_TRACE();
  grid_barrier(W.bar, gridDim.x, bar_target);
  MNTR_TRACE();
  volatile int32_t *vstatus = W.status;   // control words are re-read after every barrier
  volatile int32_t *vring = W.ring;

  unsigned long long my_nnz = 0, my_rows = 0;
  int round = 0;
  int verdict = vstatus[0];

  while (verdict == 0) {
    ++round;
    const int slot = round % 3;
    // ring[slot]: changed, ring[3+slot]: int moved
    if (tid == 0) { const int nx = (round + 1) % 3; W.ring[nx] = 0; W.ring[3 + nx] = 0; }

    // ------------------------------ rows phase ------------------------------
    {
      const SinkBox sink{W.nbox};
      process_rows<G, R>(P, W.box, W.bits, W.status, sink, tid >> 5, nthreads >> 5, round == 1, my_nnz, my_rows);
    }
    MNTR_TRACE();
    grid_barrier(W.bar, gridDim.x, bar_target);
    MNTR_TRACE();
    verdict = vstatus[0];
    if (verdict != 0) break;

    // ------------------------------ vars phase ------------------------------
    // a warp owns 32 consecutive variables; the rows of each changed variable are flagged by the
    // whole warp (lane q takes CSC entry q), so the test-and-set atomics of a variable are all in
    // flight together instead of one after the other
    int changed = 0, int_moved = 0, bad = 0;
    int n_changed = 0;
    {
      const int warp_g = tid >> 5, n_warps = nthreads >> 5;
      for (int j0 = warp_g * 32; j0 < P.n; j0 += n_warps * 32) {
        const int j = j0 + lane;
        bool ch = false;
        if (j < P.n) {
          const double2 o = W.box[j];
          double2 v = W.nbox[j];
          if (is_int_type(__ldg(P.var_type + j))) {
            if (v.x != o.x || v.y != o.y) int_moved = 1;    // row-derived mod on an int var (nintmods)
            tighten_int_bounds(v.x, v.y);
          }
          if (v.x > v.y + kETol) bad = 1;
          if (v.x != o.x || v.y != o.y) { ch = true; W.box[j] = v; W.nbox[j] = v; }
        }
        unsigned chm = __ballot_sync(0xffffffffu, ch);
        if (ch) { changed = 1; ++n_changed; }
        while (chm) {
          const int t = __ffs(chm) - 1;
          chm &= chm - 1;
          const int qb = __ldg(P.csc_ptr + j0 + t), qe = __ldg(P.csc_ptr + j0 + t + 1);
          // lane q flags CSC entry q: fire-and-forget OR into the row bit set
          for (int q = qb + lane; q < qe; q += 32) {
            const int row = __ldg(P.csc_row + q);
            atomicOr(W.bits + (row >> 5), 1u << (row & 31));
          }
        }
      }
    }
    // block-level reduction of the counters: one global atomic per block, not per thread
    n_changed = __reduce_add_sync(0xffffffffu, n_changed);
    if (lane == 0 && n_changed) atomicAdd(&s_count, n_changed);
    changed = __syncthreads_or(changed);
    int_moved = __syncthreads_or(int_moved);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
      if (s_count) { atomicAdd(&W.status[2], s_count); s_count = 0; }
      if (changed) W.ring[slot] = 1;
      if (int_moved) W.ring[3 + slot] = 1;
      if (bad) W.status[0] = 1 /* MNTR_INFEAS_BOUNDS */;
    }
    MNTR_TRACE();
    grid_barrier(W.bar, gridDim.x, bar_target);
    MNTR_TRACE();
    verdict = vstatus[0];
    const int any_changed = vring[slot];
    const int any_int = vring[3 + slot];
    if (verdict != 0 || !any_changed) break;
    if (max_rounds > 0 && round >= max_rounds) break;
    if (loop_mode == 1) {   // LinearHandler::simplePresolve truncation, :1625-1627
      if (round >= 10) break;
      if (round >= 2 && !any_int) break;
    }
  }

  // ---- epilogue: hand the box back, publish counters ----
  for (int j = tid; j < P.n; j += nthreads) {
    const double2 b = W.box[j];
    lb_io[j] = b.x;
    ub_io[j] = b.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    my_nnz += __shfl_xor_sync(0xffffffffu, my_nnz, off);
    my_rows += __shfl_xor_sync(0xffffffffu, my_rows, off);
  }
  if (lane == 0 && my_rows) { atomicAdd(&s_nnz, my_nnz); atomicAdd(&s_rows, my_rows); }
  __syncthreads();
  if (threadIdx.x == 0 && s_rows) { atomicAdd(&W.counters[0], s_nnz); atomicAdd(&W.counters[1], s_rows); }
  if (tid == 0) W.status[1] = round;
  MNTR_TRACE();
}

template <int G, class R>
cudaError_t launch_g(const LinDev &P, const SingleWs &W, double *lb, double *ub, int max_rounds,
                     int loop_mode, int sm_count, cudaStream_t stream)
{
  auto kern = fbbt_single_jacobi_kernel<G, R>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSingleThreads, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  long long want_threads = (long long)P.m * G;
  if (want_threads < P.n) want_threads = P.n;
  long long want_blocks = (want_threads + kSingleThreads - 1) / kSingleThreads;
  long long max_blocks = (long long)per_sm * sm_count;
  int blocks = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
  if (blocks < 1) blocks = 1;
  LinDev p = P; SingleWs w = W;
  void *args[] = { &p, &w, &lb, &ub, &max_rounds, &loop_mode };
  // cooperative launch: guarantees that all blocks are co-resident, which the barrier relies on
  return cudaLaunchCooperativeKernel((void *)kern, dim3(blocks), dim3(kSingleThreads), args, 0, stream);
}

template <class R>
cudaError_t launch_r(int G, const LinDev &P, const SingleWs &W, double *lb, double *ub,
                     int max_rounds, int loop_mode, int sm_count, cudaStream_t stream)
{
  switch (G) {
  case 2:  return launch_g<2, R>(P, W, lb, ub, max_rounds, loop_mode, sm_count, stream);
  case 4:  return launch_g<4, R>(P, W, lb, ub, max_rounds, loop_mode, sm_count, stream);
  case 8:  return launch_g<8, R>(P, W, lb, ub, max_rounds, loop_mode, sm_count, stream);
  case 16: return launch_g<16, R>(P, W, lb, ub, max_rounds, loop_mode, sm_count, stream);
  default: return launch_g<32, R>(P, W, lb, ub, max_rounds, loop_mode, sm_count, stream);
  }
}

}  // namespace

cudaError_t launch_single_jacobi(const LinDev &P, const SingleWs &W, double *lb_dev, double *ub_dev,
                                 int lanes_per_row, bool directed, int max_rounds, int loop_mode,
                                 int sm_count, cudaStream_t stream)
{
  if (directed)
    return launch_r<RoundDirected>(lanes_per_row, P, W, lb_dev, ub_dev, max_rounds, loop_mode, sm_count, stream);
  return launch_r<RoundNearest>(lanes_per_row, P, W, lb_dev, ub_dev, max_rounds, loop_mode, sm_count, stream);
}

}  // namespace mntr
