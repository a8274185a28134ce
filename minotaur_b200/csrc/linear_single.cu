// linear_single.cu -- K1: single-box FBBT of the linear rows, Jacobi rounds to a fixpoint inside ONE
// cooperative launch with ONE device-wide barrier per round (device-side change flag, no host round trips).
//
// Round r reads the box A = box[r&1] and merges candidate bounds into Z = box[(r+1)&1] with fp64 atomic max/min.
// Three things make a single barrier per round enough (SURVEY.md Appendix A gives the round; reference lines in
// brackets):
//   * max/min are commutative, so Z need not be a copy of A when the round starts: the variables that moved in round
//     r-1 (the only ones where Z lags behind A) are on a list, and round r brings Z up to date with the same atomics,
//     concurrently with the candidates of round r ("fix-up");
//   * integer rounding [tightenInts_ LinearHandler.cpp:415-490] is applied by the READER of a bound (bit 31 of the
//     stored column marks an integer variable), so the stored boxes hold merged, unrounded values and nobody has to
//     rewrite a box between rounds; the bound check [checkBounds_ :328-359] of a moved variable is part of its fix-up;
//   * the rows of a moved variable are flagged for the next round [changeBFlag_ :1229-1234] by the warp whose
//     candidate moved it first, through the CSC lists, into the next round's row bit set.
// The row evaluation itself (activities with outward rounding, singleton-infinity rule, implied bounds) comes in two
// forms.  RES (row_resident.cuh): every warp owns at most 32 rows, one per lane, whose heads stay in registers and
// whose entries stay in shared memory for the whole launch -- taken whenever m <= 32 x warps of the grid.  Otherwise
// row_batch.cuh: entry-parallel, one warp per staged batch of up to 32 due rows, streamed from the CSR.
// The result is exactly the two-phase Jacobi round: rows against the box of the round start, candidates merged with
// max/min, integers rounded, bounds checked, rows of changed variables flagged.
#include "device_problem.cuh"
#include "kernels.h"
#include "row_batch.cuh"
#include "row_resident.cuh"

#include <type_traits>

namespace mntr {

namespace {

constexpr int kSingleThreads = 768;      // 24 warps x 85 registers: the row evaluation must not spill (L1 is small)
constexpr int kSingleWarps = kSingleThreads / 32;

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// optional phase trace (MNTR_GPU_TRACE=1): thread 0 of block 0 stamps every phase boundary
#define MNTR_TRACE() do { if (W.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && tr < 32) W.trace[tr] = globaltimer_ns(); ++tr; } while (0)

// Device-wide barrier for a cooperative (co-resident) grid: a cumulative arrive counter; thread 0 of each block
// arrives with a release fence and polls.  The poll is ONE 16-byte acquire load of the line {arrive counter, last
// round that moved a bound, last round that moved an integer, sticky flags}: the load that sees the last arrival
// also carries the control words every block needs to decide what comes next (all of them were written, with
// atomics performed at L2, before their writers arrived), so no second trip to L2 follows the barrier.  Thread 0
// leaves the snapshot in shared memory: ONE reader per block.
constexpr int kTraceBlk = 64 + 256 * 16;   // trace buffer: per-block phase maxima start here
__device__ __forceinline__ uint4 ld_acquire_gpu_v4(const unsigned *p)
{
  uint4 v;
  asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned *sync, unsigned n_blocks, unsigned &target, unsigned *s_ctl,
                                             unsigned long long *arrive, unsigned stress_ns = 0u)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    if (arrive != nullptr) *arrive = globaltimer_ns();      // debug: when this block reached the barrier
    target += n_blocks;
    __threadfence();
    atomicAdd(sync, 1u);
    // test hook (MNTR_GPU_STRESS_BARRIER): one block takes its snapshot long after the others have gone on into the
    // next round and written its state into the same line -- the round tags must make that harmless
    if (stress_ns != 0u && blockIdx.x == 1) {
      const unsigned long long t0 = globaltimer_ns();
      while (globaltimer_ns() - t0 < stress_ns) __nanosleep(1000);
    }
    uint4 v;
    do { v = ld_acquire_gpu_v4(sync); } while (v.x < target);
    *reinterpret_cast<uint4 *>(s_ctl) = v;
  }
  __syncthreads();
}

// flag the rows of variable j (one thread; used for the few integer variables whose INCOMING bounds are fractional)
__device__ __forceinline__ void flag_rows_serial(const LinDev &P, int j, uint32_t *due)
{
  const int qb = __ldg(P.csc_ptr + j), qe = __ldg(P.csc_ptr + j + 1);
  for (int q = qb; q < qe; ++q) {
    const int row = __ldg(P.csc_row + q);
    atomicOr(due + (row >> 5), 1u << (row & 31));
  }
}

// The set bits of up to 32 words (lane l holds word wb + l) are spread over the lanes, 32 at a time, and f(j) is
// called for each: every marked variable is an independent chain, whatever word it sits in.  Convergent.
template <class F>
__device__ __forceinline__ void for_each_marked(unsigned word, int wb, int lane, F f, int first = 0)
{
  int incl = __popc(word);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  for (int base = first; base < total; base += 32) {
    const int x = base + lane;
    int lo = 0;                                      // first lane whose inclusive count exceeds x
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(kFullMask, incl, lo + step - 1);
      if (v <= x) lo += step;
    }
    const unsigned wword = __shfl_sync(kFullMask, word, lo);
    const int wincl = __shfl_sync(kFullMask, incl, lo);
    if (x < total) f((wb + lo) * 32 + (int)__fns(wword, 0, x - (wincl - __popc(wword)) + 1));
  }
}

template <class R, bool RES>
__global__ void __launch_bounds__(kSingleThreads, 1)
fbbt_single_jacobi_kernel(LinDev P, SingleWs W, double *lb_io, double *ub_io, int max_rounds, int loop_mode)
{
  using Stage = typename std::conditional<RES, ResidentStage, WarpStage>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ FixRound s_fix[2];                       // by round parity
  __shared__ __align__(16) unsigned s_ctl[4];         // SingleWs::sync as of the last barrier
  Stage &S = reinterpret_cast<Stage *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int nthreads = gridDim.x * blockDim.x;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;       // linear: coalesced sweeps over arrays
  // block-interleaved numbering: consecutive row ranges go to different SMs
  const int warp_g = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, n_warps = gridDim.x * kSingleWarps;
  unsigned bar_target = 0;
  int tr = 0;
  MNTR_TRACE();
  if (lane == 0) { S.tcount = 0; S.pad_[0] = 0; S.pad_[1] = 0; }
  if (threadIdx.x < 2) {      // round r (parity c = r & 1) merges into box[c ^ 1], marks touched[c], flags due[c ^ 1]
    const int c = threadIdx.x;
    s_fix[c] = FixRound{0u, 0u, W.box[c ^ 1], W.touched[c], W.due[c ^ 1], W.sync, P.csc_ptr, P.csc_row};
  }


  // contiguous row range of this warp
  const int rpw = (P.m + n_warps - 1) / n_warps;
  const int r0 = min(P.m, warp_g * rpw), r1 = min(P.m, r0 + rpw);
  // RES: the lane's row head stays in registers, the row's entries in the warp's slice of shared memory
  // (the head's loads are issued here and first used after the box has been set up: the trips overlap)
  RowHead head{0, -1, 0.0, 0.0};
  if constexpr (RES) head = load_head(P, r0 + lane < r1 ? r0 + lane : -1);

  // ---- phase 0: both boxes = incoming box; round 1's rows all due; moved-variable bit sets empty, except that
  //      integer variables whose INCOMING bounds are fractional move in round 1 by rounding alone: they are
  //      marked in touched[0], which round 1's fix-up scans ----
  //      The incoming box may live in pinned HOST memory (zero-copy end-to-end call): its loads are issued first and
  //      consumed last, so the trips over PCIe overlap everything else this phase does.
  {
    int cross0 = 0;
    auto take = [&](int j0, int j, bool in, double2 b, uint8_t ty) {
      bool frac = false;
      if (in) {
        W.box[0][j] = b;
        W.box[1][j] = b;
        double2 r = b;
        if (is_int_type(ty)) {
          tighten_int_bounds(r.x, r.y);
          frac = r.x != b.x || r.y != b.y;
        }
        if (r.x > r.y + kETol) cross0 = 1;       // still crossed after any round: reported after round 1
      }
      const unsigned fm = __ballot_sync(kFullMask, frac);
      if (lane == 0) { W.touched[0][j0 >> 5] = fm; W.touched[1][j0 >> 5] = 0u; W.ever[j0 >> 5] = 0u; }
    };
    // the first (for n <= threads of the grid: the only) trip of the box sweep: loads now
    const int jf0 = gtid - lane, jf = jf0 + lane;
    const bool inf = jf < P.n;
    double2 bf = make_double2(0.0, 0.0);
    uint8_t tyf = 4;
    if (inf) { bf = make_double2(lb_io[jf], ub_io[jf]); tyf = __ldg(P.var_type + jf); }
    // round 1 takes every row without looking at its bit set (:1618-1622)
    const int nw_rows = (P.m + 31) / 32;
    for (int w = gtid; w < nw_rows; w += nthreads) { W.due[0][w] = 0u; W.due[1][w] = 0u; }
    // the CSC lists (rows to flag when a variable moves) are first needed at the end of round 1, on its critical
    // path: pull them into L2 now (when they are small enough to stay there)
    if ((long long)P.csc_nnz + P.n < (8ll << 20)) {
      const char *cp = reinterpret_cast<const char *>(P.csc_ptr), *cr = reinterpret_cast<const char *>(P.csc_row);
      for (long long o = (long long)gtid * 128; o < (long long)(P.n + 1) * 4; o += (long long)nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(cp + o));
      for (long long o = (long long)gtid * 128; o < (long long)P.csc_nnz * 4; o += (long long)nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(cr + o));
    }
    if constexpr (RES) resident_fill(P, S, lane, head);
    if (jf0 < P.n) take(jf0, jf, inf, bf, tyf);
    for (int j0 = jf0 + nthreads; j0 < P.n; j0 += nthreads) {     // warp-uniform trip count: lanes = 32 consecutive j
      const int j = j0 + lane;
      const bool in = j < P.n;
      double2 b = make_double2(0.0, 0.0);
      uint8_t ty = 4;
      if (in) { b = make_double2(lb_io[j], ub_io[j]); ty = __ldg(P.var_type + j); }
      take(j0, j, in, b, ty);
    }
    if (cross0) atomicOr(W.sync + 2, kCtlInCross);
  }
  MNTR_TRACE();
  grid_barrier(W.sync, gridDim.x, bar_target, s_ctl, nullptr);
  MNTR_TRACE();

  unsigned long long my_nnz = 0, my_rows = 0;
  int round = 0, rounds_out = 0;
  int verdict = 0;
  int out_buf = 1;            // box to hand back
  bool out_round = false;     // ... with integer rounding applied
  bool final_check = false;   // the last round moved bounds that no later fix-up will check
  int my_changes = 0;         // (variable, round) pairs this thread found in the moved bit sets
  // contiguous range of words of the moved-variable bit sets this warp scans
  const int nw_vars = (P.n + 31) / 32;
  const int wpw = (nw_vars + n_warps - 1) / n_warps;
  const int fw0 = min(nw_vars, warp_g * wpw), fw1 = min(nw_vars, fw0 + wpw);

  while (verdict == 0) {
    ++round;
    const int cur = round & 1;
    const double2 *A = W.box[cur];
    double2 *Z = W.box[cur ^ 1];

    // the fix-up (below) scans the previous round's moved-variable bit set: fetch this warp's first words now, so
    // the load is in flight while the rows are evaluated
    uint32_t *Tp = W.touched[cur ^ 1];
    unsigned fw = (fw0 + lane < fw1) ? __ldcg(Tp + fw0 + lane) : 0u;

    // ---- fix-up, first half: the variables that moved in the previous round (Z lags behind A exactly there).  The
    //      first 32 of this warp's words -- one variable per lane -- have their loads issued NOW, ahead of the rows, and
    //      are merged after them: the trip to L2 rides along with the rows' own ----
    int fj = -1;
    double2 fv = make_double2(0.0, 0.0);
    uint8_t fty = 4;
    bool fmore = false;         // more than 32 marks in the first 32 words (rare)
    if (round > 1) {
      if (fw) { Tp[fw0 + lane] = 0u; atomicOr(W.ever + fw0 + lane, fw); }     // this lane owns word fw0 + lane in every round
      my_changes += __popc(fw);
      if (__any_sync(kFullMask, fw != 0u)) {
        int incl = __popc(fw);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(kFullMask, incl, d);
          if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(kFullMask, incl, 31);
        fmore = total > 32;
        int lo = 0;                                      // first lane whose inclusive count exceeds `lane`
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int v = __shfl_sync(kFullMask, incl, lo + step - 1);
          if (v <= lane) lo += step;
        }
        const unsigned wword = __shfl_sync(kFullMask, fw, lo);
        const int wincl = __shfl_sync(kFullMask, incl, lo);
        if (lane < total) {
          fj = (fw0 + lo) * 32 + (int)__fns(wword, 0, lane - (wincl - __popc(wword)) + 1);
          fv = __ldcg(A + fj);
          fty = __ldg(P.var_type + fj);
        }
      }
    }

    // ---- rows due in this round ----
    {
      const ReadPending rd{A, P.colx, round > 1};
      const SinkFix sink{&s_fix[cur], (unsigned)round, (W.trace != nullptr && blockIdx.x == 0 && threadIdx.x < 32) ? W.trace + 32 : nullptr,
                         (W.trace != nullptr && round < 16) ? W.trace + kTraceBlk + (blockIdx.x * 16 + round) * 8 : nullptr};
      if constexpr (RES) eval_due_resident<R>(P, rd, sink, S, W.due[cur], r0, r1, head, round == 1, lane, my_nnz, my_rows);
      else eval_due_stream<R>(P, rd, sink, S, W.due[cur], r0, r1, round == 1, lane, my_nnz, my_rows);
      // the objective cut-off row is evaluated in every round (the reference loops it to its own fixpoint in
      // every sweep, LinearHandler.cpp:1636-1640); the last warp takes it
      if (P.cut_cnt > 0 && warp_g == n_warps - 1) {
        const ReadPending rdc{A, P.cut_colx, round > 1};
        eval_long<R>(P.cut_val, rdc, sink, S, lane, 0, P.cut_cnt, -INFINITY, P.cut_rhs);
        if (lane == 0) { my_nnz += (unsigned long long)P.cut_cnt; ++my_rows; }
      }
    }
    // ---- fix-up, second half.  It only has to land before the barrier; the set bits of 32 words are spread over
    //      the lanes, so every moved variable is an independent chain ----
    {
      int bad = 0;
      auto fix = [&](int j, double2 v, uint8_t ty) {
        atomic_max_f64(&Z[j].x, v.x);
        atomic_min_f64(&Z[j].y, v.y);
        if (is_int_type(ty)) tighten_int_bounds(v.x, v.y);
        if (v.x > v.y + kETol) bad = 1;          // checkBounds_ of the box this round reads
      };
      if (round == 1) {
        // integer variables whose INCOMING bounds are fractional moved by rounding alone: only their rows need flagging
        for (int wb = fw0; wb < fw1; wb += 32) {
          const int w = wb + lane;
          const unsigned word = wb == fw0 ? fw : ((w < fw1) ? __ldcg(Tp + w) : 0u);
          if (word) { Tp[w] = 0u; atomicOr(W.ever + w, word); }
          my_changes += __popc(word);
          for_each_marked(word, wb, lane, [&](int j) { flag_rows_serial(P, j, W.due[0]); atomicMax(W.sync + 1, 2u); });
        }
      } else {
        if (fj >= 0) fix(fj, fv, fty);
        if (fmore) {
          for_each_marked(fw, fw0, lane, [&](int j) { fix(j, __ldcg(A + j), __ldg(P.var_type + j)); }, 32);
        }
        for (int wb = fw0 + 32; wb < fw1; wb += 32) {
          const int w = wb + lane;
          const unsigned word = (w < fw1) ? __ldcg(Tp + w) : 0u;
          if (word) { Tp[w] = 0u; atomicOr(W.ever + w, word); }
          my_changes += __popc(word);
          for_each_marked(word, wb, lane, [&](int j) { fix(j, __ldcg(A + j), __ldg(P.var_type + j)); });
        }
      }
      if (bad) atomicMax(W.sync + 2, (unsigned)round);      // round-tagged: see SingleWs::sync
    }
    if (W.trace != nullptr && round < 16 && lane == 0) atomicMax(W.trace + kTraceBlk + (blockIdx.x * 16 + round) * 8 + 6, globaltimer_ns());
    MNTR_TRACE();
    grid_barrier(W.sync, gridDim.x, bar_target, s_ctl,
                 (W.trace != nullptr && round < 16) ? W.trace + 64 + blockIdx.x * 16 + round : nullptr, W.stress_ns);
    MNTR_TRACE();

    // Decisions about round `round` from the snapshot.  A faster block may already be in round + 1 and have written
    // that round's tags into the line before this block's poller took the snapshot: every test below gives the same
    // answer in that case (a later tag proves that the loop went on after this round).
    const unsigned ctl_flags = s_ctl[2];
    if ((ctl_flags & kCtlRoundMask) == (unsigned)round) {   // the box this round read was already crossed: found after round-1
      verdict = 1; rounds_out = round - 1; out_buf = cur; out_round = true;
      break;
    }
    rounds_out = round;
    if (s_ctl[3] == (unsigned)round) {                 // activity-infeasible row: the box of the round start is handed back
      verdict = 2 /* MNTR_INFEAS_ROW */; out_buf = cur; out_round = round > 1;
      break;
    }
    out_buf = cur ^ 1; out_round = true;
    if (round == 1 && (ctl_flags & (kCtlRowCross | kCtlInCross))) { verdict = 1; break; }    // checkBounds_ after round 1
    const unsigned moved_tag = s_ctl[1];
    const bool any_changed = (moved_tag >> 1) >= (unsigned)round;
    const bool any_int = (moved_tag >> 1) > (unsigned)round || ((moved_tag >> 1) == (unsigned)round && (moved_tag & 1u));
    if (!any_changed) break;
    final_check = true;
    if (max_rounds > 0 && round >= max_rounds) break;
    if (loop_mode == 1) {   // LinearHandler::simplePresolve truncation, :1625-1627
      if (round >= 10) break;
      if (round >= 2 && !any_int) break;
    }
    final_check = false;
  }

  // ---- epilogue: hand the box back -- lb_io / ub_io already hold the bounds of every variable that never moved,
  //      so only the variables marked in `ever` (or moved in the last round) are written; publish counters ----
  if (out_round) {
    const double2 *O = W.box[out_buf];
    const uint32_t *Tl = W.touched[round & 1];
    int bad = 0;
    for (int wb = fw0; wb < fw1; wb += 32) {
      const int w = wb + lane;
      const unsigned last = (w < fw1) ? __ldcg(Tl + w) : 0u;
      const unsigned word = (w < fw1) ? (__ldcg(W.ever + w) | last) : 0u;
      if (final_check) my_changes += __popc(last);
      for_each_marked(word, wb, lane, [&](int j) {
        double2 b = __ldcg(O + j);
        if (is_int_type(__ldg(P.var_type + j))) tighten_int_bounds(b.x, b.y);
        lb_io[j] = b.x;
        ub_io[j] = b.y;
        if (b.x > b.y + kETol) bad = 1;            // bound check of what the last round moved (no fix-up follows)
      });
    }
    if (bad && final_check) atomicOr(W.sync + 2, kCtlFinalCross);
  }
  __shared__ unsigned long long s_nnz, s_rows;
  __shared__ int s_changes;
  if (threadIdx.x == 0) { s_nnz = 0ull; s_rows = 0ull; s_changes = 0; }
  __syncthreads();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    my_nnz += __shfl_xor_sync(0xffffffffu, my_nnz, off);
    my_rows += __shfl_xor_sync(0xffffffffu, my_rows, off);
  }
  my_changes = __reduce_add_sync(0xffffffffu, my_changes);
  if (lane == 0 && my_rows) { atomicAdd(&s_nnz, my_nnz); atomicAdd(&s_rows, my_rows); }
  if (lane == 0 && my_changes) atomicAdd(&s_changes, my_changes);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_rows) { atomicAdd(&W.counters[0], s_nnz); atomicAdd(&W.counters[1], s_rows); }
    if (s_changes) atomicAdd(&W.status[2], s_changes);
    if (blockIdx.x == 0) { W.status[1] = rounds_out; W.status[6] = verdict; }
    // last block out: publish the control block to the host (pinned, mapped: no copy operation follows the
    // kernel) and zero it for the next launch (no memset precedes the kernel)
    __threadfence();
    if (atomicAdd(W.done, 1u) == gridDim.x - 1) {
      __threadfence();
      const int32_t *src = reinterpret_cast<const int32_t *>(W.sync);      // the control block starts with sync[]
      int32_t v[kCtrlWords];
#pragma unroll
      for (int k = 0; k < kCtrlWords; ++k) v[k] = __ldcg(src + k);
#pragma unroll
      for (int k = 0; k < kCtrlWords; ++k) { W.result[k] = v[k]; reinterpret_cast<int32_t *>(W.sync)[k] = 0; }
    }
  }
  MNTR_TRACE();
}

template <class R>
cudaError_t launch_r(const LinDev &P, const SingleWs &W, double *lb, double *ub, int max_rounds,
                     int loop_mode, int sm_count, bool staged_only, cudaStream_t stream)
{
  // at least 8 rows per warp: small problems run on few blocks and pay a cheaper barrier
  long long want_warps = ((long long)P.m + 7) / 8;
  const long long var_warps = ((long long)P.n + 255) / 256;
  if (want_warps < var_warps) want_warps = var_warps;
  long long blocks = (want_warps + kSingleWarps - 1) / kSingleWarps;
  if (blocks > sm_count) blocks = sm_count;
  if (blocks < 1) blocks = 1;
  // resident rows: every warp owns at most 32 rows (one per lane)
  const long long n_warps = blocks * kSingleWarps;
  const bool resident = !staged_only && ((long long)P.m + n_warps - 1) / n_warps <= 32;
  auto kern = resident ? fbbt_single_jacobi_kernel<R, true> : fbbt_single_jacobi_kernel<R, false>;
  const size_t smem = (resident ? sizeof(ResidentStage) : sizeof(WarpStage)) * kSingleWarps;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  LinDev p = P; SingleWs w = W;
  void *args[] = { &p, &w, &lb, &ub, &max_rounds, &loop_mode };
  // cooperative launch: guarantees that all blocks are co-resident, which the barrier relies on
  return cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)blocks), dim3(kSingleThreads), args, smem, stream);
}

}  // namespace

cudaError_t launch_single_jacobi(const LinDev &P, const SingleWs &W, double *lb_dev, double *ub_dev,
                                 bool directed, int max_rounds, int loop_mode, int sm_count, bool staged_only,
                                 cudaStream_t stream)
{
  if (directed) return launch_r<RoundDirected>(P, W, lb_dev, ub_dev, max_rounds, loop_mode, sm_count, staged_only, stream);
  return launch_r<RoundNearest>(P, W, lb_dev, ub_dev, max_rounds, loop_mode, sm_count, staged_only, stream);
}

}  // namespace mntr
