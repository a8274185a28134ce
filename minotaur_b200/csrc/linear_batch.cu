// linear_batch.cu -- K3: thousands of B&B node boxes tightened in one launch, in the
// REFERENCE'S sweep order.
//
// Layout: boxes are node-minor, double2 {lb,ub} [n][ld]: the 32 boxes of a tile are
// contiguous for one variable, so a warp (lane = box) gathers a variable's bounds for its
// whole tile with one coalesced 512-byte request of 128-bit loads, while the CSR entries
// (col,val) are warp-uniform broadcast loads amortised over the 32 boxes.
//
// One CTA -- or, when tiles are scarce, one thread-block cluster of up to 8 CTAs -- owns one tile of 32 boxes
// for the whole call.  Because boxes are independent, the
// only ordering constraint of the reference's in-place, index-ordered Gauss-Seidel sweep
// (LinearHandler::varBndsFromCons_, LinearHandler.cpp:493-541) is between rows that share a
// variable; rows are therefore scheduled in wavefront levels (level = 1 + max level of an
// earlier row sharing a variable; rows are stored in level order at load time) and a level boundary is a plain
// __syncthreads() -- no grid-wide synchronisation, no atomics on bounds, no second buffer.
// Inside a level the warps of the CTA take rows round-robin.  With round-to-nearest
// arithmetic every lane performs exactly the reference's operation sequence (ascending
// column order, unfused mul/add), so results are bitwise those of the reference; with
// directed rounding (default) they are outward-rounded versions of the same.
//
// Per sweep, per box (reference lines in brackets):
//   rows   : linBndTighten_ [:952-1045] = getLfBnds_ [:1237-1258], getSingLfBnds_
//            [:1261-1319] (the flag machine kept as coded), infeasibility [:994-1015],
//            updateLfBoundsFromLb_ [:1048-1136], activity recomputation when the row changed
//            something [:1027-1032], updateLfBoundsFromUb_ [:1139-1226]; bFlag set on every
//            row of a changed variable [changeBFlag_ :1229-1234] through the CSC lists.
//   ints   : tightenInts_ [:415-490]   bounds: checkBounds_ [:328-359]
//   loop   : simplePresolve [:1605-1653] (<=10 rounds, rounds >=3 only while integer
//            variables moved) or fixpoint.
// Deviation, deliberate: an activity-infeasible row yields verdict MNTR_INFEAS_ROW and stops
// that box; the reference's node mode drops that status (:1631).
#include <cooperative_groups.h>
#include <cuda.h>
#include <cstdint>
#include <cstdlib>

#include "cgraph.cuh"
#include "device_problem.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace mntr {

namespace {

constexpr int kBatchWarps = 12;         // warps per CTA of the instantiation with CGraph tapes (85 registers per thread at two CTAs per SM: the interval code spills at 64)
constexpr int kLinWarps = 8;            // ... of the pure linear one (two 12 KB segment stages per warp, two CTAs per SM)
constexpr int kBatchThreads = kBatchWarps * 32;
constexpr int kGather = 4;             // independent 512-byte gathers a warp keeps in flight
constexpr unsigned kFull = 0xffffffffu;

// Per-tile control words.  They live in GLOBAL memory (BatchIo::tstate) because a tile may be worked on by a
// whole thread-block cluster: volatile accesses go to L2, cluster.sync() orders them.
struct TileShared {
  volatile int changed[32];
  volatile int nint[32];
  volatile int verdict[32];
  unsigned long long nnz[32];
};
static_assert(sizeof(TileShared) == kTileStateBytes, "BatchIo::tstate stride");

// the warps of ALL CTAs of the cluster that owns a tile, numbered consecutively
struct TileTeam {
  int gwarp, n_warps;       // this warp / warps in the cluster
  int cthread, n_threads;   // this thread / threads in the cluster
  bool solo;                // the cluster is one CTA: the level barrier is a plain bar.sync
  // bounds written with ordinary stores in this phase are read by bulk async copies (async proxy) in the next
  __device__ __forceinline__ void sync() const
  {
    asm volatile("fence.proxy.async;" ::: "memory");
    if (solo) __syncthreads();
    else cg::this_cluster().sync();
  }
};
__device__ __forceinline__ TileTeam make_team()
{
  cg::cluster_group cl = cg::this_cluster();
  TileTeam t;
  const int nb = (int)cl.num_blocks(), rk = (int)cl.block_rank();
  const int wpb = (int)(blockDim.x >> 5);
  t.gwarp = rk * wpb + (threadIdx.x >> 5);
  t.n_warps = nb * wpb;
  t.cthread = rk * (int)blockDim.x + threadIdx.x;
  t.n_threads = nb * (int)blockDim.x;
  t.solo = nb == 1;
  return t;
}

// A row is consumed in chunks of <= 32 entries: lane t fetches entry t (one coalesced request for the whole
// chunk) into this warp's slice of shared memory; every lane then reads entry after entry back as a
// broadcast (one LDS each for the coefficient and the column).  The {lb,ub} gathers of kGather consecutive
// terms are issued back to back before any of them is consumed, so a warp keeps kGather independent
// 512-byte requests in flight.  Terms are CONSUMED in ascending column order, one after the other,
// exactly like the reference's loop.  Only the row's true entries are visited (no padding), so the inner
// loops carry no per-term predicates.
struct RowStage {
  double *val;    // [32] this warp's slice
  int *col;       // [32]
  double2 *seg;   // [2][kSegEntries][32] {lb,ub} of a row's variables for the 32 boxes of the tile (TMA destination),
                  // two pipeline slots: the next row's segments are in flight while this row is evaluated
  uint64_t *bar;  // [2] mbarriers of the bulk copies into the two slots
  const void *tmap = nullptr;   // tensor map over the batch [n][2 ld] doubles (TMA gather4: four variables' segments per copy)
  int tile_x = 0;               // first double of this tile inside a variable's row: 64 * tile
};

// ---- bulk asynchronous copies (TMA, cp.async.bulk) global -> shared, completion on an mbarrier ----
constexpr int kSegEntries = 12;                      // rows of up to this many terms are staged
constexpr int kSegBytes = kTile * (int)sizeof(double2);   // 512 B: one variable, the 32 boxes of the tile
constexpr int kSegSlotBytes = kSegEntries * kSegBytes;                 // one pipeline slot of one warp: 6 KB
// dynamic shared memory of a warp of the pure linear instantiation: two segment slots, then per slot the row's
// coefficients (16 doubles) and columns (16 ints).  Addressed through the symbol below, so that the compiler knows the
// address space (LDS / STS with immediate offsets instead of generic loads with 64-bit address arithmetic).
constexpr int kEntSlotBytes = 16 * 8 + 16 * 4;
constexpr int kWarpSmemBytes = 2 * kSegSlotBytes + 2 * kEntSlotBytes;
constexpr int kSegSmemBytes = kLinWarps * kWarpSmemBytes;              // dynamic shared memory of a (pure linear) CTA
extern __shared__ __align__(128) unsigned char dyn_smem[];
// a CUtensorMap (128 opaque bytes, 64-byte aligned), passed to the kernel as a __grid_constant__ parameter
struct alignas(64) BoxTensorMap { unsigned long long opaque[16]; };
__device__ __forceinline__ double2 *slot_seg(int wl, int slot)
{
  return reinterpret_cast<double2 *>(dyn_smem + wl * kWarpSmemBytes + slot * kSegSlotBytes);
}
__device__ __forceinline__ double *slot_val(int wl, int slot)
{
  return reinterpret_cast<double *>(dyn_smem + wl * kWarpSmemBytes + 2 * kSegSlotBytes + slot * kEntSlotBytes);
}
__device__ __forceinline__ int *slot_col(int wl, int slot)
{
  return reinterpret_cast<int *>(dyn_smem + wl * kWarpSmemBytes + 2 * kSegSlotBytes + slot * kEntSlotBytes + 16 * 8);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The boxes stream through L2 once per sweep (gigabytes) while the matrix (megabytes) is re-read by every tile: the
// segment copies carry an evict-first policy and the matrix loads an evict-last one, so the stream does not push the
// matrix out (measured before: L2 hit rate 10 %, the rows' entry loads waiting on DRAM).
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, uint32_t bytes, uint64_t *bar, unsigned long long policy)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// TMA tile::gather4 (sm_100): ONE copy brings the 512-byte segments of FOUR variables (rows j0..j3 of the tensor
// [n][2 ld] doubles, 64 doubles starting at column x) into four consecutive 512-byte pieces of shared memory
__device__ __forceinline__ void tma_gather4_stream(void *dst, const void *tmap, int x, int j0, int j1, int j2, int j3,
                                                   uint64_t *bar, unsigned long long policy)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(j0), "r"(j1), "r"(j2), "r"(j3), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ int ldg_keep_i32(const int32_t *p, unsigned long long policy)
{
  int v;
  asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ double ldg_keep_f64(const double *p, unsigned long long policy)
{
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ int2 ldg_keep_i32x2(const int2 *p, unsigned long long policy)
{
  int2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.b32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ double2 ldg_keep_f64x2(const double2 *p, unsigned long long policy)
{
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Where a row's {lb,ub} come from.  BoxGlobal gathers them from the node-minor box array; BoxStaged reads the copy
// a bulk async copy put into shared memory (one 512-byte segment per term), writes a changed bound into that copy
// and through to global memory.
struct BoxGlobal {
  static constexpr bool kNeedsCol = true;
  static constexpr int kBatch = kGather;     // independent 512-byte gathers kept in flight
  double2 *bx; int64_t ld;
  __device__ __forceinline__ double2 load(int, int j) const { return bx[(int64_t)j * ld]; }
  __device__ __forceinline__ double2 *slot(int, int j) const { return bx + (int64_t)j * ld; }
  __device__ __forceinline__ void commit(int, int) const {}
};
struct BoxStaged {
  static constexpr bool kNeedsCol = false;   // the staged copy is addressed by the term's position
  static constexpr int kBatch = 2;           // shared-memory reads: no need to batch deeply, keep registers free
  double2 *seg;            // this lane's column of the staged segments: entry t at seg[t * 32]
  double2 *bx; int64_t ld;
  __device__ __forceinline__ double2 load(int t, int) const { return seg[t * kTile]; }
  __device__ __forceinline__ double2 *slot(int t, int) const { return seg + t * kTile; }
  // write-through; the fence orders this lane's store into seg before the next bulk copy into the same buffer
  __device__ __forceinline__ void commit(int t, int j) const
  {
    bx[(int64_t)j * ld] = seg[t * kTile];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
};

__device__ __forceinline__ int stage_chunk(const LinDev &P, int c0, int cnt_left, const RowStage &st, int lane)
{
  const int cnt = min(32, cnt_left);
  __syncwarp();                                   // the previous chunk has been consumed
  if (lane < cnt) { st.col[lane] = __ldg(P.col + c0 + lane); st.val[lane] = __ldg(P.val + c0 + lane); }
  __syncwarp();
  return cnt;
}

// A row of at most 32 entries is staged ONCE by its caller (stage_short) and stays staged for all passes over it
// (activity, FromLb, activity again, FromUb); longer rows are re-staged chunk by chunk.
__device__ __forceinline__ void stage_short(const LinDev &P, int beg, int cnt_row, const RowStage &st, int lane)
{
  if (cnt_row <= 32) (void)stage_chunk(P, beg, cnt_row, st, lane);
}
__device__ __forceinline__ int next_chunk(const LinDev &P, int beg, int c0, int cnt_row, const RowStage &st, int lane)
{
  return cnt_row <= 32 ? cnt_row : stage_chunk(P, beg + c0, cnt_row - c0, st, lane);
}

template <class R>
__device__ __forceinline__ void acc_term(double a, double2 b, double &ll, double &uu)
{
  const bool pos = a > 0.0;
  const double blo = pos ? b.x : b.y, bhi = pos ? b.y : b.x;
  ll = R::add_lo(ll, R::mul_lo(a, blo));
  uu = R::add_hi(uu, R::mul_hi(a, bhi));
}
// the largest reach |a| (ub - lb) of a row's terms (see row_update: a term whose reach is below the row's slack cannot
// move a bound); an undefined width (inf - inf) counts as infinite
__device__ __forceinline__ void acc_reach(double a, double2 b, double &wmax)
{
  const double d = b.y - b.x;
  const double w = fabs(a) * d * 1.000000001;
  wmax = fmax(wmax, (d == d) ? w : INFINITY);
}

// min / max activity of one row for this lane's box  [getLfBnds_].  Computed by every lane (lanes whose box
// is not due simply discard the result): no predication in the loop.
template <class R, class Box>
__device__ __forceinline__ void row_activity(const LinDev &P, int beg, int cnt_row, const Box &box,
                                             const RowStage &st, int lane, double &ll, double &uu, double &wmax)
{
  ll = 0.0; uu = 0.0; wmax = 0.0;
  for (int c0 = 0; c0 < cnt_row; c0 += 32) {
    const int cnt = next_chunk(P, beg, c0, cnt_row, st, lane);
    int t = 0;
    constexpr int kB = Box::kBatch;
    for (; t + kB <= cnt; t += kB) {
      double a[kB]; double2 b[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) { a[u] = st.val[t + u]; b[u] = box.load(t + u, Box::kNeedsCol ? st.col[t + u] : 0); }
#pragma unroll
      for (int u = 0; u < kB; ++u) { acc_term<R>(a[u], b[u], ll, uu); acc_reach(a[u], b[u], wmax); }
    }
    for (; t < cnt; ++t) {
      const double a = st.val[t];
      const double2 b = box.load(t, Box::kNeedsCol ? st.col[t] : 0);
      acc_term<R>(a, b, ll, uu); acc_reach(a, b, wmax);
    }
  }
}

// the largest reach alone (the activities of a staged row are already known to its caller)
template <class Box>
__device__ __forceinline__ double row_reach(const LinDev &P, int beg, int cnt_row, const Box &box, const RowStage &st, int lane)
{
  double wmax = 0.0;
  for (int c0 = 0; c0 < cnt_row; c0 += 32) {
    const int cnt = next_chunk(P, beg, c0, cnt_row, st, lane);
    int t = 0;
    constexpr int kB = Box::kBatch;
    for (; t + kB <= cnt; t += kB) {
      double a[kB]; double2 b[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) { a[u] = st.val[t + u]; b[u] = box.load(t + u, Box::kNeedsCol ? st.col[t + u] : 0); }
#pragma unroll
      for (int u = 0; u < kB; ++u) acc_reach(a[u], b[u], wmax);
    }
    for (; t < cnt; ++t) acc_reach(st.val[t], box.load(t, Box::kNeedsCol ? st.col[t] : 0), wmax);
  }
  return wmax;
}

// singleton-infinity activity [getSingLfBnds_], state machine as coded in the reference (rare path)
template <class R>
__device__ __forceinline__ void row_sing_activity(const LinDev &P, int beg, int end, const double2 *bx,
                                                  int64_t ld, bool need, double &slo, double &sup)
{
  double lb = 0.0, ub = 0.0;
  bool lo_sing = false, up_sing = false, lo_fin = true, up_fin = true;
  for (int t = beg; t < end; ++t) {
    const double a = __ldg(P.val + t);
    if (a == 0.0) continue;
    const int j = __ldg(P.col + t);
    if (!need) continue;
    const double2 b = bx[(int64_t)j * ld];
    if (a > kETol) {
      if (b.y < kInf20 && up_fin) ub = R::add_hi(ub, R::mul_hi(a, b.y));
      else if (up_sing) { up_sing = false; ub = INFINITY; up_fin = false; }
      else up_sing = true;
      if (b.x > -kInf20 && lo_fin) lb = R::add_lo(lb, R::mul_lo(a, b.x));
      else if (lo_sing) { lo_sing = false; lb = -INFINITY; lo_fin = false; }
      else lo_sing = true;
    } else if (a < -kETol) {
      if (b.y < kInf20 && lo_fin) lb = R::add_lo(lb, R::mul_lo(a, b.y));
      else if (lo_sing) { lo_sing = false; lb = -INFINITY; lo_fin = false; }
      else lo_sing = true;
      if (b.x > -kInf20 && up_fin) ub = R::add_hi(ub, R::mul_hi(a, b.x));
      else if (up_sing) { up_sing = false; ub = INFINITY; up_fin = false; }
      else up_sing = true;
    }
  }
  if (need) { slo = lb; sup = ub; }
}

// flag every row of variable j for the boxes in `mask`  [changeBFlag_], and remember that the
// variable moved (so the integer / bound sweep of this round only visits moved variables)
__device__ __forceinline__ void flag_rows_of(const LinDev &P, int j, unsigned mask, uint32_t *flags,
                                             uint32_t *varflag, int lane)
{
  const int b = __ldg(P.csc_ptr + j), e = __ldg(P.csc_ptr + j + 1);
  for (int q = b + lane; q < e; q += 32) atomicOr(flags + __ldg(P.csc_row + q), mask);
  if (varflag != nullptr && lane == 0) atomicOr(varflag + j, mask);
}

// exact candidate of one term for this lane's box (rare path); returns true when a bound moved
template <class R, bool FROM_LB>
__device__ __forceinline__ bool update_exact(double av, double numer, bool sing, double2 b, double2 *pb)
{
  const double aa = fabs(av);
  if (!(aa > kETol)) return false;
  const double vl = b.x, vu = b.y;
  // FromLb: a>0 raises lb, a<0 lowers ub.   FromUb: a>0 lowers ub, a<0 raises lb.
  const bool raise_lb = FROM_LB ? (av > 0.0) : (av < 0.0);
  const bool inf_side = raise_lb ? (vu >= kInf20) : (vl <= -kInf20);
  if (sing && !inf_side) return false;
  const double base = inf_side ? 0.0 : (raise_lb ? vu : vl);
  // ONE division by |a|: round_up(x/a) == -round_down(x/|a|) for a<0 (and x/a == -(x/|a|) exactly in
  // round-to-nearest), so the sign only selects the bound that moves
  if (raise_lb) {
    const double q = FROM_LB ? R::div_lo(numer, aa) : -R::div_hi(numer, aa);
    double cnd = R::add_lo(q, base);
    if (cnd > vl + kETol) {
      if (cnd > vu - kETol) cnd = vu;
      pb->x = cnd;
      return true;
    }
  } else {
    const double q = FROM_LB ? -R::div_lo(numer, aa) : R::div_hi(numer, aa);
    double cnd = R::add_hi(q, base);
    if (cnd < vu - kETol) {
      if (cnd < vl + kETol) cnd = vl;
      pb->y = cnd;
      return true;
    }
  }
  return false;
}

// updateLfBoundsFromLb_ (FROM_LB) / updateLfBoundsFromUb_ (!FROM_LB), in place.
// returns the ballot of lanes that changed something.
// slack = |row bound - activity| for lanes that take part, +inf for the others.  A candidate moves a bound of
// x_j towards the other by slack/|a|, so it can only be accepted when slack < |a|*(ub_j - lb_j): that product
// test (1e-9 relative margin; inf/NaN fall through) lets the warp skip the fp64 division for almost every term.
template <class R, bool FROM_LB, class Box>
__device__ __forceinline__ unsigned row_update(const LinDev &P, int beg, int cnt_row, const Box &box,
                                               const RowStage &st, bool doit, bool sing, double rbound, double act,
                                               uint32_t *flags, uint32_t *varflag, TileShared &sh, int lane,
                                               bool count_int = true)
{
  unsigned any = 0;
  // (row bound - activity): FromLb needs a lower estimate, FromUb an upper estimate
  const double numer = FROM_LB ? R::sub_lo(rbound, act) : R::sub_hi(rbound, act);
  const double slack = doit ? (FROM_LB ? -numer : numer) : INFINITY;
  // one term: product test, then (rarely) the exact candidate, the in-place store and the bFlag propagation
  auto term = [&](int t, double av, int j, double2 b) {
    const double reach = fabs(av) * (b.y - b.x) * 1.000000001;
    const bool maybe = !(slack > reach);
    if (!__any_sync(kFull, maybe)) return;            // nobody in the tile can move this variable
    bool chg = false;
    if (maybe && doit) chg = update_exact<R, FROM_LB>(av, numer, sing, b, box.slot(t, j));
    if (chg) box.commit(t, j);
    const unsigned m = __ballot_sync(kFull, chg);
    if (m) {
      any |= m;
      flag_rows_of(P, j, m, flags, varflag, lane);
      if (chg) {
        sh.changed[lane] = 1;
        if (count_int && is_int_type(__ldg(P.var_type + j))) sh.nint[lane] = 1;
      }
    }
  };
  for (int c0 = 0; c0 < cnt_row; c0 += 32) {
    const int cnt = next_chunk(P, beg, c0, cnt_row, st, lane);
    int t = 0;
    constexpr int kB = Box::kBatch;
    for (; t + kB <= cnt; t += kB) {
      double a[kB]; double2 b[kB]; int jj[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) { a[u] = st.val[t + u]; jj[u] = st.col[t + u]; b[u] = box.load(t + u, jj[u]); }
#pragma unroll
      for (int u = 0; u < kB; ++u) term(t + u, a[u], jj[u], b[u]);
    }
    for (; t < cnt; ++t) { const int j = st.col[t]; term(t, st.val[t], j, box.load(t, j)); }
  }
  return any;
}

// one linear row for the 32 boxes of the tile  [linBndTighten_ with apply_to_prob == false]
// (rows of at most 32 entries have been staged by the caller; `box` says where their {lb,ub} are)
// PRE: the caller (staged_row) has already added up the row's activities -- the same operations in the same order as
// row_activity -- and passes them in; only the largest reach is still to be taken.
template <class R, class Box, bool PRE = false>
__device__ __forceinline__ void process_row(const LinDev &P, int2 info, double2 bnd, const Box &box, double2 *bx,
                                            int64_t ld, const RowStage &st, bool mine, uint32_t *flags,
                                            uint32_t *varflag, TileShared &sh, int lane, unsigned long long &my_nnz,
                                            double pre_ll = 0.0, double pre_uu = 0.0)
{
  const int beg = info.x, cnt = info.y, end = beg + cnt;
  const double rl = bnd.x, ru = bnd.y;
  double ll, uu, wmax, sing_ll = -INFINITY, sing_uu = INFINITY;
  if constexpr (PRE) { ll = pre_ll; uu = pre_uu; wmax = row_reach(P, beg, cnt, box, st, lane); }
  else row_activity<R>(P, beg, cnt, box, st, lane, ll, uu, wmax);
  bool need_sing = mine && (ll < -kInf20 || uu > kInf20);
  if (__any_sync(kFull, need_sing)) row_sing_activity<R>(P, beg, end, bx, ld, need_sing, sing_ll, sing_uu);
  if (mine) my_nnz += (unsigned long long)cnt;
  if (mine && (ll > ru + kETol || uu < rl - kETol)) {       // :994-1015
    sh.verdict[lane] = 2;  /* MNTR_INFEAS_ROW */
    mine = false;
  }
  // row lb side  (:1017-1025)
  bool do_lb = false, s_lb = false; double act = 0.0;
  if (mine && rl > -kInf20) {
    if (uu < kInf20) { do_lb = true; act = uu; }
    else if (sing_uu < kInf20) { do_lb = true; s_lb = true; act = sing_uu; }
  }
  unsigned chg = 0;
  // the whole pass is skipped when no box of the tile has a slack below the row's largest reach: no term of the row
  // can move a bound then (the per-term test of row_update, taken once for the row; the common case)
  if (__any_sync(kFull, do_lb && !(-R::sub_lo(rl, act) > wmax)))
    chg = row_update<R, true>(P, beg, cnt, box, st, do_lb, s_lb, rl, act, flags, varflag, sh, lane);
  // recompute activities when FromLb changed something (:1027-1032); lanes that did not
  // change would recompute identical values, so the decision is taken per warp
  if (chg) {
    const bool redo = mine && ((chg >> lane) & 1u);
    double l2, u2, w2;
    row_activity<R>(P, beg, cnt, box, st, lane, l2, u2, w2);
    wmax = w2;          // this lane's box as it is now (lanes that changed nothing recompute what they had)
    if (redo) { ll = l2; uu = u2; }
    need_sing = redo && (ll < -kInf20 || uu > kInf20);
    if (__any_sync(kFull, need_sing)) row_sing_activity<R>(P, beg, end, bx, ld, need_sing, sing_ll, sing_uu);
  }
  // row ub side  (:1035-1043)
  bool do_ub = false, s_ub = false; act = 0.0;
  if (mine && ru < kInf20) {
    if (ll > -kInf20) { do_ub = true; act = ll; }
    else if (sing_ll > -kInf20) { do_ub = true; s_ub = true; act = sing_ll; }
  }
  if (__any_sync(kFull, do_ub && !(R::sub_hi(ru, act) > wmax)))
    (void)row_update<R, false>(P, beg, cnt, box, st, do_ub, s_ub, ru, act, flags, varflag, sh, lane);
}

// One STAGED row (at most kSegEntries terms, its {lb,ub} segments and entries in pipeline slot `slot` of warp `wl`) for
// the 32 boxes of the tile: process_row with the common case made cheap.  The activity pass reads shared memory through
// the dyn_smem symbol; the sign of a coefficient is warp-uniform (posmask, one bit per term), so the min / max selection
// is a uniform branch; and whether ANY term of the row can move a bound is decided once per row from
// amax * (largest width) >= every term's reach |a| (ub - lb) -- a conservative form of row_update's per-term test.
// Only rows that pass it run the update passes (through the generic code above).
template <class R>
__device__ __forceinline__ void staged_row(const LinDev &P, int2 info, double2 bnd, int wl, int slot, unsigned posmask,
                                           double amax, double2 *bx, int64_t ld, bool mine, uint32_t *flags, uint32_t *varflag,
                                           TileShared &sh, int lane, unsigned long long &my_nnz)
{
  const int cnt = info.y;
  const double rl = bnd.x, ru = bnd.y;
  const double2 *sg = slot_seg(wl, slot) + lane;
  const double *sv = slot_val(wl, slot);
  double ll = 0.0, uu = 0.0;
  int dhi = 0;            // high word of the largest width (ub - lb) seen: negative widths and zero never win, an
                          // infinite or undefined one makes the bound below infinite / NaN, i.e. "cannot skip"
  // one term: the sign of the coefficient is warp-uniform (posmask).  Rows whose coefficients are ALL positive
  // (knapsack, covering, packing rows: the common case) take a loop without the selection; the sums are the same
  // operations in the same order either way.
  auto term_pos = [&](double a, double2 b) {
    ll = R::add_lo(ll, R::mul_lo(a, b.x));
    uu = R::add_hi(uu, R::mul_hi(a, b.y));
    dhi = max(dhi, __double2hiint(b.y - b.x));
  };
  auto term = [&](int t, double a, double2 b) {
    const bool pos = (posmask >> t) & 1u;
    const double blo = pos ? b.x : b.y, bhi = pos ? b.y : b.x;
    ll = R::add_lo(ll, R::mul_lo(a, blo));
    uu = R::add_hi(uu, R::mul_hi(a, bhi));
    dhi = max(dhi, __double2hiint(b.y - b.x));
  };
  // four terms' loads in flight together; the remainder (0..3 terms) as a pair and a single one, straight-line
  auto sweep_terms = [&](auto &&one) {
    int t0 = 0;
#pragma unroll
    for (int c = 0; c < kSegEntries / 4; ++c) {
      if (t0 + 4 <= cnt) {
        double a[4]; double2 b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = sv[c * 4 + u]; b[u] = sg[(c * 4 + u) * kTile]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) one(c * 4 + u, a[u], b[u]);
        t0 += 4;
      }
    }
    const double *rv = sv + t0; const double2 *rg = sg + t0 * kTile;
    if (cnt - t0 >= 2) {
      const double a0 = rv[0], a1 = rv[1]; const double2 b0 = rg[0], b1 = rg[kTile];
      one(t0, a0, b0); one(t0 + 1, a1, b1);
      rv += 2; rg += 2 * kTile; t0 += 2;
    }
    if (t0 < cnt) one(t0, rv[0], rg[0]);
  };
  if (posmask == (cnt >= 32 ? kFull : ((1u << cnt) - 1u))) sweep_terms([&](int, double a, double2 b) { term_pos(a, b); });
  else sweep_terms(term);
  // >= every term's reach |a| (ub - lb) 1.000000001: amax >= |a|, and the width rounded up to the next high word
  const double wmax = amax * __hiloint2double(dhi + 1, 0) * 1.000000001;
  // SCREEN: does any box of the tile need more than the activity?  A box does when its row is activity-infeasible
  // (:994-1015), needs the singleton-infinity sums (:970-972), or has a side whose slack is within the row's largest
  // reach (only then can a term move a bound).  The common answer is no: one vote, and the row is done.  Otherwise
  // the row goes through process_row from the start (it recomputes the same activities from the staged segments).
  const double slack_lb = -R::sub_lo(rl, uu), slack_ub = R::sub_hi(ru, ll);       // as row_update forms them
  const bool more = mine && ((ll < -kInf20) | (uu > kInf20) | (ll > ru + kETol) | (uu < rl - kETol) |
                             ((rl > -kInf20) & !(slack_lb > wmax)) | ((ru < kInf20) & !(slack_ub > wmax)));
  if (!__any_sync(kFull, more)) {
    if (mine) my_nnz += (unsigned long long)cnt;
    return;
  }
  const RowStage st{slot_val(wl, slot), slot_col(wl, slot), nullptr, nullptr};
  const BoxStaged box{slot_seg(wl, slot) + lane, bx, ld};
  process_row<R, BoxStaged, true>(P, info, bnd, box, bx, ld, st, mine, flags, varflag, sh, lane, my_nnz, ll, uu);
}

// Objective cut-off row  c.x <= rhs  for the 32 boxes of the tile  [varBndsFromObj_, :544-597]: evaluated after
// the row sweep, by ONE warp (it may share variables with any row), and looped until it moves nothing more.
// Its modifications of integer variables do not count towards nintmods (the reference counts them into a local
// it never reads, :554), and an infeasible cut-off stops the box like an activity-infeasible row.
template <class R>
// (P and st BY VALUE: a reference would force the caller's copies of the kernel parameters into local memory, and the
// hot row loop would then read P.col / st.bar through the stack instead of the constant bank / registers)
// returns the nonzeros this lane's box visited (by value for the same reason: a reference to the caller's counter would
// put the counter of the hot row loop into local memory)
__device__ __noinline__ unsigned long long cutoff_row(const LinDev P, double2 *bx, int64_t ld, const RowStage st, bool run,
                                                      uint32_t *flags, uint32_t *varflag, TileShared &sh, int lane)
{
  unsigned long long my_nnz = 0ull;
  LinDev C = P;
  C.col = P.cut_col; C.val = P.cut_val;
  const int cnt = P.cut_cnt;
  bool mine = run && sh.verdict[lane] == 0;
  while (__any_sync(kFull, mine)) {
    double ll, uu, sing_ll = INFINITY, sing_uu = INFINITY;
    const BoxGlobal box{bx, ld};
    stage_short(C, 0, cnt, st, lane);
    double wmax;
    row_activity<R>(C, 0, cnt, box, st, lane, ll, uu, wmax);
    const bool need_sing = mine && (ll < -kInf20 || uu > kInf20);
    if (__any_sync(kFull, need_sing)) row_sing_activity<R>(C, 0, cnt, bx, ld, need_sing, sing_ll, sing_uu);
    if (mine) my_nnz += (unsigned long long)cnt;
    if (mine && ll > P.cut_rhs + kETol) { sh.verdict[lane] = 2; /* MNTR_INFEAS_ROW */ mine = false; }
    bool doit = false, sing = false;
    double act = 0.0;
    if (mine) {
      if (ll > -kInf20) { doit = true; act = ll; }
      else if (sing_ll > -kInf20) { doit = true; sing = true; act = sing_ll; }
    }
    unsigned chg = 0;
    if (__any_sync(kFull, doit))
      chg = row_update<R, false>(C, 0, cnt, box, st, doit, sing, P.cut_rhs, act, flags, varflag, sh, lane, false);
    mine = mine && ((chg >> lane) & 1u);
  }
  return my_nnz;
}

// integer rounding [tightenInts_] + lb>ub check [checkBounds_] of variable j for the lanes in `want`
__device__ __forceinline__ void finish_var(const LinDev &P, int j, double2 b, bool want, double2 *bx, int64_t ld,
                                           uint32_t *flags, TileShared &sh, int lane)
{
  const bool isint = is_int_type(__ldg(P.var_type + j));
  bool chg = false;
  if (want) {
    if (isint) {
      const double2 o = b;
      tighten_int_bounds(b.x, b.y);
      if (b.x != o.x || b.y != o.y) { bx[(int64_t)j * ld] = b; chg = true; }
    }
    if (b.x > b.y + kETol) sh.verdict[lane] = 1;       /* MNTR_INFEAS_BOUNDS; benign race */
  }
  if (isint) {
    const unsigned m = __ballot_sync(kFull, chg);
    if (m) {
      flag_rows_of(P, j, m, flags, nullptr, lane);
      if (chg) sh.changed[lane] = 1;
    }
  }
}

// LinearHandler::simplePresolve (loop_mode 1) or the status-honouring fixpoint (loop_mode 0) of the
// linear rows for the 32 boxes of this tile.  sh.verdict carries each box's verdict in and out;
// returns (per lane) the number of sweeps its box ran.
template <class R>
__device__ __forceinline__ int lin_tile_presolve(const LinDev &P, double2 *bx, int64_t ld, const RowStage &st,
                                                 uint32_t *flags, uint32_t *varflag, TileShared &sh, bool active, int loop_mode,
                                                 int max_rounds, int bad_row, unsigned long long &my_nnz,
                                                 bool &any_change, unsigned &seg_phase, bool prepared)
{
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const double2 *tile_base = bx - lane;          // box 0 of the tile: + j*ld is variable j's 512-byte segment
  const TileTeam team = make_team();
  const int warp = team.gwarp;
  // every row flagged for every box of the tile  (simplePresolve :1618-1622)
  for (int i = team.cthread; i < P.m; i += team.n_threads) __stcg(flags + i, __ldg(P.row_info + i).y >= 0 ? kFull : 0u);
  // (a PREPARED batch arrives with the variables its deltas set already flagged, see BatchIo::prepared)
  if (!prepared) for (int j = team.cthread; j < P.n; j += team.n_threads) __stcg(varflag + j, 0u);
  if (warp == 0) { sh.changed[lane] = 1; sh.nint[lane] = 0; }
  team.sync();

  const unsigned long long pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  int iters = 1;          // the reference's counter: starts at 1, ++ per sweep
  int my_rounds = 0;
  for (;;) {
    // ---- loop condition per box (registers are identical in every warp of the CTA) ----
    const int changed = sh.changed[lane];
    const int nint = sh.nint[lane];
    bool run = active && sh.verdict[lane] == 0 && changed;
    if (max_rounds > 0 && my_rounds >= max_rounds) run = false;
    if (loop_mode == 1) run = run && iters <= 10 && (iters <= 2 || nint > 0);   // :1625-1627
    const unsigned runmask = __ballot_sync(kFull, run);
    team.sync();                      // everybody has read the previous round's flags
    if (runmask == 0) break;
    if (warp == 0) { sh.changed[lane] = 0; sh.nint[lane] = 0; }
    const bool first_sweep = (iters == 1);
    ++iters;
    if (run) ++my_rounds;
    team.sync();

    // ---- rows, level by level; inside a level each warp owns a contiguous run of rows ----
    for (int lev = 0; lev < P.n_levels; ++lev) {
      const int qb = __ldg(P.level_ptr + lev), qe = __ldg(P.level_ptr + lev + 1);
      const int per = (qe - qb + team.n_warps - 1) / team.n_warps;
      const int q_lo = qb + warp * per, q_hi = min(qe, q_lo + per);
      // boxes already proven infeasible by a row stop sweeping (their result is final); sampled per level
      const unsigned alive = __ballot_sync(kFull, sh.verdict[lane] == 0);
      for (int q0 = q_lo; q0 < q_hi; q0 += 32) {
        // 32 rows' flag words with one coalesced request (flags live in L2: they are updated by atomics)
        const int q = q0 + lane;
        const uint32_t raw = (q < q_hi) ? __ldcg(flags + q) : 0u;
        const uint32_t fw = raw & runmask & alive;
        unsigned rows = __ballot_sync(kFull, fw != 0u);
        if (rows == 0u) continue;
        // the heads {first entry, count} / {lb, ub} of the chunk's rows come with one coalesced request each
        int2 myinfo = make_int2(0, 0);
        double2 mybnd = make_double2(0.0, 0.0);
        if (fw != 0u) { myinfo = ldg_keep_i32x2(P.row_info + q, pol_keep); mybnd = ldg_keep_f64x2(P.row_bnd + q, pol_keep); }
        if (fw != 0u) __stcg(flags + q, raw & ~fw);         // c_ptr->setBFlag(false), :513 (rows of a level share no variable,
                                                            // so nothing sets a flag of this chunk while it is worked on)
        if (st.seg == nullptr) {
          while (rows) {
            const int t = __ffs(rows) - 1;
            rows &= rows - 1;
            const int2 info = make_int2(__shfl_sync(kFull, myinfo.x, t), __shfl_sync(kFull, myinfo.y, t));
            const double2 bnd = make_double2(__shfl_sync(kFull, mybnd.x, t), __shfl_sync(kFull, mybnd.y, t));
            const bool mine = (__shfl_sync(kFull, fw, t) >> lane) & 1u;
            stage_short(P, info.x, info.y, st, lane);
            const BoxGlobal box{bx, ld};
            process_row<R>(P, info, bnd, box, bx, ld, st, mine, flags, varflag, sh, lane, my_nnz);
          }
          continue;
        }
        // ---- software pipeline over the chunk's due rows, three stages deep:
        //   A  entry t of the row after next is requested into registers (lane t),
        //   B  the next row's entries go to its slot and one 512-byte TMA bulk copy per term brings the {lb,ub} of that
        //      variable for the tile's 32 boxes into the slot's segments (all copies of the row in flight at once, landing
        //      on the slot's mbarrier),
        //   C  this row is evaluated from its slot.
        // Rows of a level share no variable, so a row's segments may be fetched while earlier rows of the level still
        // write bounds.  Rows longer than kSegEntries take no slot and gather straight from global memory.
        auto pop = [](unsigned &m) { const int t = m ? __ffs(m) - 1 : -1; m &= m - 1; return t; };
        auto entries = [&](int r, int &cnt, int &c, double &v) {        // stage A
          const int beg = __shfl_sync(kFull, myinfo.x, r);
          cnt = __shfl_sync(kFull, myinfo.y, r);
          c = 0; v = 0.0;
          if (cnt <= kSegEntries && lane < cnt) { c = ldg_keep_i32(P.col + beg + lane, pol_keep); v = ldg_keep_f64(P.val + beg + lane, pol_keep); }
        };
        // stage B; returns the row's sign mask (bit t: coefficient t positive) and largest |coefficient| (rounded up)
        auto launch = [&](int slot, int cnt, int c, double v, unsigned &posmask, double &amax) {
          posmask = __ballot_sync(kFull, v > 0.0);
          amax = (double)__uint_as_float(__reduce_max_sync(kFull, __float_as_uint(__double2float_ru(fabs(v)))));
          if (cnt > kSegEntries) return;
          if (lane < cnt) { slot_col(wl, slot)[lane] = c; slot_val(wl, slot)[lane] = v; }
          if (lane == 0) mbar_expect_tx(st.bar + slot, (uint32_t)(cnt * kSegBytes));
          if (st.tmap != nullptr) {
            // groups of four terms with one gather4 each (issued by lane g for terms 4g .. 4g+3), the 0..3 terms left
            // over with a bulk copy each
            const int j0 = __shfl_sync(kFull, c, (4 * lane) & 31), j1 = __shfl_sync(kFull, c, (4 * lane + 1) & 31),
                      j2 = __shfl_sync(kFull, c, (4 * lane + 2) & 31), j3 = __shfl_sync(kFull, c, (4 * lane + 3) & 31);
            const int full = cnt & ~3;
            if (4 * lane < full)
              tma_gather4_stream(slot_seg(wl, slot) + 4 * lane * kTile, st.tmap, st.tile_x, j0, j1, j2, j3, st.bar + slot, pol_stream);
            if (lane >= full && lane < cnt)
              bulk_g2s_stream(slot_seg(wl, slot) + lane * kTile, tile_base + (int64_t)c * ld, kSegBytes, st.bar + slot, pol_stream);
          } else
          if (lane < cnt) bulk_g2s_stream(slot_seg(wl, slot) + lane * kTile, tile_base + (int64_t)c * ld, kSegBytes, st.bar + slot, pol_stream);
          __syncwarp();
        };
        unsigned todo = rows;
        int rC = pop(todo), cntC, cA = 0, cntA = 0; double vA = 0.0;
        unsigned posC = 0u, posB = 0u; double amaxC = 0.0, amaxB = 0.0;
        int slot = 0;
        { int c; double v; entries(rC, cntC, c, v); launch(slot, cntC, c, v, posC, amaxC); }
        int rA = pop(todo);
        if (rA >= 0) entries(rA, cntA, cA, vA);
        while (rC >= 0) {
          const int rB = rA, cntB = cntA;
          if (rB >= 0) {
            launch(slot ^ 1, cntB, cA, vA, posB, amaxB);
            rA = pop(todo);
            if (rA >= 0) entries(rA, cntA, cA, vA);
          }
          const int2 info = make_int2(__shfl_sync(kFull, myinfo.x, rC), cntC);
          const double2 bnd = make_double2(__shfl_sync(kFull, mybnd.x, rC), __shfl_sync(kFull, mybnd.y, rC));
          const bool mine = (__shfl_sync(kFull, fw, rC) >> lane) & 1u;
          if (cntC <= kSegEntries) {
            mbar_wait(st.bar + slot, (seg_phase >> slot) & 1u);
            seg_phase ^= 1u << slot;
            staged_row<R>(P, info, bnd, wl, slot, posC, amaxC, bx, ld, mine, flags, varflag, sh, lane, my_nnz);
          } else {
            stage_short(P, info.x, info.y, st, lane);                 // chunk staging of long rows
            const BoxGlobal box{bx, ld};
            process_row<R>(P, info, bnd, box, bx, ld, st, mine, flags, varflag, sh, lane, my_nnz);
          }
          __syncwarp();                                               // the slot may be refilled
          rC = rB; cntC = cntB; posC = posB; amaxC = amaxB; slot ^= 1;
        }
      }
      team.sync();
    }

    // ---- objective cut-off row (only with an incumbent), :1636-1640 ----
    if (P.cut_cnt > 0) {
      if (warp == 0) my_nnz += cutoff_row<R>(P, bx, ld, st, run, flags, varflag, sh, lane);
      team.sync();
    }

    // ---- integer rounding + bound check ----
    if (first_sweep && !prepared) {
      // every variable once: the incoming box may hold fractional integer bounds or crossed bounds
      // (not for a prepared batch: its builder has checked the root box and flagged the variables the deltas set, so
      // rounding and the bound check are no-ops everywhere else and the flag-driven pass below visits the same set of
      // variables that can change or cross -- those, and the ones the rows of this sweep moved)
      for (int j0 = warp * kGather; j0 < P.n; j0 += team.n_warps * kGather) {
        double2 b[kGather];
        const bool want = run && sh.verdict[lane] != 2;     // row-infeasible boxes are frozen
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
          b[u] = make_double2(0.0, 0.0);
          if (want && j0 + u < P.n) b[u] = bx[(int64_t)(j0 + u) * ld];
        }
#pragma unroll
        for (int u = 0; u < kGather; ++u)
          if (j0 + u < P.n) finish_var(P, j0 + u, b[u], want, bx, ld, flags, sh, lane);
      }
      for (int j = team.cthread; j < P.n; j += team.n_threads) __stcg(varflag + j, 0u);
    } else {
      // later sweeps: only variables some row moved in this sweep can need rounding or can cross
      const int per = (P.n + team.n_warps - 1) / team.n_warps;
      const int j_lo = warp * per, j_hi = min(P.n, j_lo + per);
      for (int j0 = j_lo; j0 < j_hi; j0 += 32) {
        const int j = j0 + lane;
        const uint32_t vw = (j < j_hi) ? __ldcg(varflag + j) : 0u;
        if (vw != 0u) __stcg(varflag + j, 0u);
        const unsigned dead = __ballot_sync(kFull, sh.verdict[lane] == 2);
        unsigned vars = __ballot_sync(kFull, (vw & runmask & ~dead) != 0u);
        while (vars) {
          const int t = __ffs(vars) - 1;
          vars &= vars - 1;
          const uint32_t who = __shfl_sync(kFull, vw, t) & runmask & ~dead;
          const bool want = (who >> lane) & 1u;
          double2 b = make_double2(0.0, 0.0);
          if (want) b = bx[(int64_t)(j0 + t) * ld];
          finish_var(P, j0 + t, b, want, bx, ld, flags, sh, lane);
        }
      }
    }
    if (bad_row && run && warp == 0 && sh.verdict[lane] == 0) sh.verdict[lane] = 1;
    if (run && sh.changed[lane]) any_change = true;          // benign: read again after the barrier below
    team.sync();
    if (run && sh.changed[lane]) any_change = true;
  }
  return my_rounds;
}

// NlPresHandler::simplePresolve for the 32 boxes of this tile (NlPresHandler.cpp:1022-1059, no incumbent):
// at most two sweeps of { chkRed_ over all CGraph constraints ; varBndsFromCons_ in constraint index
// order, in place }.  The in-place order is reproduced by wavefront levels over the constraints
// (built at load time from the variables each constraint reads and writes).
// NlPresHandler::fixObjBins_ (NlPresHandler.cpp:1062-1121) for a LINEAR objective, for this lane's box: with olb the
// objective's lower bound over the box (LinearFunction::computeBounds, taken once), a binary with coefficient a0
// is fixed to 0 if a0>0 && olb+a0>ub, to 1 if a0<0 && olb-a0>ub; ub is the RAW incumbent value.  One warp per tile.
template <class R>
__device__ __noinline__ void fix_obj_bins(const LinDev P, double2 *bx, int64_t ld, bool run, TileShared &sh, int lane)
{
  const bool mine = run && sh.verdict[lane] == 0;
  if (!__any_sync(kFull, mine)) return;
  const double best = P.obj_ub;
  double olb = 0.0;
  for (int t = 0; t < P.cut_cnt; ++t) {
    const double a = __ldg(P.cut_val + t);
    const double2 b = bx[(int64_t)__ldg(P.cut_col + t) * ld];
    olb = R::add_lo(olb, R::mul_lo(a, a > 0.0 ? b.x : b.y));
  }
  if (!mine || olb <= -INFINITY) return;
  if (olb > best) { sh.verdict[lane] = 1; /* MNTR_INFEAS_BOUNDS */ return; }
  for (int t = 0; t < P.cut_cnt; ++t) {
    const int j = __ldg(P.cut_col + t);
    const uint8_t ty = __ldg(P.var_type + j);
    if (ty != 0 && ty != 2) continue;                    // Binary, ImplBin
    double2 *pb = bx + (int64_t)j * ld;
    const double2 b = *pb;
    if (!((b.y - b.x) > 1e-6)) continue;                 // NlPresHandler::eTol_
    const double a0 = __ldg(P.cut_val + t);
    // (olb + a0 > best) must hold in exact arithmetic for the fix to be valid: round the sum down
    if (a0 > 0.0 && R::add_lo(olb, a0) > best) { pb->y = 0.0; sh.changed[lane] = 1; }
    else if (a0 < 0.0 && R::add_lo(olb, -a0) > best) { pb->x = 1.0; sh.changed[lane] = 1; }
  }
}

template <class R, bool SHAPED>
__device__ __forceinline__ int nl_tile_presolve(const LinDev &P, const NlDev &N, double2 *bx, int64_t ld, TileShared &sh,
                                                bool active, bool &any_change, BatchStage &TS, unsigned &my_evals)
{
  const int lane = threadIdx.x & 31;
  const TileTeam team = make_team();
  const int warp = team.gwarp;
  const double2 *tile_base = bx - lane;
  double nlb[SHAPED ? 1 : kMaxTape], nub[SHAPED ? 1 : kMaxTape];      // the interpreter's node intervals
  if (warp == 0) sh.changed[lane] = 1;
  team.sync();
  int iters = 1, my_rounds = 0;
  for (;;) {
    const bool run = active && sh.verdict[lane] == 0 && sh.changed[lane] && iters <= 2;    // :1034-1035
    const unsigned runmask = __ballot_sync(kFull, run);
    team.sync();
    if (runmask == 0) break;
    if (warp == 0) sh.changed[lane] = 0;
    ++iters;
    if (run) ++my_rounds;
    team.sync();
    // chkRed_: read-only, every constraint against the box of the sweep start
    {
      // each warp takes a contiguous range of constraints, in staged batches
      const int per = (N.n_cons + team.n_warps - 1) / team.n_warps;
      const int c_lo = min(N.n_cons, warp * per), c_hi = min(N.n_cons, c_lo + per);
      for (int c0 = c_lo; c0 < c_hi;) {
        const BatchInfo B = stage_batch(N, c0, c_hi, TS, tile_base, ld, lane);
        // The box's verdict word lives in L2 (a cluster may share the tile): it is sampled once per staged batch, not
        // once per constraint -- a volatile load in front of every evaluation put a full L2 round trip before the
        // evaluation's own loads.  A finding of THIS warp stops its lane at once; another warp's finding is seen at
        // the next batch (the evaluations in between are of a box whose result is already discarded).
        bool alive = run && sh.verdict[lane] == 0;
        for (int k = 0; k < B.n; ++k) {
          const ConsView V = batch_constraint(N, TS, B, k, c0 + k);
          if (alive) {
            const int st = nl_chk_red<R, SHAPED>(V, bx, ld, nlb, nub);
            ++my_evals;
            if (st != 0) { sh.verdict[lane] = st; alive = false; }
          }
        }
        c0 += B.n;
      }
    }
    // ... and every QuadraticFunction constraint (the qf branch of chkRed_; these are only ever checked)
    if (N.n_quad > 0) {
      const int per = (N.n_quad + team.n_warps - 1) / team.n_warps;
      const int q_lo = min(N.n_quad, warp * per), q_hi = min(N.n_quad, q_lo + per);
      bool alive = run && sh.verdict[lane] == 0;
      for (int q = q_lo; q < q_hi; ++q) {
        if (alive) {
          const int st = quad_chk_red<R>(N, q, bx, ld);
          ++my_evals;
          if (st != 0) { sh.verdict[lane] = st; alive = false; }
        }
      }
    }
    team.sync();
    // varBndsFromCons_: constraints of one level touch disjoint variables
    bool moved = false;                                   // this lane's box changed in this sweep (published once, below)
    for (int lev = 0; lev < N.n_levels; ++lev) {
      const int qb = __ldg(N.level_ptr + lev), qe = __ldg(N.level_ptr + lev + 1);
      const int per = (qe - qb + team.n_warps - 1) / team.n_warps;
      const int c_lo = min(qe, qb + warp * per), c_hi = min(qe, c_lo + per);
      for (int c0 = c_lo; c0 < c_hi;) {
        const BatchInfo B = stage_batch(N, c0, c_hi, TS, tile_base, ld, lane);
        bool alive = run && sh.verdict[lane] == 0;         // (sampled per batch, see chkRed_ above)
        for (int k = 0; k < B.n; ++k) {
          const ConsView V = batch_constraint(N, TS, B, k, c0 + k);
          if (alive) {
            int n_mods = 0; unsigned dummy = 0;
            const int st = nl_var_bound_mods<R, SHAPED>(V, bx, ld, nlb, nub, n_mods, dummy);
            ++my_evals;
            if (st != 0) { sh.verdict[lane] = st; alive = false; }
            else if (n_mods > 0) moved = true;
          }
        }
        c0 += B.n;
      }
      if (lev + 1 == N.n_levels && moved) sh.changed[lane] = 1;
      team.sync();
    }
    // fixObjBins_: only with an incumbent and a linear objective (:1045-1050)
    if (P.cut_cnt > 0 && P.obj_ub < INFINITY) {
      if (warp == 0) fix_obj_bins<R>(P, bx, ld, run, sh, lane);
      team.sync();
    }
    if (run && sh.changed[lane]) any_change = true;
  }
  return my_rounds;
}

// One PCBProcessor::presolveNode_ pass per box (loop_mode 1): LinearHandler::presolveNode, then -- unless
// the box is already infeasible -- NlPresHandler::presolveNode (PCBProcessor.cpp:134-175); or (loop_mode 0)
// that pair repeated until neither handler changes a bound.
// Two instantiations.  HAS_NL = false, the pure linear one, stages the rows' {lb,ub} segments with TMA (96 KB of
// dynamic shared memory per CTA) and carries no tape interpreter.  HAS_NL = true gathers the linear rows' bounds
// directly (use_tma == 0: no dynamic shared memory), which leaves the L1 to the interpreter's per-thread node
// intervals, and stages the tapes instead.
template <class R, bool HAS_NL, bool SHAPED = false>
__global__ void __launch_bounds__(HAS_NL ? kBatchThreads : kLinWarps * 32, 2)
fbbt_batch_reference_kernel(LinDev P, NlDev N, BatchIo io, int loop_mode, int max_rounds, int lin_enabled,
                            int nl_enabled_arg, int use_tma, const __grid_constant__ BoxTensorMap tmap)
{
  constexpr int kWarps = HAS_NL ? kBatchWarps : kLinWarps;
  const int nl_enabled = HAS_NL ? nl_enabled_arg : 0;
  // per warp: chunk staging of long rows (and of the cut-off row), mbarriers of the two pipeline slots
  __shared__ double s_val[kWarps][32];
  __shared__ int s_col[kWarps][32];
  __shared__ __align__(8) uint64_t s_bar[kWarps][2];
  // dyn_smem: pure linear instantiation: kWarpSmemBytes per warp (segment slots + entries); with tapes: BatchStage per warp
  const int wl = threadIdx.x >> 5;
  const RowStage st{s_val[wl], s_col[wl], use_tma ? slot_seg(wl, 0) : nullptr, s_bar[wl],
                    use_tma == 2 ? (const void *)&tmap : nullptr,
                    64 * (int)(blockIdx.x / cg::this_cluster().num_blocks())};
  if ((threadIdx.x & 31) == 0) { mbar_init(st.bar, 1); mbar_init(st.bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  unsigned seg_phase = 0u;
  const TileTeam team = make_team();
  const int tile = blockIdx.x / (int)cg::this_cluster().num_blocks();
  TileShared &sh = *reinterpret_cast<TileShared *>(io.tstate + (int64_t)tile * kTileStateBytes);
  const int lane = threadIdx.x & 31;
  const int warp = team.gwarp;
  const int box = tile * 32 + lane;
  const bool active = box < io.n_boxes;
  double2 *bx = io.boxes + box;             // + j*ld addresses variable j of this lane's box
  const int64_t ld = io.ld;
  uint32_t *flags = io.rowflag + (int64_t)tile * P.m;
  uint32_t *varflag = io.varflag + (int64_t)tile * P.n;

  if (warp == 0) { sh.verdict[lane] = 0; sh.nnz[lane] = 0ull; }
  // checkBounds_ rows part is static: a row with lb > ub + eTol makes every box infeasible
  // (every CTA of the cluster scans all rows, so the flag is identical cluster-wide)
  int bad_row = 0;
  for (int i = threadIdx.x; i < P.m; i += (int)blockDim.x) {
    const double2 bnd = __ldg(P.row_bnd + i);
    if (__ldg(P.row_info + i).y >= 0 && bnd.x > bnd.y + kETol) bad_row = 1;
  }
  bad_row = __syncthreads_or(bad_row);
  team.sync();

  unsigned long long my_nnz = 0ull;
  unsigned my_evals = 0u;
  int my_rounds = 0;
  const bool prepared = io.prepared != nullptr && __ldg(io.prepared) == 0;
  for (int outer = 0;; ++outer) {
    bool lin_changed = false, nl_changed = false;
    if (lin_enabled)
      my_rounds += lin_tile_presolve<R>(P, bx, ld, st, flags, varflag, sh, active, loop_mode, max_rounds, bad_row,
                                        my_nnz, lin_changed, seg_phase, prepared && outer == 0);
    if constexpr (HAS_NL) {
      // the NL instantiation uses the dynamic shared memory for the tape batches (no TMA segments: use_tma == 0)
      if (nl_enabled) my_rounds += nl_tile_presolve<R, SHAPED>(P, N, bx, ld, sh, active, nl_changed, reinterpret_cast<BatchStage *>(dyn_smem)[wl], my_evals);
    }
    // fixpoint mode with both handlers: go round again while the nonlinear sweeps still move bounds
    const bool again = (loop_mode == 0) && lin_enabled && nl_enabled && nl_changed && sh.verdict[lane] == 0 &&
                       (max_rounds <= 0 || my_rounds < max_rounds) && outer < 50;
    if (!__syncthreads_or(again)) break;
  }

  // ---- per-box results ----
  if constexpr (HAS_NL) {
    const unsigned ev = __reduce_add_sync(kFull, my_evals);
    if (lane == 0 && ev != 0u && io.nl_evals != nullptr) atomicAdd(io.nl_evals, (unsigned long long)ev);
  }
  if (my_nnz) atomicAdd(&sh.nnz[lane], my_nnz);
  team.sync();
  if (warp == 0 && active) {
    io.verdict[box] = sh.verdict[lane];
    io.rounds[box] = my_rounds;
    io.nnz[box] = (long long)sh.nnz[lane];
  }
}

// ---------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------

// box-major lb/ub [nb][n]  ->  node-minor double2 [n][ld] at columns box0..box0+nb
__global__ void boxes_pack_kernel(const double *__restrict__ lb, const double *__restrict__ ub, int n,
                                  int box0, int nb, double2 *__restrict__ boxes, int64_t ld)
{
  __shared__ double tl[32][33], tu[32][33];
  const int j0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {      // r = box inside the tile
    const int b = b0 + r, j = j0 + threadIdx.x;
    if (b < nb && j < n) {
      tl[r][threadIdx.x] = lb[(int64_t)b * n + j];
      tu[r][threadIdx.x] = ub[(int64_t)b * n + j];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {      // r = variable inside the tile
    const int j = j0 + r, b = b0 + threadIdx.x;
    if (j < n && b < nb) boxes[(int64_t)j * ld + box0 + b] = make_double2(tl[threadIdx.x][r], tu[threadIdx.x][r]);
  }
}

__global__ void boxes_unpack_kernel(const double2 *__restrict__ boxes, int64_t ld, int n, int box0,
                                    int nb, double *__restrict__ lb, double *__restrict__ ub)
{
  __shared__ double tl[32][33], tu[32][33];
  const int j0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {      // r = variable
    const int j = j0 + r, b = b0 + threadIdx.x;
    if (j < n && b < nb) {
      const double2 v = boxes[(int64_t)j * ld + box0 + b];
      tl[r][threadIdx.x] = v.x;
      tu[r][threadIdx.x] = v.y;
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {      // r = box
    const int b = b0 + r, j = j0 + threadIdx.x;
    if (b < nb && j < n) {
      lb[(int64_t)b * n + j] = tl[threadIdx.x][r];
      ub[(int64_t)b * n + j] = tu[threadIdx.x][r];
    }
  }
}

__global__ void boxes_pad_kernel(double2 *boxes, int64_t ld, int n, int n_boxes)
{
  const int pad = (int)(ld - n_boxes);
  const int64_t total = (int64_t)n * pad;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = k / pad; const int p = (int)(k % pad);
    boxes[j * ld + n_boxes + p] = boxes[j * ld];
  }
}

// every box = the root box.  With `varflag` (a prepared batch, BatchIo::prepared): a variable whose ROOT bounds are not a
// fixed point of tightenInts_ + checkBounds_ (an integer variable with a fractional bound -- NlPresHandler leaves such
// bounds behind, CGraph.cpp:1627-1641 -- or crossed bounds) is flagged for every box of every tile, so the first sweep's
// flag-driven pass rounds / checks it like the variables the deltas set.
__global__ void boxes_from_root_kernel(const double *__restrict__ rl, const double *__restrict__ ru, int n,
                                       double2 *__restrict__ boxes, int64_t ld, const uint8_t *__restrict__ var_type,
                                       uint32_t *varflag)
{
  const int64_t total = (int64_t)n * ld;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = k / ld;
    const double2 b = make_double2(rl[j], ru[j]);
    boxes[k] = b;
    if (varflag != nullptr && k == j * ld) {
      double l = b.x, u = b.y;
      if (is_int_type(var_type[j])) tighten_int_bounds(l, u);
      if (l != b.x || u != b.y || b.x > b.y + kETol)
        for (int64_t t = 0; t < ld / 32; ++t) varflag[t * n + j] = kFull;
    }
  }
}

__global__ void apply_deltas_kernel(const long long *__restrict__ dptr, const int32_t *__restrict__ dvar,
                                    const uint8_t *__restrict__ dup, const double *__restrict__ dval,
                                    int n_boxes, double2 *boxes, int64_t ld, uint32_t *varflag, int n)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_boxes) return;
  for (long long q = dptr[b]; q < dptr[b + 1]; ++q) {      // in order: a later delta overrides
    double2 *p = boxes + (int64_t)dvar[q] * ld + b;
    if (dup[q]) p->y = dval[q]; else p->x = dval[q];
    // prepared batch: the variable goes through integer rounding + bound check in this box's first sweep
    if (varflag != nullptr) atomicOr(varflag + (int64_t)(b >> 5) * n + dvar[q], 1u << (b & 31));
  }
}

// mods of box b = every (variable, side) whose bound differs from the box's INITIAL bound (root box + the box's
// deltas, a later delta overriding an earlier one).  No copy of the initial boxes is kept: a block takes a strip of
// variables for one tile of 32 boxes (lane = box: coalesced 512-byte segments) and compares with the ROOT bound,
// which is the initial bound everywhere except at the box's own deltas; only where the final bound differs from
// the root's is the box's (short) delta list searched.  The sides a delta set and the sweep left EQUAL to the
// root's bound (possible only for a delta that loosens the root) are picked up by the *_delta kernels.
struct DeltaLists { const long long *ptr; const int32_t *var; const uint8_t *up; const double *val; };

__device__ __forceinline__ double2 initial_bounds(const DeltaLists &D, int b, int j, double2 root)
{
  double2 init = root;
  for (long long q = D.ptr[b]; q < D.ptr[b + 1]; ++q)
    if (D.var[q] == j) { if (D.up[q]) init.y = D.val[q]; else init.x = D.val[q]; }
  return init;
}

// grid (strips of 256 variables, tiles of 32 boxes), 8 warps: warp w walks the variables [j0 + 32 w, j0 + 32 w + 32) in
// ascending order, lane = box.  COUNT: the strip's count per box goes to strip_cnt[strip][box].  EMIT: the mods are
// written in ascending (variable, side) order at mod_ptr[box] + strip_off[strip][box] + (the warps before this one).
template <bool EMIT>
__global__ void __launch_bounds__(256)
mods_kernel(const double2 *__restrict__ boxes, const double *__restrict__ rl, const double *__restrict__ ru,
            DeltaLists D, int64_t ld, int n, int n_boxes, int32_t *strip_cnt, const long long *__restrict__ mod_ptr,
            long long cap, int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val)
{
  __shared__ int s_cnt[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const int j0 = blockIdx.x * 256 + w * 32, j1 = min(n, j0 + 32);
  const bool live = b < n_boxes;
  auto visit = [&](auto &&f) {
    for (int j = j0; j < j1; ++j) {
      const double2 v = boxes[(int64_t)j * ld + b];
      const double2 r = make_double2(__ldg(rl + j), __ldg(ru + j));
      if (v.x == r.x && v.y == r.y) continue;
      const double2 o = initial_bounds(D, b, j, r);
      const bool dl = v.x != r.x && v.x != o.x, du = v.y != r.y && v.y != o.y;
      if (dl || du) f(j, v, dl, du);
    }
  };
  int cnt = 0;
  if (live) visit([&](int, double2, bool dl, bool du) { cnt += (int)dl + (int)du; });
  s_cnt[w][lane] = cnt;
  __syncthreads();
  if (!EMIT) {
    if (w == 0 && live) {
      int tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += s_cnt[k][lane];
      strip_cnt[(int64_t)blockIdx.x * ld + b] = tot;
    }
    return;
  }
  if (!live || cnt == 0) return;
  long long at = mod_ptr[b] + (long long)strip_cnt[(int64_t)blockIdx.x * ld + b];     // here: the strip's exclusive offset
  for (int k = 0; k < w; ++k) at += s_cnt[k][lane];
  visit([&](int j, double2 v, bool dl, bool du) {
    if (dl) { if (at < cap) { mod_var[at] = j; mod_is_upper[at] = 0; mod_val[at] = v.x; } ++at; }
    if (du) { if (at < cap) { mod_var[at] = j; mod_is_upper[at] = 1; mod_val[at] = v.y; } ++at; }
  });
}

// per box: exclusive scan of the strips' counts (in place) and the box's total
__global__ void mods_scan_kernel(int32_t *strip_cnt, int n_strips, int64_t ld, int n_boxes, long long *mod_count)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_boxes) return;
  int run = 0;
  for (int s = 0; s < n_strips; ++s) {
    const int c = strip_cnt[(int64_t)s * ld + b];
    strip_cnt[(int64_t)s * ld + b] = run;
    run += c;
  }
  mod_count[b] = (long long)run;
}

// one thread per box: the sides its deltas set whose final bound EQUALS the root's but not the initial one (only a
// delta that LOOSENS the root bound can do that).  They are counted separately and emitted behind the box's ordered
// mods; the host merges them in for the (rare) boxes that have any.
template <bool EMIT>
__global__ void mods_delta_kernel(const double2 *__restrict__ boxes, const double *__restrict__ rl, const double *__restrict__ ru,
                                  DeltaLists D, int64_t ld, int n_boxes, long long *extra_count,
                                  const long long *__restrict__ extra_at, long long cap, int32_t *mod_var,
                                  uint8_t *mod_is_upper, double *mod_val)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_boxes) return;
  int cnt = 0;
  long long at = EMIT ? extra_at[b] : 0;
  for (long long q = D.ptr[b]; q < D.ptr[b + 1]; ++q) {
    const int j = D.var[q];
    const bool up = D.up[q] != 0;
    bool last = true;                 // only the last delta of a (variable, side) defines the initial bound
    for (long long q2 = q + 1; q2 < D.ptr[b + 1]; ++q2) if (D.var[q2] == j && (D.up[q2] != 0) == up) { last = false; break; }
    if (!last) continue;
    const double2 v = boxes[(int64_t)j * ld + b];
    const double fin = up ? v.y : v.x, root = up ? ru[j] : rl[j];
    if (fin != root || fin == D.val[q]) continue;
    if (!EMIT) { ++cnt; continue; }
    if (at < cap) { mod_var[at] = j; mod_is_upper[at] = up ? 1 : 0; mod_val[at] = fin; }
    ++at;
  }
  if (!EMIT) extra_count[b] = (long long)cnt;
}

}  // namespace

namespace {
// Tensor map over the batch for TMA gather4: a 2-D tensor of doubles, [n rows][2 ld], row stride 16 ld bytes, box
// {64 doubles = one tile's 512-byte segment, 1 row}: tile::gather4 takes four row indices per copy.  The driver entry
// point is fetched through the runtime (no link-time dependency on libcuda).  false: not available -> bulk copies.
bool encode_box_tensor_map(BoxTensorMap *out, const double2 *boxes, int64_t ld, int32_t n)
{
  static_assert(sizeof(BoxTensorMap) == sizeof(CUtensorMap), "CUtensorMap is 128 bytes");
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const encode_fn fn = []() -> encode_fn {          // (initialised once, thread-safe: the device group calls from several threads)
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return (encode_fn)p;
    (void)cudaGetLastError();
    return nullptr;
  }();
  if (fn == nullptr || n <= 0 || (reinterpret_cast<uintptr_t>(boxes) & 15u) != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * ld), (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)(16 * ld)};
  const cuuint32_t box[2] = {64u, 1u}, estr[2] = {1u, 1u};
  return fn(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)boxes, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class R>
cudaError_t launch_cluster(const LinDev &P, const NlDev &nl, const BatchIo &io, int loop_mode, int max_rounds,
                           int lin_enabled, int nl_enabled, int tiles, int cluster, cudaStream_t stream)
{
  BoxTensorMap tmap{};
  cudaLaunchConfig_t cfg = {};
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // one launch of `kernel` with `smem` bytes of dynamic shared memory.  A cluster of 16 CTAs is beyond the portable
  // size: the kernel opts in, and where that is refused or 16 CTAs of this shape do not fit a GPC the tile gets 8.
  auto launch = [&](auto kernel, unsigned threads, size_t smem, int use_tma) -> cudaError_t {
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int c = cluster;
    for (;;) {
      cfg.gridDim = dim3((unsigned)(tiles * c));
      attr[0].val.clusterDim.x = (unsigned)c;
      if (c <= 8) break;
      int n_clusters = 0;
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
          cudaOccupancyMaxActiveClusters(&n_clusters, kernel, &cfg) == cudaSuccess && n_clusters >= 1) break;
      (void)cudaGetLastError();
      c = 8;
    }
    return cudaLaunchKernelEx(&cfg, kernel, P, nl, io, loop_mode, max_rounds, lin_enabled, nl_enabled, use_tma, tmap);
  };
  if (nl_enabled) {
    const size_t smem = kBatchWarps * sizeof(BatchStage);
    if (nl.all_shaped)                // every tape is [Var,Var,Mult] or [Var,Var,Sqr,Sqr,SumList]: no interpreter
      return launch(fbbt_batch_reference_kernel<R, true, true>, kBatchThreads, smem, 0);
    return launch(fbbt_batch_reference_kernel<R, true>, kBatchThreads, smem, 0);
  }
  // segments by TMA: gather4 over a tensor map of the batch (default), else one bulk copy per term
  int use_tma = 1;
  const char *g4 = getenv("MNTR_GPU_GATHER4");
  if (!(g4 && atoi(g4) == 0) && encode_box_tensor_map(&tmap, io.boxes, io.ld, P.n)) use_tma = 2;
  return launch(fbbt_batch_reference_kernel<R, false>, kLinWarps * 32, kSegSmemBytes, use_tma);
}
}  // namespace

cudaError_t launch_batch_reference(const LinDev &P, const NlDev *N, const BatchIo &io, bool directed,
                                   int loop_mode, int max_rounds, int lin_enabled, int nl_enabled,
                                   int sm_count, cudaStream_t stream)
{
  const int tiles = (io.n_boxes + 31) / 32;
  if (tiles <= 0) return cudaSuccess;
  NlDev nl{};
  if (N != nullptr && nl_enabled) nl = *N; else nl_enabled = 0;
  // A tile's level-ordered sweep is a chain of short, latency-bound steps; when there are fewer tiles than
  // CTA slots, a thread-block CLUSTER of up to 8 CTAs (8 SMs) shares one tile and cluster.sync() is the
  // level barrier.  Two CTAs of this kernel fit on an SM.
  int cluster = 1;
  const int slots = 2 * sm_count;
  // Up to 8 CTAs per tile for the pure linear kernel; up to 16 -- beyond the portable cluster size, opted into at the
  // launch -- with CGraph tapes, where a level is more work per barrier (measured on 512 boxes = 16 tiles: C5 73.5 ->
  // 68.8 ms with 16, but the linear kernel on C3's rows 4.55 -> 6.70 ms: the barrier over 16 CTAs costs it more than
  // the halved shares give).  MNTR_GPU_CLUSTER_MAX overrides the limit.
  int cmax = nl_enabled ? 16 : 8;
  if (const char *e = getenv("MNTR_GPU_CLUSTER_MAX")) { const int c = atoi(e); if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) cmax = c; }
  while (cluster < cmax && tiles * cluster * 2 <= slots) cluster *= 2;
  if (const char *e = getenv("MNTR_GPU_CLUSTER")) { const int c = atoi(e); if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) cluster = c; }
  return directed ? launch_cluster<RoundDirected>(P, nl, io, loop_mode, max_rounds, lin_enabled, nl_enabled, tiles, cluster, stream)
                  : launch_cluster<RoundNearest>(P, nl, io, loop_mode, max_rounds, lin_enabled, nl_enabled, tiles, cluster, stream);
}

cudaError_t launch_boxes_pack(const double *lb_bm, const double *ub_bm, int32_t n, int32_t box0, int32_t nb,
                              double2 *boxes, int64_t ld, cudaStream_t stream)
{
  if (n <= 0 || nb <= 0) return cudaSuccess;
  dim3 grid((n + 31) / 32, (nb + 31) / 32), block(32, 8);
  boxes_pack_kernel<<<grid, block, 0, stream>>>(lb_bm, ub_bm, n, box0, nb, boxes, ld);
  return cudaGetLastError();
}

cudaError_t launch_boxes_unpack(const double2 *boxes, int64_t ld, int32_t n, int32_t box0, int32_t nb,
                                double *lb_bm, double *ub_bm, cudaStream_t stream)
{
  if (n <= 0 || nb <= 0) return cudaSuccess;
  dim3 grid((n + 31) / 32, (nb + 31) / 32), block(32, 8);
  boxes_unpack_kernel<<<grid, block, 0, stream>>>(boxes, ld, n, box0, nb, lb_bm, ub_bm);
  return cudaGetLastError();
}

cudaError_t launch_boxes_pad(double2 *boxes, int64_t ld, int32_t n, int32_t n_boxes, cudaStream_t stream)
{
  if (ld == n_boxes || n <= 0) return cudaSuccess;
  boxes_pad_kernel<<<256, 256, 0, stream>>>(boxes, ld, n, n_boxes);
  return cudaGetLastError();
}

cudaError_t launch_boxes_from_root(const double *root_lb, const double *root_ub, int32_t n, int32_t n_boxes,
                                   double2 *boxes, int64_t ld, const uint8_t *var_type, uint32_t *varflag, cudaStream_t stream)
{
  (void)n_boxes;
  if (n <= 0) return cudaSuccess;
  boxes_from_root_kernel<<<148 * 8, 256, 0, stream>>>(root_lb, root_ub, n, boxes, ld, var_type, varflag);
  return cudaGetLastError();
}

cudaError_t launch_count_mods(const double2 *boxes, const double *root_lb, const double *root_ub, const long long *delta_ptr,
                              const int32_t *delta_var, const uint8_t *delta_is_upper, const double *delta_val, int64_t ld,
                              int32_t n, int32_t n_boxes, int32_t *strip_cnt, long long *mod_count, long long *extra_count,
                              cudaStream_t stream)
{
  if (n <= 0 || n_boxes <= 0) return cudaSuccess;
  const DeltaLists D{delta_ptr, delta_var, delta_is_upper, delta_val};
  const int n_strips = (n + 255) / 256;
  dim3 grid(n_strips, (n_boxes + 31) / 32);
  mods_kernel<false><<<grid, 256, 0, stream>>>(boxes, root_lb, root_ub, D, ld, n, n_boxes, strip_cnt, nullptr, 0, nullptr, nullptr,
                                               nullptr);
  mods_scan_kernel<<<(n_boxes + 127) / 128, 128, 0, stream>>>(strip_cnt, n_strips, ld, n_boxes, mod_count);
  mods_delta_kernel<false><<<(n_boxes + 127) / 128, 128, 0, stream>>>(boxes, root_lb, root_ub, D, ld, n_boxes, extra_count, nullptr,
                                                                     0, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_emit_mods(const double2 *boxes, const double *root_lb, const double *root_ub, const long long *delta_ptr,
                             const int32_t *delta_var, const uint8_t *delta_is_upper, const double *delta_val, int64_t ld,
                             int32_t n, int32_t n_boxes, const int32_t *strip_off, const long long *mod_ptr,
                             const long long *extra_at, long long cap, int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val,
                             cudaStream_t stream)
{
  if (n <= 0 || n_boxes <= 0) return cudaSuccess;
  const DeltaLists D{delta_ptr, delta_var, delta_is_upper, delta_val};
  dim3 grid((n + 255) / 256, (n_boxes + 31) / 32);
  mods_kernel<true><<<grid, 256, 0, stream>>>(boxes, root_lb, root_ub, D, ld, n, n_boxes, const_cast<int32_t *>(strip_off), mod_ptr,
                                              cap, mod_var, mod_is_upper, mod_val);
  mods_delta_kernel<true><<<(n_boxes + 127) / 128, 128, 0, stream>>>(boxes, root_lb, root_ub, D, ld, n_boxes, nullptr, extra_at, cap,
                                                                    mod_var, mod_is_upper, mod_val);
  return cudaGetLastError();
}

cudaError_t launch_apply_deltas(const long long *delta_ptr, const int32_t *delta_var,
                                const uint8_t *delta_is_upper, const double *delta_val, int32_t n_boxes,
                                double2 *boxes, int64_t ld, uint32_t *varflag, int32_t n, cudaStream_t stream)
{
  if (n_boxes <= 0) return cudaSuccess;
  apply_deltas_kernel<<<(n_boxes + 127) / 128, 128, 0, stream>>>(delta_ptr, delta_var, delta_is_upper,
                                                                 delta_val, n_boxes, boxes, ld, varflag, n);
  return cudaGetLastError();
}

}  // namespace mntr
