// row_resident.cuh -- ROW-PARALLEL evaluation of the rows a warp OWNS, their entries resident in shared memory
// (Jacobi form of LinearHandler::linBndTighten_, LinearHandler.cpp:952-1045).
//
// The single-launch fixpoint kernel gives every warp a fixed range of at most 32 consecutive rows, one per lane, for
// the whole launch.  When the instance is small enough (m <= 32 x warps of the grid -- B&B node problems) the static
// part of a row never has to be fetched again after the first phase:
//   * the row head {first entry, term count, row lb, row ub} stays in the lane's registers;
//   * the first kRes entries {column | integer bit, coefficient} of the row stay in the warp's slice of shared memory,
//     transposed (entry t of the row of lane l at [t][l]: conflict-free); the 148 SMs together hold the matrix.
// A round then is, per due row, ONE dependent trip to L2 (the {lb,ub} gathers, up to four in flight per lane, all
// lanes of all warps at once) followed by register arithmetic:
//   pass 1  products a*blo (rounded down) / a*bhi (rounded up) added IN ASCENDING COLUMN ORDER [getLfBnds_ :1237-1258],
//           |a|(ub-lb) of every entry kept as a float rounded up; the singleton-infinity sums [getSingLfBnds_
//           :1261-1319] are a second, rare pass (only rows with an infinite activity)
//   pass 2  product test slack < |a|(ub-lb) per entry from registers; only entries that can move a bound re-derive the
//           exact candidate [updateLfBoundsFromLb_/Ub_ :1048-1226] and hand it to the Sink.
// Entries beyond kRes of a row are read from the CSR in global memory by the same lane; rows longer than kLaneMax are
// evaluated by the whole warp (eval_long in row_batch.cuh).  Compared with the staged, entry-parallel form
// (row_batch.cuh) there is no staging of products, no slot indirection, and a sparse round costs a warp a few
// hundred instructions instead of a few thousand -- sparse rounds are bound by exactly that serial path.
#pragma once
#include "row_batch.cuh"

namespace mntr {

constexpr int kRes = 12;        // resident entries per row
#ifndef MNTR_K1_GROUP
#define MNTR_K1_GROUP 4
#endif
#ifndef MNTR_K1_REACH_SMEM
#define MNTR_K1_REACH_SMEM 0
#endif
constexpr int kPassGroup = MNTR_K1_GROUP;   // gathers a lane has in flight in pass 1
constexpr int kResGroup = 4;    // exact candidates a lane derives per trip
constexpr int kLaneMax = 32;    // rows longer than this are evaluated by the whole warp
static_assert(kRes % kPassGroup == 0, "pass 1 runs in groups");
static_assert(kLaneMax <= 32, "pass 2 keeps one bit per entry");

// one warp's slice of shared memory
struct __align__(16) ResidentStage {
  double val[kRes][32];
  int32_t colx[kRes][32];
  int tcount;              // length of the Sink's list of moved variables
  int pad_[3];
  static constexpr int kListCap = 256;
  int items[kListCap];
  uint16_t work[32 * kLaneMax];     // (row lane | entry << 5) of the entries whose exact candidates are due
#if MNTR_K1_REACH_SMEM
  float reach[kRes][32];   // |a|(ub-lb) of the resident entries, rounded up (pass 1 -> pass 2)
#endif
  __device__ __forceinline__ int *list() { return items; }
};
static_assert(sizeof(ResidentStage) % 16 == 0, "slices are laid out back to back");

// phase 0: the lane's row head into registers, its first kRes entries into the slice
__device__ __forceinline__ RowHead resident_load(const LinDev &P, ResidentStage &S, int lane, int row)
{
  const RowHead h = load_head(P, row);
  if (h.cnt > 0) {
    const int keep = h.cnt < kRes ? h.cnt : kRes;
    // rows start at a multiple of four entries and are padded to one (val == 0): 128-bit loads
#pragma unroll
    for (int t = 0; t < kRes; t += 4) {
      if (t >= keep) break;
      const int4 c = __ldg(reinterpret_cast<const int4 *>(P.colx + h.beg + t));
      const double2 v0 = __ldg(reinterpret_cast<const double2 *>(P.val + h.beg + t));
      const double2 v1 = __ldg(reinterpret_cast<const double2 *>(P.val + h.beg + t + 2));
      S.colx[t][lane] = c.x; S.colx[t + 1][lane] = c.y; S.colx[t + 2][lane] = c.z; S.colx[t + 3][lane] = c.w;
      S.val[t][lane] = v0.x; S.val[t + 1][lane] = v0.y; S.val[t + 2][lane] = v1.x; S.val[t + 3][lane] = v1.y;
    }
  }
  return h;
}

// entry t of the lane's row
__device__ __forceinline__ void resident_entry(const LinDev &P, const ResidentStage &S, int lane, int beg, int t,
                                               double &a, int &cx)
{
  if (t < kRes) { a = S.val[t][lane]; cx = S.colx[t][lane]; }
  else { a = __ldg(P.val + beg + t); cx = __ldg(P.colx + beg + t); }
}

// One round of the rows of this warp: lane = row.  `due`: the lane's row is evaluated in this round (h.cnt >= 0).
// Both passes issue ALL the gathers of the resident entries before the first one is consumed: a round costs a row one
// trip to L2 for its activities and, if any of its entries can move a bound, one more for the exact candidates.
template <class R, class Sink>
__device__ __forceinline__ void eval_resident(const LinDev &P, const ReadPending &rd, const Sink &sink, ResidentStage &S,
                                              int lane, bool due, const RowHead h, bool first)
{
  if (first && due && h.rl > h.ru + kETol) sink.row_bounds_cross();     // checkBounds_, rows part (:328-359)
  const bool mine = due && h.cnt <= kLaneMax;
  unsigned longm = __ballot_sync(kFullMask, due && h.cnt > kLaneMax);
  const int cnt = mine ? h.cnt : 0;
  const int maxc = __reduce_max_sync(kFullMask, cnt);
  sink.mark(S, lane); sink.phase(lane, 0);

  // ---- pass 1: activities, terms in ascending column order ----
  double ll = 0.0, uu = 0.0;
  unsigned need = 0u;
  double slb = INFINITY, sub = INFINITY;
  uint8_t sg = 0;
  {
#if !MNTR_K1_REACH_SMEM
    float reach[kRes];
#endif
#pragma unroll
    for (int g = 0; g < kRes / kPassGroup; ++g) {
      if (g * kPassGroup >= maxc) {            // warp-uniform: no row of the warp reaches this group
#if !MNTR_K1_REACH_SMEM
#pragma unroll
        for (int u = 0; u < kPassGroup; ++u) reach[g * kPassGroup + u] = 0.f;
#endif
        continue;
      }
      double2 b[kPassGroup];
#pragma unroll
      for (int u = 0; u < kPassGroup; ++u) {
        const int t = g * kPassGroup + u;
        b[u] = make_double2(0.0, 0.0);
        if (t < cnt) b[u] = __ldcg(rd.box + (S.colx[t][lane] & kColMask));
      }
#pragma unroll
      for (int u = 0; u < kPassGroup; ++u) {
        const int t = g * kPassGroup + u;
#if !MNTR_K1_REACH_SMEM
        reach[t] = 0.f;
#endif
        if (t < cnt) {
          const double a = S.val[t][lane];
          if (rd.round_ints && S.colx[t][lane] < 0) tighten_int_bounds(b[u].x, b[u].y);
          const bool pos = a > 0.0;
          ll = R::add_lo(ll, R::mul_lo(a, pos ? b[u].x : b[u].y));
          uu = R::add_hi(uu, R::mul_hi(a, pos ? b[u].y : b[u].x));
          // |a|(ub-lb) with a 1e-9 relative margin, rounded UP to a float: conservative, inf/NaN fall through
          const float rch = __double2float_ru(fabs(a) * (b[u].y - b[u].x) * 1.000000001);
#if MNTR_K1_REACH_SMEM
          S.reach[t][lane] = rch;
#else
          reach[t] = rch;
#endif
        }
      }
    }
    for (int t = kRes; t < maxc; ++t) {         // tails of rows longer than the resident part (global CSR)
      if (t < cnt) {
        int j; bool isint;
        const double a = __ldg(P.val + h.beg + t);
        const double2 bt = rd.get(h.beg + t, j, isint);
        const bool pos = a > 0.0;
        ll = R::add_lo(ll, R::mul_lo(a, pos ? bt.x : bt.y));
        uu = R::add_hi(uu, R::mul_hi(a, pos ? bt.y : bt.x));
      }
    }
    sink.mark(S, lane); sink.phase(lane, 1);

    // ---- what the row offers its terms ----
    if (mine) {
      double sing_ll = -INFINITY, sing_uu = INFINITY;
      if (ll < -kInf20 || uu > kInf20) {          // singleton-infinity sums: finite sum + infinity count per side (rare)
        double fs_lo = 0.0, fs_hi = 0.0;
        int ninf_lo = 0, ninf_hi = 0;
        for (int t = 0; t < cnt; ++t) {
          double a; int cx, j; bool isint;
          resident_entry(P, S, lane, h.beg, t, a, cx);
          if (!(fabs(a) > kETol)) continue;
          const double2 bt = rd.fetch(cx, j, isint);
          const bool pos = a > 0.0;
          if (pos ? (bt.x <= -kInf20) : (bt.y >= kInf20)) ++ninf_lo; else fs_lo = R::add_lo(fs_lo, R::mul_lo(a, pos ? bt.x : bt.y));
          if (pos ? (bt.y >= kInf20) : (bt.x <= -kInf20)) ++ninf_hi; else fs_hi = R::add_hi(fs_hi, R::mul_hi(a, pos ? bt.y : bt.x));
        }
        sing_ll = (ninf_lo >= 2) ? -INFINITY : fs_lo;
        sing_uu = (ninf_hi >= 2) ? INFINITY : fs_hi;
      }
      if (ll > h.ru + kETol || uu < h.rl - kETol) sink.row_infeasible();     // :994-1015
      else row_offers<R>(h.rl, h.ru, ll, uu, sing_ll, sing_uu, slb, sub, sg);
    }

    // ---- pass 2: product test ----
#pragma unroll
    for (int t = 0; t < kRes; ++t) {
      if (t < cnt) {
#if MNTR_K1_REACH_SMEM
        const double rch = (double)S.reach[t][lane];
#else
        const double rch = (double)reach[t];
#endif
        if (!(slb > rch) || !(sub > rch)) need |= 1u << t;
      }
    }
  }
  if (slb < INFINITY || sub < INFINITY) {
    for (int t = kRes; t < cnt; ++t) {
      int j; bool isint;
      const double a = __ldg(P.val + h.beg + t);
      const double2 bt = rd.get(h.beg + t, j, isint);
      const double rch = fabs(a) * (bt.y - bt.x) * 1.000000001;
      if (!(slb > rch) || !(sub > rch)) need |= 1u << t;
    }
  }
  sink.mark(S, lane); sink.phase(lane, 2);

  // ---- exact candidates [updateLfBoundsFromLb_/Ub_] of the entries that passed the test.  They are few and spread
  //      unevenly over the rows, and the exact path is long (two directed fp64 divisions): the (row lane, entry)
  //      pairs of the whole warp are compacted into a work list and taken ENTRY-PARALLEL, 32 per trip, so the exact
  //      path runs with full lanes once or twice per warp instead of once per row and entry ----
  {
    const int mine_n = __popc(need);
    int incl = mine_n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(kFullMask, incl, 31);
    if (total > 0) {
      int pos = incl - mine_n;
      for (unsigned nd = need; nd; nd &= nd - 1) S.work[pos++] = (uint16_t)(lane | ((__ffs(nd) - 1) << 5));
      __syncwarp();
      for (int base = 0; base < total; base += 32) {
        if (sink.near_full(S, 32)) sink.flush(S, lane);          // warp-uniform (tcount is read behind a barrier)
        const bool on = base + lane < total;
        const int item = on ? S.work[base + lane] : 0;
        const int rl = item & 31, t = item >> 5;
        const double o_slb = __shfl_sync(kFullMask, slb, rl), o_sub = __shfl_sync(kFullMask, sub, rl);
        const int o_sg = __shfl_sync(kFullMask, (int)sg, rl), o_beg = __shfl_sync(kFullMask, h.beg, rl);
        if (on) {
          double a; int cx, j; bool isint;
          resident_entry(P, S, rl, o_beg, t, a, cx);
          const double2 bt = rd.fetch(cx, j, isint);
          emit_exact<R>(o_slb, o_sub, o_sg, a, j, isint, bt, S, sink);
        }
        __syncwarp();
      }
    }
  }
  sink.mark(S, lane); sink.phase(lane, 3);
  while (longm) {
    const int s = __ffs(longm) - 1;
    longm &= longm - 1;
    const int bg = __shfl_sync(kFullMask, h.beg, s), c = __shfl_sync(kFullMask, h.cnt, s);
    const double l = __shfl_sync(kFullMask, h.rl, s), u = __shfl_sync(kFullMask, h.ru, s);
    __syncwarp();
    if (sink.near_full(S, 32)) sink.flush(S, lane);
    eval_long<R>(P.val, rd, sink, S, lane, bg, c, l, u);
  }
  sink.flush(S, lane);
  sink.mark(S, lane); sink.phase(lane, 4);
}

// The due rows of this warp's range [r0, r0 + 32): lane l owns row r0 + l.  Which rows are due is a BIT SET (the
// reference's Constraint bFlag); what the warp takes it clears with fire-and-forget atomics [setBFlag(false), :513].
template <class R, class Sink>
__device__ __forceinline__ void eval_due_resident(const LinDev &P, const ReadPending &rd, const Sink &sink,
                                                  ResidentStage &S, uint32_t *due, int r0, int r1, const RowHead h,
                                                  bool first, int lane, unsigned long long &my_nnz,
                                                  unsigned long long &my_rows)
{
  if (r0 >= r1) return;
  unsigned m = r1 - r0 == 32 ? kFullMask : ((1u << (r1 - r0)) - 1u);     // bit l = row r0 + l
  if (!first) {                                      // in round 1 every row is due (:1618-1622)
    const int w0 = r0 >> 5, sh = r0 & 31;
    unsigned lo = 0u, hi = 0u;
    if (lane == 0) lo = __ldcg(due + w0);            // bits are set by L2 atomics: bypass L1
    if (lane == 1 && sh != 0 && (w0 + 1) * 32 < r1) hi = __ldcg(due + w0 + 1);
    lo = __shfl_sync(kFullMask, lo, 0); hi = __shfl_sync(kFullMask, hi, 1);
    m &= sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
    if (m == 0u) return;
    if (lane == 0 && (m << sh) != 0u) atomicAnd(due + w0, ~(m << sh));          // setBFlag(false), :513
    if (lane == 1 && sh != 0 && (m >> (32 - sh)) != 0u) atomicAnd(due + w0 + 1, ~(m >> (32 - sh)));
  }
  const bool is_due = ((m >> lane) & 1u) && h.cnt >= 0;      // deleted rows (term count < 0) are never evaluated
  if (is_due) { my_nnz += (unsigned long long)h.cnt; ++my_rows; }
  eval_resident<R>(P, rd, sink, S, lane, is_due, h, first);
}

}  // namespace mntr
