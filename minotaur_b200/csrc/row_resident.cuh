// row_resident.cuh -- ROW-PARALLEL evaluation of the rows a warp OWNS, their entries resident in shared memory
// (Jacobi form of LinearHandler::linBndTighten_, LinearHandler.cpp:952-1045).
//
// The single-launch fixpoint kernel gives every warp a fixed range of at most 32 consecutive rows, one per lane, for
// the whole launch.  When the instance is small enough (m <= 32 x warps of the grid -- B&B node problems) the static
// part of a row never has to be fetched again after the first phase:
//   * the row head {first entry, term count, row lb, row ub} stays in the lane's registers;
//   * the first kRes entries {column | integer bit, coefficient} of the row stay in the warp's slice of shared memory,
//     transposed (entry t of the row of lane l at [t][l]: conflict-free); the 148 SMs together hold the matrix.
// A round then takes one of two forms, picked per warp and round:
//   * lane = ROW (dense rounds, eval_resident): pass 1 gathers {lb,ub} four at a time per lane and adds the products
//     a*blo (rounded down) / a*bhi (rounded up) IN ASCENDING COLUMN ORDER [getLfBnds_ :1237-1258], keeping
//     term_reach() of every entry as a float rounded up; the singleton-infinity sums [getSingLfBnds_ :1261-1319] are a
//     second, rare pass; pass 2 is the product test from registers.  The entries that pass it are few and spread
//     unevenly over the rows, and the exact path [updateLfBoundsFromLb_/Ub_ :1048-1226] is long (two directed fp64
//     divisions), so the (row lane, entry) pairs of the whole warp are compacted into a work list and taken
//     ENTRY-PARALLEL, 32 per trip;
//   * lane = ENTRY (sparse rounds: at most 32 due entries in the warp, eval_packed): one gather per lane, the row's own
//     lane adds the products in column order through shuffles, and an entry derives its exact candidate from the
//     bounds it already holds.  A sparse round is bound by the instructions on this path, not by bandwidth.
// In both, the CSC list of an entry's variable is looked up speculatively with the gather, so a moved variable's rows
// are flagged [changeBFlag_ :1229-1234] one trip later (flag_lists: the lists laid end to end, entry-parallel).
// Entries beyond kRes of a row are read from the CSR in global memory by the same lane; rows longer than kLaneMax are
// evaluated by the whole warp (eval_long in row_batch.cuh).
#pragma once
#include "row_batch.cuh"

namespace mntr {

constexpr int kRes = 12;        // resident entries per row
constexpr int kPassGroup = 4;    // gathers a lane has in flight in pass 1
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
  static constexpr bool kSlab = true;
  static constexpr bool kL2Hints = false;
  __device__ __forceinline__ int *list() { return items; }
  __device__ __forceinline__ uint16_t *work_list() { return work; }
  static constexpr int kQueueCap = 0;                                   // (resident rows take their candidates at once)
  __device__ __forceinline__ CandItem *queue() { return nullptr; }
};
static_assert(sizeof(ResidentStage) % 16 == 0, "slices are laid out back to back");

// phase 0: the lane's row head is in registers (load_head); its first kRes entries go into the slice
__device__ __forceinline__ void resident_fill(const LinDev &P, ResidentStage &S, int lane, const RowHead h)
{
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
}

// entry t of the lane's row: from the slice when the stage keeps the rows resident, else from the CSR
template <class Stage>
__device__ __forceinline__ void resident_entry(const LinDev &P, const Stage &S, int lane, int beg, int t,
                                               double &a, int &cx)
{
  if constexpr (Stage::kSlab) {
    if (t < kRes) { a = S.val[t][lane]; cx = S.colx[t][lane]; return; }
  }
  a = __ldg(P.val + beg + t); cx = __ldg(P.colx + beg + t);
}

// singleton-infinity sums of the lane's row [getSingLfBnds_ :1261-1319]: finite sum + infinity count per side.  Rare
// (only rows with an infinite activity), serial, out of line.
template <class R, class Stage>
__device__ __noinline__ void singleton_sums(const LinDev &P, const ReadPending &rd, const Stage &S, int lane,
                                            const RowHead h, double &sing_ll, double &sing_uu)
{
  double fs_lo = 0.0, fs_hi = 0.0;
  int ninf_lo = 0, ninf_hi = 0;
  for (int t = 0; t < h.cnt; ++t) {
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

// One trip of the exact path, lane = entry: bounds of the entry's variable and (speculatively, riding along with the
// gather) the position of its CSC list; exact candidates; a moved variable is marked and its rows are flagged for the
// next round [changeBFlag_ :1229-1234].  Convergent.
template <class R, class Sink, class Stage>
__device__ __forceinline__ void exact_and_flag(const LinDev &P, const ReadPending &rd, const Sink &sink, Stage &S, int lane,
                                               bool on, double slb, double sub, int sg, double a, int cx)
{
  int qb = 0, len = 0;
  if (on) {
    int j; bool isint;
    const double2 bt = rd.fetch(cx, j, isint);
    int qe = 0;
    if constexpr (Sink::kFlagsRows) { qb = __ldg(P.csc_ptr + j); qe = __ldg(P.csc_ptr + j + 1); }
    const double2 c = exact_candidates<R>(slb, sub, sg, a, bt.x, bt.y);
    const bool up = c.x > bt.x, down = c.y < bt.y;
    if (up) sink.raise_lb(S, j, isint, c.x);
    if (down) sink.lower_ub(S, j, isint, c.y);
    if (up || down) { sink.touch(j, isint); len = qe - qb; }
  }
  if constexpr (Sink::kFlagsRows) {
    if (__any_sync(kFullMask, len != 0)) {
      if (lane == 0) sink.changed();
      flag_lists(qb, len, lane, P.csc_row, sink.due_next());
    }
  }
}

// queue of deferred exact candidates (streaming form)
template <class R, class Sink, class Stage>
__device__ __noinline__ void drain_queue(const LinDev &P, const ReadPending &rd, const Sink &sink, Stage &S, int lane)
{
  __syncwarp();
  const int n = S.pad_[1];
  for (int base = 0; base < n; base += 32) {
    const bool on = base + lane < n;
    CandItem it{0.0, 0.0, 0.0, 0, 0};
    if (on) it = S.queue()[base + lane];
    exact_and_flag<R>(P, rd, sink, S, lane, on, it.slb, it.sub, it.sg, it.a, it.cx);
  }
  __syncwarp();
  if (lane == 0) S.pad_[1] = 0;
  __syncwarp();
}

// One round of the rows of this warp: lane = row.  `due`: the lane's row is evaluated in this round (h.cnt >= 0).
// Both passes issue ALL the gathers of the resident entries before the first one is consumed: a round costs a row one
// trip to L2 for its activities and, if any of its entries can move a bound, one more for the exact candidates.
// With a stage that does not keep rows resident (Stage::kSlab false: the streaming form of large instances) the lane
// reads its row's entries from the CSR, four at a time with 128-bit loads (rows start at and are padded to a multiple
// of four entries); the lines are shared by neighbouring lanes and consecutive groups through L1.
// G: gathers a lane keeps in flight in pass 1 (a multiple of four; kRes: the whole resident part at once)
template <class R, int G = kPassGroup, class Sink, class Stage>
__device__ __forceinline__ void eval_resident(const LinDev &P, const ReadPending &rd, const Sink &sink, Stage &S,
                                              int lane, bool due, const RowHead h, bool first,
                                              const int32_t *gcol = nullptr, const double *gval = nullptr,
                                              int next_row0 = -1, int next_e0 = 0)
{
  // gcol / gval (streaming form): where the lane's row starts -- in the CSR, or in the block staged in shared memory;
  // next_row0 >= 0: the warp's next block starts at that row and at entry next_e0 -- its heads and entries are
  // requested (cp.async) as soon as pass 1 has consumed this block's staged entries
  if (first && due && h.rl > h.ru + kETol) sink.row_bounds_cross();     // checkBounds_, rows part (:328-359)
  const bool mine = due && h.cnt <= kLaneMax;
  unsigned longm = __ballot_sync(kFullMask, due && h.cnt > kLaneMax);
  const int cnt = mine ? h.cnt : 0;
  const int maxc = __reduce_max_sync(kFullMask, cnt);
  sink.mark(S, lane); sink.phase(lane, 0);

  // ---- pass 1: activities, terms in ascending column order ----
  [[maybe_unused]] unsigned long long keep = 0ull;
  if constexpr (!Stage::kSlab) { if constexpr (Stage::kL2Hints) keep = l2_policy_evict_last(); }
  double ll = 0.0, uu = 0.0;
  unsigned need = 0u;
  double slb = INFINITY, sub = INFINITY;
  uint8_t sg = 0;
  {
    float reach[kRes];
    static_assert(G % 4 == 0 && kRes % G == 0, "pass 1 runs in groups of G entries");
#pragma unroll
    for (int g = 0; g < kRes / G; ++g) {
      if (g * G >= maxc) {            // warp-uniform: no row of the warp reaches this group
#pragma unroll
        for (int u = 0; u < G; ++u) reach[g * G + u] = 0.f;
        continue;
      }
      double2 b[G];
      double av[G];
      int cv[G];
      if constexpr (Stage::kSlab) {
#pragma unroll
        for (int u = 0; u < G; ++u) {
          const int t = g * G + u;
          av[u] = 0.0; cv[u] = 0;
          if (t < cnt) { av[u] = S.val[t][lane]; cv[u] = S.colx[t][lane]; }
        }
      } else {
        // one 128-bit column load and two 128-bit value loads per four entries
#pragma unroll
        for (int q = 0; q < G; q += 4) {
          int4 c = make_int4(0, 0, 0, 0);
          double2 v0 = make_double2(0.0, 0.0), v1 = v0;
          if (g * G + q < cnt) {
            c = *reinterpret_cast<const int4 *>(gcol + g * G + q);
            v0 = *reinterpret_cast<const double2 *>(gval + g * G + q);
            v1 = *reinterpret_cast<const double2 *>(gval + g * G + q + 2);
          }
          cv[q] = c.x; cv[q + 1] = c.y; cv[q + 2] = c.z; cv[q + 3] = c.w;
          av[q] = v0.x; av[q + 1] = v0.y; av[q + 2] = v1.x; av[q + 3] = v1.y;
        }
      }
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int t = g * G + u;
        b[u] = make_double2(0.0, 0.0);
        if (t < cnt) {
          bool hinted = false;
          if constexpr (!Stage::kSlab) {
            if constexpr (Stage::kL2Hints) { b[u] = ld_box_keep(rd.box + (cv[u] & kColMask), keep); hinted = true; }   // the box stays in L2
          }
          if (!hinted) b[u] = __ldcg(rd.box + (cv[u] & kColMask));
        }
      }
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int t = g * G + u;
        reach[t] = 0.f;
        if (t < cnt) {
          const double a = av[u];
          if (rd.round_ints && cv[u] < 0) tighten_int_bounds(b[u].x, b[u].y);
          const bool pos = a > 0.0;
          ll = R::add_lo(ll, R::mul_lo(a, pos ? b[u].x : b[u].y));
          uu = R::add_hi(uu, R::mul_hi(a, pos ? b[u].y : b[u].x));
          // term_reach rounded UP to a float: conservative, inf falls through
          reach[t] = __double2float_ru(term_reach(a, b[u]));
        }
      }
    }
    if constexpr (!Stage::kSlab) {
      if (next_row0 >= 0 && longm == 0u) stage_next_block(P, S, lane, next_row0, next_e0);     // (eval_long would reuse plo[])
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
      if (ll < -kInf20 || uu > kInf20) singleton_sums<R, Stage>(P, rd, S, lane, h, sing_ll, sing_uu);     // rare
      if (ll > h.ru + kETol || uu < h.rl - kETol) sink.row_infeasible();     // :994-1015
      else row_offers<R>(h.rl, h.ru, ll, uu, sing_ll, sing_uu, slb, sub, sg);
    }

    // ---- pass 2: product test ----
#pragma unroll
    for (int t = 0; t < kRes; ++t) {
      if (t < cnt) {
        const double rch = (double)reach[t];
        if (!(slb >= rch) || !(sub >= rch)) need |= 1u << t;
      }
    }
  }
  if (slb < INFINITY || sub < INFINITY) {
    for (int t = kRes; t < cnt; ++t) {
      int j; bool isint;
      const double a = __ldg(P.val + h.beg + t);
      const double2 bt = rd.get(h.beg + t, j, isint);
      const double rch = term_reach(a, bt);
      if (!(slb >= rch) || !(sub >= rch)) need |= 1u << t;
    }
  }
  sink.mark(S, lane); sink.phase(lane, 2);

  // ---- exact candidates [updateLfBoundsFromLb_/Ub_] of the entries that passed the test.  They are few and spread
  //      unevenly over the rows, and the exact path is long (two directed fp64 divisions): the entries of the whole
  //      warp are compacted and taken ENTRY-PARALLEL, 32 per trip, so the exact path runs with full lanes.
  //      Resident rows: a work list of (row lane, entry) pairs, taken at once.  Streamed rows: the items go on a
  //      queue that is drained when it is full or the range is done, so the chain of dependent trips (bounds and
  //      CSC pointers, division, CSC rows) is paid once per ~96 candidates instead of once per 32-row block ----
  {
    const int mine_n = __popc(need);
    int incl = mine_n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(kFullMask, incl, d);
      if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(kFullMask, incl, 31);
    bool queued = false;
    if constexpr (!Stage::kSlab) {
      if (total > 0 && total <= Stage::kQueueCap) {
        __syncwarp();
        if (S.pad_[1] + total > Stage::kQueueCap) drain_queue<R>(P, rd, sink, S, lane);
        int pos = S.pad_[1] + incl - mine_n;
        for (unsigned nd = need; nd; nd &= nd - 1) {
          double a; int cx;
          resident_entry(P, S, lane, h.beg, __ffs(nd) - 1, a, cx);
          S.queue()[pos++] = CandItem{slb, sub, a, cx, (int)sg};
        }
        __syncwarp();
        if (lane == 0) S.pad_[1] += total;
        __syncwarp();
        queued = true;
      } else if (total > 0) {
        // more candidates than the queue holds (rare): each lane takes its own, one per warp-uniform trip
        drain_queue<R>(P, rd, sink, S, lane);
        while (__any_sync(kFullMask, need != 0u)) {
          double a = 0.0; int cx = 0;
          const bool on = need != 0u;
          if (on) { resident_entry(P, S, lane, h.beg, __ffs(need) - 1, a, cx); need &= need - 1; }
          exact_and_flag<R>(P, rd, sink, S, lane, on, slb, sub, (int)sg, a, cx);
        }
        queued = true;
      }
    }
    if constexpr (Stage::kSlab) if (total > 0 && !queued) {
      int pos = incl - mine_n;
      for (unsigned nd = need; nd; nd &= nd - 1) S.work_list()[pos++] = (uint16_t)(lane | ((__ffs(nd) - 1) << 5));
      __syncwarp();
      for (int base = 0; base < total; base += 32) {
        const bool on = base + lane < total;
        const int item = on ? S.work_list()[base + lane] : 0;
        const int rl = item & 31, t = item >> 5;
        const double o_slb = __shfl_sync(kFullMask, slb, rl), o_sub = __shfl_sync(kFullMask, sub, rl);
        const int o_sg = __shfl_sync(kFullMask, (int)sg, rl), o_beg = __shfl_sync(kFullMask, h.beg, rl);
        double a = 0.0; int cx = 0;
        if (on) resident_entry(P, S, rl, o_beg, t, a, cx);
        exact_and_flag<R>(P, rd, sink, S, lane, on, o_slb, o_sub, o_sg, a, cx);
      }
      __syncwarp();
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

// SPARSE rounds: the due rows of the warp have at most 32 entries in all, every one of them resident.  The entries are
// laid end to end over the lanes (lane = ENTRY: one gather per lane, each entry's arithmetic done once instead of once
// per unrolled slot), the row's own lane adds the products in ascending column order through shuffles, and an entry
// that can move a bound derives its exact candidate from the bounds it already holds: one trip to L2 for the whole
// evaluation, one more for the rows to flag.  A sparse round is bound by the instructions on this path, not by
// bandwidth -- it is a third of the lane = row form's.
template <class R, class Sink, class Stage = ResidentStage>
__device__ __forceinline__ void eval_packed(const LinDev &P, const ReadPending &rd, const Sink &sink, ResidentStage &S,
                                            int lane, bool due, const RowHead h, bool first, int incl, int total)
{
  if (first && due && h.rl > h.ru + kETol) sink.row_bounds_cross();     // checkBounds_, rows part (:328-359)
  sink.mark(S, lane); sink.phase(lane, 0);
  const int cnt = due ? h.cnt : 0;
  const int off = incl - cnt;
  // ---- lane = entry ----
  int rl = 0;                                      // the row lane of entry `lane`: first lane whose inclusive count exceeds it
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int v = __shfl_sync(kFullMask, incl, rl + step - 1);
    if (v <= lane) rl += step;
  }
  const int roff = __shfl_sync(kFullMask, off, rl);
  const bool on = lane < total;
  double a = 0.0, plo = 0.0, phi = 0.0, rch = 0.0;
  double2 b = make_double2(0.0, 0.0);
  int j = 0, qb = 0, qe = 0;
  bool isint = false;
  if (on) {
    const int t = lane - roff;
    a = S.val[t][rl];
    b = rd.fetch(S.colx[t][rl], j, isint);
    qb = __ldg(P.csc_ptr + j);                     // speculative: the variable's row list, should it move
    qe = __ldg(P.csc_ptr + j + 1);
    const bool pos = a > 0.0;
    plo = R::mul_lo(a, pos ? b.x : b.y);
    phi = R::mul_hi(a, pos ? b.y : b.x);
    rch = term_reach(a, b);
  }
  // ---- lane = row: products added in ascending column order ----
  double ll = 0.0, uu = 0.0;
  const int maxc = __reduce_max_sync(kFullMask, cnt);
  for (int t = 0; t < maxc; ++t) {
    const double vlo = __shfl_sync(kFullMask, plo, off + t), vhi = __shfl_sync(kFullMask, phi, off + t);
    if (t < cnt) { ll = R::add_lo(ll, vlo); uu = R::add_hi(uu, vhi); }
  }
  sink.mark(S, lane); sink.phase(lane, 1);
  double slb = INFINITY, sub = INFINITY;
  int sg = 0;
  if (due) {
    double sing_ll = -INFINITY, sing_uu = INFINITY;
    if (ll < -kInf20 || uu > kInf20) singleton_sums<R, Stage>(P, rd, S, lane, h, sing_ll, sing_uu);     // rare
    uint8_t sg8 = 0;
    if (ll > h.ru + kETol || uu < h.rl - kETol) sink.row_infeasible();     // :994-1015
    else row_offers<R>(h.rl, h.ru, ll, uu, sing_ll, sing_uu, slb, sub, sg8);
    sg = sg8;
  }
  // ---- lane = entry: product test, rarely the exact candidate [updateLfBoundsFromLb_/Ub_ :1048-1226] ----
  const double o_slb = __shfl_sync(kFullMask, slb, rl), o_sub = __shfl_sync(kFullMask, sub, rl);
  const int o_sg = __shfl_sync(kFullMask, sg, rl);
  const bool need = on && (!(o_slb >= rch) || !(o_sub >= rch));
  sink.mark(S, lane); sink.phase(lane, 2);
  int len = 0;
  if (__any_sync(kFullMask, need)) {
    if (need) {
      const double2 c = exact_candidates<R>(o_slb, o_sub, o_sg, a, b.x, b.y);
      const bool up = c.x > b.x, down = c.y < b.y;
      if (up) sink.raise_lb(S, j, isint, c.x);
      if (down) sink.lower_ub(S, j, isint, c.y);
      if (up || down) { sink.touch(j, isint); len = qe - qb; }
    }
    sink.mark(S, lane); sink.phase(lane, 3);
    if (__any_sync(kFullMask, len != 0)) {
      if (lane == 0) sink.changed();
      flag_lists(qb, len, lane, P.csc_row, sink.due_next());
    }
  } else {
    sink.mark(S, lane); sink.phase(lane, 3);
  }
  sink.mark(S, lane); sink.phase(lane, 4);
}

// Streaming form: the rows of a 32-row block lie back to back in the CSR.  When most of them are due and the block is
// not too long, its entries are copied to shared memory in ONE coalesced, asynchronous burst (cp.async, 16 bytes per
// lane and instruction, no registers held), so pass 1 starts from shared memory with one trip to DRAM behind it instead
// of one per group of four entries.  Returns where the lane's row starts (staged or in the CSR).
// (Measured and dropped: pulling the NEXT block's entries into L2 with prefetch.global.L2, and holding the next
// block's row heads in registers -- both made the 2.5M-row fixpoint slower, 631 -> 720 us.)
template <class Stage>
__device__ __forceinline__ void stage_block(const LinDev &P, Stage &S, int lane, bool is_due, const RowHead h,
                                            const int32_t *&gcol, const double *&gval)
{
  gcol = P.colx + h.beg; gval = P.val + h.beg;
  const int e1 = __reduce_max_sync(kFullMask, is_due ? row_end(make_int2(h.beg, h.cnt)) : 0);
  const int e0 = __reduce_min_sync(kFullMask, is_due ? h.beg : 0x7fffffff);
  const int extent = e1 - e0;
  if (extent <= 0 || extent > kStageEntries || __popc(__ballot_sync(kFullMask, is_due)) < 16) return;
  int32_t *sc = S.stage_col();
  double *sv = S.stage_val();
  __syncwarp();
  if constexpr (Stage::kL2Hints) {
    const unsigned long long pol = l2_policy_evict_first();
    for (int o = lane * 4; o < extent; o += 128) {
      const unsigned dc = (unsigned)__cvta_generic_to_shared(sc + o), dv = (unsigned)__cvta_generic_to_shared(sv + o);
      cp_async16_stream(dc, P.colx + e0 + o, pol);
      cp_async16_stream(dv, P.val + e0 + o, pol);
      cp_async16_stream(dv + 16, P.val + e0 + o + 2, pol);
    }
  } else {
    for (int o = lane * 4; o < extent; o += 128) {
      const unsigned dc = (unsigned)__cvta_generic_to_shared(sc + o), dv = (unsigned)__cvta_generic_to_shared(sv + o);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dc), "l"(P.colx + e0 + o) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dv), "l"(P.val + e0 + o) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dv + 16), "l"(P.val + e0 + o + 2) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  gcol = sc + (h.beg - e0); gval = sv + (h.beg - e0);
}

// ---- streaming form, staging AHEAD ----
// The rows of a warp's range lie back to back in the CSR, so the next 32-row block's entries start where this block's
// end -- known before the next heads are.  Right after pass 1 (the staged entries of the current block are consumed)
// the next block's row heads AND a window of kStageEntries entries from that position are requested together with
// cp.async: by the time the current block's product test and candidates are done they have landed, and the next block
// starts with shared-memory reads instead of two dependent trips to DRAM (heads, then entries).
template <class Stage>
__device__ __forceinline__ void stage_next_block(const LinDev &P, Stage &S, int lane, int row0, int e0)
{
  __syncwarp();
  {
    const unsigned di = (unsigned)__cvta_generic_to_shared(S.stage_info() + lane);
    const unsigned db = (unsigned)__cvta_generic_to_shared(S.stage_bnd() + lane);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(di), "l"(P.row_info + row0 + lane) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(db), "l"(P.row_bnd + row0 + lane) : "memory");
  }
  const int extent = min(kStageEntries, P.nnz_pad - e0);      // (a multiple of four: rows are padded)
  int32_t *sc = S.stage_col();
  double *sv = S.stage_val();
  const unsigned long long pol = l2_policy_evict_first();
  for (int o = lane * 4; o < extent; o += 128) {
    const unsigned dc = (unsigned)__cvta_generic_to_shared(sc + o), dv = (unsigned)__cvta_generic_to_shared(sv + o);
    cp_async16_stream(dc, P.colx + e0 + o, pol);
    cp_async16_stream(dv, P.val + e0 + o, pol);
    cp_async16_stream(dv + 16, P.val + e0 + o + 2, pol);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// STREAMING form (rows not resident: large instances).  The due rows of [r0, r1): 32-row blocks (one word of the bit
// set) in which at least kDenseRows rows are due are evaluated lane = row straight from the CSR (eval_resident with a
// non-resident stage): in the dense rounds, which carry almost all the work of a large instance, that costs about two
// warp instructions per nonzero where the staged, entry-parallel batches of row_batch.cuh cost about twelve.  The
// bits taken are cleared; what remains (sparse blocks) goes to the packing path of eval_due_range.
constexpr int kDenseRows = 6;
template <class R, class Sink>
__device__ __forceinline__ void eval_due_stream(const LinDev &P, const ReadPending &rd, const Sink &sink, WarpStage &S,
                                                uint32_t *due, int r0, int r1, bool first, int lane,
                                                unsigned long long &my_nnz, unsigned long long &my_rows)
{
  if (r0 >= r1) return;
  if (r1 - r0 > 32) {
    bool rest = false;
    for (int wb = r0 >> 5; wb * 32 < r1; wb += 32) {
      const int w = wb + lane;
      unsigned word = 0u;
      if (w * 32 < r1) {
        word = first ? kFullMask : __ldcg(due + w);
        if (w * 32 < r0) word &= ~0u << (r0 - w * 32);
        if (r1 - w * 32 < 32) word &= (1u << (r1 - w * 32)) - 1u;
      }
      const bool dense = first ? word != 0u : __popc(word) >= kDenseRows;
      if (dense && !first) atomicAnd(due + w, ~word);            // setBFlag(false), :513
      if (word != 0u && !dense) rest = true;
      unsigned dm = __ballot_sync(kFullMask, dense);
      int spec_row0 = -1, spec_e0 = 0;         // the block staged ahead (first row, first entry of the window), if any
      while (dm) {
        const int k = __ffs(dm) - 1;
        dm &= dm - 1;
        const unsigned wk = __shfl_sync(kFullMask, word, k);
        const int row0 = (wb + k) * 32, row = row0 + lane;
        const bool bit = (wk >> lane) & 1u;
        // heads of ALL the block's rows (the end of its last row is where the next block's entries start)
        RowHead h{0, -1, 0.0, 0.0};
        const bool ahead = spec_row0 == row0;
        if (ahead) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          const int2 info = S.stage_info()[lane];
          const double2 bnd = S.stage_bnd()[lane];
          h = RowHead{info.x, info.y, bnd.x, bnd.y};
        } else if (row < P.m) {
          h = load_head(P, row);
        }
        const bool is_due = bit && h.cnt >= 0;                     // deleted rows (term count < 0) are never evaluated
        if (is_due) { my_nnz += (unsigned long long)h.cnt; ++my_rows; }
        // where the lane's row starts: in the window staged ahead if it covers the block's due rows, else staged now
        // (one burst) or read from the CSR
        const int32_t *gcol; const double *gval;
        const int d1 = __reduce_max_sync(kFullMask, is_due ? row_end(make_int2(h.beg, h.cnt)) : 0);
        const int d0 = __reduce_min_sync(kFullMask, is_due ? h.beg : 0x7fffffff);
        if (ahead && d0 >= spec_e0 && d1 <= spec_e0 + kStageEntries) {
          gcol = S.stage_col() + (h.beg - spec_e0); gval = S.stage_val() + (h.beg - spec_e0);
        } else {
          stage_block(P, S, lane, is_due, h, gcol, gval);
        }
        // the next dense block is the adjacent one and this block has no deleted rows: stage it ahead
        const bool whole = __all_sync(kFullMask, row < P.m && h.cnt >= 0);
        const int blk_end = __shfl_sync(kFullMask, row_end(make_int2(h.beg, h.cnt < 0 ? 0 : h.cnt)), 31);
#ifdef MNTR_K1_NOSPEC
        const bool adj = false;
#else
        const bool adj = dm != 0u && __ffs(dm) - 1 == k + 1 && whole && row0 + 64 <= P.m && blk_end < P.nnz_pad;
#endif
        spec_row0 = adj ? row0 + 32 : -1;
        spec_e0 = blk_end;
        eval_resident<R>(P, rd, sink, S, lane, is_due, h, first, gcol, gval, spec_row0, spec_e0);
        if (spec_row0 >= 0 && __ballot_sync(kFullMask, is_due && h.cnt > kLaneMax) != 0u) spec_row0 = -1;   // (not issued)
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    drain_queue<R>(P, rd, sink, S, lane);
    if (first || !__any_sync(kFullMask, rest)) return;
  }
  eval_due_range<R>(P, rd, sink, S, due, r0, r1, first, lane, my_nnz, my_rows);
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
  // few entries in all, every one resident: the entry-parallel form (sparse rounds); else lane = row
  const int c = is_due ? h.cnt : 0;
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  if (total <= 32 && !__any_sync(kFullMask, c > kRes)) eval_packed<R>(P, rd, sink, S, lane, is_due, h, first, incl, total);
  else eval_resident<R>(P, rd, sink, S, lane, is_due, h, first);
}

}  // namespace mntr
