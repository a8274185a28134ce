// row_batch.cuh -- ENTRY-PARALLEL evaluation of up to 32 linear rows by one warp against ONE box
// (Jacobi form of LinearHandler::linBndTighten_, LinearHandler.cpp:952-1045).
//
// A warp takes a batch of due rows, one row per lane, and stages the batch in its shared-memory slice:
//   pass A  (lane = ENTRY): the entries of all rows of the batch are laid end to end; lane l takes entries
//           l, l+32, ... -- coalesced (col,val) loads and up to ten independent 128-bit {lb,ub} gathers per lane in
//           flight -- and stores the two activity products a*blo (rounded down) / a*bhi (rounded up) of its
//           entries [getLfBnds_ :1237-1258] plus the infinity bits of the singleton rule [getSingLfBnds_ :1261-1319]
//   pass B  (lane = ROW): the row's lane adds its products IN ASCENDING COLUMN ORDER (the reference's order),
//           tests activity infeasibility [:994-1015] and publishes what the row offers its terms: the two slacks
//   pass C  (lane = ENTRY): product test slack < |a|(ub-lb) per entry (|a|(ub-lb) kept in a register, as a float
//           rounded up); only entries that can move a bound re-derive the exact candidate [updateLfBoundsFromLb_/
//           Ub_ :1048-1226] and hand it to the Sink.
// Rows longer than the staging capacity are evaluated by the whole warp, strided, with a butterfly reduction.
//
// This is the LATENCY form, used by the single-launch fixpoint kernel (linear_single.cu): one staged batch per warp
// per round, every load of the batch in flight at once.  The per-round kernels of the row-partitioned multi-GPU path
// stream hundreds of rows per warp and use the lane = row streaming form (row_resident.cuh).  How the box is read
// (Reader) and where candidates go (Sink) are policies.
#pragma once
#include "device_problem.cuh"

namespace mntr {

constexpr int kCap = 384;                 // entries a warp stages per batch
constexpr int kPerLane = kCap / 32;       // entries per lane per batch
constexpr int kGroup = 4;                 // gathers a lane issues back to back
static_assert(kPerLane % kGroup == 0, "pass A runs in groups");
constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kColMask = 0x7fffffff;

constexpr uint8_t kLoInf = 1, kHiInf = 2, kTiny = 4;   // per-entry bits of the singleton-infinity rule

// a deferred exact candidate of the streaming form (row_resident.cuh): what its row offers, the entry
struct __align__(16) CandItem { double slb, sub, a; int cx, sg; };
constexpr int kStageEntries = 384;      // entries of a 32-row block staged in shared memory (streaming form)

// one warp's slice of shared memory
struct __align__(16) WarpStage {
  double plo[kCap];        // a * blo, rounded down   (after pass B: first-touch list of the Sink)
  double phi[kCap];        // a * bhi, rounded up
  double slack_lb[32];     // max activity - row lb  (>= -eTol) or +inf: nothing to derive   [updateLfBoundsFromLb_]
  double slack_ub[32];     // row ub - min activity  (>= -eTol) or +inf                       [updateLfBoundsFromUb_]
  int beg[32];             // first CSR entry of the row in slot s
  int off[32];             // first staged entry of the row in slot s
  int tcount;              // length of the first-touch list
  int pad_[3];             // [0] debug mark counter  [1] length of the candidate queue
  uint8_t slot[kCap];      // staged entry -> row slot (= lane of the row)
  uint8_t flag[kCap];      // kLoInf | kHiInf | kTiny
  uint8_t sing[32];        // bit 0: lb side runs in singleton-infinity mode, bit 1: ub side
  // the Sink's list of moved variables lives in plo[] once pass B has consumed the products
  static constexpr int kListCap = 2 * kCap;
  __device__ __forceinline__ int *list() { return reinterpret_cast<int *>(plo); }
  // The lane = row form of dense 32-row blocks (row_resident.cuh, streaming form) reuses the slice: the block's
  // entries staged by cp.async (values in plo[], columns in the first half of phi[]), the queue of deferred exact
  // candidates in the second half of phi[]; the work list of a block with more candidates than the queue holds
  // overwrites the staged values (they have been consumed by then).
  static constexpr bool kSlab = false;
  static constexpr bool kL2Hints = true;     // matrix stream evict-first, box gathers evict-last (common.cuh)
  static constexpr int kQueueCap = 48;
  __device__ __forceinline__ double *stage_val() { return plo; }
  __device__ __forceinline__ int32_t *stage_col() { return reinterpret_cast<int32_t *>(phi); }
  __device__ __forceinline__ CandItem *queue() { return reinterpret_cast<CandItem *>(phi + kCap / 2); }
  __device__ __forceinline__ uint16_t *work_list() { return reinterpret_cast<uint16_t *>(plo); }
  // row heads of the NEXT 32-row block, staged ahead together with its entries (slot[] / flag[] are only used by the
  // staged batches, which run after the dense blocks)
  __device__ __forceinline__ int2 *stage_info() { return reinterpret_cast<int2 *>(slot); }
  __device__ __forceinline__ double2 *stage_bnd() { return reinterpret_cast<double2 *>(slot + 256); }
};
static_assert(2 * kCap >= 32 * (sizeof(int2) + sizeof(double2)), "staged row heads fit slot[] + flag[]");
static_assert(kCap >= kStageEntries && sizeof(double) * (kCap / 2) >= sizeof(int32_t) * kStageEntries, "staged block fits");
static_assert(sizeof(double) * (kCap / 2) >= sizeof(CandItem) * WarpStage::kQueueCap, "the candidate queue fits phi[]");
static_assert(sizeof(double) * kCap >= sizeof(uint16_t) * 32 * 32, "the work list of a 32-row block fits plo[]");

// ---------------------------------------------------------------------------------------------------------------
// Readers: how an entry's column and the bounds of its variable are fetched.  Boxes are read with ld.global.cg
// (L2): they are updated by other SMs between rounds, and a gather has no reuse inside an SM.
// ---------------------------------------------------------------------------------------------------------------

// merged-but-unrounded box of the one-barrier fixpoint kernel: bit 31 of the column marks an integer variable, whose
// bounds are rounded on the fly [tightenInts_ :415-490] from round 2 on (round 1 sees the incoming box as is)
struct ReadPending {
  const double2 *box;
  const int32_t *colx;
  bool round_ints;
  // bounds of the variable of a packed column (bit 31: integer)
  __device__ __forceinline__ double2 fetch(int cx, int &j, bool &isint) const
  {
    j = cx & kColMask;
    isint = cx < 0;
    double2 b = __ldcg(box + j);
    if (round_ints && isint) tighten_int_bounds(b.x, b.y);
    return b;
  }
  __device__ __forceinline__ double2 get(int src, int &j, bool &isint) const { return fetch(__ldg(colx + src), j, isint); }
};

// ---------------------------------------------------------------------------------------------------------------
// Sinks: where accepted candidates go.
// ---------------------------------------------------------------------------------------------------------------

// The rows of up to 32 moved variables, one per lane (len == 0: none), are flagged for the next round [changeBFlag_
// :1229-1234]: their CSC lists [qb, qb + len) are laid end to end and taken ENTRY-PARALLEL, one trip to L2 per 32
// rows to flag.  Convergent.
__device__ __forceinline__ void flag_lists(int qb, int len, int lane, const int32_t *csc_row, uint32_t *due_next)
{
  int incl = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  for (int e0 = 0; e0 < total; e0 += 32) {
    const int x = e0 + lane;
    int lo = 0;                                      // first lane whose inclusive count exceeds x
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(kFullMask, incl, lo + step - 1);
      if (v <= x) lo += step;
    }
    const int oqb = __shfl_sync(kFullMask, qb, lo), oincl = __shfl_sync(kFullMask, incl, lo);
    const int olen = __shfl_sync(kFullMask, len, lo);
    if (x < total) {
      const int row = __ldg(csc_row + oqb + (x - (oincl - olen)));
      atomicOr(due_next + (row >> 5), 1u << (row & 31));
    }
  }
}

struct FixRound;
template <class Stage> static __device__ __noinline__ void flush_moved(Stage &S, int lane, const FixRound *rc, unsigned round);

// One-barrier fixpoint kernel: candidates are merged into the NEXT round's box; a moved variable is marked in the
// round's bit set (the next round's fix-up scans it) and its rows are flagged for the next round [changeBFlag_
// :1229-1234] by a warp-cooperative walk of the CSC lists in flush().  Everything is fire-and-forget: no atomic
// returns a value, so nothing waits for a round trip.  (Two rows moving the same variable in one round walk its
// list twice; the flags are idempotent.)
// The round's pointers live in shared memory (one copy per block): they are needed on rare paths only and must
// not occupy registers during the row evaluation.
struct FixRound {
  unsigned seen_changed, seen_int;   // last round for which THIS BLOCK has already reported a move / an integer move
  double2 *next_box;
  uint32_t *touched;       // bit set: variable moved in this round
  uint32_t *due_next;      // row bit set of the next round   (Constraint bFlag)
  unsigned *sync;          // SingleWs::sync
  const int32_t *csc_ptr, *csc_row;
};

struct SinkFix {
  static constexpr bool kFlagsRows = true;      // a moved variable's rows are flagged by the emitting warp
  __device__ __forceinline__ uint32_t *due_next() const { return rc->due_next; }
  const FixRound *rc;          // in shared memory (one record per round parity, built once)
  unsigned round;
  unsigned long long *probe;   // debug (MNTR_GPU_TRACE): time stamps of one warp's passes, or nullptr
  unsigned long long *blk;     // debug: per-block, per-round latest time any warp passed phase k, or nullptr
  __device__ __forceinline__ void phase(int lane, int k) const
  {
    if (blk != nullptr && lane == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMax(blk + k, t);
    }
  }

  template <class Stage> __device__ __forceinline__ void mark(Stage &S, int lane) const
  {
    if (probe != nullptr && lane == 0 && S.pad_[0] < 32) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      probe[S.pad_[0]++] = t;
    }
  }
  template <class Stage> __device__ __forceinline__ void raise_lb(Stage &, int j, bool, double c) const { atomic_max_f64(&rc->next_box[j].x, c); }
  template <class Stage> __device__ __forceinline__ void lower_ub(Stage &, int j, bool, double c) const { atomic_min_f64(&rc->next_box[j].y, c); }
  // the moved variable goes on the warp's list; flush() walks the lists of 32 variables at a time
  template <class Stage> __device__ __forceinline__ void moved(Stage &S, int j, bool isint) const
  {
    S.list()[atomicAdd(&S.tcount, 1)] = j;
    if (isint) int_moved();      // a row moved an integer variable (nintmods, :1070-1133)
  }
  // without the list: mark a moved variable now; its rows are flagged by flag_lists()
  // "a bound moved in this round" / "a row moved an integer variable in this round" are device-wide words that every
  // block reads after the barrier; thousands of warps would report the same thing to the same address, and
  // same-address atomics are served one after the other: a block reports each once per round
  __device__ __forceinline__ void int_moved() const
  {
    FixRound *w = const_cast<FixRound *>(rc);
    if (w->seen_int != round && atomicExch(&w->seen_int, round) != round) atomicMax(rc->sync + 1, 2u * round + 1u);
  }
  __device__ __forceinline__ void changed() const
  {
    FixRound *w = const_cast<FixRound *>(rc);
    if (w->seen_changed != round && atomicExch(&w->seen_changed, round) != round) atomicMax(rc->sync + 1, 2u * round);
  }
  __device__ __forceinline__ void touch(int j, bool isint) const
  {
    atomicOr(rc->touched + (j >> 5), 1u << (j & 31));
    if (isint) int_moved();      // nintmods, :1070-1133
  }
  // round-tagged (see SingleWs::sync): a block that is already in round r+1 can never make a block that is still
  // taking its snapshot of barrier r believe that round r found an infeasible row
  __device__ __forceinline__ void row_infeasible() const { atomicMax(rc->sync + 3, round); }
  __device__ __forceinline__ void row_bounds_cross() const { atomicOr(rc->sync + 2, kCtlRowCross); }
  // no room for `room` more entries?
  template <class Stage> __device__ __forceinline__ bool near_full(const Stage &S, int room = 32) const { return S.tcount + room > Stage::kListCap; }

  // convergent: mark the warp's moved variables and flag their rows
  template <class Stage> __device__ __forceinline__ void flush(Stage &S, int lane) const
  {
    __syncwarp();
    if (S.tcount != 0) flush_moved(S, lane, rc, round);
  }
};

// out of line: rare, and called from several places
template <class Stage>
static __device__ void flush_moved(Stage &S, int lane, const FixRound *rc, unsigned round)
{
  uint32_t *due_next = rc->due_next, *touched = rc->touched;
  const int32_t *csc_ptr = rc->csc_ptr, *csc_row = rc->csc_row;
  const int n = S.tcount;
  const int *tl = S.list();
  if (lane == 0) SinkFix{rc, round, nullptr, nullptr}.changed();
  for (int base = 0; base < n; base += 32) {
    const int idx = base + lane;
    int qb = 0, len = 0;
    if (idx < n) {
      const int j = tl[idx];
      qb = __ldg(csc_ptr + j); len = __ldg(csc_ptr + j + 1) - qb;
      atomicOr(touched + (j >> 5), 1u << (j & 31));
    }
    flag_lists(qb, len, lane, csc_row, due_next);
  }
  __syncwarp();
  if (lane == 0) S.tcount = 0;
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Exact candidates of one term (rare path, taken only when the product test cannot rule them out).
//
// Both row sides and both coefficient signs obey one rule: the candidate moves ONE bound of x_j towards the other by
// slack/|a|:
//     a>0, lb side: new lb = ub_j - slack/a        a<0, lb side: new ub = lb_j + slack/|a|
//     a>0, ub side: new ub = lb_j + slack/a        a<0, ub side: new lb = ub_j - slack/|a|
// With directed rounding the division is taken on |a| (round_up(x/a) == -round_down(x/|a|) for a<0; in
// round-to-nearest x/a == -(x/|a|) exactly), so the sign only picks the bound that moves.
// A candidate is handed on only when it really moves the bound (after the clamp to the opposite bound it may not).
// ---------------------------------------------------------------------------------------------------------------
// Returns {new lb or -inf, new ub or +inf} (pure arithmetic, out of line: the fp64 divisions are long).
template <class R>
__device__ __noinline__ double2 exact_candidates(double slack_lb, double slack_ub, int sing, double a, double vl, double vu)
{
  double2 out = make_double2(-INFINITY, INFINITY);
  const double aa = fabs(a);
  if (!(aa > kETol)) return out;
  const bool pos = a > 0.0;
  if (slack_lb < INFINITY) {
    const bool inf_side = pos ? (vu >= kInf20) : (vl <= -kInf20);
    if (!(sing & 1) || inf_side) {
      const double base = inf_side ? 0.0 : (pos ? vu : vl);
      const double t = R::div_lo(-slack_lb, aa);        // round_down((rl - act)/|a|)
      if (pos) {
        double c = R::add_lo(t, base);
        if (c > vl + kETol) { if (c > vu - kETol) c = vu; out.x = c; }
      } else {
        double c = R::add_hi(-t, base);
        if (c < vu - kETol) { if (c < vl + kETol) c = vl; out.y = c; }
      }
    }
  }
  if (slack_ub < INFINITY) {
    const bool inf_side = pos ? (vl <= -kInf20) : (vu >= kInf20);
    if (!(sing & 2) || inf_side) {
      const double base = inf_side ? 0.0 : (pos ? vl : vu);
      const double sq = R::div_hi(slack_ub, aa);        // round_up((ru - act)/|a|)
      if (pos) {
        double c = R::add_hi(sq, base);
        if (c < vu - kETol) { if (c < vl + kETol) c = vl; out.y = c; }
      } else {
        double c = R::add_lo(-sq, base);
        if (c > vl + kETol) { if (c > vu - kETol) c = vu; out.x = c; }
      }
    }
  }
  return out;
}

// A candidate is handed to the sink only when it really moves the bound (after the clamp it may not).
template <class R, class Sink, class Stage>
__device__ __forceinline__ void emit_exact(double slack_lb, double slack_ub, int sing, double a, int j, bool isint,
                                           double2 b, Stage &S, const Sink &sink)
{
  const double2 c = exact_candidates<R>(slack_lb, slack_ub, sing, a, b.x, b.y);
  const bool up = c.x > b.x, down = c.y < b.y;
  if (up) sink.raise_lb(S, j, isint, c.x);
  if (down) sink.lower_ub(S, j, isint, c.y);
  if (up || down) sink.moved(S, j, isint);
}

// what a row offers its terms once its activities are known (:1017-1043); +inf slack = this side yields nothing
template <class R>
__device__ __forceinline__ void row_offers(double rl, double ru, double ll, double uu, double sing_ll, double sing_uu,
                                           double &slack_lb, double &slack_ub, uint8_t &sing)
{
  slack_lb = INFINITY; slack_ub = INFINITY; sing = 0;
  if (rl > -kInf20) {
    if (uu < kInf20) slack_lb = -R::sub_lo(rl, uu);
    else if (sing_uu < kInf20) { slack_lb = -R::sub_lo(rl, sing_uu); sing |= 1; }
  }
  if (ru < kInf20) {
    if (ll > -kInf20) slack_ub = R::sub_hi(ru, ll);
    else if (sing_ll > -kInf20) { slack_ub = R::sub_hi(ru, sing_ll); sing |= 2; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One staged batch: the rows of the lanes with `in` set; their entries occupy staged positions [loff, loff+cnt),
// T entries in all (T <= kCap).
// ---------------------------------------------------------------------------------------------------------------
template <class R, class Reader, class Sink>
__device__ __forceinline__ void eval_staged(const LinDev &P, const Reader &rd, const Sink &sink, WarpStage &S, int lane,
                                            bool in, int beg, int cnt, int loff, int T, double rl, double ru)
{
  if (in) {
    S.beg[lane] = beg; S.off[lane] = loff;
    for (int t = 0; t < cnt; ++t) S.slot[loff + t] = (uint8_t)lane;
  }
  __syncwarp();
  sink.mark(S, lane);

  // ---- pass A: lane = entry ----
  float reach[kPerLane];
#pragma unroll
  for (int h = 0; h < kPerLane / kGroup; ++h) {
    if (h * kGroup * 32 >= T) {             // warp-uniform: nothing staged beyond here
#pragma unroll
      for (int u = 0; u < kGroup; ++u) reach[h * kGroup + u] = 0.f;
      continue;
    }
    double a[kGroup];
    double2 b[kGroup];
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int e = (h * kGroup + u) * 32 + lane;
      a[u] = 0.0; b[u] = make_double2(0.0, 0.0);
      if (e < T) {
        const int s = S.slot[e];
        const int src = S.beg[s] + (e - S.off[s]);
        int j; bool isint;
        a[u] = __ldg(P.val + src);
        b[u] = rd.get(src, j, isint);
      }
    }
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int k = h * kGroup + u;
      const int e = k * 32 + lane;
      reach[k] = 0.f;
      if (e < T) {
        const bool pos = a[u] > 0.0;
        const double blo = pos ? b[u].x : b[u].y, bhi = pos ? b[u].y : b[u].x;
        S.plo[e] = R::mul_lo(a[u], blo);
        S.phi[e] = R::mul_hi(a[u], bhi);
        uint8_t f = 0;
        if (pos ? (b[u].x <= -kInf20) : (b[u].y >= kInf20)) f |= kLoInf;
        if (pos ? (b[u].y >= kInf20) : (b[u].x <= -kInf20)) f |= kHiInf;
        if (!(fabs(a[u]) > kETol)) f |= kTiny;
        S.flag[e] = f;
        // term_reach rounded UP to a float: conservative, inf falls through
        reach[k] = __double2float_ru(term_reach(a[u], b[u]));
      }
    }
  }
  __syncwarp();
  sink.mark(S, lane);

  // ---- pass B: lane = row; products added in ascending column order ----
  if (in) {
    double ll = 0.0, uu = 0.0;
    for (int t = 0; t < cnt; ++t) {
      ll = R::add_lo(ll, S.plo[loff + t]);
      uu = R::add_hi(uu, S.phi[loff + t]);
    }
    double sing_ll = -INFINITY, sing_uu = INFINITY;
    if (ll < -kInf20 || uu > kInf20) {            // singleton-infinity sums: finite sum + infinity count per side
      double fs_lo = 0.0, fs_hi = 0.0;
      int ninf_lo = 0, ninf_hi = 0;
      for (int t = 0; t < cnt; ++t) {
        const uint8_t f = S.flag[loff + t];
        if (f & kTiny) continue;
        if (f & kLoInf) ++ninf_lo; else fs_lo = R::add_lo(fs_lo, S.plo[loff + t]);
        if (f & kHiInf) ++ninf_hi; else fs_hi = R::add_hi(fs_hi, S.phi[loff + t]);
      }
      sing_ll = (ninf_lo >= 2) ? -INFINITY : fs_lo;
      sing_uu = (ninf_hi >= 2) ? INFINITY : fs_hi;
    }
    double slb = INFINITY, sub = INFINITY;
    uint8_t sg = 0;
    if (ll > ru + kETol || uu < rl - kETol) sink.row_infeasible();     // :994-1015
    else row_offers<R>(rl, ru, ll, uu, sing_ll, sing_uu, slb, sub, sg);
    S.slack_lb[lane] = slb; S.slack_ub[lane] = sub; S.sing[lane] = sg;
  }
  __syncwarp();
  sink.mark(S, lane);

  // ---- pass C: lane = entry; product test, rarely the exact candidate ----
  unsigned need = 0u;
#pragma unroll
  for (int k = 0; k < kPerLane; ++k) {
    const int e = k * 32 + lane;
    if (e < T) {
      const int s = S.slot[e];
      const double rch = (double)reach[k];
      if (!(S.slack_lb[s] >= rch) || !(S.slack_ub[s] >= rch)) need |= 1u << k;
    }
  }
  while (need) {                          // one copy of the exact path, whatever the entry
    const int k = __ffs(need) - 1;
    need &= need - 1;
    const int e = k * 32 + lane;
    const int s = S.slot[e];
    const int src = S.beg[s] + (e - S.off[s]);
    int j; bool isint;
    const double a = __ldg(P.val + src);
    const double2 b = rd.get(src, j, isint);
    emit_exact<R>(S.slack_lb[s], S.slack_ub[s], S.sing[s], a, j, isint, b, S, sink);
  }
  sink.mark(S, lane);
  sink.flush(S, lane);
  sink.mark(S, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// One row longer than the staging capacity, by the whole warp (all arguments warp-uniform).
// ---------------------------------------------------------------------------------------------------------------
template <class R, bool LO>
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(kFullMask, v, off);
    v = LO ? R::add_lo(v, o) : R::add_hi(v, o);
  }
  return v;
}

template <class R, class Reader, class Sink, class Stage>
__device__ __forceinline__ void eval_long(const double *val, const Reader &rd, const Sink &sink, Stage &S, int lane,
                                          int beg, int cnt, double rl, double ru)
{
  double ll = 0.0, uu = 0.0, fs_lo = 0.0, fs_hi = 0.0;
  int ninf_lo = 0, ninf_hi = 0;
  for (int t = lane; t < cnt; t += 32) {
    int j; bool isint;
    const double a = __ldg(val + beg + t);
    const double2 b = rd.get(beg + t, j, isint);
    const bool pos = a > 0.0;
    const double plo = R::mul_lo(a, pos ? b.x : b.y), phi = R::mul_hi(a, pos ? b.y : b.x);
    ll = R::add_lo(ll, plo);
    uu = R::add_hi(uu, phi);
    if (fabs(a) > kETol) {
      if (pos ? (b.x <= -kInf20) : (b.y >= kInf20)) ++ninf_lo; else fs_lo = R::add_lo(fs_lo, plo);
      if (pos ? (b.y >= kInf20) : (b.x <= -kInf20)) ++ninf_hi; else fs_hi = R::add_hi(fs_hi, phi);
    }
  }
  ll = warp_sum<R, true>(ll); uu = warp_sum<R, false>(uu);
  fs_lo = warp_sum<R, true>(fs_lo); fs_hi = warp_sum<R, false>(fs_hi);
  ninf_lo = __reduce_add_sync(kFullMask, ninf_lo); ninf_hi = __reduce_add_sync(kFullMask, ninf_hi);
  if (ll > ru + kETol || uu < rl - kETol) {
    if (lane == 0) sink.row_infeasible();
    return;
  }
  double sing_ll = -INFINITY, sing_uu = INFINITY;
  if (ll < -kInf20 || uu > kInf20) {
    sing_ll = (ninf_lo >= 2) ? -INFINITY : fs_lo;
    sing_uu = (ninf_hi >= 2) ? INFINITY : fs_hi;
  }
  double slb, sub;
  uint8_t sg;
  row_offers<R>(rl, ru, ll, uu, sing_ll, sing_uu, slb, sub, sg);
  for (int t0 = 0; t0 < cnt; t0 += 32) {
    __syncwarp();
    if (sink.near_full(S)) sink.flush(S, lane);          // warp-uniform: tcount is read after the barrier
    const int t = t0 + lane;
    if (t < cnt) {
      int j; bool isint;
      const double a = __ldg(val + beg + t);
      const double2 b = rd.get(beg + t, j, isint);
      const double rch = term_reach(a, b);
      if (!(slb >= rch) || !(sub >= rch)) emit_exact<R>(slb, sub, sg, a, j, isint, b, S, sink);
    }
  }
  sink.flush(S, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// Up to 32 rows, one per lane (row < 0: none).  Deleted rows (term count < 0) are never evaluated.
// ---------------------------------------------------------------------------------------------------------------
// what a lane knows about its row
struct RowHead { int beg, cnt; double rl, ru; };     // cnt < 0: no row / deleted row

__device__ __forceinline__ RowHead load_head(const LinDev &P, int row)
{
  RowHead h{0, -1, 0.0, 0.0};
  if (row >= 0) {                                   // both loads in flight together
    const int2 info = __ldg(P.row_info + row);
    const double2 bnd = __ldg(P.row_bnd + row);
    h.beg = info.x; h.cnt = info.y; h.rl = bnd.x; h.ru = bnd.y;
  }
  return h;
}

template <class R, class Reader, class Sink>
__device__ __forceinline__ void eval_rows(const LinDev &P, const Reader &rd, const Sink &sink, WarpStage &S, int lane,
                                          const RowHead h, bool first)
{
  const bool valid = h.cnt >= 0;
  const int beg = h.beg, cnt = valid ? h.cnt : 0;
  const double rl = h.rl, ru = h.ru;
  if (first && valid && rl > ru + kETol) sink.row_bounds_cross();     // checkBounds_, rows part (:328-359)
  sink.mark(S, lane);
  const bool staged = valid && cnt <= kCap;
  unsigned longm = __ballot_sync(kFullMask, valid && cnt > kCap);
  unsigned pending = __ballot_sync(kFullMask, staged);
  const int bc = staged ? cnt : 0;
  int incl = bc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, d);
    if (lane >= d) incl += v;
  }
  const int off = incl - bc;
  // batches that exceed the capacity are split at row boundaries (the staged lanes form a prefix of `pending`)
  while (pending) {
    const int f = __ffs(pending) - 1;
    const int lo = __shfl_sync(kFullMask, off, f);
    const bool in = ((pending >> lane) & 1u) && (incl - lo <= kCap);
    const unsigned inm = __ballot_sync(kFullMask, in);
    const int last = 31 - __clz(inm);
    const int T = __shfl_sync(kFullMask, incl, last) - lo;
    eval_staged<R>(P, rd, sink, S, lane, in, beg, cnt, off - lo, T, rl, ru);
    pending &= ~inm;
  }
  while (longm) {
    const int s = __ffs(longm) - 1;
    longm &= longm - 1;
    const int b = __shfl_sync(kFullMask, beg, s), c = __shfl_sync(kFullMask, cnt, s);
    const double l = __shfl_sync(kFullMask, rl, s), u = __shfl_sync(kFullMask, ru, s);
    eval_long<R>(P.val, rd, sink, S, lane, b, c, l, u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The due rows of [r0, r1): which rows are due is a BIT SET (one bit per row, the reference's Constraint bFlag).
// The warp reads up to 32 words at once, clears what it takes with fire-and-forget atomics [setBFlag(false), :513]
// and packs the due rows into batches of 32, so sparse rounds keep every lane of pass B busy.
// ---------------------------------------------------------------------------------------------------------------
template <class R, class Reader, class Sink>
__device__ __forceinline__ void eval_due_range(const LinDev &P, const Reader &rd, const Sink &sink, WarpStage &S,
                                               uint32_t *due, int r0, int r1, bool first, int lane,
                                               unsigned long long &my_nnz, unsigned long long &my_rows)
{
  if (r0 >= r1) return;
  // One loop, one call site of the (large) row evaluation.  Short ranges -- the usual case when the instance fits
  // the GPU in one pass -- take a single trip: the row heads are fetched speculatively, together with the due
  // bits, and handed to the lanes the due rows are packed into.  Long ranges read 32 words of the bit set per
  // refill and pack the due rows of those 1024 rows into batches of 32.
  const bool short_range = r1 - r0 <= 32;
  bool done = false;
  int wb = r0 >> 5, wcur = 0, base = 0, total = 0, incl = 0;
  unsigned word = 0u;
  for (;;) {
    RowHead h;
    if (short_range) {
      if (done) break;
      done = true;
      const RowHead sp = load_head(P, r0 + lane < r1 ? r0 + lane : -1);
      unsigned m = r1 - r0 == 32 ? kFullMask : ((1u << (r1 - r0)) - 1u);     // bit l = row r0 + l
      if (!first) {                                    // in round 1 every row is due (:1618-1622)
        const int w0 = r0 >> 5, sh = r0 & 31;
        unsigned lo = 0u, hi = 0u;
        if (lane == 0) lo = __ldcg(due + w0);          // bits are set by L2 atomics: bypass L1
        if (lane == 1 && sh != 0 && (w0 + 1) * 32 < r1) hi = __ldcg(due + w0 + 1);
        lo = __shfl_sync(kFullMask, lo, 0); hi = __shfl_sync(kFullMask, hi, 1);
        m &= sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        if (m != 0u) {                                 // setBFlag(false), :513
          if (lane == 0 && (m << sh) != 0u) atomicAnd(due + w0, ~(m << sh));
          if (lane == 1 && sh != 0 && (m >> (32 - sh)) != 0u) atomicAnd(due + w0 + 1, ~(m >> (32 - sh)));
        }
      }
      sink.mark(S, lane);
      if (m == 0u) break;
      const int nd = __popc(m);
      const int src = lane < nd ? (int)__fns(m, 0, lane + 1) : 0;      // lane x takes the x-th due row
      h.beg = __shfl_sync(kFullMask, sp.beg, src); h.cnt = __shfl_sync(kFullMask, sp.cnt, src);
      h.rl = __shfl_sync(kFullMask, sp.rl, src); h.ru = __shfl_sync(kFullMask, sp.ru, src);
      if (lane >= nd) h.cnt = -1;
    } else {
      if (base >= total) {                             // refill: the next 32 words
        if (wb * 32 >= r1) break;
        const int w = wb + lane;
        word = 0u;
        if (w * 32 < r1) {
          word = first ? kFullMask : __ldcg(due + w);
          if (w * 32 < r0) word &= ~0u << (r0 - w * 32);
          if (r1 - w * 32 < 32) word &= (1u << (r1 - w * 32)) - 1u;
          if (word && !first) atomicAnd(due + w, ~word);
        }
        incl = __popc(word);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(kFullMask, incl, d);
          if (lane >= d) incl += v;
        }
        total = __shfl_sync(kFullMask, incl, 31);
        wcur = wb; wb += 32; base = 0;
        if (total == 0) continue;
      }
      const int x = base + lane;
      base += 32;
      int lo = 0;                                      // first lane whose inclusive count exceeds x
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(kFullMask, incl, lo + step - 1);
        if (v <= x) lo += step;
      }
      const unsigned wword = __shfl_sync(kFullMask, word, lo);
      const int wincl = __shfl_sync(kFullMask, incl, lo);
      int row = -1;
      if (x < total) row = (wcur + lo) * 32 + (int)__fns(wword, 0, x - (wincl - __popc(wword)) + 1);
      h = load_head(P, row);
    }
    if (h.cnt >= 0) { my_nnz += (unsigned long long)h.cnt; ++my_rows; }
    eval_rows<R>(P, rd, sink, S, lane, h, first);
  }
}

}  // namespace mntr
