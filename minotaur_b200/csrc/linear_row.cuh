// linear_row.cuh -- evaluation of ONE linear row against ONE box by a sub-warp group of G lanes
// (Jacobi form of LinearHandler::linBndTighten_, LinearHandler.cpp:952-1045), registers and shuffles only.
// Used by the per-round kernels of the row-partitioned multi-GPU path (linear_rounds.cu) for the due rows of SPARSE
// 32-row blocks and the cut-off row; dense blocks are evaluated lane = row (row_resident.cuh, streaming form).
#pragma once
#include "device_problem.cuh"

namespace mntr {

// candidates merged into separate lb / ub arrays: each is contiguous, so the cross-GPU merge is one
// NCCL MAX all-reduce on nlb and one MIN all-reduce on nub.  Slot n carries the row-infeasible flag.
struct SinkSplit {
  double *nlb, *nub;
  int n;
  __device__ __forceinline__ void raise_lb(int j, double c) const { atomic_max_f64(&nlb[j], c); }
  __device__ __forceinline__ void lower_ub(int j, double c) const { atomic_min_f64(&nub[j], c); }
  __device__ __forceinline__ void row_infeasible(int32_t *) const { nlb[n] = 2.0; }
};

// butterfly over the G lanes of a group; every lane ends with bitwise the same total
template <int G, class R, bool LO>
__device__ __forceinline__ double group_reduce(double v, unsigned mask)
{
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    double o = __shfl_xor_sync(mask, v, off, G);
    v = LO ? R::add_lo(v, o) : R::add_hi(v, o);
  }
  return v;
}

template <int G>
__device__ __forceinline__ int group_reduce_int(int v, unsigned mask)
{
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
  return v;
}

// What a row offers its terms once its activities are known.  slack_* = +inf means "this side yields
// nothing", so the per-term product test below rejects without looking at the side at all.
struct RowCtx {
  double slack_lb;     // max activity - row lb   (>= -eTol)        [updateLfBoundsFromLb_]
  double slack_ub;     // row ub - min activity   (>= -eTol)        [updateLfBoundsFromUb_]
  bool sing_lb, sing_ub;   // singleton-infinity mode of that side
};

// branch-free term of getLfBnds_: an absent entry has a == 0 and b == {0,0}, contributing +0
template <class R>
__device__ __forceinline__ void accumulate(double a, double2 b, double &ll, double &uu)
{
  const bool pos = a > 0.0;
  const double blo = pos ? b.x : b.y, bhi = pos ? b.y : b.x;
  ll = R::add_lo(ll, R::mul_lo(a, blo));
  uu = R::add_hi(uu, R::mul_hi(a, bhi));
}

// Exact candidates of one term (rare path, taken only when the product test cannot rule them out).
//
// Both row sides and both coefficient signs obey one rule: the candidate moves ONE bound of x_j
// towards the other by slack/|a|:
//     a>0, lb side: new lb = ub_j - slack/a        a<0, lb side: new ub = lb_j + slack/|a|
//     a>0, ub side: new ub = lb_j + slack/a        a<0, ub side: new lb = ub_j - slack/|a|
// With directed rounding the division is taken on |a| (round_up(x/a) == -round_down(x/|a|) for
// a<0; in round-to-nearest x/a == -(x/|a|) exactly), so the sign only picks the bound that moves.
// numer = rl - act (lb side, rounded down) / ru - act (ub side, rounded up) as the reference forms it.
template <class R, class Sink>
__device__ __forceinline__ void emit_exact(const RowCtx rc, double a, int j, double2 b, const Sink sink)
{
  const double vl = b.x, vu = b.y;
  const double aa = fabs(a);
  if (!(aa > kETol)) return;
  const bool pos = a > 0.0;
  if (rc.slack_lb < INFINITY) {
    const double numer = -rc.slack_lb;
    const bool inf_side = pos ? (vu >= kInf20) : (vl <= -kInf20);
    if (!rc.sing_lb || inf_side) {
      const double base = inf_side ? 0.0 : (pos ? vu : vl);
      const double t = R::div_lo(numer, aa);        // round_down((rl - act)/|a|)
      if (pos) {
        double c = R::add_lo(t, base);
        if (c > vl + kETol) { if (c > vu - kETol) c = vu; sink.raise_lb(j, c); }
      } else {
        double c = R::add_hi(-t, base);
        if (c < vu - kETol) { if (c < vl + kETol) c = vl; sink.lower_ub(j, c); }
      }
    }
  }
  if (rc.slack_ub < INFINITY) {
    const double numer = rc.slack_ub;
    const bool inf_side = pos ? (vl <= -kInf20) : (vu >= kInf20);
    if (!rc.sing_ub || inf_side) {
      const double base = inf_side ? 0.0 : (pos ? vl : vu);
      const double sq = R::div_hi(numer, aa);       // round_up((ru - act)/|a|)
      if (pos) {
        double c = R::add_hi(sq, base);
        if (c < vu - kETol) { if (c < vl + kETol) c = vl; sink.lower_ub(j, c); }
      } else {
        double c = R::add_lo(-sq, base);
        if (c > vl + kETol) { if (c > vu - kETol) c = vu; sink.raise_lb(j, c); }
      }
    }
  }
}

// Product test: a candidate can only be accepted when slack < |a| * (ub_j - lb_j).  One DADD and two
// DMULs shared by both sides reject almost every term of a round without the fp64 division; the test
// is conservative (1e-9 relative margin; an infinite or NaN reach falls through to the exact path), so
// results are unchanged.
template <class R, class Sink>
__device__ __forceinline__ void emit_candidates(const RowCtx &rc, double a, int j, double2 b, const Sink &sink)
{
  const double reach = term_reach(a, b);
  if (!(rc.slack_lb >= reach) || !(rc.slack_ub >= reach)) emit_exact<R, Sink>(rc, a, j, b, sink);
}

// this lane's first four entries of a row (a == 0: none)
struct RowData { double2 a01, a23; int4 c; };

__device__ __forceinline__ RowData load_data(const LinDev &P, int beg, int cnt, int lane_g)
{
  RowData d;
  d.a01 = make_double2(0.0, 0.0); d.a23 = d.a01; d.c = make_int4(0, 0, 0, 0);
  const int t = 4 * lane_g;
  if (t < cnt) {       // rows are padded to a multiple of 4 entries: the three 128-bit loads stay inside the row
    d.a01 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + beg + t));
    d.a23 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + beg + t + 2));
    d.c = __ldg(reinterpret_cast<const int4 *>(P.col + beg + t));
  }
  return d;
}

// singleton-infinity sums (getSingLfBnds_, LinearHandler.cpp:1261-1319) as finite-sum + infinity-count
// per side; rare path
template <int G, class R>
__device__ __noinline__ void sing_activity(const LinDev &P, const double2 *box, int beg, int end, int lane_g,
                                           unsigned gmask, double &sing_ll, double &sing_uu)
{
  double fs_lo = 0.0, fs_hi = 0.0;
  int ninf_lo = 0, ninf_hi = 0;
  for (int t = beg + lane_g; t < end; t += G) {
    {
      const double a = __ldg(P.val + t);
      const int j = __ldg(P.col + t);
      if (a > kETol) {
        const double2 b = box[j];
        if (b.y < kInf20) fs_hi = R::add_hi(fs_hi, R::mul_hi(a, b.y)); else ++ninf_hi;
        if (b.x > -kInf20) fs_lo = R::add_lo(fs_lo, R::mul_lo(a, b.x)); else ++ninf_lo;
      } else if (a < -kETol) {
        const double2 b = box[j];
        if (b.y < kInf20) fs_lo = R::add_lo(fs_lo, R::mul_lo(a, b.y)); else ++ninf_lo;
        if (b.x > -kInf20) fs_hi = R::add_hi(fs_hi, R::mul_hi(a, b.x)); else ++ninf_hi;
      }
    }
  }
  fs_lo = group_reduce<G, R, true>(fs_lo, gmask);
  fs_hi = group_reduce<G, R, false>(fs_hi, gmask);
  ninf_lo = group_reduce_int<G>(ninf_lo, gmask);
  ninf_hi = group_reduce_int<G>(ninf_hi, gmask);
  sing_ll = (ninf_lo >= 2) ? -INFINITY : fs_lo;
  sing_uu = (ninf_hi >= 2) ? INFINITY : fs_hi;
}

// one flagged row: activities, infeasibility, candidates   [linBndTighten_, Jacobi form]
// Row i has `cnt` terms starting at entry `beg`; each lane owns four consecutive entries (two 128-bit
// value loads + one 128-bit column load), so a group of G lanes covers 4G entries per step and most rows
// take a single step.  The four {lb,ub} gathers of a lane are issued together.
template <int G, class R, class Sink>
__device__ __forceinline__ void process_row(const LinDev &P, const double2 *box, int32_t *status,
                                            const Sink &sink, int i, int beg, int cnt, int lane_g, unsigned gmask,
                                            unsigned long long &my_nnz, unsigned long long &my_rows)
{
  const RowData d = load_data(P, beg, cnt, lane_g);
  const double2 bnd = __ldg(P.row_bnd + i);
  const double rl = bnd.x, ru = bnd.y;
  if (lane_g == 0) { my_nnz += (unsigned long long)cnt; ++my_rows; }
  // pass 1: an absent entry (a == 0) reads a zero box so that it contributes exactly +0
  const double2 zero = make_double2(0.0, 0.0);
  const double2 b0 = (d.a01.x != 0.0) ? box[d.c.x] : zero;
  const double2 b1 = (d.a01.y != 0.0) ? box[d.c.y] : zero;
  const double2 b2 = (d.a23.x != 0.0) ? box[d.c.z] : zero;
  const double2 b3 = (d.a23.y != 0.0) ? box[d.c.w] : zero;
  double ll = 0.0, uu = 0.0;
  accumulate<R>(d.a01.x, b0, ll, uu);
  accumulate<R>(d.a01.y, b1, ll, uu);
  accumulate<R>(d.a23.x, b2, ll, uu);
  accumulate<R>(d.a23.y, b3, ll, uu);
  const bool long_row = cnt > 4 * G;                        // group-uniform
  const int end = row_end(make_int2(beg, cnt));
  if (long_row) {
    for (int t = beg + 4 * G + lane_g; t < end; t += G) {
      const double a = __ldg(P.val + t);
      if (a != 0.0) accumulate<R>(a, box[__ldg(P.col + t)], ll, uu);
    }
  }
  ll = group_reduce<G, R, true>(ll, gmask);
  uu = group_reduce<G, R, false>(uu, gmask);

  // singleton-infinity sums, only when an activity is beyond +-1e20 (:970-972)
  double sing_ll = -INFINITY, sing_uu = INFINITY;
  if (ll < -kInf20 || uu > kInf20) sing_activity<G, R>(P, box, beg, end, lane_g, gmask, sing_ll, sing_uu);

  if (ll > ru + kETol || uu < rl - kETol) {     // activity-infeasible row (:994-1015)
    if (lane_g == 0) sink.row_infeasible(status);
    return;
  }
  RowCtx rc;
  rc.slack_lb = INFINITY; rc.slack_ub = INFINITY;
  rc.sing_lb = false; rc.sing_ub = false;
  if (rl > -kInf20) {                            // :1017-1025
    if (uu < kInf20) rc.slack_lb = -R::sub_lo(rl, uu);
    else if (sing_uu < kInf20) { rc.slack_lb = -R::sub_lo(rl, sing_uu); rc.sing_lb = true; }
  }
  if (ru < kInf20) {                             // :1035-1043
    if (ll > -kInf20) rc.slack_ub = R::sub_hi(ru, ll);
    else if (sing_ll > -kInf20) { rc.slack_ub = R::sub_hi(ru, sing_ll); rc.sing_ub = true; }
  }
  // pass 2: implied bounds; the lane's entries and their bounds are still in registers
  emit_candidates<R>(rc, d.a01.x, d.c.x, b0, sink);
  emit_candidates<R>(rc, d.a01.y, d.c.y, b1, sink);
  emit_candidates<R>(rc, d.a23.x, d.c.z, b2, sink);
  emit_candidates<R>(rc, d.a23.y, d.c.w, b3, sink);
  if (long_row) {
    for (int t = beg + 4 * G + lane_g; t < end; t += G) {
      const double a = __ldg(P.val + t);
      const int j = __ldg(P.col + t);
      if (a != 0.0) emit_candidates<R>(rc, a, j, box[j], sink);
    }
  }
}

}  // namespace mntr
