// linear_row.cuh -- evaluation of ONE linear row against ONE box by a sub-warp group of G lanes
// (Jacobi form of LinearHandler::linBndTighten_, LinearHandler.cpp:952-1045).  Shared by the
// cooperative single-launch fixpoint kernel (linear_single.cu) and the per-round kernels of the
// row-partitioned multi-GPU path (linear_rounds.cu); they differ only in where candidate bounds go
// (the Sink).
#pragma once
#include "device_problem.cuh"

namespace mntr {

// candidates merged into an interleaved {lb,ub} box
struct SinkBox {
  double2 *nbox;
  __device__ __forceinline__ void raise_lb(int j, double c) const { atomic_max_f64(&nbox[j].x, c); }
  __device__ __forceinline__ void lower_ub(int j, double c) const { atomic_min_f64(&nbox[j].y, c); }
  __device__ __forceinline__ void row_infeasible(int32_t *status) const { status[0] = 2; /* MNTR_INFEAS_ROW */ }
};

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

struct RowCtx {
  double rl, ru;       // row bounds
  double act_lb;       // activity used by FromLb (uu or sing_uu)
  double act_ub;       // activity used by FromUb (ll or sing_ll)
  bool do_lb, do_ub, sing_lb, sing_ub;
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

// one term of pass 2: candidates of updateLfBoundsFromLb_ / updateLfBoundsFromUb_.
//
// Both row sides and both coefficient signs obey one rule: the candidate moves ONE bound of x_j
// towards the other by  slack/|a|,  slack = (max activity - row lb)  or  (row ub - min activity):
//     a>0, lb side: new lb = ub_j - slack/a        a<0, lb side: new ub = lb_j + slack/|a|
//     a>0, ub side: new ub = lb_j + slack/a        a<0, ub side: new lb = ub_j - slack/|a|
// so it can only be accepted when slack < |a| * (ub_j - lb_j).  That product test (one DADD, one
// DMUL, shared by both sides) rejects almost every term of a round without the fp64 division; it
// is conservative (1e-9 relative margin, NaN/inf fall through), so results are unchanged.
// With directed rounding the division is taken on |a| (round_up(x/a) == -round_down(x/|a|) for
// a<0; in round-to-nearest x/a == -(x/|a|) exactly), so the sign only picks the bound that moves.
template <class R, class Sink>
__device__ __forceinline__ void emit_candidates(const RowCtx &rc, double a, int j, double2 b,
                                                const Sink &sink)
{
  const double vl = b.x, vu = b.y;
  const double aa = fabs(a);
  if (!(aa > kETol)) return;
  const bool pos = a > 0.0;
  const double reach = aa * (vu - vl) * 1.000000001;      // inf or NaN when a bound is infinite
  if (rc.do_lb) {                                   // row lb side: numer = rl - max activity = -slack
    const double numer = R::sub_lo(rc.rl, rc.act_lb);
    const bool inf_side = pos ? (vu >= kInf20) : (vl <= -kInf20);
    if ((!rc.sing_lb || inf_side) && !(-numer > reach)) {
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
  if (rc.do_ub) {                                   // row ub side: numer = ru - min activity = slack
    const double numer = R::sub_hi(rc.ru, rc.act_ub);
    const bool inf_side = pos ? (vl <= -kInf20) : (vu >= kInf20);
    if ((!rc.sing_ub || inf_side) && !(numer > reach)) {
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

// pipeline registers of one row
struct RowMeta { int i, beg, cnt; };                  // i < 0: nothing to do; cnt = true term count
struct RowData { double2 a2; int2 c2; };              // this lane's first entry pair (a == 0: none)

__device__ __forceinline__ RowMeta load_meta(const LinDev &P, const int32_t *list, int idx, int count, bool first)
{
  RowMeta r; r.i = -1; r.beg = 0; r.cnt = 0;
  if (idx < count) {
    const int i = first ? idx : list[idx];
    const int2 info = __ldg(P.row_info + i);
    if (info.y >= 0) { r.i = i; r.beg = info.x; r.cnt = info.y; }   // deleted rows are never evaluated
  }
  return r;
}

__device__ __forceinline__ RowData load_data(const LinDev &P, const RowMeta &r, int lane_g)
{
  RowData d; d.a2 = make_double2(0.0, 0.0); d.c2 = make_int2(0, 0);
  const int t = 2 * lane_g;
  if (r.i >= 0 && t < r.cnt) {
    d.a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + r.beg + t));
    d.c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + r.beg + t));
  }
  return d;
}

// one flagged row: activities, infeasibility, candidates   [linBndTighten_, Jacobi form]
template <int G, class R, class Sink>
__device__ __forceinline__ void process_row(const LinDev &P, const double2 *box, uint32_t *bits, int32_t *status,
                                            const Sink &sink, const RowMeta &r, const RowData &d,
                                            int lane_g, unsigned gmask, bool first, unsigned long long &my_nnz,
                                            unsigned long long &my_rows)
{
  const int i = r.i;
  const int end = r.beg + ((r.cnt + 1) & ~1);
  RowCtx rc;
  const double2 bnd = __ldg(P.row_bnd + i);
  rc.rl = bnd.x;
  rc.ru = bnd.y;
  if (lane_g == 0) {
    my_nnz += (unsigned long long)r.cnt; ++my_rows;
    if (!first) atomicAnd(bits + (i >> 5), ~(1u << (i & 31)));       // setBFlag(false), :513
  }
  // pass 1: the first entry pair is already in registers, its two gathers go out together
  double2 b0 = make_double2(0.0, 0.0), b1 = b0;
  if (d.a2.x != 0.0) b0 = box[d.c2.x];
  if (d.a2.y != 0.0) b1 = box[d.c2.y];
  double ll = 0.0, uu = 0.0;
  accumulate<R>(d.a2.x, b0, ll, uu);
  accumulate<R>(d.a2.y, b1, ll, uu);
  const bool long_row = r.cnt > 2 * G;                      // group-uniform
  if (long_row) {
    for (int t = r.beg + 2 * lane_g + 2 * G; t < end; t += 2 * G) {
      const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
      const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
      if (a2.x != 0.0) accumulate<R>(a2.x, box[c2.x], ll, uu);
      if (a2.y != 0.0) accumulate<R>(a2.y, box[c2.y], ll, uu);
    }
  }
  ll = group_reduce<G, R, true>(ll, gmask);
  uu = group_reduce<G, R, false>(uu, gmask);

  // singleton-infinity sums, only when an activity is beyond +-1e20 (:970-972)
  double sing_ll = -INFINITY, sing_uu = INFINITY;
  if (ll < -kInf20 || uu > kInf20) {
    double fs_lo = 0.0, fs_hi = 0.0;
    int ninf_lo = 0, ninf_hi = 0;
    for (int t = r.beg + 2 * lane_g; t < end; t += 2 * G) {
      const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
      const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double a = h ? a2.y : a2.x;
        const int j = h ? c2.y : c2.x;
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

  if (ll > rc.ru + kETol || uu < rc.rl - kETol) {     // activity-infeasible row
    if (lane_g == 0) sink.row_infeasible(status);
    return;
  }
  rc.do_lb = rc.do_ub = rc.sing_lb = rc.sing_ub = false;
  rc.act_lb = rc.act_ub = 0.0;
  if (rc.rl > -kInf20) {
    if (uu < kInf20) { rc.do_lb = true; rc.act_lb = uu; }
    else if (sing_uu < kInf20) { rc.do_lb = true; rc.sing_lb = true; rc.act_lb = sing_uu; }
  }
  if (rc.ru < kInf20) {
    if (ll > -kInf20) { rc.do_ub = true; rc.act_ub = ll; }
    else if (sing_ll > -kInf20) { rc.do_ub = true; rc.sing_ub = true; rc.act_ub = sing_ll; }
  }
  if (!rc.do_lb && !rc.do_ub) return;
  // pass 2: implied bounds; the first pair and its bounds are still in registers
  if (d.a2.x != 0.0) emit_candidates<R>(rc, d.a2.x, d.c2.x, b0, sink);
  if (d.a2.y != 0.0) emit_candidates<R>(rc, d.a2.y, d.c2.y, b1, sink);
  if (long_row) {
    for (int t = r.beg + 2 * lane_g + 2 * G; t < end; t += 2 * G) {
      const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
      const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
      if (a2.x != 0.0) emit_candidates<R>(rc, a2.x, c2.x, box[c2.x], sink);
      if (a2.y != 0.0) emit_candidates<R>(rc, a2.y, c2.y, box[c2.y], sink);
    }
  }
}

}  // namespace mntr
