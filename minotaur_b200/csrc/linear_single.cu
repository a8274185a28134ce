// linear_single.cu -- K1: single-box FBBT of the linear rows, Jacobi rounds to a fixpoint
// inside ONE cooperative launch (device-side change flag, no host round trips).
//
// Per round (reference lines in brackets):
//   rows phase : sub-warp group per flagged CSR row -- 128-bit loads of (val,val) and
//                64-bit loads of (col,col), 128-bit gathers of {lb,ub}; min/max activity
//                with outward rounding [getLfBnds_ LinearHandler.cpp:1237-1258]; singleton-
//                infinity sums by finite-sum + infinity-count [getSingLfBnds_ :1261-1319];
//                warp-shuffle butterfly reduction; activity infeasibility [:994-1015];
//                implied bounds [updateLfBoundsFromLb_/Ub_ :1048-1226] merged with fp64
//                atomic max/min into the next box.
//   vars phase : integer rounding [tightenInts_ :415-490], lb>ub check [checkBounds_
//                :328-359], change detection, CSC flagging of the rows to re-evaluate
//                [changeBFlag_ :1229-1234], device-side change flag.
// grid.sync() separates the phases; the loop condition is evaluated on the device.
#include <cooperative_groups.h>

#include "device_problem.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace mntr {

namespace {

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

// one term of pass 2: candidates of updateLfBoundsFromLb_ / updateLfBoundsFromUb_
template <class R>
__device__ __forceinline__ void emit_candidates(const RowCtx &rc, double a, int j, double2 b,
                                                double2 *nbox)
{
  const double vl = b.x, vu = b.y;
  if (rc.do_lb) {                                   // row lb side, activity = max activity
    const double numer = R::sub_lo(rc.rl, rc.act_lb);
    if (a > kETol && (!rc.sing_lb || vu >= kInf20)) {
      const double base = (vu >= kInf20) ? 0.0 : vu;
      double c = R::add_lo(R::div_lo(numer, a), base);
      if (c > vl + kETol) {
        if (c > vu - kETol) c = vu;
        atomic_max_f64(&nbox[j].x, c);
      }
    } else if (a < -kETol && (!rc.sing_lb || vl <= -kInf20)) {
      const double base = (vl <= -kInf20) ? 0.0 : vl;
      double c = R::add_hi(R::div_hi(numer, a), base);
      if (c < vu - kETol) {
        if (c < vl + kETol) c = vl;
        atomic_min_f64(&nbox[j].y, c);
      }
    }
  }
  if (rc.do_ub) {                                   // row ub side, activity = min activity
    const double numer = R::sub_hi(rc.ru, rc.act_ub);
    if (a > kETol && (!rc.sing_ub || vl <= -kInf20)) {
      const double base = (vl <= -kInf20) ? 0.0 : vl;
      double c = R::add_hi(R::div_hi(numer, a), base);
      if (c < vu - kETol) {
        if (c < vl + kETol) c = vl;
        atomic_min_f64(&nbox[j].y, c);
      }
    } else if (a < -kETol && (!rc.sing_ub || vu >= kInf20)) {
      const double base = (vu >= kInf20) ? 0.0 : vu;
      double c = R::add_lo(R::div_lo(numer, a), base);
      if (c > vl + kETol) {
        if (c > vu - kETol) c = vu;
        atomic_max_f64(&nbox[j].x, c);
      }
    }
  }
}

template <int G, class R>
__global__ void __launch_bounds__(256)
fbbt_single_jacobi_kernel(LinDev P, SingleWs W, double *lb_io, double *ub_io, int max_rounds,
                          int loop_mode)
{
  cg::grid_group grid = cg::this_grid();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int lane_g = lane % G;
  const int group = tid / G;
  const int n_groups = nthreads / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - lane_g));

  // ---- phase 0: build {lb,ub} boxes, flag every active row (simplePresolve :1618-1622) ----
  int infeasible0 = 0;
  for (int j = tid; j < P.n; j += nthreads) {
    double2 b = make_double2(lb_io[j], ub_io[j]);
    W.box[j] = b;
    W.nbox[j] = b;
  }
  for (int i = tid; i < P.m; i += nthreads) {
    const bool act = P.row_active[i] != 0;
    W.flag_a[i] = act ? 1 : 0;
    W.flag_b[i] = 0;
    if (act && P.row_lb[i] > P.row_ub[i] + kETol) infeasible0 = 1;   // checkBounds_, rows part
  }
  if (infeasible0) W.status[0] = 1 /* MNTR_INFEAS_BOUNDS */;
  grid.sync();
  volatile int32_t *vstatus = W.status;   // control words are re-read after every grid.sync
  volatile int32_t *vring = W.ring;

  uint8_t *fcur = W.flag_a, *fnext = W.flag_b;
  unsigned long long my_nnz = 0, my_rows = 0;
  int round = 0;
  int verdict = vstatus[0];

  while (verdict == 0) {
    ++round;
    const int slot = round % 3;
    if (tid == 0) { W.ring[(round + 1) % 3] = 0; W.ring[3 + (round + 1) % 3] = 0; }

    // ------------------------------ rows phase ------------------------------
    for (int i = group; i < P.m; i += n_groups) {
      int f = 0;
      if (lane_g == 0) { f = fcur[i]; if (f) fcur[i] = 0; }
      f = __shfl_sync(gmask, f, 0, G);
      if (!f) continue;
      const int beg = P.row_ptr[i], end = P.row_ptr[i + 1];
      RowCtx rc;
      rc.rl = P.row_lb[i];
      rc.ru = P.row_ub[i];
      // pass 1: min / max activity
      double ll = 0.0, uu = 0.0;
      for (int t = beg + 2 * lane_g; t < end; t += 2 * G) {
        const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
        const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
        if (a2.x != 0.0) {
          const double2 b = W.box[c2.x];
          if (a2.x > 0) { ll = R::add_lo(ll, R::mul_lo(a2.x, b.x)); uu = R::add_hi(uu, R::mul_hi(a2.x, b.y)); }
          else          { ll = R::add_lo(ll, R::mul_lo(a2.x, b.y)); uu = R::add_hi(uu, R::mul_hi(a2.x, b.x)); }
        }
        if (a2.y != 0.0) {
          const double2 b = W.box[c2.y];
          if (a2.y > 0) { ll = R::add_lo(ll, R::mul_lo(a2.y, b.x)); uu = R::add_hi(uu, R::mul_hi(a2.y, b.y)); }
          else          { ll = R::add_lo(ll, R::mul_lo(a2.y, b.y)); uu = R::add_hi(uu, R::mul_hi(a2.y, b.x)); }
        }
      }
      ll = group_reduce<G, R, true>(ll, gmask);
      uu = group_reduce<G, R, false>(uu, gmask);
      if (lane_g == 0) { my_nnz += (unsigned long long)P.row_nnz[i]; ++my_rows; }

      // singleton-infinity sums, only when an activity is beyond +-1e20 (:970-972)
      double sing_ll = -INFINITY, sing_uu = INFINITY;
      if (ll < -kInf20 || uu > kInf20) {
        double fs_lo = 0.0, fs_hi = 0.0;
        int ninf_lo = 0, ninf_hi = 0;
        for (int t = beg + 2 * lane_g; t < end; t += 2 * G) {
          const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
          const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double a = h ? a2.y : a2.x;
            const int j = h ? c2.y : c2.x;
            if (a > kETol) {
              const double2 b = W.box[j];
              if (b.y < kInf20) fs_hi = R::add_hi(fs_hi, R::mul_hi(a, b.y)); else ++ninf_hi;
              if (b.x > -kInf20) fs_lo = R::add_lo(fs_lo, R::mul_lo(a, b.x)); else ++ninf_lo;
            } else if (a < -kETol) {
              const double2 b = W.box[j];
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
        if (lane_g == 0) W.status[0] = 2 /* MNTR_INFEAS_ROW */;
        continue;
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
      if (!rc.do_lb && !rc.do_ub) continue;
      // pass 2: implied bounds (the row's entries are L1/L2 hits now)
      for (int t = beg + 2 * lane_g; t < end; t += 2 * G) {
        const double2 a2 = ldg_f64x2(reinterpret_cast<const double2 *>(P.val + t));
        const int2 c2 = ldg_i32x2(reinterpret_cast<const int2 *>(P.col + t));
        if (a2.x != 0.0) emit_candidates<R>(rc, a2.x, c2.x, W.box[c2.x], W.nbox);
        if (a2.y != 0.0) emit_candidates<R>(rc, a2.y, c2.y, W.box[c2.y], W.nbox);
      }
    }
    grid.sync();
    verdict = vstatus[0];
    if (verdict != 0) break;

    // ------------------------------ vars phase ------------------------------
    int changed = 0, int_moved = 0, bad = 0;
    int n_changed = 0;
    for (int j = tid; j < P.n; j += nthreads) {
      const double2 o = W.box[j];
      double2 v = W.nbox[j];
      const bool isint = is_int_type(P.var_type[j]);
      if (isint) {
        if (v.x != o.x || v.y != o.y) int_moved = 1;      // row-derived mod on an int var (nintmods)
        tighten_int_bounds(v.x, v.y);
      }
      if (v.x > v.y + kETol) bad = 1;
      if (v.x != o.x || v.y != o.y) {
        changed = 1;
        ++n_changed;
        W.box[j] = v;
        W.nbox[j] = v;
        for (int q = P.csc_ptr[j]; q < P.csc_ptr[j + 1]; ++q) fnext[P.csc_row[q]] = 1;
      }
    }
    if (n_changed) atomicAdd(&W.status[2], n_changed);
    changed = __syncthreads_or(changed);
    int_moved = __syncthreads_or(int_moved);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
      if (changed) W.ring[slot] = 1;
      if (int_moved) W.ring[3 + slot] = 1;
      if (bad) W.status[0] = 1 /* MNTR_INFEAS_BOUNDS */;
    }
    grid.sync();
    verdict = vstatus[0];
    const int any_changed = vring[slot];
    const int any_int = vring[3 + slot];
    { uint8_t *t = fcur; fcur = fnext; fnext = t; }
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
  if (lane == 0 && my_rows) { atomicAdd(&W.counters[0], my_nnz); atomicAdd(&W.counters[1], my_rows); }
  if (tid == 0) W.status[1] = round;
}

template <int G, class R>
cudaError_t launch_g(const LinDev &P, const SingleWs &W, double *lb, double *ub, int max_rounds,
                     int loop_mode, int sm_count, cudaStream_t stream)
{
  auto kern = fbbt_single_jacobi_kernel<G, R>;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  long long want_threads = (long long)P.m * G;
  if (want_threads < P.n) want_threads = P.n;
  long long want_blocks = (want_threads + 255) / 256;
  long long max_blocks = (long long)per_sm * sm_count;
  int blocks = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
  if (blocks < 1) blocks = 1;
  LinDev p = P; SingleWs w = W;
  void *args[] = { &p, &w, &lb, &ub, &max_rounds, &loop_mode };
  return cudaLaunchCooperativeKernel((void *)kern, dim3(blocks), dim3(256), args, 0, stream);
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
