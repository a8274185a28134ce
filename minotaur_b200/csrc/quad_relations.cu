// quad_relations.cu -- QuadHandler::simplePresolve (QuadHandler.cpp:1146-1201) for a batch of node boxes: ONE in-place
// sweep over the relations the handler holds after the reformulation, y = x^2 (x2Funs_, ascending x) and then
// y = x0 * x1 (x0x1Funs_, ascending (x0, x1)), every step through updatePBounds_ (:3218-3246).
//
// Layout and decomposition are the batch kernel's (linear_batch.cu): boxes node-minor double2 {lb,ub} [n][ld], one
// CTA per tile of 32 boxes, lane = box; the sweep is sequential in the reference, so the relations are scheduled in
// wavefront levels (level = 1 + the highest level of an earlier relation sharing a variable -- every relation reads
// and may write all of its variables), a level's relations are dealt to the warps, a CTA barrier ends a level.  With
// round-to-nearest every lane performs the reference's operations in the reference's order: results are bitwise the
// reference's.  With directed rounding the forward bounds and the square roots are rounded outward.
//
// Second kernel: the propagation loop of QuadHandler::presolveNode (:1204-1239) -- the same sweep repeated to its fixpoint
// with the relaxation-aware updatePBounds_ (:3248-3320) and an infeasibility verdict per box.
#include "cgraph.cuh"
#include "device_problem.cuh"
#include "kernels.h"

namespace mntr {

namespace {

constexpr int kQrelWarps = 8;

// updatePBounds_(p, v, lb, ub, mods), QuadHandler.cpp:3218-3246: integer rounding, consistency, then each side moves
// if it improves by more than the absolute AND the relative tolerance (ub: bTol 1e-8, lb: aTol 1e-6; rTol 1e-7).
// returns -1 (inconsistent: nothing changes), else the number of sides that moved (b is updated)
__device__ __forceinline__ int qh_update(uint8_t ty, double2 &b, double lb, double ub)
{
  const double aTol = 1e-6, bTol = 1e-8, rTol = 1e-7;
  if (ty != 4) { ub = floor(ub); lb = ceil(lb); }          // Binary, Integer, ImplBin, ImplInt (Types.h:83-89)
  if (ub < b.x - bTol || lb > b.y + bTol) return -1;
  int moved = 0;
  if (ub < b.y - bTol && (b.y == INFINITY || ub < b.y - fabs(b.y) * rTol)) { b.y = ub; ++moved; }
  if (lb > b.x + aTol && (b.x == -INFINITY || lb > b.x + fabs(b.x) * rTol)) { b.x = lb; ++moved; }
  return moved;
}

template <class R>
__global__ void __launch_bounds__(kQrelWarps * 32)
quad_relations_kernel(QRelDev Q, double2 *boxes, int64_t ld, int32_t n_boxes, int32_t *n_mods, int32_t *n_bad)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int box = blockIdx.x * 32 + lane;
  double2 *bx = boxes + box;                       // + j * ld: variable j of this lane's box (padding boxes are harmless copies)
  int mods = 0, bad = 0;
  for (int lev = 0; lev < Q.n_levels; ++lev) {
    const int rb = __ldg(Q.level_ptr + lev), re = __ldg(Q.level_ptr + lev + 1);
    for (int r = rb + warp; r < re; r += kQrelWarps) {
      const int a = __ldg(Q.a + r), b = __ldg(Q.b + r), y = __ldg(Q.y + r);
      double2 *pa = bx + (int64_t)a * ld, *py = bx + (int64_t)y * ld;
      double2 va = *pa, vy = *py;
      const uint8_t ty_a = __ldg(Q.var_type + a), ty_y = __ldg(Q.var_type + y);
      double lb, ub;
      if (b < 0) {
        // ---- y = x^2 (:1153-1177) ----
        bounds_on_square<R>(va.x, va.y, lb, ub);
        int k = qh_update(ty_y, vy, lb, ub);
        if (k < 0) ++bad; else if (k) { mods += k; *py = vy; }
        const double bTol = 1e-8;
        if (vy.y > bTol) {
          ub = R::sqrt_hi(vy.y);
          lb = -ub;
          if (va.x > -R::sqrt_lo(vy.x) + bTol) lb = R::sqrt_lo(vy.x);      // (the test with the root rounded toward 0: harder to pass)
          k = qh_update(ty_a, va, lb, ub);
          if (k < 0) ++bad; else if (k) { mods += k; *pa = va; }
        } else if (vy.y < -bTol) {
          ++bad;
        } else {
          k = qh_update(ty_a, va, 0.0, 0.0);
          if (k < 0) ++bad; else if (k) { mods += k; *pa = va; }
        }
      } else {
        // ---- y = x0 * x1 (:1179-1198): forward, then x1 from y / x0, then x0 from y / x1 (the already tightened x1) ----
        double2 *pb = bx + (int64_t)b * ld;
        double2 vb = *pb;
        const uint8_t ty_b = __ldg(Q.var_type + b);
        bounds_on_product<R>(true, va.x, va.y, vb.x, vb.y, lb, ub);
        int k = qh_update(ty_y, vy, lb, ub);
        if (k < 0) ++bad; else if (k) { mods += k; *py = vy; }
        bounds_on_div<R>(vy.x, vy.y, va.x, va.y, lb, ub);
        k = qh_update(ty_b, vb, lb, ub);
        if (k < 0) ++bad; else if (k) { mods += k; *pb = vb; }
        bounds_on_div<R>(vy.x, vy.y, vb.x, vb.y, lb, ub);
        k = qh_update(ty_a, va, lb, ub);
        if (k < 0) ++bad; else if (k) { mods += k; *pa = va; }
      }
    }
    __syncthreads();
  }
  // per box: the warps' counts
  __shared__ int s_mods[kQrelWarps][32], s_bad[kQrelWarps][32];
  s_mods[warp][lane] = mods; s_bad[warp][lane] = bad;
  __syncthreads();
  if (warp == 0 && box < n_boxes) {
    int m = 0, d = 0;
#pragma unroll
    for (int w = 0; w < kQrelWarps; ++w) { m += s_mods[w][lane]; d += s_bad[w][lane]; }
    n_mods[box] = m; n_bad[box] = d;
  }
}

// updatePBounds_(v, lb, ub, rel, mod_rel, changed, p_mods, r_mods), QuadHandler.cpp:3248-3320, the relaxation-aware
// variant presolveNode uses: bTol on both sides; each side moves exactly when its own absolute and relative tests pass
// (the reference's three branches -- both sides as one VarBoundMod2, else lower, else upper -- only decide how many
// Modification objects are pushed: one per call that moves anything).
// returns -1 (inconsistent), 0 (nothing moved) or 1 (b updated: one Modification)
__device__ __forceinline__ int qh_update_node(uint8_t ty, double2 &b, double lb, double ub)
{
  const double bTol = 1e-8, rTol = 1e-7;
  if (ty != 4) { ub = floor(ub); lb = ceil(lb); }
  if (lb > b.y + bTol || ub < b.x - bTol) return -1;
  const bool lo = lb > b.x + bTol && (b.x == -INFINITY || lb > b.x + rTol * fabs(b.x));
  const bool up = ub < b.y - bTol && (b.y == INFINITY || ub < b.y - rTol * fabs(b.y));
  if (lo) b.x = lb;
  if (up) b.y = ub;
  return (lo || up) ? 1 : 0;
}

// The propagation loop of QuadHandler::presolveNode (QuadHandler.cpp:1214-1239) for a tile of 32 boxes: sweeps
// { propSqrBnds_ (:1361-1395) over the squares, propBilBnds_ (:1271-1301) over the products } in place -- in wavefront
// levels, like the single sweep above -- repeated per box while its last sweep moved a bound; the first inconsistent
// step makes the box infeasible and ends its propagation (the reference returns there, :1224 / :1233: the bounds of
// such a box are not a result).  A box's state lives in shared memory: changed (any warp sets it), infeasible (any
// warp sets it; sampled by all warps at level boundaries -- between them a warp stops at its own finding only, the
// relations other warps still evaluate in that level are ones the reference would not have reached, and they can only
// touch a box that is already discarded).
template <class R>
__global__ void __launch_bounds__(kQrelWarps * 32)
quad_node_kernel(QRelDev Q, double2 *boxes, int64_t ld, int32_t n_boxes, int32_t max_sweeps, int32_t *verdict, int32_t *n_mods,
                 int32_t *n_sweeps)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int box = blockIdx.x * 32 + lane;
  double2 *bx = boxes + box;
  __shared__ int s_changed[32], s_inf[32];
  __shared__ int s_mods[kQrelWarps][32];
  if (warp == 0) { s_changed[lane] = 1; s_inf[lane] = 0; }
  int mods = 0, sweeps = 0;
  for (;;) {
    __syncthreads();
    const bool run = s_changed[lane] != 0 && s_inf[lane] == 0 && (max_sweeps <= 0 || sweeps < max_sweeps);
    const unsigned runmask = __ballot_sync(0xffffffffu, run);
    __syncthreads();                                     // every warp has read the flags of the last sweep
    if (runmask == 0u) break;
    if (warp == 0) s_changed[lane] = 0;
    if (run) ++sweeps;
    bool dead = false, moved = false;
    for (int lev = 0; lev < Q.n_levels; ++lev) {
      const int rb = __ldg(Q.level_ptr + lev), re = __ldg(Q.level_ptr + lev + 1);
      __syncthreads();
      const bool live0 = run && s_inf[lane] == 0;
      if (__ballot_sync(0xffffffffu, live0) == 0u) continue;        // (uniform per CTA: every warp reads the same words)
      for (int r = rb + warp; r < re; r += kQrelWarps) {
        const int a = __ldg(Q.a + r), b = __ldg(Q.b + r), y = __ldg(Q.y + r);
        double2 *pa = bx + (int64_t)a * ld, *py = bx + (int64_t)y * ld;
        double2 va = *pa, vy = *py;
        const uint8_t ty_a = __ldg(Q.var_type + a), ty_y = __ldg(Q.var_type + y);
        const bool live = live0 && !dead;
        double lb, ub;
        int k;
        if (b < 0) {
          // ---- propSqrBnds_ ----
          bounds_on_square<R>(va.x, va.y, lb, ub);
          k = qh_update_node(ty_y, vy, lb, ub);
          bool stop = k < 0;
          if (k > 0 && live) { ++mods; moved = true; *py = vy; }
          if (!stop) {
            const double bTol = 1e-8;
            if (vy.y > bTol) {
              ub = R::sqrt_hi(vy.y);
              lb = -ub;
              if (va.x > -R::sqrt_lo(vy.x) + bTol) lb = R::sqrt_lo(vy.x);
              k = qh_update_node(ty_a, va, lb, ub);
            } else if (vy.y < -bTol) {
              k = -1;
            } else {
              k = qh_update_node(ty_a, va, 0.0, 0.0);
            }
            if (k < 0) stop = true;
            else if (k > 0 && live) { ++mods; moved = true; *pa = va; }
          }
          if (stop && live) dead = true;
        } else {
          // ---- propBilBnds_: y from x0 * x1, x1 from y / x0, x0 from y / x1 ----
          double2 *pb = bx + (int64_t)b * ld;
          double2 vb = *pb;
          const uint8_t ty_b = __ldg(Q.var_type + b);
          bounds_on_product<R>(true, va.x, va.y, vb.x, vb.y, lb, ub);
          k = qh_update_node(ty_y, vy, lb, ub);
          bool stop = k < 0;
          if (k > 0 && live) { ++mods; moved = true; *py = vy; }
          if (!stop) {
            bounds_on_div<R>(vy.x, vy.y, va.x, va.y, lb, ub);
            k = qh_update_node(ty_b, vb, lb, ub);
            if (k < 0) stop = true;
            else if (k > 0 && live) { ++mods; moved = true; *pb = vb; }
          }
          if (!stop) {
            bounds_on_div<R>(vy.x, vy.y, vb.x, vb.y, lb, ub);
            k = qh_update_node(ty_a, va, lb, ub);
            if (k < 0) stop = true;
            else if (k > 0 && live) { ++mods; moved = true; *pa = va; }
          }
          if (stop && live) dead = true;
        }
      }
      if (dead) s_inf[lane] = 1;
    }
    if (moved) s_changed[lane] = 1;
  }
  s_mods[warp][lane] = mods;
  __syncthreads();
  if (warp == 0 && box < n_boxes) {
    int m = 0;
#pragma unroll
    for (int w = 0; w < kQrelWarps; ++w) m += s_mods[w][lane];
    n_mods[box] = m; verdict[box] = s_inf[lane] ? 1 : 0; n_sweeps[box] = sweeps;
  }
}

}  // namespace

cudaError_t launch_quad_node(const QRelDev &Q, double2 *boxes, int64_t ld, int32_t n_boxes, bool directed, int32_t max_sweeps,
                             int32_t *verdict, int32_t *n_mods, int32_t *n_sweeps, cudaStream_t stream)
{
  const int tiles = (n_boxes + 31) / 32;
  if (tiles <= 0) return cudaSuccess;
  if (directed) quad_node_kernel<RoundDirected><<<tiles, kQrelWarps * 32, 0, stream>>>(Q, boxes, ld, n_boxes, max_sweeps, verdict, n_mods, n_sweeps);
  else quad_node_kernel<RoundNearest><<<tiles, kQrelWarps * 32, 0, stream>>>(Q, boxes, ld, n_boxes, max_sweeps, verdict, n_mods, n_sweeps);
  return cudaGetLastError();
}

cudaError_t launch_quad_relations(const QRelDev &Q, double2 *boxes, int64_t ld, int32_t n_boxes, bool directed, int32_t *n_mods,
                                  int32_t *n_bad, cudaStream_t stream)
{
  const int tiles = (n_boxes + 31) / 32;
  if (tiles <= 0) return cudaSuccess;
  if (directed) quad_relations_kernel<RoundDirected><<<tiles, kQrelWarps * 32, 0, stream>>>(Q, boxes, ld, n_boxes, n_mods, n_bad);
  else quad_relations_kernel<RoundNearest><<<tiles, kQrelWarps * 32, 0, stream>>>(Q, boxes, ld, n_boxes, n_mods, n_bad);
  return cudaGetLastError();
}

}  // namespace mntr
