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

}  // namespace

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
