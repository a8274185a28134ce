// common.cuh -- shared device helpers of the B200 FBBT engine (sm_100a).
//
// Tolerances and the "three infinities" follow the reference:
//   LinearHandler.cpp:69-71   intTol_=1e-6, eTol_=1e-8, infty_=1e20
//   CNode.cpp:26              MINFTY=1e25 clamp of CGraph node bounds
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mntr {

constexpr double kIntTol = 1e-6;
constexpr double kETol   = 1e-8;
constexpr double kInf20  = 1e20;
constexpr double kCoefDrop = 1e-9;   // LinearFunction::tol_, LinearFunction.cpp:20-23
constexpr int    kTile   = 32;       // boxes per warp / per CTA tile in the batched kernels

__host__ __device__ inline bool is_int_type(uint8_t t) { return t <= 1; }  // Binary, Integer

// ---------------------------------------------------------------------------------------
// Rounding policies.  "lo" results may only err downward, "hi" results only upward, so a
// lower bound built from lo-ops and an upper bound built from hi-ops are never tighter than
// exact arithmetic.  Nearest reproduces the reference's plain, unfused IEEE arithmetic
// (the reference builds without FMA contraction; __dmul_rn/__dadd_rn are never fused).
// ---------------------------------------------------------------------------------------
struct RoundDirected {
  static __device__ __forceinline__ double mul_lo(double a, double b) { return __dmul_rd(a, b); }
  static __device__ __forceinline__ double mul_hi(double a, double b) { return __dmul_ru(a, b); }
  static __device__ __forceinline__ double add_lo(double a, double b) { return __dadd_rd(a, b); }
  static __device__ __forceinline__ double add_hi(double a, double b) { return __dadd_ru(a, b); }
  static __device__ __forceinline__ double sub_lo(double a, double b) { return __dadd_rd(a, -b); }
  static __device__ __forceinline__ double sub_hi(double a, double b) { return __dadd_ru(a, -b); }
  static __device__ __forceinline__ double div_lo(double a, double b) { return __ddiv_rd(a, b); }
  static __device__ __forceinline__ double div_hi(double a, double b) { return __ddiv_ru(a, b); }
  static __device__ __forceinline__ double sqrt_lo(double a) { return __dsqrt_rd(a); }
  static __device__ __forceinline__ double sqrt_hi(double a) { return __dsqrt_ru(a); }
};

struct RoundNearest {
  static __device__ __forceinline__ double mul_lo(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double mul_hi(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add_lo(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double add_hi(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub_lo(double a, double b) { return __dadd_rn(a, -b); }
  static __device__ __forceinline__ double sub_hi(double a, double b) { return __dadd_rn(a, -b); }
  static __device__ __forceinline__ double div_lo(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double div_hi(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt_lo(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double sqrt_hi(double a) { return __dsqrt_rn(a); }
};

// ---------------------------------------------------------------------------------------
// fp64 atomic max / min on raw IEEE storage (no native double atomicMax): non-negative
// doubles order like signed int64, negative doubles order inversely like uint64.
// NaN candidates are never passed in (callers test them away).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_f64(double *addr, double v)
{
  v += 0.0;  // -0.0 -> +0.0
  if (v >= 0.0) atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
  else atomicMin(reinterpret_cast<unsigned long long *>(addr),
                 static_cast<unsigned long long>(__double_as_longlong(v)));
}

__device__ __forceinline__ void atomic_min_f64(double *addr, double v)
{
  v += 0.0;
  if (v >= 0.0) atomicMin(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
  else atomicMax(reinterpret_cast<unsigned long long *>(addr),
                 static_cast<unsigned long long>(__double_as_longlong(v)));
}

// 128-bit read-only loads (LDG.E.128)
__device__ __forceinline__ double2 ldg_f64x2(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ int2 ldg_i32x2(const int2 *p) { return __ldg(p); }

// Product test of the Jacobi kernels.  A term a*x_j of a row with slack s (row bound minus the activity of the OTHER
// bound side) moves a bound of x_j towards the other by s/|a|, and updateLfBoundsFromLb_/Ub_ accept the new bound only
// if it improves by more than eTol (LinearHandler.cpp:1070): in exact arithmetic iff  s < |a| (ub - lb - eTol).
// term_reach() is an upper bound of that threshold with room for the rounding of the exact path (half of eTol
// absolute, 1e-9 relative; bounds beyond 1e6 in magnitude, infinities and NaN always pass): a term with
// s >= term_reach() cannot be accepted, so the fp64 divisions are skipped for it.  Test it as  !(s >= reach).
__device__ __forceinline__ double term_reach(double a, double2 b)
{
  const double w = fabs(a) * ((b.y - b.x) - 0.5 * kETol) * 1.000000001;
  return (fmax(fabs(b.x), fabs(b.y)) > 1e6) ? INFINITY : w;
}

// integer rounding of one bound pair: LinearHandler::tightenInts_, LinearHandler.cpp:415-490
__device__ __forceinline__ void tighten_int_bounds(double &lb, double &ub)
{
  if (lb > -kInf20 && fabs(lb - floor(lb + 0.5)) > kIntTol) lb = ceil(lb);
  if (ub < kInf20 && fabs(ub - floor(ub + 0.5)) > kIntTol) ub = floor(ub);
}

}  // namespace mntr
