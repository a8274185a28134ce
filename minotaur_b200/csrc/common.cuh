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

// ---------------------------------------------------------------------------------------
// L2 residency of the streaming forms.  A large instance streams hundreds of MB of matrix through L2 in every dense
// round while its rows gather {lb,ub} from a box of tens of MB at random: left to the default policy the stream
// evicts the box and three gathers in four go to DRAM as 32-byte reads (measured: 1.30 GB of DRAM reads for 0.45 GB
// of matrix).  The stream is therefore marked evict-first (cp.async with an L2 cache-hint policy, ld.global.cs for
// register loads) and the gathers evict-last.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async16_stream(unsigned dst_smem, const void *src, unsigned long long policy)
{
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(dst_smem), "l"(src), "l"(policy) : "memory");
}
// {lb,ub} of one variable, kept in L2 (bypasses L1 like ld.cg: boxes are updated by other SMs between rounds)
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_box_keep(const double2 *p, unsigned long long policy)
{
  double2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
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
