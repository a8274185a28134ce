// cgraph.cuh -- K4 device code: interval evaluation of CGraph expression tapes.
//
// One lane evaluates one (constraint, box) pair; all lanes of a warp walk the SAME tape, so the
// opcode dispatch is warp-uniform and the tape entries are broadcast loads.  Node intervals live
// in per-thread local arrays (indexed by the warp-uniform node index: coalesced local memory).
//
// Reference rules reproduced (file:line under /root/reference/src/base/):
//   forward  CNode::updateBnd        CNode.cpp:1701-1904   (incl. the (-inf,inf) "TODO" opcodes)
//   reverse  CNode::propBounds       CNode.cpp:1259-1501   (incl. the OpSqr / OpSqrt / OpPowK / OpSumList
//            CNode::propBounds_      CNode.cpp:1504-1526    quirks listed in SURVEY.md section 7, hard part 3)
//   helpers  BoundsOnProduct/Recip/Div/Square   Operations.cpp:100-246
//   driver   CGraph::varBoundMods    CGraph.cpp:1605-1644  (1e-4 acceptance, 1e-5 relaxation)
// With RoundNearest every lane performs the reference's operation sequence, so +,-,*,/,sqrt results are
// bitwise the reference's; exp/log/log10/pow come from CUDA's libm (<= 2 ulp from glibc).  With
// RoundDirected lower bounds are rounded down and upper bounds up (transcendentals widened by 2 ulp).
#pragma once
#include "device_problem.cuh"

namespace mntr {

constexpr int kMaxTape = 48;          // longest supported tape (nodes per constraint)
constexpr double kMinfty = 1e25;      // CNode.cpp:26

// Minotaur::OpCode (OpCode.h:17-53)
enum : int {
  OpAbs = 0, OpAcos, OpAcosh, OpAsin, OpAsinh, OpAtan, OpAtanh, OpCeil, OpCos, OpCosh, OpCPow, OpDiv, OpExp,
  OpFloor, OpInt, OpIntDiv, OpLog, OpLog10, OpMinus, OpMult, OpNone, OpNum, OpPlus, OpPow, OpPowK, OpRound,
  OpSin, OpSinh, OpSqr, OpSqrt, OpSumList, OpTan, OpTanh, OpUMinus, OpVar
};

// std::min / std::max as the reference uses them (first argument wins ties: zero signs preserved)
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }

__device__ __forceinline__ double widen_lo(double x, bool directed)
{
  // two representable steps down: adding the smallest subnormal with round-down moves one step
  return (directed && isfinite(x)) ? __dadd_rd(__dadd_rd(x, -4.9406564584124654e-324), -4.9406564584124654e-324) : x;
}
__device__ __forceinline__ double widen_hi(double x, bool directed)
{
  return (directed && isfinite(x)) ? __dadd_ru(__dadd_ru(x, 4.9406564584124654e-324), 4.9406564584124654e-324) : x;
}

template <class R> struct RoundTraits { static constexpr bool directed = true; };
template <> struct RoundTraits<RoundNearest> { static constexpr bool directed = false; };

// BoundsOnProduct, Operations.cpp:117-179
template <class R>
__device__ __forceinline__ void bounds_on_product(bool zero_x_inf_zero, double l0, double u0, double l1, double u1,
                                                  double &lb, double &ub)
{
  if (fabs(l1) <= 1e-10 && fabs(u1) <= 1e-10) {
    double t = l1; l1 = l0; l0 = t;
    t = u1; u1 = u0; u0 = t;
  }
  if (fabs(l0) <= 1e-10 && fabs(u0) <= 1e-10) {
    if (zero_x_inf_zero) { lb = 0.0; ub = 0.0; }
    else {
      lb = (l1 == -INFINITY) ? -INFINITY : 0.0;
      ub = (u1 == INFINITY) ? INFINITY : 0.0;
    }
  } else if ((l1 == -INFINITY && u1 == INFINITY) || (l0 == -INFINITY && u0 == INFINITY)) {
    lb = -INFINITY; ub = INFINITY;
  } else {
    // four corner products; 0*inf (NaN) is mapped as the reference maps it
    double pl = R::mul_lo(l0, l1), ph = R::mul_hi(l0, l1);
    if (isnan(pl)) { pl = -INFINITY; ph = -INFINITY; }
    double l = pl, u = ph;
    pl = R::mul_lo(u0, l1); ph = R::mul_hi(u0, l1);
    if (isnan(pl)) { pl = INFINITY; ph = INFINITY; }
    l = std_min(l, pl); u = std_max(u, ph);
    pl = R::mul_lo(u0, u1); ph = R::mul_hi(u0, u1);
    if (isnan(pl)) { pl = -INFINITY; ph = -INFINITY; }
    l = std_min(l, pl); u = std_max(u, ph);
    pl = R::mul_lo(l0, u1); ph = R::mul_hi(l0, u1);
    if (isnan(pl)) { pl = INFINITY; ph = INFINITY; }
    l = std_min(l, pl); u = std_max(u, ph);
    lb = l; ub = u;
  }
}

// BoundsOnRecip, Operations.cpp:182-212
template <class R>
__device__ __forceinline__ void bounds_on_recip(double l0, double u0, double &lb, double &ub)
{
  if ((fabs(u0) < 1e-10) && (fabs(l0) < 1e-10)) { lb = -INFINITY; ub = INFINITY; }
  else if (l0 < -1e-10 && u0 > 1e-10) { lb = -INFINITY; ub = INFINITY; }
  else if ((fabs(u0) < 1e-10) && l0 < 0) { lb = -INFINITY; ub = R::div_hi(1.0, l0); }
  else if ((fabs(l0) < 1e-10) && u0 < 0) { lb = R::div_lo(1.0, u0); ub = INFINITY; }
  else { lb = R::div_lo(1.0, u0); ub = R::div_hi(1.0, l0); }
}

// BoundsOnDiv, Operations.cpp:100-106
template <class R>
__device__ __forceinline__ void bounds_on_div(double l0, double u0, double l1, double u1, double &lb, double &ub)
{
  double tl, tu;
  bounds_on_recip<R>(l1, u1, tl, tu);
  bounds_on_product<R>(false, l0, u0, tl, tu, lb, ub);
}

// BoundsOnSquare, Operations.cpp:233-246
template <class R>
__device__ __forceinline__ void bounds_on_square(double l1, double u1, double &lb, double &ub)
{
  if (u1 < 0.) { lb = R::mul_lo(u1, u1); ub = R::mul_hi(l1, l1); }
  else if (l1 > 0.) { lb = R::mul_lo(l1, l1); ub = R::mul_hi(u1, u1); }
  else { lb = 0.; ub = std_max(R::mul_hi(l1, l1), R::mul_hi(u1, u1)); }
}

// isInt(v, 1e-12), Operations.cpp:80-83
__device__ __forceinline__ bool is_int_val(double v) { return fabs(floor(v + 0.5) - v) < 1e-12; }

// libm calls with the errno behaviour the reference observes (domain / range errors make
// CGraph::computeBounds / varBoundMods return SolveError)
__device__ __forceinline__ double chk(double r, double x, int &error)
{
  if (isnan(r) && !isnan(x)) error = 33;                    // EDOM
  else if (isinf(r) && isfinite(x)) error = 34;             // ERANGE (overflow / pole)
  return r;
}
__device__ __forceinline__ double exp_chk(double x, int &error)
{
  const double r = exp(x);
  if (isfinite(x) && (isinf(r) || r < 2.2250738585072014e-308)) error = 34;   // overflow / underflow
  return r;
}
__device__ __forceinline__ double pow_chk(double x, double y, int &error)
{
  const double r = pow(x, y);
  if (isnan(r) && !isnan(x) && !isnan(y)) error = 33;
  else if (isinf(r) && isfinite(x) && isfinite(y)) error = 34;
  else if (r == 0.0 && x != 0.0 && isfinite(x) && isfinite(y)) error = 34;
  return r;
}

// tape shapes with a specialised evaluator (same operations in the same order as the interpreter, node intervals in
// registers instead of local memory, no opcode dispatch)
constexpr int kShapeGeneric = 0;
constexpr int kShapeBilinear = 1;   // [Var, Var, Mult(a, b)]                              x_i * x_j (+ linear part)
constexpr int kShapeSumSq = 2;      // [Var, Var, Sqr(a), Sqr(b), SumList(2, 3)]           x_i^2 + x_j^2

struct TapeView {
  const uint8_t *op; const int32_t *a0; const int32_t *a1; const double *cn; const int32_t *child;
  int nn;
};

// One warp's staging area for a BATCH of consecutive constraints.  All lanes of a warp walk the same tapes, so the
// nodes, linear parts and bounds of up to 31 constraints are fetched ONCE, cooperatively (coalesced requests), the
// interpreter reads them from shared memory, and the {lb,ub} segments of every variable the batch touches are
// prefetched into L2 before the first constraint is evaluated: a constraint costs L2 hits instead of a chain of
// DRAM round trips.
constexpr int kStageNodes = 128;
constexpr int kStageChild = 64;
constexpr int kStageLin = 64;
struct BatchStage {
  double cn[kStageNodes];
  double lin_val[kStageLin];
  double c_lb[32], c_ub[32];
  int32_t a0[kStageNodes], a1[kStageNodes];
  int32_t child[kStageChild];
  int32_t lin_col[kStageLin];
  int32_t tape_off[33], lin_off[33];     // offsets of the batch's constraints inside the staged arrays
  uint8_t op[kStageNodes];
  uint8_t shape[32], nlead[32];          // per staged constraint: kShape*, leading OpVar nodes (classified once per batch)
};

// a staged constraint: tape view (pointers into the BatchStage), linear part, bounds
struct ConsView {
  TapeView t;
  const int32_t *lin_col; const double *lin_val;   // [n_lin]
  int n_lin;
  int n_lead_vars;          // leading OpVar nodes (the tape starts with the variable nodes, ascending id)
  double c_lb, c_ub;
  int shape;                // kShape*: tapes of the two dominant forms are evaluated by straight-line code
};


struct BatchInfo {
  int n;                    // constraints staged (1..32)
  const int32_t *child;     // child list base to index with the tape's absolute offsets
  const int32_t *lin_col; const double *lin_val; int lin_base;   // linear terms: staged arrays (lin_base = first term)
};

__device__ __forceinline__ void prefetch_segment(const double2 *tile_base, int64_t ld, int var)
{
  const char *p = reinterpret_cast<const char *>(tile_base + (int64_t)var * ld);
#pragma unroll
  for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128 * k));
}

// Stages constraints c0, c0+1, ... (at most 31, at most c_end - c0, as many as fit the staging arrays) and prefetches
// their variables' segments.  Convergent.
__device__ __forceinline__ BatchInfo stage_batch(const NlDev &N, int c0, int c_end, BatchStage &S, const double2 *tile_base,
                                                 int64_t ld, int lane)
{
  const int avail = min(31, c_end - c0);               // lane `avail` carries the end markers
  // heads: lane l holds the offsets of constraint c0 + l (lane `avail` the end markers)
  int tp = 0, lp = 0;
  if (lane <= avail) { tp = __ldg(N.tape_ptr + c0 + lane); lp = __ldg(N.lin_ptr + c0 + lane); }
  double clb = 0.0, cub = 0.0;
  if (lane < avail) { clb = __ldg(N.c_lb + c0 + lane); cub = __ldg(N.c_ub + c0 + lane); }
  const int tp0 = __shfl_sync(0xffffffffu, tp, 0), lp0 = __shfl_sync(0xffffffffu, lp, 0);
  // the largest k with nodes(c0..c0+k) and lin terms within capacity; a single constraint always fits its nodes
  // (kMaxTape <= kStageNodes); its linear part may not, then it is read from global memory
  const bool fits = lane >= 1 && lane <= avail && tp - tp0 <= kStageNodes && lp - lp0 <= kStageLin;
  const unsigned fm = __ballot_sync(0xffffffffu, fits);
  const unsigned gaps = ~(fm | 1u);
  int n = (fm & 2u) ? (gaps ? __ffs(gaps) - 2 : 31) : 1;     // length of the run of set bits starting at bit 1
  if (n > avail) n = avail;
  const int tp_end = __shfl_sync(0xffffffffu, tp, n), lp_end = __shfl_sync(0xffffffffu, lp, n);
  const int nn = tp_end - tp0, nl = lp_end - lp0;
  __syncwarp();                                        // the previous batch has been consumed
  if (lane <= n) { S.tape_off[lane] = tp - tp0; S.lin_off[lane] = lp - lp0; }
  if (lane < n) { S.c_lb[lane] = clb; S.c_ub[lane] = cub; }
  int clo = 0x7fffffff, chi = 0;
  for (int i = lane; i < nn; i += 32) {
    const int op = __ldg(N.op + tp0 + i);
    const int x0 = __ldg(N.arg0 + tp0 + i), x1 = __ldg(N.arg1 + tp0 + i);
    S.op[i] = (uint8_t)op; S.a0[i] = x0; S.a1[i] = x1; S.cn[i] = __ldg(N.cnst + tp0 + i);
    if (op == OpSumList) { clo = min(clo, x0); chi = max(chi, x1); }
    if (op == OpVar) prefetch_segment(tile_base, ld, x0);
  }
  BatchInfo B;
  B.n = n;
  B.lin_col = N.lin_col; B.lin_val = N.lin_val; B.lin_base = 0;       // global fallback: absolute offsets
  if (nl <= kStageLin) {
    for (int q = lane; q < nl; q += 32) {
      const int col = __ldg(N.lin_col + lp0 + q);
      S.lin_col[q] = col; S.lin_val[q] = __ldg(N.lin_val + lp0 + q);
      prefetch_segment(tile_base, ld, col);
    }
    B.lin_col = S.lin_col; B.lin_val = S.lin_val; B.lin_base = lp0;
  }
  clo = __reduce_min_sync(0xffffffffu, clo); chi = __reduce_max_sync(0xffffffffu, chi);
  B.child = N.child;
  if (chi > clo && chi - clo <= kStageChild) {         // the SumList child lists of consecutive constraints are contiguous
    for (int k = lane; k < chi - clo; k += 32) S.child[k] = __ldg(N.child + clo + k);
    B.child = S.child - clo;
  }
  __syncwarp();
  // classify the staged tapes, one constraint per lane (every lane would otherwise repeat this for every constraint)
  if (lane < n) {
    const int o = S.tape_off[lane], nn = S.tape_off[lane + 1] - o;
    const uint8_t *op = S.op + o; const int32_t *a0 = S.a0 + o, *a1 = S.a1 + o;
    int nv = 0;
    while (nv < nn && op[nv] == OpVar) ++nv;
    int shape = kShapeGeneric;
    if (nv == 2 && nn == 3 && op[2] == OpMult && (unsigned)a0[2] < 2u && (unsigned)a1[2] < 2u && a0[2] != a1[2])
      shape = kShapeBilinear;
    else if (nv == 2 && nn == 5 && op[2] == OpSqr && op[3] == OpSqr && op[4] == OpSumList &&
             (unsigned)a0[2] < 2u && (unsigned)a0[3] < 2u && a1[4] - a0[4] == 2 &&
             B.child[a0[4]] == 2 && B.child[a0[4] + 1] == 3)
      shape = kShapeSumSq;
    S.shape[lane] = (uint8_t)shape; S.nlead[lane] = (uint8_t)nv;
  }
  __syncwarp();
  return B;
}

// constraint k of the staged batch (k < B.n); c = its global index.  Warp-uniform.
__device__ __forceinline__ ConsView batch_constraint(const NlDev &N, const BatchStage &S, const BatchInfo &B, int k, int c)
{
  ConsView V;
  const int o = S.tape_off[k];
  V.t.op = S.op + o; V.t.a0 = S.a0 + o; V.t.a1 = S.a1 + o; V.t.cn = S.cn + o; V.t.child = B.child;
  V.t.nn = S.tape_off[k + 1] - o;
  V.n_lin = S.lin_off[k + 1] - S.lin_off[k];
  if (B.lin_col == S.lin_col) { V.lin_col = S.lin_col + S.lin_off[k]; V.lin_val = S.lin_val + S.lin_off[k]; }
  else { const int q0 = __ldg(N.lin_ptr + c); V.lin_col = N.lin_col + q0; V.lin_val = N.lin_val + q0; }
  V.c_lb = S.c_lb[k]; V.c_ub = S.c_ub[k];
  V.n_lead_vars = S.nlead[k];
  V.shape = S.shape[k];
  return V;
}

// CNode::updateBnd for node i.  Constants: OpNum keeps [d,d] (CNode::setVal :1693-1699), OpInt keeps the
// constructor's (-inf,inf) (CGraph.cpp:1238-1245); neither gets the 1e25 clamp (they are not in vq_/dq_).
template <class R>
__device__ __forceinline__ void node_forward(const TapeView &t, int i, double *nlb, double *nub, const double2 *bx,
                                             int64_t ld, int &error)
{
  constexpr bool D = RoundTraits<R>::directed;
  const int op = t.op[i];
  if (op == OpNum) { nlb[i] = nub[i] = t.cn[i]; return; }
  if (op == OpInt) { nlb[i] = -INFINITY; nub[i] = INFINITY; return; }
  double llb = 0, lub = 0, rlb = 0, rub = 0;
  if (op != OpVar && op != OpSumList) {
    llb = nlb[t.a0[i]]; lub = nub[t.a0[i]];
    if (t.a1[i] >= 0) { rlb = nlb[t.a1[i]]; rub = nub[t.a1[i]]; }
  }
  double lb, ub;
  switch (op) {
  case OpAbs:
    if (lub < 0) { lb = -lub; ub = -llb; }
    else if (llb < 0) { if (-llb > lub) { lb = 0.0; ub = -llb; } else { lb = 0.0; ub = lub; } }
    else { lb = llb; ub = lub; }
    break;
  case OpAcos: lb = 0.0; ub = 3.141592653589793; break;
  case OpAsin: case OpAtan: lb = -3.141592653589793 / 2; ub = 3.141592653589793 / 2; break;
  case OpCeil: lb = ceil(llb); ub = ceil(lub); break;
  case OpCos: case OpSin: lb = -1.0; ub = 1.0; break;
  case OpDiv: bounds_on_div<R>(llb, lub, rlb, rub, lb, ub); break;
  case OpExp:
    lb = (llb == -INFINITY) ? 0.0 : widen_lo(exp_chk(llb, error), D);
    ub = (lub == INFINITY) ? INFINITY : widen_hi(exp_chk(lub, error), D);
    if (D && lb < 0.0) lb = 0.0;
    break;
  case OpFloor: lb = floor(llb); ub = floor(lub); break;
  case OpLog:
    lb = (llb <= 0.0) ? -INFINITY : widen_lo(chk(log(llb), llb, error), D);
    ub = widen_hi(chk(log(lub), lub, error), D);
    break;
  case OpLog10:
    lb = (llb <= 0.0) ? -INFINITY : widen_lo(chk(log10(llb), llb, error), D);
    ub = widen_hi(chk(log10(lub), lub, error), D);
    break;
  case OpMinus: lb = R::sub_lo(llb, rub); ub = R::sub_hi(lub, rlb); break;
  case OpMult: bounds_on_product<R>(true, llb, lub, rlb, rub, lb, ub); break;
  case OpNone: return;
  case OpPlus: lb = R::add_lo(llb, rlb); ub = R::add_hi(lub, rub); break;
  case OpSqr: bounds_on_square<R>(llb, lub, lb, ub); break;
  case OpSqrt:
    lb = (llb < 1e-12) ? 0.0 : R::sqrt_lo(llb);
    ub = R::sqrt_hi(lub);
    if (isnan(ub) && !isnan(lub)) error = 33;
    break;
  case OpSumList: {
    double l = 0.0, u = 0.0;
    for (int c = t.a0[i]; c < t.a1[i]; ++c) { l = R::add_lo(l, nlb[t.child[c]]); u = R::add_hi(u, nub[t.child[c]]); }
    lb = l; ub = u;
  } break;
  case OpUMinus: lb = -lub; ub = -llb; break;
  case OpVar: { const double2 b = bx[(int64_t)t.a0[i] * ld]; lb = b.x; ub = b.y; } break;
  default:   // Acosh Asinh Atanh Cosh CPow IntDiv Pow PowK Round Sinh Tan Tanh: "TODO" in the reference
    lb = -INFINITY; ub = INFINITY; break;
  }
  if (lb < -kMinfty) lb = -INFINITY;
  if (ub > kMinfty) ub = INFINITY;
  nlb[i] = lb; nub[i] = ub;
}

// CNode::propBounds_, CNode.cpp:1504-1526 (a NaN trips an assert in the reference -> error)
__device__ __forceinline__ void prop_child(int c, double lb, double ub, double *nlb, double *nub, bool &is_inf,
                                           int &error)
{
  const double etol = 1e-7;
  if (isnan(lb) || isnan(ub)) { error = 9999; return; }
  if (lb < -kMinfty) lb = -INFINITY;
  if (ub > kMinfty) ub = INFINITY;
  if (lb > ub + etol || ub < nlb[c] - etol || lb > nub[c] + etol) is_inf = true;
  else { if (lb > nlb[c]) nlb[c] = lb; if (ub < nub[c]) nub[c] = ub; }
}

// CNode::propBounds for node i, reference quirks kept (see header)
template <class R>
__device__ __forceinline__ void node_reverse(const TapeView &t, int i, double *nlb, double *nub, bool &is_inf,
                                             int &error)
{
  constexpr bool D = RoundTraits<R>::directed;
  const int op = t.op[i];
  const int l = t.a0[i], r = t.a1[i];
  const double lb_ = nlb[i], ub_ = nub[i];
  double lb = -INFINITY, ub = INFINITY;
  switch (op) {
  case OpAbs: prop_child(l, -ub_, ub_, nlb, nub, is_inf, error); break;
  case OpAcos: case OpAsin: prop_child(l, -1.0, 1.0, nlb, nub, is_inf, error); break;
  case OpCeil: prop_child(l, floor(lb_), floor(ub_), nlb, nub, is_inf, error); break;
  case OpDiv:
    bounds_on_product<R>(false, nlb[r], nub[r], lb_, ub_, lb, ub);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    bounds_on_div<R>(nlb[l], nub[l], lb_, ub_, lb, ub);
    prop_child(r, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpExp:
    lb = widen_lo(chk(log(lb_), lb_, error), D); ub = widen_hi(chk(log(ub_), ub_, error), D);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpFloor: prop_child(l, ceil(lb_), ceil(ub_), nlb, nub, is_inf, error); break;
  case OpLog:
    lb = widen_lo(exp_chk(lb_, error), D); ub = widen_hi(exp_chk(ub_, error), D);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpLog10:
    lb = widen_lo(pow_chk(10.0, lb_, error), D); ub = widen_hi(pow_chk(10.0, ub_, error), D);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpMinus:
    lb = R::add_lo(lb_, nlb[r]); ub = R::add_hi(ub_, nub[r]);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    lb = R::sub_lo(nlb[l], ub_); ub = R::sub_hi(nub[l], lb_);
    prop_child(r, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpMult:
    bounds_on_div<R>(lb_, ub_, nlb[r], nub[r], lb, ub);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    bounds_on_div<R>(lb_, ub_, nlb[l], nub[l], lb, ub);
    prop_child(r, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpPlus:
    lb = R::sub_lo(lb_, nub[r]); ub = R::sub_hi(ub_, nlb[r]);
    prop_child(l, lb, ub, nlb, nub, is_inf, error);
    lb = R::sub_lo(lb_, nub[l]); ub = R::sub_hi(ub_, nlb[l]);
    prop_child(r, lb, ub, nlb, nub, is_inf, error);
    break;
  case OpPowK: {
    const double k = t.cn[r];                       // r_->val_
    if (k > 0) {
      if (is_int_val(k / 2.0)) {
        if (ub_ < -1e-12) error = 3141;
        else {
          ub = widen_hi(pow_chk(ub_, 1.0 / k, error), D); lb = -ub;
          prop_child(l, lb, ub, nlb, nub, is_inf, error);
        }
      } else if (is_int_val((k + 1) / 2.0)) {
        // the reference tests the LOCALS (lb = -inf, ub = +inf), CNode.cpp:1377-1386
        lb = -pow_chk(-lb_, 1.0 / k, error);
        ub = pow_chk(ub_, 1.0 / k, error);
        lb = widen_lo(lb, D); ub = widen_hi(ub, D);
        prop_child(l, lb, ub, nlb, nub, is_inf, error);
      }
    }
  } break;
  case OpSqr:    // local ub = +inf: a no-op in the reference, CNode.cpp:1399-1403
    prop_child(l, -INFINITY, INFINITY, nlb, nub, is_inf, error);
    break;
  case OpSqrt:   // local lb = -inf: only child >= 0 is ever derived, CNode.cpp:1404-1412
    if (ub_ < 0.0) is_inf = true;
    else prop_child(l, 0.0, INFINITY, nlb, nub, is_inf, error);
    break;
  case OpSumList: {   // CNode.cpp:1413-1481, including the tub = -inf defect at :1471-1473
    bool inf_lb = false, inf_ub = false;
    const int c0 = t.a0[i], c1 = t.a1[i];
    lb = 0.0;
    for (int c = c0; c < c1; ++c) {
      const double cl = nlb[t.child[c]];
      if (cl > -INFINITY) lb = R::add_lo(lb, cl); else if (inf_lb) { lb = -INFINITY; break; } else inf_lb = true;
    }
    ub = 0.0;
    for (int c = c0; c < c1; ++c) {
      const double cu = nub[t.child[c]];
      if (cu < INFINITY) ub = R::add_hi(ub, cu); else if (inf_ub) { ub = INFINITY; break; } else inf_ub = true;
    }
    if (lb > -INFINITY || ub < INFINITY) {
      for (int c = c0; c < c1; ++c) {
        const int ch = t.child[c];
        double tlb, tub;
        if (ub < INFINITY) {
          if (!inf_ub) tlb = R::sub_lo(lb_, R::sub_hi(ub, nub[ch]));
          else if (nub[ch] < INFINITY) tlb = -INFINITY;
          else tlb = R::sub_lo(lb_, ub);
        } else tlb = -INFINITY;
        if (lb > -INFINITY) {
          if (!inf_lb) tub = R::sub_hi(ub_, R::sub_lo(lb, nlb[ch]));
          else if (nlb[ch] > -INFINITY) tub = INFINITY;
          else tub = R::sub_hi(ub_, lb);
        } else tub = -INFINITY;
        prop_child(ch, tlb, tub, nlb, nub, is_inf, error);
        if (is_inf) break;
      }
    }
  } break;
  case OpUMinus: prop_child(l, -ub_, -lb_, nlb, nub, is_inf, error); break;
  default: break;      // unimplemented in the reference: no-op
  }
}

__device__ __forceinline__ bool is_leaf_op(int op) { return op == OpVar || op == OpNum || op == OpInt; }

// LinearFunction::computeBounds of the constraint's linear part (LinearFunction.cpp:178-195)
template <class R>
__device__ __forceinline__ void lin_part_bounds(const ConsView &V, const double2 *bx, int64_t ld, double &lo, double &up)
{
  lo = 0.0; up = 0.0;
  for (int q0 = 0; q0 < V.n_lin; q0 += 2) {             // two gathers in flight
    const double a0 = V.lin_val[q0];
    const double2 b0 = bx[(int64_t)V.lin_col[q0] * ld];
    const bool two = q0 + 1 < V.n_lin;
    const double a1 = two ? V.lin_val[q0 + 1] : 0.0;
    const double2 b1 = two ? bx[(int64_t)V.lin_col[q0 + 1] * ld] : make_double2(0.0, 0.0);
    if (a0 > 0) { lo = R::add_lo(lo, R::mul_lo(a0, b0.x)); up = R::add_hi(up, R::mul_hi(a0, b0.y)); }
    else        { lo = R::add_lo(lo, R::mul_lo(a0, b0.y)); up = R::add_hi(up, R::mul_hi(a0, b0.x)); }
    if (two) {
      if (a1 > 0) { lo = R::add_lo(lo, R::mul_lo(a1, b1.x)); up = R::add_hi(up, R::mul_hi(a1, b1.y)); }
      else        { lo = R::add_lo(lo, R::mul_lo(a1, b1.y)); up = R::add_hi(up, R::mul_hi(a1, b1.x)); }
    }
  }
}

// CGraph::computeBounds forward sweep.  The leading variable nodes are gathered four at a time (independent
// requests in flight together) before the operator nodes are interpreted.
template <class R>
__device__ __forceinline__ void tape_forward(const ConsView &V, double *nlb, double *nub, const double2 *bx, int64_t ld,
                                             int &error)
{
  const TapeView &t = V.t;
  for (int i0 = 0; i0 < V.n_lead_vars; i0 += 4) {
    double2 b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i0 + u < V.n_lead_vars) b[u] = bx[(int64_t)t.a0[i0 + u] * ld];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < V.n_lead_vars) {
        nlb[i0 + u] = (b[u].x < -kMinfty) ? -INFINITY : b[u].x;        // the clamp of updateBnd, :1898-1903
        nub[i0 + u] = (b[u].y > kMinfty) ? INFINITY : b[u].y;
      }
  }
  for (int i = V.n_lead_vars; i < t.nn; ++i) node_forward<R>(t, i, nlb, nub, bx, ld, error);
}

// ---- straight-line evaluators of the two dominant tape shapes.  Every arithmetic operation, comparison and store is
//      the interpreter's (node_forward / node_reverse / prop_child above), in the same order: results are bitwise the
//      same; node intervals live in registers. ----
__device__ __forceinline__ void clamp25(double &lb, double &ub)      // the clamp of CNode::updateBnd, :1898-1903
{
  if (lb < -kMinfty) lb = -INFINITY;
  if (ub > kMinfty) ub = INFINITY;
}
// CNode::propBounds_ on an interval held in registers
__device__ __forceinline__ void prop_reg(double lb, double ub, double &nl, double &nu, bool &is_inf, int &error)
{
  const double etol = 1e-7;
  if (isnan(lb) || isnan(ub)) { error = 9999; return; }
  if (lb < -kMinfty) lb = -INFINITY;
  if (ub > kMinfty) ub = INFINITY;
  if (lb > ub + etol || ub < nl - etol || lb > nu + etol) is_inf = true;
  else { if (lb > nl) nl = lb; if (ub < nu) nu = ub; }
}

// forward sweep of [Var, Var, Mult]: leaves clamped, product of node a0 and node a1
template <class R>
__device__ __forceinline__ void bilinear_forward(const TapeView &t, double2 b0, double2 b1, double nl[2], double nu[2],
                                                 double &ol, double &ou)
{
  nl[0] = (b0.x < -kMinfty) ? -INFINITY : b0.x; nu[0] = (b0.y > kMinfty) ? INFINITY : b0.y;
  nl[1] = (b1.x < -kMinfty) ? -INFINITY : b1.x; nu[1] = (b1.y > kMinfty) ? INFINITY : b1.y;
  const bool swap = t.a0[2] != 0;            // warp-uniform: which variable node is the left operand
  bounds_on_product<R>(true, swap ? nl[1] : nl[0], swap ? nu[1] : nu[0], swap ? nl[0] : nl[1], swap ? nu[0] : nu[1], ol, ou);
  clamp25(ol, ou);
}

template <class R>
__device__ __forceinline__ void sumsq_forward(const TapeView &t, double2 b0, double2 b1, double nl[2], double nu[2],
                                              double sl[2], double su[2], double &ol, double &ou)
{
  nl[0] = (b0.x < -kMinfty) ? -INFINITY : b0.x; nu[0] = (b0.y > kMinfty) ? INFINITY : b0.y;
  nl[1] = (b1.x < -kMinfty) ? -INFINITY : b1.x; nu[1] = (b1.y > kMinfty) ? INFINITY : b1.y;
  const bool s2 = t.a0[2] != 0, s3 = t.a0[3] != 0;
  bounds_on_square<R>(s2 ? nl[1] : nl[0], s2 ? nu[1] : nu[0], sl[0], su[0]); clamp25(sl[0], su[0]);
  bounds_on_square<R>(s3 ? nl[1] : nl[0], s3 ? nu[1] : nu[0], sl[1], su[1]); clamp25(sl[1], su[1]);
  ol = R::add_lo(R::add_lo(0.0, sl[0]), sl[1]);
  ou = R::add_hi(R::add_hi(0.0, su[0]), su[1]);
  clamp25(ol, ou);
}

// NlPresHandler::chkRed_ for one constraint (NlPresHandler.cpp:101-208, nlf branch):
// returns 0 ok, 3 infeasible, 4 evaluation error
// SHAPED: every tape of the problem has one of the two straight-line shapes (classified at load time): the interpreter
// and its per-thread node arrays are not compiled in (fewer registers, no spills: 12-warp CTAs at 80 registers spill
// 400 bytes with the interpreter)
template <class R, bool SHAPED = false>
__device__ __forceinline__ int nl_chk_red(const ConsView &V, const double2 *bx, int64_t ld, double *nlb, double *nub)
{
  const TapeView &t = V.t;
  if (SHAPED || V.shape != kShapeGeneric) {
    const double2 b0 = bx[(int64_t)t.a0[0] * ld], b1 = bx[(int64_t)t.a0[1] * ld];     // both gathers in flight
    double nl[2], nu[2], ol, ou;
    if (V.shape == kShapeBilinear) bilinear_forward<R>(t, b0, b1, nl, nu, ol, ou);
    else { double sl[2], su[2]; sumsq_forward<R>(t, b0, b1, nl, nu, sl, su, ol, ou); }
    double lfl, lfu;
    lin_part_bounds<R>(V, bx, ld, lfl, lfu);
    const double impl_lb = R::add_lo(ol, lfl), impl_ub = R::add_hi(ou, lfu);
    return (impl_ub + 1e-6 < V.c_lb || impl_lb - 1e-6 > V.c_ub) ? 3 : 0;
  }
  if constexpr (SHAPED) return 4;
  int error = 0;
  tape_forward<R>(V, nlb, nub, bx, ld, error);
  if (error != 0) return 4;
  double lfl, lfu;
  lin_part_bounds<R>(V, bx, ld, lfl, lfu);
  const double impl_lb = R::add_lo(nlb[t.nn - 1], lfl), impl_ub = R::add_hi(nub[t.nn - 1], lfu);
  if (impl_ub + 1e-6 < V.c_lb || impl_lb - 1e-6 > V.c_ub) return 3;
  return 0;
}

// NlPresHandler::chkRed_ for one QuadraticFunction constraint (NlPresHandler.cpp:127-150): QuadraticFunction::computeBounds
// (QuadraticFunction.cpp:156-180) -- per term the four corner products (coef * x1) * x2, evaluated left to right, their min
// added to the lower and their max to the upper bound -- plus the linear part.  With RoundNearest these are the
// reference's operations in the reference's order; with RoundDirected every corner is an interval product, so the lower
// bound only errs downward and the upper bound only upward.  returns 0 ok, 3 infeasible
template <class R>
__device__ __forceinline__ int quad_chk_red(const NlDev &N, int q, const double2 *bx, int64_t ld)
{
  double lb = 0.0, ub = 0.0;
  const int tb = __ldg(N.q_ptr + q), te = __ldg(N.q_ptr + q + 1);
  for (int t = tb; t < te; ++t) {
    const double w = __ldg(N.q_coef + t);
    const double2 b1 = bx[(int64_t)__ldg(N.q_v1 + t) * ld], b2 = bx[(int64_t)__ldg(N.q_v2 + t) * ld];
    // (w * x1) as an interval per end point of x1, then times the end points of x2
    const double wl_lo = R::mul_lo(w, b1.x), wl_hi = R::mul_hi(w, b1.x), wu_lo = R::mul_lo(w, b1.y), wu_hi = R::mul_hi(w, b1.y);
    const double a_lo = std_min(R::mul_lo(wl_lo, b2.x), R::mul_lo(wl_hi, b2.x)), a_hi = std_max(R::mul_hi(wl_lo, b2.x), R::mul_hi(wl_hi, b2.x));
    const double b_lo = std_min(R::mul_lo(wl_lo, b2.y), R::mul_lo(wl_hi, b2.y)), b_hi = std_max(R::mul_hi(wl_lo, b2.y), R::mul_hi(wl_hi, b2.y));
    const double c_lo = std_min(R::mul_lo(wu_lo, b2.x), R::mul_lo(wu_hi, b2.x)), c_hi = std_max(R::mul_hi(wu_lo, b2.x), R::mul_hi(wu_hi, b2.x));
    const double d_lo = std_min(R::mul_lo(wu_lo, b2.y), R::mul_lo(wu_hi, b2.y)), d_hi = std_max(R::mul_hi(wu_lo, b2.y), R::mul_hi(wu_hi, b2.y));
    double m = std_min(a_lo, b_lo); m = std_min(m, c_lo); m = std_min(m, d_lo);
    lb = R::add_lo(lb, m);
    m = std_max(a_hi, b_hi); m = std_max(m, c_hi); m = std_max(m, d_hi);
    ub = R::add_hi(ub, m);
  }
  double lfl = 0.0, lfu = 0.0;
  for (int k = __ldg(N.q_lin_ptr + q); k < __ldg(N.q_lin_ptr + q + 1); ++k) {      // LinearFunction::computeBounds
    const double a = __ldg(N.q_lin_val + k);
    const double2 b = bx[(int64_t)__ldg(N.q_lin_col + k) * ld];
    if (a > 0) { lfl = R::add_lo(lfl, R::mul_lo(a, b.x)); lfu = R::add_hi(lfu, R::mul_hi(a, b.y)); }
    else       { lfl = R::add_lo(lfl, R::mul_lo(a, b.y)); lfu = R::add_hi(lfu, R::mul_hi(a, b.x)); }
  }
  const double impl_lb = R::add_lo(lb, lfl), impl_ub = R::add_hi(ub, lfu);
  return (impl_ub + 1e-6 < __ldg(N.q_lb + q) || impl_lb - 1e-6 > __ldg(N.q_ub + q)) ? 3 : 0;
}

// NlPresHandler::varBndsFromCons_ for one constraint (NlPresHandler.cpp:1771-1803) = lf bounds +
// CGraph::varBoundMods + in-place application of the mods.  returns 0 ok, 3 infeasible, 4 error;
// n_mods receives the number of bound changes.
template <class R, bool SHAPED = false>
__device__ __forceinline__ int nl_var_bound_mods(const ConsView &V, double2 *bx, int64_t ld, double *nlb,
                                                 double *nub, int &n_mods, unsigned &moved_int)
{
  const double bslack = 1e-5, bslack10 = 1e-4;
  const TapeView &t = V.t;
  double lfl, lfu;
  lin_part_bounds<R>(V, bx, ld, lfl, lfu);
  const double ub_in = R::sub_hi(V.c_ub, lfl), lb_in = R::sub_lo(V.c_lb, lfu);
  int error = 0;
  if (SHAPED || V.shape != kShapeGeneric) {
    double2 *p0 = bx + (int64_t)t.a0[0] * ld, *p1 = bx + (int64_t)t.a0[1] * ld;
    double2 b0 = *p0, b1 = *p1;
    double nl[2], nu[2], ol, ou;
    bool is_inf = false;
    if (V.shape == kShapeBilinear) {
      bilinear_forward<R>(t, b0, b1, nl, nu, ol, ou);
      ol = fmax(lb_in, ol); ou = fmin(ub_in, ou);
      // reverse rule of OpMult (CNode.cpp:1344-1349): left child from out / right, then right child from out / the
      // ALREADY TIGHTENED left child
      const bool swap = t.a0[2] != 0;            // warp-uniform: node 1 is the left operand
      double Ll = swap ? nl[1] : nl[0], Lu = swap ? nu[1] : nu[0], Rl = swap ? nl[0] : nl[1], Ru = swap ? nu[0] : nu[1];
      double lb, ub;
      bounds_on_div<R>(ol, ou, Rl, Ru, lb, ub);
      prop_reg(lb, ub, Ll, Lu, is_inf, error);
      bounds_on_div<R>(ol, ou, Ll, Lu, lb, ub);
      prop_reg(lb, ub, Rl, Ru, is_inf, error);
      if (is_inf) return 3;
      if (error > 0) return 4;
      nl[0] = swap ? Rl : Ll; nu[0] = swap ? Ru : Lu; nl[1] = swap ? Ll : Rl; nu[1] = swap ? Lu : Ru;
    } else {
      double sl[2], su[2];
      sumsq_forward<R>(t, b0, b1, nl, nu, sl, su, ol, ou);
      ol = fmax(lb_in, ol); ou = fmin(ub_in, ou);
      // reverse rule of OpSumList over the two squares (CNode.cpp:1413-1481, its tub = -inf defect included); the
      // reverse rule of OpSqr is a no-op in the reference (:1399-1403), so the variable nodes never move
      bool inf_lb = false, inf_ub = false;
      double lb = 0.0, ub = 0.0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (sl[c] > -INFINITY) lb = R::add_lo(lb, sl[c]); else if (inf_lb) { lb = -INFINITY; break; } else inf_lb = true;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (su[c] < INFINITY) ub = R::add_hi(ub, su[c]); else if (inf_ub) { ub = INFINITY; break; } else inf_ub = true;
      }
      if (lb > -INFINITY || ub < INFINITY) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double tlb, tub;
          if (ub < INFINITY) {
            if (!inf_ub) tlb = R::sub_lo(ol, R::sub_hi(ub, su[c]));
            else if (su[c] < INFINITY) tlb = -INFINITY;
            else tlb = R::sub_lo(ol, ub);
          } else tlb = -INFINITY;
          if (lb > -INFINITY) {
            if (!inf_lb) tub = R::sub_hi(ou, R::sub_lo(lb, sl[c]));
            else if (sl[c] > -INFINITY) tub = INFINITY;
            else tub = R::sub_hi(ou, lb);
          } else tub = -INFINITY;
          prop_reg(tlb, tub, sl[c], su[c], is_inf, error);
          if (is_inf) break;
        }
      }
      if (is_inf) return 3;
      if (error > 0) return 4;
    }
    // harvest (CGraph.cpp:1627-1641): variable nodes in ascending id; both tests read the variable's bounds before
    // any mod of this call is applied (the two variables are distinct)
    bool ch0 = false, ch1 = false;
    const double2 o0 = b0, o1 = b1;
    if (nl[0] > o0.x + bslack10) { b0.x = nl[0] - bslack; ++n_mods; ch0 = true; }
    if (nu[0] < o0.y - bslack10) { b0.y = nu[0] + bslack; ++n_mods; ch0 = true; }
    if (ch0) *p0 = b0;
    if (nl[1] > o1.x + bslack10) { b1.x = nl[1] - bslack; ++n_mods; ch1 = true; }
    if (nu[1] < o1.y - bslack10) { b1.y = nu[1] + bslack; ++n_mods; ch1 = true; }
    if (ch1) *p1 = b1;
    return 0;
  }
  if constexpr (SHAPED) return 4;
  tape_forward<R>(V, nlb, nub, bx, ld, error);
  if (error > 0) return 4;
  const int o = t.nn - 1;
  nlb[o] = fmax(lb_in, nlb[o]); nub[o] = fmin(ub_in, nub[o]);
  bool is_inf = false;
  for (int i = t.nn - 1; i >= 0; --i) {
    if (is_leaf_op(t.op[i])) continue;               // leaves are not in dq_
    node_reverse<R>(t, i, nlb, nub, is_inf, error);
    if (is_inf) return 3;
    if (error > 0) return 4;
  }
  (void)moved_int;
  for (int i = 0; i < t.nn; ++i) {
    if (t.op[i] != OpVar) continue;
    double2 *pb = bx + (int64_t)t.a0[i] * ld;
    double2 b = *pb;
    bool ch = false;
    // both tests read the variable's bounds before any mod of this call is applied
    if (nlb[i] > b.x + bslack10) { b.x = nlb[i] - bslack; ++n_mods; ch = true; }
    if (nub[i] < pb->y - bslack10) { b.y = nub[i] + bslack; ++n_mods; ch = true; }
    if (ch) *pb = b;
  }
  return 0;
}

}  // namespace mntr
