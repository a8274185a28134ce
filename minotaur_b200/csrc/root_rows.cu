// root_rows.cu -- row-parallel pieces of LinearHandler::presolve (the ROOT presolve) over the CSR already in HBM
// (SURVEY.md 8f-3).  Only the data-parallel detection runs here; what mutates Minotaur's object graph (merging two
// duplicate rows, deleting a row) stays on the host, driven by the lists these kernels return.
//
//   dupRows_ (LinearHandler.cpp:882-949): every linear row is hashed with two random vectors (h = row . r, terms added
//     in ascending column order exactly like Constraint::getActivity / LinearFunction::eval, LinearFunction.cpp:151-158),
//     then ALL pairs i < j are compared: |h1j - h1i| < 1e-10 or |h1j + h1i| < 1e-10 (same row up to sign), else
//     |h1i/h1j - h2i/h2j| < 1e-10 (a multiple).  The reference does this with an O(m^2) loop on one core; here one thread
//     owns a row i and streams the hashes of the rows j > i through shared memory.  The pairs that pass are the
//     CANDIDATES the reference hands to treatDupRows_ -- the same tests on the same numbers, so the same list.
//   redundancy (linBndTighten_ with apply_to_prob, :974-985): a row whose activity range [ll, uu] lies inside its bounds
//     (with eTol) is redundant.
#include "device_problem.cuh"
#include "kernels.h"

namespace mntr {

namespace {

constexpr int kPairThreads = 256;

// h1[i], h2[i] of every stored row q (i = perm[q], the caller's row index); deleted rows get 1e30 like the reference's
// non-linear constraints
__global__ void row_hash_kernel(LinDev P, const int32_t *__restrict__ perm, const double *__restrict__ r1,
                                const double *__restrict__ r2, double *h1, double *h2)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.m) return;
  const int2 info = __ldg(P.row_info + q);
  double a = 1e30, b = 1e30;
  if (info.y >= 0) {
    a = 0.0; b = 0.0;
    for (int t = info.x; t < info.x + info.y; ++t) {
      const int j = __ldg(P.col + t);
      const double v = __ldg(P.val + t);
      a = __dadd_rn(a, __dmul_rn(__ldg(r1 + j), v));
      b = __dadd_rn(b, __dmul_rn(__ldg(r2 + j), v));
    }
  }
  const int i = __ldg(perm + q);
  h1[i] = a; h2[i] = b;
}

// all pairs i < j; kind 1 = same / negated row (mult 1.0), kind 2 = a multiple (mult h1i / h1j)
__global__ void __launch_bounds__(kPairThreads)
dup_pairs_kernel(int m, const double *__restrict__ h1, const double *__restrict__ h2, long long cap, int32_t *pair_i,
                 int32_t *pair_j, uint8_t *pair_kind, unsigned long long *count)
{
  __shared__ double s1[kPairThreads], s2[kPairThreads];
  const int i = blockIdx.x * kPairThreads + threadIdx.x;
  const bool live = i < m;
  const double a1 = live ? h1[i] : 1e30, a2 = live ? h2[i] : 1e30;
  const bool ok_i = live && a1 < 1e29;
  // tiles of j; tile jb covers [jb * T, jb * T + T): only tiles that hold some j > the block's first i
  for (int jb = blockIdx.x; jb * kPairThreads < m; ++jb) {
    const int j0 = jb * kPairThreads;
    __syncthreads();
    s1[threadIdx.x] = (j0 + threadIdx.x < m) ? h1[j0 + threadIdx.x] : 1e30;
    s2[threadIdx.x] = (j0 + threadIdx.x < m) ? h2[j0 + threadIdx.x] : 1e30;
    __syncthreads();
    if (!ok_i) continue;
    const int jn = min(kPairThreads, m - j0);
    for (int t = 0; t < jn; ++t) {
      const int j = j0 + t;
      if (j <= i) continue;
      const double b1 = s1[t], b2 = s2[t];
      int kind = 0;
      if (fabs(b1 - a1) < 1e-10 || fabs(b1 + a1) < 1e-10) kind = 1;
      else if (b1 < 1e29) {
        // |a1/b1 - a2/b2| < 1e-10 needs two divisions: a cross-multiplied form with a generous margin rules almost
        // every pair out first; the survivors take the reference's exact test
        const double cross = fabs(a1 * b2 - a2 * b1);
        const double lim = 2e-10 * fabs(b1 * b2) + 1e-13 * (fabs(a1 * b2) + fabs(a2 * b1));
        if (!(cross > lim) && fabs(__ddiv_rn(a1, b1) - __ddiv_rn(a2, b2)) < 1e-10) kind = 2;
      }
      if (kind) {
        const unsigned long long at = atomicAdd(count, 1ull);
        if ((long long)at < cap) { pair_i[at] = i; pair_j[at] = j; pair_kind[at] = (uint8_t)kind; }
      }
    }
  }
}

// getLfBnds_ (LinearHandler.cpp:1237-1258) of every row on one box, round to nearest like the reference; redundant iff
// ll >= lb - eTol && uu <= ub + eTol (:974)
__global__ void redundant_rows_kernel(LinDev P, const int32_t *__restrict__ perm, const double *__restrict__ lb,
                                      const double *__restrict__ ub, uint8_t *flag, unsigned long long *count)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.m) return;
  const int2 info = __ldg(P.row_info + q);
  const int i = __ldg(perm + q);
  uint8_t f = 0;
  if (info.y >= 0) {
    double ll = 0.0, uu = 0.0;
    for (int t = info.x; t < info.x + info.y; ++t) {
      const int j = __ldg(P.col + t);
      const double a = __ldg(P.val + t), l = __ldg(lb + j), u = __ldg(ub + j);
      if (a > 0) { ll = __dadd_rn(ll, __dmul_rn(a, l)); uu = __dadd_rn(uu, __dmul_rn(a, u)); }
      else       { ll = __dadd_rn(ll, __dmul_rn(a, u)); uu = __dadd_rn(uu, __dmul_rn(a, l)); }
    }
    const double2 bnd = __ldg(P.row_bnd + q);
    f = (ll >= bnd.x - kETol && uu <= bnd.y + kETol) ? 1 : 0;
  }
  flag[i] = f;
  if (f) atomicAdd(count, 1ull);
}


// ---- LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) + computeImpBounds_ (:707-783) -----------------------
// One thread per row of the level: the row's work is the reference's, term by term in variable order, the first
// binary that can be improved ends the row.  Terms with |a| <= 1e-9 do not exist (LinearFunction::addTerm :89-95).
constexpr int kImplicTerms = 50;       // rows with fewer terms use the implications (:625-629)
__device__ __forceinline__ bool ci_term(double a) { return fabs(a) > 1e-9; }

// weights of z and v in the 2-term row c2 (0: absent); the version of c2 the row c may see
__device__ __forceinline__ void ci_pair(const CoeffProb &Q, int c, int c2, int z, int v, double &b2, double &a2, double &clb, double &cub)
{
  const bool cur = c2 < c;             // improved earlier in the pass: its current state; else as it came
  const double *V = cur ? Q.val : Q.val0;
  b2 = 0.0; a2 = 0.0;
  for (int t = Q.row_ptr[c2]; t < Q.row_ptr[c2 + 1]; ++t) {
    const double a = V[t];
    if (!ci_term(a)) continue;
    const int j = Q.col[t];
    if (j == z) b2 = a; else if (j == v) a2 = a;
  }
  clb = cur ? Q.rlb[c2] : Q.rlb0[c2];
  cub = cur ? Q.rub[c2] : Q.rub0[c2];
}

// computeImpBounds_: activity of row c with z at zval and every other variable tightened by the 2-term rows it shares
// with z; then LinearFunction::computeBounds (LinearFunction.cpp:178-195)
__device__ void ci_imp_bounds(const CoeffProb &Q, int c, int z, double zval, double &out_l, double &out_u)
{
  double lo = 0.0, up = 0.0;
  for (int t = Q.row_ptr[c]; t < Q.row_ptr[c + 1]; ++t) {
    const double a = Q.val[t];
    if (!ci_term(a)) continue;
    const int v = Q.col[t];
    double l1 = Q.lb[v], u1 = Q.ub[v];
    if (v == z) {
      if (zval < 0.5) u1 = 0.0; else l1 = 1.0;
    } else {
      const double vl = l1, vu = u1;
      for (int q = Q.cptr[v]; q < Q.cptr[v + 1]; ++q) {
        const int c2 = Q.crow[q];
        if (!Q.is2[c2]) continue;
        double b2, a2, clb, cub;
        ci_pair(Q, c, c2, z, v, b2, a2, clb, cub);
        if (b2 == 0.0 || a2 == 0.0) continue;
        if (a2 > 0 && (cub - zval * b2) / a2 < u1) u1 = (cub - zval * b2) / a2;
        if (a2 < 0 && (cub - zval * b2) / a2 > l1) l1 = (cub - zval * b2) / a2;
        if (a2 > 0 && (clb - zval * b2) / a2 > l1) l1 = (clb - zval * b2) / a2;
        if (a2 < 0 && (clb - zval * b2) / a2 < u1) u1 = (clb - zval * b2) / a2;
      }
      if (!(l1 > vl)) l1 = vl;
      if (!(u1 < vu)) u1 = vu;
    }
    if (a > 0) { lo += a * l1; up += a * u1; }
    else       { lo += a * u1; up += a * l1; }
  }
  out_l = lo; out_u = up;
}

__global__ void coeff_imp_kernel(CoeffProb Q, const int32_t *rows, int32_t n_rows, long long cap, int32_t *out_row,
                                 int32_t *out_var, double *out_coef, int32_t *out_side, double *out_bnd, double *out_delta,
                                 unsigned long long *count, int32_t *n_erased)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int c = rows[i];
  const double coeftol = 1e-4, bslack = 1e-4;
  const double lb = Q.rlb[c], ub = Q.rub[c];
  int nt = 0;
  double ll = 0.0, uu = 0.0;                        // getLfBnds_ :1237-1258
  for (int t = Q.row_ptr[c]; t < Q.row_ptr[c + 1]; ++t) {
    const double a = Q.val[t];
    if (!ci_term(a)) continue;
    ++nt;
    const int j = Q.col[t];
    if (a > 0) { ll += a * Q.lb[j]; uu += a * Q.ub[j]; }
    else       { ll += a * Q.ub[j]; uu += a * Q.lb[j]; }
  }
  if (nt < 2) return;
  const bool implic = nt < kImplicTerms;
  for (int t = Q.row_ptr[c]; t < Q.row_ptr[c + 1]; ++t) {
    const double a0 = Q.val[t];
    if (!ci_term(a0)) continue;
    const int v = Q.col[t];
    const int ty = Q.var_type[v];
    if (!((ty == 0 || ty == 2) && Q.ub[v] > Q.lb[v] + 0.5)) continue;      // Binary / ImplBin, not fixed (:635-637)
    double delta = 0.0, nb = 0.0;
    int side = 0;
    bool hit = false;
    if (implic) {
      ci_imp_bounds(Q, c, v, 1.0, ll, uu);
      ll -= bslack; uu += bslack;
      if (a0 > 0) ll -= a0; else uu -= a0;
    }
    if (uu + a0 < ub - coeftol && uu >= ub) { delta = ub - uu - a0; hit = true; }
    else if (ll + a0 > lb + coeftol && ll <= lb) { delta = lb - ll - a0; hit = true; }
    if (!hit) {
      if (implic) {
        ci_imp_bounds(Q, c, v, 0.0, ll, uu);
        if (a0 > 0) uu += a0; else ll += a0;
      }
      if (uu - a0 < ub - coeftol && uu >= ub) { delta = uu - a0 - ub; side = 2; nb = uu - a0; hit = true; }
      else if (ll - a0 > lb + coeftol && ll <= lb) { delta = ll - a0 - lb; side = 1; nb = ll - a0; hit = true; }
    }
    if (!hit) continue;
    double nv = a0;                                  // LinearFunction::incTerm :133-142
    if (fabs(delta) > 1e-9) { nv = a0 + delta; if (fabs(nv) < 1e-9) { nv = 0.0; atomicAdd(n_erased, 1); } }
    Q.val[t] = nv;
    if (side == 2) Q.rub[c] = nb; else if (side == 1) Q.rlb[c] = nb;
    const unsigned long long k = atomicAdd(count, 1ull);
    if ((long long)k < cap) { out_row[k] = c; out_var[k] = v; out_coef[k] = nv; out_side[k] = side; out_bnd[k] = nb; out_delta[k] = delta; }
    return;
  }
}

}  // namespace

cudaError_t launch_row_hash(const LinDev &P, const int32_t *perm, const double *r1, const double *r2, double *h1, double *h2,
                            cudaStream_t stream)
{
  if (P.m <= 0) return cudaSuccess;
  row_hash_kernel<<<(P.m + 255) / 256, 256, 0, stream>>>(P, perm, r1, r2, h1, h2);
  return cudaGetLastError();
}

cudaError_t launch_dup_pairs(int m, const double *h1, const double *h2, long long cap, int32_t *pair_i, int32_t *pair_j,
                             uint8_t *pair_kind, unsigned long long *count, cudaStream_t stream)
{
  if (m <= 1) return cudaSuccess;
  dup_pairs_kernel<<<(m + kPairThreads - 1) / kPairThreads, kPairThreads, 0, stream>>>(m, h1, h2, cap, pair_i, pair_j, pair_kind, count);
  return cudaGetLastError();
}

cudaError_t launch_redundant_rows(const LinDev &P, const int32_t *perm, const double *lb, const double *ub, uint8_t *flag,
                                  unsigned long long *count, cudaStream_t stream)
{
  if (P.m <= 0) return cudaSuccess;
  redundant_rows_kernel<<<(P.m + 255) / 256, 256, 0, stream>>>(P, perm, lb, ub, flag, count);
  return cudaGetLastError();
}

cudaError_t launch_coeff_imp(const CoeffProb &Q, const int32_t *rows, int32_t n_rows, long long cap, int32_t *out_row,
                             int32_t *out_var, double *out_coef, int32_t *out_side, double *out_bnd, double *out_delta,
                             unsigned long long *count, int32_t *n_erased, cudaStream_t stream)
{
  if (n_rows <= 0) return cudaSuccess;
  coeff_imp_kernel<<<(n_rows + 127) / 128, 128, 0, stream>>>(Q, rows, n_rows, cap, out_row, out_var, out_coef, out_side, out_bnd,
                                                             out_delta, count, n_erased);
  return cudaGetLastError();
}

}  // namespace mntr
