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

}  // namespace mntr
