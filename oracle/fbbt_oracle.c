/*
 * fbbt_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * Plain-C restatement of the reference's FBBT hot path.  "ref:" comments cite
 * files under /root/reference/src/base/.  Arithmetic is IEEE fp64 round-to-nearest
 * with no FMA contraction (compile with -ffp-contract=off), terms of a row are
 * walked in ascending variable id exactly like the reference's std::map.
 *
 * Pinned against the reference's own objects (oracle/_ref) -- see fbbt_oracle.h.
 */
#include "fbbt_oracle.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ref: LinearHandler.cpp:69-71 */
static const double INT_TOL = 1e-6;
static const double E_TOL   = 1e-8;
static const double INF20   = 1e20;

static int is_int_type(uint8_t t) { return t == ORC_INTEGER || t == ORC_BINARY; }

/* ===========================================================================
 *                               linear rows
 * ========================================================================= */

/* ref: LinearHandler.cpp:1237-1258 getLfBnds_ (== LinearFunction.cpp:178-195) */
static void lf_bnds(int32_t k, const int32_t *col, const double *val, const double *lb,
                    const double *ub, double *lo, double *up)
{
  double l = 0, u = 0;
  for (int32_t t = 0; t < k; ++t) {
    double a = val[t], vlb = lb[col[t]], vub = ub[col[t]];
    if (a > 0) { l += a * vlb; u += a * vub; }
    else       { l += a * vub; u += a * vlb; }
  }
  *lo = l; *up = u;
}

/* ref: LinearHandler.cpp:1261-1319 getSingLfBnds_ -- the flag machine is kept as coded */
static void lf_sing_bnds(int32_t k, const int32_t *col, const double *val, const double *lbv,
                         const double *ubv, double *lo, double *up)
{
  double lb = 0, ub = 0;
  int lo_is_sing = 0, up_is_sing = 0, lo_is_finite = 1, up_is_finite = 1;
  for (int32_t t = 0; t < k; ++t) {
    double coef = val[t], vlb = lbv[col[t]], vub = ubv[col[t]];
    if (coef > E_TOL) {
      if (vub < INF20 && up_is_finite) ub += coef * vub;
      else if (up_is_sing) { up_is_sing = 0; ub = INFINITY; up_is_finite = 0; }
      else up_is_sing = 1;
      if (vlb > -INF20 && lo_is_finite) lb += coef * vlb;
      else if (lo_is_sing) { lo_is_sing = 0; lb = -INFINITY; lo_is_finite = 0; }
      else lo_is_sing = 1;
    } else if (coef < -E_TOL) {
      if (vub < INF20 && lo_is_finite) lb += coef * vub;
      else if (lo_is_sing) { lo_is_sing = 0; lb = -INFINITY; lo_is_finite = 0; }
      else lo_is_sing = 1;
      if (vlb > -INF20 && up_is_finite) ub += coef * vlb;
      else if (up_is_sing) { up_is_sing = 0; ub = INFINITY; up_is_finite = 0; }
      else up_is_sing = 1;
    }
  }
  *lo = lb; *up = ub;
}

void orc_lin_row_activity(const orc_lin_t *p, int32_t row, const double *lb, const double *ub,
                          double out[4])
{
  int32_t b = p->row_ptr[row], k = p->row_ptr[row + 1] - b;
  lf_bnds(k, p->col + b, p->val + b, lb, ub, &out[0], &out[1]);
  out[2] = -INFINITY; out[3] = INFINITY;   /* ref: LinearHandler.cpp:960 */
  lf_sing_bnds(k, p->col + b, p->val + b, lb, ub, &out[2], &out[3]);
}

/* --- in-place (Gauss-Seidel) state: the reference mutates Variable::lb_/ub_ and the
 *     per-row bFlag while it sweeps ------------------------------------------------ */
typedef struct {
  const orc_lin_t *p;
  double *lb, *ub;
  uint8_t *bflag;            /* ref: Constraint::bTemp_ */
  int32_t *csc_ptr, *csc_row;/* var -> rows, ref: Variable::cons_ */
  int64_t n_mods, nnz_updates;
  uint32_t nintmods;
} gs_t;

static void build_csc(const orc_lin_t *p, int32_t **ptr_out, int32_t **row_out)
{
  int32_t n = p->n, m = p->m;
  int32_t *ptr = (int32_t *)calloc((size_t)n + 2, sizeof(int32_t));
  int32_t nnz = p->row_ptr[m];
  int32_t *rows = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  for (int32_t t = 0; t < nnz; ++t) ptr[p->col[t] + 2]++;
  for (int32_t j = 0; j < n; ++j) ptr[j + 2] += ptr[j + 1];
  for (int32_t i = 0; i < m; ++i)
    for (int32_t t = p->row_ptr[i]; t < p->row_ptr[i + 1]; ++t) rows[ptr[p->col[t] + 1]++] = i;
  *ptr_out = ptr; *row_out = rows;
}

/* ref: LinearHandler.cpp:1229-1234 changeBFlag_ */
static void change_bflag(gs_t *s, int32_t j)
{
  for (int32_t t = s->csc_ptr[j]; t < s->csc_ptr[j + 1]; ++t) s->bflag[s->csc_row[t]] = 1;
}

/* ref: LinearHandler.cpp:1048-1136 updateLfBoundsFromLb_ (apply_to_prob == false) */
static void gs_from_lb(gs_t *s, int32_t k, const int32_t *col, const double *val, double rlb,
                       double uu, int is_sing, int *changed)
{
  for (int32_t t = 0; t < k; ++t) {
    int32_t j = col[t];
    double coef = val[t], vlb = s->lb[j], vub = s->ub[j];
    if (coef > E_TOL && (!is_sing || vub >= INF20)) {
      if (vub >= INF20) vub = 0.;
      double nlb = (rlb - uu) / coef + vub;
      if (nlb > vlb + E_TOL) {
        if (nlb > s->ub[j] - E_TOL) nlb = s->ub[j];
        change_bflag(s, j);
        s->lb[j] = nlb; s->n_mods++;
        if (is_int_type(s->p->var_type[j])) s->nintmods++;
        *changed = 1;
      }
    } else if (coef < -E_TOL && (!is_sing || vlb <= -INF20)) {
      if (vlb <= -INF20) vlb = 0.;
      double nub = (rlb - uu) / coef + vlb;
      if (nub < vub - E_TOL) {
        if (nub < s->lb[j] + E_TOL) nub = s->lb[j];
        change_bflag(s, j);
        s->ub[j] = nub; s->n_mods++;
        if (is_int_type(s->p->var_type[j])) s->nintmods++;
        *changed = 1;
      }
    }
  }
}

/* ref: LinearHandler.cpp:1139-1226 updateLfBoundsFromUb_ (apply_to_prob == false) */
static void gs_from_ub(gs_t *s, int32_t k, const int32_t *col, const double *val, double rub,
                       double ll, int is_sing, int *changed)
{
  for (int32_t t = 0; t < k; ++t) {
    int32_t j = col[t];
    double coef = val[t], vlb = s->lb[j], vub = s->ub[j];
    if (coef > E_TOL && (!is_sing || vlb <= -INF20)) {
      if (vlb <= -INF20) vlb = 0.;
      double nub = (rub - ll) / coef + vlb;
      if (nub < vub - E_TOL) {
        if (nub < s->lb[j] + E_TOL) nub = s->lb[j];
        change_bflag(s, j);
        s->ub[j] = nub; s->n_mods++;
        if (is_int_type(s->p->var_type[j])) s->nintmods++;
        *changed = 1;
      }
    } else if (coef < -E_TOL && (!is_sing || vub >= INF20)) {
      if (vub >= INF20) vub = 0.;
      double nlb = (rub - ll) / coef + vub;
      if (nlb > vlb + E_TOL) {
        if (nlb > s->ub[j] - E_TOL) nlb = s->ub[j];
        change_bflag(s, j);
        s->lb[j] = nlb; s->n_mods++;
        if (is_int_type(s->p->var_type[j])) s->nintmods++;
        *changed = 1;
      }
    }
  }
}

/* ref: LinearHandler.cpp:952-1045 linBndTighten_ with apply_to_prob == false.
 * returns 1 if the row is activity-infeasible. */
static int gs_row(gs_t *s, int32_t k, const int32_t *col, const double *val, double lb,
                  double ub, int *changed)
{
  double ll, uu, sing_ll = -INFINITY, sing_uu = INFINITY;
  *changed = 0;
  lf_bnds(k, col, val, s->lb, s->ub, &ll, &uu);
  if (ll < -INF20 || uu > INF20) lf_sing_bnds(k, col, val, s->lb, s->ub, &sing_ll, &sing_uu);
  s->nnz_updates += k;
  if (ll > ub + E_TOL) return 1;                 /* :994 */
  if (uu < lb - E_TOL) return 1;                 /* :1005 */
  if (lb > -INF20) {                             /* :1017-1025 */
    if (uu < INF20) gs_from_lb(s, k, col, val, lb, uu, 0, changed);
    else if (sing_uu < INF20) gs_from_lb(s, k, col, val, lb, sing_uu, 1, changed);
  }
  if (*changed) {                                /* :1027-1032 */
    lf_bnds(k, col, val, s->lb, s->ub, &ll, &uu);
    if (ll < -INF20 || uu > INF20) lf_sing_bnds(k, col, val, s->lb, s->ub, &sing_ll, &sing_uu);
  }
  if (ub < INF20) {                              /* :1035-1043 */
    if (ll > -INF20) gs_from_ub(s, k, col, val, ub, ll, 0, changed);
    else if (sing_ll > -INF20) gs_from_ub(s, k, col, val, ub, sing_ll, 1, changed);
  }
  return 0;
}

/* ref: LinearHandler.cpp:493-541 varBndsFromCons_ with apply_to_prob == false: every
 * flagged, non-deleted linear row once, in index order; stops at the first
 * activity-infeasible row (:517-519) */
static int gs_rows_sweep(gs_t *s, int *changed)
{
  const orc_lin_t *p = s->p;
  for (int32_t i = 0; i < p->m; ++i) {
    if (!s->bflag[i] || (p->row_active && !p->row_active[i])) continue;
    s->bflag[i] = 0;
    int32_t b = p->row_ptr[i], k = p->row_ptr[i + 1] - b;
    int t_changed = 0;
    if (gs_row(s, k, p->col + b, p->val + b, p->row_lb[i], p->row_ub[i], &t_changed)) return 1;
    if (t_changed) *changed = 1;
  }
  return 0;
}

/* ref: LinearHandler.cpp:544-597 varBndsFromObj_ (cut-off row, looped to its own fixpoint) */
static int gs_cutoff(gs_t *s, int *changed)
{
  const orc_lin_t *p = s->p;
  if (p->cut_k <= 0) return 0;
  int t_changed = 1;
  const uint32_t nintmods_keep = s->nintmods;  /* varBndsFromObj_ counts into a local it never reads (:554) */
  int status = 0;
  while (t_changed) {
    double ll, uu, sing_ll = INFINITY, sing_uu = INFINITY;
    t_changed = 0;
    lf_bnds(p->cut_k, p->cut_col, p->cut_val, s->lb, s->ub, &ll, &uu);
    if (ll < -INF20 || uu > INF20)
      lf_sing_bnds(p->cut_k, p->cut_col, p->cut_val, s->lb, s->ub, &sing_ll, &sing_uu);
    s->nnz_updates += p->cut_k;
    if (ll > p->cut_rhs + E_TOL) { status = 1; break; }
    if (ll > -INF20) gs_from_ub(s, p->cut_k, p->cut_col, p->cut_val, p->cut_rhs, ll, 0, &t_changed);
    else if (sing_ll > -INF20)
      gs_from_ub(s, p->cut_k, p->cut_col, p->cut_val, p->cut_rhs, sing_ll, 1, &t_changed);
    if (t_changed) *changed = 1;
  }
  s->nintmods = nintmods_keep;
  return status;
}

/* ref: LinearHandler.cpp:415-490 tightenInts_ with apply_to_prob == false */
static void gs_tighten_ints(gs_t *s, int *changed)
{
  for (int32_t j = 0; j < s->p->n; ++j) {
    if (!is_int_type(s->p->var_type[j])) continue;
    double lb = s->lb[j], ub = s->ub[j];
    if (lb > -INF20 && fabs(lb - floor(lb + 0.5)) > INT_TOL) {
      change_bflag(s, j); s->lb[j] = ceil(lb); s->n_mods++; *changed = 1;
    }
    if (ub < INF20 && fabs(ub - floor(ub + 0.5)) > INT_TOL) {
      s->ub[j] = floor(ub); change_bflag(s, j); s->n_mods++; *changed = 1;
    }
  }
}

/* ref: LinearHandler.cpp:328-359 checkBounds_ */
static int check_bounds(const orc_lin_t *p, const double *lb, const double *ub)
{
  for (int32_t j = 0; j < p->n; ++j) if (lb[j] > ub[j] + E_TOL) return 1;
  for (int32_t i = 0; i < p->m; ++i) {
    if (p->row_active && !p->row_active[i]) continue;  /* deleted rows are gone from cons_ */
    if (p->row_lb[i] > p->row_ub[i] + E_TOL) return 1;
  }
  return 0;
}

static void gs_init(gs_t *s, const orc_lin_t *p, double *lb, double *ub)
{
  memset(s, 0, sizeof(*s));
  s->p = p; s->lb = lb; s->ub = ub;
  s->bflag = (uint8_t *)malloc((size_t)(p->m > 0 ? p->m : 1));
  memset(s->bflag, 1, (size_t)p->m);     /* ref: LinearHandler.cpp:1618-1622 */
  build_csc(p, &s->csc_ptr, &s->csc_row);
}

static void gs_free(gs_t *s) { free(s->bflag); free(s->csc_ptr); free(s->csc_row); }

/* ref: LinearHandler.cpp:1605-1653 simplePresolve.  The return value of the row sweep
 * is dropped (:1631); only checkBounds_ sets the verdict. */
void orc_lin_simple_presolve(const orc_lin_t *p, double *lb, double *ub, orc_result_t *res)
{
  gs_t s; gs_init(&s, p, lb, ub);
  int changed = 1, infeasible = 0;
  uint32_t iters = 1; const uint32_t max_iters = 10, min_iters = 2;
  int32_t rounds = 0;
  while (changed && iters <= max_iters && (iters <= min_iters || s.nintmods > 0) && !infeasible) {
    s.nintmods = 0; changed = 0; ++iters; ++rounds;
    (void)gs_rows_sweep(&s, &changed);
    (void)gs_cutoff(&s, &changed);           /* status ignored too, :1636-1640 */
    gs_tighten_ints(&s, &changed);
    infeasible = check_bounds(p, lb, ub);
  }
  res->verdict = infeasible ? ORC_INFEASIBLE : ORC_OK;
  res->rounds = rounds; res->nnz_updates = s.nnz_updates; res->n_mods = s.n_mods;
  gs_free(&s);
}

/* SURVEY.md section 8c parity driver: same sweeps, status honoured, to a fixpoint */
void orc_lin_fixpoint_inplace(const orc_lin_t *p, double *lb, double *ub, orc_result_t *res)
{
  gs_t s; gs_init(&s, p, lb, ub);
  int changed = 1, infeasible = 0;
  int32_t rounds = 0;
  while (changed && !infeasible) {
    changed = 0; ++rounds;
    if (gs_rows_sweep(&s, &changed)) { infeasible = 1; break; }
    if (gs_cutoff(&s, &changed)) { infeasible = 1; break; }
    gs_tighten_ints(&s, &changed);
    infeasible = check_bounds(p, lb, ub);
  }
  res->verdict = infeasible ? ORC_INFEASIBLE : ORC_OK;
  res->rounds = rounds; res->nnz_updates = s.nnz_updates; res->n_mods = s.n_mods;
  gs_free(&s);
}

/* --- Jacobi round: the rule the CUDA kernels implement ----- */

typedef struct { const double *L, *U; double *NL, *NU; } jac_t;

static void jac_from_lb(jac_t *s, int32_t k, const int32_t *col, const double *val, double rlb,
                        double act, int sing)
{
  for (int32_t t = 0; t < k; ++t) {
    int32_t j = col[t];
    double a = val[t], vl = s->L[j], vu = s->U[j];
    if (a > E_TOL && (!sing || vu >= INF20)) {
      if (vu >= INF20) vu = 0.;
      double c = (rlb - act) / a + vu;
      if (c > vl + E_TOL) { if (c > s->U[j] - E_TOL) c = s->U[j]; if (c > s->NL[j]) s->NL[j] = c; }
    } else if (a < -E_TOL && (!sing || vl <= -INF20)) {
      if (vl <= -INF20) vl = 0.;
      double c = (rlb - act) / a + vl;
      if (c < vu - E_TOL) { if (c < s->L[j] + E_TOL) c = s->L[j]; if (c < s->NU[j]) s->NU[j] = c; }
    }
  }
}

static void jac_from_ub(jac_t *s, int32_t k, const int32_t *col, const double *val, double rub,
                        double act, int sing)
{
  for (int32_t t = 0; t < k; ++t) {
    int32_t j = col[t];
    double a = val[t], vl = s->L[j], vu = s->U[j];
    if (a > E_TOL && (!sing || vl <= -INF20)) {
      if (vl <= -INF20) vl = 0.;
      double c = (rub - act) / a + vl;
      if (c < vu - E_TOL) { if (c < s->L[j] + E_TOL) c = s->L[j]; if (c < s->NU[j]) s->NU[j] = c; }
    } else if (a < -E_TOL && (!sing || vu >= INF20)) {
      if (vu >= INF20) vu = 0.;
      double c = (rub - act) / a + vu;
      if (c > vl + E_TOL) { if (c > s->U[j] - E_TOL) c = s->U[j]; if (c > s->NL[j]) s->NL[j] = c; }
    }
  }
}

static int jac_row(jac_t *s, int32_t k, const int32_t *col, const double *val, double rl, double ru)
{
  double ll, uu, sing_ll = -INFINITY, sing_uu = INFINITY;
  lf_bnds(k, col, val, s->L, s->U, &ll, &uu);
  if (ll < -INF20 || uu > INF20) lf_sing_bnds(k, col, val, s->L, s->U, &sing_ll, &sing_uu);
  if (ll > ru + E_TOL || uu < rl - E_TOL) return 1;
  if (rl > -INF20) {
    if (uu < INF20) jac_from_lb(s, k, col, val, rl, uu, 0);
    else if (sing_uu < INF20) jac_from_lb(s, k, col, val, rl, sing_uu, 1);
  }
  if (ru < INF20) {
    if (ll > -INF20) jac_from_ub(s, k, col, val, ru, ll, 0);
    else if (sing_ll > -INF20) jac_from_ub(s, k, col, val, ru, sing_ll, 1);
  }
  return 0;
}

void orc_lin_fixpoint_jacobi(const orc_lin_t *p, double *lb, double *ub, int32_t max_rounds,
                             orc_result_t *res)
{
  int32_t n = p->n, m = p->m;
  double *NL = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *NU = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  uint8_t *flag = (uint8_t *)malloc((size_t)(m > 0 ? m : 1));
  int32_t *csc_ptr, *csc_row;
  build_csc(p, &csc_ptr, &csc_row);
  memset(flag, 1, (size_t)m);
  jac_t s = { lb, ub, NL, NU };
  int changed = 1, infeasible = 0;
  int32_t rounds = 0; int64_t nnz_updates = 0, n_mods = 0;
  while (changed && !infeasible && (max_rounds <= 0 || rounds < max_rounds)) {
    changed = 0; ++rounds;
    memcpy(NL, lb, sizeof(double) * (size_t)n);
    memcpy(NU, ub, sizeof(double) * (size_t)n);
    for (int32_t i = 0; i < m && !infeasible; ++i) {
      if (!flag[i] || (p->row_active && !p->row_active[i])) continue;
      flag[i] = 0;
      int32_t b = p->row_ptr[i], k = p->row_ptr[i + 1] - b;
      nnz_updates += k;
      if (jac_row(&s, k, p->col + b, p->val + b, p->row_lb[i], p->row_ub[i])) infeasible = 1;
    }
    if (!infeasible && p->cut_k > 0) {
      nnz_updates += p->cut_k;
      if (jac_row(&s, p->cut_k, p->cut_col, p->cut_val, -INFINITY, p->cut_rhs)) infeasible = 1;
    }
    if (infeasible) break;
    for (int32_t j = 0; j < n; ++j) {             /* L7 */
      if (!is_int_type(p->var_type[j])) continue;
      if (NL[j] > -INF20 && fabs(NL[j] - floor(NL[j] + 0.5)) > INT_TOL) NL[j] = ceil(NL[j]);
      if (NU[j] < INF20 && fabs(NU[j] - floor(NU[j] + 0.5)) > INT_TOL) NU[j] = floor(NU[j]);
    }
    for (int32_t j = 0; j < n; ++j) {
      int ch = 0;
      if (NL[j] != lb[j]) { ch = 1; n_mods++; }
      if (NU[j] != ub[j]) { ch = 1; n_mods++; }
      if (ch) {
        changed = 1;
        for (int32_t t = csc_ptr[j]; t < csc_ptr[j + 1]; ++t) flag[csc_row[t]] = 1;
      }
      lb[j] = NL[j]; ub[j] = NU[j];
    }
    infeasible = check_bounds(p, lb, ub);        /* L8 */
  }
  res->verdict = infeasible ? ORC_INFEASIBLE : ORC_OK;
  res->rounds = rounds; res->nnz_updates = nnz_updates; res->n_mods = n_mods;
  free(NL); free(NU); free(flag); free(csc_ptr); free(csc_row);
}

/* One Jacobi round split at the point where the row-partitioned multi-GPU mode merges candidate
 * bounds across ranks: (1) rows -> candidates NL/NU (combined with max/min, so per-block results can be
 * merged with an element-wise MAX / MIN all-reduce), (2) integer rounding + bound check + commit.
 * Used by the world_size-2 gloo test of the merge protocol. */
int32_t orc_lin_jacobi_round_rows(const orc_lin_t *p, const double *lb, const double *ub, double *NL, double *NU)
{
  jac_t s = { lb, ub, NL, NU };
  memcpy(NL, lb, sizeof(double) * (size_t)p->n);
  memcpy(NU, ub, sizeof(double) * (size_t)p->n);
  for (int32_t i = 0; i < p->m; ++i) {
    if (p->row_active && !p->row_active[i]) continue;
    int32_t b = p->row_ptr[i], k = p->row_ptr[i + 1] - b;
    if (jac_row(&s, k, p->col + b, p->val + b, p->row_lb[i], p->row_ub[i])) return 1;
  }
  return 0;
}

int32_t orc_lin_jacobi_round_vars(const orc_lin_t *p, double *lb, double *ub, double *NL, double *NU,
                                  int32_t *changed)
{
  *changed = 0;
  for (int32_t j = 0; j < p->n; ++j) {
    if (is_int_type(p->var_type[j])) {
      if (NL[j] > -INF20 && fabs(NL[j] - floor(NL[j] + 0.5)) > INT_TOL) NL[j] = ceil(NL[j]);
      if (NU[j] < INF20 && fabs(NU[j] - floor(NU[j] + 0.5)) > INT_TOL) NU[j] = floor(NU[j]);
    }
    if (NL[j] != lb[j] || NU[j] != ub[j]) *changed = 1;
    lb[j] = NL[j]; ub[j] = NU[j];
  }
  for (int32_t j = 0; j < p->n; ++j) if (lb[j] > ub[j] + E_TOL) return 1;
  return 0;
}

/* ===========================================================================
 *                         interval helpers (Operations.cpp)
 * ========================================================================= */

/* std::min / std::max exactly as the reference uses them (first argument wins ties, so the sign of
 * a zero is preserved the way libstdc++ does) */
#define STD_MIN(a, b) (((b) < (a)) ? (b) : (a))
#define STD_MAX(a, b) (((a) < (b)) ? (b) : (a))

/* ref: Operations.cpp:117-179 */
void orc_bounds_on_product(int zero_x_inf_zero, double l0, double u0, double l1, double u1,
                           double *lb, double *ub)
{
  double prod;
  if (fabs(l1) <= 1e-10 && fabs(u1) <= 1e-10) {
    prod = l1; l1 = l0; l0 = prod;
    prod = u1; u1 = u0; u0 = prod;
  }
  if (fabs(l0) <= 1e-10 && fabs(u0) <= 1e-10) {
    if (zero_x_inf_zero) { *lb = 0.0; *ub = 0.0; }
    else {
      *lb = (l1 == -INFINITY) ? -INFINITY : 0.0;
      *ub = (u1 == INFINITY) ? INFINITY : 0.0;
    }
  } else if ((l1 == -INFINITY && u1 == INFINITY) || (l0 == -INFINITY && u0 == INFINITY)) {
    *lb = -INFINITY; *ub = INFINITY;
  } else {
    double l, u;
    prod = l0 * l1; if (isnan(prod)) prod = -INFINITY;
    l = prod; u = prod;
    prod = u0 * l1; if (isnan(prod)) prod = INFINITY;
    l = STD_MIN(l, prod); u = STD_MAX(u, prod);
    prod = u0 * u1; if (isnan(prod)) prod = -INFINITY;
    l = STD_MIN(l, prod); u = STD_MAX(u, prod);
    prod = l0 * u1; if (isnan(prod)) prod = INFINITY;
    l = STD_MIN(l, prod); u = STD_MAX(u, prod);
    *lb = l; *ub = u;
  }
}

/* ref: Operations.cpp:182-212 */
void orc_bounds_on_recip(double l0, double u0, double *lb, double *ub)
{
  if ((fabs(u0) < 1e-10) && (fabs(l0) < 1e-10)) { *lb = -INFINITY; *ub = INFINITY; }
  else if (l0 < -1e-10 && u0 > 1e-10) { *lb = -INFINITY; *ub = INFINITY; }
  else if ((fabs(u0) < 1e-10) && l0 < 0) { *lb = -INFINITY; *ub = 1.0 / l0; }
  else if ((fabs(l0) < 1e-10) && u0 < 0) { *lb = 1.0 / u0; *ub = INFINITY; }
  else { *lb = 1.0 / u0; *ub = 1.0 / l0; }
}

/* ref: Operations.cpp:100-106 */
void orc_bounds_on_div(double l0, double u0, double l1, double u1, double *lb, double *ub)
{
  double tl, tu;
  orc_bounds_on_recip(l1, u1, &tl, &tu);
  orc_bounds_on_product(0, l0, u0, tl, tu, lb, ub);
}

/* ref: Operations.cpp:233-246 */
void orc_bounds_on_square(double l1, double u1, double *lb, double *ub)
{
  if (u1 < 0.) { *lb = u1 * u1; *ub = l1 * l1; }
  else if (l1 > 0.) { *lb = l1 * l1; *ub = u1 * u1; }
  else { double p = l1 * l1, q = u1 * u1; *lb = 0.; *ub = STD_MAX(p, q); }
}

/* ref: Operations.cpp:80-83 */
static int orc_is_int(double v) { return fabs(floor(v + 0.5) - v) < 1e-12; }

/* ===========================================================================
 *                               CGraph tapes
 * ========================================================================= */

#define ORC_PI 3.141592653589793      /* ref: CNode.cpp:25 */
#define ORC_MINFTY 1e25               /* ref: CNode.cpp:26 */
#define ORC_MAX_NODES 4096

typedef struct {
  const orc_nl_t *g; int32_t base, nn;
  double *nlb, *nub;       /* per local node */
} tape_t;

/* ref: CNode.cpp:1701-1904 updateBnd, one node.  Constants are not part of vq_/dq_ in
 * the reference: OpNum keeps [d,d] from CNode::setVal (:1693-1699) and OpInt keeps the
 * constructor's (-inf,inf) (:57-79, CGraph.cpp:1238-1245); neither gets the 1e25 clamp.
 * NOTE: the reference never resets constant-node bounds between calls, so reverse
 * propagation can leave a drift of <=1e-7 on an OpNum node for later calls; this
 * restatement evaluates every call from a fresh tape (documented deviation). */
static void node_forward(tape_t *t, int32_t i, const double *vlb, const double *vub, int *error)
{
  const orc_nl_t *g = t->g; int32_t q = t->base + i;
  int op = g->op[q];
  double *lb_ = &t->nlb[i], *ub_ = &t->nub[i];
  double llb = 0, lub = 0, rlb = 0, rub = 0;
  if (op == ORC_OpNum) { *lb_ = *ub_ = g->cnst[q]; return; }
  if (op == ORC_OpInt) { *lb_ = -INFINITY; *ub_ = INFINITY; return; }
  if (op != ORC_OpVar && op != ORC_OpSumList) {
    llb = t->nlb[g->arg0[q]]; lub = t->nub[g->arg0[q]];
    if (g->arg1[q] >= 0) { rlb = t->nlb[g->arg1[q]]; rub = t->nub[g->arg1[q]]; }
  }
  errno = 0;
  switch (op) {
  case ORC_OpAbs:
    if (lub < 0) { *lb_ = -lub; *ub_ = -llb; }
    else if (llb < 0) { if (-llb > lub) { *lb_ = 0.0; *ub_ = -llb; } else { *lb_ = 0.0; *ub_ = lub; } }
    else { *lb_ = llb; *ub_ = lub; }
    break;
  case ORC_OpAcos: *lb_ = 0.0; *ub_ = ORC_PI; break;
  case ORC_OpAsin: case ORC_OpAtan: *lb_ = -ORC_PI / 2; *ub_ = ORC_PI / 2; break;
  case ORC_OpCeil: *lb_ = ceil(llb); *ub_ = ceil(lub); break;
  case ORC_OpCos: case ORC_OpSin: *lb_ = -1.0; *ub_ = 1.0; break;
  case ORC_OpDiv: orc_bounds_on_div(llb, lub, rlb, rub, lb_, ub_); break;
  case ORC_OpExp:
    *lb_ = (llb == -INFINITY) ? 0.0 : exp(llb);
    *ub_ = (lub == INFINITY) ? INFINITY : exp(lub);
    break;
  case ORC_OpFloor: *lb_ = floor(llb); *ub_ = floor(lub); break;
  case ORC_OpLog: *lb_ = (llb <= 0.0) ? -INFINITY : log(llb); *ub_ = log(lub); break;
  case ORC_OpLog10: *lb_ = (llb <= 0.0) ? -INFINITY : log10(llb); *ub_ = log10(lub); break;
  case ORC_OpMinus: *lb_ = llb - rub; *ub_ = lub - rlb; break;
  case ORC_OpMult: orc_bounds_on_product(1, llb, lub, rlb, rub, lb_, ub_); break;
  case ORC_OpNone: break;
  case ORC_OpPlus: *lb_ = llb + rlb; *ub_ = lub + rub; break;
  case ORC_OpSqr: orc_bounds_on_square(llb, lub, lb_, ub_); break;
  case ORC_OpSqrt: *lb_ = (llb < 1e-12) ? 0.0 : sqrt(llb); *ub_ = sqrt(lub); break;
  case ORC_OpSumList: {
    double l = 0.0, u = 0.0;
    for (int32_t c = g->arg0[q]; c < g->arg1[q]; ++c) { l += t->nlb[g->child[c]]; u += t->nub[g->child[c]]; }
    *lb_ = l; *ub_ = u;
  } break;
  case ORC_OpUMinus: *lb_ = -lub; *ub_ = -llb; break;
  case ORC_OpVar: *lb_ = vlb[g->arg0[q]]; *ub_ = vub[g->arg0[q]]; break;
  default:  /* Acosh Asinh Atanh Cosh CPow IntDiv Pow PowK Round Sinh Tan Tanh: "TODO" in the ref */
    *lb_ = -INFINITY; *ub_ = INFINITY; break;
  }
  if (errno != 0) *error = errno;
  if (*lb_ < -ORC_MINFTY) *lb_ = -INFINITY;
  if (*ub_ > ORC_MINFTY) *ub_ = INFINITY;
}

/* ref: CNode.cpp:1504-1526 propBounds_ ; NaN trips an assert in the reference -> error */
static void prop_child(tape_t *t, int32_t c, double lb, double ub, int *is_inf, int *error)
{
  const double etol = 1e-7;
  if (isnan(lb) || isnan(ub)) { *error = 9999; return; }
  if (lb < -ORC_MINFTY) lb = -INFINITY;
  if (ub > ORC_MINFTY) ub = INFINITY;
  if (lb > ub + etol || ub < t->nlb[c] - etol || lb > t->nub[c] + etol) *is_inf = 1;
  else { if (lb > t->nlb[c]) t->nlb[c] = lb; if (ub < t->nub[c]) t->nub[c] = ub; }
}

/* ref: CNode.cpp:1259-1501 propBounds, one node, reference quirks kept */
static void node_reverse(tape_t *t, int32_t i, int *is_inf, int *error)
{
  const orc_nl_t *g = t->g; int32_t q = t->base + i;
  int op = g->op[q];
  int32_t l = g->arg0[q], r = g->arg1[q];
  double lb_ = t->nlb[i], ub_ = t->nub[i];
  double lb = -INFINITY, ub = INFINITY;
  errno = 0;
  switch (op) {
  case ORC_OpAbs: lb = -ub_; ub = ub_; prop_child(t, l, lb, ub, is_inf, error); break;
  case ORC_OpAcos: case ORC_OpAsin: prop_child(t, l, -1.0, 1.0, is_inf, error); break;
  case ORC_OpCeil: prop_child(t, l, floor(lb_), floor(ub_), is_inf, error); break;
  case ORC_OpDiv:
    orc_bounds_on_product(0, t->nlb[r], t->nub[r], lb_, ub_, &lb, &ub);
    prop_child(t, l, lb, ub, is_inf, error);
    orc_bounds_on_div(t->nlb[l], t->nub[l], lb_, ub_, &lb, &ub);
    prop_child(t, r, lb, ub, is_inf, error);
    break;
  case ORC_OpExp: lb = log(lb_); ub = log(ub_); prop_child(t, l, lb, ub, is_inf, error); break;
  case ORC_OpFloor: prop_child(t, l, ceil(lb_), ceil(ub_), is_inf, error); break;
  case ORC_OpLog: lb = exp(lb_); ub = exp(ub_); prop_child(t, l, lb, ub, is_inf, error); break;
  case ORC_OpLog10: lb = pow(10.0, lb_); ub = pow(10.0, ub_); prop_child(t, l, lb, ub, is_inf, error); break;
  case ORC_OpMinus:
    lb = lb_ + t->nlb[r]; ub = ub_ + t->nub[r];
    prop_child(t, l, lb, ub, is_inf, error);
    lb = t->nlb[l] - ub_; ub = t->nub[l] - lb_;
    prop_child(t, r, lb, ub, is_inf, error);
    break;
  case ORC_OpMult:
    orc_bounds_on_div(lb_, ub_, t->nlb[r], t->nub[r], &lb, &ub);
    prop_child(t, l, lb, ub, is_inf, error);
    orc_bounds_on_div(lb_, ub_, t->nlb[l], t->nub[l], &lb, &ub);
    prop_child(t, r, lb, ub, is_inf, error);
    break;
  case ORC_OpPlus:
    lb = lb_ - t->nub[r]; ub = ub_ - t->nlb[r];
    prop_child(t, l, lb, ub, is_inf, error);
    lb = lb_ - t->nub[l]; ub = ub_ - t->nlb[l];
    prop_child(t, r, lb, ub, is_inf, error);
    break;
  case ORC_OpPowK: {
    double k = g->cnst[t->base + r];           /* r_->val_ */
    if (k > 0) {
      if (orc_is_int(k / 2.0)) {
        if (ub_ < -1e-12) *error = 3141;
        else { ub = pow(ub_, 1.0 / k); lb = -ub; prop_child(t, l, lb, ub, is_inf, error); }
      } else if (orc_is_int((k + 1) / 2.0)) {
        /* the reference tests the LOCALS (lb=-inf, ub=+inf), :1377-1386 */
        if (lb < 0) lb = -pow(-lb_, 1.0 / k); else lb = pow(lb_, 1.0 / k);
        if (ub < 0) ub = -pow(-ub_, 1.0 / k); else ub = pow(ub_, 1.0 / k);
        prop_child(t, l, lb, ub, is_inf, error);
      }
    }
  } break;
  case ORC_OpSqr:   /* local ub = +inf: a no-op in the reference, :1399-1403 */
    ub = sqrt(ub); lb = -ub; prop_child(t, l, lb, ub, is_inf, error); break;
  case ORC_OpSqrt:  /* local lb = -inf: only child >= 0 is ever derived, :1404-1412 */
    if (ub_ < 0.0) *is_inf = 1;
    else if (lb >= 0.0) prop_child(t, l, lb * lb, ub * ub, is_inf, error);
    else prop_child(t, l, 0.0, ub * ub, is_inf, error);
    break;
  case ORC_OpSumList: {   /* :1413-1481, including the tub = -inf defect at :1471-1473 */
    int inf_lb = 0, inf_ub = 0; int32_t c0 = g->arg0[q], c1 = g->arg1[q];
    lb = 0.0;
    for (int32_t c = c0; c < c1; ++c) {
      double cl = t->nlb[g->child[c]];
      if (cl > -INFINITY) lb += cl; else if (inf_lb) { lb = -INFINITY; break; } else inf_lb = 1;
    }
    ub = 0.0;
    for (int32_t c = c0; c < c1; ++c) {
      double cu = t->nub[g->child[c]];
      if (cu < INFINITY) ub += cu; else if (inf_ub) { ub = INFINITY; break; } else inf_ub = 1;
    }
    if (lb > -INFINITY || ub < INFINITY) {
      for (int32_t c = c0; c < c1; ++c) {
        int32_t ch = g->child[c]; double tlb, tub;
        if (ub < INFINITY) {
          if (!inf_ub) tlb = lb_ - (ub - t->nub[ch]);
          else if (t->nub[ch] < INFINITY) tlb = -INFINITY;
          else tlb = lb_ - ub;
        } else tlb = -INFINITY;
        if (lb > -INFINITY) {
          if (!inf_lb) tub = ub_ - (lb - t->nlb[ch]);
          else if (t->nlb[ch] > -INFINITY) tub = INFINITY;
          else tub = ub_ - lb;
        } else tub = -INFINITY;
        prop_child(t, ch, tlb, tub, is_inf, error);
        if (*is_inf) break;
      }
    }
  } break;
  case ORC_OpUMinus: prop_child(t, l, -ub_, -lb_, is_inf, error); break;
  default: break;      /* unimplemented in the reference: no-op */
  }
  if (errno != 0) *error = errno;
}

static int is_leaf(int op) { return op == ORC_OpVar || op == ORC_OpNum || op == ORC_OpInt; }

/* ref: CGraph.cpp:172-183 computeBounds */
static int tape_forward(tape_t *t, const double *lb, const double *ub)
{
  int error = 0;
  for (int32_t i = 0; i < t->nn; ++i) node_forward(t, i, lb, ub, &error);
  return error;
}

static int tape_open(tape_t *t, const orc_nl_t *g, int32_t c, double *bufl, double *bufu)
{
  t->g = g; t->base = g->tape_ptr[c]; t->nn = g->tape_ptr[c + 1] - t->base;
  t->nlb = bufl; t->nub = bufu;
  return t->nn > 0 && t->nn <= ORC_MAX_NODES;
}

int32_t orc_nl_compute_bounds(const orc_nl_t *g, int32_t c, const double *lb, const double *ub,
                              double *out_lb, double *out_ub)
{
  double bl[ORC_MAX_NODES], bu[ORC_MAX_NODES]; tape_t t;
  if (!tape_open(&t, g, c, bl, bu)) return -1;
  int err = tape_forward(&t, lb, ub);
  *out_lb = t.nlb[t.nn - 1]; *out_ub = t.nub[t.nn - 1];
  return err;
}

/* ref: CGraph.cpp:1605-1644 varBoundMods; mods applied as NlPresHandler.cpp:1787-1803 does */
int32_t orc_nl_var_bound_mods(const orc_nl_t *g, int32_t c, double lb_in, double ub_in,
                              double *lb, double *ub, int32_t *n_mods)
{
  const double bslack = 1e-5, bslack10 = 1e-4;
  double bl[ORC_MAX_NODES], bu[ORC_MAX_NODES]; tape_t t;
  *n_mods = 0;
  if (!tape_open(&t, g, c, bl, bu)) return ORC_ERROR;
  int error = tape_forward(&t, lb, ub);
  if (error > 0) return ORC_ERROR;
  int32_t o = t.nn - 1; int is_inf = 0;
  t.nlb[o] = fmax(lb_in, t.nlb[o]); t.nub[o] = fmin(ub_in, t.nub[o]);
  for (int32_t i = t.nn - 1; i >= 0; --i) {
    if (is_leaf(g->op[t.base + i])) continue;        /* leaves are not in dq_ */
    node_reverse(&t, i, &is_inf, &error);
    if (is_inf) return ORC_INFEASIBLE;
    if (error > 0) return ORC_ERROR;
  }
  for (int32_t i = 0; i < t.nn; ++i) {
    if (g->op[t.base + i] != ORC_OpVar) continue;
    int32_t j = g->arg0[t.base + i];
    /* both tests read the variable's bounds BEFORE any mod of this call is applied */
    double ol = lb[j], ou = ub[j];
    if (t.nlb[i] > ol + bslack10) { lb[j] = t.nlb[i] - bslack; ++*n_mods; }
    if (t.nub[i] < ou - bslack10) { ub[j] = t.nub[i] + bslack; ++*n_mods; }
  }
  return ORC_OK;
}

/* ref: QuadraticFunction.cpp:156-180 computeBounds: per term the four corner products (coef*x1)*x2 -- evaluated left to
 * right -- their min added to the lower, their max to the upper bound; std::min / std::max as called there */
void orc_quad_compute_bounds(const orc_nl_t *g, int32_t q, const double *lbv, const double *ubv, double *out_lb, double *out_ub)
{
  double lb = 0, ub = 0;
  for (int32_t t = g->q_ptr[q]; t < g->q_ptr[q + 1]; ++t) {
    const double w = g->q_coef[t];
    const double l1 = lbv[g->q_v1[t]], u1 = ubv[g->q_v1[t]], l2 = lbv[g->q_v2[t]], u2 = ubv[g->q_v2[t]];
    const double a = w * l1 * l2, b = w * l1 * u2, c = w * u1 * l2, d = w * u1 * u2;
    double m = (b < a) ? b : a; m = (c < m) ? c : m; m = (d < m) ? d : m;
    lb += m;
    m = (a < b) ? b : a; m = (m < c) ? c : m; m = (m < d) ? d : m;
    ub += m;
  }
  *out_lb = lb; *out_ub = ub;
}

/* ref: NlPresHandler.cpp:101-208 chkRed_ (nlf branch, and the qf branch for QuadraticFunction constraints),
 * tolerance eTol_ = 1e-6 (:72).  The reference walks the constraints in index order and stops at the first
 * infeasible one: only the verdict is observable, so the two families are checked one after the other. */
int32_t orc_nl_chk_red(const orc_nl_t *g, const double *lb, const double *ub)
{
  for (int32_t q = 0; q < g->n_quad; ++q) {
    double lfl = 0, lfu = 0, ql = 0, qu = 0;
    int32_t b = g->q_lin_ptr[q], k = g->q_lin_ptr[q + 1] - b;
    if (k > 0) lf_bnds(k, g->q_lin_col + b, g->q_lin_val + b, lb, ub, &lfl, &lfu);
    orc_quad_compute_bounds(g, q, lb, ub, &ql, &qu);
    double impl_lb = ql + lfl, impl_ub = qu + lfu;
    if (impl_ub + 1e-6 < g->q_lb[q] || impl_lb - 1e-6 > g->q_ub[q]) return ORC_INFEASIBLE;
  }
  for (int32_t c = 0; c < g->n_cons; ++c) {
    double lfl = 0, lfu = 0, nl = 0, nu = 0;
    int32_t b = g->lin_ptr[c], k = g->lin_ptr[c + 1] - b;
    if (k > 0) lf_bnds(k, g->lin_col + b, g->lin_val + b, lb, ub, &lfl, &lfu);
    int err = orc_nl_compute_bounds(g, c, lb, ub, &nl, &nu);
    if (err != 0) return ORC_ERROR;              /* assert(error==0) in the reference */
    double impl_lb = nl + lfl, impl_ub = nu + lfu;
    if (impl_ub + 1e-6 < g->c_lb[c] || impl_lb - 1e-6 > g->c_ub[c]) return ORC_INFEASIBLE;
  }
  return ORC_OK;
}

/* ref: NlPresHandler.cpp:1686-1806 varBndsFromCons_ (nlf branch), in place, index order */
int32_t orc_nl_sweep(const orc_nl_t *g, double *lb, double *ub, int64_t *n_mods)
{
  for (int32_t c = 0; c < g->n_cons; ++c) {
    double lfl = 0, lfu = 0;
    int32_t b = g->lin_ptr[c], k = g->lin_ptr[c + 1] - b;
    if (k > 0) lf_bnds(k, g->lin_col + b, g->lin_val + b, lb, ub, &lfl, &lfu);
    double cub = g->c_ub[c] - lfl, clb = g->c_lb[c] - lfu;     /* :1775-1776 */
    int32_t nm = 0;
    int32_t st = orc_nl_var_bound_mods(g, c, clb, cub, lb, ub, &nm);
    if (st != ORC_OK) return st;
    *n_mods += nm;
  }
  return ORC_OK;
}

/* ref: NlPresHandler.cpp:1062-1121 fixObjBins_ for a LINEAR objective (obj = the cut-off row's coefficients):
 * with olb the objective's lower bound over the box (LinearFunction::computeBounds, LinearFunction.cpp:178-195,
 * taken ONCE), a binary z with coefficient a0 is fixed to 0 if a0>0 && olb+a0>ub, to 1 if a0<0 && olb-a0>ub. */
int32_t orc_nl_fix_obj_bins(const orc_lin_t *obj, double *lb, double *ub, int64_t *n_mods)
{
  if (!obj || obj->cut_k <= 0 || !(obj->obj_ub < INFINITY)) return ORC_OK;
  const double best = obj->obj_ub;
  double olb = 0.0;
  for (int32_t t = 0; t < obj->cut_k; ++t) {
    double a = obj->cut_val[t];
    olb += (a > 0) ? a * lb[obj->cut_col[t]] : a * ub[obj->cut_col[t]];
  }
  if (olb <= -INFINITY) return ORC_OK;
  if (olb > best) return ORC_INFEASIBLE;
  for (int32_t t = 0; t < obj->cut_k; ++t) {
    int32_t j = obj->cut_col[t];
    uint8_t ty = obj->var_type[j];
    if ((ty == 0 /* Binary */ || ty == 2 /* ImplBin */) && (ub[j] - lb[j]) > 1e-6) {
      double a0 = obj->cut_val[t];
      if (a0 > 0 && olb + a0 > best) { ub[j] = 0.0; ++*n_mods; }
      else if (a0 < 0 && olb - a0 > best) { lb[j] = 1.0; ++*n_mods; }
    }
  }
  return ORC_OK;
}

/* ref: NlPresHandler.cpp:1022-1059 simplePresolve; obj (may be NULL) carries the objective and the incumbent */
void orc_nl_simple_presolve_obj(const orc_nl_t *g, const orc_lin_t *obj, double *lb, double *ub, orc_result_t *res)
{
  int changed = 1; uint32_t iters = 1; const uint32_t max_iters = 4, min_iters = 2;
  int32_t verdict = ORC_OK, rounds = 0; int64_t n_mods = 0;
  while (changed && iters <= max_iters && iters <= min_iters && verdict != ORC_INFEASIBLE) {
    changed = 0; ++iters; ++rounds;
    int32_t st = orc_nl_chk_red(g, lb, ub);
    if (st == ORC_INFEASIBLE) { verdict = ORC_INFEASIBLE; break; }
    int64_t nm = 0;
    st = orc_nl_sweep(g, lb, ub, &nm);
    n_mods += nm; if (nm > 0) changed = 1;
    if (st == ORC_INFEASIBLE) { verdict = ORC_INFEASIBLE; break; }
    /* SolveError: the sweep stopped early; the loop condition only tests Infeasible */
    nm = 0;
    st = orc_nl_fix_obj_bins(obj, lb, ub, &nm);      /* :1045-1050, only with an incumbent */
    n_mods += nm; if (nm > 0) changed = 1;
    if (st == ORC_INFEASIBLE) { verdict = ORC_INFEASIBLE; break; }
  }
  res->verdict = verdict; res->rounds = rounds; res->n_mods = n_mods; res->nnz_updates = 0;
}

void orc_nl_simple_presolve(const orc_nl_t *g, double *lb, double *ub, orc_result_t *res)
{
  orc_nl_simple_presolve_obj(g, NULL, lb, ub, res);
}

/* ref: PCBProcessor.cpp:134-175 presolveNode_: handlers in order, stop at first infeasible */
void orc_node_presolve(const orc_lin_t *p, const orc_nl_t *g, double *lb, double *ub,
                       orc_result_t *res)
{
  orc_result_t a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
  if (p) orc_lin_simple_presolve(p, lb, ub, &a);
  if (a.verdict == ORC_OK && g) orc_nl_simple_presolve_obj(g, p, lb, ub, &b);
  res->verdict = (a.verdict != ORC_OK) ? a.verdict : b.verdict;
  res->rounds = a.rounds + b.rounds;
  res->nnz_updates = a.nnz_updates; res->n_mods = a.n_mods + b.n_mods;
}

/* --- batches of node boxes given as branching deltas on a root box ------------------------------------------
 * TEST / BASELINE DRIVER (no reference counterpart: the reference tightens one node at a time).  Box b is the
 * root box with deltas dptr[b] .. dptr[b+1) applied in order (a later delta overrides), exactly like
 * mntr_gpu_tighten_nodes.  mode 0 = orc_lin_fixpoint_inplace, 1 = orc_lin_simple_presolve, 2 = orc_node_presolve.
 * Boxes are dealt to n_threads OpenMP threads (each with its own dense box).  Per box: verdict, rounds,
 * nnz-updates and the bound changes relative to the box's INITIAL bounds, ascending (variable, side), at most
 * mod_cap of them in slot b of mod_var / mod_up / mod_val (mod_cnt[b] is the true count).  Returns the wall
 * time of the parallel region in seconds. */
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
double orc_batch_deltas(const orc_lin_t *p, const orc_nl_t *g, int32_t mode, int32_t n_boxes,
                        const double *root_lb, const double *root_ub, const int64_t *dptr,
                        const int32_t *dvar, const uint8_t *dup, const double *dval, int32_t n_threads,
                        int32_t *verdict, int32_t *rounds, int64_t *nnz, int32_t mod_cap, int32_t *mod_cnt,
                        int32_t *mod_var, uint8_t *mod_up, double *mod_val)
{
  const int32_t n = p ? p->n : 0;
  struct timespec t0, t1;
  if (n_threads < 1) n_threads = 1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel num_threads(n_threads)
  {
    double *lb = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *ub = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *l0 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *u0 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
    for (int32_t b = 0; b < n_boxes; ++b) {
      memcpy(lb, root_lb, sizeof(double) * (size_t)n);
      memcpy(ub, root_ub, sizeof(double) * (size_t)n);
      for (int64_t q = dptr[b]; q < dptr[b + 1]; ++q) {
        if (dup[q]) ub[dvar[q]] = dval[q]; else lb[dvar[q]] = dval[q];
      }
      memcpy(l0, lb, sizeof(double) * (size_t)n);
      memcpy(u0, ub, sizeof(double) * (size_t)n);
      orc_result_t r = {0, 0, 0, 0};
      if (mode == 0) orc_lin_fixpoint_inplace(p, lb, ub, &r);
      else if (mode == 1) orc_lin_simple_presolve(p, lb, ub, &r);
      else orc_node_presolve((p && p->m > 0) ? p : NULL, g, lb, ub, &r);
      if (verdict) verdict[b] = r.verdict;
      if (rounds) rounds[b] = r.rounds;
      if (nnz) nnz[b] = r.nnz_updates;
      if (mod_cnt) {
        int32_t cnt = 0;
        for (int32_t j = 0; j < n; ++j) {
          if (lb[j] != l0[j]) {
            if (cnt < mod_cap) { mod_var[(size_t)b * mod_cap + cnt] = j; mod_up[(size_t)b * mod_cap + cnt] = 0; mod_val[(size_t)b * mod_cap + cnt] = lb[j]; }
            ++cnt;
          }
          if (ub[j] != u0[j]) {
            if (cnt < mod_cap) { mod_var[(size_t)b * mod_cap + cnt] = j; mod_up[(size_t)b * mod_cap + cnt] = 1; mod_val[(size_t)b * mod_cap + cnt] = ub[j]; }
            ++cnt;
          }
        }
        mod_cnt[b] = cnt;
      }
    }
    free(lb); free(ub); free(l0); free(u0);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* --- root presolve row operations (LinearHandler::presolve) ------------------------------------------------------ */

/* ref: LinearHandler.cpp:882-949 dupRows_, the detection part: row hashes h = Constraint::getActivity(r) (terms in
 * ascending variable id, value += x * coef, LinearFunction.cpp:151-158) for two random vectors, then every pair i < j
 * through the reference's tests.  Returns the number of candidate pairs (stored up to cap, in (i, j) order -- the order
 * of the reference's double loop).  kind 1: treatDupRows_ is called with mult 1.0, kind 2: with h1i / h1j.
 * The reference also skips rows deleted earlier in the same loop; that depends on treatDupRows_ (the object graph) and
 * is applied by the caller while walking the list. */
int64_t orc_root_dup_rows(const orc_lin_t *p, const double *r1, const double *r2, double *h1, double *h2, int64_t cap,
                          int32_t *pair_i, int32_t *pair_j, uint8_t *pair_kind)
{
  const int32_t m = p->m;
  for (int32_t i = 0; i < m; ++i) {
    if (p->row_active && !p->row_active[i]) { h1[i] = h2[i] = 1e30; continue; }
    double a = 0, b = 0;
    for (int32_t t = p->row_ptr[i]; t < p->row_ptr[i + 1]; ++t) {
      if (fabs(p->val[t]) <= 1e-9) continue;                 /* LinearFunction::addTerm drops these */
      a += r1[p->col[t]] * p->val[t];
      b += r2[p->col[t]] * p->val[t];
    }
    h1[i] = a; h2[i] = b;
  }
  int64_t k = 0;
  for (int32_t i = 0; i < m; ++i) {
    if (!(h1[i] < 1e29)) continue;
    for (int32_t j = i + 1; j < m; ++j) {
      int kind = 0;
      if (fabs(h1[j] - h1[i]) < 1e-10 || fabs(h1[j] + h1[i]) < 1e-10) kind = 1;
      else if (h1[j] < 1e29 && fabs(h1[i] / h1[j] - h2[i] / h2[j]) < 1e-10) kind = 2;
      if (kind) {
        if (k < cap) { pair_i[k] = i; pair_j[k] = j; pair_kind[k] = (uint8_t)kind; }
        ++k;
      }
    }
  }
  return k;
}

/* ref: LinearHandler.cpp:974-985: a row is redundant on a box when getLfBnds_ gives ll >= lb - eTol && uu <= ub + eTol */
int64_t orc_root_redundant_rows(const orc_lin_t *p, const double *lb, const double *ub, uint8_t *redundant)
{
  int64_t k = 0;
  for (int32_t i = 0; i < p->m; ++i) {
    redundant[i] = 0;
    if (p->row_active && !p->row_active[i]) continue;
    double ll, uu;
    int32_t b = p->row_ptr[i];
    lf_bnds(p->row_ptr[i + 1] - b, p->col + b, p->val + b, lb, ub, &ll, &uu);
    if (ll >= p->row_lb[i] - E_TOL && uu <= p->row_ub[i] + E_TOL) { redundant[i] = 1; ++k; }
  }
  return k;
}


/* ---- LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) with its implications computeImpBounds_ (:707-783) ----
 * Sequential, in the reference's order: rows by index, the terms of a row by variable id, the first binary that can
 * be improved ends the row (`break`).  Works on private copies of the coefficients and row bounds, which it edits as
 * the reference edits the problem (LinearFunction::incTerm :133-142: a change of at most 1e-9 is ignored, a result
 * below 1e-9 erases the term; Problem::changeBound on the row), so later rows see earlier improvements through the
 * 2-term rows their implications read.  Terms with |a| <= 1e-9 do not exist (LinearFunction::addTerm :89-95).
 * Output per improved row: the row, the variable, the NEW coefficient, which row bound moved (0 none, 1 lower,
 * 2 upper) and its new value.  Returns the number of improved rows (counted beyond cap too). */
static int ci_is_term(double a) { return fabs(a) > 1e-9; }

typedef struct {
  const orc_lin_t *p;
  double *val, *rlb, *rub;      /* edited copies */
  const int32_t *cptr, *crow, *cpos;   /* CSC: rows of a variable and the entry's position in the row */
  double *lb, *ub;              /* variable bounds (edited temporarily by the implications) */
} ci_state;

static int ci_num_terms(const ci_state *s, int32_t r)
{
  int k = 0;
  for (int32_t t = s->p->row_ptr[r]; t < s->p->row_ptr[r + 1]; ++t) k += ci_is_term(s->val[t]);
  return k;
}

/* weight of variable j in row r (0 when absent) */
static double ci_weight(const ci_state *s, int32_t r, int32_t j)
{
  for (int32_t t = s->p->row_ptr[r]; t < s->p->row_ptr[r + 1]; ++t)
    if (s->p->col[t] == j) return ci_is_term(s->val[t]) ? s->val[t] : 0.0;
  return 0.0;
}

static void ci_lf_bounds(const ci_state *s, int32_t r, double *l, double *u)     /* LinearFunction::computeBounds :178-195 */
{
  double lo = 0.0, up = 0.0;
  for (int32_t t = s->p->row_ptr[r]; t < s->p->row_ptr[r + 1]; ++t) {
    const double a = s->val[t];
    if (!ci_is_term(a)) continue;
    const int32_t j = s->p->col[t];
    if (a > 0) { lo += a * s->lb[j]; up += a * s->ub[j]; }
    else       { lo += a * s->ub[j]; up += a * s->lb[j]; }
  }
  *l = lo; *u = up;
}

/* computeImpBounds_ :707-783: activity of row c with z fixed at zval and every other variable of the row tightened
 * by the 2-term rows it shares with z */
static void ci_imp_bounds(ci_state *s, int32_t c, int32_t z, double zval, double *out_l, double *out_u)
{
  const orc_lin_t *p = s->p;
  const double zl0 = s->lb[z], zu0 = s->ub[z];
  if (zval < 0.5) s->ub[z] = 0.0; else s->lb[z] = 1.0;
  /* the other variables: their tightened bounds are applied after each is computed; a variable's implied bounds
   * depend on its own bounds and on z alone, so the order does not matter */
  const int32_t b = p->row_ptr[c], e = p->row_ptr[c + 1];
  double *sl = (double *)malloc(sizeof(double) * (size_t)(e - b + 1)), *su = (double *)malloc(sizeof(double) * (size_t)(e - b + 1));
  for (int32_t t = b; t < e; ++t) {
    const int32_t v = p->col[t];
    sl[t - b] = s->lb[v]; su[t - b] = s->ub[v];
    if (!ci_is_term(s->val[t]) || v == z) continue;
    double l1 = s->lb[v], u1 = s->ub[v];
    for (int32_t q = s->cptr[v]; q < s->cptr[v + 1]; ++q) {
      const int32_t c2 = s->crow[q];
      if (p->row_active && !p->row_active[c2]) continue;
      if (ci_num_terms(s, c2) != 2) continue;
      const double b2 = ci_weight(s, c2, z);
      if (b2 == 0.0) continue;                       /* lf2->hasVar(z) */
      const double a2 = ci_weight(s, c2, v);
      if (a2 == 0.0) continue;
      const double cub = s->rub[c2], clb = s->rlb[c2];
      if (a2 > 0 && (cub - zval * b2) / a2 < u1) u1 = (cub - zval * b2) / a2;
      if (a2 < 0 && (cub - zval * b2) / a2 > l1) l1 = (cub - zval * b2) / a2;
      if (a2 > 0 && (clb - zval * b2) / a2 > l1) l1 = (clb - zval * b2) / a2;
      if (a2 < 0 && (clb - zval * b2) / a2 < u1) u1 = (clb - zval * b2) / a2;
    }
    if (l1 > s->lb[v]) s->lb[v] = l1;
    if (u1 < s->ub[v]) s->ub[v] = u1;
  }
  ci_lf_bounds(s, c, out_l, out_u);
  for (int32_t t = b; t < e; ++t) { s->lb[p->col[t]] = sl[t - b]; s->ub[p->col[t]] = su[t - b]; }
  s->lb[z] = zl0; s->ub[z] = zu0;
  free(sl); free(su);
}

int64_t orc_root_coeff_imp(const orc_lin_t *p, const double *lb_in, const double *ub_in, int64_t cap, int32_t *out_row,
                           int32_t *out_var, double *out_coef, int32_t *out_side, double *out_bnd)
{
  const int32_t m = p->m, n = p->n;
  const int64_t nnz = p->row_ptr[m];
  const double coeftol = 1e-4, bslack = 1e-4;
  ci_state s;
  s.p = p;
  s.val = (double *)malloc(sizeof(double) * (size_t)(nnz + 1));
  s.rlb = (double *)malloc(sizeof(double) * (size_t)(m + 1)); s.rub = (double *)malloc(sizeof(double) * (size_t)(m + 1));
  s.lb = (double *)malloc(sizeof(double) * (size_t)(n + 1)); s.ub = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  memcpy(s.val, p->val, sizeof(double) * (size_t)nnz);
  memcpy(s.rlb, p->row_lb, sizeof(double) * (size_t)m); memcpy(s.rub, p->row_ub, sizeof(double) * (size_t)m);
  memcpy(s.lb, lb_in, sizeof(double) * (size_t)n); memcpy(s.ub, ub_in, sizeof(double) * (size_t)n);
  int32_t *cptr = (int32_t *)calloc((size_t)n + 2, sizeof(int32_t)), *crow = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1)),
          *cpos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
  for (int64_t t = 0; t < nnz; ++t) cptr[p->col[t] + 2]++;
  for (int32_t j = 0; j < n; ++j) cptr[j + 2] += cptr[j + 1];
  for (int32_t i = 0; i < m; ++i)
    for (int32_t t = p->row_ptr[i]; t < p->row_ptr[i + 1]; ++t) { const int32_t q = cptr[p->col[t] + 1]++; crow[q] = i; cpos[q] = t; }
  s.cptr = cptr; s.crow = crow; s.cpos = cpos;

  int64_t k = 0;
  for (int32_t c = 0; c < m; ++c) {
    if (p->row_active && !p->row_active[c]) continue;
    if (!(s.rlb[c] <= -INFINITY || s.rub[c] >= INFINITY)) continue;
    const int nt = ci_num_terms(&s, c);
    if (nt < 2) continue;
    const int implic = nt < 50;
    const double lb = s.rlb[c], ub = s.rub[c];
    double ll, uu;
    ci_lf_bounds(&s, c, &ll, &uu);                  /* getLfBnds_ :1237-1258: the same sums */
    for (int32_t t = p->row_ptr[c]; t < p->row_ptr[c + 1]; ++t) {
      const double a0 = s.val[t];
      if (!ci_is_term(a0)) continue;
      const int32_t v = p->col[t];
      const int ty = p->var_type[v];
      if (!((ty == ORC_BINARY || ty == ORC_IMPLBIN) && s.ub[v] > s.lb[v] + 0.5)) continue;
      double delta = 0.0, nb = 0.0; int side = 0, hit = 0;
      if (implic) {
        ll = uu = 0;
        ci_imp_bounds(&s, c, v, 1.0, &ll, &uu);
        ll -= bslack; uu += bslack;
        if (a0 > 0) ll -= a0; else uu -= a0;
      }
      if (uu + a0 < ub - coeftol && uu >= ub) { delta = ub - uu - a0; hit = 1; }
      else if (ll + a0 > lb + coeftol && ll <= lb) { delta = lb - ll - a0; hit = 1; }
      if (!hit) {
        if (implic) {
          ll = uu = 0;
          ci_imp_bounds(&s, c, v, 0.0, &ll, &uu);
          if (a0 > 0) uu += a0; else ll += a0;
        }
        if (uu - a0 < ub - coeftol && uu >= ub) { delta = uu - a0 - ub; side = 2; nb = uu - a0; hit = 1; }
        else if (ll - a0 > lb + coeftol && ll <= lb) { delta = ll - a0 - lb; side = 1; nb = ll - a0; hit = 1; }
      }
      if (!hit) continue;
      /* lf->incTerm(v, delta) */
      double nv = a0;
      if (fabs(delta) > 1e-9) { nv = a0 + delta; if (fabs(nv) < 1e-9) nv = 0.0; }
      s.val[t] = nv;
      if (side == 2) s.rub[c] = nb; else if (side == 1) s.rlb[c] = nb;
      if (k < cap) { out_row[k] = c; out_var[k] = v; out_coef[k] = nv; out_side[k] = side; out_bnd[k] = nb; }
      ++k;
      break;
    }
  }
  free(s.val); free(s.rlb); free(s.rub); free(s.lb); free(s.ub); free(cptr); free(crow); free(cpos);
  return k;
}


/* ---- QuadHandler::simplePresolve (QuadHandler.cpp:1146-1201): ONE in-place sweep over the relations y = x^2 (x2Funs_,
 *      a map keyed by x: ascending variable id) and then y = x0 * x1 (x0x1Funs_, ordered by (x0, x1), x0 < x1 by
 *      index: LinBil.cpp:26-35, 57-68), every step through updatePBounds_ (QuadHandler.cpp:3218-3246).  The status the
 *      reference computes is overwritten with Finished at the end (:1200): nothing is ever reported infeasible; the
 *      number of steps that found inconsistent bounds is returned in *n_inconsistent for information. ---- */
static int qh_update(const uint8_t *var_type, double *lbv, double *ubv, int32_t v, double lb, double ub, int64_t *n_mods)
{
  const double aTol = 1e-6, bTol = 1e-8, rTol = 1e-7;
  const int ty = var_type[v];
  if (ty == ORC_BINARY || ty == ORC_IMPLBIN || ty == ORC_INTEGER || ty == ORC_IMPLINT) { ub = floor(ub); lb = ceil(lb); }
  if (ub < lbv[v] - bTol || lb > ubv[v] + bTol) return -1;
  if (ub < ubv[v] - bTol && (ubv[v] == INFINITY || ub < ubv[v] - fabs(ubv[v]) * rTol)) { ubv[v] = ub; ++*n_mods; }
  if (lb > lbv[v] + aTol && (lbv[v] == -INFINITY || lb > lbv[v] + fabs(lbv[v]) * rTol)) { lbv[v] = lb; ++*n_mods; }
  return 0;
}

int64_t orc_quad_simple_presolve(int32_t n_sq, const int32_t *sq_x, const int32_t *sq_y, int32_t n_bil, const int32_t *b_x0,
                                 const int32_t *b_x1, const int32_t *b_y, const uint8_t *var_type, double *lbv, double *ubv,
                                 int64_t *n_inconsistent)
{
  const double bTol = 1e-8;
  int64_t n_mods = 0, bad = 0;
  for (int32_t k = 0; k < n_sq; ++k) {
    const int32_t x = sq_x[k], y = sq_y[k];
    double lb, ub;
    orc_bounds_on_square(lbv[x], ubv[x], &lb, &ub);
    if (qh_update(var_type, lbv, ubv, y, lb, ub, &n_mods) < 0) ++bad;
    if (ubv[y] > bTol) {
      ub = sqrt(ubv[y]);
      lb = -ub;
      if (lbv[x] > -sqrt(lbv[y]) + bTol) lb = sqrt(lbv[y]);
      if (qh_update(var_type, lbv, ubv, x, lb, ub, &n_mods) < 0) ++bad;
    } else if (ubv[y] < -bTol) {
      ++bad;
    } else {
      if (qh_update(var_type, lbv, ubv, x, 0.0, 0.0, &n_mods) < 0) ++bad;
    }
  }
  for (int32_t k = 0; k < n_bil; ++k) {
    const int32_t x1 = b_x0[k], x2 = b_x1[k], y = b_y[k];
    double lb, ub;
    orc_bounds_on_product(1, lbv[x1], ubv[x1], lbv[x2], ubv[x2], &lb, &ub);
    if (qh_update(var_type, lbv, ubv, y, lb, ub, &n_mods) < 0) ++bad;
    orc_bounds_on_div(lbv[y], ubv[y], lbv[x1], ubv[x1], &lb, &ub);
    if (qh_update(var_type, lbv, ubv, x2, lb, ub, &n_mods) < 0) ++bad;
    orc_bounds_on_div(lbv[y], ubv[y], lbv[x2], ubv[x2], &lb, &ub);
    if (qh_update(var_type, lbv, ubv, x1, lb, ub, &n_mods) < 0) ++bad;
  }
  if (n_inconsistent) *n_inconsistent = bad;
  return n_mods;
}


/* ---- QuadHandler::presolveNode's propagation loop (QuadHandler.cpp:1214-1239): sweeps { propSqrBnds_ over x2Funs_
 *      (:1361-1395), propBilBnds_ over x0x1Funs_ (:1271-1301) } in place, repeated while a sweep changed a bound; the
 *      first inconsistent step ends the call with "infeasible" (:1224, :1233).  Every step goes through the
 *      relaxation-aware updatePBounds_ (:3248-3320): integer rounding, consistency test with bTol, then
 *        both sides move   if each side improves by more than bTol AND by more than rTol relative   (one VarBoundMod2),
 *        else the lower    if it does                                                                (one VarBoundMod),
 *        else the upper    if it does                                                                (one VarBoundMod)
 *      -- i.e. each side moves exactly when its own two tests pass; the three-way split only decides how many
 *      Modification objects are pushed.  (The Problem-side variant of simplePresolve uses aTol = 1e-6 for the lower
 *      bound; this one uses bTol = 1e-8 for both.)  What follows the loop in presolveNode -- tightenQuad_ on the first
 *      call (:1241-1250) and the McCormick row refresh upSqCon_/upBilCon_ (:1252-1257) -- is not part of this path.
 *      Returns 1 when infeasible (the box is then undefined: the reference stops mid-sweep), else 0;
 *      *n_mods = Modification objects pushed, *n_sweeps = pStats_.iters increments.  max_sweeps <= 0: no cap. ---- */
static int qh_update_node(const uint8_t *var_type, double *lbv, double *ubv, int32_t v, double lb, double ub, int64_t *n_mods,
                          int *changed)
{
  const double bTol = 1e-8, rTol = 1e-7;
  const int ty = var_type[v];
  if (ty == ORC_BINARY || ty == ORC_IMPLBIN || ty == ORC_INTEGER || ty == ORC_IMPLINT) { ub = floor(ub); lb = ceil(lb); }
  if (lb > ubv[v] + bTol || ub < lbv[v] - bTol) return -1;
  const int lo = lb > lbv[v] + bTol && (lbv[v] == -INFINITY || lb > lbv[v] + rTol * fabs(lbv[v]));
  const int up = ub < ubv[v] - bTol && (ubv[v] == INFINITY || ub < ubv[v] - rTol * fabs(ubv[v]));
  if (lo && up) { lbv[v] = lb; ubv[v] = ub; }
  else if (lo) lbv[v] = lb;
  else if (up) ubv[v] = ub;
  if (lo || up) { ++*n_mods; *changed = 1; }
  return 0;
}

int32_t orc_quad_presolve_node(int32_t n_sq, const int32_t *sq_x, const int32_t *sq_y, int32_t n_bil, const int32_t *b_x0,
                               const int32_t *b_x1, const int32_t *b_y, const uint8_t *var_type, double *lbv, double *ubv,
                               int32_t max_sweeps, int64_t *n_mods_out, int32_t *n_sweeps_out)
{
  const double bTol = 1e-8;
  int64_t n_mods = 0;
  int32_t sweeps = 0, inf = 0;
  int changed = 1;
  while (changed && !inf && (max_sweeps <= 0 || sweeps < max_sweeps)) {
    ++sweeps;
    changed = 0;
    for (int32_t k = 0; k < n_sq && !inf; ++k) {                                       /* propSqrBnds_ :1361-1395 */
      const int32_t x = sq_x[k], y = sq_y[k];
      double lb, ub;
      orc_bounds_on_square(lbv[x], ubv[x], &lb, &ub);
      if (qh_update_node(var_type, lbv, ubv, y, lb, ub, &n_mods, &changed) < 0) { inf = 1; break; }
      if (ubv[y] > bTol) {
        ub = sqrt(ubv[y]);
        lb = -ub;
        if (lbv[x] > -sqrt(lbv[y]) + bTol) lb = sqrt(lbv[y]);
        if (qh_update_node(var_type, lbv, ubv, x, lb, ub, &n_mods, &changed) < 0) inf = 1;
      } else if (ubv[y] < -bTol) {
        inf = 1;
      } else {
        if (qh_update_node(var_type, lbv, ubv, x, 0.0, 0.0, &n_mods, &changed) < 0) inf = 1;
      }
    }
    for (int32_t k = 0; k < n_bil && !inf; ++k) {                                      /* propBilBnds_ :1271-1301 */
      const int32_t x0 = b_x0[k], x1 = b_x1[k], y = b_y[k];
      double lb, ub;
      orc_bounds_on_product(1, lbv[x0], ubv[x0], lbv[x1], ubv[x1], &lb, &ub);
      if (qh_update_node(var_type, lbv, ubv, y, lb, ub, &n_mods, &changed) < 0) { inf = 1; break; }
      orc_bounds_on_div(lbv[y], ubv[y], lbv[x0], ubv[x0], &lb, &ub);
      if (qh_update_node(var_type, lbv, ubv, x1, lb, ub, &n_mods, &changed) < 0) { inf = 1; break; }
      orc_bounds_on_div(lbv[y], ubv[y], lbv[x1], ubv[x1], &lb, &ub);
      if (qh_update_node(var_type, lbv, ubv, x0, lb, ub, &n_mods, &changed) < 0) inf = 1;
    }
  }
  if (n_mods_out) *n_mods_out = n_mods;
  if (n_sweeps_out) *n_sweeps_out = sweeps;
  return inf;
}
