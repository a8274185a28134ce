/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A small C-ABI driver around the REFERENCE'S OWN objects (compiled by oracle/Makefile
 * from /root/reference/src/base/*.cpp, never copied).  It builds a Minotaur::Problem
 * from the same flat CSR / tape description the CUDA path takes and runs the
 * reference's LinearHandler / NlPresHandler / CGraph code on it.  Output goes only to
 * oracle/_ref/libminotaur_ref.so.
 *
 * Used (a) to pin oracle/fbbt_oracle.c, (b) to generate tests/golden/ fixtures,
 * (c) as the "reference" CPU baseline of bench.py.
 */
#include "MinotaurConfig.h"

// NlPresHandler::chkRed_ is private (NlPresHandler.h:226); the harness calls it directly to pin the oracle's
// QuadraticFunction check in isolation (the handler's public simplePresolve also runs the qf branch of varBndsFromCons_,
// which writes to the ORIGINAL problem -- a reference defect).  The objects were compiled from the
// untouched headers: access specifiers do not change the layout.
#define private public
#include "NlPresHandler.h"
#include "QuadHandler.h"
#undef private

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "CGraph.h"
#include "CNode.h"
#include "Constraint.h"
#include "Environment.h"
#include "Function.h"
#include "LinearFunction.h"
#include "LinearHandler.h"
#include "Logger.h"
#include "NlPresHandler.h"
#include "NonlinearFunction.h"
#include "Objective.h"
#include "Option.h"
#include "Problem.h"
#include "QuadraticFunction.h"
#include "Reader.h"
#include "Relaxation.h"
#include "SolutionPool.h"
#include "Types.h"
#include "VarBoundMod.h"
#include "Variable.h"

using namespace Minotaur;

namespace {

/* exposes the protected per-round routines of LinearHandler (LinearHandler.h:154-273) */
class LinProbe : public LinearHandler {
public:
  LinProbe(EnvPtr env, ProblemPtr p) : LinearHandler(env, p) {}

  void flagAll(ProblemPtr p)
  {
    for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) (*it)->setBFlag(true);
  }

  /* cut-off value of the incumbent, as simplePresolve forms it (LinearHandler.cpp:1636-1640) */
  static bool cutoff(ProblemPtr p, SolutionPool *spool, double *ub)
  {
    if (!spool || spool->getNumSols() == 0 || !p->getObjective()) return false;
    *ub = spool->getBestSolutionValue() - p->getObjective()->getConstant();
    return true;
  }

  /* SURVEY.md section 8c parity driver: the reference's sweeps, status honoured, to fixpoint */
  int fixpoint(ProblemPtr p, SolutionPool *spool, int *rounds, int64_t *nmods)
  {
    ModQ mods; bool ch = true; UInt ni = 0; int inf = 0; *rounds = 0;
    double cut = 0;
    flagAll(p);
    while (ch && !inf) {
      ch = false; ++*rounds;
      if (varBndsFromCons_(p, false, &ch, &mods, &ni) == SolvedInfeasible) { inf = 1; break; }
      if (cutoff(p, spool, &cut) && varBndsFromObj_(p, cut, false, &ch, &mods) == SolvedInfeasible) { inf = 1; break; }
      tightenInts_(p, false, &ch, &mods);
      if (checkBounds_(p) == SolvedInfeasible) inf = 1;
    }
    *nmods = (int64_t)mods.size();
    for (ModQ::iterator it = mods.begin(); it != mods.end(); ++it) delete *it;
    return inf;
  }

  /* same driver with the row loop of varBndsFromCons_ (LinearHandler.cpp:506-539,
   * apply_to_prob == false) unrolled here so the visited nnz can be counted */
  int fixpointCounted(ProblemPtr p, SolutionPool *spool, int *rounds, int64_t *nmods, int64_t *nnz, int max_rounds)
  {
    ModQ mods; bool ch = true; UInt ni = 0; int inf = 0; *rounds = 0; *nnz = 0;
    double cut = 0;
    flagAll(p);
    while (ch && !inf && (max_rounds <= 0 || *rounds < max_rounds)) {
      ch = false; ++*rounds;
      for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) {
        ConstraintPtr c = *it;
        if (c->getBFlag() && c->getFunctionType() == Linear && c->getQuadraticFunction() == 0 &&
            c->getNonlinearFunction() == 0 && DeletedCons != c->getState()) {
          bool t = false;
          c->setBFlag(false);
          *nnz += c->getLinearFunction()->getNumTerms();
          if (linBndTighten_(p, false, c, &t, &mods, &ni) == SolvedInfeasible) { inf = 1; break; }
          if (t) ch = true;
        }
      }
      if (inf) break;
      if (cutoff(p, spool, &cut)) {
        /* the loop of varBndsFromObj_ (LinearHandler.cpp:565-593) spelled out with the reference's own
         * pieces, so that the terms it visits can be counted */
        ObjectivePtr o = p->getObjective();
        LinearFunctionPtr lf = o->getLinearFunction();
        if (lf && o->getFunctionType() == Linear) {
          bool t = true; UInt unused = 0;
          while (t && !inf) {
            double ll, uu, sll = INFINITY, suu = INFINITY;
            t = false;
            getLfBnds_(lf, &ll, &uu);
            if (ll < -1e20 || uu > 1e20) getSingLfBnds_(lf, &sll, &suu);
            *nnz += lf->getNumTerms();
            if (ll > cut + 1e-8) { inf = 1; break; }
            if (ll > -1e20) updateLfBoundsFromUb_(p, false, lf, cut, ll, false, &t, &mods, &unused);
            else if (sll > -1e20) updateLfBoundsFromUb_(p, false, lf, cut, sll, true, &t, &mods, &unused);
            if (t) ch = true;
          }
        }
      }
      if (inf) break;
      tightenInts_(p, false, &ch, &mods);
      if (checkBounds_(p) == SolvedInfeasible) inf = 1;
    }
    *nmods = (int64_t)mods.size();
    for (ModQ::iterator it = mods.begin(); it != mods.end(); ++it) delete *it;
    return inf;
  }

  /* LinearHandler::dupRows_ itself (protected, :882-949).  It draws its two random vectors with rand(): the caller
   * seeds the generator, so that the same vectors can be drawn again outside (drawDupVectors). */
  void dupRows(bool *changed) { dupRows_(changed); }
  bool treatDup(ConstraintPtr c1, ConstraintPtr c2, double mult, bool *changed) { return treatDupRows_(c1, c2, mult, changed); }
  void coeffImp(bool *changed) { coeffImp_(changed); }
  bool redundantOnBox(ConstraintPtr c)
  {
    double ll, uu;
    getLfBnds_(c->getLinearFunction(), &ll, &uu);
    return ll >= c->getLb() - 1e-8 && uu <= c->getUb() + 1e-8;       /* :974 */
  }

  void rowActivity(ConstraintPtr c, double out[4])
  {
    LinearFunctionPtr lf = c->getLinearFunction();
    getLfBnds_(lf, &out[0], &out[1]);
    out[2] = -INFINITY; out[3] = INFINITY;
    getSingLfBnds_(lf, &out[2], &out[3]);
  }
};

struct RefProblem {
  EnvPtr env = 0;
  ProblemPtr p = 0;
  LinProbe *lh = 0;
  NlPresHandler *nh = 0;
  SolutionPool *spool = 0;       /* holds the incumbent when a cut-off is set */
  std::vector<VariablePtr> vars;
  std::vector<ConstraintPtr> lin_rows;
  std::vector<ConstraintPtr> nl_rows;
  std::vector<ConstraintPtr> quad_rows;
  std::vector<CGraph *> graphs;
};

void freeMods(ModVector &mods)
{
  for (ModVector::iterator it = mods.begin(); it != mods.end(); ++it) delete *it;
  mods.clear();
}

}  // namespace

extern "C" {

void *ref_create(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col,
                 const double *val, const double *row_lb, const double *row_ub,
                 const uint8_t *var_type, const double *lb, const double *ub)
{
  RefProblem *h = new RefProblem();
  int err = 0;
  h->env = (EnvPtr) new Environment();
  h->env->startTimer(err);
  h->env->setLogLevel(LogNone);
  h->p = (ProblemPtr) new Problem(h->env);
  h->vars.reserve(n);
  for (int32_t j = 0; j < n; ++j)
    h->vars.push_back(h->p->newVariable(lb[j], ub[j], (VariableType)var_type[j]));
  h->lin_rows.reserve(m);
  for (int32_t i = 0; i < m; ++i) {
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    for (int32_t t = row_ptr[i]; t < row_ptr[i + 1]; ++t) lf->addTerm(h->vars[col[t]], val[t]);
    FunctionPtr f = (FunctionPtr) new Function(lf);
    h->lin_rows.push_back(h->p->newConstraint(f, row_lb[i], row_ub[i]));
  }
  return h;
}

/* one nonlinear constraint  c_lb <= tape(x) + lin.x <= c_ub ; tape in the flat layout
 * of fbbt_oracle.h (local indices).  Returns its index among nonlinear constraints. */
int32_t ref_add_nl(void *hv, int32_t n_nodes, const uint8_t *op, const int32_t *arg0,
                   const int32_t *arg1, const double *cnst, const int32_t *child, int32_t lin_k,
                   const int32_t *lin_col, const double *lin_val, double c_lb, double c_ub)
{
  RefProblem *h = (RefProblem *)hv;
  CGraph *cg = new CGraph();
  std::vector<CNode *> nodes(n_nodes, (CNode *)0);
  for (int32_t i = 0; i < n_nodes; ++i) {
    OpCode o = (OpCode)op[i];
    if (o == OpVar) nodes[i] = cg->newNode(h->vars[arg0[i]]);
    else if (o == OpNum) nodes[i] = cg->newNode(cnst[i]);
    else if (o == OpInt) nodes[i] = cg->newNode((int)cnst[i]);
    else if (o == OpSumList) {
      std::vector<CNode *> ch;
      for (int32_t c = arg0[i]; c < arg1[i]; ++c) ch.push_back(nodes[child[c]]);
      nodes[i] = cg->newNode(OpSumList, &ch[0], (UInt)ch.size());
    } else {
      nodes[i] = cg->newNode(o, nodes[arg0[i]], arg1[i] >= 0 ? nodes[arg1[i]] : (CNode *)0);
    }
  }
  cg->setOut(nodes[n_nodes - 1]);
  cg->finalize();
  LinearFunctionPtr lf = 0;
  if (lin_k > 0) {
    lf = (LinearFunctionPtr) new LinearFunction();
    for (int32_t t = 0; t < lin_k; ++t) lf->addTerm(h->vars[lin_col[t]], lin_val[t]);
  }
  FunctionPtr f = (FunctionPtr) new Function(lf, (NonlinearFunctionPtr)cg);
  h->nl_rows.push_back(h->p->newConstraint(f, c_lb, c_ub));
  h->graphs.push_back(cg);
  return (int32_t)h->nl_rows.size() - 1;
}

/* one constraint  lb <= sum_k coef_k x_{v1_k} x_{v2_k} + lin.x <= ub  with a QuadraticFunction */
int32_t ref_add_quad(void *hv, int32_t k, const int32_t *v1, const int32_t *v2, const double *coef, int32_t lin_k,
                     const int32_t *lin_col, const double *lin_val, double lb, double ub)
{
  RefProblem *h = (RefProblem *)hv;
  QuadraticFunctionPtr qf = (QuadraticFunctionPtr) new QuadraticFunction();
  for (int32_t t = 0; t < k; ++t) qf->addTerm(h->vars[v1[t]], h->vars[v2[t]], coef[t]);
  LinearFunctionPtr lf = 0;
  if (lin_k > 0) {
    lf = (LinearFunctionPtr) new LinearFunction();
    for (int32_t t = 0; t < lin_k; ++t) lf->addTerm(h->vars[lin_col[t]], lin_val[t]);
  }
  FunctionPtr f = (FunctionPtr) new Function(lf, qf, (NonlinearFunctionPtr)0);
  h->quad_rows.push_back(h->p->newConstraint(f, lb, ub));
  return (int32_t)h->quad_rows.size() - 1;
}

/* The reference's own QuadHandler::simplePresolve (QuadHandler.cpp:1146-1201) on a fresh problem of n variables with
 * the given types and box: the relations y = x^2 and y = x0 x1 are handed to QuadHandler::addConstraint (:127-179) as
 * constraints  y - x0 x1 = 0  (one linear term, one quadratic term); the handler orders them itself (x2Funs_ by x,
 * x0x1Funs_ by (x0, x1)).  lb / ub are updated in place; returns the number of modifications. */
int64_t ref_quad_simple_presolve(int32_t n, const uint8_t *var_type, double *lb, double *ub, int32_t n_sq, const int32_t *sq_x,
                                 const int32_t *sq_y, int32_t n_bil, const int32_t *b_x0, const int32_t *b_x1, const int32_t *b_y)
{
  EnvPtr env = (EnvPtr) new Environment();
  int err = 0;
  env->startTimer(err);
  env->setLogLevel(LogNone);
  ProblemPtr p = (ProblemPtr) new Problem(env);
  std::vector<VariablePtr> vars;
  for (int32_t j = 0; j < n; ++j) vars.push_back(p->newVariable(lb[j], ub[j], (VariableType)var_type[j]));
  QuadHandler *qh = new QuadHandler(env, p);
  for (int32_t k = 0; k < n_sq + n_bil; ++k) {
    const bool sq = k < n_sq;
    VariablePtr y = vars[sq ? sq_y[k] : b_y[k - n_sq]];
    VariablePtr x0 = vars[sq ? sq_x[k] : b_x0[k - n_sq]], x1 = vars[sq ? sq_x[k] : b_x1[k - n_sq]];
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    lf->addTerm(y, 1.0);
    QuadraticFunctionPtr qf = (QuadraticFunctionPtr) new QuadraticFunction();
    qf->addTerm(x0, x1, -1.0);
    FunctionPtr f = (FunctionPtr) new Function(lf, qf, (NonlinearFunctionPtr)0);
    qh->addConstraint(p->newConstraint(f, 0.0, 0.0));
  }
  ModVector mods;
  SolveStatus status = Started;
  qh->simplePresolve(p, (SolutionPoolPtr)0, mods, status);
  const int64_t n_mods = (int64_t)mods.size();
  for (int32_t j = 0; j < n; ++j) { lb[j] = vars[j]->getLb(); ub[j] = vars[j]->getUb(); }
  freeMods(mods);
  delete qh;
  delete p;
  delete env;
  return n_mods;
}

/* The reference's own QuadHandler::presolveNode (QuadHandler.cpp:1204-1269) on a batch of boxes.  A problem of n
 * variables with the ROOT box (root_lb, root_ub) receives the relations through QuadHandler::addConstraint; the
 * relaxation is built the way NodeIncRelaxer does for this handler (variables cloned, then QuadHandler::relaxInitInc:
 * the McCormick rows).  For every box the bounds of the problem's and the relaxation's variables are set and
 * presolveNode is called with setModFlags(false, true), Bnb's setting.  bStats_.niters is preset to 1: the handler
 * behaves as at every node after the first, i.e. without tightenQuad_ (:1241, doQT_ is false by default), so the call
 * is the propagation loop plus the refresh of the McCormick rows (upSqCon_/upBilCon_, which do not touch variable
 * bounds).  lb/ub [n_boxes][n] are updated in place; verdict[b] = 1 when presolveNode returned true (infeasible: the
 * box is then whatever the reference had reached); n_mods[b] = p_mods.size(). */
int32_t ref_quad_presolve_node(int32_t n, const uint8_t *var_type, const double *root_lb, const double *root_ub, int32_t n_sq,
                               const int32_t *sq_x, const int32_t *sq_y, int32_t n_bil, const int32_t *b_x0, const int32_t *b_x1,
                               const int32_t *b_y, int32_t n_boxes, double *lb, double *ub, int32_t *verdict, int64_t *n_mods)
{
  EnvPtr env = (EnvPtr) new Environment();
  int err = 0;
  env->startTimer(err);
  env->setLogLevel(LogNone);
  ProblemPtr p = (ProblemPtr) new Problem(env);
  std::vector<VariablePtr> vars;
  for (int32_t j = 0; j < n; ++j) vars.push_back(p->newVariable(root_lb[j], root_ub[j], (VariableType)var_type[j]));
  QuadHandler *qh = new QuadHandler(env, p);
  for (int32_t k = 0; k < n_sq + n_bil; ++k) {
    const bool sq = k < n_sq;
    VariablePtr y = vars[sq ? sq_y[k] : b_y[k - n_sq]];
    VariablePtr x0 = vars[sq ? sq_x[k] : b_x0[k - n_sq]], x1 = vars[sq ? sq_x[k] : b_x1[k - n_sq]];
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    lf->addTerm(y, 1.0);
    QuadraticFunctionPtr qf = (QuadraticFunctionPtr) new QuadraticFunction();
    qf->addTerm(x0, x1, -1.0);
    FunctionPtr f = (FunctionPtr) new Function(lf, qf, (NonlinearFunctionPtr)0);
    qh->addConstraint(p->newConstraint(f, 0.0, 0.0));
  }
  qh->setModFlags(false, true);
  RelaxationPtr rel = (RelaxationPtr) new Relaxation(env);
  rel->setProblem(p);
  for (int32_t j = 0; j < n; ++j) rel->newVariable(root_lb[j], root_ub[j], (VariableType)var_type[j], vars[j]->getName(), vars[j]->getSrcType());
  bool is_inf = false;
  qh->relaxInitInc(rel, &is_inf);
  SolutionPoolPtr pool = (SolutionPoolPtr) new SolutionPool(env, p, 1);
  for (int32_t b = 0; b < n_boxes; ++b) {
    double *bl = lb + (size_t)b * n, *bu = ub + (size_t)b * n;
    for (int32_t j = 0; j < n; ++j) {
      p->changeBound(vars[j], bl[j], bu[j]);
      rel->changeBound(rel->getRelaxationVar(vars[j]), bl[j], bu[j]);
    }
    qh->bStats_.niters = 1;
    ModVector p_mods, r_mods;
    const bool inf = qh->presolveNode(rel, (NodePtr)0, pool, p_mods, r_mods);
    verdict[b] = inf ? 1 : 0;
    n_mods[b] = (int64_t)p_mods.size();
    for (int32_t j = 0; j < n; ++j) { bl[j] = vars[j]->getLb(); bu[j] = vars[j]->getUb(); }
    freeMods(p_mods);
    freeMods(r_mods);
  }
  delete pool;
  delete qh;
  delete rel;
  delete p;
  delete env;
  return 0;
}

/* NlPresHandler::chkRed_ alone (NlPresHandler.cpp:101-208) on the current box: 1 = infeasible */
int32_t ref_nl_chk_red(void *hv)
{
  RefProblem *h = (RefProblem *)hv;
  bool changed = false; ModQ mods; SolveStatus st = Started;
  h->nh->chkRed_(h->p, false, &changed, &mods, st);
  return st == SolvedInfeasible ? 1 : 0;
}

/* QuadraticFunction::computeBounds of quadratic constraint q on the current box */
void ref_quad_compute_bounds(void *hv, int32_t q, double *lb, double *ub)
{
  RefProblem *h = (RefProblem *)hv;
  h->quad_rows[q]->getFunction()->getQuadraticFunction()->computeBounds(lb, ub);
}

/* the two random vectors LinearHandler::dupRows_ draws after srand(seed): r1[i], r2[i] alternately (:904-907) */
void ref_draw_dup_vectors(void *hv, uint32_t seed, double *r1, double *r2)
{
  RefProblem *h = (RefProblem *)hv;
  srand(seed);
  for (size_t i = 0; i < h->vars.size(); ++i) {
    r1[i] = (double)rand() / (RAND_MAX) * 10.0;
    r2[i] = (double)rand() / (RAND_MAX) * 10.0;
  }
}

static void dumpRows(RefProblem *h, uint8_t *deleted, double *row_lb, double *row_ub)
{
  for (size_t i = 0; i < h->lin_rows.size(); ++i) {
    deleted[i] = h->p->isMarkedDel(h->lin_rows[i]) ? 1 : 0;
    row_lb[i] = h->lin_rows[i]->getLb(); row_ub[i] = h->lin_rows[i]->getUb();
  }
}

/* the reference's own LinearHandler::dupRows_ with the generator seeded: which rows it marks deleted, the row bounds
 * it leaves (the problem is modified: use a fresh handle per call) */
int32_t ref_dup_rows(void *hv, uint32_t seed, uint8_t *deleted, double *row_lb, double *row_ub)
{
  RefProblem *h = (RefProblem *)hv;
  bool changed = false;
  srand(seed);
  h->lh->dupRows(&changed);
  dumpRows(h, deleted, row_lb, row_ub);
  return changed ? 1 : 0;
}

/* the same loop driven by an externally computed candidate list (mntr_gpu_root_dup_rows / orc_root_dup_rows): pairs in
 * (i, j) order; a pair is skipped when one of its rows was deleted by an earlier pair, exactly what overwriting
 * h1[j] = 1e30 does in the reference (:934-947); treatDupRows_ is the reference's own */
int32_t ref_dup_rows_replay(void *hv, int64_t n_pairs, const int32_t *pair_i, const int32_t *pair_j, const uint8_t *pair_kind,
                            const double *h1, uint8_t *deleted, double *row_lb, double *row_ub)
{
  RefProblem *h = (RefProblem *)hv;
  bool changed = false;
  std::vector<char> gone(h->lin_rows.size(), 0);
  for (int64_t k = 0; k < n_pairs; ++k) {
    const int32_t i = pair_i[k], j = pair_j[k];
    if (gone[i] || gone[j]) continue;
    const double mult = pair_kind[k] == 1 ? 1.0 : h1[i] / h1[j];
    if (h->lh->treatDup(h->lin_rows[i], h->lin_rows[j], mult, &changed)) gone[j] = 1;
  }
  dumpRows(h, deleted, row_lb, row_ub);
  return changed ? 1 : 0;
}

/* rows that the reference's root-mode linBndTighten_ would find redundant on the current box (:974) */
void ref_redundant_rows(void *hv, uint8_t *redundant)
{
  RefProblem *h = (RefProblem *)hv;
  for (size_t i = 0; i < h->lin_rows.size(); ++i) redundant[i] = h->lh->redundantOnBox(h->lin_rows[i]) ? 1 : 0;
}

/* The reference's own LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) on the current box.  MODIFIES the problem
 * (coefficients and row bounds); what it changed is found by comparing every linear row before and after: per changed
 * row the variable, its new coefficient (0: the term was erased), which row bound moved (0 none / 1 lower / 2 upper)
 * and its new value.  Returns the number of changed rows. */
int64_t ref_coeff_imp(void *hv, int64_t cap, int32_t *out_row, int32_t *out_var, double *out_coef, int32_t *out_side,
                      double *out_bnd)
{
  RefProblem *h = (RefProblem *)hv;
  const size_t m = h->lin_rows.size();
  std::vector<std::vector<std::pair<int, double> > > before(m);
  std::vector<double> lb0(m), ub0(m);
  for (size_t i = 0; i < m; ++i) {
    LinearFunctionPtr lf = h->lin_rows[i]->getLinearFunction();
    if (lf) for (VariableGroupConstIterator it = lf->termsBegin(); it != lf->termsEnd(); ++it)
      before[i].push_back(std::make_pair((int)it->first->getIndex(), it->second));
    lb0[i] = h->lin_rows[i]->getLb(); ub0[i] = h->lin_rows[i]->getUb();
  }
  bool changed = false;
  h->lh->coeffImp(&changed);
  int64_t k = 0;
  for (size_t i = 0; i < m; ++i) {
    ConstraintPtr c = h->lin_rows[i];
    LinearFunctionPtr lf = c->getLinearFunction();
    int var = -1; double coef = 0.0;
    for (size_t t = 0; t < before[i].size(); ++t) {
      const double w = lf->getWeight(h->vars[(size_t)before[i][t].first]);
      if (w != before[i][t].second) { var = before[i][t].first; coef = w; }
    }
    int side = 0; double nb = 0.0;
    if (c->getLb() != lb0[i]) { side = 1; nb = c->getLb(); }
    if (c->getUb() != ub0[i]) { side = 2; nb = c->getUb(); }
    if (var < 0 && side == 0) continue;
    if (k < cap) { out_row[k] = (int32_t)i; out_var[k] = var; out_coef[k] = coef; out_side[k] = side; out_bnd[k] = nb; }
    ++k;
  }
  return k;
}

/* linear objective  min c.x + constant  (the cut-off row of LinearHandler::varBndsFromObj_) */
void ref_set_objective(void *hv, int32_t k, const int32_t *col, const double *val, double constant)
{
  RefProblem *h = (RefProblem *)hv;
  LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
  for (int32_t t = 0; t < k; ++t) lf->addTerm(h->vars[col[t]], val[t]);
  FunctionPtr f = (FunctionPtr) new Function(lf);
  h->p->newObjective(f, constant, Minimize);
}

/* an incumbent of objective value `value`; has_incumbent == 0 removes it */
void ref_set_incumbent(void *hv, int32_t has_incumbent, double value)
{
  RefProblem *h = (RefProblem *)hv;
  delete h->spool; h->spool = 0;
  if (has_incumbent) {
    std::vector<double> x(h->vars.size(), 0.0);
    h->spool = new SolutionPool(h->env, h->p, 10);
    h->spool->addSolution(x.empty() ? 0 : &x[0], value);
  }
}

void ref_finish(void *hv)
{
  RefProblem *h = (RefProblem *)hv;
  h->p->calculateSize();
  h->lh = new LinProbe(h->env, h->p);
  h->nh = new NlPresHandler(h->env, h->p);
}

void ref_set_box(void *hv, const double *lb, const double *ub)
{
  RefProblem *h = (RefProblem *)hv;
  for (size_t j = 0; j < h->vars.size(); ++j) h->p->changeBound(h->vars[j], lb[j], ub[j]);
}

void ref_get_box(void *hv, double *lb, double *ub)
{
  RefProblem *h = (RefProblem *)hv;
  for (size_t j = 0; j < h->vars.size(); ++j) { lb[j] = h->vars[j]->getLb(); ub[j] = h->vars[j]->getUb(); }
}

int32_t ref_lin_fixpoint(void *hv, int32_t *rounds, int64_t *nmods)
{
  RefProblem *h = (RefProblem *)hv;
  return h->lh->fixpoint(h->p, h->spool, rounds, nmods);
}

int32_t ref_lin_fixpoint_counted(void *hv, int32_t max_rounds, int32_t *rounds, int64_t *nmods,
                                 int64_t *nnz)
{
  RefProblem *h = (RefProblem *)hv;
  return h->lh->fixpointCounted(h->p, h->spool, rounds, nmods, nnz, max_rounds);
}

/* raw LinearHandler::simplePresolve (LinearHandler.cpp:1605-1653): what B&B pays per node */
int32_t ref_lin_simple_presolve(void *hv, int64_t *nmods)
{
  RefProblem *h = (RefProblem *)hv;
  ModVector mods; SolveStatus st = Started;
  h->lh->simplePresolve(h->p, (SolutionPoolPtr)h->spool, mods, st);
  *nmods = (int64_t)mods.size();
  freeMods(mods);
  return st == SolvedInfeasible ? 1 : 0;
}

/* raw NlPresHandler::simplePresolve (NlPresHandler.cpp:1022-1059) */
int32_t ref_nl_simple_presolve(void *hv, int64_t *nmods)
{
  RefProblem *h = (RefProblem *)hv;
  ModVector mods; SolveStatus st = Started;
  h->nh->simplePresolve(h->p, (SolutionPoolPtr)h->spool, mods, st);
  *nmods = (int64_t)mods.size();
  freeMods(mods);
  return st == SolvedInfeasible ? 1 : (st == SolveError ? 2 : 0);
}

/* LinearHandler then NlPresHandler, as PCBProcessor::presolveNode_ (PCBProcessor.cpp:134-175) */
int32_t ref_node_presolve(void *hv, int64_t *nmods)
{
  int64_t a = 0, b = 0;
  int32_t inf = ref_lin_simple_presolve(hv, &a);
  if (!inf) inf = (ref_nl_simple_presolve(hv, &b) == 1);
  *nmods = a + b;
  return inf;
}

/* CGraph::computeBounds of nonlinear constraint c */
int32_t ref_nl_compute_bounds(void *hv, int32_t c, double *lb, double *ub)
{
  RefProblem *h = (RefProblem *)hv; int err = 0;
  h->graphs[c]->computeBounds(lb, ub, &err);
  return err;
}

/* CGraph::varBoundMods(lb_in, ub_in) of constraint c, mods applied to the problem the way
 * NlPresHandler::varBndsFromCons_ applies them (NlPresHandler.cpp:1787-1803) */
int32_t ref_nl_var_bound_mods(void *hv, int32_t c, double lb_in, double ub_in, int32_t *n_mods)
{
  RefProblem *h = (RefProblem *)hv;
  VarBoundModVector mods; SolveStatus st = Started;
  h->graphs[c]->varBoundMods(lb_in, ub_in, mods, &st);
  *n_mods = (int32_t)mods.size();
  for (VarBoundModVector::iterator it = mods.begin(); it != mods.end(); ++it) {
    if (st != SolvedInfeasible && st != SolveError) (*it)->applyToProblem(h->p);
    delete *it;
  }
  return st == SolvedInfeasible ? 1 : (st == SolveError ? 2 : 0);
}

/* opcodes of the dependent nodes in the order CGraph::finalize produced (dq_) */
int32_t ref_nl_dq_ops(void *hv, int32_t c, int32_t cap, int32_t *ops)
{
  RefProblem *h = (RefProblem *)hv;
  CNodeQ dq = h->graphs[c]->dNodes();
  int32_t k = 0;
  for (CNodeQ::iterator it = dq.begin(); it != dq.end(); ++it, ++k) if (k < cap) ops[k] = (int32_t)(*it)->getOp();
  return k;
}

void ref_row_activity(void *hv, int32_t row, double out[4])
{
  RefProblem *h = (RefProblem *)hv;
  h->lh->rowActivity(h->lin_rows[row], out);
}

/* CPU baseline legs: repeat a call on a list of boxes, steady_clock around the
 * tighten calls only (box reset excluded).  mode 0 = status-honouring fixpoint
 * (counted), 1 = raw LinearHandler::simplePresolve, 2 = raw node presolve (lin + nl).
 * Returns seconds; accumulates nnz / verdict counts. */
double ref_time_boxes(void *hv, int32_t mode, int32_t n_boxes, const double *lbs, const double *ubs,
                      int64_t *nnz_total, int64_t *n_infeasible)
{
  RefProblem *h = (RefProblem *)hv;
  size_t n = h->vars.size();
  double secs = 0; *nnz_total = 0; *n_infeasible = 0;
  for (int32_t b = 0; b < n_boxes; ++b) {
    ref_set_box(hv, lbs + (size_t)b * n, ubs + (size_t)b * n);
    int32_t rounds = 0, inf = 0; int64_t nmods = 0, nnz = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (mode == 0) inf = h->lh->fixpointCounted(h->p, h->spool, &rounds, &nmods, &nnz, 0);
    else if (mode == 1) inf = ref_lin_simple_presolve(hv, &nmods);
    else inf = ref_node_presolve(hv, &nmods);
    auto t1 = std::chrono::steady_clock::now();
    secs += std::chrono::duration<double>(t1 - t0).count();
    *nnz_total += nnz; *n_infeasible += inf;
  }
  return secs;
}

/* all nonlinear constraints of a flat tape description at once (the Python loop over ref_add_nl costs 60 us per
 * constraint, a minute for the 1M constraints of config C5) */
int32_t ref_add_nl_batch(void *hv, int32_t n_cons, const int32_t *tape_ptr, const uint8_t *op, const int32_t *arg0,
                         const int32_t *arg1, const double *cnst, const int32_t *child, const int32_t *lin_ptr,
                         const int32_t *lin_col, const double *lin_val, const double *c_lb, const double *c_ub)
{
  for (int32_t c = 0; c < n_cons; ++c) {
    const int32_t b = tape_ptr[c], lb = lin_ptr[c];
    ref_add_nl(hv, tape_ptr[c + 1] - b, op + b, arg0 + b, arg1 + b, cnst + b, child, lin_ptr[c + 1] - lb,
               lin_col + lb, lin_val + lb, c_lb[c], c_ub[c]);
  }
  return n_cons;
}

/* ref_time_boxes for boxes given as branching deltas on a root box (the form of mntr_gpu_tighten_nodes): box b is
 * the root box with deltas dptr[b] .. dptr[b+1) applied in order.  Only the tighten calls are timed. */
double ref_time_deltas(void *hv, int32_t mode, int32_t n_boxes, const double *root_lb, const double *root_ub,
                       const int64_t *dptr, const int32_t *dvar, const uint8_t *dup, const double *dval,
                       int64_t *nnz_total, int64_t *n_infeasible)
{
  RefProblem *h = (RefProblem *)hv;
  double secs = 0; *nnz_total = 0; *n_infeasible = 0;
  for (int32_t b = 0; b < n_boxes; ++b) {
    ref_set_box(hv, root_lb, root_ub);
    for (int64_t q = dptr[b]; q < dptr[b + 1]; ++q) {
      VariablePtr v = h->vars[dvar[q]];
      if (dup[q]) h->p->changeBound(v, Upper, dval[q]); else h->p->changeBound(v, Lower, dval[q]);
    }
    int32_t rounds = 0, inf = 0; int64_t nmods = 0, nnz = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (mode == 0) inf = h->lh->fixpointCounted(h->p, h->spool, &rounds, &nmods, &nnz, 0);
    else if (mode == 1) inf = ref_lin_simple_presolve(hv, &nmods);
    else inf = ref_node_presolve(hv, &nmods);
    auto t1 = std::chrono::steady_clock::now();
    secs += std::chrono::duration<double>(t1 - t0).count();
    *nnz_total += nnz; *n_infeasible += inf;
  }
  return secs;
}

// ---- Reader::readMps (Reader.cpp:42-473): read an MPS file with the reference's own reader and dump the problem
//      flat -- the fixture generator of minotaur_b200/mps_reader.py.  Two-call protocol: sizes first (arrays null),
//      then the arrays.  Returns the reader's error code (0 = ok), or -1 if the file gave no problem.
int32_t ref_read_mps(const char *path, int32_t *m, int32_t *n, int32_t *nnz, int32_t *row_ptr, int32_t *col,
                     double *val, double *row_lb, double *row_ub, uint8_t *var_type, double *lb, double *ub,
                     int32_t *obj_k, int32_t *obj_col, double *obj_val, double *obj_const)
{
  EnvPtr env = new Environment();
  int err = 0;
  env->getLogger()->setMaxLevel(LogNone);
  env->startTimer(err);
  Reader rd(env);
  ProblemPtr p = rd.readMps(path, err);
  if (!p) { delete env; return err ? err : -1; }
  *m = (int32_t)p->getNumCons(); *n = (int32_t)p->getNumVars();
  int32_t cnt = 0, k = 0;
  if (row_ptr) row_ptr[0] = 0;
  for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it, ++k) {
    LinearFunctionPtr lf = (*it)->getLinearFunction();
    if (lf) {
      for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
        if (col) { col[cnt] = (int32_t)t->first->getIndex(); val[cnt] = t->second; }
        ++cnt;
      }
    }
    if (row_ptr) { row_ptr[k + 1] = cnt; row_lb[k] = (*it)->getLb(); row_ub[k] = (*it)->getUb(); }
  }
  *nnz = cnt;
  if (var_type) {
    int j = 0;
    for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it, ++j) {
      var_type[j] = (uint8_t)(*it)->getType(); lb[j] = (*it)->getLb(); ub[j] = (*it)->getUb();
    }
  }
  *obj_k = 0; *obj_const = 0.0;
  if (p->getObjective()) {
    *obj_const = p->getObjective()->getConstant();
    LinearFunctionPtr lf = p->getObjective()->getLinearFunction();
    if (lf) {
      for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
        if (obj_col) { obj_col[*obj_k] = (int32_t)t->first->getIndex(); obj_val[*obj_k] = t->second; }
        ++*obj_k;
      }
    }
  }
  delete p;
  delete env;
  return err;
}

void ref_destroy(void *hv)
{
  RefProblem *h = (RefProblem *)hv;
  delete h->lh; delete h->nh; delete h->spool;
  delete h->p;
  delete h->env;
  delete h;
}

}  // extern "C"
