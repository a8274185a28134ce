// handler_test.cpp -- TEST: GpuBoundHandler against the reference's own LinearHandler / NlPresHandler
// through Minotaur's Handler::presolveNode interface, on the same relaxation objects.
//
// Built by oracle/Makefile (target handler_test) against the reference objects in oracle/_ref; run on a GPU
// box by tests/test_gpu_handler.py.  Exit code 0 = every comparison held.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "MinotaurConfig.h"
#include "CGraph.h"
#include "Constraint.h"
#include "Environment.h"
#include "Function.h"
#include "GpuBoundHandler.h"
#include "LinearFunction.h"
#include "LinearHandler.h"
#include "NlPresHandler.h"
#include "Objective.h"
#include "Presolver.h"
#include "Problem.h"
#include "Relaxation.h"
#include "SolutionPool.h"
#include "VarBoundMod.h"
#include "Variable.h"

using namespace Minotaur;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double urand() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static int irand(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }

// random MILP with a planted point (the C2 generator's recipe) + a few CGraph constraints
static ProblemPtr makeProblem(EnvPtr env, int n, int m, int k, int n_nl, std::vector<double> &xstar)
{
  ProblemPtr p = (ProblemPtr) new Problem(env);
  xstar.resize(n);
  std::vector<VariablePtr> v(n);
  for (int j = 0; j < n; ++j) {
    const bool isint = urand() < 0.5;
    const double ub = irand(1, 10);
    v[j] = p->newVariable(0.0, ub, isint ? (ub == 1.0 ? Binary : Integer) : Continuous);
    xstar[j] = isint ? std::floor(urand() * (ub + 1)) : std::floor(urand() * ub * 4) / 4;
    if (xstar[j] > ub) xstar[j] = ub;
  }
  for (int i = 0; i < m; ++i) {
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    double act = 0;
    for (int t = 0; t < k; ++t) {
      const int j = irand(0, n - 1);
      if (lf->getWeight(v[j]) != 0.0) continue;
      const double a = irand(1, 9) * (urand() < 0.3 ? -1.0 : 1.0);
      lf->addTerm(v[j], a);
      act += a * xstar[j];
    }
    const bool eq = urand() < 0.3;
    p->newConstraint((FunctionPtr) new Function(lf), eq ? act : -INFINITY, eq ? act : act + irand(0, 3));
  }
  for (int c = 0; c < n_nl; ++c) {
    int i = irand(0, n - 1), j = irand(0, n - 1), l = irand(0, n - 1);
    if (i == j) j = (j + 1) % n;
    CGraph *cg = new CGraph();
    CNode *ni = cg->newNode(v[i]), *nj = cg->newNode(v[j]);
    CNode *out = cg->newNode(OpMult, ni, nj);
    cg->setOut(out);
    cg->finalize();
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    const double a = irand(1, 5);
    lf->addTerm(v[l], a);
    const double val = xstar[i] * xstar[j] + a * xstar[l];
    p->newConstraint((FunctionPtr) new Function(lf, (NonlinearFunctionPtr)cg), -INFINITY, val + irand(0, 3));
  }
  // Relaxation's cloning constructor dereferences the objective (Relaxation.cpp:139-140)
  // a linear objective over ~30 variables: with an incumbent it becomes the cut-off row of varBndsFromObj_
  LinearFunctionPtr of = (LinearFunctionPtr) new LinearFunction();
  of->addTerm(v[0], 1.0);
  for (int t = 0; t < 30; ++t) {
    const int j = irand(1, n - 1);
    if (of->getWeight(v[j]) == 0.0) of->addTerm(v[j], (double)irand(1, 5));
  }
  p->newObjective((FunctionPtr) new Function(of), 2.5, Minimize);
  p->calculateSize();
  return p;
}

static void branch(ProblemPtr p, int depth)
{
  for (int d = 0; d < depth; ++d) {
    VariablePtr v = p->getVariable(irand(0, p->getNumVars() - 1));
    if ((v->getType() != Integer && v->getType() != Binary) || v->getUb() - v->getLb() < 1) continue;
    const double x = v->getLb() + (v->getUb() - v->getLb()) * urand();
    if (urand() < 0.5) p->changeBound(v, Upper, std::floor(x)); else p->changeBound(v, Lower, std::ceil(x + 1e-9));
  }
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
                              fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

// tests/golden/tls4_flat.txt (written by tests/golden/make_tls4_flat.py): test_instances/tls4.nl as Minotaur objects
static ProblemPtr readFlat(EnvPtr env, const char *path)
{
  std::ifstream in(path);
  if (!in) return 0;
  auto num = [&]() { std::string t; in >> t; return t == "inf" ? INFINITY : (t == "-inf" ? -INFINITY : atof(t.c_str())); };
  int n, m, nc;
  in >> n >> m >> nc;
  ProblemPtr p = (ProblemPtr) new Problem(env);
  std::vector<VariablePtr> v(n);
  for (int j = 0; j < n; ++j) { int ty; in >> ty; const double lb = num(), ub = num(); v[j] = p->newVariable(lb, ub, (VariableType)ty); }
  for (int i = 0; i < m; ++i) {
    const double lb = num(), ub = num(); int k; in >> k;
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    for (int t = 0; t < k; ++t) { int c; in >> c; lf->addTerm(v[c], num()); }
    p->newConstraint((FunctionPtr) new Function(lf), lb, ub);
  }
  for (int c = 0; c < nc; ++c) {
    const double clb = num(), cub = num(); int nn, klin, nchild; in >> nn >> klin >> nchild;
    std::vector<int> op(nn), a0(nn), a1(nn); std::vector<double> cn(nn);
    for (int i = 0; i < nn; ++i) { in >> op[i] >> a0[i] >> a1[i]; cn[i] = num(); }
    std::vector<std::pair<int, double> > lin;
    if (klin == 0) { std::string dash; in >> dash; }
    for (int t = 0; t < klin; ++t) { int col; in >> col; lin.push_back(std::make_pair(col, num())); }
    std::vector<int> child(nchild);
    if (nchild == 0) { std::string dash; in >> dash; }
    for (int t = 0; t < nchild; ++t) in >> child[t];
    CGraph *cg = new CGraph();
    std::vector<CNode *> nodes(nn, (CNode *)0);
    for (int i = 0; i < nn; ++i) {
      const OpCode o = (OpCode)op[i];
      if (o == OpVar) nodes[i] = cg->newNode(v[a0[i]]);
      else if (o == OpNum) nodes[i] = cg->newNode(cn[i]);
      else if (o == OpInt) nodes[i] = cg->newNode((int)cn[i]);
      else if (o == OpSumList) {
        std::vector<CNode *> ch;
        for (int q = a0[i]; q < a1[i]; ++q) ch.push_back(nodes[child[q]]);
        nodes[i] = cg->newNode(OpSumList, &ch[0], (UInt)ch.size());
      } else nodes[i] = cg->newNode(o, nodes[a0[i]], a1[i] >= 0 ? nodes[a1[i]] : (CNode *)0);
    }
    cg->setOut(nodes[nn - 1]);
    cg->finalize();
    LinearFunctionPtr lf = 0;
    if (klin > 0) { lf = (LinearFunctionPtr) new LinearFunction(); for (int t = 0; t < klin; ++t) lf->addTerm(v[lin[t].first], lin[t].second); }
    p->newConstraint((FunctionPtr) new Function(lf, (NonlinearFunctionPtr)cg), clb, cub);
  }
  int k; in >> k; const double oc = num();
  LinearFunctionPtr of = (LinearFunctionPtr) new LinearFunction();
  for (int t = 0; t < k; ++t) { int c; in >> c; of->addTerm(v[c], num()); }
  p->newObjective((FunctionPtr) new Function(of), oc, Minimize);
  p->calculateSize();
  return p;
}

// LinearHandler::coeffImp_ is protected: a probe
struct CoeffProbe : public LinearHandler {
  CoeffProbe(EnvPtr env, ProblemPtr p) : LinearHandler(env, p) {}
  void run(bool *changed) { coeffImp_(changed); }
};

// variable upper / lower bound rows with loose big-M coefficients plus mixed one-sided rows (what coeffImp_ improves)
static ProblemPtr makeBigM(EnvPtr env, int nBin, int nCont, int nRows)
{
  ProblemPtr p = (ProblemPtr) new Problem(env);
  std::vector<VariablePtr> v;
  for (int j = 0; j < nBin; ++j) v.push_back(p->newVariable(0.0, 1.0, Binary));
  for (int j = 0; j < nCont; ++j) {
    const double l = (double)((int)(urand() * 5) - 3);
    v.push_back(p->newVariable(l, l + 1.0 + (int)(urand() * 11), Continuous));
  }
  int made = 0;
  for (int j = nBin; j < nBin + nCont && made < nRows; ++j) {
    if (urand() < 0.6) {
      LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
      lf->addTerm(v[(size_t)(urand() * nBin)], -(v[(size_t)j]->getUb() + (int)(urand() * 6)));
      lf->addTerm(v[(size_t)j], 1.0);
      p->newConstraint((FunctionPtr) new Function(lf), -INFINITY, 0.0);
      ++made;
    }
  }
  while (made < nRows) {
    LinearFunctionPtr lf = (LinearFunctionPtr) new LinearFunction();
    const int k = 2 + (int)(urand() * 7);
    double lo = 0.0, hi = 0.0;
    for (int t = 0; t < k; ++t) {
      const size_t j = (size_t)(urand() * (nBin + nCont));
      if (lf->hasVar(v[j])) continue;
      const double a = (j < (size_t)nBin ? 1.0 + (int)(urand() * 29) : 1.0 + (int)(urand() * 5)) * (urand() < 0.5 ? 1.0 : -1.0);
      lf->addTerm(v[j], a);
      lo += a > 0 ? a * v[j]->getLb() : a * v[j]->getUb();
      hi += a > 0 ? a * v[j]->getUb() : a * v[j]->getLb();
    }
    if (lf->getNumTerms() < 2) { delete lf; continue; }
    if (urand() < 0.5) p->newConstraint((FunctionPtr) new Function(lf), -INFINITY, std::floor(lo + (hi - lo) * (0.3 + 0.8 * urand())));
    else p->newConstraint((FunctionPtr) new Function(lf), std::ceil(lo + (hi - lo) * (-0.1 + 0.8 * urand())), INFINITY);
    ++made;
  }
  p->calculateSize();
  return p;
}

// Root presolve through Presolver::solve (Presolver.cpp:91-182): the reference pair LinearHandler + NlPresHandler on one
// copy of the problem, LinearHandler + GpuBoundHandler (reference order, round to nearest) on another.  Returns the
// number of bounds that differ; prints what each run did.
static int rootPresolve(EnvPtr env, ProblemPtr pA, ProblemPtr pB, const char *what, int *tightA, int *tightB)
{
  std::vector<double> l0, u0;
  for (VariableConstIterator it = pA->varsBegin(); it != pA->varsEnd(); ++it) { l0.push_back((*it)->getLb()); u0.push_back((*it)->getUb()); }
  HandlerVector hA, hB;
  LinearHandler *lhA = new LinearHandler(env, pA); NlPresHandler *nhA = new NlPresHandler(env, pA);
  lhA->setModFlags(false, true); nhA->setModFlags(false, true);
  hA.push_back(lhA); hA.push_back(nhA);
  LinearHandler *lhB = new LinearHandler(env, pB); GpuBoundHandler *ghB = new GpuBoundHandler(env, pB, 0);
  lhB->setModFlags(false, true);
  ghB->setMode(GpuBoundHandler::ReferenceOrder); ghB->setRoundNearest(true);
  hB.push_back(lhB); hB.push_back(ghB);
  Presolver prA(pA, env, hA), prB(pB, env, hB);
  const SolveStatus sA = prA.solve(), sB = prB.solve();
  int diff = 0; *tightA = 0; *tightB = 0;
  const UInt n = pA->getNumVars();
  if (pB->getNumVars() != n) { printf("root presolve %s: variable counts differ (%u vs %u)\n", what, n, pB->getNumVars()); return -1; }
  VariableConstIterator ia = pA->varsBegin(), ib = pB->varsBegin();
  for (UInt j = 0; j < n; ++j, ++ia, ++ib) {
    if ((*ia)->getLb() != (*ib)->getLb() || (*ia)->getUb() != (*ib)->getUb()) ++diff;
  }
  printf("root presolve %s: reference pair status %d, %u vars %u cons left; with GpuBoundHandler status %d, %u vars %u cons left; "
         "%d of %u variables end with different bounds; GPU handler: %d calls, %d uploads, %d mods\n", what, (int)sA,
         pA->getNumVars(), pA->getNumCons(), (int)sB, pB->getNumVars(), pB->getNumCons(), diff, n, ghB->getStats()->calls,
         ghB->getStats()->uploads, ghB->getStats()->nMods);
  (void)l0; (void)u0;
  delete lhA; delete nhA; delete lhB; delete ghB;
  if (sA != sB) return -2;
  return (pA->getNumCons() == pB->getNumCons()) ? diff : -1;
}

static void onAbort(int)
{
  void *frames[48];
  const int nf = backtrace(frames, 48);
  backtrace_symbols_fd(frames, nf, 2);
  _exit(134);
}

int main(int argc, char **argv)
{
  setvbuf(stdout, 0, _IOLBF, 0);
  signal(SIGABRT, onAbort);
  EnvPtr env = (EnvPtr) new Environment();
  int err = 0;
  env->startTimer(err);
  env->setLogLevel(LogNone);
  int n_cmp = 0, n_inf = 0, n_mods = 0, n_cut = 0;
  for (int trial = 0; trial < 6; ++trial) {
    std::vector<double> xstar;
    const int n = 300 + 50 * trial, m = 350, n_nl = (trial % 2) ? 40 : 0;
    ProblemPtr p = makeProblem(env, n, m, 6, n_nl, xstar);
    for (int box = 0; box < 6; ++box) {
      RelaxationPtr relA = (RelaxationPtr) new Relaxation(p, env);
      RelaxationPtr relB = (RelaxationPtr) new Relaxation(p, env);
      relA->calculateSize(); relB->calculateSize();
      const uint64_t keep = rng_state;
      branch(relA, 2 * box);
      rng_state = keep;
      branch(relB, 2 * box);
      std::vector<double> lb0(n), ub0(n);
      for (int j = 0; j < n; ++j) { lb0[j] = relB->getVariable(j)->getLb(); ub0[j] = relB->getVariable(j)->getUb(); }

      // every other box has an incumbent: the objective becomes the cut-off row of LinearHandler::varBndsFromObj_
      // and NlPresHandler::fixObjBins_ fixes binaries of the objective
      SolutionPool *spool = 0;
      if (box & 1) {
        std::vector<double> x(n, 0.0);
        spool = new SolutionPool(env, p, 10);
        spool->addSolution(&x[0], 2.5 + irand(15, 45));
        ++n_cut;
      }

      // reference: LinearHandler then NlPresHandler, one PCBProcessor::presolveNode_ pass
      LinearHandler lh(env, p);
      NlPresHandler nh(env, p);
      // (Handler::modProb_ / modRel_ are uninitialised until set -- Handler.h:380-383 -- and a LinearHandler that finds
      //  garbage `true` in modProb_ mirrors the relaxation's bounds into p: the solver mains always set them, Bnb.cpp:111,121)
      lh.setModFlags(false, true);
      nh.setModFlags(false, true);
      ModVector pm, rmA;
      bool infA = lh.presolveNode(relA, (NodePtr)0, (SolutionPoolPtr)spool, pm, rmA);
      if (!infA && n_nl) infA = nh.presolveNode(relA, (NodePtr)0, (SolutionPoolPtr)spool, pm, rmA);

      // GPU handler in reference-order, round-to-nearest mode: must reproduce it bit for bit
      GpuBoundHandler gh(env, p, 0);
      gh.setMode(GpuBoundHandler::ReferenceOrder);
      gh.setRoundNearest(true);
      ModVector rmB;
      const bool infB = gh.presolveNode(relB, (NodePtr)0, (SolutionPoolPtr)spool, pm, rmB);
      ++n_cmp;
      if (infB && !infA) {
        // allowed only when the GPU proved an activity-infeasible row, which the reference's node mode drops
        // (LinearHandler.cpp:1631); such a box must really be infeasible: the reference's own sweeps agree
        // once their status is honoured -- checked by the parity tests; here we only count it
        ++n_inf;
      } else {
        CHECK(infA == infB, "trial %d box %d: verdict ref %d gpu %d", trial, box, (int)infA, (int)infB);
        if (!infA) {
          for (int j = 0; j < n; ++j) {
            CHECK(relA->getVariable(j)->getLb() == relB->getVariable(j)->getLb() &&
                  relA->getVariable(j)->getUb() == relB->getVariable(j)->getUb(),
                  "trial %d box %d var %d: ref [%.17g,%.17g] gpu [%.17g,%.17g]", trial, box, j,
                  relA->getVariable(j)->getLb(), relA->getVariable(j)->getUb(), relB->getVariable(j)->getLb(),
                  relB->getVariable(j)->getUb());
          }
        } else ++n_inf;
      }
      n_mods += (int)rmB.size();
      if (getenv("HANDLER_TEST_VERBOSE"))
        printf("trial %d box %d: ref inf %d (%d mods), gpu inf %d (%d mods), cutoff %d\n", trial, box, (int)infA, (int)rmA.size(),
               (int)infB, (int)rmB.size(), spool ? 1 : 0);
      // undo in reverse order restores the incoming box (Node.cpp:318-340)
      for (ModVector::reverse_iterator it = rmB.rbegin(); it != rmB.rend(); ++it) (*it)->undoToProblem(relB);
      for (int j = 0; j < n; ++j)
        CHECK(relB->getVariable(j)->getLb() == lb0[j] && relB->getVariable(j)->getUb() == ub0[j],
              "trial %d box %d var %d: undo did not restore the box", trial, box, j);

      // default mode (Jacobi fixpoint, directed rounding): valid, never cuts off the planted point on the root box
      if (box == 0) {
        GpuBoundHandler gf(env, p, 0);
        ModVector rmC;
        const bool infC = gf.presolveNode(relB, (NodePtr)0, (SolutionPoolPtr)0, pm, rmC);
        CHECK(!infC, "trial %d: root box declared infeasible by the fast mode", trial);
        for (int j = 0; j < n && !infC; ++j)
          CHECK(relB->getVariable(j)->getLb() <= xstar[j] + 1e-6 && relB->getVariable(j)->getUb() >= xstar[j] - 1e-6,
                "trial %d var %d: planted point cut off [%g,%g] x*=%g", trial, j, relB->getVariable(j)->getLb(),
                relB->getVariable(j)->getUb(), xstar[j]);
        for (ModVector::iterator it = rmC.begin(); it != rmC.end(); ++it) delete *it;
      }
      for (ModVector::iterator it = rmA.begin(); it != rmA.end(); ++it) delete *it;
      for (ModVector::iterator it = rmB.begin(); it != rmB.end(); ++it) delete *it;
      delete relA; delete relB;
      delete spool;
    }
    delete p;
  }
  // ---- batch form: all strong-branching candidates of a node in one call (StrongBrancher.cpp:505-585) ----
  int n_cand = 0, n_cand_inf = 0;
  {
    std::vector<double> xstar;
    const int n = 400;
    ProblemPtr p = makeProblem(env, n, 350, 6, 30, xstar);
    RelaxationPtr rel = (RelaxationPtr) new Relaxation(p, env);
    rel->calculateSize();
    branch(rel, 4);
    GpuBoundHandler gh(env, p, 0);
    gh.setMode(GpuBoundHandler::ReferenceOrder);
    gh.setRoundNearest(true);
    // candidates: down and up branch of 40 integer variables
    std::vector<std::vector<GpuBoundHandler::BoundChange> > deltas;
    for (int j = 0; j < n && (int)deltas.size() < 80; ++j) {
      VariablePtr v = rel->getVariable(j);
      if ((v->getType() != Integer && v->getType() != Binary) || v->getUb() - v->getLb() < 1) continue;
      const double x = std::floor(v->getLb() + 0.5 * (v->getUb() - v->getLb()));
      GpuBoundHandler::BoundChange dn = { (UInt)j, Upper, x }, up = { (UInt)j, Lower, x + 1 };
      deltas.push_back(std::vector<GpuBoundHandler::BoundChange>(1, dn));
      deltas.push_back(std::vector<GpuBoundHandler::BoundChange>(1, up));
    }
    std::vector<GpuBoundHandler::BoxOutcome> outc;
    gh.tightenCandidates(rel, (SolutionPoolPtr)0, deltas, outc);
    ModVector pm;
    for (size_t b = 0; b < deltas.size(); ++b) {
      // the same box the way StrongBrancher does it: apply the branching mod, presolveNode, undo
      VariablePtr bv = rel->getVariable(deltas[b][0].var);
      VarBoundMod br(bv, deltas[b][0].lu, deltas[b][0].val);
      br.applyToProblem(rel);
      std::vector<double> l0(n), u0(n);
      for (int j = 0; j < n; ++j) { l0[j] = rel->getVariable(j)->getLb(); u0[j] = rel->getVariable(j)->getUb(); }
      ModVector rm;
      const bool inf = gh.presolveNode(rel, (NodePtr)0, (SolutionPoolPtr)0, pm, rm);
      ++n_cand;
      CHECK(inf == outc[b].infeasible, "candidate %d: verdict single %d batch %d", (int)b, (int)inf, (int)outc[b].infeasible);
      if (!inf && !outc[b].infeasible) {
        std::vector<double> l(l0), u(u0);
        for (size_t k = 0; k < outc[b].changes.size(); ++k)
          (outc[b].changes[k].lu == Upper ? u : l)[outc[b].changes[k].var] = outc[b].changes[k].val;
        for (int j = 0; j < n; ++j)
          CHECK(rel->getVariable(j)->getLb() == l[j] && rel->getVariable(j)->getUb() == u[j],
                "candidate %d var %d: single [%.17g,%.17g] batch [%.17g,%.17g]", (int)b, j,
                rel->getVariable(j)->getLb(), rel->getVariable(j)->getUb(), l[j], u[j]);
      } else ++n_cand_inf;
      for (ModVector::reverse_iterator it = rm.rbegin(); it != rm.rend(); ++it) { (*it)->undoToProblem(rel); delete *it; }
      br.undoToProblem(rel);
    }
    // ---- the same loop the way StrongBrancher runs it with strong_brancher_prefetch.patch: one prefetch, then every
    //      getStrongerMods -> presolveNode is answered from the prefetched outcomes (no device call per candidate) ----
    {
      const uint64_t keepRng = rng_state;   // the sections below keep the random problems they were written against
      GpuBoundHandler g2(env, p, 0);
      g2.setMode(GpuBoundHandler::ReferenceOrder);
      g2.setRoundNearest(true);
      g2.prefetchCandidates(rel, (SolutionPoolPtr)0, deltas);
      int hits = 0;
      for (size_t b = 0; b < deltas.size(); ++b) {
        VariablePtr bv = rel->getVariable(deltas[b][0].var);
        VarBoundMod br(bv, deltas[b][0].lu, deltas[b][0].val);
        br.applyToProblem(rel);
        std::vector<double> l0(n), u0(n);
        for (int j = 0; j < n; ++j) { l0[j] = rel->getVariable(j)->getLb(); u0[j] = rel->getVariable(j)->getUb(); }
        ModVector rm;
        const bool inf = g2.presolveNode(rel, (NodePtr)0, (SolutionPoolPtr)0, pm, rm);
        CHECK(inf == outc[b].infeasible, "prefetched candidate %d: verdict %d, batch %d", (int)b, (int)inf, (int)outc[b].infeasible);
        if (!inf) {
          std::vector<double> l(l0), u(u0);
          for (size_t k = 0; k < outc[b].changes.size(); ++k)
            (outc[b].changes[k].lu == Upper ? u : l)[outc[b].changes[k].var] = outc[b].changes[k].val;
          for (int j = 0; j < n; ++j)
            CHECK(rel->getVariable(j)->getLb() == l[j] && rel->getVariable(j)->getUb() == u[j],
                  "prefetched candidate %d var %d differs from the batch outcome", (int)b, j);
        }
        for (ModVector::reverse_iterator it = rm.rbegin(); it != rm.rend(); ++it) { (*it)->undoToProblem(rel); delete *it; }
        br.undoToProblem(rel);
      }
      hits = g2.getStats()->cacheHits;
      CHECK(hits == (int)deltas.size(), "prefetch: %d of %d candidate calls were answered from the prefetched outcomes", hits, (int)deltas.size());
      // a box that is not one of the candidates drops the outcomes and takes the device path again
      branch(rel, 3);
      ModVector rm;
      (void)g2.presolveNode(rel, (NodePtr)0, (SolutionPoolPtr)0, pm, rm);
      CHECK(g2.getStats()->cacheHits == hits, "prefetch: a foreign box was answered from the cache");
      for (ModVector::iterator it = rm.begin(); it != rm.end(); ++it) delete *it;
      printf("handler_test: prefetch of %d candidate boxes, %d presolveNode calls answered without a device call\n", (int)deltas.size(), hits);
      rng_state = keepRng;
    }
    delete rel;
    delete p;
  }
  // ---- the device copy follows the problem: a row bound changed between two calls (ConBoundMod, a cut tightened) ----
  int n_stale = 0;
  {
    std::vector<double> xstar;
    const uint64_t keep = rng_state;
    ProblemPtr pA = makeProblem(env, 300, 320, 6, 20, xstar);
    rng_state = keep;
    ProblemPtr pB = makeProblem(env, 300, 320, 6, 20, xstar);
    RelaxationPtr relA = (RelaxationPtr) new Relaxation(pA, env), relB = (RelaxationPtr) new Relaxation(pB, env);
    relA->calculateSize(); relB->calculateSize();
    GpuBoundHandler gh(env, pB, 0);
    gh.setMode(GpuBoundHandler::ReferenceOrder); gh.setRoundNearest(true);
    ModVector pm;
    for (int step = 0; step < 3; ++step) {
      if (step > 0) {
        // tighten the upper bound of a few linear rows by one (the planted point may be cut off: fine, both sides see
        // the same rows) -- same change on both copies
        for (int k = 0; k < 12; ++k) {
          const int i = irand(0, 319);
          ConstraintPtr ca = relA->getConstraint(i), cb = relB->getConstraint(i);
          if (ca->getUb() < 1e20) { relA->changeBound(ca, Upper, ca->getUb() - 1.0); relB->changeBound(cb, Upper, cb->getUb() - 1.0); }
        }
      }
      LinearHandler lh(env, pA); NlPresHandler nh(env, pA);
      lh.setModFlags(false, true); nh.setModFlags(false, true);
      ModVector rmA, rmB;
      bool infA = lh.presolveNode(relA, (NodePtr)0, (SolutionPoolPtr)0, pm, rmA);
      if (!infA) infA = nh.presolveNode(relA, (NodePtr)0, (SolutionPoolPtr)0, pm, rmA);
      const bool infB = gh.presolveNode(relB, (NodePtr)0, (SolutionPoolPtr)0, pm, rmB);
      ++n_stale;
      if (!(infB && !infA)) {
        CHECK(infA == infB, "stale step %d: verdict ref %d gpu %d", step, (int)infA, (int)infB);
        for (int j = 0; j < 300 && !infA; ++j)
          CHECK(relA->getVariable(j)->getLb() == relB->getVariable(j)->getLb() && relA->getVariable(j)->getUb() == relB->getVariable(j)->getUb(),
                "stale step %d var %d: ref [%.17g,%.17g] gpu [%.17g,%.17g]", step, j, relA->getVariable(j)->getLb(),
                relA->getVariable(j)->getUb(), relB->getVariable(j)->getLb(), relB->getVariable(j)->getUb());
      }
      for (ModVector::reverse_iterator it = rmA.rbegin(); it != rmA.rend(); ++it) { (*it)->undoToProblem(relA); delete *it; }
      for (ModVector::reverse_iterator it = rmB.rbegin(); it != rmB.rend(); ++it) { (*it)->undoToProblem(relB); delete *it; }
    }
    CHECK(gh.getStats()->uploads == 1 && gh.getStats()->rowBoundUpdates == 2,
          "row-bound changes: %d uploads, %d row-bound refreshes (expected 1 and 2)", gh.getStats()->uploads, gh.getStats()->rowBoundUpdates);
    delete relA; delete relB; delete pA; delete pB;
  }
  // ---- setModFlags(true, true): the bounds found on the relaxation are mirrored into the original problem as p_mods
  //      (LinearHandler::copyBndsFromRel_, LinearHandler.cpp:108-132) ----
  int n_pmods = 0;
  {
    std::vector<double> xstar;
    const uint64_t keep = rng_state;
    ProblemPtr pA = makeProblem(env, 260, 300, 6, 0, xstar);
    rng_state = keep;
    ProblemPtr pB = makeProblem(env, 260, 300, 6, 0, xstar);
    RelaxationPtr relA = (RelaxationPtr) new Relaxation(pA, env), relB = (RelaxationPtr) new Relaxation(pB, env);
    relA->calculateSize(); relB->calculateSize();
    const uint64_t keep2 = rng_state;
    branch(relA, 6); rng_state = keep2; branch(relB, 6);
    LinearHandler lh(env, pA);
    lh.setModFlags(true, true);
    GpuBoundHandler gh(env, pB, 0);
    gh.setMode(GpuBoundHandler::ReferenceOrder); gh.setRoundNearest(true);
    gh.setModFlags(true, true);
    ModVector pmA, pmB, rmA, rmB;
    const bool infA = lh.presolveNode(relA, (NodePtr)0, (SolutionPoolPtr)0, pmA, rmA);
    const bool infB = gh.presolveNode(relB, (NodePtr)0, (SolutionPoolPtr)0, pmB, rmB);
    CHECK(infA == infB || (infB && !infA), "modProb: verdicts differ");
    if (!infA && !infB) {
      CHECK(pmA.size() == pmB.size(), "modProb: %d p_mods from LinearHandler, %d from GpuBoundHandler", (int)pmA.size(), (int)pmB.size());
      for (int j = 0; j < 260; ++j)
        CHECK(pA->getVariable(j)->getLb() == pB->getVariable(j)->getLb() && pA->getVariable(j)->getUb() == pB->getVariable(j)->getUb(),
              "modProb var %d: original problem bounds differ", j);
      n_pmods = (int)pmB.size();
    }
    for (ModVector::iterator it = pmA.begin(); it != pmA.end(); ++it) delete *it;
    for (ModVector::iterator it = pmB.begin(); it != pmB.end(); ++it) delete *it;
    for (ModVector::iterator it = rmA.begin(); it != rmA.end(); ++it) delete *it;
    for (ModVector::iterator it = rmB.begin(); it != rmB.end(); ++it) delete *it;
    delete relA; delete relB; delete pA; delete pB;
  }
  // ---- root: Presolver::solve with LinearHandler + GpuBoundHandler against the reference pair ----
  {
    std::vector<double> xstar;
    const uint64_t keep = rng_state;
    ProblemPtr pA = makeProblem(env, 300, 320, 6, 30, xstar);
    rng_state = keep;
    ProblemPtr pB = makeProblem(env, 300, 320, 6, 30, xstar);
    int ta, tb;
    // (the reference's NlPresHandler::bin2Lin_, NlPresHandler.cpp:424-540, adds constraints while it walks them with
    //  `mult` sized for the old count: a random problem with a binary x binary product aborts inside the REFERENCE run
    //  under _FORTIFY_SOURCE -- this seed has none)
    // (informational: NlPresHandler::presolve also DELETES constraints it finds redundant at the root, chkRed_ with
    //  apply_to_prob -- a structure change GpuBoundHandler leaves to it by design, SURVEY.md 8a L12 -- so the two runs
    //  need not end with the same rows or bit-identical boxes; both must finish)
    const int d = rootPresolve(env, pA, pB, "random MINLP (300 variables, 320 rows, 30 bilinear constraints)", &ta, &tb);
    CHECK(d != -2, "root presolve (random MINLP): the two runs end with different statuses");
    delete pA; delete pB;
    if (argc > 1) {
      ProblemPtr tA = readFlat(env, argv[1]), tB = readFlat(env, argv[1]);
      CHECK(tA && tB, "cannot read %s", argv[1]);
      if (tA && tB) {
        const int dt = rootPresolve(env, tA, tB, "tls4 (BASELINE config 1)", &ta, &tb);
        CHECK(dt == 0, "root presolve (tls4): %d variables end with different bounds", dt);
      }
      delete tA; delete tB;
    }
  }
  // ---- root step: coefficient improvement (LinearHandler::coeffImp_) through GpuBoundHandler::coeffImprove ----
  {
    int n_imp = 0;
    for (int trial = 0; trial < 3; ++trial) {
      const uint64_t keep = rng_state;
      ProblemPtr pA = makeBigM(env, 30, 90, 300), pB = 0;
      rng_state = keep;
      pB = makeBigM(env, 30, 90, 300);
      CoeffProbe lhA(env, pA);
      bool chA = false, chB = false;
      lhA.run(&chA);
      GpuBoundHandler gB(env, pB, 0);
      const int k = gB.coeffImprove(&chB);
      CHECK(chA == chB, "coeffImprove: changed flags differ");
      CHECK(pA->getNumCons() == pB->getNumCons(), "coeffImprove: constraint counts differ");
      ConstraintConstIterator ia = pA->consBegin(), ib = pB->consBegin();
      int diff = 0;
      for (; ia != pA->consEnd() && ib != pB->consEnd(); ++ia, ++ib) {
        if ((*ia)->getLb() != (*ib)->getLb() || (*ia)->getUb() != (*ib)->getUb()) ++diff;
        LinearFunctionPtr la = (*ia)->getLinearFunction(), lb2 = (*ib)->getLinearFunction();
        if (la->getNumTerms() != lb2->getNumTerms()) { ++diff; continue; }
        VariableGroupConstIterator ta = la->termsBegin(), tb = lb2->termsBegin();
        for (; ta != la->termsEnd(); ++ta, ++tb)
          if (ta->first->getIndex() != tb->first->getIndex() || ta->second != tb->second) { ++diff; break; }
      }
      CHECK(diff == 0, "coeffImprove: %d rows differ from LinearHandler::coeffImp_", diff);
      CHECK(k > 10, "coeffImprove: only %d rows improved", k);
      n_imp += k;
      delete pA; delete pB;
    }
    printf("handler_test: coeffImprove = LinearHandler::coeffImp_ on 3 big-M problems (%d improved rows, coefficients and row bounds bit for bit)\n", n_imp);
  }
  printf("handler_test: %d calls with row bounds changed in between, %d p_mods mirrored into the original problem\n", n_stale, n_pmods);
  printf("handler_test: %d comparisons (%d with an incumbent cut-off), %d infeasible, %d mods emitted; "
         "%d strong-branching candidates in one batch (%d infeasible); %d failures\n",
         n_cmp, n_cut, n_inf, n_mods, n_cand, n_cand_inf, failures);
  return failures ? 1 : 0;
}
