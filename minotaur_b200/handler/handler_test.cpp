// handler_test.cpp -- TEST: GpuBoundHandler against the reference's own LinearHandler / NlPresHandler
// through Minotaur's Handler::presolveNode interface, on the same relaxation objects.
//
// Built by oracle/Makefile (target handler_test) against the reference objects in oracle/_ref; run on a GPU
// box by tests/test_gpu_handler.py.  Exit code 0 = every comparison held.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

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

int main()
{
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
    delete rel;
    delete p;
  }
  printf("handler_test: %d comparisons (%d with an incumbent cut-off), %d infeasible, %d mods emitted; "
         "%d strong-branching candidates in one batch (%d infeasible); %d failures\n",
         n_cmp, n_cut, n_inf, n_mods, n_cand, n_cand_inf, failures);
  return failures ? 1 : 0;
}
