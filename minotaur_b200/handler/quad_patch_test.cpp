// quad_patch_test.cpp -- the reference's QuadHandler with handler/quad_handler_gpu.patch applied (built by
// `make -C oracle quad_patch_test` from the reference's own sources, the patched copy living in the untracked build
// directory), run twice on the same problem and node boxes: once as it is (no GPU context attached: the host loop of
// presolveNode, QuadHandler.cpp:1214-1239, untouched by the patch) and once with a context of the engine attached
// (the loop on the device through mntr_gpu_quad_presolve_node).  Verdicts must agree on every box, and on the feasible
// ones every bound bit for bit.  Input: tests/golden/quad_node_case.txt (tests/golden/make_quad_node_case.py).
#include "MinotaurConfig.h"

// bStats_.niters is preset so that the handler behaves as at every node after its first (no tightenQuad_, :1241)
#define private public
#include "QuadHandler.h"
#undef private

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "Constraint.h"
#include "Environment.h"
#include "Function.h"
#include "LinearFunction.h"
#include "Problem.h"
#include "QuadraticFunction.h"
#include "Relaxation.h"
#include "SolutionPool.h"
#include "Types.h"
#include "Variable.h"
#include "mntr_gpu.h"

using namespace Minotaur;

namespace {

struct Case {
  int n, nSq, nBil, nBoxes;
  std::vector<uint8_t> type;
  std::vector<double> lb, ub, L, U;
  std::vector<int32_t> sqX, sqY, bX0, bX1, bY;
};

bool readCase(const char* path, Case& c)
{
  FILE* f = fopen(path, "r");
  if(!f) return false;
  bool ok = fscanf(f, "%d %d %d %d", &c.n, &c.nSq, &c.nBil, &c.nBoxes) == 4;
  auto ints = [&](std::vector<int32_t>& v, int k) { v.resize(k); for(int i = 0; i < k && ok; ++i) ok = fscanf(f, "%d", &v[i]) == 1; };
  auto dbls = [&](std::vector<double>& v, size_t at, int k) { for(int i = 0; i < k && ok; ++i) ok = fscanf(f, "%lf", &v[at + i]) == 1; };
  std::vector<int32_t> t;
  ints(t, c.n);
  c.type.assign(t.begin(), t.end());
  c.lb.resize(c.n); c.ub.resize(c.n);
  dbls(c.lb, 0, c.n); dbls(c.ub, 0, c.n);
  ints(c.sqX, c.nSq); ints(c.sqY, c.nSq); ints(c.bX0, c.nBil); ints(c.bX1, c.nBil); ints(c.bY, c.nBil);
  c.L.resize((size_t)c.nBoxes * c.n); c.U.resize((size_t)c.nBoxes * c.n);
  for(int b = 0; b < c.nBoxes; ++b) { dbls(c.L, (size_t)b * c.n, c.n); dbls(c.U, (size_t)b * c.n, c.n); }
  fclose(f);
  return ok;
}

// presolveNode on every box; gpu != 0: with the engine attached.  lb / ub receive the boxes as the handler leaves them.
void run(const Case& c, mntr_gpu_ctx* gpu, std::vector<double>& lb, std::vector<double>& ub, std::vector<int>& verdict,
         std::vector<int>& nmods)
{
  EnvPtr env = (EnvPtr) new Environment();
  int err = 0;
  env->startTimer(err);
  env->setLogLevel(LogNone);
  ProblemPtr p = (ProblemPtr) new Problem(env);
  std::vector<VariablePtr> vars;
  for(int j = 0; j < c.n; ++j) vars.push_back(p->newVariable(c.lb[j], c.ub[j], (VariableType)c.type[j]));
  QuadHandler* qh = new QuadHandler(env, p);
  for(int k = 0; k < c.nSq + c.nBil; ++k) {
    const bool sq = k < c.nSq;
    VariablePtr y = vars[sq ? c.sqY[k] : c.bY[k - c.nSq]];
    VariablePtr x0 = vars[sq ? c.sqX[k] : c.bX0[k - c.nSq]], x1 = vars[sq ? c.sqX[k] : c.bX1[k - c.nSq]];
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
  for(int j = 0; j < c.n; ++j) rel->newVariable(c.lb[j], c.ub[j], (VariableType)c.type[j], vars[j]->getName(), vars[j]->getSrcType());
  bool isInf = false;
  qh->relaxInitInc(rel, &isInf);
  if(gpu) qh->setGpuContext(gpu);
  SolutionPoolPtr pool = (SolutionPoolPtr) new SolutionPool(env, p, 1);
  lb = c.L; ub = c.U;
  verdict.assign(c.nBoxes, 0); nmods.assign(c.nBoxes, 0);
  for(int b = 0; b < c.nBoxes; ++b) {
    double *bl = &lb[(size_t)b * c.n], *bu = &ub[(size_t)b * c.n];
    for(int j = 0; j < c.n; ++j) {
      p->changeBound(vars[j], bl[j], bu[j]);
      rel->changeBound(rel->getRelaxationVar(vars[j]), bl[j], bu[j]);
    }
    qh->bStats_.niters = 1;
    ModVector pMods, rMods;
    verdict[b] = qh->presolveNode(rel, (NodePtr)0, pool, pMods, rMods) ? 1 : 0;
    nmods[b] = (int)pMods.size();
    for(int j = 0; j < c.n; ++j) {
      bl[j] = vars[j]->getLb(); bu[j] = vars[j]->getUb();
      // the relaxation's copy moved with the problem's (modRel_)
      if(!verdict[b] && (rel->getRelaxationVar(vars[j])->getLb() != bl[j] || rel->getRelaxationVar(vars[j])->getUb() != bu[j])) nmods[b] = -1;
    }
    for(size_t k = 0; k < pMods.size(); ++k) delete pMods[k];
    for(size_t k = 0; k < rMods.size(); ++k) delete rMods[k];
  }
  delete pool;
  delete qh;
  delete rel;
  delete p;
  delete env;
}

} // namespace

int main(int argc, char** argv)
{
  setvbuf(stdout, 0, _IOLBF, 0);
  if(argc < 2) { fprintf(stderr, "usage: quad_patch_test <tests/golden/quad_node_case.txt>\n"); return 2; }
  Case c;
  if(!readCase(argv[1], c)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  mntr_gpu_ctx* gpu = 0;
  if(mntr_gpu_create(0, &gpu) != MNTR_OK) { fprintf(stderr, "no CUDA device: %s\n", mntr_gpu_last_error(0)); return 2; }
  // the engine needs the problem's variables (their types); this handler has no linear rows for it
  const int32_t rowPtr0 = 0;
  if(mntr_gpu_load_linear(gpu, 0, c.n, &rowPtr0, 0, 0, 0, 0, &c.type[0], 0) != MNTR_OK) {
    fprintf(stderr, "load_linear: %s\n", mntr_gpu_last_error(gpu));
    return 2;
  }
  std::vector<double> hl, hu, gl, gu;
  std::vector<int> hv, gv, hm, gm;
  run(c, 0, hl, hu, hv, hm);
  run(c, gpu, gl, gu, gv, gm);
  int failures = 0, nInf = 0, moved = 0;
  for(int b = 0; b < c.nBoxes; ++b) {
    if(hv[b] != gv[b]) { printf("box %d: verdict host %d, device %d\n", b, hv[b], gv[b]); ++failures; continue; }
    if(hv[b]) { ++nInf; continue; }
    if(gm[b] < 0) { printf("box %d: the relaxation's bounds differ from the problem's\n", b); ++failures; }
    for(int j = 0; j < c.n; ++j) {
      const size_t k = (size_t)b * c.n + j;
      if(memcmp(&hl[k], &gl[k], 8) || memcmp(&hu[k], &gu[k], 8)) {
        if(failures < 10) printf("box %d var %d: host [%.17g, %.17g] device [%.17g, %.17g]\n", b, j, hl[k], hu[k], gl[k], gu[k]);
        ++failures;
      }
      moved += hl[k] != c.L[k] || hu[k] != c.U[k];
    }
  }
  mntr_gpu_destroy(gpu);
  printf("quad_patch_test: %d boxes (%d infeasible), %d bounds moved by QuadHandler::presolveNode, patched handler on the device "
         "vs the same handler on the host: %d failures\n", c.nBoxes, nInf, moved, failures);
  return failures ? 1 : 0;
}
