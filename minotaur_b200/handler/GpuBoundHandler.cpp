//
//     GpuBoundHandler -- B200 bound-tightening handler for Minotaur
//
/**
 * \file GpuBoundHandler.cpp
 * \brief Host adapter between Minotaur's Handler plugin API and the C ABI of libmntr_gpu.so.
 * No bound arithmetic happens here: this file only flattens the object graph, moves bounds and turns
 * the tightened box into VarBoundMods.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <map>
#include <cstring>
#include <stdexcept>

#include "MinotaurConfig.h"
#include "CGraph.h"
#include "CNode.h"
#include "Constraint.h"
#include "Environment.h"
#include "Function.h"
#include "GpuBoundHandler.h"
#include "LinearFunction.h"
#include "Logger.h"
#include "Node.h"
#include "NonlinearFunction.h"
#include "Objective.h"
#include "Problem.h"
#include "QuadraticFunction.h"
#include "Relaxation.h"
#include "SolutionPool.h"
#include "Timer.h"
#include "VarBoundMod.h"
#include "Variable.h"
#include "mntr_gpu.h"

using namespace Minotaur;

const std::string GpuBoundHandler::me_ = "GpuBoundHandler: ";

GpuBoundHandler::GpuBoundHandler(EnvPtr env, ProblemPtr problem, int device)
  : env_(env), problem_(problem), ctx_(0), group_(0), mode_(FastFixpoint), roundNearest_(false), checkStructure_(true),
    loadedFor_(0), loadedVars_(0), loadedCons_(0), cutoffOn_(false), sigStruct_(0), sigBounds_(0), lb_(0), ub_(0), boxCap_(0),
    cacheFor_(0)
{
  init_(std::vector<int>(1, device));
}

GpuBoundHandler::GpuBoundHandler(EnvPtr env, ProblemPtr problem, const std::vector<int> &devices)
  : env_(env), problem_(problem), ctx_(0), group_(0), mode_(FastFixpoint), roundNearest_(false), checkStructure_(true),
    loadedFor_(0), loadedVars_(0), loadedCons_(0), cutoffOn_(false), sigStruct_(0), sigBounds_(0), lb_(0), ub_(0), boxCap_(0),
    cacheFor_(0)
{
  init_(devices);
}

void GpuBoundHandler::init_(const std::vector<int> &devices)
{
  logger_ = env_->getLogger();
  // Handler leaves modProb_ / modRel_ uninitialised (Handler.h:380-383; LinearHandler's constructor does not set them
  // either): start from what Bnb sets for the node handlers (setModFlags(false, true), Bnb.cpp:111,121)
  modProb_ = false;
  modRel_ = true;
  stats_.calls = stats_.uploads = stats_.nMods = stats_.nInf = 0;
  stats_.rowBoundUpdates = stats_.skippedCons = stats_.engineErrors = stats_.cacheHits = 0;
  stats_.nnzUpdates = 0;
  stats_.timeHost = stats_.timeDevice = 0.;
  int rc;
  if (devices.size() > 1) {
    rc = mntr_gpu_group_create((int)devices.size(), &devices[0], &group_);
    if (rc == MNTR_OK) ctx_ = mntr_gpu_group_member(group_, 0);
  } else {
    rc = mntr_gpu_create(devices.empty() ? 0 : devices[0], &ctx_);
  }
  if (rc != MNTR_OK || !ctx_) {
    ctx_ = 0; group_ = 0;
    // the reference's convention for unusable components is an assert / exception at set-up time (never later)
    throw std::runtime_error("GpuBoundHandler: no usable CUDA device (mntr_gpu_create failed); "
                             "use LinearHandler/NlPresHandler instead");
  }
}

GpuBoundHandler::~GpuBoundHandler()
{
  if (ctx_) { mntr_gpu_free_host(ctx_, lb_); mntr_gpu_free_host(ctx_, ub_); }
  if (group_) mntr_gpu_group_destroy(group_);
  else if (ctx_) mntr_gpu_destroy(ctx_);
}

// an engine call failed: say so once per call site, count it, and let the caller carry on without tightening
void GpuBoundHandler::engineFailed_(const char *where)
{
  ++stats_.engineErrors;
  logger_->errStream() << me_ << where << ": " << (ctx_ ? mntr_gpu_last_error(ctx_) : "no context")
                       << " -- no bound tightening from this call" << std::endl;
  loadedFor_ = 0;
}

int GpuBoundHandler::coeffImprove(bool *changed)
{
  ProblemPtr p = problem_;
  const UInt n = p->getNumVars();
  std::vector<int> rowPtr(1, 0), col;
  std::vector<double> val, rowLb, rowUb, lb(n), ub(n);
  std::vector<unsigned char> vtype(n);
  std::vector<ConstraintPtr> rows;
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    const UInt j = (*it)->getIndex();
    vtype[j] = (unsigned char)(*it)->getType(); lb[j] = (*it)->getLb(); ub[j] = (*it)->getUb();
  }
  // the rows coeffImp_ and computeImpBounds_ look at: Linear constraints, in index order (:616-620, :739-742)
  for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) {
    ConstraintPtr c = *it;
    if (p->isMarkedDel(c) || c->getFunctionType() != Linear) continue;
    LinearFunctionPtr lf = c->getLinearFunction();
    if (!lf) continue;
    for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
      col.push_back((int)t->first->getIndex());
      val.push_back(t->second);
    }
    rowPtr.push_back((int)col.size());
    rowLb.push_back(c->getLb()); rowUb.push_back(c->getUb());
    rows.push_back(c);
  }
  const int m = (int)rows.size();
  if (m == 0) return 0;
  const long long cap = m;
  std::vector<int> oRow((size_t)cap), oVar((size_t)cap), oSide((size_t)cap);
  std::vector<double> oCoef((size_t)cap), oBnd((size_t)cap), oDelta((size_t)cap);
  long long cnt = 0;
  int levels = 0, erased = 0;
  if (mntr_gpu_root_coeff_imp(ctx_, m, (int)n, &rowPtr[0], col.empty() ? 0 : &col[0], val.empty() ? 0 : &val[0], &rowLb[0],
                              &rowUb[0], &vtype[0], &lb[0], &ub[0], cap, &oRow[0], &oVar[0], &oCoef[0], &oSide[0], &oBnd[0],
                              &oDelta[0], (int64_t *)&cnt, &levels, &erased) != 0) {
    engineFailed_("coeffImprove");
    return 0;
  }
  for (long long k = 0; k < cnt; ++k) {
    ConstraintPtr c = rows[(size_t)oRow[(size_t)k]];
    c->getLinearFunction()->incTerm(p->getVariable((UInt)oVar[(size_t)k]), oDelta[(size_t)k]);     // :652, :660, :680, :689
    if (oSide[(size_t)k] == 2) p->changeBound(c, Upper, oBnd[(size_t)k]);
    else if (oSide[(size_t)k] == 1) p->changeBound(c, Lower, oBnd[(size_t)k]);
    c->setBFlag(true);
  }
  if (cnt > 0) { if (changed) *changed = true; invalidate(); }
  return (int)cnt;
}

std::string GpuBoundHandler::getName() const { return "GpuBoundHandler (FBBT on B200)"; }

namespace {

// CGraph -> tape in the evaluation order of CGraph::computeBounds: variable nodes by ascending variable id,
// constants, then the dependent nodes dq_ (public accessor dNodes(), CGraph.h:202).
bool flattenCGraph(CGraph *cg, std::vector<unsigned char> &op, std::vector<int> &a0, std::vector<int> &a1,
                   std::vector<double> &cn, std::vector<int> &child)
{
  CNodeQ dq = cg->dNodes();
  if (dq.empty()) return false;
  std::map<const CNode *, int> index;
  std::map<UInt, const CNode *> vars;          // ascending variable id
  std::vector<const CNode *> consts;
  for (CNodeQ::iterator it = dq.begin(); it != dq.end(); ++it) {
    const CNode *nd = *it;
    std::vector<const CNode *> kids;
    if (nd->numChild() > 2 || nd->getOp() == OpSumList) {
      for (CNode **c = nd->getListL(); c != nd->getListR(); ++c) kids.push_back(*c);
    } else {
      if (nd->getL()) kids.push_back(nd->getL());
      if (nd->getR()) kids.push_back(nd->getR());
    }
    for (size_t k = 0; k < kids.size(); ++k) {
      const CNode *c = kids[k];
      if (c->getOp() == OpVar) vars[c->getV()->getId()] = c;
      else if ((c->getOp() == OpNum || c->getOp() == OpInt) && index.find(c) == index.end()) {
        index[c] = -1;
        consts.push_back(c);
      }
    }
  }
  const size_t base = op.size();
  int next = 0;
  for (std::map<UInt, const CNode *>::iterator it = vars.begin(); it != vars.end(); ++it) {
    index[it->second] = next++;
    op.push_back((unsigned char)OpVar); a0.push_back((int)it->second->getV()->getIndex()); a1.push_back(-1);
    cn.push_back(0.);
  }
  for (size_t k = 0; k < consts.size(); ++k) {
    index[consts[k]] = next++;
    op.push_back((unsigned char)consts[k]->getOp()); a0.push_back(-1); a1.push_back(-1);
    cn.push_back(consts[k]->getVal());
  }
  for (CNodeQ::iterator it = dq.begin(); it != dq.end(); ++it) {
    const CNode *nd = *it;
    index[nd] = next++;
    op.push_back((unsigned char)nd->getOp());
    cn.push_back(0.);
    if (nd->getOp() == OpSumList) {
      a0.push_back((int)child.size());
      for (CNode **c = nd->getListL(); c != nd->getListR(); ++c) child.push_back(index[*c]);
      a1.push_back((int)child.size());
    } else {
      a0.push_back(nd->getL() ? index[nd->getL()] : -1);
      a1.push_back(nd->getR() ? index[nd->getR()] : -1);
    }
  }
  (void)base;
  return true;
}

}  // namespace

// FNV-1a over the words that decide whether the device copy is current
static inline void fnv(unsigned long long &h, unsigned long long x)
{
  for (int k = 0; k < 8; ++k) { h ^= (x >> (8 * k)) & 0xffull; h *= 1099511628211ull; }
}

void GpuBoundHandler::signature_(ProblemPtr p, unsigned long long &structure, unsigned long long &bounds) const
{
  structure = 1469598103934665603ull; bounds = 1469598103934665603ull;
  fnv(structure, p->getNumVars()); fnv(structure, p->getNumCons());
  for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) {
    ConstraintPtr c = *it;
    LinearFunctionPtr lf = c->getLinearFunction();
    fnv(structure, (unsigned long long)c->getState() * 8u + (unsigned long long)c->getFunctionType());
    fnv(structure, lf ? lf->getNumTerms() : 0u);
    fnv(structure, (unsigned long long)(size_t)c->getNonlinearFunction());
    double lb = c->getLb(), ub = c->getUb();
    unsigned long long bl, bu;
    memcpy(&bl, &lb, 8); memcpy(&bu, &ub, 8);
    // the bounds of the linear rows can be refreshed alone; those of CGraph constraints travel with their tapes
    if (c->getNonlinearFunction() || c->getQuadraticFunction()) { fnv(structure, bl); fnv(structure, bu); }
    else { fnv(bounds, bl); fnv(bounds, bu); }
  }
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) fnv(structure, (unsigned long long)(*it)->getType());
}

void GpuBoundHandler::sync_(ProblemPtr p)
{
  if (loadedFor_ != p || loadedVars_ != p->getNumVars() || loadedCons_ != p->getNumCons()) { upload_(p); return; }
  if (!checkStructure_) return;
  unsigned long long st, bd;
  signature_(p, st, bd);
  if (st != sigStruct_) { upload_(p); return; }
  if (bd != sigBounds_) {
    // same rows, new bounds of LINEAR rows (a ConBoundMod, LinearHandler's row-bound tightening): refresh them alone
    // (the bounds of CGraph constraints are part of the structure signature: they re-flatten)
    std::vector<double> rl, ru;
    for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) {
      ConstraintPtr c = *it;
      if (DeletedCons == c->getState()) continue;
      if (c->getFunctionType() == Linear && c->getQuadraticFunction() == 0 && c->getNonlinearFunction() == 0) {
        rl.push_back(c->getLb()); ru.push_back(c->getUb());
      }
    }
    int rc = MNTR_OK;
    const int k = group_ ? mntr_gpu_group_size(group_) : 1;
    for (int i = 0; i < k && rc == MNTR_OK; ++i)
      rc = mntr_gpu_update_row_bounds(group_ ? mntr_gpu_group_member(group_, i) : ctx_, (int)rl.size(),
                                      rl.empty() ? 0 : &rl[0], ru.empty() ? 0 : &ru[0]);
    if (rc != MNTR_OK) { upload_(p); return; }
    sigBounds_ = bd;
    ++stats_.rowBoundUpdates;
  }
}

void GpuBoundHandler::upload_(ProblemPtr p)
{
  const UInt n = p->getNumVars();
  std::vector<int> rowPtr(1, 0), col, tapePtr(1, 0), a0, a1, child, linPtr(1, 0), linCol;
  std::vector<double> val, rowLb, rowUb, cn, linVal, cLb, cUb;
  std::vector<unsigned char> vtype(n), op;
  std::vector<int> qPtr(1, 0), qV1, qV2, qLinPtr(1, 0), qLinCol;
  std::vector<double> qCoef, qLinVal, qLb, qUb;
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it)
    vtype[(*it)->getIndex()] = (unsigned char)(*it)->getType();
  stats_.skippedCons = 0;

  for (ConstraintConstIterator it = p->consBegin(); it != p->consEnd(); ++it) {
    ConstraintPtr c = *it;
    if (DeletedCons == c->getState()) continue;
    LinearFunctionPtr lf = c->getLinearFunction();
    if (c->getFunctionType() == Linear && c->getQuadraticFunction() == 0 && c->getNonlinearFunction() == 0) {
      // LinearHandler::varBndsFromCons_ row filter, LinearHandler.cpp:509-511; terms in ascending variable id
      if (lf)
        for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
          col.push_back((int)t->first->getIndex());
          val.push_back(t->second);
        }
      rowPtr.push_back((int)col.size());
      rowLb.push_back(c->getLb());
      rowUb.push_back(c->getUb());
    } else if (c->getFunctionType() != Constant && c->getNonlinearFunction() && !c->getQuadraticFunction()) {
      // NlPresHandler::varBndsFromCons_ nlf branch, NlPresHandler.cpp:1771-1778 (native CGraph only).  A constraint
      // the engine cannot take -- a tape longer than MNTR_GPU_MAX_TAPE nodes, a non-CGraph function -- is SKIPPED
      // and counted: skipping a constraint only loses tightening, it never makes a derived bound invalid.
      CGraph *cg = dynamic_cast<CGraph *>(c->getNonlinearFunction());
      const size_t keep = op.size(), keepChild = child.size();
      if (!cg || !flattenCGraph(cg, op, a0, a1, cn, child) || op.size() - keep > (size_t)MNTR_GPU_MAX_TAPE) {
        op.resize(keep); a0.resize(keep); a1.resize(keep); cn.resize(keep); child.resize(keepChild);
        ++stats_.skippedCons;
        continue;
      }
      tapePtr.push_back((int)op.size());
      if (lf)
        for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
          linCol.push_back((int)t->first->getIndex());
          linVal.push_back(t->second);
        }
      linPtr.push_back((int)linCol.size());
      cLb.push_back(c->getLb());
      cUb.push_back(c->getUb());
    } else if (c->getFunctionType() != Constant && c->getQuadraticFunction()) {
      // QuadraticFunction constraints are only CHECKED by NlPresHandler (chkRed_, NlPresHandler.cpp:127-150: the qf
      // bounds take precedence over an nlf part, which chkRed_ then ignores); terms in VariablePairGroup order
      QuadraticFunctionPtr qf = c->getQuadraticFunction();
      for (VariablePairGroupConstIterator t = qf->begin(); t != qf->end(); ++t) {
        qV1.push_back((int)t->first.first->getIndex());
        qV2.push_back((int)t->first.second->getIndex());
        qCoef.push_back(t->second);
      }
      qPtr.push_back((int)qV1.size());
      if (lf)
        for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t) {
          qLinCol.push_back((int)t->first->getIndex());
          qLinVal.push_back(t->second);
        }
      qLinPtr.push_back((int)qLinCol.size());
      qLb.push_back(c->getLb());
      qUb.push_back(c->getUb());
    } else if (c->getFunctionType() != Constant && c->getFunctionType() != Linear) {
      ++stats_.skippedCons;
    }
  }
  // CSR columns must ascend by variable INDEX; LinearFunction orders by id, which equals the index unless
  // variables were deleted -- sort defensively
  for (size_t i = 0; i + 1 < rowPtr.size(); ++i) {
    bool sorted = true;
    for (int t = rowPtr[i] + 1; t < rowPtr[i + 1]; ++t) if (col[t] <= col[t - 1]) { sorted = false; break; }
    if (!sorted) {
      std::vector<std::pair<int, double> > tmp;
      for (int t = rowPtr[i]; t < rowPtr[i + 1]; ++t) tmp.push_back(std::make_pair(col[t], val[t]));
      std::sort(tmp.begin(), tmp.end());
      for (int t = rowPtr[i]; t < rowPtr[i + 1]; ++t) { col[t] = tmp[t - rowPtr[i]].first; val[t] = tmp[t - rowPtr[i]].second; }
    }
  }
  const int m = (int)rowPtr.size() - 1;
  const int nc = (int)tapePtr.size() - 1;
  if (child.empty()) child.push_back(0);
  if (linCol.empty()) { linCol.push_back(0); linVal.push_back(0.); }
  const int nq = (int)qPtr.size() - 1;
  if (qV1.empty()) { qV1.push_back(0); qV2.push_back(0); qCoef.push_back(0.); }
  if (qLinCol.empty()) { qLinCol.push_back(0); qLinVal.push_back(0.); }
  int rc;
  if (group_) {
    rc = mntr_gpu_group_load_linear(group_, m, (int)n, &rowPtr[0], col.empty() ? 0 : &col[0], val.empty() ? 0 : &val[0],
                                    rowLb.empty() ? 0 : &rowLb[0], rowUb.empty() ? 0 : &rowUb[0], vtype.empty() ? 0 : &vtype[0], 0);
    if (rc == MNTR_OK && nc > 0)
      rc = mntr_gpu_group_load_cgraph(group_, nc, &tapePtr[0], &op[0], &a0[0], &a1[0], &cn[0], &child[0], &linPtr[0],
                                      &linCol[0], &linVal[0], &cLb[0], &cUb[0]);
    if (rc == MNTR_OK && nq > 0)
      rc = mntr_gpu_group_load_quad(group_, nq, &qPtr[0], &qV1[0], &qV2[0], &qCoef[0], &qLinPtr[0], &qLinCol[0], &qLinVal[0],
                                    &qLb[0], &qUb[0]);
  } else {
    rc = mntr_gpu_load_linear(ctx_, m, (int)n, &rowPtr[0], col.empty() ? 0 : &col[0], val.empty() ? 0 : &val[0],
                              rowLb.empty() ? 0 : &rowLb[0], rowUb.empty() ? 0 : &rowUb[0], vtype.empty() ? 0 : &vtype[0], 0);
    if (rc == MNTR_OK && nc > 0)
      rc = mntr_gpu_load_cgraph(ctx_, nc, &tapePtr[0], &op[0], &a0[0], &a1[0], &cn[0], &child[0], &linPtr[0],
                                &linCol[0], &linVal[0], &cLb[0], &cUb[0]);
    if (rc == MNTR_OK && nq > 0)
      rc = mntr_gpu_load_quad(ctx_, nq, &qPtr[0], &qV1[0], &qV2[0], &qCoef[0], &qLinPtr[0], &qLinCol[0], &qLinVal[0], &qLb[0], &qUb[0]);
  }
  if (rc != MNTR_OK) throw std::runtime_error(std::string(me_) + (group_ ? mntr_gpu_group_last_error(group_) : mntr_gpu_last_error(ctx_)));
  loadedFor_ = p;
  loadedVars_ = p->getNumVars();
  loadedCons_ = p->getNumCons();
  signature_(p, sigStruct_, sigBounds_);
  cutoffOn_ = false;
  if (n > boxCap_) {
    mntr_gpu_free_host(ctx_, lb_); mntr_gpu_free_host(ctx_, ub_);
    lb_ = (double *)mntr_gpu_alloc_host(ctx_, (long long)sizeof(double) * std::max<UInt>(n, 1));
    ub_ = (double *)mntr_gpu_alloc_host(ctx_, (long long)sizeof(double) * std::max<UInt>(n, 1));
    if (!lb_ || !ub_) throw std::runtime_error(std::string(me_) + "mntr_gpu_alloc_host failed");
    boxCap_ = n;
  }
  lb0_.resize(n); ub0_.resize(n);
  if (stats_.skippedCons > 0)
    logger_->msgStream(LogInfo) << me_ << stats_.skippedCons << " constraint(s) are not taken by the GPU engine (tape longer than "
                                << MNTR_GPU_MAX_TAPE << " nodes, quadratic or non-CGraph function); they stay with NlPresHandler / QuadHandler"
                                << std::endl;
  ++stats_.uploads;
}

// The incumbent's cut-off row  c.x <= best - constant  (LinearHandler::simplePresolve, LinearHandler.cpp:1636-1640,
// varBndsFromObj_ :544-597): only for a linear objective, only when the pool holds a solution.
void GpuBoundHandler::setCutoff_(ProblemPtr p, SolutionPoolPtr spool)
{
  ObjectivePtr o = p->getObjective();
  LinearFunctionPtr lf = o ? o->getLinearFunction() : LinearFunctionPtr();
  const bool on = spool && spool->getNumSols() > 0 && lf && o->getFunctionType() == Linear;
  if (!on) {
    if (cutoffOn_) {
      const int rc0 = group_ ? mntr_gpu_group_set_cutoff(group_, 0, 0, 0, 0.0) : mntr_gpu_set_cutoff(ctx_, 0, 0, 0, 0.0);
      if (rc0 != MNTR_OK) throw std::runtime_error(std::string(me_) + mntr_gpu_last_error(ctx_));
      cutoffOn_ = false;
    }
    return;
  }
  std::vector<std::pair<int, double> > terms;
  for (VariableGroupConstIterator t = lf->termsBegin(); t != lf->termsEnd(); ++t)
    terms.push_back(std::make_pair((int)t->first->getIndex(), t->second));
  std::sort(terms.begin(), terms.end());
  std::vector<int> col(terms.size());
  std::vector<double> val(terms.size());
  for (size_t t = 0; t < terms.size(); ++t) { col[t] = terms[t].first; val[t] = terms[t].second; }
  const double rhs = spool->getBestSolutionValue() - o->getConstant();
  const int rc1 = group_ ? mntr_gpu_group_set_cutoff(group_, (int)col.size(), col.empty() ? 0 : &col[0], val.empty() ? 0 : &val[0], rhs)
                         : mntr_gpu_set_cutoff(ctx_, (int)col.size(), col.empty() ? 0 : &col[0], val.empty() ? 0 : &val[0], rhs);
  if (rc1 != MNTR_OK) throw std::runtime_error(std::string(me_) + mntr_gpu_last_error(ctx_));
  cutoffOn_ = !col.empty();
  // NlPresHandler::fixObjBins_ compares against the raw pool value (NlPresHandler.cpp:1030)
  const int rc2 = group_ ? mntr_gpu_group_set_incumbent(group_, spool->getBestSolutionValue())
                         : mntr_gpu_set_incumbent(ctx_, spool->getBestSolutionValue());
  if (rc2 != MNTR_OK) throw std::runtime_error(std::string(me_) + mntr_gpu_last_error(ctx_));
}

bool GpuBoundHandler::tighten_(ProblemPtr p, SolutionPoolPtr spool, ModVector &mods, bool truncated)
{
  // Nothing on this path may abort the solve (no exceptions on the node path): an engine failure is
  // logged and counted, and the call reports "no tightening".
  try {
    sync_(p);
    setCutoff_(p, spool);
  } catch (const std::exception &) {
    engineFailed_("upload");
    return false;
  }
  const UInt n = p->getNumVars();
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    const UInt j = (*it)->getIndex();
    lb0_[j] = lb_[j] = (*it)->getLb();
    ub0_[j] = ub_[j] = (*it)->getUb();
  }
  mntr_gpu_options o;
  o.rounding = roundNearest_ ? MNTR_ROUND_NEAREST : MNTR_ROUND_DIRECTED;
  o.order = (mode_ == ReferenceOrder) ? MNTR_ORDER_REFERENCE : MNTR_ORDER_JACOBI;
  o.loop = (mode_ == ReferenceOrder && truncated) ? MNTR_LOOP_SIMPLEPRESOLVE : MNTR_LOOP_FIXPOINT;
  o.max_rounds = 0;
  o.handlers = MNTR_HANDLERS_ALL;
  o.flags = 0;
  o.reserved[0] = o.reserved[1] = 0;
  int verdict = 0, rounds = 0;
  long long nnz = 0;
  int rc = mntr_gpu_tighten(ctx_, 1, lb_, ub_, &o, &verdict, &rounds, (int64_t *)&nnz);
  if (rc == MNTR_E_UNSUPPORTED && o.order == MNTR_ORDER_JACOBI) {
    // CGraph constraints are evaluated by the reference-order kernel
    o.order = MNTR_ORDER_REFERENCE;
    rc = mntr_gpu_tighten(ctx_, 1, lb_, ub_, &o, &verdict, &rounds, (int64_t *)&nnz);
  }
  if (rc != MNTR_OK) { engineFailed_("tighten"); return false; }
  mntr_gpu_stats st;
  mntr_gpu_get_stats(ctx_, &st);
  stats_.timeDevice += st.kernel_ms + st.h2d_ms + st.d2h_ms;
  stats_.nnzUpdates += nnz;
  ++stats_.calls;
  if (verdict != MNTR_FEASIBLE) { ++stats_.nInf; return true; }
  // one VarBoundMod per changed (variable, side), applied at once: the final box is what matters
  // (VarBoundMod::oldVal_ is captured by its constructor, VarBoundMod.cpp:27-43)
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    VariablePtr v = *it;
    const UInt j = v->getIndex();
    if (lb_[j] != lb0_[j]) {
      VarBoundModPtr mod = (VarBoundModPtr) new VarBoundMod(v, Lower, lb_[j]);
      mod->applyToProblem(p);
      mods.push_back(mod);
      ++stats_.nMods;
    }
    if (ub_[j] != ub0_[j]) {
      VarBoundModPtr mod = (VarBoundModPtr) new VarBoundMod(v, Upper, ub_[j]);
      mod->applyToProblem(p);
      mods.push_back(mod);
      ++stats_.nMods;
    }
  }
  return false;
}

void GpuBoundHandler::tightenCandidates(RelaxationPtr rel, SolutionPoolPtr spool,
                                        const std::vector<std::vector<BoundChange> > &deltas,
                                        std::vector<BoxOutcome> &out)
{
  ProblemPtr p = rel;
  const int nb = (int)deltas.size();
  out.assign(deltas.size(), BoxOutcome());
  if (nb == 0) return;
  try {
    sync_(p);
    setCutoff_(p, spool);
  } catch (const std::exception &) {
    engineFailed_("upload");
    return;                          // every outcome: feasible, no changes
  }
  const UInt n = p->getNumVars();
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    const UInt j = (*it)->getIndex();
    lb_[j] = (*it)->getLb();
    ub_[j] = (*it)->getUb();
  }
  std::vector<int64_t> dptr(1, 0);
  std::vector<int> dvar;
  std::vector<unsigned char> dup;
  std::vector<double> dval;
  for (int b = 0; b < nb; ++b) {
    for (size_t k = 0; k < deltas[b].size(); ++k) {
      dvar.push_back((int)deltas[b][k].var);
      dup.push_back(deltas[b][k].lu == Upper ? 1 : 0);
      dval.push_back(deltas[b][k].val);
    }
    dptr.push_back((int64_t)dvar.size());
  }
  if (dvar.empty()) { dvar.push_back(0); dup.push_back(0); dval.push_back(0.); }
  mntr_gpu_options o;
  o.rounding = roundNearest_ ? MNTR_ROUND_NEAREST : MNTR_ROUND_DIRECTED;
  o.order = MNTR_ORDER_REFERENCE;
  o.loop = (mode_ == ReferenceOrder) ? MNTR_LOOP_SIMPLEPRESOLVE : MNTR_LOOP_FIXPOINT;
  o.max_rounds = 0;
  o.handlers = MNTR_HANDLERS_ALL;
  o.flags = 0;
  o.reserved[0] = o.reserved[1] = 0;
  std::vector<int> verdict(nb), rounds(nb);
  std::vector<int64_t> mptr(nb + 1, 0);
  int64_t cap = std::max<int64_t>(1024, 8 * (int64_t)nb), total = 0;
  std::vector<int> mvar;
  std::vector<unsigned char> mup;
  std::vector<double> mval;
  for (int attempt = 0; attempt < 2; ++attempt) {          // second trip only when the tuple buffer was too small
    mvar.assign((size_t)cap, 0); mup.assign((size_t)cap, 0); mval.assign((size_t)cap, 0.);
    const int rc = group_
        ? mntr_gpu_group_tighten_nodes(group_, nb, lb_, ub_, &dptr[0], &dvar[0], &dup[0], &dval[0], &o, &verdict[0],
                                       &rounds[0], &mptr[0], &mvar[0], &mup[0], &mval[0], cap, &total)
        : mntr_gpu_tighten_nodes(ctx_, nb, lb_, ub_, &dptr[0], &dvar[0], &dup[0], &dval[0], &o, &verdict[0], &rounds[0],
                                 &mptr[0], &mvar[0], &mup[0], &mval[0], cap, &total);
    if (rc != MNTR_OK) { engineFailed_("tighten_nodes"); return; }
    if (total <= cap) break;
    cap = total;
  }
  mntr_gpu_stats st;
  mntr_gpu_get_stats(ctx_, &st);
  stats_.timeDevice += st.kernel_ms + st.h2d_ms + st.d2h_ms;
  stats_.calls += nb;
  for (int b = 0; b < nb; ++b) {
    out[b].infeasible = verdict[b] != MNTR_FEASIBLE;
    if (out[b].infeasible) { ++stats_.nInf; continue; }
    for (int64_t q = mptr[b]; q < mptr[b + 1]; ++q) {
      BoundChange c;
      c.var = (UInt)mvar[(size_t)q]; c.lu = mup[(size_t)q] ? Upper : Lower; c.val = mval[(size_t)q];
      out[b].changes.push_back(c);
      ++stats_.nMods;
    }
  }
}

void GpuBoundHandler::prefetchCandidates(RelaxationPtr rel, SolutionPoolPtr spool,
                                         const std::vector<std::vector<BoundChange> > &deltas)
{
  clearCandidates();
  if (deltas.empty()) return;
  tightenCandidates(rel, spool, deltas, cacheOut_);
  if (cacheOut_.size() != deltas.size()) { cacheOut_.clear(); return; }
  ProblemPtr p = rel;
  const UInt n = p->getNumVars();
  cacheLb_.resize(n); cacheUb_.resize(n);
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    cacheLb_[(*it)->getIndex()] = (*it)->getLb();
    cacheUb_[(*it)->getIndex()] = (*it)->getUb();
  }
  cacheDeltas_ = deltas;
  cacheFor_ = p;
}

void GpuBoundHandler::clearCandidates()
{
  cacheFor_ = 0;
  cacheDeltas_.clear(); cacheOut_.clear();
}

int GpuBoundHandler::findCached_(ProblemPtr p) const
{
  if (cacheFor_ != p || cacheLb_.size() != p->getNumVars()) return -1;
  // the bounds that differ from the cached base box
  std::vector<BoundChange> diff;
  for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd(); ++it) {
    const UInt j = (*it)->getIndex();
    if ((*it)->getLb() != cacheLb_[j]) { BoundChange c = { j, Lower, (*it)->getLb() }; diff.push_back(c); }
    if ((*it)->getUb() != cacheUb_[j]) { BoundChange c = { j, Upper, (*it)->getUb() }; diff.push_back(c); }
    if (diff.size() > 8) return -1;
  }
  if (diff.empty()) return -1;
  for (size_t b = 0; b < cacheDeltas_.size(); ++b) {
    const std::vector<BoundChange> &d = cacheDeltas_[b];
    if (d.size() != diff.size()) continue;
    bool same = true;
    for (size_t k = 0; k < d.size() && same; ++k) {
      bool found = false;
      for (size_t q = 0; q < diff.size(); ++q)
        if (diff[q].var == d[k].var && diff[q].lu == d[k].lu && diff[q].val == d[k].val) { found = true; break; }
      same = found;
    }
    if (same) return (int)b;
  }
  return -1;
}

void GpuBoundHandler::simplePresolve(ProblemPtr p, SolutionPoolPtr spool, ModVector &t_mods, SolveStatus &status)
{
  Timer *timer = env_->getNewTimer();
  timer->start();
  if (cacheFor_) {
    const int b = findCached_(p);
    if (b >= 0) {
      // the outcome mntr_gpu_tighten_nodes computed for exactly this box: emit it as applied VarBoundMods
      ++stats_.cacheHits; ++stats_.calls;
      if (cacheOut_[(size_t)b].infeasible) { ++stats_.nInf; status = SolvedInfeasible; }
      else
        for (size_t k = 0; k < cacheOut_[(size_t)b].changes.size(); ++k) {
          const BoundChange &c = cacheOut_[(size_t)b].changes[k];
          VarBoundModPtr mod = (VarBoundModPtr) new VarBoundMod(p->getVariable(c.var), c.lu, c.val);
          mod->applyToProblem(p);
          t_mods.push_back(mod);
          ++stats_.nMods;
        }
      stats_.timeHost += timer->query();
      delete timer;
      return;
    }
    if (cacheFor_ == p) {
      // `p` is neither the base box nor one of the candidates: the outcomes are stale
      bool isBase = true;
      for (VariableConstIterator it = p->varsBegin(); it != p->varsEnd() && isBase; ++it)
        isBase = (*it)->getLb() == cacheLb_[(*it)->getIndex()] && (*it)->getUb() == cacheUb_[(*it)->getIndex()];
      if (!isBase) clearCandidates();
    }
  }
  const bool inf = tighten_(p, spool, t_mods, true);
  if (inf) status = SolvedInfeasible;
  stats_.timeHost += timer->query();
  delete timer;
}

bool GpuBoundHandler::presolveNode(RelaxationPtr rel, NodePtr, SolutionPoolPtr spool, ModVector &p_mods,
                                   ModVector &r_mods)
{
  SolveStatus status = Started;
  simplePresolve(rel, spool, r_mods, status);
  if (true == modProb_) copyBndsFromRel_(rel, p_mods);      // as LinearHandler::presolveNode, :1592-1602
  return (status == SolvedInfeasible);
}

// LinearHandler::copyBndsFromRel_, LinearHandler.cpp:108-132
void GpuBoundHandler::copyBndsFromRel_(RelaxationPtr rel, ModVector &p_mods)
{
  const double eTol = 1e-8;
  for (VariableConstIterator it = problem_->varsBegin(); it != problem_->varsEnd(); ++it) {
    VariablePtr xp = *it;
    VariablePtr xr = rel->getRelaxationVar(xp);
    if (!xr) continue;
    if (xr->getLb() > xp->getLb() + eTol) {
      VarBoundModPtr mod = (VarBoundModPtr) new VarBoundMod(xp, Lower, xr->getLb());
      mod->applyToProblem(problem_);
      p_mods.push_back(mod);
    }
    if (xr->getUb() < xp->getUb() - eTol) {
      VarBoundModPtr mod = (VarBoundModPtr) new VarBoundMod(xp, Upper, xr->getUb());
      mod->applyToProblem(problem_);
      p_mods.push_back(mod);
    }
  }
}

SolveStatus GpuBoundHandler::presolve(PreModQ *, bool *changed, Solution **)
{
  ModVector mods;
  Timer *timer = env_->getNewTimer();
  timer->start();
  invalidate();       // LinearHandler::presolve edits coefficients and rows in place between the calls (coeffImp_, dupRows_)
  const bool inf = tighten_(problem_, SolutionPoolPtr(), mods, false);     // root presolve: no incumbent yet
  if (!mods.empty()) *changed = true;
  // root mode keeps no undo information: the mods are already applied (LinearHandler deletes them too,
  // LinearHandler.cpp:1093-1099)
  for (ModVector::iterator it = mods.begin(); it != mods.end(); ++it) delete *it;
  stats_.timeHost += timer->query();
  delete timer;
  return inf ? SolvedInfeasible : Finished;
}

void GpuBoundHandler::writeStats(std::ostream &out) const
{
  out << me_ << "Statistics for GPU bound tightening:" << std::endl
      << me_ << "Calls                        = " << stats_.calls << std::endl
      << me_ << "Structure uploads            = " << stats_.uploads << std::endl
      << me_ << "Row-bound refreshes          = " << stats_.rowBoundUpdates << std::endl
      << me_ << "Constraints left to others   = " << stats_.skippedCons << std::endl
      << me_ << "Engine errors                = " << stats_.engineErrors << std::endl
      << me_ << "Answered from prefetched set = " << stats_.cacheHits << std::endl
      << me_ << "Bound modifications          = " << stats_.nMods << std::endl
      << me_ << "Times infeasibility detected = " << stats_.nInf << std::endl
      << me_ << "nnz-updates                  = " << stats_.nnzUpdates << std::endl
      << me_ << "Host time (s)                = " << stats_.timeHost << std::endl
      << me_ << "Device time (ms)             = " << stats_.timeDevice << std::endl;
}
