//
//     GpuBoundHandler -- B200 bound-tightening handler for Minotaur
//
/**
 * \file GpuBoundHandler.h
 * \brief A Minotaur::Handler that runs activity-based FBBT (what LinearHandler::simplePresolve /
 * presolveNode and NlPresHandler::simplePresolve / presolveNode do) on a B200 through the C ABI of
 * libmntr_gpu.so (include/mntr_gpu.h).
 *
 * It implements the reference's plugin interface, src/base/Handler.h:48-384: all twelve pure virtuals,
 * of which only presolve(), presolveNode() and getName() do work -- exactly the shape of NlPresHandler
 * (src/base/NlPresHandler.h:74-112) -- plus the non-virtual simplePresolve() that callers such as
 * MINLPDiving invoke on a LinearHandler directly (MINLPDiving.cpp:203).  It is meant to live in the Minotaur
 * tree as src/gpu/GpuBoundHandler.{h,cpp}; INTEGRATION.md shows the wiring in Bnb.cpp / QG.cpp.
 *
 * Data flow of presolveNode(rel, ...): the relaxation is flattened once (and again whenever its size
 * changes) into CSR rows + CGraph tapes and uploaded; per call only the current variable bounds go to the
 * device and the tightened box comes back; every changed (variable, side) becomes one VarBoundMod, already
 * applied to rel, appended to r_mods (ownership passes to the Node, PCBProcessor.cpp:153-160).
 */
#ifndef MINOTAURGPUBOUNDHANDLER_H
#define MINOTAURGPUBOUNDHANDLER_H

#include <vector>

#include "Handler.h"

struct mntr_gpu_ctx;

namespace Minotaur {

/// Counters, in the spirit of LinPresolveStats (LinearHandler.h:22-36).
struct GpuBoundStats {
  int calls;          ///< presolveNode / simplePresolve calls
  int uploads;        ///< times the problem structure was (re)flattened and uploaded
  int nMods;          ///< VarBoundMods emitted
  int nInf;           ///< calls that proved infeasibility
  long long nnzUpdates;
  double timeHost;    ///< host seconds in flatten + gather + mod emission
  double timeDevice;  ///< device milliseconds reported by the engine
};

class GpuBoundHandler : public Handler {
public:
  /// Which sweep the device runs for a single node box.
  enum Mode {
    FastFixpoint,   ///< Jacobi rounds to the fixpoint, directed rounding (default)
    ReferenceOrder  ///< the reference's in-place index-ordered sweep with simplePresolve's loop truncation
  };

  GpuBoundHandler(EnvPtr env, ProblemPtr problem, int device = 0);
  ~GpuBoundHandler();

  // ---- Handler interface: does nothing, like NlPresHandler ----
  void relaxInitFull(RelaxationPtr, bool *) {}
  void relaxInitInc(RelaxationPtr, bool *) {}
  void relaxNodeFull(NodePtr, RelaxationPtr, bool *) {}
  void relaxNodeInc(NodePtr, RelaxationPtr, bool *) {}
  bool isFeasible(ConstSolutionPtr, RelaxationPtr, bool &, double &) { return true; }
  void separate(ConstSolutionPtr, NodePtr, RelaxationPtr, CutManager *, SolutionPoolPtr, ModVector &,
                ModVector &, bool *, SeparationStatus *) {}
  void getBranchingCandidates(RelaxationPtr, const DoubleVector &, ModVector &, BrVarCandSet &,
                              BrCandVector &, bool &) {}
  ModificationPtr getBrMod(BrCandPtr, DoubleVector &, RelaxationPtr, BranchDirection)
  { return ModificationPtr(); }
  Branches getBranches(BrCandPtr, DoubleVector &, RelaxationPtr, SolutionPoolPtr) { return Branches(); }

  // ---- Handler interface: the hot path ----
  /// Root: tighten the bounds of problem_ in place (bounds only; structure changes stay with LinearHandler).
  SolveStatus presolve(PreModQ *pre_mods, bool *changed, Solution **sol);

  /// Node: returns true iff the node is proven infeasible (Handler.h:229-231).
  bool presolveNode(RelaxationPtr rel, NodePtr node, SolutionPoolPtr s_pool, ModVector &p_mods,
                    ModVector &r_mods);

  /// Same entry LinearHandler offers to heuristics (not virtual in Handler, Handler.h:356-361).
  void simplePresolve(ProblemPtr p, SolutionPoolPtr spool, ModVector &t_mods, SolveStatus &status);

  // ---- batch form: all candidate boxes of a node in ONE device call ----
  /// A bound change that distinguishes a candidate box from the current box of the relaxation.
  struct BoundChange { UInt var; BoundType lu; double val; };
  /// What FBBT makes of one candidate box: infeasible, or the bound changes it derives (NOT applied anywhere).
  struct BoxOutcome { bool infeasible; std::vector<BoundChange> changes; };
  /**
   * Strong(er)-branching producers (StrongBrancher::strongBranch_, StrongBrancher.cpp:505-585;
   * WeakBrancher::getStrongerMods loops, WeakBrancher.cpp:317-341) tighten 2 boxes per candidate one after the
   * other through Handler::getStrongerMods.  This call takes all of them at once: box b = the current bounds of
   * `rel` with deltas[b] applied (mntr_gpu_tighten_nodes: branching deltas in, VarBoundMod tuples out).  `rel` is
   * not modified.  out[b] is what presolveNode would have produced on box b.
   */
  void tightenCandidates(RelaxationPtr rel, SolutionPoolPtr spool,
                         const std::vector<std::vector<BoundChange> > &deltas, std::vector<BoxOutcome> &out);

  std::string getName() const;
  void writeStats(std::ostream &out) const;

  // ---- options ----
  void setMode(Mode m) { mode_ = m; }
  /// Reproduce the reference's round-to-nearest arithmetic instead of outward rounding.
  void setRoundNearest(bool b) { roundNearest_ = b; }
  const GpuBoundStats *getStats() const { return &stats_; }
  /// False when no CUDA device / library is usable; every call then reports "no tightening" is NOT done:
  /// the handler throws at construction instead (there is no CPU fallback).
  bool isReady() const { return ctx_ != 0; }

private:
  EnvPtr env_;
  ProblemPtr problem_;
  LoggerPtr logger_;
  mntr_gpu_ctx *ctx_;
  Mode mode_;
  bool roundNearest_;
  GpuBoundStats stats_;
  static const std::string me_;

  // identity of the structure currently on the device
  const Problem *loadedFor_;
  UInt loadedVars_, loadedCons_;
  bool cutoffOn_;      // an objective cut-off row is installed on the device
  std::vector<double> lb_, ub_, lb0_, ub0_;

  /// Flatten p (linear rows -> CSR, CGraph constraints -> tapes) and upload it.
  void upload_(ProblemPtr p);
  /// One tighten call on the current bounds of p; emits applied VarBoundMods.  Returns infeasible?
  bool tighten_(ProblemPtr p, SolutionPoolPtr spool, ModVector &mods, bool truncated);
  void setCutoff_(ProblemPtr p, SolutionPoolPtr spool);
  void copyBndsFromRel_(RelaxationPtr rel, ModVector &p_mods);
};

typedef GpuBoundHandler *GpuBoundHandlerPtr;
}  // namespace Minotaur

#endif
