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
struct mntr_gpu_group;

namespace Minotaur {

/// Counters, in the spirit of LinPresolveStats (LinearHandler.h:22-36).
struct GpuBoundStats {
  int calls;          ///< presolveNode / simplePresolve calls
  int uploads;        ///< times the problem structure was (re)flattened and uploaded
  int nMods;          ///< VarBoundMods emitted
  int nInf;           ///< calls that proved infeasibility
  int rowBoundUpdates; ///< times only the row bounds were refreshed on the device (structure unchanged)
  int skippedCons;    ///< constraints of the last upload the engine does not take (left to NlPresHandler / QuadHandler)
  int engineErrors;   ///< engine calls that failed (the call then reports "no tightening", it never aborts the solve)
  int cacheHits;      ///< presolveNode calls answered from prefetched candidate outcomes (no device call)
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
  /// Several GPUs of the box: node boxes (presolveNode, simplePresolve) run on devices[0]; tightenCandidates splits
  /// its batch over all of them (mntr_gpu_group_*: one process, one host thread per device, no collective).
  GpuBoundHandler(EnvPtr env, ProblemPtr problem, const std::vector<int> &devices);
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

  /**
   * The same, for callers that cannot consume a batch at once: StrongBrancher / WeakBrancher tighten one candidate
   * box after the other through Handler::getStrongerMods -> presolveNode (StrongBrancher.cpp:499-585,
   * WeakBrancher.cpp:300-350).  Called once before such a loop with the branching modifications of ALL candidates
   * (Handler::getBrMod for DownBranch and UpBranch), this tightens all their boxes in one device call and keeps the
   * outcomes; every presolveNode that follows on `rel` with exactly one of those modifications applied is answered
   * from them (same VarBoundMods, same verdict, no device call).  The outcomes are dropped by clearCandidates(), by the
   * next prefetch, and whenever a presolveNode finds the relaxation in a state it does not know.
   */
  void prefetchCandidates(RelaxationPtr rel, SolutionPoolPtr spool, const std::vector<std::vector<BoundChange> > &deltas);
  void clearCandidates();

  /**
   * LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) for problem_, the detection on the device
   * (mntr_gpu_root_coeff_imp: the reference's sequential semantics through dependency levels), the changes applied
   * here exactly as the reference applies them: LinearFunction::incTerm with the same argument,
   * Problem::changeBound on the row, Constraint::setBFlag(true).  Returns the number of improved rows; *changed is
   * set when there is one.  Meant for a root presolve loop that runs this step on the GPU handler instead of in
   * LinearHandler (setPreOptCoeffImp(false) there).
   */
  int coeffImprove(bool *changed);

  std::string getName() const;
  void writeStats(std::ostream &out) const;

  // ---- options ----
  void setMode(Mode m) { mode_ = m; }
  /// Reproduce the reference's round-to-nearest arithmetic instead of outward rounding.
  void setRoundNearest(bool b) { roundNearest_ = b; }
  /**
   * Before every call the handler compares a signature of the problem's rows (state, bounds, number of terms of every
   * constraint: one pass over the constraints, no pass over the terms) with the one of the device copy: changed row
   * BOUNDS alone are refreshed (mntr_gpu_update_row_bounds), anything else re-flattens the problem.  Coefficients
   * edited in place with the term count unchanged (LinearHandler::coeffImp_) are not seen by the signature: root
   * presolve() therefore always re-flattens, and callers that edit coefficients between node calls call invalidate().
   * A caller that guarantees a static structure may switch the check off.
   */
  void setStructureCheck(bool b) { checkStructure_ = b; }
  /// Forget the device copy: the next call re-flattens and uploads the problem.
  void invalidate() { loadedFor_ = 0; }
  const GpuBoundStats *getStats() const { return &stats_; }
  /// False when no CUDA device / library is usable; every call then reports "no tightening" is NOT done:
  /// the handler throws at construction instead (there is no CPU fallback).
  bool isReady() const { return ctx_ != 0; }

private:
  EnvPtr env_;
  ProblemPtr problem_;
  LoggerPtr logger_;
  mntr_gpu_ctx *ctx_;
  mntr_gpu_group *group_;   // non-null with several devices; ctx_ is then its first member
  Mode mode_;
  bool roundNearest_;
  bool checkStructure_;
  GpuBoundStats stats_;
  static const std::string me_;

  // identity of the structure currently on the device
  const Problem *loadedFor_;
  UInt loadedVars_, loadedCons_;
  bool cutoffOn_;      // an objective cut-off row is installed on the device
  unsigned long long sigStruct_, sigBounds_;   // signature of the uploaded rows: structure / row bounds
  // the box that travels: page-locked, device-mapped host memory (mntr_gpu_alloc_host), so that a single-box call is
  // ONE kernel launch that reads the bounds over PCIe and writes back only those that moved
  double *lb_, *ub_;
  UInt boxCap_;
  std::vector<double> lb0_, ub0_;
  // prefetched candidate outcomes (prefetchCandidates): the box they were computed from and, per candidate, its deltas
  const Problem *cacheFor_;
  std::vector<double> cacheLb_, cacheUb_;
  std::vector<std::vector<BoundChange> > cacheDeltas_;
  std::vector<BoxOutcome> cacheOut_;
  /// the candidate whose deltas are exactly the difference between p's bounds and the cached base box, or -1
  int findCached_(ProblemPtr p) const;

  /// Flatten p (linear rows -> CSR, CGraph constraints -> tapes) and upload it.
  void upload_(ProblemPtr p);
  /// Make the device copy current for p (signature check, see setStructureCheck).
  void sync_(ProblemPtr p);
  void signature_(ProblemPtr p, unsigned long long &structure, unsigned long long &bounds) const;
  void init_(const std::vector<int> &devices);
  void engineFailed_(const char *where);
  /// One tighten call on the current bounds of p; emits applied VarBoundMods.  Returns infeasible?
  bool tighten_(ProblemPtr p, SolutionPoolPtr spool, ModVector &mods, bool truncated);
  void setCutoff_(ProblemPtr p, SolutionPoolPtr spool);
  void copyBndsFromRel_(RelaxationPtr rel, ModVector &p_mods);
};

typedef GpuBoundHandler *GpuBoundHandlerPtr;
}  // namespace Minotaur

#endif
