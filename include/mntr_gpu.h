/*
 * mntr_gpu.h -- C ABI of the B200 bound-propagation engine (libmntr_gpu.so).
 *
 * This is the drop-in boundary for Minotaur's activity-based FBBT hot path.  The
 * reference has no C ABI for it: the path lives behind the C++ plugin class
 * Minotaur::Handler (/root/reference/src/base/Handler.h:48-384).  The Minotaur-side
 * adapter (minotaur_b200/handler/GpuBoundHandler.{h,cpp}) implements that class and
 * calls the entry points below; INTEGRATION.md shows the wiring.  Each entry point
 * names the reference routine(s) whose work it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer unless the name says "_dev";
 *   - return 0 on success, <0 on error (MNTR_E_*); mntr_gpu_last_error() gives text;
 *   - a context is used by one host thread at a time (the reference creates one
 *     handler per B&B thread, examples/simple-bnb/McBnb.cpp:84-140) and owns its
 *     device, stream and workspaces; there is no global mutable state;
 *   - there is NO CPU fallback: without a usable CUDA device every call fails.
 *
 * Arithmetic: IEEE fp64.  With MNTR_ROUND_DIRECTED (default) activities and implied
 * bounds are rounded outward (__dadd_rd/_ru, __dmul_rd/_ru, __ddiv_rd/_ru) so no
 * bound is ever tighter than exact arithmetic allows; MNTR_ROUND_NEAREST reproduces
 * the reference's round-to-nearest, unfused arithmetic bit for bit where the
 * evaluation order is also the reference's (MNTR_ORDER_REFERENCE).
 */
#ifndef MNTR_GPU_H
#define MNTR_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNTR_GPU_ABI_VERSION 1
/* longest CGraph tape (nodes of one constraint) mntr_gpu_load_cgraph accepts; callers leave longer ones to NlPresHandler */
#define MNTR_GPU_MAX_TAPE 48

typedef struct mntr_gpu_ctx mntr_gpu_ctx;

/* error codes */
enum {
  MNTR_OK          = 0,
  MNTR_E_ARG       = -1,   /* bad argument */
  MNTR_E_CUDA      = -2,   /* CUDA runtime error (see mntr_gpu_last_error) */
  MNTR_E_STATE     = -3,   /* call out of order (e.g. tighten before load) */
  MNTR_E_NOMEM     = -4,
  MNTR_E_UNSUPPORTED = -5,
  MNTR_E_NCCL      = -6
};

/* variable types: Minotaur::VariableType, Types.h:83-89 */
enum { MNTR_BINARY = 0, MNTR_INTEGER = 1, MNTR_IMPLBIN = 2, MNTR_IMPLINT = 3, MNTR_CONTINUOUS = 4 };

/* per-box verdicts */
enum {
  MNTR_FEASIBLE      = 0,
  MNTR_INFEAS_BOUNDS = 1,  /* lb > ub + 1e-8 : LinearHandler::checkBounds_, LinearHandler.cpp:328-359 */
  MNTR_INFEAS_ROW    = 2,  /* activity-infeasible linear row : linBndTighten_, :994-1015.  The
                              reference's node mode drops this status (:1631); it is still a
                              valid proof of infeasibility. */
  MNTR_INFEAS_NL     = 3,  /* CGraph: NlPresHandler::chkRed_ (:139-150) or CNode::propBounds_ */
  MNTR_ERROR_NL      = 4   /* CGraph evaluation error (SolveError in the reference) */
};

/* rounding of activities / implied bounds */
enum { MNTR_ROUND_DIRECTED = 0, MNTR_ROUND_NEAREST = 1 };

/* sweep order of the linear rows */
enum {
  MNTR_ORDER_JACOBI    = 0,  /* all rows read the box of the round start; candidates merged
                                with exact max/min.  Result is
                                independent of row order and of the number of GPUs. */
  MNTR_ORDER_REFERENCE = 1   /* the reference's in-place, index-ordered sweep
                                (LinearHandler::varBndsFromCons_, :493-541) reproduced by
                                wavefront scheduling: rows that share no variable run
                                concurrently, levels run in order. */
};

/* loop control */
enum {
  MNTR_LOOP_FIXPOINT       = 0,  /* sweep until nothing changes (or max_rounds) */
  MNTR_LOOP_SIMPLEPRESOLVE = 1   /* LinearHandler::simplePresolve's truncation: <=10 rounds,
                                    rounds >=3 only while integer variables moved (:1625-1627);
                                    NlPresHandler::simplePresolve: <=2 sweeps (:1034-1035) */
};

/* which of the reference's presolve handlers a tighten call stands in for */
enum {
  MNTR_HANDLERS_ALL       = 0,  /* LinearHandler, then NlPresHandler when tapes are loaded */
  MNTR_HANDLERS_LINEAR    = 1,  /* LinearHandler only */
  MNTR_HANDLERS_NONLINEAR = 2   /* NlPresHandler only */
};

typedef struct {
  int32_t rounding;     /* MNTR_ROUND_*  */
  int32_t order;        /* MNTR_ORDER_*  (-1: Jacobi for one box, reference order for a batch) */
  int32_t loop;         /* MNTR_LOOP_*   */
  int32_t max_rounds;   /* 0 = no extra cap */
  int32_t handlers;     /* MNTR_HANDLERS_* */
  int32_t flags;        /* MNTR_FLAG_* */
  int32_t reserved[2];  /* must be 0 */
} mntr_gpu_options;

/* Single-box Jacobi with one kernel launch per phase and the host reading the change flag between rounds
 * (the kernels of the row-partitioned multi-GPU mode) instead of the single cooperative launch.  Always on
 * when a communicator is attached; this flag selects it on one GPU too. */
#define MNTR_FLAG_PER_ROUND_KERNELS 1
/* Single-box cooperative launch: always stream the rows from the CSR in global memory (staged, entry-parallel
 * batches) instead of keeping them resident in shared memory, which is the default whenever every warp of the grid
 * owns at most 32 rows.  Same results; for testing the streaming form on small instances. */
#define MNTR_FLAG_STAGED_ROWS 2

/* per-call statistics (LinPresolveStats / NlPresStats counterparts, LinearHandler.h:22-36) */
typedef struct {
  int64_t nnz_updates;   /* sum over evaluated (row, box) pairs of the row's term count */
  int64_t rows_evaluated;
  int64_t n_infeasible;  /* boxes with verdict != MNTR_FEASIBLE */
  int64_t n_changes;     /* (variable, round) pairs whose bounds moved (single-box path) */
  int32_t max_rounds;    /* largest number of rounds any box took */
  int32_t sparse_rounds; /* row-partitioned mode: rounds whose bounds were merged by the sparse exchange (changed
                            candidates all-gathered) instead of the dense MAX/MIN all-reduce */
  double  kernel_ms;     /* device time of the tighten kernels (CUDA events) */
  double  h2d_ms, d2h_ms;
  double  comm_ms;       /* device time of the per-round NCCL bound all-reduces (row-partitioned mode) */
  double  rows_ms, vars_ms; /* per-round kernels of the row-partitioned mode */
  int64_t nl_evals;      /* (constraint, box) evaluations of CGraph tapes, counted on the device: one per chkRed_
                            check and one per varBoundMods call that actually ran (boxes that stop early stop counting) */
} mntr_gpu_stats;

/* ---- lifetime -------------------------------------------------------------------- */

/* Creates a context on CUDA device `device`.  Replaces: LinearHandler / NlPresHandler
 * construction (LinearHandler.cpp:66-97, NlPresHandler.cpp:69-91). */
int mntr_gpu_create(int device, mntr_gpu_ctx **out);
void mntr_gpu_destroy(mntr_gpu_ctx *ctx);
const char *mntr_gpu_last_error(const mntr_gpu_ctx *ctx);
int mntr_gpu_abi_version(void);
/* number of CUDA devices visible, <0 if the runtime is unusable */
int mntr_gpu_device_count(void);

/* ---- problem upload --------------------------------------------------------------- */

/* Linear rows  row_lb <= A x <= row_ub  as CSR; columns strictly ascending inside a row
 * (the term order of LinearFunction, Types.h:496).  Entries with |a| <= 1e-9 are dropped
 * as LinearFunction::addTerm does (LinearFunction.cpp:20-23,89-95).  row_active may be
 * NULL (all rows active); 0 marks a deleted row (Constraint state DeletedCons).
 * Flattens: Problem / Constraint / LinearFunction / Variable::cons_ object graph. */
int mntr_gpu_load_linear(mntr_gpu_ctx *ctx, int32_t m, int32_t n, const int32_t *row_ptr,
                         const int32_t *col, const double *val, const double *row_lb,
                         const double *row_ub, const uint8_t *var_type,
                         const uint8_t *row_active);

/* CGraph constraints  c_lb <= f_c(x) + lin_c.x <= c_ub  as expression tapes.  Node order
 * inside a tape is the reference's evaluation order: variable nodes by ascending variable
 * id (vq_), constants, then operator nodes in the post-order of CGraph::finalize (dq_,
 * CGraph.cpp:557-644); the last node is the output.  op = Minotaur::OpCode (OpCode.h);
 * arg0/arg1 = local node indices (OpVar: arg0 = variable; OpSumList: child[arg0..arg1)).
 * Requires a previous mntr_gpu_load_linear (m may be 0) for n and the variable types.
 * Flattens: CGraph / CNode. */
int mntr_gpu_load_cgraph(mntr_gpu_ctx *ctx, int32_t n_cons, const int32_t *tape_ptr,
                         const uint8_t *op, const int32_t *arg0, const int32_t *arg1,
                         const double *cnst, const int32_t *child, const int32_t *lin_ptr,
                         const int32_t *lin_col, const double *lin_val, const double *c_lb,
                         const double *c_ub);

/* Constraints whose nonlinear part is a QuadraticFunction:  q_lb <= sum_k coef_k x_{v1_k} x_{v2_k} + lin.x <= q_ub.
 * Terms of a constraint in the order of the reference's VariablePairGroup (ascending (v1, v2), v1 <= v2;
 * Types.cpp CompareVariablePair); terms with |coef| < 1e-8 are dropped as QuadraticFunction::addTerm drops them.
 * NlPresHandler only CHECKS these constraints: chkRed_ (NlPresHandler.cpp:101-208) with
 * QuadraticFunction::computeBounds (QuadraticFunction.cpp:156-180) -- an infeasible one gives verdict MNTR_INFEAS_NL;
 * no bound is derived from them (the qf branch of NlPresHandler::varBndsFromCons_ is a reference defect, SURVEY.md 8a
 * N5).  Requires a previous mntr_gpu_load_linear; n_quad = 0 removes them.  Flattens: QuadraticFunction. */
int mntr_gpu_load_quad(mntr_gpu_ctx *ctx, int32_t n_quad, const int32_t *q_ptr, const int32_t *v1,
                       const int32_t *v2, const double *coef, const int32_t *lin_ptr,
                       const int32_t *lin_col, const double *lin_val, const double *q_lb,
                       const double *q_ub);

/* New bounds for the linear rows already on the device (same rows, same order as the mntr_gpu_load_linear call):
 * what a ConBoundMod or LinearHandler's row-bound tightening changes.  m doubles each way instead of re-flattening the
 * whole problem.  Replaces: Constraint::lb_/ub_ being read live by linBndTighten_ (LinearHandler.cpp:952-1045). */
int mntr_gpu_update_row_bounds(mntr_gpu_ctx *ctx, int32_t m, const double *row_lb, const double *row_ub);

/* Objective cut-off row  c.x <= rhs  (rhs = incumbent value - objective constant);
 * k = 0 removes it.  Replaces: LinearHandler::varBndsFromObj_ (LinearHandler.cpp:544-597). */
int mntr_gpu_set_cutoff(mntr_gpu_ctx *ctx, int32_t k, const int32_t *col, const double *val,
                        double rhs);

/* The incumbent's objective value as NlPresHandler::fixObjBins_ compares it: the RAW pool value -- that rule
 * does not subtract the objective constant (NlPresHandler.cpp:1030,1045-1050).  The objective's linear function
 * is the cut-off row's (col,val), so this only has an effect after mntr_gpu_set_cutoff with k > 0, in the
 * reference-order kernel with CGraph tapes loaded and the nonlinear handler enabled; +inf (the state after
 * mntr_gpu_set_cutoff) switches the rule off.  Nonlinear objectives are not supported.
 * Replaces: NlPresHandler::fixObjBins_ (NlPresHandler.cpp:1062-1121). */
int mntr_gpu_set_incumbent(mntr_gpu_ctx *ctx, double best_value);

/* ---- the hot path ------------------------------------------------------------------ */

/* Tightens n_boxes independent node boxes.  lb/ub are box-major [n_boxes][n] host arrays,
 * updated in place.  Per box: verdict[b] (MNTR_FEASIBLE ...), rounds[b] (sweeps run) and
 * nnz_updates[b]; any of the three output arrays may be NULL.  opts may be NULL (directed
 * rounding; Jacobi order for a single box, reference order for a batch; fixpoint loop).
 * Replaces: LinearHandler::simplePresolve / presolveNode (LinearHandler.cpp:1592-1653) and,
 * when tapes are loaded, NlPresHandler::simplePresolve / presolveNode (:1009-1059), i.e. one
 * PCBProcessor::presolveNode_ pass (PCBProcessor.cpp:134-175) per box. */
int mntr_gpu_tighten(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub,
                     const mntr_gpu_options *opts, int32_t *verdict, int32_t *rounds,
                     int64_t *nnz_updates);

/* Same work with boxes given as sparse branching deltas on a common root box and results
 * returned as the bound changes to emit as VarBoundMods (VarBoundMod.cpp:27-43): the form in
 * which B&B nodes actually differ (Node r_mods, NodeIncRelaxer.cpp:94-155).
 *   root_lb/root_ub [n]; box b applies deltas delta_ptr[b] .. delta_ptr[b+1):
 *   (delta_var, delta_is_upper, delta_val).
 * Outputs: mod_ptr [n_boxes+1] offsets into (mod_var, mod_is_upper, mod_val), capacity
 * mod_cap entries; *n_mods_out receives the total produced (may exceed mod_cap: call again
 * with a larger buffer).  A mod is emitted for every (var, side) whose final bound differs
 * from the box's initial bound. */
int mntr_gpu_tighten_nodes(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *root_lb,
                           const double *root_ub, const int64_t *delta_ptr,
                           const int32_t *delta_var, const uint8_t *delta_is_upper,
                           const double *delta_val, const mntr_gpu_options *opts,
                           int32_t *verdict, int32_t *rounds, int64_t *mod_ptr,
                           int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val,
                           int64_t mod_cap, int64_t *n_mods_out);

/* Device-resident variant for callers that keep boxes in HBM (bench "value" leg, and the
 * multi-GPU drivers): boxes_dev is the engine's own layout, a double2 {lb,ub} array
 * [n][ld] with ld = mntr_gpu_box_ld(n_boxes) (node-minor: the boxes of one variable are
 * contiguous).  verdict_dev/rounds_dev/nnz_dev are device arrays of n_boxes entries. */
int64_t mntr_gpu_box_ld(int32_t n_boxes);
int mntr_gpu_tighten_dev(mntr_gpu_ctx *ctx, int32_t n_boxes, void *boxes_dev,
                         const mntr_gpu_options *opts, int32_t *verdict_dev,
                         int32_t *rounds_dev, int64_t *nnz_dev);
/* box-major host arrays <-> engine layout in HBM (uses the context's stream) */
int mntr_gpu_boxes_upload(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *lb,
                          const double *ub, void *boxes_dev);
int mntr_gpu_boxes_download(mntr_gpu_ctx *ctx, int32_t n_boxes, const void *boxes_dev,
                            double *lb, double *ub);

/* Single box already resident in HBM (lb_dev/ub_dev: device arrays of n doubles, tightened in
 * place) with the Jacobi fixpoint kernel; verdict/rounds/nnz_updates are host outputs.  This is
 * the whole hot path of a single-box call minus the host<->device copies. */
int mntr_gpu_tighten_single_dev(mntr_gpu_ctx *ctx, double *lb_dev, double *ub_dev,
                                const mntr_gpu_options *opts, int32_t *verdict, int32_t *rounds,
                                int64_t *nnz_updates);
/* the context's cudaStream_t, for callers that time or order work against it */
void *mntr_gpu_stream(mntr_gpu_ctx *ctx);

/* Node boxes in the engine's layout built on the device from a root box and branching deltas (the input form of
 * mntr_gpu_tighten_nodes; host arrays in, boxes_dev [n][mntr_gpu_box_ld(n_boxes)] double2 out).  For callers
 * that keep a batch resident in HBM -- a dense box-major host copy of config C5's batch would be 65 GB.
 * The batch is PREPARED: the call also notes on the device which variables are not at a fixed point of
 * LinearHandler::tightenInts_ / checkBounds_ (LinearHandler.cpp:415-490, 328-359) -- root variables with a fractional
 * integer bound or crossed bounds, for every box, and the variables a box's deltas set -- so the next
 * mntr_gpu_tighten_dev on exactly these boxes (same pointer, same n_boxes) applies integer rounding and the bound check
 * of its first sweep only where they can do anything -- there and at the variables the rows move -- instead of at all n
 * variables of every box.  Results are the same bit for bit.  The caller must not write to the boxes between the two
 * calls; any other batch call on the context in between falls back to the full pass.
 * MNTR_GPU_NO_PREPARED=1 in the environment switches the shortcut off. */
int mntr_gpu_boxes_from_deltas(mntr_gpu_ctx *ctx, int32_t n_boxes, const double *root_lb,
                               const double *root_ub, const int64_t *delta_ptr, const int32_t *delta_var,
                               const uint8_t *delta_is_upper, const double *delta_val, void *boxes_dev);

/* Page-locked, device-mapped host memory for the caller's bound arrays.  mntr_gpu_tighten on a single box whose
 * lb / ub live in such memory takes the zero-copy path: the kernel reads the box over PCIe and writes back only
 * the bounds that moved -- no staging copies.  (Any cudaHostAlloc'ed / cudaHostRegister'ed memory qualifies; these
 * two calls spare the Minotaur-side adapter a CUDA dependency.)  Replaces: nothing in the reference (its bounds
 * live in Variable objects, Variable.h:164-191). */
void *mntr_gpu_alloc_host(mntr_gpu_ctx *ctx, int64_t bytes);
void mntr_gpu_free_host(mntr_gpu_ctx *ctx, void *p);

/* ---- root presolve: row operations of LinearHandler::presolve over the rows on the device ------------------ */

/* Duplicate-row CANDIDATES as LinearHandler::dupRows_ finds them (LinearHandler.cpp:882-949): every loaded row is
 * hashed with the two random vectors r1, r2 [n] (the caller draws them the way the reference does: rand()/RAND_MAX*10),
 * all pairs i < j are compared, and the pairs that pass the reference's tests are returned sorted by (i, j):
 * kind 1 = |h1j - h1i| < 1e-10 or |h1j + h1i| < 1e-10 (treatDupRows_ with mult 1.0), kind 2 = |h1i/h1j - h2i/h2j| <
 * 1e-10 (mult h1i / h1j).  Row indices are the caller's (the order of mntr_gpu_load_linear).  h1_out / h2_out [m]
 * receive the hashes (may be NULL).  The O(m^2) compare runs on the GPU; merging / deleting rows (treatDupRows_,
 * :1322-1395) mutates Minotaur's object graph and stays with the caller, who walks the list in order and skips pairs
 * whose row was deleted by an earlier pair -- exactly the reference's loop.  *n_pairs_out may exceed cap. */
int mntr_gpu_root_dup_rows(mntr_gpu_ctx *ctx, const double *r1, const double *r2, double *h1_out,
                           double *h2_out, int64_t cap, int32_t *pair_i, int32_t *pair_j,
                           uint8_t *pair_kind, int64_t *n_pairs_out);

/* Rows that are redundant on the box (lb, ub): the activity range of getLfBnds_ lies inside the row bounds,
 * ll >= row_lb - 1e-8 && uu <= row_ub + 1e-8 -- the test of linBndTighten_ in root mode (LinearHandler.cpp:974-985).
 * redundant [m] (caller's row order) receives 0 / 1.  Round-to-nearest, the reference's operation order: bit-exact. */
int mntr_gpu_root_redundant_rows(mntr_gpu_ctx *ctx, const double *lb, const double *ub, uint8_t *redundant,
                                 int64_t *n_redundant);

/* LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) with its implications computeImpBounds_ (:707-783), for the
 * problem given here in the caller's row order (CSR as in mntr_gpu_load_linear; no problem needs to be loaded) on the box
 * (lb, ub).  For every one-sided row with at least two terms the first binary (Binary / ImplBin, not fixed) whose
 * coefficient the reference would improve is reported: out_row, out_var, out_coef = the NEW coefficient (0: the term
 * is erased, LinearFunction::incTerm :133-142), out_side = which row bound moves with it (0 none, 1 lower, 2 upper)
 * and out_bnd its new value; out_delta (may be NULL) = the argument the reference passes to
 * LinearFunction::incTerm, so that a caller applying the change gets the very same coefficient; sorted by row.  The reference's pass is sequential through the 2-term rows its
 * implications read (a row sees the improved version of those before it): rows run in dependency levels
 * (*n_levels_out launches), so the result is the reference's, bit for bit.  Applying the changes (incTerm,
 * changeBound, bFlags) mutates Minotaur's object graph and stays with the caller.  *n_erased_out counts improvements
 * that erased a term: the set of 2-term rows changed under the pass, the caller re-runs after applying them.
 * *n_out may exceed cap. */
int mntr_gpu_root_coeff_imp(mntr_gpu_ctx *ctx, int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col,
                            const double *val, const double *row_lb, const double *row_ub, const uint8_t *var_type,
                            const double *lb, const double *ub, int64_t cap, int32_t *out_row, int32_t *out_var,
                            double *out_coef, int32_t *out_side, double *out_bnd, double *out_delta, int64_t *n_out,
                            int32_t *n_levels_out, int32_t *n_erased_out);

/* QuadHandler::simplePresolve (QuadHandler.cpp:1146-1201), the (f)-4 slice: the relations the handler holds after the
 * reformulation -- y = x^2 (QuadHandler::x2Funs_: one per x, given ascending in x) and y = x0 * x1 (x0x1Funs_: given
 * ascending in (x0, x1), x0 < x1; LinBil.cpp:26-35, 57-68) -- for the variables of the loaded problem (load_linear
 * first; m may be 0).  n_sq = n_bil = 0 removes them. */
int mntr_gpu_load_quad_relations(mntr_gpu_ctx *ctx, int32_t n_sq, const int32_t *sq_x, const int32_t *sq_y,
                                 int32_t n_bil, const int32_t *b_x0, const int32_t *b_x1, const int32_t *b_y);
/* ONE in-place sweep over the squares, then the products, on each of n_boxes boxes (lb / ub box-major [n_boxes][n],
 * updated in place): BoundsOnSquare / the square-root rule / BoundsOnProduct / BoundsOnDiv, every step through
 * updatePBounds_ (:3218-3246: integer rounding, tolerances 1e-8 / 1e-6 absolute and 1e-7 relative).  n_mods [n_boxes]
 * receives the bound changes per box.  The reference overwrites its status with Finished (:1200), so no box is ever
 * reported infeasible; n_inconsistent [n_boxes] (may be NULL) counts the steps that found crossing bounds.  With
 * MNTR_ROUND_NEAREST the result is the reference's bit for bit; MNTR_ROUND_DIRECTED rounds outward. */
int mntr_gpu_quad_simple_presolve(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub, int32_t rounding,
                                  int32_t *n_mods, int32_t *n_inconsistent);

/* The propagation loop of QuadHandler::presolveNode (QuadHandler.cpp:1204-1239) over the loaded relations, on each of
 * n_boxes boxes (lb / ub box-major [n_boxes][n], updated in place): the sweep { propSqrBnds_ (:1361-1395) over the
 * squares, propBilBnds_ (:1271-1301) over the products } repeated while it moves a bound, every step through the
 * relaxation-aware updatePBounds_ (:3248-3320: integer rounding; a side moves when it improves by more than 1e-8
 * absolute and 1e-7 relative).  verdict [n_boxes]: 1 = a step found inconsistent bounds (the reference returns
 * "infeasible" there; the bounds of such a box are not a result), else 0.  n_mods (may be NULL): Modification objects
 * the reference would push to p_mods (one per step that moves a variable, whether one side or both).  n_sweeps (may
 * be NULL): sweeps run (pStats_.iters).  max_sweeps <= 0: no cap, like the reference -- whose loop, like this one, ends
 * only because an accepted step improves a bound by more than 1e-8 + 1e-7 |bound|; a caller that cannot rule out slowly
 * contracting relations passes a cap (a box cut short is valid, just not at the fixpoint; n_sweeps tells).  Not included, because they
 * change structure or belong to another algorithm: tightenQuad_ of a handler's first node (:1241-1250) and the refresh
 * of the McCormick rows (upSqCon_ / upBilCon_, :1252-1257) -- the caller runs the latter on the boxes it keeps.
 * MNTR_ROUND_NEAREST: the reference's result bit for bit; MNTR_ROUND_DIRECTED: rounded outward. */
int mntr_gpu_quad_presolve_node(mntr_gpu_ctx *ctx, int32_t n_boxes, double *lb, double *ub, int32_t rounding,
                                int32_t max_sweeps, int32_t *verdict, int32_t *n_mods, int32_t *n_sweeps);

/* statistics of the last tighten call */
int mntr_gpu_get_stats(const mntr_gpu_ctx *ctx, mntr_gpu_stats *out);

/* ---- row-partitioned multi-GPU (one process per GPU) --------------------------------- */

/* Every rank loads ITS row block with mntr_gpu_load_linear (all n columns) and holds a
 * full replica of the box.  After this call mntr_gpu_tighten on a single box runs Jacobi
 * rounds whose candidate bounds are merged across ranks each round: rounds that move few bounds
 * exchange only the changed candidates (all-gathered, merged with exact max / min), the others take
 * a grouped NCCL all-reduce (MAX on the lower-bound candidates, whose extra slot carries the
 * row-infeasible flag, MIN on the upper-bound candidates); integer rounding and the bound check run
 * replicated after the merge,
 * so every rank ends with bit-identical boxes, independent of the number of ranks.
 * nccl_unique_id: 128 bytes from mntr_gpu_nccl_unique_id on rank 0, broadcast by the
 * caller (MPI / torch.distributed / files). */
int mntr_gpu_nccl_unique_id(void *id128);
int mntr_gpu_comm_init(mntr_gpu_ctx *ctx, int32_t n_ranks, int32_t rank, const void *id128);
int mntr_gpu_comm_destroy(mntr_gpu_ctx *ctx);

/* ---- node batches over several GPUs of one box, from ONE process --------------------------------------- */

/* A group holds one context per listed device; the problem is replicated on every member (the CSR / tapes are
 * small next to the boxes), and mntr_gpu_group_tighten_nodes splits a node batch contiguously over the members,
 * one host thread per device, no collective (north star: "independent node-box batches split across GPUs";
 * "multi-GPU init taking a device list").  The single-process counterpart of the
 * one-process-per-GPU split bench.py makes under torchrun; a B&B thread that owns a group hands all strong-
 * branching candidates of a node to all GPUs at once.  Signatures mirror the single-context calls. */
typedef struct mntr_gpu_group mntr_gpu_group;
int mntr_gpu_group_create(int32_t n_devices, const int32_t *devices, mntr_gpu_group **out);
void mntr_gpu_group_destroy(mntr_gpu_group *g);
int32_t mntr_gpu_group_size(const mntr_gpu_group *g);
mntr_gpu_ctx *mntr_gpu_group_member(mntr_gpu_group *g, int32_t i);
const char *mntr_gpu_group_last_error(const mntr_gpu_group *g);
int mntr_gpu_group_load_linear(mntr_gpu_group *g, int32_t m, int32_t n, const int32_t *row_ptr,
                               const int32_t *col, const double *val, const double *row_lb,
                               const double *row_ub, const uint8_t *var_type, const uint8_t *row_active);
int mntr_gpu_group_load_cgraph(mntr_gpu_group *g, int32_t n_cons, const int32_t *tape_ptr, const uint8_t *op,
                               const int32_t *arg0, const int32_t *arg1, const double *cnst,
                               const int32_t *child, const int32_t *lin_ptr, const int32_t *lin_col,
                               const double *lin_val, const double *c_lb, const double *c_ub);
int mntr_gpu_group_load_quad(mntr_gpu_group *g, int32_t n_quad, const int32_t *q_ptr, const int32_t *v1,
                             const int32_t *v2, const double *coef, const int32_t *lin_ptr,
                             const int32_t *lin_col, const double *lin_val, const double *q_lb,
                             const double *q_ub);
int mntr_gpu_group_set_cutoff(mntr_gpu_group *g, int32_t k, const int32_t *col, const double *val, double rhs);
int mntr_gpu_group_set_incumbent(mntr_gpu_group *g, double best_value);
int mntr_gpu_group_tighten_nodes(mntr_gpu_group *g, int32_t n_boxes, const double *root_lb,
                                 const double *root_ub, const int64_t *delta_ptr,
                                 const int32_t *delta_var, const uint8_t *delta_is_upper,
                                 const double *delta_val, const mntr_gpu_options *opts,
                                 int32_t *verdict, int32_t *rounds, int64_t *mod_ptr,
                                 int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val,
                                 int64_t mod_cap, int64_t *n_mods_out);

#ifdef __cplusplus
}
#endif
#endif /* MNTR_GPU_H */
