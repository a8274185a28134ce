// kernels.h -- host-callable launchers of the CUDA kernels (internal to libmntr_gpu.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_problem.cuh"

namespace mntr {

// K1 (linear_single.cu): single-box Jacobi fixpoint, one cooperative launch.
// lb_dev/ub_dev [n] are read at the start and overwritten with the tightened box.  staged_only: never take the
// resident-rows form (rows in shared memory, row_resident.cuh), even when the instance is small enough.
cudaError_t launch_single_jacobi(const LinDev &P, const SingleWs &W, double *lb_dev, double *ub_dev,
                                 bool directed, int max_rounds, int loop_mode, int sm_count, bool staged_only,
                                 cudaStream_t stream);

// K5 (linear_rounds.cu): one Jacobi round as separate launches (row-partitioned multi-GPU mode)
cudaError_t launch_rounds_init(const LinDev &P, const RoundsWs &W, const double *lb_dev, const double *ub_dev,
                               int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_rows(const LinDev &P, const RoundsWs &W, int lanes_per_row, bool directed,
                               int first, int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_vars(const LinDev &P, const RoundsWs &W, int sm_count, cudaStream_t stream);
// sparse bound exchange: compact this rank's changed candidates into W.xsend; after the all-gather into W.xrecv,
// merge the other ranks' candidates (or raise ctrl[5] if any rank's message overflowed)
// cap: the capacity tier of this round's messages (<= W.xcap, the same on every rank)
cudaError_t launch_rounds_compact(const LinDev &P, const RoundsWs &W, int cap, int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_apply(const LinDev &P, const RoundsWs &W, int rank, int cap, int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_finish(const LinDev &P, const RoundsWs &W, double *lb_dev, double *ub_dev, int sm_count,
                                 cudaStream_t stream);
// list-based phases: work proportional to the variables that received a candidate in the round
cudaError_t launch_rounds_vars_list(const LinDev &P, const RoundsWs &W, int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_finalize(const LinDev &P, const RoundsWs &W, int max_rounds, int loop_mode, cudaStream_t stream);
cudaError_t launch_rounds_clear_list(const RoundsWs &W, int sm_count, cudaStream_t stream);
// bound exchange over NVLink peer memory: push this rank's touched candidates into every peer's inbox and raise the
// round tag there; wait for the peers' tags and merge their candidates
cudaError_t launch_rounds_push(const LinDev &P, const RoundsWs &W, unsigned tag, int sm_count, cudaStream_t stream);
cudaError_t launch_rounds_pull(const LinDev &P, const RoundsWs &W, unsigned tag, int sm_count, cudaStream_t stream);

// per-box outputs / controls of the batched kernels
struct BatchIo {
  double2 *boxes;        // [n][ld] {lb,ub}, node-minor
  int64_t ld;            // boxes per variable row (multiple of 32)
  int32_t n_boxes;
  uint32_t *rowflag;     // [tiles][m] bit b = row flagged for box tile*32+b   (bFlag)
  uint32_t *varflag;     // [tiles][n] bit b = variable moved in the current sweep of box tile*32+b
  int32_t *verdict;      // [n_boxes]
  int32_t *rounds;       // [n_boxes]
  long long *nnz;        // [n_boxes]
  unsigned char *tstate; // [tiles][kTileStateBytes] per-tile control words (global: shared by a cluster)
  unsigned long long *nl_evals;  // [1] (constraint, box) evaluations of CGraph tapes (chkRed_ and varBoundMods each count one)
  // PREPARED batch (built by boxes_from_root + apply_deltas in this context, untouched since): non-null and *prepared == 0
  // say that every variable whose bounds are not a fixed point of tightenInts_ / checkBounds_ in some box -- a root
  // variable with a fractional integer bound or crossed bounds, a variable one of the box's deltas set -- has its bit up
  // in varflag already.  The first sweep's integer rounding / bound check then visits flagged variables only instead of
  // all n (LinearHandler.cpp:415-490, 328-359 are no-ops elsewhere).
  const int32_t *prepared = nullptr;
};
constexpr int kTileStateBytes = 640;

// K3 (linear_batch.cu): batched node boxes, the reference's in-place index-ordered sweep
// reproduced by wavefront levels; one CTA per tile of 32 boxes.
cudaError_t launch_batch_reference(const LinDev &P, const NlDev *N, const BatchIo &io, bool directed,
                                   int loop_mode, int max_rounds, int lin_enabled, int nl_enabled,
                                   int sm_count, cudaStream_t stream);

// box-major [n_boxes][n] lb/ub (device staging) <-> node-minor double2 boxes
cudaError_t launch_boxes_pack(const double *lb_bm, const double *ub_bm, int32_t n, int32_t box0,
                              int32_t nb, double2 *boxes, int64_t ld, cudaStream_t stream);
cudaError_t launch_boxes_unpack(const double2 *boxes, int64_t ld, int32_t n, int32_t box0, int32_t nb,
                                double *lb_bm, double *ub_bm, cudaStream_t stream);
// fill padding boxes [n_boxes, ld) with a harmless copy of box 0
cudaError_t launch_boxes_pad(double2 *boxes, int64_t ld, int32_t n, int32_t n_boxes, cudaStream_t stream);

// node form: boxes from root + deltas, and extraction of the resulting mods
// (varflag [tiles][n], zeroed by the caller, for BatchIo::prepared: root variables that need rounding / the bound check are
// flagged for every box, then the deltas' variables for their boxes)
cudaError_t launch_boxes_from_root(const double *root_lb, const double *root_ub, int32_t n,
                                   int32_t n_boxes, double2 *boxes, int64_t ld, const uint8_t *var_type, uint32_t *varflag,
                                   cudaStream_t stream);
cudaError_t launch_apply_deltas(const long long *delta_ptr, const int32_t *delta_var,
                                const uint8_t *delta_is_upper, const double *delta_val,
                                int32_t n_boxes, double2 *boxes, int64_t ld, uint32_t *varflag, int32_t n, cudaStream_t stream);
// mods = final bounds that differ from the box's initial bounds (root + the box's deltas); no copy of the initial boxes.
// count: strip_cnt [ceil(n/256)][ld] receives, per box, the EXCLUSIVE offsets of the 256-variable strips, mod_count the
// box's ordered mods, extra_count the unordered extras (deltas that loosen the root).  emit: ascending (variable, side)
// at mod_ptr[b] + strip offset; the extras at extra_at[b] ...
cudaError_t launch_count_mods(const double2 *boxes, const double *root_lb, const double *root_ub, const long long *delta_ptr,
                              const int32_t *delta_var, const uint8_t *delta_is_upper, const double *delta_val, int64_t ld,
                              int32_t n, int32_t n_boxes, int32_t *strip_cnt, long long *mod_count, long long *extra_count,
                              cudaStream_t stream);
cudaError_t launch_emit_mods(const double2 *boxes, const double *root_lb, const double *root_ub, const long long *delta_ptr,
                             const int32_t *delta_var, const uint8_t *delta_is_upper, const double *delta_val, int64_t ld,
                             int32_t n, int32_t n_boxes, const int32_t *strip_off, const long long *mod_ptr,
                             const long long *extra_at, long long cap, int32_t *mod_var, uint8_t *mod_is_upper, double *mod_val,
                             cudaStream_t stream);

// root-presolve row operations (root_rows.cu): duplicate-row candidates and redundant rows of LinearHandler::presolve
cudaError_t launch_row_hash(const LinDev &P, const int32_t *perm, const double *r1, const double *r2, double *h1, double *h2,
                            cudaStream_t stream);
cudaError_t launch_dup_pairs(int m, const double *h1, const double *h2, long long cap, int32_t *pair_i, int32_t *pair_j,
                             uint8_t *pair_kind, unsigned long long *count, cudaStream_t stream);
cudaError_t launch_redundant_rows(const LinDev &P, const int32_t *perm, const double *lb, const double *ub, uint8_t *flag,
                                  unsigned long long *count, cudaStream_t stream);


// LinearHandler::coeffImp_ (LinearHandler.cpp:600-704) for the rows of one dependency level (root_rows.cu).  Rows are in
// the CALLER's order (the pass is sequential in row index: a row's implications read 2-term rows, and see the improved
// version of those with a smaller index, the original of the others).
struct CoeffProb {
  int32_t m, n;
  const int32_t *row_ptr, *col;
  const double *val0, *rlb0, *rub0;     // the problem as it came
  double *val, *rlb, *rub;              // ... as the pass has changed it so far
  const uint8_t *var_type;
  const double *lb, *ub;
  const int32_t *cptr, *crow;           // rows of a variable
  const uint8_t *is2;                   // the row has exactly two terms
};
cudaError_t launch_coeff_imp(const CoeffProb &Q, const int32_t *rows, int32_t n_rows, long long cap, int32_t *out_row,
                             int32_t *out_var, double *out_coef, int32_t *out_side, double *out_bnd, double *out_delta,
                             unsigned long long *count, int32_t *n_erased, cudaStream_t stream);


// QuadHandler::simplePresolve (quad_relations.cu): the relations y = x^2 (b < 0) / y = a * b, stored in wavefront-level
// order of the handler's sweep (squares by x, then products by (x0, x1))
struct QRelDev {
  int32_t n_rel;
  const int32_t *a, *b, *y;
  int32_t n_levels;
  const int32_t *level_ptr;
  const uint8_t *var_type;
};
cudaError_t launch_quad_relations(const QRelDev &Q, double2 *boxes, int64_t ld, int32_t n_boxes, bool directed, int32_t *n_mods,
                                  int32_t *n_bad, cudaStream_t stream);
// the propagation loop of QuadHandler::presolveNode over the same relations: verdict 1 = infeasible, Modification count, sweeps
cudaError_t launch_quad_node(const QRelDev &Q, double2 *boxes, int64_t ld, int32_t n_boxes, bool directed, int32_t max_sweeps,
                             int32_t *verdict, int32_t *n_mods, int32_t *n_sweeps, cudaStream_t stream);

}  // namespace mntr
