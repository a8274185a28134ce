// device_problem.cuh -- flattened problem data as the kernels see it (all device pointers).
#pragma once
#include "common.cuh"

namespace mntr {

// Linear rows, stored in (wavefront level, original index) order with deleted rows last.
// CSR whose rows are padded to a multiple of FOUR entries so a lane can fetch four (col,val) pairs
// with one 128-bit column load and two 128-bit value loads; padding entries have val == 0 and
// a valid column, and are skipped by value.  CSC (var -> rows) replaces Variable::cons_
// (Variable.h:164-191) for the bFlag propagation of LinearHandler::changeBFlag_ (:1229-1234).
struct LinDev {
  int32_t m, n;
  const int2    *row_info;  // [m] {first entry (multiple of 4), true term count; count < 0 marks a deleted row}
  const double2 *row_bnd;   // [m] {row lb, row ub}
  const int32_t *col;       // [nnz_padded]  a row occupies entries [beg, beg + roundup4(count))
  const double  *val;       // [nnz_padded]
  const uint8_t *var_type;  // [n]
  const int32_t *csc_ptr;   // [n+1]
  const int32_t *csc_row;   // [nnz]
  // wavefront schedule of the reference's index-ordered in-place sweep
  int32_t n_levels;
  const int32_t *level_ptr; // [n_levels+1] ranges of STORED rows (rows are stored in level order)
};

// rows are padded to a multiple of kRowPad entries (padding: val == 0, a valid column)
constexpr int kRowPad = 4;
__host__ __device__ __forceinline__ int row_end(int2 info) { return info.x + ((info.y + kRowPad - 1) & ~(kRowPad - 1)); }

// CGraph tapes (see include/mntr_gpu.h for the node order contract)
struct NlDev {
  int32_t n_cons;
  const int32_t *tape_ptr;  // [n_cons+1]
  const uint8_t *op;
  const int32_t *arg0, *arg1;
  const double  *cnst;
  const int32_t *child;
  const int32_t *lin_ptr, *lin_col;
  const double  *lin_val;
  const double  *c_lb, *c_ub;
  int32_t max_nodes;        // longest tape
  int32_t n_levels;
  const int32_t *level_ptr; // [n_levels+1] ranges of STORED constraints (stored in level order)
};

// workspace of the single-box Jacobi fixpoint kernel
struct SingleWs {
  double2 *box;    // [n] {lb, ub} of the round start
  double2 *nbox;   // [n] candidates of the round (atomic max / min)
  uint32_t *bits;  // [(m+31)/32] row-is-on-the-next-work-list bit set   (Constraint bFlag)
  int32_t *list;   // [m] work list of flagged rows
  // control block (128 bytes, zeroed before every launch)
  int32_t *ring;   // [12] per-round words: ring[r%3] changed, ring[3+r%3] int moved, ring[6+r%3] list length
  int32_t *status; // [4]  [0] verdict, [1] rounds, [2] changed (variable, round) pairs
  unsigned long long *counters;  // [2] [0] nnz_updates, [1] rows evaluated
  unsigned *bar;   // device-wide barrier arrive counter
  unsigned long long *trace;     // [64] optional phase timestamps (globaltimer ns), or nullptr
};

// workspace of the per-round kernels (row-partitioned multi-GPU mode)
struct RoundsWs {
  double2 *box;    // [n] {lb, ub} of the round start (replicated on every rank)
  double *nlb;     // [n+1] lower-bound candidates (all-reduced with MAX); [n] = row-infeasible flag
  double *nub;     // [n]   upper-bound candidates (all-reduced with MIN)
  uint32_t *bits;  // [(m+31)/32] this rank's rows on the next work list
  int32_t *list;   // [m] work list
  int32_t *ctrl;   // [8] [0] changed [1] int moved [2] next list length [3] verdict [4] changed pairs
  unsigned long long *counters;  // [2] nnz_updates, rows evaluated (this rank)
};

}  // namespace mntr
