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
  const int32_t *colx;      // [nnz_padded]  col with bit 31 set when the variable is integer (single-box fixpoint kernel)
  const double  *val;       // [nnz_padded]
  const uint8_t *var_type;  // [n]
  const int32_t *csc_ptr;   // [n+1]
  const int32_t *csc_row;   // [nnz]
  int32_t csc_nnz;          // length of csc_row
  int32_t nnz_pad;          // length of col / colx / val (rows padded to kRowPad entries)
  // wavefront schedule of the reference's index-ordered in-place sweep
  int32_t n_levels;
  const int32_t *level_ptr; // [n_levels+1] ranges of STORED rows (rows are stored in level order)
  // objective cut-off row  c.x <= rhs  (LinearHandler::varBndsFromObj_, LinearHandler.cpp:544-597); cut_cnt 0 = none
  int32_t cut_cnt;
  const int32_t *cut_col;   // [roundup4(cut_cnt)]
  const int32_t *cut_colx;  // same with the integer bit
  const double  *cut_val;
  const double2 *cut_bnd;   // [1] {-inf, rhs}
  double cut_rhs;
  double obj_ub;            // raw incumbent value for NlPresHandler::fixObjBins_ (+inf: none)
};

// rows are padded to a multiple of kRowPad entries (padding: val == 0, a valid column)
constexpr int kRowPad = 4;
__host__ __device__ __forceinline__ int row_end(int2 info) { return info.x + ((info.y + kRowPad - 1) & ~(kRowPad - 1)); }

// CGraph tapes (see include/mntr_gpu.h for the node order contract)
struct NlDev {
  int32_t n_cons;
  int32_t all_shaped;       // every tape has one of the two straight-line shapes of cgraph.cuh (classified at load time)
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
  // constraints with a QuadraticFunction  q_lb <= sum_k coef_k x_{v1_k} x_{v2_k} + lin.x <= q_ub: only checked
  // (NlPresHandler::chkRed_ with QuadraticFunction::computeBounds, QuadraticFunction.cpp:156-180)
  int32_t n_quad;
  const int32_t *q_ptr;     // [n_quad+1]
  const int32_t *q_v1, *q_v2;
  const double  *q_coef;
  const int32_t *q_lin_ptr, *q_lin_col;
  const double  *q_lin_val;
  const double  *q_lb, *q_ub;
};

// workspace of the single-box Jacobi fixpoint kernel (linear_single.cu).  Everything is double-buffered by round
// parity: round r reads box[r&1] / due[r&1] and writes box[(r+1)&1] / due[(r+1)&1] / touched[r&1].
struct SingleWs {
  double2 *box[2];      // [n] {lb, ub}: merged candidates, integer bounds not yet rounded
  uint32_t *due[2];     // [(m+31)/32] row-is-due bit sets   (Constraint bFlag)
  uint32_t *touched[2]; // [(n+31)/32] variable-moved-in-the-round bit sets
  uint32_t *ever;       // [(n+31)/32] variable moved in some round: the only entries the epilogue writes back
  // control block (128 bytes; zero when a launch starts: the previous launch's last block resets it)
  unsigned *sync;  // [4]  ONE 16-byte line, read with one load by the barrier's poller.  Every word is ROUND-TAGGED and
                   //      only ever grows (atomicMax / atomicOr), so a block that has already left barrier r and writes
                   //      round r+1 state cannot change what a slower block, still taking its snapshot of barrier r,
                   //      concludes about round r (a later tag means "round r went on", which implies the same decision):
                   //      [0] device-wide barrier arrive counter
                   //      [1] 2 x (last round in which a bound moved) + (a row moved an integer variable in it,
                   //          nintmods > 0, :1625-1627)
                   //      [2] bits 0..28: last round whose fix-up found crossed bounds; kCtl* bits above
                   //      [3] last round in which a row was activity-infeasible
  int32_t *status; // [8]  results: [1] rounds  [2] changed (variable, round) pairs  [6] verdict of the loop
  unsigned long long *counters;  // [2] [0] nnz_updates, [1] rows evaluated
  unsigned *done;  // blocks that have left the kernel: the last one publishes and resets the control block
  int32_t *result; // [kCtrlWords] pinned, mapped host copy of the control block, written by the last block
  unsigned long long *trace;     // [64] optional phase timestamps (globaltimer ns), or nullptr
  unsigned stress_ns;            // test hook: block 1's barrier poller sleeps this long after arriving (0: off)
};
constexpr int kCtrlWords = 32;   // the control block: 128 bytes
// flag bits of sync[2] (above the round tag of the bound check)
constexpr unsigned kCtlRoundMask = (1u << 29) - 1u;
constexpr unsigned kCtlFinalCross = 1u << 29;  // the epilogue's bound check of the last round's moved variables failed
constexpr unsigned kCtlRowCross = 1u << 30;    // a row's bounds cross        (raised in round 1 only)
constexpr unsigned kCtlInCross = 1u << 31;     // the incoming bounds cross   (raised before round 1 only)

// one changed candidate of the sparse exchange; the header of a rank's message reuses the layout:
// j = number of changed candidates (may exceed the capacity: overflow), lb = the rank's row-infeasible flag
struct __align__(8) BoundMsg { double lb, ub; long long j; };

// workspace of the per-round kernels (row-partitioned multi-GPU mode)
struct RoundsWs {
  double2 *box;    // [n] {lb, ub} of the round start (replicated on every rank)
  double *nlb;     // [n+1] lower-bound candidates (all-reduced with MAX); [n] = row-infeasible flag
  double *nub;     // [n]   upper-bound candidates (all-reduced with MIN)
  uint32_t *bits;  // [(m+31)/32] this rank's due rows   (Constraint bFlag)
  int32_t *ctrl;   // [16] see kRc*
  unsigned long long *counters;  // [2] nnz_updates, rows evaluated (this rank)
  // Per-round work proportional to the CHANGES: a variable that receives a candidate (from this rank's rows or, after
  // the exchange, from another rank's) goes on the round's touched list -- once: tbits says who is on it -- and the
  // merge, the integer rounding, the bound check and the row flagging walk that list instead of all n variables.
  uint32_t *tbits; // [(n+31)/32]
  int32_t *tlist;  // [n]   (ctrl[kRcTouched] entries)
  uint32_t *ebits; // [(n+31)/32] variable moved in some round of this call: the only ones the finish kernel writes back
  int32_t *elist;  // [n]   (ctrl[kRcEver] entries)
  int32_t *progress;  // [4] pinned, mapped host words {round finished, stop, verdict, changed}: the host polls them
                      //     WITHOUT synchronising the stream and keeps enqueuing rounds while the loop goes on
  // sparse bound exchange over NCCL (fallback when peer memory is unavailable): this rank's changed candidates
  // {count, row-infeasible flag | entries {lb, ub, j}} are all-gathered instead of all-reducing 16 bytes per variable
  BoundMsg *xsend; // [1 + xcap]  header + entries of this rank
  BoundMsg *xrecv; // [n_ranks][1 + xcap]
  int32_t xcap;    // entries a rank can send (0: no sparse exchange)
  int32_t n_ranks;
  int32_t rank;
  // bound exchange over NVLink peer memory (one process per GPU, buffers shared with CUDA IPC): every rank PUSHES its
  // touched candidates into every peer's inbox and raises the peer's round tag; no host in the loop, no collective
  // call.  Inboxes are double-buffered by round parity (a peer can be at most one round ahead).
  int32_t p2p;               // 1: peer-memory exchange is set up
  int64_t inbox_stride;      // entries per (parity, sender) slot: 1 header + capacity (= n: a message always fits)
  BoundMsg *inbox;           // [2][n_ranks][inbox_stride]  this rank's inbox
  unsigned *inbox_tag;       // [n_ranks] last round whose message from rank r is complete
  BoundMsg *const *peer_inbox;   // [n_ranks] device array: peers' inbox base pointers (own entry unused)
  unsigned *const *peer_tag;     // [n_ranks] peers' tag arrays
};
// ctrl words of the per-round kernels
constexpr int kRcChanged = 0;   // a bound moved in this round
constexpr int kRcIntMoved = 1;  // a row moved an integer variable in this round
constexpr int kRcTouched = 2;   // length of tlist
constexpr int kRcVerdict = 3;
constexpr int kRcPairs = 4;     // changed (variable, round) pairs of the call
constexpr int kRcOverflow = 5;  // NCCL sparse exchange overflowed: the vars kernel did nothing, redo the merge densely
constexpr int kRcStop = 6;      // the loop is over: every later kernel of the call returns at once
constexpr int kRcRound = 7;     // rounds finished
constexpr int kRcEver = 8;      // length of elist
constexpr int kRcDoneA = 9;     // blocks of the current push kernel that have finished
constexpr int kRcDoneB = 10;    // ... of the current vars kernel
constexpr int kRcWords = 16;

}  // namespace mntr
