// mdb_common.cuh -- shared definitions of libmdb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/mdb200.h"
#include "md_rule.h"

#define MDB_NB 128  // panel width = distribution block = K of the trailing update
#define MDB_VINV_DOUBLES (4 * 32 * 32 + 8)

struct mdb_ctx {
  int device;
  cudaStream_t stream;
  cudaStream_t own_stream;
  cudaStream_t side_stream;  // low priority: the bulk of the trailing update under lookahead
  cudaStream_t copy_stream;  // finalize + device -> host copy of finished rows of L while the factorization goes on
  cudaEvent_t ev_copy;
  cudaEvent_t ev_panel[2], ev_upd[2];
  int lookahead;
  int use_tma;  // MDB200_TMA=0 falls back to the cp.async tile kernel
  int use_cluster_gather;  // MDB200_CLUSTER_GATHER=0: plain 8-byte gather instead of the cluster / DSMEM kernel
  int use_trsv; // MDB200_TRSV=0: per-block launches instead of the persistent triangular sweep (<= 8 right-hand sides)
  int agg_fixed;            // panels per outer block forced by MDB200_AGG (0 = choose by remaining size)
  int64_t agg_thresh[3];    // remaining sizes from which 2, 4, 8 panels are aggregated
  char err[512];
  int64_t launches;
  int profiling;
  double phase_ms[4];
  void* stage[3];             // pinned staging buffers for large pageable host <-> device copies
  cudaEvent_t stage_ev[3];
  size_t stage_bytes;
  std::thread* reaper;        // background release of the previous one-shot call's device buffers (cudaFree of tens
                              // of GB takes 0.2-0.4 s and nothing the caller waits for depends on it)
  void* big_ptr[3];           // released working matrix / staging / permutation buffers kept for the next call (mdb_ctx_trim)
  size_t big_bytes[3];
  int use_cache;
  uint8_t* tail_flags; size_t tail_flags_cap;   // per-row 'sent too early' flags of the streamed download (mdb_factor_run)
  int stage_next; bool stage_busy[3];   // rotation of the staging buffers across calls; buffer has a recorded, unwaited event
  struct mdb_factor* parked;  // workspace (everything but W) of the last one-shot call, reused by a call of the same shape
  int last_input_nonfinite;   // set by mdb_factor_set_matrix: the matrix held an inf / NaN
  double gemm_flop[2];        // algorithmic flop of the bulk / inner trailing updates of the last mdb_factor_run
  int64_t gemm_launches[2];
  int num_sms;
  cudaEvent_t ev[2];
};

static inline int mdb_fail(mdb_ctx* ctx, int code, const char* fmt, const char* a = "", const char* b = "") {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b);
  return code;
}

#define MDB_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s failed: %s (%s:%d)", #call,             \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                \
      return MDB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define MDB_CHECK(ctx, cond, msg)                                        \
  do {                                                                   \
    if (!(cond)) return mdb_fail((ctx), MDB_ERR_INVALID, "%s", (msg));   \
  } while (0)

#define MDB_LAUNCHED(ctx) ((ctx)->launches++)

// cudaFuncSetAttribute is per device: the "already configured" marks of the launchers are bit masks over devices
static inline bool mdb_device_marked(const uint64_t& mask, const mdb_ctx* ctx) { return (mask >> (ctx->device & 63)) & 1ull; }
static inline void mdb_mark_device(uint64_t& mask, const mdb_ctx* ctx) { mask |= 1ull << (ctx->device & 63); }

static inline int64_t mdb_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// device-side description of one factorization (passed by value to kernels)
// ------------------------------------------------------------------------------------------------
struct FactorDev {
  int64_t n;      // matrix dimension
  int64_t ldw;    // leading dimension of W
  double* W;      // n x ldw: strict lower = S / Lu (unscaled rows), strict upper = permuted original A
  double* gamma;  // n, diag of the permuted A (by position)
  double* alpha;  // n, sum_k Lu[j,k]^2 d_k accumulated column by column (Reimer.py:679-684)
  double* beta;   // n, sum_k 2 A[j,k]^2 (Reimer.py:602-604)
  double* rowmax; // n, max_k |Lu[j,k]| (for the all-dropped shortcut)
  double* d;      // n, by position
  double* omega;  // n, by position
  double* delta;  // n, by position
  uint8_t* clamped;  // n, by position
  const double* minB;  // n by position, or null (scalar bound in prm)
  const double* maxB;
  int* info;      // first failing pivot + 1 (MDB_RULE_EXACT), 0 = none
  double* Ut;     // NB x NB scratch: Ut[m * NB + k] = d_m * Ltilde[k][m]  (m < k)
  double* hdr;    // 5 x NB scratch: d, omega, drop flag, delta, clamped of the current panel
  double* Vinv;   // MDB_VINV_DOUBLES scratch: inverses of the four 32 x 32 scaled unit triangles of the diagonal block
                  // (row-major, [sub-block][c][m]) + [4096] = 1.0 when the panel must not use them
  // strides between consecutive matrices of a batch (elements)
  int64_t sW, sVec, sUt, sHdr, sPanel, sV;
  MdRuleParams prm;
};
