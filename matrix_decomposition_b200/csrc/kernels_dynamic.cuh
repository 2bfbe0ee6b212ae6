// kernels_dynamic.cuh -- dynamic pivoting ('minimal_difference', 'maximal_stability'; Reimer.py:497-523).
//
// Every step evaluates the scalar rule for ALL remaining rows and picks the minimum of the reference's
// key -- (f, -d, omega, j) or (-d, f, omega, j) -- so alpha must be current for every remaining row at every
// column: the whole column below the pivot has to be computed at every step (no row-local diagonal block as in
// the static path).  The Schur complement, however, is only updated once per panel of 128 columns: inside a
// panel column i is read as  S[j, i] + sum_c X[j, c] Y[i, c]  over the panel's earlier columns c (X = unscaled
// rows, Y = -X D, kept for every row in the operand buffers; O(n * 64) per column instead of the O(n^2) rank-1
// update), and after the panel the trailing matrix gets the same K = 128 tensor-core update as the static path.
// Pivot swaps exchange the rows of X and Y together with the rows / columns of W.  One kernel per column
// (k_dyn_column), no host round trip; n^2 / 2 rule evaluations in total.
#pragma once
#include "mdb_common.cuh"

struct DynSel {      // decision of one step, written by dyn_commit: everything the swap and column kernels of that step
  double d, omega, drop;   // need sits in one 40-byte record (one memory round trip; a separate step counter cost a
  long long jstar;         // second, dependent one in both kernels of every column)
  long long step;          // the column index this decision is for
};

struct DynCand {
  double d, omega, f;
  long long j;
  int clamped;
};

// key order of Reimer.py:499-506; true when a is strictly better than b
__device__ __forceinline__ bool dyn_better(int method, const DynCand& a, const DynCand& b) {
  if (b.j < 0) return a.j >= 0;
  if (a.j < 0) return false;
  if (method == 1) {  // minimal_difference: (f, -d, omega, j)
    if (a.f != b.f) return a.f < b.f;
    if (a.d != b.d) return a.d > b.d;
  } else {            // maximal_stability: (-d, f, omega, j)
    if (a.d != b.d) return a.d > b.d;
    if (a.f != b.f) return a.f < b.f;
  }
  if (a.omega != b.omega) return a.omega < b.omega;
  return a.j < b.j;
}

__device__ __forceinline__ DynCand dyn_shfl_xor(const DynCand& c, int o) {
  DynCand r;
  r.d = __shfl_xor_sync(0xffffffffu, c.d, o);
  r.omega = __shfl_xor_sync(0xffffffffu, c.omega, o);
  r.f = __shfl_xor_sync(0xffffffffu, c.f, o);
  r.j = __shfl_xor_sync(0xffffffffu, c.j, o);
  r.clamped = __shfl_xor_sync(0xffffffffu, c.clamped, o);
  return r;
}

// Best candidate of the first `warps` warps of the block (the others only take part in the barrier): butterfly inside
// each warp, one shared-memory slot per warp, the final few merged by every thread itself.  The key is a strict total
// order (j is unique), so the result does not depend on the order of the comparisons.
__device__ __forceinline__ DynCand dyn_block_reduce(int method, DynCand c, DynCand* sh, int warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < warps) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const DynCand r = dyn_shfl_xor(c, o);
      if (dyn_better(method, r, c)) c = r;
    }
    if (lane == 0) sh[warp] = c;
  }
  __syncthreads();
  DynCand best = sh[0];
  for (int w = 1; w < warps; ++w) {
    const DynCand o = sh[w];
    if (dyn_better(method, o, best)) best = o;
  }
  __syncthreads();   // sh may be reused by the caller
  return best;
}

// stage 1: one rule evaluation per remaining row, best per block
__global__ void __launch_bounds__(256) k_dyn_select1(FactorDev F, int method, int64_t i, DynCand* __restrict__ cand) {
  __shared__ DynCand sh[8];
  const int64_t j = i + (int64_t)blockIdx.x * 256 + threadIdx.x;
  DynCand c;
  c.j = -1; c.d = 0; c.omega = 0; c.f = 0; c.clamped = 0;
  if (j < F.n) {
    const MdDecision dec = md_decide(F.prm, F.alpha[j], F.beta[j], F.gamma[j], F.minB[j], F.maxB[j]);
    c.d = dec.d; c.omega = dec.omega; c.f = dec.f; c.clamped = dec.clamped; c.j = j;
  }
  c = dyn_block_reduce(method, c, sh, 8);
  if (threadIdx.x == 0) cand[blockIdx.x] = c;
}

// the decision for step i, given the best candidate c: swap the position-indexed vectors i <-> j*, record d / omega /
// delta / clamped and what the swap and column kernels need.  One thread.
__device__ __forceinline__ void dyn_commit(const FactorDev& F, int64_t i, const DynCand& c, double* __restrict__ minB,
                                           double* __restrict__ maxB, int64_t* __restrict__ p, DynSel* __restrict__ sel) {
  const int64_t js = c.j;
  // Reimer.py:509-512 (p) -- and every per-row quantity kept in position order.  All loads first (independent, one
  // memory latency), then the stores: one thread swapping vector after vector paid ~14 dependent round trips per column.
  // (alpha / beta / rowmax of row j* were written by another block of the same launch: loads that bypass L1)
  const double g_i = __ldcg(F.gamma + i), g_s = __ldcg(F.gamma + js), a_i = __ldcg(F.alpha + i), a_s = __ldcg(F.alpha + js);
  const double b_i = __ldcg(F.beta + i), b_s = __ldcg(F.beta + js), r_i = __ldcg(F.rowmax + i), r_s = __ldcg(F.rowmax + js);
  const double m_i = __ldcg(minB + i), m_s = __ldcg(minB + js), M_i = __ldcg(maxB + i), M_s = __ldcg(maxB + js);
  const int64_t p_i = (int64_t)__ldcg((const long long*)p + i), p_s = (int64_t)__ldcg((const long long*)p + js);
  if (js != i) {
    F.gamma[i] = g_s; F.gamma[js] = g_i;
    F.alpha[i] = a_s; F.alpha[js] = a_i;
    F.beta[i] = b_s; F.beta[js] = b_i;
    F.rowmax[i] = r_s; F.rowmax[js] = r_i;
    minB[i] = m_s; minB[js] = m_i;
    maxB[i] = M_s; maxB[js] = M_i;
    p[i] = p_s; p[js] = p_i;
  }
  const double alpha_i = a_s, gamma_i = g_s, rowmax_i = r_s;   // (js == i: the same values)
  F.d[i] = c.d;
  F.omega[i] = c.omega;
  F.delta[i] = (c.d - gamma_i) + (c.omega != 0 ? c.omega * c.omega * alpha_i : 0.0);  // Reimer.py:541-545
  F.clamped[i] = (uint8_t)c.clamped;
  sel->d = c.d;
  sel->omega = c.omega;
  sel->drop = (c.omega == 0 || c.omega * rowmax_i < F.prm.min_abs_value_L) ? 1.0 : 0.0;
  sel->jstar = js;
  sel->step = i;
}

// stage 2: best over blocks, then dyn_commit.  Only the first column goes through select1 / select2: afterwards
// k_dyn_column evaluates the rule for the next step itself.
__global__ void __launch_bounds__(256) k_dyn_select2(FactorDev F, int method, int64_t i, int nblocks,
                                                     const DynCand* __restrict__ cand, double* __restrict__ minB,
                                                     double* __restrict__ maxB, int64_t* __restrict__ p,
                                                     DynSel* __restrict__ sel) {
  __shared__ DynCand sh[8];
  DynCand c;
  c.j = -1; c.d = 0; c.omega = 0; c.f = 0; c.clamped = 0;
  for (int b = threadIdx.x; b < nblocks; b += 256)
    if (dyn_better(method, cand[b], c)) c = cand[b];
  c = dyn_block_reduce(method, c, sh, 8);
  if (threadIdx.x == 0) dyn_commit(F, i, c, minB, maxB, p, sel + (i & 1));
}

// One kernel per column.  The pivot swap i <-> j* (Reimer.py:509-518: p, the rows of L; here also the Schur complement
// in the lower triangle of W, S(a, b) at W[max][min], the original matrix in the upper one, A(a, b) at W[min][max], and
// the pending rows of X / Y) is not a pass of its own: the column computation READS THROUGH the swap -- new row j is old
// row (j == j* ? i : j), new column i is old column j* -- and each thread then puts the two pairs of elements of its own
// row where the swap wants them (the value it consumed is replaced by what stood in column i; x goes to column i).  A
// separate swap kernel in front of every column kernel cost a launch gap and ~3 us of dependent strided accesses per
// column.  What no row thread owns is done by
//   * the blocks behind the row blocks: positions m < i (the finished part of rows i and j* of L, coalesced, and of
//     columns i and j* of A);
//   * the block that finishes last: rows i and j* of X / Y (every block reads Y[j*] as the new Y[i], so they can only
//     move when all are done), then the decision for step i + 1.
//
// Column i for every row below: current Schur complement entry (stored value + the pending updates of this panel),
// mix, divide, alpha / beta / rowmax; x and y = -x d go into column kc of the operand buffers.  A CTA owns
// MDB_DYN_ROWS = 64 rows (4 CTAs per 256 rows: every SM gets work also at n = 16384):
//   pass 1  a warp per row, lanes over the pending columns: sum_c X[j, c] Y[i, c] with coalesced loads of the row of X
//           (a thread per row walked its row alone: 32 different cache lines per load instruction) and a fixed butterfly;
//   pass 2  a thread per row: the mix, x, the row state and -- its alpha / beta being current -- the rule for step i + 1.
// The block minima are merged by the last block to finish, which commits the decision for step i + 1 into the other
// decision record (the two alternate, so no block can still be reading the one being written).
#define MDB_DYN_ROWS 64

__global__ void __launch_bounds__(256) k_dyn_column(FactorDev F, int method, const DynSel* __restrict__ sel2, int slot,
                                                    int row_blocks, double* __restrict__ PX, double* __restrict__ PY,
                                                    int64_t ldp, DynCand* cand, unsigned* counter,
                                                    double* __restrict__ minB, double* __restrict__ maxB,
                                                    int64_t* __restrict__ p) {
  // sel2 = the pair of decision records; this step reads record `slot` (the parity of its column).  In a replayed CUDA
  // graph the grid is sized for the first column of the panel: blocks without rows only take part in the count.
  const DynSel* sel = sel2 + slot;
  DynSel* sel_next = const_cast<DynSel*>(sel2) + (slot ^ 1);
  const int64_t i = sel->step;
  const int64_t js = sel->jstar;
  const int kc = (int)(i % MDB_NB);
  const double d = sel->d, w = sel->omega;
  const bool drop = sel->drop != 0.0;
  const bool sw = js != i;
  double* __restrict__ W = F.W;
  const int64_t ld = F.ldw;
  __shared__ double s_sum[MDB_DYN_ROWS];
  __shared__ DynCand sh[8];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  DynCand best;
  best.j = -1; best.d = 0; best.omega = 0; best.f = 0; best.clamped = 0;
  if ((int)blockIdx.x >= row_blocks) {
    // positions m < i of the swap
    const int64_t m = (int64_t)(blockIdx.x - row_blocks) * 256 + tid;
    if (sw && m < i) {
      const double l_i = W[i * ld + m], l_s = W[js * ld + m], a_i = W[m * ld + i], a_s = W[m * ld + js];
      W[i * ld + m] = l_s; W[js * ld + m] = l_i;
      W[m * ld + i] = a_s; W[m * ld + js] = a_i;
    }
  } else {
    const int64_t j0 = i + 1 + (int64_t)blockIdx.x * MDB_DYN_ROWS;
    // pass 2's operands are requested first: their latency (strided elements) overlaps pass 1
    const int64_t j = j0 + tid;
    const bool mine = tid < MDB_DYN_ROWS && j < F.n;
    const bool moved = mine && sw && j != js;    // this row's elements of columns / rows i and j* change places
    double a = 0, sji = 0, old_s = 0, old_a = 0, al0 = 0, be0 = 0, ga = 0, mb = 0, Mb = 0, rm = 0;
    double *loc_s = nullptr, *loc_a = nullptr;
    if (mine) {
      loc_s = &W[j * ld + i];
      loc_a = &W[i * ld + j];
      if (moved) {
        loc_s = &W[(js > j ? js : j) * ld + (js > j ? j : js)];   // old S(j, j*)
        loc_a = &W[(js > j ? j : js) * ld + (js > j ? js : j)];   // old A(j*, j)
        old_s = W[j * ld + i];
        old_a = W[i * ld + j];
      }
      sji = *loc_s;
      a = *loc_a;
      al0 = F.alpha[j]; be0 = F.beta[j]; ga = F.gamma[j]; rm = F.rowmax[j];
      mb = minB[j]; Mb = maxB[j];
    }
    if (!drop) {   // pass 1: a warp per row, 8 consecutive rows per warp, every load issued before the first use
      double y[MDB_NB / 32];
#pragma unroll
      for (int e = 0; e < MDB_NB / 32; ++e) {
        const int c = lane + 32 * e;
        y[e] = (c < kc) ? PY[js * ldp + c] : 0.0;   // new Y[i, 0:kc] = old Y[j*, 0:kc]
      }
      double xv[MDB_DYN_ROWS / 8][MDB_NB / 32];
#pragma unroll
      for (int q = 0; q < MDB_DYN_ROWS / 8; ++q) {
        const int64_t jr = j0 + warp * (MDB_DYN_ROWS / 8) + q;
        const int64_t pr = (jr == js) ? i : jr;     // where the row still stands
#pragma unroll
        for (int e = 0; e < MDB_NB / 32; ++e) {
          const int c = lane + 32 * e;
          xv[q][e] = (jr < F.n && c < kc) ? PX[pr * ldp + c] : 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < MDB_DYN_ROWS / 8; ++q) {
        double sacc = 0;
#pragma unroll
        for (int e = 0; e < MDB_NB / 32; ++e) {
          const int c = lane + 32 * e;
          if (c < kc) sacc = fma(xv[q][e], y[e], sacc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (lane == 0) s_sum[warp * (MDB_DYN_ROWS / 8) + q] = sacc;
      }
    }
    __syncthreads();
    if (mine) {   // pass 2: a thread per row
      double r = a;
      if (!drop) r = (1.0 - w) * a + w * (sji + s_sum[tid]);
      const double x = (d != 0) ? r / d : 0.0;
      W[j * ld + i] = x;
      if (moved) {
        *loc_s = old_s;          // new S(j, j*) = old S(j, i)
        W[i * ld + j] = a;       // new A(i, j) = old A(j*, j)
        *loc_a = old_a;          // new A(j*, j) = old A(i, j)
      }
      const double al = al0 + x * x * d, be = be0 + 2.0 * a * a;
      F.alpha[j] = al;
      F.beta[j] = be;
      F.rowmax[j] = fmax(rm, fabs(x));
      const int64_t pj = (sw && j == js) ? i : j;   // row j* of X / Y still stands at i (the last block moves it)
      PX[pj * ldp + kc] = x;
      PY[pj * ldp + kc] = -x * d;
      const MdDecision dec = md_decide(F.prm, al, be, ga, mb, Mb);
      best.d = dec.d; best.omega = dec.omega; best.f = dec.f; best.clamped = dec.clamped; best.j = j;
    }
    best = dyn_block_reduce(method, best, sh, MDB_DYN_ROWS / 32);
  }
  if (tid == 0) {
    cand[blockIdx.x] = best;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // rows i and j* of X / Y, columns 0 .. kc (written by other blocks of this launch: loads that bypass L1)
  double x_i = 0, x_s = 0, y_i = 0, y_s = 0;
  const bool mv = sw && tid <= kc;
  if (mv) {
    x_i = __ldcg(PX + i * ldp + tid); x_s = __ldcg(PX + js * ldp + tid);
    y_i = __ldcg(PY + i * ldp + tid); y_s = __ldcg(PY + js * ldp + tid);
  }
  DynCand c;
  c.j = -1; c.d = 0; c.omega = 0; c.f = 0; c.clamped = 0;
  for (int b = tid; b < row_blocks; b += 256) {
    const volatile DynCand* vc = cand + b;   // written by other blocks of this launch
    DynCand o;
    o.d = vc->d; o.omega = vc->omega; o.f = vc->f; o.j = vc->j; o.clamped = vc->clamped;
    if (dyn_better(method, o, c)) c = o;
  }
  if (mv) {
    PX[i * ldp + tid] = x_s; PX[js * ldp + tid] = x_i;
    PY[i * ldp + tid] = y_s; PY[js * ldp + tid] = y_i;
  }
  c = dyn_block_reduce(method, c, sh, 8);
  if (tid == 0) {
    *counter = 0;
    dyn_commit(F, i + 1, c, minB, maxB, p, sel_next);
  }
}

__global__ void k_fill_iota(int64_t n, int64_t* __restrict__ p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
