// kernels_dynamic.cuh -- dynamic pivoting ('minimal_difference', 'maximal_stability'; Reimer.py:497-523).
//
// Every step evaluates the scalar rule for ALL remaining rows and picks the minimum of the reference's
// key -- (f, -d, omega, j) or (-d, f, omega, j) -- so alpha must be current for every remaining row at every
// column.  That rules out the blocked panel scheme; this path is an unblocked right-looking factorization
// (rank-1 update per column), meant for the sizes at which the reference's O(n^2) rule evaluations are
// practical at all (n up to ~10^4).  Five small kernels per column, no host round trip.
#pragma once
#include "mdb_common.cuh"

struct DynSel {      // decision of the current step, written by k_dyn_select2
  double d, omega, drop;
  long long jstar;
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

__device__ __forceinline__ DynCand dyn_block_reduce(int method, DynCand c, DynCand* sh) {
  const int tid = threadIdx.x;
  sh[tid] = c;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (tid < s) {
      DynCand o = sh[tid + s];
      if (dyn_better(method, o, sh[tid])) sh[tid] = o;
    }
    __syncthreads();
  }
  return sh[0];
}

// stage 1: one rule evaluation per remaining row, best per block
__global__ void __launch_bounds__(256) k_dyn_select1(FactorDev F, int method, int64_t i, DynCand* __restrict__ cand) {
  __shared__ DynCand sh[256];
  const int64_t j = i + (int64_t)blockIdx.x * 256 + threadIdx.x;
  DynCand c;
  c.j = -1; c.d = 0; c.omega = 0; c.f = 0; c.clamped = 0;
  if (j < F.n) {
    const MdDecision dec = md_decide(F.prm, F.alpha[j], F.beta[j], F.gamma[j], F.minB[j], F.maxB[j]);
    c.d = dec.d; c.omega = dec.omega; c.f = dec.f; c.clamped = dec.clamped; c.j = j;
  }
  c = dyn_block_reduce(method, c, sh);
  if (threadIdx.x == 0) cand[blockIdx.x] = c;
}

// stage 2: best over blocks; swap the position-indexed vectors i <-> j*; record the decision
__global__ void __launch_bounds__(256) k_dyn_select2(FactorDev F, int method, int64_t i, int nblocks,
                                                     const DynCand* __restrict__ cand, double* __restrict__ minB,
                                                     double* __restrict__ maxB, int64_t* __restrict__ p,
                                                     DynSel* __restrict__ sel) {
  __shared__ DynCand sh[256];
  DynCand c;
  c.j = -1; c.d = 0; c.omega = 0; c.f = 0; c.clamped = 0;
  for (int b = threadIdx.x; b < nblocks; b += 256)
    if (dyn_better(method, cand[b], c)) c = cand[b];
  c = dyn_block_reduce(method, c, sh);
  if (threadIdx.x == 0) {
    const int64_t js = c.j;
    if (js != i) {  // Reimer.py:509-512 (p) -- and every per-row quantity kept in position order
      double t;
      t = F.gamma[i]; F.gamma[i] = F.gamma[js]; F.gamma[js] = t;
      t = F.alpha[i]; F.alpha[i] = F.alpha[js]; F.alpha[js] = t;
      t = F.beta[i]; F.beta[i] = F.beta[js]; F.beta[js] = t;
      t = F.rowmax[i]; F.rowmax[i] = F.rowmax[js]; F.rowmax[js] = t;
      t = minB[i]; minB[i] = minB[js]; minB[js] = t;
      t = maxB[i]; maxB[i] = maxB[js]; maxB[js] = t;
      const int64_t tp = p[i]; p[i] = p[js]; p[js] = tp;
    }
    const double alpha_i = F.alpha[i], gamma_i = F.gamma[i];
    F.d[i] = c.d;
    F.omega[i] = c.omega;
    F.delta[i] = (c.d - gamma_i) + (c.omega != 0 ? c.omega * c.omega * alpha_i : 0.0);  // Reimer.py:541-545
    F.clamped[i] = (uint8_t)c.clamped;
    sel->d = c.d;
    sel->omega = c.omega;
    sel->drop = (c.omega == 0 || c.omega * F.rowmax[i] < F.prm.min_abs_value_L) ? 1.0 : 0.0;
    sel->jstar = js;
  }
}

// symmetric swap of indices i and j* in both triangles of W: S(a, b) lives at W[max][min], A(a, b) at W[min][max]
// (Reimer.py:513-518 swaps the rows of L; the Schur complement and the original matrix must follow)
__global__ void __launch_bounds__(256) k_dyn_swap(int64_t n, double* __restrict__ W, int64_t ld, int64_t i,
                                                  const DynSel* __restrict__ sel) {
  const int64_t js = sel->jstar;
  if (js == i) return;
  const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (m >= n || m == i || m == js) return;
  {  // lower triangle: S
    double* a = &W[(i > m ? i : m) * ld + (i > m ? m : i)];
    double* b = &W[(js > m ? js : m) * ld + (js > m ? m : js)];
    const double t = *a; *a = *b; *b = t;
  }
  {  // upper triangle: A
    double* a = &W[(i > m ? m : i) * ld + (i > m ? i : m)];
    double* b = &W[(js > m ? m : js) * ld + (js > m ? js : m)];
    const double t = *a; *a = *b; *b = t;
  }
}

// column i for every row below: mix, divide, alpha / beta / rowmax; x and y = -x d for the rank-1 update
__global__ void __launch_bounds__(256) k_dyn_column(FactorDev F, int64_t i, const DynSel* __restrict__ sel,
                                                    double* __restrict__ xcol, double* __restrict__ ycol) {
  const int64_t j = i + 1 + (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= F.n) return;
  const double d = sel->d, w = sel->omega;
  const bool drop = sel->drop != 0.0;
  const double a = F.W[i * F.ldw + j];
  double r = a;
  if (!drop) r = (1.0 - w) * a + w * F.W[j * F.ldw + i];
  const double x = (d != 0) ? r / d : 0.0;
  F.W[j * F.ldw + i] = x;
  F.alpha[j] += x * x * d;
  F.beta[j] += 2.0 * a * a;
  F.rowmax[j] = fmax(F.rowmax[j], fabs(x));
  xcol[j] = x;
  ycol[j] = -x * d;
}

// S(j, j') += x_j y_j' for j > j' > i
__global__ void __launch_bounds__(256) k_dyn_rank1(int64_t n, double* __restrict__ W, int64_t ld, int64_t i,
                                                   const double* __restrict__ xcol, const double* __restrict__ ycol) {
  const int64_t j0 = i + 1 + (int64_t)blockIdx.y * 32, c0 = i + 1 + (int64_t)blockIdx.x * 32;
  if (c0 > j0 + 31) return;  // tile above the diagonal
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = c0 + tx;
  if (c >= n) return;
  const double y = ycol[c];
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = j0 + r;
    if (j < n && j > c) W[j * ld + c] += xcol[j] * y;
  }
}

__global__ void k_fill_iota(int64_t n, int64_t* __restrict__ p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
