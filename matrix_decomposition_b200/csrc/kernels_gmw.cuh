// kernels_gmw.cuh -- the GMW family of modified Cholesky factorizations (GMW_SE.py:6-35, :67-152 of the reference):
// right-looking rank-1 LDL^T without pivoting in which d_k is chosen from the current column c = S[k+1:, k]
// (|c|_inf) and the diagonal of the trailing matrix, in two phases.  d_k depends on the WHOLE column below the
// pivot, so this is an unblocked factorization: per column a reduction, a scalar decision and a rank-1 update
// (HBM bound: 16 (n-k)^2 / 2 bytes per column), meant for n up to ~10^4 like the dynamic-pivoting path.
// Arithmetic mirrors numpy operation by operation (explicitly rounded products / quotients, no FMA contraction), so
// d and delta equal the reference's bit for bit.
#pragma once
#include "mdb_common.cuh"

struct GmwState {
  double a, cmax, m1, m2;       // pivot, |c|_inf, min_j (S_jj - c_j^2 / a), min_j S_jj  (j > k)
  double beta2, prev_delta, dk;
  double min_d_next, min_d_factor, min_diag_D;
  unsigned long long max_diag_bits, max_off_bits;   // max |.| of the trailing matrix (non-negative doubles as bits)
  int phase;                    // 10 = phase 1, 15 = phase 1.5, 20 = phase 2   (GMW_SE.py:84, :100-110)
  int need_init;
  int use_abs, nondecreasing;
  unsigned int init_done;       // blocks of k_gmw_init_decide that have contributed their maxima
  // SE family (GMW_SE.py:155-192): d from the 1-norm of the column, tau * max |diag| of the trailing matrix at the switch
  // to phase 2, a 2 x 2 eigenvalue step for the last two columns
  int family;                   // 0 = GMW, 1 = SE
  double c1;                    // |c|_1
  double tau, eps13, tau_max, d_last;
  int have_d_last;
};

// phase 1 test (GMW_SE.py:100-104): keep d = a, or switch to phase 1.5 and ask for the trailing-matrix maxima
// (one thread, at the end of k_gmw_reduce)
__device__ __forceinline__ void gmw_phase(GmwState* st) {
  if (st->phase == 10) {
    const double a = st->a;
    // np.all over an empty trailing matrix is True: m1 = m2 = +inf then
    if (a >= st->min_diag_D && st->m1 >= st->min_d_next && st->m2 >= __dmul_rn(st->min_d_factor, a)) {
      st->dk = a;
    } else {
      st->phase = 15;
    }
  }
  if (st->phase == 15) {
    st->need_init = 1;
    st->max_diag_bits = 0ull;
    st->max_off_bits = 0ull;
  }
  st->init_done = 0u;
}

// column k below the pivot -> cvec (contiguous), plus the reductions the decision needs and the phase test.  One CTA.
__global__ void __launch_bounds__(1024) k_gmw_reduce(int64_t n, const double* __restrict__ W, int64_t ld, int64_t k,
                                                     double* __restrict__ cvec, GmwState* st) {
  __shared__ double s_cmax[32], s_m1[32], s_m2[32], s_c1[32];
  const double a = W[k * ld + k];
  double cmax = 0.0, m1 = INFINITY, m2 = INFINITY, c1 = 0.0;
  for (int64_t j = k + 1 + threadIdx.x; j < n; j += blockDim.x) {
    const double c = W[j * ld + k];
    const double djj = W[j * ld + j];
    cvec[j] = c;
    cmax = fmax(cmax, fabs(c));
    c1 += fabs(c);
    m1 = fmin(m1, __dsub_rn(djj, __ddiv_rn(__dmul_rn(c, c), a)));   // A.diagonal() - c**2 / a   (:101)
    m2 = fmin(m2, djj);
  }
  for (int o = 16; o > 0; o >>= 1) {
    cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    m1 = fmin(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_cmax[w] = cmax; s_m1[w] = m1; s_m2[w] = m2; s_c1[w] = c1; }
  __syncthreads();
  if (w == 0) {
    cmax = (l < (int)(blockDim.x >> 5)) ? s_cmax[l] : 0.0;
    m1 = (l < (int)(blockDim.x >> 5)) ? s_m1[l] : INFINITY;
    m2 = (l < (int)(blockDim.x >> 5)) ? s_m2[l] : INFINITY;
    c1 = (l < (int)(blockDim.x >> 5)) ? s_c1[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      m1 = fmin(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if (l == 0) {
      st->a = a; st->cmax = cmax; st->m1 = m1; st->m2 = m2; st->c1 = c1;
      gmw_phase(st);
    }
  }
}

// d_k (GMW_SE.py:106-117, :123-148), delta_k = d_k - a (:29).  One thread.
__device__ __forceinline__ void gmw_decide(int64_t n, int64_t k, GmwState* st, double* __restrict__ d, double* __restrict__ delta,
                                           const double* __restrict__ W, int64_t ld) {
  const double eps = 2.220446049250313e-16;
  const int64_t m = n - k - 1;  // size of the trailing matrix
  if (st->need_init) {
    const double max_diag = __longlong_as_double((long long)st->max_diag_bits);
    const double max_off = __longlong_as_double((long long)st->max_off_bits);
    double beta2;
    if (m > 1) {
      const double x = st->use_abs ? (double)(m * m - 1) : (double)(m * m - m);
      beta2 = fmax(fmax(eps, max_diag), __ddiv_rn(max_off, sqrt(x)));
    } else if (m == 1) {
      beta2 = fmax(eps, max_diag);
    } else {
      beta2 = eps;
    }
    st->beta2 = beta2;
    st->tau_max = __dmul_rn(st->tau, (m > 0) ? max_diag : 0.0);   // SE: init_choose_d_state, GMW_SE.py:163-169
    st->need_init = 0;
    st->phase = 20;
  }
  double dk = st->dk;
  const double a = st->a;
  if (st->phase == 20) {
    if (st->family == 0) {          // _GMW.choose_d, GMW_SE.py:145-148
      dk = (m > 0) ? __ddiv_rn(__dmul_rn(st->cmax, st->cmax), st->beta2) : 0.0;
    } else if (m > 1) {             // _SE.choose_d, GMW_SE.py:171-190
      dk = fmax(st->c1, st->tau_max);
    } else if (m == 1) {
      // eigenvalues of [[a, c], [c, A]] (np.linalg.eigvalsh in the reference; closed form here)
      const double c0 = W[(k + 1) * ld + k], a11 = W[(k + 1) * ld + (k + 1)];
      const double mean = 0.5 * (a + a11), r = hypot(0.5 * (a - a11), c0);
      const double l1 = mean - r, l2 = mean + r;
      dk = (a - l1) + fmax(st->tau * (l2 - l1) / (1.0 - st->tau), st->tau_max);
      if (st->use_abs) dk = fmax(dk, a - 2.0 * l1);
      st->d_last = dk;
      st->have_d_last = 1;
    } else {
      dk = st->have_d_last ? st->d_last : fmax((-st->eps13 * a) / (1.0 - st->eps13), st->tau_max);
    }
    dk = fmax(dk, st->min_diag_D);
    double a_changed = st->use_abs ? fabs(a) : a;
    if (st->nondecreasing) a_changed = __dadd_rn(a_changed, st->prev_delta);
    dk = fmax(dk, a_changed);
    st->prev_delta = __dsub_rn(dk, a);
    st->dk = dk;
  }
  d[k] = dk;
  delta[k] = __dsub_rn(dk, a);
}

// One launch for the decision of column k: when phase 1.5 asked for them (once per factorization), max |diag| and
// max |strict lower| of the trailing matrix S[k+1:, k+1:] (GMW_SE.py:127-133), one row per block, and the last block to
// finish decides; otherwise block 0 decides at once and the others leave.  grid = max(n - k - 1, 1) blocks.
__global__ void __launch_bounds__(256) k_gmw_init_decide(int64_t n, const double* __restrict__ W, int64_t ld, int64_t k,
                                                         GmwState* st, double* __restrict__ d, double* __restrict__ delta) {
  __shared__ bool s_last;
  if (!st->need_init) {
    if (blockIdx.x == 0 && threadIdx.x == 0) gmw_decide(n, k, st, d, delta, W, ld);
    return;
  }
  const int64_t i = k + 1 + blockIdx.x;
  if (i < n) {
    double moff = 0.0;
    if (st->family == 0)   // (the SE family only needs the diagonal)
      for (int64_t j = k + 1 + threadIdx.x; j < i; j += blockDim.x) moff = fmax(moff, fabs(W[i * ld + j]));
    for (int o = 16; o > 0; o >>= 1) moff = fmax(moff, __shfl_xor_sync(0xffffffffu, moff, o));
    if ((threadIdx.x & 31) == 0 && moff > 0.0) atomicMax(&st->max_off_bits, (unsigned long long)__double_as_longlong(moff));
    if (threadIdx.x == 0) atomicMax(&st->max_diag_bits, (unsigned long long)__double_as_longlong(fabs(W[i * ld + i])));
  }
  __syncthreads();   // this block's atomics have been issued
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&st->init_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    volatile GmwState* vs = st;   // the maxima were written by other blocks of this launch
    st->max_off_bits = vs->max_off_bits;
    st->max_diag_bits = vs->max_diag_bits;
    gmw_decide(n, k, st, d, delta, W, ld);
  }
}

// S[i, j] -= c_i c_j / d_k  for i >= j > k  (A -= np.outer(c, c) / d[k], GMW_SE.py:31), 32 x 32 tiles of the lower triangle
__global__ void __launch_bounds__(256) k_gmw_rank1(int64_t n, double* __restrict__ W, int64_t ld, int64_t k,
                                                   const double* __restrict__ cvec, const GmwState* __restrict__ st) {
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const double dk = st->dk;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t j = k + 1 + bj * 32 + tx;
  if (j >= n) return;
  const double cj = cvec[j];
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = k + 1 + bi * 32 + r;
    if (i < n && i >= j) {
      const double v = W[i * ld + j];
      W[i * ld + j] = __dsub_rn(v, __ddiv_rn(__dmul_rn(cvec[i], cj), dk));
    }
  }
}
