// kernels_factor.cuh -- gather, generator, diagonal-block, panel and finalize kernels of the blocked
// right-looking modified LDL^T (the restatement validated by oracle/blocked_model.py).
#pragma once
#include <cooperative_groups.h>

#include "mdb_common.cuh"

// ------------------------------------------------------------------------------------------------
// Symmetric gather  W[i, j] = A[p[i], p[j]]   (dense/permute.py:26), gamma[i] = A[p[i], p[i]].
// HBM bound: 8 n^2 written coalesced; reads are 8-byte gathers inside row p[i] (sector efficiency 1/4
// for a random p, 1 for the identity).  grid = (n rows, column chunks of 1024, batch).  Also raises
// *nonfinite when an element is inf / NaN (the reference's check_finite scan, dense/util.py:15-21).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_sym(int64_t n, const double* __restrict__ A, int64_t lda,
                                                    int64_t sA, const int64_t* __restrict__ p, int64_t sP,
                                                    double* __restrict__ W, int64_t ldw, int64_t sW,
                                                    double* __restrict__ gamma, int64_t sVec,
                                                    int* __restrict__ nonfinite, int upper_only) {
  // upper_only: A is symmetric and only its upper triangle (row <= column) is valid in the source buffer
  const int64_t b = blockIdx.z;
  bool bad = false;
  const int64_t i = blockIdx.x;
  A += b * sA;
  W += b * sW;
  const int64_t* pb = p ? p + b * sP : nullptr;
  const int64_t pi = pb ? pb[i] : i;
  const double* Arow = A + pi * lda;
  double* Wrow = W + i * ldw;
  const int64_t j0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t j = j0 + 256 * e;
    if (j < n) {
      const int64_t pj = pb ? pb[j] : j;
      const double v = (upper_only && pj < pi) ? A[pj * lda + pi] : Arow[pj];
      Wrow[j] = v;
      bad |= !isfinite(v);
      if (gamma && j == i) gamma[b * sVec + i] = v;
    }
  }
  if (bad && nonfinite) *nonfinite = 1;  // dense/util.py:15-21 (is_finite), evaluated where the data already is
}

// ------------------------------------------------------------------------------------------------
// The same gather with the source row staged in shared memory.  A thread-block cluster of C CTAs owns one output row:
// CTA r loads chunk r of source row p[i] with coalesced loads (the row is read from HBM exactly once), then every
// CTA writes its slice of the output row, fetching A[p[i], p[j]] from whichever CTA of the cluster holds column p[j]
// (distributed shared memory).  Traffic = the algorithmic 16 n^2 bytes for any permutation; the 8-byte gathers of
// k_gather_sym (sector efficiency 1/4) become shared-memory reads.  Chunks are 2^shift doubles (64 KB, three CTAs
// per SM, for n <= 65536; 128 KB beyond), C = 1, 2, 4, 8.  grid = (n * C, 1, batch), cluster (C, 1, 1).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_sym_cluster(int64_t n, const double* __restrict__ A, int64_t lda,
                                                            int64_t sA, const int64_t* __restrict__ p, int64_t sP,
                                                            double* __restrict__ W, int64_t ldw, int64_t sW,
                                                            double* __restrict__ gamma, int64_t sVec,
                                                            int* __restrict__ nonfinite, int shift) {
  namespace cg = cooperative_groups;
  extern __shared__ double smem_d[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int r = (int)cluster.block_rank();
  const int64_t chunk = (int64_t)1 << shift;
  const int64_t b = blockIdx.z;
  const int64_t i = blockIdx.x / C;
  A += b * sA;
  W += b * sW;
  const int64_t* pb = p ? p + b * sP : nullptr;
  const int64_t pi = pb ? pb[i] : i;
  const double* Arow = A + pi * lda;
  bool bad = false;
  {  // my chunk of the source row: coalesced, 4 independent loads in flight per thread
    const int64_t c0 = (int64_t)r * chunk;
    const int64_t len = (c0 < n) ? ((n - c0 < chunk) ? n - c0 : chunk) : 0;
    const double* src = Arow + c0;
    int64_t k = threadIdx.x;
    for (; k + 3 * 256 < len; k += 4 * 256) {
      const double v0 = src[k], v1 = src[k + 256], v2 = src[k + 512], v3 = src[k + 768];
      smem_d[k] = v0; smem_d[k + 256] = v1; smem_d[k + 512] = v2; smem_d[k + 768] = v3;
      bad |= !(isfinite(v0) && isfinite(v1) && isfinite(v2) && isfinite(v3));
    }
    for (; k < len; k += 256) {
      const double v = src[k];
      smem_d[k] = v;
      bad |= !isfinite(v);
    }
  }
  cluster.sync();
  {  // my slice of the output row
    const int64_t per = (((n + C - 1) / C) + 3) & ~(int64_t)3;
    const int64_t j0 = (int64_t)r * per, j1 = (j0 + per < n) ? j0 + per : n;
    double* Wrow = W + i * ldw;
    const int64_t mask = chunk - 1;
    auto fetch = [&](int64_t pj) -> double {
      const int owner = (int)(pj >> shift);
      const double* src = (owner == r) ? smem_d : cluster.map_shared_rank(smem_d, owner);
      return src[pj & mask];
    };
    int64_t j = j0 + threadIdx.x;
    for (; j + 3 * 256 < j1; j += 4 * 256) {
      const int64_t q0 = pb ? pb[j] : j, q1 = pb ? pb[j + 256] : j + 256, q2 = pb ? pb[j + 512] : j + 512,
                    q3 = pb ? pb[j + 768] : j + 768;
      const double v0 = fetch(q0), v1 = fetch(q1), v2 = fetch(q2), v3 = fetch(q3);
      Wrow[j] = v0; Wrow[j + 256] = v1; Wrow[j + 512] = v2; Wrow[j + 768] = v3;
    }
    for (; j < j1; j += 256) Wrow[j] = fetch(pb ? pb[j] : j);
    if (gamma && i >= j0 && i < j1 && threadIdx.x == 0) gamma[b * sVec + i] = fetch(pi);
  }
  if (bad && nonfinite) *nonfinite = 1;
  cluster.sync();  // nobody leaves while its shared memory may still be read
}

// lower triangle <- transpose of the upper triangle (the host upload sends only the upper triangle of a symmetric A)
__global__ void __launch_bounds__(256) k_mirror_upper_to_lower(int64_t n, double* __restrict__ A, int64_t lda) {
  __shared__ double tile[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;  // tile (bi, bj) of the UPPER triangle, bi <= bj
  if (bi > bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = bi * 32 + rr, j = bj * 32 + tx;
    tile[rr][tx] = (i < n && j < n) ? A[i * lda + j] : 0.0;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = bj * 32 + rr, j = bi * 32 + tx;  // destination (row in block bj, column in block bi)
    if (i < n && j < n && i > j) A[i * lda + j] = tile[tx][rr];
  }
}

// ------------------------------------------------------------------------------------------------
// Synthetic family F2: A = (G + G^T)/2 + mu I, G ~ N(0,1) iid.  Counter-based: element
// (i, j) depends only on (min, max, seed), so any rank can generate any column block, in any
// permuted order.  Off-diagonal ~ N(0, 1/2), diagonal ~ N(0, 1) + mu.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mdb_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ double mdb_f2_element(int64_t i, int64_t j, uint64_t seed, double mu) {
  const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
  const uint64_t h1 = mdb_splitmix64(seed ^ mdb_splitmix64(((uint64_t)lo << 32) ^ (uint64_t)hi));
  const uint64_t h2 = mdb_splitmix64(h1);
  const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0, 1)
  const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  // explicit roundings: no FMA contraction across the select, so every kernel that evaluates an element
  // (matrix, permuted matrix, diagonal only) produces the same bits
  const double z = __dmul_rn(sqrt(-2.0 * log(u1)), cospi(2.0 * u2));
  return (i == j) ? __dadd_rn(z, mu) : __dmul_rn(z, 0.70710678118654752440);
}

// out[r, c] = elem(prow[row0 + r], pcol[c]) for r < rows, c < cols (p null = identity)
__global__ void __launch_bounds__(256) k_generate_f2(int64_t rows, int64_t cols, int64_t row0,
                                                     const int64_t* __restrict__ prow,
                                                     const int64_t* __restrict__ pcol, uint64_t seed, double mu,
                                                     double* __restrict__ out, int64_t ld) {
  const int64_t r = blockIdx.x;
  const int64_t gi = prow ? prow[row0 + r] : row0 + r;
  const int64_t c0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t c = c0 + 256 * e;
    if (c < cols) {
      const int64_t gj = pcol ? pcol[c] : c;
      out[r * ld + c] = mdb_f2_element(gi, gj, seed, mu);
    }
  }
}

__global__ void k_generate_f2_diag(int64_t n, uint64_t seed, double mu, double* __restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = mdb_f2_element(i, i, seed, mu);
}

__global__ void k_extract_diag(int64_t n, const double* __restrict__ W, int64_t ldw, double* __restrict__ gamma) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gamma[i] = W[i * ldw + i];
}

// 8-byte asynchronous global -> shared copy (LDGSTS): lets one thread keep a whole tile column in flight
__device__ __forceinline__ void mdb_cp_async8(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void mdb_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Diagonal block: columns K = k0 .. k0+kb-1, rows of the block only.  One CTA, the whole block in shared
// memory, 4 threads per row.  Sequential over the columns because the decision for column k needs alpha_k,
// which contains Lu[k, k0:k]^2 d from this very block.  The kernel is pure latency, and per column the two
// expensive things are the rule (~1300 dependent scalar instructions when the column is clamped: two cubic
// solves, Reimer.py:95-115) and the length-k dot products of the rows below.  They overlap:
//   phase 1   lanes (k, 0) and (k, 1) : (d, omega) = rule(alpha_k, beta_k, gamma_k)     Reimer.py:479-494 / :42-176
//             (the candidates of d = lo and d = max_diag_D side by side, one instruction stream, merged with the
//             reference's key, md_rule.h)  WHILE
//             rows j > k : G_j = sum_{m < k-1} Lu[j, m] d_m Lu[k, m]   -- the dot product with the UNscaled row k,
//             which does not need the decision.  (d_m Lu[k, m] sits at the transposed position (m, k) of the tile:
//             the A entry stored there is dead once column m has been processed.)
//   barrier
//   phase 2   x[j, k] = ((1 - omega) A[j, k] + omega (S0[j, k] - G_j)) / d    (= A[j, k] / d when the scaled row k
//             is dropped), alpha / beta / rowmax updates in column order (Reimer.py:604,684),
//             threads m < k : Ut[m][k] = d_m thresh(omega Lu[k, m])  (Reimer.py:553-560, for the panel kernel).
// omega G ignores the reference's thresholding of tiny entries of the scaled row (Reimer.py:558).  A column falls
// back to the explicit thresholded dot product ("slow", exact w.r.t. the reference inside the block) whenever that
// could matter -- omega_k * min nonzero |Lu[k, .]| < min_abs_value_L -- or G could overflow (explosive regime,
// |Lu| > 1e100).  Outputs: d, omega, delta, clamped; Ut; hdr = (d, omega, drop, delta, clamped) of the panel.
// ------------------------------------------------------------------------------------------------
// Sblk / Ablk point at element (k0, k0) of the storage that holds S (lower part is read and written) and
// the permuted original A (upper part is read): the same array W on one GPU, S_loc / A_loc when distributed.
#define MDB_DIAG_THREADS(NB) (4 * (NB) + 32)
#define MDB_DIAG_SMEM_DOUBLES(NB) ((NB) * ((NB) + 4) + 9 * (NB))

template <int NB>
__global__ void __launch_bounds__(4 * NB + 32) k_diag_block(FactorDev F, int64_t k0, int kb, double* __restrict__ Sblk,
                                                            int64_t ldS, const double* __restrict__ Ablk, int64_t ldA) {
  // Compute warps: thread (j, h) = (tid >> 2, tid & 3), tid < 4 NB, sums the terms m = h (mod 4) of row j's dot
  // product; lane h == 0 carries the row's state (alpha, beta, rowmax, rowmin) and publishes it in shared memory.
  // Rule warp (the last one): lanes 0 and 1 evaluate the rule for column k from the published state of row k, so
  // the rule really runs beside the Gram sums instead of serialising with the lanes of its own warp.
  extern __shared__ double smem_d[];
  constexpr int LDT = NB + 4;   // 264 words: the 4 rows of a half warp land in 4 disjoint groups of 8 banks
  constexpr double BIG = 1e100;
  double* T = smem_d;           // T[r * LDT + c]: c < r: S0 -> Lu;  c >= r: A, then (r, c) <- d_r Lu[c, r]
  double* s_u = T + NB * LDT;   // NB (slow path)
  double* s_d = s_u + NB;       // NB
  double* s_alpha = s_d + NB;   // per row, published by lane (j, 0) after every column
  double* s_beta = s_alpha + NB;
  double* s_rowmax = s_beta + NB;
  double* s_rowmin = s_rowmax + NB;
  double* s_gamma = s_rowmin + NB;
  double* s_mB = s_gamma + NB;
  double* s_MB = s_mB + NB;
  __shared__ double s_dec[2][5];  // d, omega, drop, slow, 1 / d of column k (slot k & 1)
  __shared__ int s_big;
  __shared__ volatile int s_ready;  // rows 0 .. s_ready have published their state for their own column

  const int64_t b = blockIdx.z;
  Sblk += b * F.sW;
  Ablk += b * F.sW;
  const int64_t vo = b * F.sVec;
  double* Ut = F.Ut + b * F.sUt;
  double* hdr = F.hdr + b * F.sHdr;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool rule_warp = tid >= 4 * NB;
  const int j = tid >> 2, h = tid & 3;
  const bool owner = !rule_warp && (h == 0) && (j < kb);
  const bool exact = F.prm.rule == MD_RULE_EXACT;
  const double mavL = F.prm.min_abs_value_L;

  // whole block in flight at once (asynchronous 8-byte copies, coalesced per row):
  // strict lower part from S, diagonal and upper part from A
  if (!rule_warp) {
    const int c = tid & (NB - 1);
    if (c < kb)
      for (int r = tid / NB; r < kb; r += 4)
        mdb_cp_async8(&T[r * LDT + c], (c < r) ? &Sblk[r * ldS + c] : &Ablk[r * ldA + c]);
  }
  double alpha_j = 0, beta_j = 0, rowmax_j = 0, rowmin_j = INFINITY;
  if (owner) {
    alpha_j = F.alpha[vo + k0 + j];
    beta_j = F.beta[vo + k0 + j];
    rowmax_j = F.rowmax[vo + k0 + j];
    s_alpha[j] = alpha_j; s_beta[j] = beta_j; s_rowmax[j] = rowmax_j; s_rowmin[j] = rowmin_j;
    s_gamma[j] = F.gamma[vo + k0 + j];
    s_mB[j] = F.minB ? F.minB[vo + k0 + j] : F.prm.min_diag_B;
    s_MB[j] = F.maxB ? F.maxB[vo + k0 + j] : F.prm.max_diag_B;
  }
  if (tid == 0) { s_big = 0; s_ready = 0; }
  mdb_cp_async_wait_all();
  __syncthreads();
  if (owner && rowmax_j > BIG) s_big = 1;
  __syncthreads();

  const double* Trow = T + (rule_warp ? 0 : j) * LDT;
  for (int k = 0; k < kb; ++k) {
    const bool active = !rule_warp && (j > k && j < kb);
    double G = 0;
    if (rule_warp) {
      // ---------------- the rule for column k (lanes 0 and 1: d = lo and d = max_diag_D side by side) ----------------
      while (s_ready < k) { }  // row k has published alpha_k (it contains column k - 1)
      __syncwarp();
      const double al = s_alpha[k], be = s_beta[k], ga = s_gamma[k], mB = s_mB[k], MB = s_MB[k];
      MdBest cand;
      cand.d = cand.w = cand.f = 0;
      cand.have = 0;
      int fb = 0;
      bool uncl = true;
      double mD = 0;
      if (lane < 2 && !exact) {
        mD = md_effective_min_diag_D(F.prm, ga, mB, MB);
        uncl = md_unclamped(al, ga, mD, F.prm.max_diag_D, mB, MB, F.prm.min_abs_value_D);
        if (!uncl) {
          double dd;
          const bool en = md_d_candidate(lane, al, be, mD, F.prm.max_diag_D, mB, MB, F.prm.min_abs_value_D, &dd);
          cand = md_candidates_for_d(en, dd, al, be, ga, mB, MB, &fb);
        }
      }
      MdBest other;
      other.d = __shfl_sync(0xffffffffu, cand.d, 1);
      other.w = __shfl_sync(0xffffffffu, cand.w, 1);
      other.f = __shfl_sync(0xffffffffu, cand.f, 1);
      other.have = __shfl_sync(0xffffffffu, cand.have, 1);
      const int fb1 = __shfl_sync(0xffffffffu, fb, 1);
      if (lane == 0) {
        MdDecision dec;
        if (exact || uncl) {
          dec.d = ga - al; dec.omega = 1; dec.f = 0;
          dec.clamped = exact ? !(dec.d > 0) : 0;
        } else {
          dec = md_decision_from_best(md_assemble_minimal_change(cand, other, fb | fb1, al, be, ga, mD, F.prm.max_diag_D,
                                                                 mB, MB, F.prm.min_abs_value_D));
        }
        const int64_t K = k0 + k;
        const double delta = (dec.d - ga) + (dec.omega != 0 ? dec.omega * dec.omega * al : 0.0);  // :541-545
        F.d[vo + K] = dec.d;
        F.omega[vo + K] = dec.omega;
        F.delta[vo + K] = delta;
        F.clamped[vo + K] = (uint8_t)dec.clamped;
        if (exact && dec.clamped && F.info[b] == 0) F.info[b] = (int)(K + 1);
        const bool drop = (dec.omega == 0 || dec.omega * s_rowmax[k] < mavL);
        const bool slow = !drop && k > 0 && (dec.omega * s_rowmin[k] < mavL || s_big != 0);
        double* sd = s_dec[k & 1];
        sd[0] = dec.d;
        sd[1] = dec.omega;
        sd[2] = drop ? 1.0 : 0.0;
        sd[3] = slow ? 1.0 : 0.0;
        sd[4] = (dec.d != 0) ? 1.0 / dec.d : 0.0;   // the division leaves the dependency chain of the rows
        s_d[k] = dec.d;
        hdr[k] = dec.d;
        hdr[NB + k] = dec.omega;
        hdr[2 * NB + k] = drop ? 1.0 : 0.0;
        hdr[3 * NB + k] = delta;
        hdr[4 * NB + k] = (double)dec.clamped;
      }
    } else {
      // ---------------- the Gram dot products G_j = sum_{m < k-1} Lu[j, m] (d_m Lu[k, m]) ----------------
      double g0 = 0, g1 = 0;
      if (active) {
        const double* Tk = T + k;  // (m, k) holds d_m Lu[k, m]
        int m = h;
        for (; m + 4 < k - 1; m += 8) {
          g0 = fma(Trow[m], Tk[m * LDT], g0);
          g1 = fma(Trow[m + 4], Tk[(m + 4) * LDT], g1);
        }
        if (m < k - 1) g0 = fma(Trow[m], Tk[m * LDT], g0);
      }
      G = g0 + g1;
      // every lane takes part in the shuffles (rows of one warp become active at different k)
      G += __shfl_xor_sync(0xffffffffu, G, 1);
      G += __shfl_xor_sync(0xffffffffu, G, 2);
    }
    __syncthreads();

    // ---------------- column k ----------------
    const double dk = s_dec[k & 1][0], wk = s_dec[k & 1][1], inv_dk = s_dec[k & 1][4];
    const bool drop = s_dec[k & 1][2] != 0.0, slow = s_dec[k & 1][3] != 0.0;
    double a = 0, s0k = 0;
    if (active) {
      a = T[k * LDT + j];   // A[j, k]; read before thread (j, 0) overwrites it below
      s0k = Trow[k];
      if (k >= 1) G = fma(Trow[k - 1], T[(k - 1) * LDT + k], G);  // the term m = k - 1, published just before the barrier
    }
    if (tid < k) {  // the scaled, thresholded row k times D: for the panel kernel, and for the slow path
      double lt = wk * T[k * LDT + tid];
      if (wk == 0 || fabs(lt) < mavL) lt = 0;
      const double u = s_d[tid] * lt;
      Ut[tid * NB + k] = u;
      if (slow) s_u[tid] = u;
    }
    double acc = 0;
    if (slow) {  // uniform over the block
      __syncthreads();
      if (active) {
        double a0 = 0, a1 = 0;
        int m = h;
        for (; m + 4 < k; m += 8) {
          a0 = fma(Trow[m], s_u[m], a0);
          a1 = fma(Trow[m + 4], s_u[m + 4], a1);
        }
        if (m < k) a0 = fma(Trow[m], s_u[m], a0);
        acc = a0 + a1;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    }
    __syncwarp();  // all four lanes of a row have read T[k][j] and T[j][k]
    if (active && h == 0) {
      // x = r * (1 / d_k): the reciprocal comes from the rule warp (<= 1 ulp from r / d_k)
      double x = 0;
      if (dk != 0) {
        if (drop) x = a * inv_dk;
        else if (slow) x = (((1.0 - wk) * a + wk * s0k) - acc) * inv_dk;
        else x = fma(wk, s0k - G, (1.0 - wk) * a) * inv_dk;
      }
      T[j * LDT + k] = x;
      T[k * LDT + j] = x * dk;        // d_k Lu[j, k] at the transposed position
      alpha_j += x * x * dk;          // Reimer.py:679-684
      beta_j += 2.0 * a * a;          // Reimer.py:602-604
      const double ax = fabs(x);
      rowmax_j = fmax(rowmax_j, ax);
      if (ax != 0) rowmin_j = fmin(rowmin_j, ax);
      if (ax > BIG) s_big = 1;
      if (j == k + 1) {               // the next column's row: hand its state to the rule warp right away
        s_alpha[j] = alpha_j; s_beta[j] = beta_j; s_rowmax[j] = rowmax_j; s_rowmin[j] = rowmin_j;
        __threadfence_block();
        s_ready = k + 1;
      }
    }
    __syncwarp();  // x[j, k] is visible to the other lanes of the row before they use it in the next Gram sum
    // no block barrier here: what the next Gram sums read was published before the barrier above; the entries written
    // just now are first read after the next barrier (the m = k term, s_dec) or through s_ready (the rule warp)
  }
  __syncthreads();
  if (owner) {
    F.alpha[vo + k0 + j] = alpha_j;
    F.beta[vo + k0 + j] = beta_j;
    F.rowmax[vo + k0 + j] = rowmax_j;
  }
  // strict lower part of the block back to S (unscaled rows Lu)
  if (!rule_warp) {
    const int c = tid & (NB - 1);
    for (int r = tid / NB; r < kb; r += 4)
      if (c < r) Sblk[r * ldS + c] = T[r * LDT + c];
  }
}

#define MDB_PANEL_LDX(R) ((R) + 8)
#define MDB_PANEL_SMEM_DOUBLES(NB, R) ((NB) * MDB_PANEL_LDX(R) + 4 * (NB) + 2 * (NB) * 8)

template <int NB, int R>
__global__ void __launch_bounds__(R) k_panel(FactorDev F, int64_t k0, int64_t row_begin, int64_t n_rows,
                                             const double* __restrict__ Asrc, int64_t a_rs, int64_t a_cs,
                                             double* __restrict__ Wcol, int64_t w_ld,
                                             double* __restrict__ PX, double* __restrict__ PY, int64_t p_row0,
                                             int64_t p_ld) {
  // Wcol: pointer such that S0[j, k0 + c] lives at Wcol[j * w_ld + c]  (j = global row)
  // PX / PY: row (j - p_row0) of the operand buffers, leading dimension p_ld (>= NB: the buffers hold the
  // panels of one outer block side by side so that the bulk trailing update contracts over all of them)
  extern __shared__ double smem_d[];
  constexpr int LDX = MDB_PANEL_LDX(R);  // 72: the DMMA fragment loads below then take the minimal two wavefronts
  double* xs = smem_d;          // xs[c * LDX + r]
  double* s_d = xs + NB * LDX;  // NB
  double* s_w = s_d + NB;       // NB
  double* s_drop = s_w + NB;    // NB
  double* s_inv = s_drop + NB;  // NB: 1 / d (0 where d == 0): x = r * (1 / d), <= 1 ulp from r / d, one division per column
  double* us = s_inv + NB;      // 2 x (NB x 8): double-buffered 8-column slices of U (L1 is too small to hold U
                                // next to the X tiles, so U is staged through shared memory with cp.async)

  const int64_t b = blockIdx.z;
  const int64_t vo = b * F.sVec;
  const double* Ut = F.Ut + b * F.sUt;
  const double* hdr = F.hdr + b * F.sHdr;
  Asrc += b * F.sW;
  Wcol += b * F.sW;
  PX += b * F.sPanel;
  PY += b * F.sPanel;
  const int tid = threadIdx.x;
  const int64_t jbase = row_begin + (int64_t)blockIdx.x * R;
  const int64_t jend = row_begin + n_rows;
  const int64_t j = jbase + tid;
  const bool valid = j < jend;

  // slice sb = columns [8 sb, 8 sb + 8) of U, rows m < 8 sb + 8 (rows m >= column hold stale data and are not used)
  auto load_slice = [&](int sb, int buf) {
    const int c0 = sb * 8;
    const int chunks = (c0 + 8) * 4;  // 16-byte chunks
    for (int idx = tid; idx < chunks; idx += R) {
      const int m = idx >> 2, q = idx & 3;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(us + buf * (NB * 8) + m * 8 + q * 2);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ut + m * NB + c0 + q * 2));
    }
  };

  for (int c = tid; c < NB; c += R) {
    const double dc = hdr[c];
    s_d[c] = dc;
    s_inv[c] = (dc != 0) ? 1.0 / dc : 0.0;
    s_w[c] = hdr[NB + c];
    s_drop[c] = hdr[2 * NB + c];
  }
  load_slice(0, 0);
  // S0 tile, transposed into shared memory: coalesced row segments, all rows in flight at once
  for (int rr = 0; rr < R; ++rr) {
    const int64_t jj = jbase + rr;
    if (jj < jend)
      for (int c = tid; c < NB; c += R) mdb_cp_async8(&xs[c * LDX + rr], &Wcol[jj * w_ld + c]);
  }
  mdb_cp_async_wait_all();
  __syncthreads();

  double alpha_j = 0, beta_j = 0, rowmax_j = 0;
  if (valid) {
    alpha_j = F.alpha[vo + j];
    beta_j = F.beta[vo + j];
    rowmax_j = F.rowmax[vo + j];
    // mix with the original A column entries; beta in column order
    const double* Aj = Asrc + j * a_rs + k0 * a_cs;
    for (int cb = 0; cb < NB; cb += 32) {
      double av[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) av[i] = Aj[(int64_t)(cb + i) * a_cs];  // 32 independent loads in flight
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = cb + i;
        const double a = av[i];
        const double w = s_w[c];
        double r = a;
        if (s_drop[c] == 0.0) r = (1.0 - w) * a + w * xs[c * LDX + tid];
        xs[c * LDX + tid] = r;
        beta_j += 2.0 * a * a;
      }
    }
  }
  // forward substitution, 8 columns at a time; the next U slice streams in while this one is used
  for (int sb = 0; sb < NB / 8; ++sb) {
    const int c0 = sb * 8;
    if (sb + 1 < NB / 8) load_slice(sb + 1, (sb + 1) & 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    {
      // acc = R - X U on the FP64 tensor cores: each warp owns its 32 rows (two 16 x 8 tiles), K = c0.
      // mma.m16n8k4: a0 = A[g][t4], a1 = A[g+8][t4], b0 = B[t4][g], c = C[g | g+8][2 t4 | 2 t4 + 1].
      const double* U = us + (sb & 1) * (NB * 8);
      const int lane = tid & 31, g = lane >> 2, t4 = lane & 3;
      const int rb = (tid & ~31);  // first row of this warp inside the CTA tile
      if (jbase + rb < jend && c0 > 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = rb + 16 * mt + g;
          double cfrag[4];
          cfrag[0] = xs[(c0 + 2 * t4) * LDX + r0];
          cfrag[1] = xs[(c0 + 2 * t4 + 1) * LDX + r0];
          cfrag[2] = xs[(c0 + 2 * t4) * LDX + r0 + 8];
          cfrag[3] = xs[(c0 + 2 * t4 + 1) * LDX + r0 + 8];
#pragma unroll 4
          for (int k = 0; k < c0; k += 4) {
            const double a0 = xs[(k + t4) * LDX + r0];
            const double a1 = xs[(k + t4) * LDX + r0 + 8];
            const double b0 = -U[(k + t4) * 8 + g];
            asm volatile(
                "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                : "+d"(cfrag[0]), "+d"(cfrag[1]), "+d"(cfrag[2]), "+d"(cfrag[3])
                : "d"(a0), "d"(a1), "d"(b0));
          }
          xs[(c0 + 2 * t4) * LDX + r0] = cfrag[0];
          xs[(c0 + 2 * t4 + 1) * LDX + r0] = cfrag[1];
          xs[(c0 + 2 * t4) * LDX + r0 + 8] = cfrag[2];
          xs[(c0 + 2 * t4 + 1) * LDX + r0 + 8] = cfrag[3];
        }
      }
      __syncwarp();
      if (valid) {
        // the 8 x 8 triangle of the slice, one thread per row
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = xs[(c0 + i) * LDX + tid];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = c0 + i;
          const double dk = s_d[c];
          const double x = (dk != 0) ? acc[i] * s_inv[c] : 0.0;
          xs[c * LDX + tid] = x;
          alpha_j += x * x * dk;          // Reimer.py:679-684
          rowmax_j = fmax(rowmax_j, fabs(x));
#pragma unroll
          for (int i2 = i + 1; i2 < 8; ++i2) acc[i2] -= x * U[c * 8 + i2];
        }
      }
      __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
  }
  if (valid) {
    F.alpha[vo + j] = alpha_j;
    F.beta[vo + j] = beta_j;
    F.rowmax[vo + j] = rowmax_j;
  }
  // write back (coalesced rows): Lu into W, X and Y = -X d into the compact panel buffers
  for (int rr = 0; rr < R; ++rr) {
    const int64_t jj = jbase + rr;
    if (jj < jend) {
      const int64_t pr = jj - p_row0;
      for (int c = tid; c < NB; c += R) {
        const double x = xs[c * LDX + rr];
        Wcol[jj * w_ld + c] = x;
        PX[pr * p_ld + c] = x;
        PY[pr * p_ld + c] = -x * s_d[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Finalize in place: strict lower L[j, m] = thresh(omega_j Lu[j, m]) (Reimer.py:553-560), diagonal 1
// (LDL, Reimer.py:689-692), d (LDL_compressed, decompositions.py:875-882) or sqrt(d) with columns
// scaled by sqrt(d_m) (LL, decompositions.py:898-918); upper triangle zero (Reimer.py:698-701).
// HBM bound: 16 n^2 bytes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finalize(int64_t n, double* __restrict__ W, int64_t ldw, int64_t sW,
                                                  const double* __restrict__ d, const double* __restrict__ omega,
                                                  int64_t sVec, double mavL, int out_type, double* __restrict__ out,
                                                  int64_t ldo, int64_t sOut, int64_t row0) {
  const int64_t b = blockIdx.z;
  const int64_t i = row0 + blockIdx.x;   // rows [row0, row0 + gridDim.x)
  const double* Wrow = W + b * sW + i * ldw;
  double* Orow = out + b * sOut + i * ldo;
  const double w = omega[b * sVec + i];
  const double* db = d + b * sVec;
  const int64_t j0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t j = j0 + 256 * e;
    if (j < n) {
      double v;
      if (j < i) {
        v = w * Wrow[j];
        if (w == 0 || fabs(v) < mavL) v = 0;
        if (out_type == MDB_TYPE_LL) v *= sqrt(db[j]);
      } else if (j == i) {
        v = (out_type == MDB_TYPE_LDL) ? 1.0 : (out_type == MDB_TYPE_LL ? sqrt(db[i]) : db[i]);
      } else {
        v = 0.0;
      }
      Orow[j] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// B of positive_semidefinite_matrix (Reimer.py:899-992) from the UN-finalised state (upper triangle of W = the
// permuted A, strict lower = unscaled rows Lu), in ORIGINAL index order:
//   B[i, i] = A[i, i] + delta_i clipped to [min_diag_B_i, max_diag_B_i]                       :924-945
//   B[i, j] = A[i, j] * omega[later of (i, j)]          when d[earlier] != 0                  :947-957
//           = 0                                          when omega[later] == 0
//           = sum_{k < earlier} L[q_i, k] d_k L[q_j, k]  otherwise (rows of the scaled, thresholded L)  :958-975
// q = inverse permutation (position of an original index).  HBM bound: 8 n^2 written coalesced, reads are
// 8-byte gathers from the upper triangle.  grid = (n rows, column chunks of 1024).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_psd_matrix(int64_t n, const double* __restrict__ W, int64_t ldw,
                                                    const double* __restrict__ gamma, const double* __restrict__ delta,
                                                    const double* __restrict__ omega, const double* __restrict__ d,
                                                    const int64_t* __restrict__ q, const double* __restrict__ minB,
                                                    const double* __restrict__ maxB, int64_t strideB, double mavL,
                                                    double* __restrict__ B, int64_t ldb) {
  const int64_t i = blockIdx.x;
  const int64_t qi = q ? q[i] : i;
  const int64_t j0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t j = j0 + 256 * e;
    if (j >= n) continue;
    double v;
    if (j == i) {
      v = gamma[qi] + delta[qi];
      if (minB) { const double m = minB[i * strideB]; if (v < m) v = m; }
      if (maxB) { const double m = maxB[i * strideB]; if (v > m) v = m; }
    } else {
      const int64_t qj = q ? q[j] : j;
      const int64_t lo = qi < qj ? qi : qj, hi = qi < qj ? qj : qi;
      const double om = omega[hi];
      if (d[lo] != 0) {
        v = W[lo * ldw + hi] * om;
      } else if (om == 0) {
        v = 0.0;
      } else {  // rare: inner product of two rows of the scaled, thresholded L over the columns before `lo`
        const double wi = omega[qi], wj = omega[qj];
        const double* Li = W + qi * ldw;
        const double* Lj = W + qj * ldw;
        double acc = 0;
        for (int64_t k = 0; k < lo; ++k) {
          double a = wi * Li[k], bb = wj * Lj[k];
          if (wi == 0 || fabs(a) < mavL) a = 0;
          if (wj == 0 || fabs(bb) < mavL) bb = 0;
          acc += a * d[k] * bb;
        }
        v = acc;
      }
    }
    B[i * ldb + j] = v;
  }
}

// q[p[i]] = i
__global__ void k_invert_permutation(int64_t n, const int64_t* __restrict__ p, int64_t* __restrict__ q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) q[p[i]] = i;
}

// position-indexed -> original-indexed scatter of omega / delta (Reimer.py:538,545): out[p[i]] = in[i]
__global__ void k_scatter_by_p(int64_t n, const int64_t* __restrict__ p, const double* __restrict__ in,
                               double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[p ? p[i] : i] = in[i];
}

__global__ void k_gather_by_p(int64_t n, const int64_t* __restrict__ p, const double* __restrict__ in, int64_t stride,
                              double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[(p ? p[i] : i) * stride];
}

// Distributed finalize: the local column slab S_loc (n x ld_loc) in place.  Local column lc belongs to global
// column gc = ((lc / NB) * world + rank) * NB + lc % NB.
__global__ void __launch_bounds__(256) k_finalize_cols(int64_t n, double* __restrict__ S, int64_t ld, int64_t ncols,
                                                       int rank, int world, const double* __restrict__ omega,
                                                       double mavL) {
  const int64_t i = blockIdx.x;
  const double w = omega[i];
  const int64_t l0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t lc = l0 + 256 * e;
    if (lc < ncols) {
      const int64_t gc = ((lc / MDB_NB) * world + rank) * MDB_NB + lc % MDB_NB;
      double v;
      if (gc < i) {
        v = w * S[i * ld + lc];
        if (w == 0 || fabs(v) < mavL) v = 0;
      } else {
        v = (gc == i) ? 1.0 : 0.0;
      }
      S[i * ld + lc] = v;
    }
  }
}
