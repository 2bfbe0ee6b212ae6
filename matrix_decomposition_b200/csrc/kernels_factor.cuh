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
// memory.  Sequential over the columns because the decision for column k needs alpha_k, which contains
// Lu[k, k0:k]^2 d from this very block: the kernel is one dependency chain, and everything is arranged to keep the
// per-column part of it short (round 2: 1750 -> ~X cycles per column).
//
// Threads: 16 compute warps, thread (j, h) = (tid >> 2, tid & 3): the four lanes of a row share its dot products,
// lanes h = 0 and 1 both carry the row's state (alpha, beta, rowmax, rowmin) redundantly -- identical arithmetic --
// so that the row itself can decide its column without a hand-over: the unclamped test, MD_RULE_EXACT and the upper
// clamp (md_upper_clamp_shortcut) are a few comparisons in lane h = 0; a column that needs the two cubic solves
// (Reimer.py:95-115) evaluates the candidates of d = lo and d = max_diag_D side by side in lanes h = 0 / 1
// (mdb_rule_full_pair, one instruction stream, merged with the reference's key).  Warp 16 is a helper (below).
// Per column k, compute warps only (named barrier, the helper does not take part):
//   phase A   rows j > k : G_j = sum_{sb0 <= m < k-1} Lu[j, m] d_m Lu[k, m]   -- the dot product with the UNscaled row
//             k (d_m Lu[k, m] sits at the transposed position (m, k) of the tile: the A entry stored there is dead
//             once column m has been processed).  Does not need the decision for column k.
//   barrier
//   phase B   x[j, k] = ((1 - omega) A[j, k] + omega (S'[j, k] - G_j)) / d    (= A[j, k] / d when the scaled row k
//             is dropped), alpha / beta / rowmax updates in column order (Reimer.py:604,684); row k + 1 then decides
//             column k + 1 at once (d, omega, 1 / d, flags into the double-buffered s_dec).
// Two-level inside the block: sub-blocks of SB = 32 columns.  The dot products only run over the columns of the
// current sub-block (sb0 = its first column): when a sub-block is complete, its contribution
// S'[j, c] -= sum_{m in sub-block} Lu[j, m] d_m Lu[c, m]  to every later column c of the block is applied at once on
// the FP64 tensor cores (mma.m16n8k4 from shared memory).
// omega G ignores the reference's thresholding of tiny entries of the scaled row (Reimer.py:558).  A column falls
// back to the explicit thresholded dot product over ALL earlier columns of the block and the block-entry value of S
// (re-read from global memory, which this kernel only overwrites at its end) -- "slow", exact w.r.t. the reference
// inside the block -- whenever that could matter: omega_k * min nonzero |Lu[k, .]| < min_abs_value_L, or the
// products could overflow (explosive regime, |Lu| > 1e100; from then on S' is not used any more).
// Helper warp: trails the column loop (polls s_progress) and forms, row by row, the inverse of the scaled unit
// triangle Ltilde[c][m] = thresh(omega_c Lu[c, m]) of each sub-block; the panel kernel k_panel_gemm multiplies by
// these 32 x 32 inverses instead of substituting column by column.  Vinv[4096] = 1 tells it not to (an entry of an
// inverse above 1e8 or not finite, |Lu| > 1e100 anywhere, or a zero d): it then substitutes.
// Outputs: d, omega, delta, clamped; Ut[m][k] = d_m thresh(omega_k Lu[k, m]) (Reimer.py:553-560); hdr = (d, omega,
// drop, delta, clamped) of the panel; Vinv.
// ------------------------------------------------------------------------------------------------
// Sblk / Ablk point at element (k0, k0) of the storage that holds S (lower part is read and written) and
// the permuted original A (upper part is read): the same array W on one GPU, S_loc / A_loc when distributed.
#define MDB_DIAG_THREADS(NB) (4 * (NB) + 32)
#define MDB_DIAG_SMEM_DOUBLES(NB) ((NB) * ((NB) + 4) + 3 * (NB) + 8 + 32 * 33 + 32)
#define MDB_DIAG_SB 32

// S'[j, c] -= sum_{m0 <= m < m0 + SB} T[j][m] T[m][c]  for c0 <= c < j < kb (c0 = m0 + SB), tiles of 16 x 8 dealt
// round-robin to the 16 compute warps.  mma.m16n8k4: a0 = A[g][t4], a1 = A[g+8][t4], b0 = B[t4][g],
// c = C[g | g+8][2 t4 | 2 t4 + 1].
template <int NB>
__device__ __forceinline__ void mdb_diag_subblock_update(double* __restrict__ T, int m0, int kb, int warp, int lane) {
  constexpr int LDT = NB + 4;
  constexpr int SB = MDB_DIAG_SB;
  const int c0 = m0 + SB;
  const int g = lane >> 2, t4 = lane & 3;
  const int n_rt = (kb - c0 + 15) >> 4, n_ct_all = (kb - c0 + 7) >> 3;
  int t = 0;
  for (int ri = 0; ri < n_rt; ++ri) {
    const int r0 = c0 + 16 * ri;
    const int n_ct = (2 * ri + 2 < n_ct_all) ? 2 * ri + 2 : n_ct_all;   // tiles that reach below the diagonal
    for (int ci = 0; ci < n_ct; ++ci, ++t) {
      if ((t & 15) != warp) continue;
      const int cc0 = c0 + 8 * ci;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* Ap0 = T + (r0 + g) * LDT + m0 + t4;
      const double* Ap1 = Ap0 + 8 * LDT;
      const double* Bp = T + (m0 + t4) * LDT + cc0 + g;
#pragma unroll
      for (int kk = 0; kk < SB; kk += 4) {
        const double a0 = Ap0[kk], a1 = Ap1[kk], b0 = Bp[kk * LDT];
        asm volatile(
            "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
            : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3])
            : "d"(a0), "d"(a1), "d"(b0));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = r0 + g + ((e & 2) ? 8 : 0), col = cc0 + 2 * t4 + (e & 1);
        if (row < kb && col < row) T[row * LDT + col] -= acc[e];
      }
    }
  }
}

// The clamped branch with both cubic solves (Reimer.py:89-135), lanes h = 0 / 1 of one row group: lane h evaluates the
// candidates of d = lo (h = 0) or d = max_diag_D (h = 1), the pair exchanges them and both lanes assemble the same
// decision.  Kept out of line: it is ~1300 instructions and the column loop calls it rarely.
__device__ __noinline__ MdDecision mdb_rule_full_pair(MdRuleParams prm, int h, unsigned pairmask, double al, double be,
                                                      double ga, double mD, double mB, double MB) {
  double dd;
  int fb = 0;
  const bool en = md_d_candidate(h, al, be, mD, prm.max_diag_D, mB, MB, prm.min_abs_value_D, &dd);
  const MdBest mine = md_candidates_for_d(en, dd, al, be, ga, mB, MB, &fb);
  MdBest other;
  other.d = __shfl_xor_sync(pairmask, mine.d, 1);
  other.w = __shfl_xor_sync(pairmask, mine.w, 1);
  other.f = __shfl_xor_sync(pairmask, mine.f, 1);
  other.have = __shfl_xor_sync(pairmask, mine.have, 1);
  const int fb_other = __shfl_xor_sync(pairmask, fb, 1);
  const MdBest from_lo = (h == 0) ? mine : other, from_max = (h == 0) ? other : mine;
  return md_decision_from_best(md_assemble_minimal_change(from_lo, from_max, fb | fb_other, al, be, ga, mD, prm.max_diag_D,
                                                          mB, MB, prm.min_abs_value_D));
}

template <int NB>
__global__ void __launch_bounds__(4 * NB + 32) k_diag_block(FactorDev F, int64_t k0, int kb, double* __restrict__ Sblk,
                                                            int64_t ldS, const double* __restrict__ Ablk, int64_t ldA) {
  extern __shared__ double smem_d[];
  constexpr int LDT = NB + 4;   // 264 words: the 4 rows of a half warp land in 4 disjoint groups of 8 banks
  constexpr int SB = MDB_DIAG_SB;
  constexpr int NCOMPUTE = 4 * NB;
  constexpr double BIG = 1e100;
  double* T = smem_d;           // T[r * LDT + c]: c < r: S0 -> S' -> Lu;  c >= r: A, then (r, c) <- d_r Lu[c, r]
  double* s_u = T + NB * LDT;   // NB (slow path)
  double* s_d = s_u + NB;       // NB: d of the finished columns
  double* s_w = s_d + NB;       // NB: omega of the finished columns
  double* s_dec = s_w + NB;     // 2 x 4: d, omega, 1 / d, flags (1 = drop, 2 = slow) of column k in slot k & 1
  double* s_inv = s_dec + 8;    // 32 x 33: inverse of the current sub-block's scaled unit triangle (helper warp)
  double* s_lt = s_inv + 32 * 33;  // 32: the scaled, thresholded row the helper is working on
  __shared__ int s_bigf[2], s_unsafe;   // s_bigf[k & 1]: some |Lu| > 1e100 seen up to phase B of a column of that parity
  __shared__ volatile int s_progress;  // columns whose decision and row are final (the helper polls it)

  const int64_t b = blockIdx.z;
  Sblk += b * F.sW;
  Ablk += b * F.sW;
  const int64_t vo = b * F.sVec;
  double* Ut = F.Ut + b * F.sUt;
  double* hdr = F.hdr + b * F.sHdr;
  double* Vout = F.Vinv + b * F.sV;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool helper = tid >= NCOMPUTE;
  const int j = tid >> 2, h = tid & 3;
  const bool keeper = !helper && (h < 2) && (j < kb);   // lanes that carry the row's state
  const bool exact = F.prm.rule == MD_RULE_EXACT;
  const double mavL = F.prm.min_abs_value_L;
  const unsigned pairmask = 0x3u << (lane & 28);
#define MDB_DIAG_BAR() asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory")

  // whole block in flight at once (asynchronous 8-byte copies, coalesced per row):
  // strict lower part from S, diagonal and upper part from A
  if (!helper) {
    const int c = tid & (NB - 1);
    if (c < kb)
      for (int r = tid / NB; r < kb; r += 4)
        mdb_cp_async8(&T[r * LDT + c], (c < r) ? &Sblk[r * ldS + c] : &Ablk[r * ldA + c]);
  }
  double alpha_j = 0, beta_j = 0, rowmax_j = 0, rowmin_j = INFINITY, gamma_j = 0, mB_j = 0, MB_j = 0;
  if (keeper) {
    alpha_j = F.alpha[vo + k0 + j];
    beta_j = F.beta[vo + k0 + j];
    rowmax_j = F.rowmax[vo + k0 + j];
    gamma_j = F.gamma[vo + k0 + j];
    mB_j = F.minB ? F.minB[vo + k0 + j] : F.prm.min_diag_B;
    MB_j = F.maxB ? F.maxB[vo + k0 + j] : F.prm.max_diag_B;
  }
  if (tid == 0) { s_bigf[0] = s_bigf[1] = 0; s_unsafe = 0; s_progress = 0; }
  mdb_cp_async_wait_all();
  __syncthreads();
  if (keeper && h == 0 && rowmax_j > BIG) s_bigf[0] = s_bigf[1] = 1;
  __syncthreads();

  if (helper) {
    // ---------------- inverses of the scaled unit triangles, one row per finished column ----------------
    double vmax = 0;
    for (int k = 0; k < kb; ++k) {
      const int kk = k & (SB - 1), sb0 = k - kk;
      while (s_progress <= k) __nanosleep(40);
      __threadfence_block();
      const double wk = s_w[k];
      double lt = 0;
      if (lane < kk) {
        lt = wk * T[k * LDT + sb0 + lane];
        if (wk == 0 || fabs(lt) < mavL) lt = 0;
      }
      s_lt[lane] = lt;
      __syncwarp();
      double x = (lane == kk) ? 1.0 : 0.0;
      if (lane < kk) {
        double a0 = 0, a1 = 0;
        int m = lane;
        for (; m + 1 < kk; m += 2) {
          a0 = fma(s_lt[m], s_inv[m * 33 + lane], a0);
          a1 = fma(s_lt[m + 1], s_inv[(m + 1) * 33 + lane], a1);
        }
        if (m < kk) a0 = fma(s_lt[m], s_inv[m * 33 + lane], a0);
        x = -(a0 + a1);
      }
      s_inv[kk * 33 + lane] = x;
      if (!(fabs(x) <= 1e8)) vmax = INFINITY;   // also catches NaN
      __syncwarp();
      if (kk == SB - 1 || k == kb - 1) {
        double* V = Vout + (sb0 / SB) * (SB * SB);
        for (int r = 0; r < SB; ++r) V[r * SB + lane] = (r <= kk) ? s_inv[r * 33 + lane] : ((r == lane) ? 1.0 : 0.0);
        __syncwarp();
      }
    }
    const unsigned bad = __ballot_sync(0xffffffffu, vmax != 0.0);
    __syncthreads();
    if (lane == 0) Vout[4 * SB * SB] = (bad != 0 || s_bigf[0] != 0 || s_bigf[1] != 0 || s_unsafe != 0) ? 1.0 : 0.0;
    return;
  }

  // decision of column kc by the two keeper lanes of row kc (their state is complete); lane h == 0 publishes it
  auto decide = [&](int kc, bool big) {
    double d, w;
    int clamped;
    if (exact) {
      d = gamma_j - alpha_j; w = 1.0;
      clamped = !(d > 0);
    } else {
      const double mD = md_effective_min_diag_D(F.prm, gamma_j, mB_j, MB_j);
      MdBest quick;
      if (md_unclamped(alpha_j, gamma_j, mD, F.prm.max_diag_D, mB_j, MB_j, F.prm.min_abs_value_D)) {
        d = gamma_j - alpha_j; w = 1.0; clamped = 0;
      } else if (md_upper_clamp_shortcut(alpha_j, beta_j, gamma_j, mD, F.prm.max_diag_D, mB_j, MB_j, F.prm.min_abs_value_D,
                                         &quick)) {
        d = quick.d; w = quick.w; clamped = 1;
      } else {
        const MdDecision dec = mdb_rule_full_pair(F.prm, h, pairmask, alpha_j, beta_j, gamma_j, mD, mB_j, MB_j);
        d = dec.d; w = dec.omega; clamped = 1;
      }
    }
    if (h == 0) {
      const bool drop = (w == 0 || w * rowmax_j < mavL);
      const bool slow = !drop && kc > 0 && (w * rowmin_j < mavL || big);
      double* sd = s_dec + 4 * (kc & 1);
      sd[0] = d;
      sd[1] = w;
      sd[2] = (d != 0) ? 1.0 / d : 0.0;
      sd[3] = (drop ? 1.0 : 0.0) + (slow ? 2.0 : 0.0);
      s_d[kc] = d;
      s_w[kc] = w;
      if (d == 0) s_unsafe = 1;
      // (global results after the shared ones: nothing on the chip waits for them)
      const int64_t K = k0 + kc;
      F.d[vo + K] = d;
      F.omega[vo + K] = w;
      const double delta = (d - gamma_j) + (w != 0 ? w * w * alpha_j : 0.0);  // :541-545
      F.delta[vo + K] = delta;
      F.clamped[vo + K] = (uint8_t)clamped;
      if (exact && clamped && F.info[b] == 0) F.info[b] = (int)(K + 1);
      hdr[kc] = d;
      hdr[NB + kc] = w;
      hdr[2 * NB + kc] = drop ? 1.0 : 0.0;
      hdr[3 * NB + kc] = delta;
      hdr[4 * NB + kc] = (double)clamped;
    }
  };

  bool big = s_bigf[0] != 0;   // sticky; refreshed after every column barrier from the flag of the previous column's parity
  if (keeper && j == 0) decide(0, big);
  const double* Trow = T + j * LDT;
  for (int k = 0; k < kb; ++k) {
    const int kk = k & (SB - 1), sb0 = k - kk;
    if (kk == 0 && k > 0) {
      MDB_DIAG_BAR();   // every x of the finished sub-block is in shared memory
      mdb_diag_subblock_update<NB>(T, sb0 - SB, kb, tid >> 5, lane);
      // (made visible to the readers of column k by the column barrier below; the Gram sum of a sub-block's first
      // column is empty)
    }
    const bool active = (j > k && j < kb);
    // ---------------- phase A: G_j = sum_{sb0 <= m < k-1} Lu[j, m] (d_m Lu[k, m]), lane h takes m = sb0 + h + 4 i ----------------
    double a = 0, g0 = 0, g1 = 0;
    if (active) {
      a = T[k * LDT + j];   // A[j, k]; lane (j, 0) overwrites it in phase B
      const double* Tr = Trow + sb0 + h;
      const double* Tk = T + (sb0 + h) * LDT + k;   // (m, k) holds d_m Lu[k, m]
      const int lim = kk - 1 - h;                   // terms 4 i < lim
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        if (4 * i < lim) g0 = fma(Tr[4 * i], Tk[4 * i * LDT], g0);
        if (4 * i + 4 < lim) g1 = fma(Tr[4 * i + 4], Tk[(4 * i + 4) * LDT], g1);
      }
    }
    double G = g0 + g1;
    // every lane takes part in the shuffles (rows of one warp become active at different k)
    G += __shfl_xor_sync(0xffffffffu, G, 1);
    G += __shfl_xor_sync(0xffffffffu, G, 2);
    MDB_DIAG_BAR();
    if (tid == 0) s_progress = k + 1;

    // ---------------- phase B: column k ----------------
    const double2 dec01 = *reinterpret_cast<const double2*>(s_dec + 4 * (k & 1));
    const double2 dec23 = *reinterpret_cast<const double2*>(s_dec + 4 * (k & 1) + 2);
    const double dk = dec01.x, wk = dec01.y, inv_dk = dec23.x;
    const int flags = (int)dec23.y;
    const bool drop = (flags & 1) != 0, slow = (flags & 2) != 0;
    big = big || (s_bigf[(k + 1) & 1] != 0);   // written in phase B of column k - 1, complete since the barrier: the same
                                               // value in every thread (the flag of this column's parity is being written now)
    double s0k = 0;
    if (active) {
      s0k = Trow[k];        // S'[j, k]: the block-entry value minus the finished sub-blocks
      if (kk >= 1) G = fma(Trow[k - 1], T[(k - 1) * LDT + k], G);  // the term m = k - 1, published just before the barrier
    }
    double acc = 0;
    if (slow) {  // uniform over the block
      if (tid < k) {  // the scaled, thresholded row k times D
        double lt = wk * T[k * LDT + tid];
        if (wk == 0 || fabs(lt) < mavL) lt = 0;
        s_u[tid] = s_d[tid] * lt;
      }
      if (active) s0k = Sblk[j * ldS + k];   // the block-entry value of S (global memory is still untouched)
      MDB_DIAG_BAR();
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
    __syncwarp();  // all four lanes of a row have read T[k][j] and T[j][k] before lane (j, 0) overwrites them
    if (active && h < 2) {
      // x = r * (1 / d_k): the reciprocal comes with the decision (<= 1 ulp from r / d_k)
      double x = 0;
      if (dk != 0) {
        if (drop) x = a * inv_dk;
        else if (slow) x = (((1.0 - wk) * a + wk * s0k) - acc) * inv_dk;
        else x = fma(wk, s0k - G, (1.0 - wk) * a) * inv_dk;
      }
      if (h == 0) {
        T[j * LDT + k] = x;
        T[k * LDT + j] = x * dk;      // d_k Lu[j, k] at the transposed position
      }
      alpha_j += x * x * dk;          // Reimer.py:679-684
      beta_j += 2.0 * a * a;          // Reimer.py:602-604
      const double ax = fabs(x);
      rowmax_j = fmax(rowmax_j, ax);
      if (ax != 0) rowmin_j = fmin(rowmin_j, ax);
      if (ax > BIG && h == 0) s_bigf[k & 1] = 1;
      if (j == k + 1) decide(k + 1, big || rowmax_j > BIG);   // the next column's row: its state is complete now
    }
    __syncwarp();  // x[j, k] is visible to the other lanes of the row before they use it in the next Gram sum
    // no block barrier here: what the next Gram sums read was published before the barrier above; the entries written
    // just now are first read after the next barrier
  }
  __syncthreads();
  if (keeper && h == 0) {
    F.alpha[vo + k0 + j] = alpha_j;
    F.beta[vo + k0 + j] = beta_j;
    F.rowmax[vo + k0 + j] = rowmax_j;
  }
  {
    const int c = tid & (NB - 1);
    // strict lower part of the block back to S (unscaled rows Lu)
    for (int r = tid / NB; r < kb; r += 4)
      if (c < r) Sblk[r * ldS + c] = T[r * LDT + c];
    // Ut[m][k] = d_m thresh(omega_k Lu[k, m]) for m < k: the scaled, thresholded rows times D (Reimer.py:553-560),
    // written row by row (coalesced over k)
    if (c < kb) {
      const double wk = s_w[c];
      for (int m = tid / NB; m < c; m += 4) {
        double lt = wk * T[c * LDT + m];
        if (wk == 0 || fabs(lt) < mavL) lt = 0;
        Ut[m * NB + c] = s_d[m] * lt;
      }
    }
  }
#undef MDB_DIAG_BAR
}

// ------------------------------------------------------------------------------------------------
// Panel: rows j below the diagonal block, columns of the panel.  Per row, X (D Ltilde^T) = R with
// R[j, c] = (1 - omega_c) A[j, c] + omega_c S0[j, c]  (A[j, c] when the scaled row c is dropped; the mix identity of
// DESIGN.md 2), i.e. forward substitution over the 128 columns -- done four columns-blocks of 32 at a time as matrix
// products on the FP64 tensor cores:
//     Z_s = R_s - X_{<s} U[<s, s]          (U[m][c] = d_m Ltilde[c][m] = Ut, written by k_diag_block)
//     X_s = (Z_s Ltilde_ss^-T) D_s^-1      (the 32 x 32 inverses Vinv, formed by k_diag_block's helper warp)
// A CTA owns 64 rows, a warp 16 of them (one m16 tile; 4 n8 tiles = 4 independent accumulator chains per step).
// R / Z / X live in one shared tile (row-major), the A entries of the mix are staged TRANSPOSED so that both source
// layouts (column-major inside W on one GPU, row-major in A_loc when distributed) load coalesced and the warp-per-row
// passes read them without bank conflicts.  Ut and Vinv (160 KB for all CTAs together) are read straight from global
// memory: L1 / L2 hits.  alpha / beta / rowmax: one pass per row with lanes over the columns and a fixed butterfly
// (deterministic).  When Vinv[4096] != 0 (explosive regime, zero d) the rows substitute column by column instead.
// HBM traffic: reads 2 x rows x 128 doubles (S0, A), writes 3 x (Lu, X, Y).
// ------------------------------------------------------------------------------------------------
#define MDB_PANEL_RT 64
#define MDB_PANEL_THREADS 128
#define MDB_PANEL_LDX(NB) ((NB) + 4)
#define MDB_PANEL_LDA (MDB_PANEL_RT + 1)
#define MDB_PANEL_SMEM_DOUBLES(NB) (MDB_PANEL_RT * MDB_PANEL_LDX(NB) + (NB) * MDB_PANEL_LDA + 4 * (NB))

__device__ __forceinline__ void mdb_dmma(double (&c)[4], double a0, double a1, double b0) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(b0));
}

template <int NB>
__global__ void __launch_bounds__(MDB_PANEL_THREADS) k_panel_gemm(FactorDev F, int64_t k0, int64_t row_begin,
                                                                 int64_t n_rows, const double* __restrict__ Asrc,
                                                                 int64_t a_rs, int64_t a_cs, double* __restrict__ Wcol,
                                                                 int64_t w_ld, double* __restrict__ PX,
                                                                 double* __restrict__ PY, int64_t p_row0, int64_t p_ld) {
  // Wcol: pointer such that S0[j, k0 + c] lives at Wcol[j * w_ld + c]  (j = global row)
  // PX / PY: row (j - p_row0) of the operand buffers, leading dimension p_ld (>= NB: the buffers hold the
  // panels of one outer block side by side so that the bulk trailing update contracts over all of them)
  extern __shared__ double smem_d[];
  constexpr int RT = MDB_PANEL_RT, LDX = MDB_PANEL_LDX(NB), LDA = MDB_PANEL_LDA, SB = 32;
  double* xs = smem_d;             // xs[r * LDX + c]: S0 -> R -> Z -> X
  double* as_t = xs + RT * LDX;    // as_t[c * LDA + r]: A[j, k0 + c]
  double* s_d = as_t + NB * LDA;   // NB
  double* s_w = s_d + NB;          // NB
  double* s_drop = s_w + NB;       // NB
  double* s_inv = s_drop + NB;     // NB: 1 / d (0 where d == 0)

  const int64_t b = blockIdx.z;
  const int64_t vo = b * F.sVec;
  const double* __restrict__ U = F.Ut + b * F.sUt;
  const double* __restrict__ V = F.Vinv + b * F.sV;
  const double* hdr = F.hdr + b * F.sHdr;
  Asrc += b * F.sW;
  Wcol += b * F.sW;
  PX += b * F.sPanel;
  PY += b * F.sPanel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t jbase = row_begin + (int64_t)blockIdx.x * RT;
  const int64_t jend = row_begin + n_rows;
  const int nrow = (int)((jend - jbase < RT) ? jend - jbase : RT);

  for (int c = tid; c < NB; c += MDB_PANEL_THREADS) {
    const double dc = hdr[c];
    s_d[c] = dc;
    s_inv[c] = (dc != 0) ? 1.0 / dc : 0.0;
    s_w[c] = hdr[NB + c];
    s_drop[c] = hdr[2 * NB + c];
  }
  {  // S0 rows: 16-byte chunks, coalesced
    const int q = tid & 63;
    for (int rr = tid >> 6; rr < nrow; rr += 2) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + rr * LDX + 2 * q);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Wcol + (jbase + rr) * w_ld + 2 * q));
    }
  }
  if (a_rs == 1) {  // A is column-major in (j, c): consecutive threads take consecutive rows
    const int jr = tid & 63;
    if (jr < nrow)
      for (int c = tid >> 6; c < NB; c += 2) mdb_cp_async8(&as_t[c * LDA + jr], Asrc + (jbase + jr) + (k0 + c) * a_cs);
  } else {          // row-major: consecutive threads take consecutive columns
    for (int jr = 0; jr < nrow; ++jr) mdb_cp_async8(&as_t[tid * LDA + jr], Asrc + (jbase + jr) * a_rs + (k0 + tid) * a_cs);
  }
  mdb_cp_async_wait_all();
  __syncthreads();
  const bool unsafe = V[4 * SB * SB] != 0.0;

  // ---- R = mix(A, S0), beta += sum_c 2 A^2: the warp's own 16 rows, lanes over the columns ----
  const int r0 = 16 * warp;
  for (int rr = r0; rr < r0 + 16 && rr < nrow; ++rr) {
    double bsum = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = lane + 32 * e;
      const double a = as_t[c * LDA + rr];
      const double w = s_w[c];
      double r = a;
      if (s_drop[c] == 0.0) r = (1.0 - w) * a + w * xs[rr * LDX + c];
      xs[rr * LDX + c] = r;
      bsum += 2.0 * a * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if (lane == 0) F.beta[vo + jbase + rr] += bsum;   // Reimer.py:602-604
  }
  __syncwarp();

  const int g = lane >> 2, t4 = lane & 3;
  if (!unsafe) {
    double* X0 = xs + (r0 + g) * LDX;   // rows r0 + g and r0 + g + 8 of the tile
    double* X1 = X0 + 8 * LDX;
#pragma unroll
    for (int s = 0; s < NB / SB; ++s) {
      const int cs = SB * s;
      double acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        acc[nt][0] = X0[col]; acc[nt][1] = X0[col + 1]; acc[nt][2] = X1[col]; acc[nt][3] = X1[col + 1];
      }
      // Z_s = R_s - X_{<s} U[<s, s]
#pragma unroll 4
      for (int kq = 0; kq < cs; kq += 4) {
        const double a0 = -X0[kq + t4], a1 = -X1[kq + t4];
        const double* Up = U + (kq + t4) * NB + cs + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mdb_dmma(acc[nt], a0, a1, __ldg(Up + 8 * nt));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        X0[col] = acc[nt][0]; X0[col + 1] = acc[nt][1]; X1[col] = acc[nt][2]; X1[col + 1] = acc[nt][3];
      }
      __syncwarp();
      // W_s = Z_s Ltilde_ss^-T  (B[m][c] = inverse[c][m], zero for m > c: tile nt only needs m < 8 nt + 8)
      double w2[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) w2[nt][0] = w2[nt][1] = w2[nt][2] = w2[nt][3] = 0.0;
      const double* Vs = V + s * SB * SB;
#pragma unroll
      for (int kq = 0; kq < SB; kq += 4) {
        const double a0 = X0[cs + kq + t4], a1 = X1[cs + kq + t4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          if (kq < 8 * nt + 8) mdb_dmma(w2[nt], a0, a1, __ldg(Vs + (8 * nt + g) * SB + kq + t4));
      }
      __syncwarp();  // every lane has read Z before X replaces it
      // X_s = W_s D_s^-1
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        const double i0 = s_inv[col], i1 = s_inv[col + 1];
        X0[col] = w2[nt][0] * i0; X0[col + 1] = w2[nt][1] * i1; X1[col] = w2[nt][2] * i0; X1[col + 1] = w2[nt][3] * i1;
      }
      __syncwarp();
    }
  } else if (lane < 16 && r0 + lane < nrow) {
    // column-by-column substitution, one lane per row (explosive regime / zero d: rare, small matrices)
    double* Xr = xs + (r0 + lane) * LDX;
    for (int c = 0; c < NB; ++c) {
      double acc = Xr[c];
      for (int m = 0; m < c; ++m) acc = fma(-Xr[m], __ldg(U + m * NB + c), acc);
      Xr[c] = (s_d[c] != 0) ? acc * s_inv[c] : 0.0;
    }
  }
  __syncwarp();

  // ---- alpha, rowmax, and the results: Lu into W, X and Y = -X d into the operand buffers (coalesced rows) ----
  for (int rr = r0; rr < r0 + 16 && rr < nrow; ++rr) {
    const int64_t jj = jbase + rr;
    const int64_t pr = jj - p_row0;
    double asum = 0, mx = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = lane + 32 * e;
      const double x = xs[rr * LDX + c], dc = s_d[c];
      asum += x * x * dc;          // Reimer.py:679-684
      mx = fmax(mx, fabs(x));
      Wcol[jj * w_ld + c] = x;
      PX[pr * p_ld + c] = x;
      PY[pr * p_ld + c] = -x * dc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      asum += __shfl_xor_sync(0xffffffffu, asum, o);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      F.alpha[vo + jj] += asum;
      F.rowmax[vo + jj] = fmax(F.rowmax[vo + jj], mx);
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
