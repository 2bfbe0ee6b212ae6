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

// ------------------------------------------------------------------------------------------------
// The permutation of large matrices (n > 8192, one matrix) in TWO passes of "gather rows, transpose":
//     pass 1:  B[c, i] = A[p[i], c]            pass 2:  W[i, j] = B[p[j], i]  ( = A[p[i], p[j]] )
// A random p turns any one-pass formulation into 8-byte accesses on one side -- gathers inside the source row or
// scattered writes -- and those are what the memory system cannot do: measured at n = 65536 (tools/permute_bench.py),
// 8-byte gathers from L2 85 ms, scattered 8-byte writes 169 ms (partial sectors on ECC memory), the source row staged in
// the distributed shared memory of a cluster 61 ms (random ld.shared::cluster: ~0.17 accesses per clock and SM), the same
// with persistent clusters 99 ms.  A transposing tile kernel only ever touches 512-byte pieces of rows: tile (I, C) of
// 64 x 64 loads the 64 rows map[i] at the contiguous columns C (16-byte loads, 8 in flight per thread), turns it in
// shared memory and writes 64 contiguous pieces of the rows c of the destination.  Two such passes move 32 n^2 bytes
// instead of the algorithmic 16 n^2; pass 2 uses the symmetry of the result -- tile (I, J) and its mirror image come
// from one read -- so it reads only half of B: 28 n^2 bytes in total, all of them full sectors.
// The finite check of the input (dense/util.py:15-21) rides on pass 1, gamma[i] = W[i, i] on the diagonal tiles of pass 2.
// ------------------------------------------------------------------------------------------------
#define MDB_TT 64            // tile edge
#define MDB_TT_PITCH 65      // odd pitch: the transposed reads are conflict-free
#define MDB_TT_THREADS 256

__global__ void __launch_bounds__(MDB_TT_THREADS) k_gather_rows_transpose(int64_t n, const double* __restrict__ src, int64_t lds,
                                                                         const int64_t* __restrict__ map,
                                                                         double* __restrict__ dst, int64_t ldd, int symmetric,
                                                                         double* __restrict__ gamma, int* __restrict__ nonfinite) {
  // dst[c, i] = src[map[i], c] for i in tile blockIdx.x, c in tile blockIdx.y;  symmetric: only tiles with
  // blockIdx.y >= blockIdx.x, which also write dst[i, c] = dst[c, i]
  __shared__ double tile[MDB_TT * MDB_TT_PITCH];
  const int ti = blockIdx.x, tc = blockIdx.y;
  if (symmetric && tc < ti) return;
  const int64_t i0 = (int64_t)ti * MDB_TT, c0 = (int64_t)tc * MDB_TT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool bad = false;
  {  // rows map[i0 + r], r = warp + 8 k: 64 contiguous doubles each, two per lane
    const bool vec = ((lds & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (c0 + MDB_TT <= n);
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = warp + 8 * k;
      v[k] = make_double2(0.0, 0.0);
      if (i0 + r < n) {
        const int64_t row = map ? map[i0 + r] : i0 + r;
        const double* sp = src + row * lds + c0 + 2 * lane;
        if (vec) {
          v[k] = *reinterpret_cast<const double2*>(sp);
        } else {
          if (c0 + 2 * lane < n) v[k].x = sp[0];
          if (c0 + 2 * lane + 1 < n) v[k].y = sp[1];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = warp + 8 * k;
      tile[r * MDB_TT_PITCH + 2 * lane] = v[k].x;
      tile[r * MDB_TT_PITCH + 2 * lane + 1] = v[k].y;
      bad |= !(isfinite(v[k].x) && isfinite(v[k].y));
    }
  }
  __syncthreads();
  // dst rows c0 + cc, cc = warp + 8 k: columns i0 + lane, i0 + lane + 32 (the tile read down its columns)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int cc = warp + 8 * k;
    if (c0 + cc < n) {
      double* dp = dst + (c0 + cc) * ldd + i0;
      const double a = tile[lane * MDB_TT_PITCH + cc], b = tile[(lane + 32) * MDB_TT_PITCH + cc];
      if (i0 + lane < n) dp[lane] = a;
      if (i0 + lane + 32 < n) dp[lane + 32] = b;
      if (gamma && ti == tc) {
        if (lane == cc) gamma[c0 + cc] = a;
        if (lane + 32 == cc) gamma[c0 + cc] = b;
      }
    }
  }
  if (symmetric && tc != ti) {
    // the mirror image: dst rows i0 + r, columns c0 .. c0 + 63 (the tile read along its rows)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = warp + 8 * k;
      if (i0 + r < n) {
        double* dp = dst + (i0 + r) * ldd + c0;
        if (c0 + lane < n) dp[lane] = tile[r * MDB_TT_PITCH + lane];
        if (c0 + lane + 32 < n) dp[lane + 32] = tile[r * MDB_TT_PITCH + lane + 32];
      }
    }
  }
  if (bad && nonfinite) *nonfinite = 1;
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
// Lu[k, k0:k]^2 d from this very block: the kernel is ONE dependency chain
//     decision k  ->  x[k+1, k]  ->  alpha_{k+1}  ->  decision k+1,
// and only row k + 1 is on it.  Everything is arranged so that a column costs that chain and nothing else
// (round 1: one CTA-wide barrier per column, ~2200 cycles per column; now one warp carries the chain through
// register shuffles, ~250 cycles per column, and every other row trails it on a progress counter).
//
// Two-level inside the block: sub-blocks of SB = 32 columns.  Roles per sub-block (first column sb0):
//   chain warp (warp 0), lane = row sb0 + lane: owns the 32 x 32 triangle.  Per column k the lane of row k holds
//     the decision (d, omega, 1 / d, flags) in registers; one shuffle broadcasts it; every lane j > k forms
//         x[j, k] = ((1 - omega) A[j, k] + omega S''[j, k]) / d        (= A[j, k] / d when the scaled row k is dropped)
//     where S''[j, c], c = k .. sb0 + 31, is the row's part of the Schur complement kept up to date in REGISTERS
//     (right-looking inside the sub-block: after column k, S''[j, c] -= x[j, k] d_k x[c, k]; the register file is
//     rotated by one column per step so that the loop body has static indices), updates alpha / beta / rowmax /
//     rowmin in column order (Reimer.py:604,684), and every lane evaluates the cheap part of the rule on its own row
//     -- the unclamped test, MD_RULE_EXACT and the division-free upper-clamp shortcut (md_upper_clamp_shortcut_cheap) --
//     so that lane k + 1 has the next decision a few dependent instructions after alpha_{k+1}.  A column that needs
//     the cubic solves (Reimer.py:95-115) calls md_decide in that one lane (rare in the benign regime).
//   trailing warps (warps 1..3), one thread per row below the sub-block: the same column step, decision read from
//     shared memory once s_progress says column k is published (they never hold the chain up).
//   helper warp (warp 5): trails as well and forms, row by row, the inverse of the scaled unit triangle
//     Ltilde[c][m] = thresh(omega_c Lu[c, m]) of each sub-block; the panel kernel k_panel_gemm multiplies by these
//     32 x 32 inverses instead of substituting column by column.  Vinv[4096] = 1 tells it not to (an entry of an inverse
//     above 1e8 or not finite, |Lu| > 1e100 anywhere, or a zero d): it then substitutes.
//   When a sub-block is complete (named barrier of the 7 compute warps), its contribution
//     S'[j, c] -= sum_{m in sub-block} Lu[j, m] d_m Lu[c, m]  to every later column c of the block is applied at once on
//     the FP64 tensor cores (mma.m16n8k4 from shared memory; d_m Lu[c, m] sits at the transposed tile position (m, c),
//     where the A entry is dead once row c has passed column m).
// omega S'' ignores the reference's thresholding of tiny entries of the scaled row (Reimer.py:558).  A row falls back
// to the explicit thresholded dot product over ALL earlier columns of the block and the block-entry value of S
// (re-read from global memory, which this kernel only overwrites at its end) -- "slow", exact w.r.t. the reference
// inside the block -- whenever that could matter: omega_k * min nonzero |Lu[k, .]| < min_abs_value_L, or the products
// could overflow (|Lu| > 1e100 in row k or in the row itself: both are properties of the two rows involved, so the
// choice does not depend on how far the trailing warps lag).
// Outputs: d, omega, delta, clamped; Ut[m][k] = d_m thresh(omega_k Lu[k, m]) (Reimer.py:553-560); hdr = (d, omega,
// drop, delta, clamped) of the panel; Vinv.
// ------------------------------------------------------------------------------------------------
// Sblk / Ablk point at element (k0, k0) of the storage that holds S (lower part is read and written) and
// the permuted original A (upper part is read): the same array W on one GPU, S_loc / A_loc when distributed.
#define MDB_DIAG_WARPS 8
#define MDB_DIAG_THREADS(NB) (32 * MDB_DIAG_WARPS)
#define MDB_DIAG_SB 32
#define MDB_DIAG_YLD 36
#define MDB_DIAG_HELPER_WARP 4

// shared-memory layout of k_diag_block (offsets in doubles from the dynamic shared base)
template <int NB>
struct DiagLayout {
  static constexpr int LDT = NB + 4;            // 264 words: the DMMA fragment loads of the sub-block update are conflict-free
  static constexpr int T = 0;                   // T[r * LDT + c]: c < r: S0 -> S' -> Lu;  c >= r: A, then (r, c) <- d_r Lu[c, r]
  static constexpr int YT = T + NB * LDT;       // 33 x YLD: Yt[kk * YLD + c + ((kk + 1) & 1)] = d_k x[c, k] of the chain rows c
  static constexpr int DEC = YT + 33 * MDB_DIAG_YLD;  // 4 per column: d, omega, 1 / d, flags (1 = drop, 2 = slow, 4 = clamped)
  static constexpr int ALPHA = DEC + 4 * NB;    // row state, by row of the block
  static constexpr int BETA = ALPHA + NB;
  static constexpr int ROWMAX = BETA + NB;
  static constexpr int ROWMIN = ROWMAX + NB;
  static constexpr int GAMMA = ROWMIN + NB;
  static constexpr int MINB = GAMMA + NB;
  static constexpr int MAXB = MINB + NB;
  static constexpr int INV = MAXB + NB;         // 32 x 33: inverse of the current sub-block's scaled unit triangle (helper warp)
  static constexpr int LT = INV + 32 * 33;      // 36 (16-byte aligned): the scaled, thresholded row the helper is working on
  static constexpr int FLAGS = LT + 36;         // int progress: columns whose x of the chain rows and whose decision are published
  static constexpr int TOTAL = FLAGS + 2;
  static_assert((YT % 2) == 0 && (DEC % 2) == 0 && (LT % 2) == 0, "16-byte alignment of the vector accesses");
};
#define MDB_DIAG_SMEM_DOUBLES(NB) (DiagLayout<NB>::TOTAL)

// Shared-memory accesses of the column loop by 32-bit shared address (one base register, immediate offsets: the
// compiler otherwise re-derives the shared window base inside the loop).  asm volatile keeps their program order.
__device__ __forceinline__ double mdb_lds(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
template <int OFF>
__device__ __forceinline__ double2 mdb_lds2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ double mdb_lds1(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ void mdb_sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
template <int OFF>
__device__ __forceinline__ void mdb_sts2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0+%3], {%1, %2};" ::"r"(a), "d"(x), "d"(y), "n"(OFF) : "memory");
}
__device__ __forceinline__ int mdb_lds_flag(uint32_t a) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mdb_sts_flag(uint32_t a, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// S'[j, c] -= sum_{m0 <= m < m0 + SB} T[j][m] T[m][c]  for c0 <= c < j < kb (c0 = m0 + SB): items of 16 x 16 (one A
// fragment stream, two independent accumulator chains) dealt round-robin to the warps.  mma.m16n8k4: a0 = A[g][t4],
// a1 = A[g+8][t4], b0 = B[t4][g], c = C[g | g+8][2 t4 | 2 t4 + 1].
template <int NB>
__device__ __forceinline__ void mdb_diag_subblock_update(double* __restrict__ T, int m0, int kb, int warp, int lane) {
  constexpr int LDT = NB + 4;
  constexpr int SB = MDB_DIAG_SB;
  const int c0 = m0 + SB;
  const int g = lane >> 2, t4 = lane & 3;
  const int n_rt = (kb - c0 + 15) >> 4;
  int t = 0, next = warp;
  for (int ri = 0; ri < n_rt; ++ri) {
    for (int cj = 0; cj <= ri; ++cj, ++t) {
      if (t != next) continue;
      next += MDB_DIAG_WARPS;
      const int r0 = c0 + 16 * ri, cc0 = c0 + 16 * cj;
      double acc0[4] = {0.0, 0.0, 0.0, 0.0}, acc1[4] = {0.0, 0.0, 0.0, 0.0};
      const double* Ap0 = T + (r0 + g) * LDT + m0 + t4;
      const double* Ap1 = Ap0 + 8 * LDT;
      const double* Bp = T + (m0 + t4) * LDT + cc0 + g;
#pragma unroll
      for (int kk = 0; kk < SB; kk += 4) {
        const double a0 = Ap0[kk], a1 = Ap1[kk], b0 = Bp[kk * LDT], b1 = Bp[kk * LDT + 8];
        asm volatile(
            "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
            : "+d"(acc0[0]), "+d"(acc0[1]), "+d"(acc0[2]), "+d"(acc0[3])
            : "d"(a0), "d"(a1), "d"(b0));
        asm volatile(
            "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
            : "+d"(acc1[0]), "+d"(acc1[1]), "+d"(acc1[2]), "+d"(acc1[3])
            : "d"(a0), "d"(a1), "d"(b1));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = r0 + g + ((e & 2) ? 8 : 0), col = cc0 + 2 * t4 + (e & 1);
        if (row < kb && col < row) T[row * LDT + col] -= acc0[e];
        if (row < kb && col + 8 < row) T[row * LDT + col + 8] -= acc1[e];
      }
    }
  }
}

// x[j, k] of a "slow" column: the reference's own form, Reimer.py:553-560 and :586-600 inside the block,
//   x = ((1 - w) A[j, k] + w S0[j, k] - sum_{m < k} Lu[j, m] d_m thresh(w Lu[k, m])) / d_k
__device__ __noinline__ double mdb_diag_slow_x(const double* __restrict__ T, int ldt, const double* __restrict__ dec, int j,
                                               int k, double w, double inv, double a, double s0k, double mavL) {
  const double* Tj = T + j * ldt;
  const double* Tk = T + k * ldt;
  double a0 = 0, a1 = 0;
  int m = 0;
  for (; m + 1 < k; m += 2) {
    double l0 = w * Tk[m], l1 = w * Tk[m + 1];
    if (w == 0 || fabs(l0) < mavL) l0 = 0;
    if (w == 0 || fabs(l1) < mavL) l1 = 0;
    a0 = fma(Tj[m], dec[4 * m] * l0, a0);
    a1 = fma(Tj[m + 1], dec[4 * m + 4] * l1, a1);
  }
  if (m < k) {
    double l0 = w * Tk[m];
    if (w == 0 || fabs(l0) < mavL) l0 = 0;
    a0 = fma(Tj[m], dec[4 * m] * l0, a0);
  }
  return (((1.0 - w) * a + w * s0k) - (a0 + a1)) * inv;
}

// 1 / d for the chain: MUFU.RCP64H and the Newton steps of __drcp_rn without its special-case branch, so that the
// compiler can interleave it with the register update.  Only valid for 1e-280 <= |d| <= 1e280 (the caller redoes
// everything else with __drcp_rn).
__device__ __forceinline__ double mdb_rcp_normal(double d) {
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
  double e = fma(-d, x0, 1.0);
  e = fma(e, e, e);
  const double x1 = fma(x0, e, x0);
  const double e2 = fma(-d, x1, 1.0);
  return fma(x1, e2, x1);
}

__device__ __forceinline__ void mdb_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// What the chain keeps out of its loop body (one lane, rare): x of a column that is dropped, slow or has d = 0 ...
__device__ __noinline__ double mdb_diag_special_x(const double* __restrict__ T, int ldt, const double* __restrict__ dec, int j,
                                                  int k, double dk, double wk, double inv, int flags, bool row_big, double a,
                                                  double r0, const double* __restrict__ Sblk, int64_t ldS, double mavL) {
  if (dk == 0) return 0.0;
  if (flags & 1) return a * inv;   // the scaled row k is dropped: column k of L is A[:, k] / d
  if (k > 0 && ((flags & 2) || row_big)) return mdb_diag_slow_x(T, ldt, dec, j, k, wk, inv, a, Sblk[j * ldS + k], mavL);
  return fma(wk, r0, (1.0 - wk) * a) * inv;
}

// ... and a decision that is not one of the cheap cases (or whose d is outside the range of mdb_rcp_normal):
// out = (d, omega, 1 / d), returns the flags
__device__ __noinline__ int mdb_diag_special_decision(MdRuleParams prm, bool rare, double nd, int ncl, int kc, double al,
                                                      double be, double ga, double mB, double MB, double rowmax,
                                                      double rowmin, double* __restrict__ out) {
  double nw = 1.0;
  if (rare) {
    const MdDecision dd = md_decide(prm, al, be, ga, mB, MB);
    nd = dd.d; nw = dd.omega; ncl = dd.clamped;
  }
  const double ninv = (nd != 0) ? __drcp_rn(nd) : 0.0;
  const double mavL = prm.min_abs_value_L;
  const bool drop = (nw == 0 || nw * rowmax < mavL);
  const bool slow = !drop && kc > 0 && (nw * rowmin < mavL || rowmax > 1e100);
  out[0] = nd; out[1] = nw; out[2] = ninv;
  return (drop ? 1 : 0) | (slow ? 2 : 0) | (ncl ? 4 : 0);
}

// The columns of one sub-block for one row (thread): CHAIN = the rows of the sub-block itself (one warp, decisions made
// here), otherwise a row below it.  The body of the column loop is what the whole factorization of small matrices
// waits for.
template <int NB, bool EXACT, bool CHAIN>
__device__ __forceinline__ void mdb_diag_columns(const FactorDev& F, double* __restrict__ sm, int sb0, int kb, int j, int lane,
                                                 const double* __restrict__ Sblk, int64_t ldS) {
  using L = DiagLayout<NB>;
  constexpr int LDT = L::LDT, SB = MDB_DIAG_SB, YLD = MDB_DIAG_YLD;
  constexpr double BIG = 1e100;
  const unsigned FULL = 0xffffffffu;
  const bool valid = j < kb;
  const double mavL = F.prm.min_abs_value_L;
  const bool track_min = mavL > 0;   // rowmin only matters for the threshold test
  double* T = sm + L::T;
  double* dec = sm + L::DEC;
  const int jr = valid ? j : 0;   // rows beyond the block only go through the motions
  const double* Trow = T + jr * LDT;

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t a_col = sbase + 8u * (uint32_t)(L::T + sb0 * LDT + jr);   // &T[k][j]: A[j, k], then d_k Lu[j, k]
  uint32_t a_x = sbase + 8u * (uint32_t)(L::T + jr * LDT + sb0);     // &T[j][k]: Lu[j, k]
  uint32_t a_dec = sbase + 8u * (uint32_t)(L::DEC + 4 * sb0);        // decision of column k
  uint32_t a_yst = sbase + 8u * (uint32_t)(L::YT + lane);            // row kk of Yt, this lane's slot (before the parity shift)
  uint32_t a_yld = sbase + 8u * (uint32_t)(L::YT + 1);               // row kk of Yt from column kk + 1 on (before the parity shift)
  const uint32_t a_prog = sbase + 8u * (uint32_t)L::FLAGS;

  double r[SB];   // r[i] = S''[j, k + i] at the top of column step k (rotated by one per step)
#pragma unroll
  for (int c = 0; c < SB; c += 2) {
    const double2 v = *reinterpret_cast<const double2*>(Trow + sb0 + c);
    r[c] = v.x; r[c + 1] = v.y;
  }
  double alpha = sm[L::ALPHA + jr], beta = sm[L::BETA + jr], rowmax = sm[L::ROWMAX + jr], rowmin = sm[L::ROWMIN + jr];
  double gamma = 0, mB = 0, MB = 0;
  const double MD = F.prm.max_diag_D;
  MdRowConst rc;
  rc.lo = 0; rc.alpha_hi = 0; rc.ok = 0;
  bool d_in_range = false;   // lo and max_diag_D inside the range of mdb_rcp_normal: then so is every cheap d
  double dec_d = 0, dec_w = 0, dec_inv = 0;
  int dec_fl = 0;
  if (CHAIN) {
    gamma = sm[L::GAMMA + jr]; mB = sm[L::MINB + jr]; MB = sm[L::MAXB + jr];
    if (!EXACT) {
      rc = md_row_const(gamma, md_effective_min_diag_D(F.prm, gamma, mB, MB), MD, mB, F.prm.min_abs_value_D);
      d_in_range = rc.lo >= 1e-280 && MD <= 1e280;
    }
  }

  // The decision of this lane's own row from its state: the cheap cases -- the unclamped test (Reimer.py:68-72),
  // MD_RULE_EXACT, the division-free upper clamp -- are arithmetic only and evaluated by every lane after every column
  // (only the lane of the next column uses the result); everything else goes through mdb_diag_special_decision in
  // that one lane.  a_d = shared address of the column's decision record.
  auto decide = [&](int kc, bool mine, uint32_t a_d) {
    const double dstar = gamma - alpha;
    double nd = dstar;
    int ncl;
    bool rule_rare = false, special;
    if (EXACT) {
      ncl = !(dstar > 0);
      special = !(fabs(nd) >= 1e-280 && fabs(nd) <= 1e280);
    } else {
      const double a_lim = md_pymax(rc.lo, mB - alpha), t2 = MB - alpha, b_lim = md_pymin(MD, t2);
      const bool uncl = a_lim <= dstar && dstar <= b_lim;   // then lo <= d <= max_diag_D
      const bool sc = md_upper_clamp_shortcut_row(rc, alpha, beta, dstar, a_lim, t2, MD);
      nd = uncl ? dstar : MD;
      ncl = !uncl;
      rule_rare = !(uncl || sc);
      special = rule_rare || !d_in_range;
    }
    double ninv = mdb_rcp_normal(nd);
    const bool drop = rowmax < mavL;   // omega = 1 in every cheap case
    const bool slow = !drop && kc > 0 && ((track_min && rowmin < mavL) || rowmax > BIG);
    int fl = (drop ? 1 : 0) | (slow ? 2 : 0) | (ncl ? 4 : 0);
    double nw = 1.0;
    if (mine) {
      if (special) {
        double o[3];
        fl = mdb_diag_special_decision(F.prm, rule_rare, nd, ncl, kc, alpha, beta, gamma, mB, MB, rowmax, rowmin, o);
        nd = o[0]; nw = o[1]; ninv = o[2];
      }
      dec_d = nd; dec_w = nw; dec_inv = ninv; dec_fl = fl;
      mdb_sts2<0>(a_d, nd, nw);
      mdb_sts2<16>(a_d, ninv, (double)fl);
    }
  };

  if (CHAIN) decide(sb0, lane == 0 && valid, a_dec);
  const int kend = (kb - sb0 < SB) ? kb - sb0 : SB;
  for (int kk = 0; kk < kend; ++kk) {
    const int k = sb0 + kk;
    double dk, wk, inv;
    int flags;
    if (CHAIN) {
      dk = __shfl_sync(FULL, dec_d, kk);
      wk = __shfl_sync(FULL, dec_w, kk);
      inv = __shfl_sync(FULL, dec_inv, kk);
      flags = __shfl_sync(FULL, dec_fl, kk);
    } else {
      while (mdb_lds_flag(a_prog) <= k) __nanosleep(20);
      mdb_fence_cta();
      const double2 d01 = mdb_lds2<0>(a_dec), d23 = mdb_lds2<16>(a_dec);
      dk = d01.x; wk = d01.y; inv = d23.x; flags = (int)d23.y;
    }
    const bool active = valid && j > k;
    const double a = mdb_lds(a_col);   // A[j, k] (rows j > k)
    // the reciprocal comes with the decision (<= 1 ulp from r / d_k)
    double x = fma(wk, r[0], (1.0 - wk) * a) * inv;
    if (active && ((flags & 3) != 0 || dk == 0 || rowmax > BIG))
      x = mdb_diag_special_x(T, LDT, dec, j, k, dk, wk, inv, flags, rowmax > BIG, a, r[0], Sblk, ldS, mavL);
    if (!active) x = 0.0;
    const double y = x * dk;
    if (active) {
      mdb_sts(a_x, x);
      mdb_sts(a_col, y);                // d_k Lu[j, k] at the transposed position
      alpha = fma(x, y, alpha);         // Reimer.py:679-684
      beta = fma(2.0 * a, a, beta);     // Reimer.py:602-604
      const double ax = fabs(x);
      rowmax = fmax(rowmax, ax);
      if (track_min && ax != 0) rowmin = fmin(rowmin, ax);
    }
    const uint32_t par8 = 8u * (uint32_t)((kk + 1) & 1);
    if (CHAIN) {
      mdb_sts(a_yst + par8, y);
      __syncwarp();
    }
    // S''[j, c] -= x[j, k] (d_k x[c, k]) for the remaining columns c of the sub-block, registers rotated by one
    {
      const uint32_t ay = a_yld + par8;
      const double nx = -x;
#define MDB_ROT2(I)                                   \
  {                                                   \
    const double2 yy = mdb_lds2<8 * (I)>(ay);         \
    r[I] = fma(nx, yy.x, r[(I) + 1]);                 \
    r[(I) + 1] = fma(nx, yy.y, r[(I) + 2]);           \
  }
      MDB_ROT2(0) MDB_ROT2(2) MDB_ROT2(4) MDB_ROT2(6) MDB_ROT2(8) MDB_ROT2(10) MDB_ROT2(12) MDB_ROT2(14)
      MDB_ROT2(16) MDB_ROT2(18) MDB_ROT2(20) MDB_ROT2(22) MDB_ROT2(24) MDB_ROT2(26) MDB_ROT2(28)
#undef MDB_ROT2
      r[SB - 2] = fma(nx, mdb_lds1<8 * (SB - 2)>(ay), r[SB - 1]);
    }
    if (CHAIN) {
      decide(k + 1, lane == kk + 1 && valid, a_dec + 32u);   // the next column's row: its state is complete now
      __syncwarp();
      mdb_fence_cta();
      if (lane == 0) mdb_sts_flag(a_prog, k + 1);
    }
    a_col += 8u * LDT; a_x += 8u; a_dec += 32u; a_yst += 8u * YLD; a_yld += 8u * (YLD + 1);
  }
  if (valid) {
    sm[L::ALPHA + j] = alpha; sm[L::BETA + j] = beta; sm[L::ROWMAX + j] = rowmax; sm[L::ROWMIN + j] = rowmin;
  }
}

// Helper warp, one sub-block: trails the chain and forms, row by row, the inverse of the scaled unit triangle
// Ltilde[c][m] = thresh(omega_c Lu[c, m]):  row kk of the inverse = -(ltilde row kk) x (rows < kk of the inverse).
// Returns true when an entry is above 1e8 or not finite.
template <int NB>
__device__ __forceinline__ bool mdb_diag_inverse_rows(double* __restrict__ sm, int sb0, int kb, int lane, double mavL,
                                                      double* __restrict__ Vout) {
  using L = DiagLayout<NB>;
  constexpr int LDT = L::LDT, SB = MDB_DIAG_SB;
  double* T = sm + L::T;
  const double* dec = sm + L::DEC;
  double* s_inv = sm + L::INV;
  double* s_lt = sm + L::LT;
  const uint32_t a_prog = (uint32_t)__cvta_generic_to_shared(sm) + 8u * (uint32_t)L::FLAGS;
  bool bad = false;
  for (int r = 0; r < SB; ++r) s_inv[r * 33 + lane] = 0.0;   // rows beyond the current one count as zero below
  if (lane < 4) s_lt[32 + lane] = 0.0;
  __syncwarp();
  const int kend = (kb - sb0 < SB) ? kb - sb0 : SB;
  double* V = Vout + (sb0 / SB) * (SB * SB);
  for (int kk = 0; kk < kend; ++kk) {
    const int k = sb0 + kk;
    while (mdb_lds_flag(a_prog) <= k) __nanosleep(40);
    mdb_fence_cta();
    const double wk = dec[4 * k + 1];
    double lt = 0;
    if (lane < kk) {
      lt = wk * T[k * LDT + sb0 + lane];
      if (wk == 0 || fabs(lt) < mavL) lt = 0;
    }
    s_lt[lane] = lt;
    __syncwarp();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int m = 0; m < kk; m += 4) {   // (lt and the rows of the inverse are zero from kk on)
      const double2 l01 = *reinterpret_cast<const double2*>(s_lt + m), l23 = *reinterpret_cast<const double2*>(s_lt + m + 2);
      a0 = fma(l01.x, s_inv[m * 33 + lane], a0);
      a1 = fma(l01.y, s_inv[(m + 1) * 33 + lane], a1);
      a2 = fma(l23.x, s_inv[(m + 2) * 33 + lane], a2);
      a3 = fma(l23.y, s_inv[(m + 3) * 33 + lane], a3);
    }
    const double x = (lane < kk) ? -((a0 + a1) + (a2 + a3)) : ((lane == kk) ? 1.0 : 0.0);
    s_inv[kk * 33 + lane] = x;
    V[kk * SB + lane] = x;
    bad = bad || !(fabs(x) <= 1e8);   // also catches NaN
    __syncwarp();
  }
  for (int r = kend; r < SB; ++r) V[r * SB + lane] = (r == lane) ? 1.0 : 0.0;
  return bad;
}

template <int NB, bool EXACT>
__global__ void __launch_bounds__(32 * MDB_DIAG_WARPS, 1) k_diag_block(FactorDev F, int64_t k0, int kb, double* __restrict__ Sblk,
                                                                      int64_t ldS, const double* __restrict__ Ablk, int64_t ldA) {
  extern __shared__ double smem_d[];
  using L = DiagLayout<NB>;
  constexpr int LDT = L::LDT;
  constexpr int SB = MDB_DIAG_SB;
  constexpr int NT = 32 * MDB_DIAG_WARPS;
  static_assert(NB == 4 * SB, "three trailing warps cover the rows below the first sub-block");
  double* sm = smem_d;
  double* T = sm + L::T;
  double* dec = sm + L::DEC;

  const int64_t b = blockIdx.z;
  Sblk += b * F.sW;
  Ablk += b * F.sW;
  const int64_t vo = b * F.sVec;
  double* Ut = F.Ut + b * F.sUt;
  double* hdr = F.hdr + b * F.sHdr;
  double* Vout = F.Vinv + b * F.sV;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double mavL = F.prm.min_abs_value_L;
  const bool vec_ok = ((ldS & 1) == 0) && ((reinterpret_cast<uintptr_t>(Sblk) & 15) == 0);

  // whole block in flight at once (asynchronous copies, coalesced per row): strict lower part from S, diagonal and
  // upper part from A (one array on a single GPU: 16-byte copies)
  if (Sblk == Ablk && vec_ok) {
    const int c2 = 2 * (tid & (NB / 2 - 1));
    if (c2 < kb)   // (a pair that straddles kb copies one element of padding: the rows of W are padded to 128)
      for (int r = tid / (NB / 2); r < kb; r += NT / (NB / 2)) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&T[r * LDT + c2]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(&Sblk[r * ldS + c2]));
      }
  } else {
    const int c = tid & (NB - 1);
    if (c < kb)
      for (int r = tid / NB; r < kb; r += NT / NB)
        mdb_cp_async8(&T[r * LDT + c], (c < r) ? &Sblk[r * ldS + c] : &Ablk[r * ldA + c]);
  }
  if (tid < NB) {
    const bool in = tid < kb;
    sm[L::ALPHA + tid] = in ? F.alpha[vo + k0 + tid] : 0.0;
    sm[L::BETA + tid] = in ? F.beta[vo + k0 + tid] : 0.0;
    sm[L::ROWMAX + tid] = in ? F.rowmax[vo + k0 + tid] : 0.0;
    sm[L::ROWMIN + tid] = INFINITY;
    sm[L::GAMMA + tid] = in ? F.gamma[vo + k0 + tid] : 0.0;
    sm[L::MINB + tid] = (in && F.minB) ? F.minB[vo + k0 + tid] : F.prm.min_diag_B;
    sm[L::MAXB + tid] = (in && F.maxB) ? F.maxB[vo + k0 + tid] : F.prm.max_diag_B;
  }
  if (tid == NT - 1) *reinterpret_cast<volatile int*>(sm + L::FLAGS) = 0;
  mdb_cp_async_wait_all();
  __syncthreads();

  bool inv_bad = false;
  for (int sb0 = 0; sb0 < kb; sb0 += SB) {
    if (warp == 0) {
      mdb_diag_columns<NB, EXACT, true>(F, sm, sb0, kb, sb0 + lane, lane, Sblk, ldS);
    } else if (warp <= 3) {
      if (sb0 + SB * warp < kb) mdb_diag_columns<NB, EXACT, false>(F, sm, sb0, kb, sb0 + SB * warp + lane, lane, Sblk, ldS);
    } else if (warp == MDB_DIAG_HELPER_WARP) {
      inv_bad = mdb_diag_inverse_rows<NB>(sm, sb0, kb, lane, mavL, Vout) || inv_bad;
    }
    __syncthreads();   // every x of the finished sub-block is in shared memory
    if (sb0 + SB < kb) {
      mdb_diag_subblock_update<NB>(T, sb0, kb, warp, lane);
      __syncthreads();
    }
  }
  if (warp == MDB_DIAG_HELPER_WARP) {
    // the panel kernel must substitute instead of multiplying by the inverses: an entry of an inverse above 1e8 or
    // not finite, |Lu| > 1e100 in a row of the block (the row maxima are complete now), a zero d
    bool bad = inv_bad;
    for (int c = lane; c < kb; c += 32) bad = bad || sm[L::ROWMAX + c] > 1e100 || dec[4 * c] == 0;
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) Vout[4 * SB * SB] = anybad ? 1.0 : 0.0;
  }
  // results of the block: the row state, the decisions (kept in shared memory while the chain ran: nothing on the chip
  // waits for the global copies), the unscaled rows, Ut
  if (tid < kb) {
    const int64_t K = k0 + tid;
    const double alpha_k = sm[L::ALPHA + tid];   // alpha of row k is final when column k is decided
    F.alpha[vo + K] = alpha_k;
    F.beta[vo + K] = sm[L::BETA + tid];
    F.rowmax[vo + K] = sm[L::ROWMAX + tid];
    const double dk = dec[4 * tid], wk = dec[4 * tid + 1];
    const int fl = (int)dec[4 * tid + 3];
    const double delta = (dk - sm[L::GAMMA + tid]) + (wk != 0 ? wk * wk * alpha_k : 0.0);  // Reimer.py:541-545
    F.d[vo + K] = dk;
    F.omega[vo + K] = wk;
    F.delta[vo + K] = delta;
    F.clamped[vo + K] = (uint8_t)((fl >> 2) & 1);
    hdr[tid] = dk;
    hdr[NB + tid] = wk;
    hdr[2 * NB + tid] = (fl & 1) ? 1.0 : 0.0;
    hdr[3 * NB + tid] = delta;
    hdr[4 * NB + tid] = (double)((fl >> 2) & 1);
  }
  if (EXACT && warp == 1) {   // first pivot that is not positive, like LAPACK's info
    int first = NB;
    for (int c = lane; c < kb; c += 32)
      if (((int)dec[4 * c + 3]) & 4) { first = c; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (lane == 0 && first < NB && F.info[b] == 0) F.info[b] = (int)(k0 + first + 1);
  }
  // strict lower part of the block back to S (unscaled rows Lu)
  if (vec_ok) {
    const int c2 = 2 * (tid & (NB / 2 - 1));
#pragma unroll 4
    for (int r = tid / (NB / 2); r < kb; r += NT / (NB / 2)) {
      if (c2 + 1 < r) *reinterpret_cast<double2*>(&Sblk[r * ldS + c2]) = *reinterpret_cast<const double2*>(&T[r * LDT + c2]);
      else if (c2 < r) Sblk[r * ldS + c2] = T[r * LDT + c2];
    }
  } else {
    const int c = tid & (NB - 1);
#pragma unroll 4
    for (int r = tid / NB; r < kb; r += NT / NB)
      if (c < r) Sblk[r * ldS + c] = T[r * LDT + c];
  }
  {
    // Ut[m][k] = d_m thresh(omega_k Lu[k, m]) for m < k: the scaled, thresholded rows times D (Reimer.py:553-560),
    // written row by row (coalesced over k)
    const int c = tid & (NB - 1);
    if (c < kb) {
      const double wk = dec[4 * c + 1];
#pragma unroll 4
      for (int m = tid / NB; m < c; m += NT / NB) {
        double lt = wk * T[c * LDT + m];
        if (wk == 0 || fabs(lt) < mavL) lt = 0;
        Ut[m * NB + c] = dec[4 * m] * lt;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Panel: rows j below the diagonal block, columns of the panel.  Per row, X (D Ltilde^T) = R with
// R[j, c] = (1 - omega_c) A[j, c] + omega_c S0[j, c]  (A[j, c] when the scaled row c is dropped; the mix identity of
// DESIGN.md 2), i.e. forward substitution over the 128 columns -- done four columns-blocks of 32 at a time as matrix
// products on the FP64 tensor cores:
//     Z_s = R_s - X_{<s} U[<s, s]          (U[m][c] = d_m Ltilde[c][m] = Ut, written by k_diag_block)
//     X_s = (Z_s Ltilde_ss^-T) D_s^-1      (the 32 x 32 inverses Vinv, formed by k_diag_block's helper warp)
// A CTA owns 64 rows, a warp 16 of them (one m16 tile; 4 n8 tiles = 4 independent accumulator chains per step), and
// everything a warp does after the one CTA-wide barrier is private to its rows.  What the tensor-core loops read is
// in shared memory: the R / Z / X tile (row-major), the six 32 x 32 blocks of Ut above the diagonal blocks and the four
// inverses (row pitch 36 doubles: conflict-free fragment loads) -- staged by cp.async while the A entries of the mix
// travel from global memory into registers (64 independent loads per thread, issued first).  The A entries are
// addressed through (a_rs, a_cs): column-major inside W on one GPU (lanes take 16 consecutive rows of two columns),
// row-major in A_loc when distributed (lanes take consecutive columns of one row).  alpha / beta / rowmax of the rows
// are read at the start and written once at the end; the row sums use a fixed butterfly (deterministic).
// When Vinv[4096] != 0 (explosive regime, zero d) the rows substitute column by column instead.
// HBM traffic: reads 2 x rows x 128 doubles (S0, A), writes 3 x (Lu, X, Y).
// ------------------------------------------------------------------------------------------------
#define MDB_PANEL_RT 64
#define MDB_PANEL_THREADS 128
#define MDB_PANEL_LDX(NB) ((NB) + 4)
#define MDB_PANEL_LDB 36
#define MDB_PANEL_U_ROWS(NB) (3 * (NB) / 2)   // 32 + 64 + 96 rows of Ut above the diagonal blocks of column blocks 1, 2, 3
#define MDB_PANEL_SMEM_DOUBLES(NB) (MDB_PANEL_RT * MDB_PANEL_LDX(NB) + (MDB_PANEL_U_ROWS(NB) + (NB)) * MDB_PANEL_LDB + 6 * (NB) + 3 * MDB_PANEL_RT)

__device__ __forceinline__ void mdb_dmma(double (&c)[4], double a0, double a1, double b0) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(b0));
}

__device__ __forceinline__ void mdb_cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}

// Reduction of 16 per-lane values (one per row of a warp's strip) over the 32 lanes: lanes 2 R and 2 R + 1 end up with
// the total of value R.  Lanes are paired with strides 16, 8, 4, 2, 1 -- the tree of the xor butterfly, hence the same
// bits -- but every step halves the number of values a lane carries: 16 + 1 shuffles instead of 16 x 5.
template <bool MAX>
__device__ __forceinline__ double mdb_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const int o = 16 >> st, n = 8 >> st;
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double send = up ? v[i] : v[i + n];
      const double keep = up ? v[i + n] : v[i];
      const double recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[i] = MAX ? fmax(keep, recv) : keep + recv;
    }
  }
  const double recv = __shfl_xor_sync(0xffffffffu, v[0], 1);
  return MAX ? fmax(v[0], recv) : v[0] + recv;
}

template <int NB, bool A_ROW_MAJOR>
__global__ void __launch_bounds__(MDB_PANEL_THREADS) k_panel_gemm(FactorDev F, int64_t k0, int64_t row_begin,
                                                                 int64_t n_rows, const double* __restrict__ Asrc,
                                                                 int64_t a_rs, int64_t a_cs, double* __restrict__ Wcol,
                                                                 int64_t w_ld, double* __restrict__ PX,
                                                                 double* __restrict__ PY, int64_t p_row0, int64_t p_ld) {
  // Wcol: pointer such that S0[j, k0 + c] lives at Wcol[j * w_ld + c]  (j = global row)
  // PX / PY: row (j - p_row0) of the operand buffers, leading dimension p_ld (>= NB: the buffers hold the
  // panels of one outer block side by side so that the bulk trailing update contracts over all of them)
  extern __shared__ double smem_d[];
  constexpr int RT = MDB_PANEL_RT, LDX = MDB_PANEL_LDX(NB), LDB = MDB_PANEL_LDB, SB = 32;
  static_assert(NB == 4 * SB, "four column blocks of 32");
  double* xs = smem_d;                           // xs[r * LDX + c]: S0 -> R -> Z -> X
  double* us = xs + RT * LDX;                    // column block s = 1, 2, 3: rows m < 32 s of Ut[m][32 s ..], stacked
  double* vs = us + MDB_PANEL_U_ROWS(NB) * LDB;  // vs[(32 s + c) * LDB + m] = inverse_s[c][m]
  double* s_d = vs + NB * LDB;                   // NB
  double* s_w = s_d + NB;                        // NB
  double* s_drop = s_w + NB;                     // NB
  double* s_inv = s_drop + NB;                   // NB: 1 / d (0 where d == 0)
  double2* s_mix = reinterpret_cast<double2*>(s_inv + NB);   // NB pairs (1 - omega_c, omega_c); (1, 0) for a dropped column
  double* s_alpha = s_inv + 3 * NB;              // RT: alpha of the rows at entry
  double* s_rowmax = s_alpha + RT;               // RT: rowmax at entry
  double* s_beta = s_rowmax + RT;                // RT: beta at entry, then the new beta

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
  const int r0 = 16 * warp;          // the warp's rows of the tile
  const int64_t jw = jbase + r0;     // ... and of the matrix

  // ---- A entries of the warp's own 16 rows: 64 loads per lane, all in flight before anything waits ----
  // (A_ROW_MAJOR: a_cs == 1, A[j, k0 + c] = Asrc[j * a_rs + k0 + c]; otherwise a_rs == 1, Asrc[j + (k0 + c) * a_cs])
  double areg[64];
  if (!A_ROW_MAJOR) {  // column-major in (j, c): lane (cc, r) = (lane >> 4, lane & 15) takes columns 2 i + cc of row r
    const int r = lane & 15, cc = lane >> 4;
    const bool ok = r0 + r < nrow;
    const double* ap = Asrc + (jw + r) + (k0 + cc) * a_cs;
    const int64_t step = 2 * a_cs;
#pragma unroll
    for (int i = 0; i < 64; ++i, ap += step) areg[i] = ok ? __ldg(ap) : 0.0;
  } else {          // row-major: lanes take columns lane + 32 e of row rr
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const bool ok = r0 + rr < nrow;
      const double* ap = Asrc + (jw + rr) * a_rs + (k0 + lane);
#pragma unroll
      for (int e = 0; e < 4; ++e) areg[4 * rr + e] = ok ? __ldg(ap + 32 * e) : 0.0;
    }
  }
  // row state at entry (written back once at the end)
  if (tid < RT) {
    const bool ok = tid < nrow;
    s_alpha[tid] = ok ? F.alpha[vo + jbase + tid] : 0.0;
    s_rowmax[tid] = ok ? F.rowmax[vo + jbase + tid] : 0.0;
    s_beta[tid] = ok ? F.beta[vo + jbase + tid] : 0.0;
  }

  // ---- staged by cp.async: S0 rows, the blocks of Ut and the inverses, the panel header ----
  {
    const int q = tid & 63;
    for (int rr = tid >> 6; rr < nrow; rr += 2) mdb_cp_async16(xs + rr * LDX + 2 * q, Wcol + (jbase + rr) * w_ld + 2 * q);
    const int q16 = tid & 15;   // 16-byte chunk of a 32-column block row
    for (int row = tid >> 4; row < MDB_PANEL_U_ROWS(NB); row += MDB_PANEL_THREADS / 16) {
      // stacked row -> (column block s, m): rows [0, 32) -> s = 1, [32, 96) -> s = 2, [96, 192) -> s = 3
      const int s = (row < 32) ? 1 : ((row < 96) ? 2 : 3);
      const int m = row - ((s == 1) ? 0 : ((s == 2) ? 32 : 96));
      mdb_cp_async16(us + row * LDB + 2 * q16, U + m * NB + SB * s + 2 * q16);
    }
    for (int row = tid >> 4; row < NB; row += MDB_PANEL_THREADS / 16)
      mdb_cp_async16(vs + row * LDB + 2 * q16, V + row * SB + 2 * q16);
  }
  int dropped = 0;
  for (int c = tid; c < NB; c += MDB_PANEL_THREADS) {
    const double dc = hdr[c];
    s_d[c] = dc;
    s_inv[c] = (dc != 0) ? 1.0 / dc : 0.0;
    const double wc = hdr[NB + c], dropc = hdr[2 * NB + c];
    s_w[c] = wc;
    s_drop[c] = dropc;
    s_mix[c] = (dropc == 0.0) ? make_double2(1.0 - wc, wc) : make_double2(1.0, 0.0);
    dropped |= (dropc != 0.0);
  }
  const bool unsafe = V[4 * SB * SB] != 0.0;
  mdb_cp_async_wait_all();
  const int any_drop = __syncthreads_or(dropped);
  if (any_drop) {
    // a dropped column takes A[j, c] alone: its S0 entries (possibly not finite) must not enter the mix as 0 x S0
    for (int c = lane; c < NB; c += 32)
      if (s_drop[c] != 0.0)
        for (int rr = 0; rr < 16; ++rr) xs[(r0 + rr) * LDX + c] = 0.0;
    __syncwarp();
  }

  // ---- R = (1 - omega_c) A + omega_c S0, beta += sum_c 2 A^2 (Reimer.py:602-604) ----
  // (the partial sums of both branches are combined in the same order -- columns c = l + 32 e summed over e, then the
  // butterfly tree over l with strides 16, 8, 4, 2, 1 -- so that one GPU and the distributed path produce the same bits)
  if (!A_ROW_MAJOR) {
    const int r = lane & 15, cc = lane >> 4;
    double* xr = xs + (r0 + r) * LDX + cc;
    const double2* mx2 = s_mix + cc;
    double p[16];
#pragma unroll
    for (int li = 0; li < 16; ++li) p[li] = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
      for (int li = 0; li < 16; ++li) {
        const int i = 16 * e + li;   // column c = 2 i + cc = l + 32 e with l = 2 li + cc
        const double a = areg[i];
        const double2 m = mx2[2 * i];
        xr[2 * i] = fma(m.y, xr[2 * i], m.x * a);
        p[li] = fma(2.0 * a, a, p[li]);
      }
    }
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
      for (int li = 0; li < h; ++li) p[li] += p[li + h];   // l <-> l + 2 h: strides 16, 8, 4, 2
    }
    const double bsum = p[0] + __shfl_xor_sync(0xffffffffu, p[0], 16);   // l <-> l ^ 1: the lane of the other parity
    if (cc == 0) s_beta[r0 + r] += bsum;
  } else {
    double2 m[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) m[e] = s_mix[lane + 32 * e];
    double p[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      double* xr = xs + (r0 + rr) * LDX + lane;
      double bs = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double a = areg[4 * rr + e];
        xr[32 * e] = fma(m[e].y, xr[32 * e], m[e].x * a);
        bs = fma(2.0 * a, a, bs);
      }
      p[rr] = bs;
    }
    const double bsum = mdb_reduce16<false>(p, lane);   // lanes 2 R and 2 R + 1: row R
    if ((lane & 1) == 0) s_beta[r0 + (lane >> 1)] += bsum;
  }
  __syncwarp();

  const int g = lane >> 2, t4 = lane & 3;
  if (!unsafe) {
    double* X0 = xs + (r0 + g) * LDX;   // rows r0 + g and r0 + g + 8 of the tile
    double* X1 = X0 + 8 * LDX;
#pragma unroll 1
    for (int s = 0; s < NB / SB; ++s) {   // (rolled: the body is executed four times from the instruction cache)
      const int cs = SB * s;
      double acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        const double2 v0 = *reinterpret_cast<const double2*>(X0 + col), v1 = *reinterpret_cast<const double2*>(X1 + col);
        acc[nt][0] = v0.x; acc[nt][1] = v0.y; acc[nt][2] = v1.x; acc[nt][3] = v1.y;
      }
      // Z_s = R_s - X_{<s} U[<s, s]
      const double* Ub = us + (16 * s * (s - 1)) * LDB + g;   // stacked rows start at 0, 32, 96
#pragma unroll 2
      for (int kq = 0; kq < cs; kq += 4) {
        const double a0 = -X0[kq + t4], a1 = -X1[kq + t4];
        const double* Up = Ub + (kq + t4) * LDB;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mdb_dmma(acc[nt], a0, a1, Up[8 * nt]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        *reinterpret_cast<double2*>(X0 + col) = make_double2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<double2*>(X1 + col) = make_double2(acc[nt][2], acc[nt][3]);
      }
      __syncwarp();
      // W_s = Z_s Ltilde_ss^-T  (B[m][c] = inverse[c][m], zero for m > c: tile nt only needs m < 8 nt + 8)
      double w2[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) w2[nt][0] = w2[nt][1] = w2[nt][2] = w2[nt][3] = 0.0;
      const double* Vs = vs + (cs + g) * LDB + t4;
#pragma unroll
      for (int kq = 0; kq < SB; kq += 4) {
        const double a0 = X0[cs + kq + t4], a1 = X1[cs + kq + t4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          if (kq < 8 * nt + 8) mdb_dmma(w2[nt], a0, a1, Vs[(8 * nt) * LDB + kq]);
      }
      __syncwarp();  // every lane has read Z before X replaces it
      // X_s = W_s D_s^-1
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = cs + 8 * nt + 2 * t4;
        const double i0 = s_inv[col], i1 = s_inv[col + 1];
        *reinterpret_cast<double2*>(X0 + col) = make_double2(w2[nt][0] * i0, w2[nt][1] * i1);
        *reinterpret_cast<double2*>(X1 + col) = make_double2(w2[nt][2] * i0, w2[nt][3] * i1);
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

  // ---- alpha, rowmax, and the results: Lu into W, X and Y = -X d into the operand buffers (16-byte stores) ----
  // Four rows per (rolled) iteration; their sums are reduced over the lanes with strides 16, 8 (halving the number of
  // values per lane) and then 4, 2, 1: lane (b4, b3, 0, 0, 0) ends up with row 2 b4 + b3 of the group.
  {
    const double2 d01 = *reinterpret_cast<const double2*>(s_d + 2 * lane), d23 = *reinterpret_cast<const double2*>(s_d + 64 + 2 * lane);
    const double* xp = xs + r0 * LDX + 2 * lane;
    double* wp = Wcol + jw * w_ld + 2 * lane;
    double* pxp = PX + (jw - p_row0) * p_ld + 2 * lane;
    double* pyp = PY + (jw - p_row0) * p_ld + 2 * lane;
    const int row_in_grp = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {
      double as4[4], mx4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q, xp += LDX, wp += w_ld, pxp += p_ld, pyp += p_ld) {
        as4[q] = 0.0; mx4[q] = 0.0;
        if (r0 + 4 * grp + q < nrow) {   // (uniform)
          const double2 x01 = *reinterpret_cast<const double2*>(xp), x23 = *reinterpret_cast<const double2*>(xp + 64);
          *reinterpret_cast<double2*>(wp) = x01;
          *reinterpret_cast<double2*>(wp + 64) = x23;
          *reinterpret_cast<double2*>(pxp) = x01;
          *reinterpret_cast<double2*>(pxp + 64) = x23;
          *reinterpret_cast<double2*>(pyp) = make_double2(-x01.x * d01.x, -x01.y * d01.y);
          *reinterpret_cast<double2*>(pyp + 64) = make_double2(-x23.x * d23.x, -x23.y * d23.y);
          double a = x01.x * x01.x * d01.x;          // Reimer.py:679-684
          a = fma(x01.y * x01.y, d01.y, a);
          a = fma(x23.x * x23.x, d23.x, a);
          a = fma(x23.y * x23.y, d23.y, a);
          as4[q] = a;
          mx4[q] = fmax(fmax(fabs(x01.x), fabs(x01.y)), fmax(fabs(x23.x), fabs(x23.y)));
        }
      }
      {
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double sa = up16 ? as4[i] : as4[i + 2], ka = up16 ? as4[i + 2] : as4[i];
          const double sm = up16 ? mx4[i] : mx4[i + 2], km = up16 ? mx4[i + 2] : mx4[i];
          as4[i] = ka + __shfl_xor_sync(0xffffffffu, sa, 16);
          mx4[i] = fmax(km, __shfl_xor_sync(0xffffffffu, sm, 16));
        }
        const double sa = up8 ? as4[0] : as4[1], ka = up8 ? as4[1] : as4[0];
        const double sm = up8 ? mx4[0] : mx4[1], km = up8 ? mx4[1] : mx4[0];
        double a = ka + __shfl_xor_sync(0xffffffffu, sa, 8);
        double m = fmax(km, __shfl_xor_sync(0xffffffffu, sm, 8));
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        }
        const int R = r0 + 4 * grp + row_in_grp;
        if ((lane & 7) == 0 && R < nrow) {
          F.alpha[vo + jbase + R] = s_alpha[R] + a;
          F.beta[vo + jbase + R] = s_beta[R];
          F.rowmax[vo + jbase + R] = fmax(s_rowmax[R], m);
        }
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
                                                  int64_t ldo, int64_t sOut, int64_t row0,
                                                  uint8_t* __restrict__ changed = nullptr, int64_t changed_cols = 0) {
  // changed (optional): changed[i] = 1 when an entry of row i left of column changed_cols differs from the unscaled
  // value that was there (omega_i != 1 or an entry lost to the threshold) -- the one-shot call ships those columns of
  // the late rows before the rows are final and must send such rows again (mdb_factor_run).
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
        const double lu = Wrow[j];
        v = w * lu;
        if (w == 0 || fabs(v) < mavL) v = 0;
        if (changed && j < changed_cols && !(v == lu)) changed[i] = 1;
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

// Streamed download, late rows (mdb_factor_run): the left part [0, cols) of a row went home before the row was final;
// where k_finalize found that the final values differ (changed[r]) it is written again, straight into the pinned host
// matrix (zero-copy stores over PCIe) from the copy stream -- no host round trip, so it rides behind the factorization
// like every other row block.  A SMALL grid (MDB_RESEND_CTAS) walks the (row, 1024-double segment) pairs: the kernel is
// bound by PCIe, and one CTA per segment (tens of thousands, each waiting on its stores) took the registers of half of
// the tile kernel's CTAs for 0.1 s (measured at n = 65536, F2 s = 0.9: factorization 0.11 s longer).
#define MDB_RESEND_CTAS 32
__global__ void __launch_bounds__(256) k_resend_rows(const uint8_t* __restrict__ changed, const double* __restrict__ W,
                                                     int64_t ldw, double* __restrict__ out, int64_t ldo, int64_t row0,
                                                     int64_t rows, int64_t cols) {
  const int64_t segs = (cols + 1023) / 1024;
  for (int64_t t = blockIdx.x; t < rows * segs; t += gridDim.x) {
    const int64_t r = row0 + t / segs, c0 = (t % segs) * 1024;
    if (!changed[r]) continue;
    const double* __restrict__ src = W + r * ldw + c0;
    double* __restrict__ dst = out + r * ldo + c0;
    const int64_t m = cols - c0 < 1024 ? cols - c0 : 1024;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
      const int64_t ca = 2 * threadIdx.x, cb = ca + 512;
      double2 va = make_double2(0, 0), vb = va;
      if (ca + 1 < m) va = *reinterpret_cast<const double2*>(src + ca);
      if (cb + 1 < m) vb = *reinterpret_cast<const double2*>(src + cb);
      if (ca + 1 < m) *reinterpret_cast<double2*>(dst + ca) = va;
      else if (ca < m) dst[ca] = src[ca];
      if (cb + 1 < m) *reinterpret_cast<double2*>(dst + cb) = vb;
      else if (cb < m) dst[cb] = src[cb];
    } else {
      for (int64_t c = threadIdx.x; c < m; c += 256) dst[c] = src[c];
    }
  }
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
