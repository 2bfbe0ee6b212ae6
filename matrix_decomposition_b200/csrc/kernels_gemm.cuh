// kernels_gemm.cuh -- FP64 tensor-core (DMMA) tile kernel  C = beta C +/- A B^T  with both operands
// K-major ("TN"): A is M x K (lda), B is N x K (ldb), C is M x N (ldc), all row-major float64.
//
// This is the only dense contraction of the path: the trailing Schur-complement update
//   S[j, j'] -= sum_k Lu[j, k] d_k Lu[j', k]      (A = PX, B = PY = -X D, lower triangle only)
// and, with other flags, the block updates of the multi-right-hand-side triangular solves.
//
// Design (B200, sm_100a):  FP64 has no tcgen05 kind -- the FP64 tensor path on Blackwell is mma.sync
// (SASS DMMA.8x8x4; 37.1 TFLOP/s measured, tools/fp64_peak.cu).  CTA tile 128 x 64, 8 warps of 32 x 32,
// K staged 16 doubles (= 128 B rows) at a time through a 4-stage cp.async ring, two CTAs per SM so one
// CTA's C-tile traffic hides under the other's DMMA stream.  Shared-memory rows carry the
// SWIZZLE_128B pattern (16-byte chunk index XOR (row & 7)), i.e. exactly what a TMA tensor map would
// write, and MMA row g is mapped to tile row rho(g) = ((g & 1) << 2) | (g >> 1), which makes every
// LDS.128 fragment load bank-conflict free: the 8 lanes of a quarter warp (rows g, g+1) then touch
// chunk sets that differ in bit 2.  The k index inside a 16-wide stage is permuted
// (thread t supplies k = 2t, 2t+1, 2t+8, 2t+9 in the four k4 steps) -- the same permutation for A
// and B, so the contraction is unchanged.
#pragma once
#include "mdb_common.cuh"

struct GemmArgs {
  const double* A; int64_t lda; int64_t sA;  // sX: batch strides (elements)
  const double* B; int64_t ldb; int64_t sB;
  double* C; int64_t ldc; int64_t sC;
  int64_t M, N;
  int K;                 // multiple of 16
  int64_t row_glob0;     // global row index of A row 0 and global column index of B row 0:
  int64_t col_glob0;     //   with LOWER only elements with (row_glob0 + i) > (col_glob0 + j) are touched
  int64_t m_tile0;       // first M tile (units of 128) this launch covers (grid.y bands start here)
  int64_t n_tiles;       // number of N tiles (units of 64) this launch covers, starting at n_tile0
  int64_t n_tile0;
  // Column mapping.  Local N tile nt (C columns nt*64 ...) has the global column
  //   cyc_G <= 1 : col_glob0 + nt*64                                  (one GPU)
  //   cyc_G  > 1 : ((nt / 2) * cyc_G + cyc_r) * 128 + (nt % 2) * 64   (1-D block-cyclic column slab)
  // and its B rows start at (global column - b_glob0); b_rows = number of valid B rows (0: N).
  int cyc_G, cyc_r;
  int64_t b_glob0;
  int64_t b_rows;
};

namespace gemm {
constexpr int BM = 128, BN = 64, BK = 16, STAGES = 4, THREADS = 256;
constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;  // 98304

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma_m16n8k4(double (&c)[4], double a0, double a1, double b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a0), "d"(a1), "d"(b0));
}
}  // namespace gemm

template <bool LOWER, bool BETA1, bool NEG>
__global__ void __launch_bounds__(gemm::THREADS, 2) k_gemm_tn(GemmArgs g) {
  using namespace gemm;
  extern __shared__ __align__(1024) unsigned char smem_raw[];

  // tile mapping: blockIdx.y = band of 8 M-tiles, blockIdx.x = (N tile) * 8 + (M tile in band);
  // consecutive CTAs share the band's A rows and walk the B tiles (L2 locality of the panel).
  const int64_t mt = g.m_tile0 + (int64_t)blockIdx.y * 8 + (blockIdx.x & 7);
  const int64_t nt = g.n_tile0 + (blockIdx.x >> 3);
  const int64_t i0 = mt * BM, j0 = nt * BN;
  if (i0 >= g.M || j0 >= g.N) return;
  // global column of the tile's first column, and the B row that belongs to it
  const int64_t gj0 = (g.cyc_G > 1) ? ((nt / (MDB_NB / BN)) * g.cyc_G + g.cyc_r) * MDB_NB + (nt % (MDB_NB / BN)) * BN
                                    : g.col_glob0 + j0;
  const int64_t brow0 = gj0 - g.b_glob0;
  const int64_t Nb = g.b_rows > 0 ? g.b_rows : g.N;  // number of valid B rows
  if (LOWER && (g.row_glob0 + i0 + BM - 1) <= gj0) return;  // tile entirely on/above the diagonal
  if (brow0 < 0) return;                                     // column block already factored

  const int64_t bz = blockIdx.z;
  const double* A = g.A + bz * g.sA;
  const double* B = g.B + bz * g.sB;
  double* C = g.C + bz * g.sC;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int gq = lane >> 2, t = lane & 3;
  const int rho = ((gq & 1) << 2) | (gq >> 1);
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);

  // ---- producer mapping: 8 consecutive threads fetch one 128-byte row ----
  // (source pointers are recomputed per stage: keeping them live costs 12+ registers under the 128 cap)
  const int ld_row = tid >> 3, ld_chunk = tid & 7;
  const uint32_t ld_dst = ld_row * 128 + ((ld_chunk ^ (ld_row & 7)) << 4);  // rows ld_row + 32 e share (row & 7)
  auto load_stage = [&](int slot, int kt) {
    const uint32_t sb = smem_base + slot * STAGE_BYTES + ld_dst;
    const int64_t ko = (int64_t)kt * BK + ld_chunk * 2;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int64_t gr = i0 + ld_row + 32 * e;
      gr = gr < g.M ? gr : g.M - 1;
      cp_async16(sb + e * 32 * 128, A + gr * g.lda + ko);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      int64_t gr = brow0 + ld_row + 32 * e;
      gr = gr < Nb ? gr : Nb - 1;
      cp_async16(sb + A_BYTES + e * 32 * 128, B + gr * g.ldb + ko);
    }
  };

  const int KT = g.K / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  // ---- accumulators, initialised from C (C tile traffic overlaps the pipeline fill) ----
  double acc[2][4][4];
  const bool interior = (i0 + BM <= g.M) && (j0 + BN <= g.N) && (!LOWER || (g.row_glob0 + i0) > (gj0 + BN - 1));
  const int64_t cg_off = gj0 - j0;  // global column = local column + cg_off
  const int64_t crow0 = i0 + wm * 32 + rho;
  const int64_t ccol0 = j0 + wn * 32 + t;
  if (BETA1) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t r = crow0 + mb * 16 + (q >> 1) * 8;
          const int64_t c = ccol0 + nb * 8 + (q & 1) * 4;
          double v = 0.0;
          if (interior || (r < g.M && c < g.N && (!LOWER || (g.row_glob0 + r) > (cg_off + c))))
            v = C[r * g.ldc + c];
          acc[mb][nb][q] = NEG ? -v : v;
        }
  } else {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mb][nb][q] = 0.0;
  }

  // ---- consumer addresses (bytes inside a stage) ----
  uint32_t a_off[2][2], b_off[4];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) a_off[mb][h2] = (wm * 32 + mb * 16 + h2 * 8 + rho) * 128;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) b_off[nb] = A_BYTES + (wn * 32 + nb * 8 + rho) * 128;
  const uint32_t ch[2] = {(uint32_t)(((t) ^ rho) << 4), (uint32_t)(((t + 4) ^ rho) << 4)};

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const uint32_t sb = smem_base + (kt % STAGES) * STAGE_BYTES;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      double2 af[2][2], bf[4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) af[mb][h2] = lds128(sb + a_off[mb][h2] + ch[hh]);
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) bf[nb] = lds128(sb + b_off[nb] + ch[hh]);
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          dmma_m16n8k4(acc[mb][nb], af[mb][0].x, af[mb][1].x, bf[nb].x);
          dmma_m16n8k4(acc[mb][nb], af[mb][0].y, af[mb][1].y, bf[nb].y);
        }
    }
  }
  cp_async_wait<0>();

  // ---- store ----
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t r = crow0 + mb * 16 + (q >> 1) * 8;
        const int64_t c = ccol0 + nb * 8 + (q & 1) * 4;
        if (interior || (r < g.M && c < g.N && (!LOWER || (g.row_glob0 + r) > (cg_off + c))))
          C[r * g.ldc + c] = NEG ? -acc[mb][nb][q] : acc[mb][nb][q];
      }
}

// host-side launcher
template <bool LOWER, bool BETA1, bool NEG>
static int launch_gemm_tn(mdb_ctx* ctx, const GemmArgs& g0, int64_t batch) {
  using namespace gemm;
  static uint64_t configured = 0;
  auto kern = k_gemm_tn<LOWER, BETA1, NEG>;
  if (!mdb_device_marked(configured, ctx)) {
    MDB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MDB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    mdb_mark_device(configured, ctx);
  }
  GemmArgs g = g0;
  const int64_t m_tiles = (g.M + BM - 1) / BM - g.m_tile0;
  const int64_t n_tiles_total = (g.N + BN - 1) / BN;
  if (g.n_tiles <= 0) g.n_tiles = n_tiles_total - g.n_tile0;
  if (m_tiles <= 0 || g.n_tiles <= 0) return MDB_OK;
  const int64_t bands = (m_tiles + 7) / 8;
  // grid.y is limited to 65535 bands (8.3M rows): fine for any matrix that fits in HBM
  dim3 grid((unsigned)(g.n_tiles * 8), (unsigned)bands, (unsigned)batch);
  kern<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(g);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}
