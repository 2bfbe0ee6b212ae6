// kernels_gemm_tma.cuh -- TMA-fed variant of the FP64 DMMA tile kernel for the trailing update
//   S[i, j] += sum_k X[i, k] Y[j, k]   (strict lower triangle; X, Y = compact (rows x 128) panel operands).
//
// Same tile shape, fragment mapping and shared-memory layout as k_gemm_tn (kernels_gemm.cuh): the layout was
// chosen to be what a SWIZZLE_128B tensor map writes.  What changes is the feed and the synchronisation:
//   * operands arrive by cp.async.bulk.tensor.2d (TMA, one elected thread, 2 instructions per 24 KB stage)
//     instead of 6 cp.async + address arithmetic per thread per stage; out-of-range rows are zero-filled by
//     the TMA unit, so no clamping;
//   * stages are handed over with mbarriers (full: TMA transaction bytes; empty: one arrival per warp), so
//     there is no CTA-wide barrier in the main loop -- a warp only ever waits for data;
//   * the grid is 1-D over exactly the tiles that intersect the lower triangle (per-band prefix table),
//     instead of a rectangle of which half the CTAs exit immediately.
#pragma once
#include <cuda.h>

#include "kernels_gemm.cuh"

#define MDB_MAX_BANDS 160

struct GemmTmaArgs {
  double* C; int64_t ldc;
  int64_t M, N;
  int K;
  int64_t row_glob0, col_glob0;
  int64_t m_tile0, n_tile0;
  int cyc_G, cyc_r;
  int64_t b_glob0;
  int variant;   // debugging: 1 = CTA-wide barrier before every refill
  int mode;      // bit 0: C -= A B^T (instead of +=), bit 1: C is not read (C = A B^T)
  int nbands;
  int band_prefix[MDB_MAX_BANDS + 1];  // tiles before band b (band = 8 M-tiles), tiles ordered (N tile, M tile in band)
};

namespace gemm {
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MDB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MDB_DONE;\n"
      "bra MDB_WAIT;\n"
      "MDB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
}  // namespace gemm

__global__ void __launch_bounds__(gemm::THREADS, 2)
    k_gemm_tn_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ GemmTmaArgs g) {
  using namespace gemm;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * STAGES];  // full[STAGES], empty[STAGES]

  // ---- tile lookup: band by binary search in the prefix table ----
  const int t = blockIdx.x;
  int lo = 0, hi = g.nbands;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (g.band_prefix[mid] <= t) lo = mid; else hi = mid;
  }
  const int local = t - g.band_prefix[lo];
  const int64_t mt = g.m_tile0 + (int64_t)lo * 8 + (local & 7);
  const int64_t nt = g.n_tile0 + (local >> 3);
  const int64_t i0 = mt * BM, j0 = nt * BN;
  if (i0 >= g.M || j0 >= g.N) return;
  const int64_t gj0 = (g.cyc_G > 1) ? ((nt / (MDB_NB / BN)) * g.cyc_G + g.cyc_r) * MDB_NB + (nt % (MDB_NB / BN)) * BN
                                    : g.col_glob0 + j0;
  const int64_t brow0 = gj0 - g.b_glob0;
  if ((g.row_glob0 + i0 + BM - 1) <= gj0) return;  // tile entirely on/above the diagonal
  if (brow0 < 0) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int gq = lane >> 2, tq = lane & 3;
  const int rho = ((gq & 1) << 2) | (gq >> 1);
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };

  const int KT = g.K / BK;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), (g.variant & 2) ? THREADS : THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue_stage = [&](int kt) {  // one elected thread
    const int slot = kt % STAGES;
    const uint32_t sb = smem_base + slot * STAGE_BYTES;
    mbar_expect_tx(full_bar(slot), STAGE_BYTES);
    tma_load_2d(sb, &tmA, full_bar(slot), kt * BK, (int)i0);
    tma_load_2d(sb + A_BYTES, &tmB, full_bar(slot), kt * BK, (int)brow0);
  };
  if (tid == 0) {
    const int npre = KT < STAGES ? KT : STAGES;
    for (int s = 0; s < npre; ++s) issue_stage(s);
  }

  // ---- accumulators, initialised from C (C tile traffic overlaps the pipeline fill) ----
  double acc[2][4][4];
  const bool interior = (i0 + BM <= g.M) && (j0 + BN <= g.N) && ((g.row_glob0 + i0) > (gj0 + BN - 1));
  const int64_t cg_off = gj0 - j0;
  const int64_t crow0 = i0 + wm * 32 + rho;
  const int64_t ccol0 = j0 + wn * 32 + tq;
  double* __restrict__ C = g.C;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t r = crow0 + mb * 16 + (q >> 1) * 8;
        const int64_t c = ccol0 + nb * 8 + (q & 1) * 4;
        double v = 0.0;
        if (!(g.mode & 2) && (interior || (r < g.M && c < g.N && (g.row_glob0 + r) > (cg_off + c)))) v = C[r * g.ldc + c];
        acc[mb][nb][q] = (g.mode & 1) ? -v : v;   // C - A B^T = -(-C + A B^T), exactly (rounding is symmetric)
      }

  uint32_t a_off[2][2], b_off[4];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) a_off[mb][h2] = (wm * 32 + mb * 16 + h2 * 8 + rho) * 128;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) b_off[nb] = A_BYTES + (wn * 32 + nb * 8 + rho) * 128;
  const uint32_t ch[2] = {(uint32_t)(((tq) ^ rho) << 4), (uint32_t)(((tq + 4) ^ rho) << 4)};

  for (int kt = 0; kt < KT; ++kt) {
    const int slot = kt % STAGES;
    // producer: the slot that held stage kt - 2 is refilled with stage kt + 2 (two stages of slack on both sides)
    if (g.variant & 1) __syncthreads();
    if (g.variant & 8) __syncwarp();
    if (tid == 0) {
      const int nk = kt + 2;
      if (nk >= STAGES && nk < KT) {
        const int ps = nk % STAGES;
        mbar_wait(empty_bar(ps), ((nk / STAGES) - 1) & 1);
        issue_stage(nk);
      }
    }
    mbar_wait(full_bar(slot), (kt / STAGES) & 1);
    const uint32_t sb = smem_base + slot * STAGE_BYTES;
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
    if (g.variant & 4) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (g.variant & 2) {
      mbar_arrive(empty_bar(slot));   // debugging variant: every thread arrives
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
    }
  }

#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t r = crow0 + mb * 16 + (q >> 1) * 8;
        const int64_t c = ccol0 + nb * 8 + (q & 1) * 4;
        if (interior || (r < g.M && c < g.N && (g.row_glob0 + r) > (cg_off + c)))
          C[r * g.ldc + c] = (g.mode & 1) ? -acc[mb][nb][q] : acc[mb][nb][q];
      }
}

// ---- host side ----
typedef CUresult (*mdb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static mdb_encode_tiled_fn mdb_get_encode_tiled() {
  static mdb_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (mdb_encode_tiled_fn)p;
  }
  return fn;
}

// (rows x K) row-major FP64 operand, box = box_rows x 16 doubles (128 bytes), SWIZZLE_128B, zero fill out of range
static bool mdb_make_operand_map(CUtensorMap* map, const double* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  mdb_encode_tiled_fn enc = mdb_get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Number of local N tiles (from tile 0) whose first global column is < limit.
static int64_t mdb_tiles_below(const GemmArgs& g, int64_t limit, int64_t n_tiles_total) {
  if (g.cyc_G > 1) {
    int64_t lo = 0, hi = n_tiles_total;  // first tile with gcol >= limit
    const int64_t tpb = MDB_NB / gemm::BN;
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      const int64_t gc = ((mid / tpb) * g.cyc_G + g.cyc_r) * MDB_NB + (mid % tpb) * gemm::BN;
      if (gc < limit) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  const int64_t x = limit - g.col_glob0;  // tiles with col_glob0 + nt*64 < limit
  if (x <= 0) return 0;
  return std::min<int64_t>((x + gemm::BN - 1) / gemm::BN, n_tiles_total);
}

// Returns MDB_OK when launched, -1 when the TMA path does not apply (caller falls back to k_gemm_tn).
static int launch_gemm_tn_tma(mdb_ctx* ctx, const GemmArgs& g0, int64_t a_rows, int64_t b_rows, int mode = 0) {
  using namespace gemm;
  if (!ctx->use_tma || g0.K % BK != 0 || g0.lda % 2 != 0 || g0.ldb % 2 != 0) return -1;
  if (((uintptr_t)g0.A & 15) || ((uintptr_t)g0.B & 15)) return -1;
  static uint64_t configured = 0;
  if (!mdb_device_marked(configured, ctx)) {
    if (cudaFuncSetAttribute(k_gemm_tn_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -1;
    cudaFuncSetAttribute(k_gemm_tn_tma, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    mdb_mark_device(configured, ctx);
  }
  CUtensorMap tmA, tmB;
  if (!mdb_make_operand_map(&tmA, g0.A, a_rows, g0.K, g0.lda, BM)) return -1;
  if (!mdb_make_operand_map(&tmB, g0.B, b_rows, g0.K, g0.ldb, BN)) return -1;
  GemmTmaArgs g;
  memset(&g, 0, sizeof(g));
  g.C = g0.C; g.ldc = g0.ldc; g.M = g0.M; g.N = g0.N; g.K = g0.K;
  g.row_glob0 = g0.row_glob0; g.col_glob0 = g0.col_glob0;
  g.m_tile0 = g0.m_tile0; g.n_tile0 = g0.n_tile0;
  g.cyc_G = g0.cyc_G; g.cyc_r = g0.cyc_r; g.b_glob0 = g0.b_glob0;
  g.mode = mode;
  { const char* v = getenv("MDB200_TMA_VARIANT"); g.variant = v ? atoi(v) : 0; }
  const int64_t n_tiles_total = (g.N + BN - 1) / BN;
  const int64_t nt_limit = g0.n_tiles > 0 ? std::min<int64_t>(g0.n_tile0 + g0.n_tiles, n_tiles_total) : n_tiles_total;
  const int64_t m_tiles = (g.M + BM - 1) / BM - g.m_tile0;
  if (m_tiles <= 0 || nt_limit <= g.n_tile0) return MDB_OK;
  const int64_t bands = (m_tiles + 7) / 8;
  if (bands > MDB_MAX_BANDS) return -1;
  int64_t total = 0;
  for (int64_t b = 0; b < bands; ++b) {
    g.band_prefix[b] = (int)total;
    const int64_t row_end = std::min<int64_t>(g.M, (g.m_tile0 + 8 * b + 8) * BM);
    // a tile is needed when its first global column is < (largest global row of the band)
    const int64_t hi = std::min<int64_t>(mdb_tiles_below(g0, g.row_glob0 + row_end - 1, n_tiles_total), nt_limit);
    const int64_t cnt = std::max<int64_t>(0, hi - g.n_tile0);
    total += 8 * cnt;
    if (total > 0x7fffffff) return -1;
  }
  g.band_prefix[bands] = (int)total;
  g.nbands = (int)bands;
  if (total == 0) return MDB_OK;
  k_gemm_tn_tma<<<(unsigned)total, THREADS, SMEM_BYTES, ctx->stream>>>(tmA, tmB, g);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}
