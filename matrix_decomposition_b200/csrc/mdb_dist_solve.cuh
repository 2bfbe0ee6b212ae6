// mdb_dist_solve.cuh -- consumers of the block-cyclic multi-GPU factor ("solve = pipelined block TRSV
// with the same layout (or gather L)").  Included at the end of mdb_api.cu, after mdb_dist.cuh.
//
// (1) Gather: mdb_dist_place_columns writes the column blocks of one rank's finalised slab to their global positions
//     in a full n x n matrix on this device; mdb_factor_upload_device wraps such a matrix in a factor object, so the
//     result of the distributed factorization becomes an ordinary LDL_Decomposition (Reimer.py:810-812).  For sizes
//     whose factor fits one GPU.
//
// (2) Distributed solve  x = A^-1 b = P^T L^-T D^-1 L^-1 P b  (decompositions.py:983-991, dense/util.py:24-29) on the
//     slabs where they are.  The n-vectors (right-hand sides, <= MDB_DSOLVE_KMAX at a time) are replicated; L never
//     moves.  Rows / columns are processed in WINDOWS of MDB_DSOLVE_W = 8 column blocks (1024 columns):
//       * once per factor, the diagonal band  T_w = L[window w, window w]  (n x 1024 doubles in total) is replicated
//         on every rank: each rank writes its own columns into a zeroed buffer (k_dsolve_fill_band), the caller
//         all-reduces it (disjoint supports: the sum is exact), k_dsolve_invert inverts the 128 x 128 diagonal blocks.
//       * forward  L y = b:  t_r = sum over this rank's columns of earlier windows of L[:, J] y_J (partial sums, local).
//         Per window: the caller all-reduces t[window]  ->  every rank solves the band  y_w = T_w^-1 (b_w - t_w)
//         redundantly (same kernel, same data: identical bits everywhere)  ->  t[rows beyond] += L[rows beyond, own
//         columns of w] y_w[own]  (k_dsolve_gemv: streams this rank's part of L once -- the HBM-bound part).
//       * z = y / d.
//       * backward  L^T x = z:  a column of L is wholly local, so  s[c] = sum_{j beyond w} L[j, c] x[j]  is complete on
//         the owner of column c (k_dsolve_colsum + k_dsolve_colsum_reduce, fixed order).  Per window (last to first):
//         each rank fills its own columns of a zeroed 1024-vector, the caller all-reduces it (= all-gather), every
//         rank solves  x_w = T_w^-T (z_w - s_w).
//     One small collective per window and sweep (n / 1024 of them), no other exchange.  Algorithmic traffic per rank:
//     its share of L once per sweep, 8 n^2 / world bytes per solve.
#pragma once

#define MDB_DSOLVE_W 8
#define MDB_DSOLVE_WD (MDB_DSOLVE_W * MDB_NB)
#define MDB_DSOLVE_KMAX 8

// ---- (1) gather ----
__global__ void __launch_bounds__(256) k_dist_place_cols(int64_t n, int rank, int world, const double* __restrict__ slab,
                                                         int64_t ld_slab, int64_t ncols, double* __restrict__ dst,
                                                         int64_t ld_dst) {
  const int64_t i = blockIdx.x;
  const int64_t l0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t lc = l0 + 256 * e;
    if (lc < ncols) {
      const int64_t gc = ((lc / MDB_NB) * world + rank) * MDB_NB + lc % MDB_NB;
      if (gc < n) dst[i * ld_dst + gc] = slab[i * ld_slab + lc];
    }
  }
}

extern "C" int mdb_dist_place_columns(mdb_ctx* ctx, int64_t n, int src_rank, int world, const double* slab,
                                      int64_t ld_slab, double* dst, int64_t ld_dst) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, n >= 0 && world >= 1 && src_rank >= 0 && src_rank < world && slab && dst && ld_dst >= n,
            "mdb_dist_place_columns: bad arguments");
  const int64_t ncols = dist_local_blocks(n, src_rank, world) * NB;
  MDB_CHECK(ctx, ld_slab >= ncols, "mdb_dist_place_columns: ld_slab < local columns of that rank");
  if (n == 0 || ncols == 0) return MDB_OK;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  dim3 grid((unsigned)n, (unsigned)((ncols + 1023) / 1024));
  k_dist_place_cols<<<grid, 256, 0, ctx->stream>>>(n, src_rank, world, slab, ld_slab, ncols, dst, ld_dst);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_factor_upload_device(mdb_ctx* ctx, int64_t n, int type, const double* L_dev, int64_t ldl,
                                        const double* d, const int64_t* p, mdb_factor** out) {
  if (!ctx || !out) return MDB_ERR_INVALID;
  *out = nullptr;
  MDB_CHECK(ctx, type == MDB_TYPE_LDL || type == MDB_TYPE_LL, "mdb_factor_upload_device: type must be LDL or LL");
  MDB_CHECK(ctx, (L_dev || n == 0) && ldl >= n, "mdb_factor_upload_device: bad L");
  MDB_CHECK(ctx, type == MDB_TYPE_LL || d || n == 0, "mdb_factor_upload_device: LDL needs d");
  mdb_factor* f = nullptr;
  int rc = mdb_factor_create(ctx, n, 1, &f);
  if (rc) return rc;
  rc = factor_set_p(f, p);
  cudaError_t e = cudaSuccess;
  if (!rc && n > 0)
    e = cudaMemcpy2DAsync(f->W, (size_t)f->ld * 8, L_dev, (size_t)ldl * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToDevice,
                          ctx->stream);
  if (!rc && e == cudaSuccess && n > 0 && d) e = cudaMemcpyAsync(f->d, d, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (!rc && e != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "%s", cudaGetErrorString(e));
  if (rc) { mdb_factor_destroy(f); return rc; }
  f->state = ST_FINAL;
  f->out_type = type;
  *out = f;
  return MDB_OK;
}

// ---- (2) distributed solve ----
extern "C" int64_t mdb_dist_solve_band_elems(const mdb_dist* s) { return s ? s->n * (int64_t)MDB_DSOLVE_WD : 0; }
// the inverses of the diagonal blocks, row-major, followed by their transposes (the backward sweep reads those)
extern "C" int64_t mdb_dist_solve_tinv_elems(const mdb_dist* s) { return s ? 2 * s->nblk_glob * (int64_t)NB * NB : 0; }
extern "C" int64_t mdb_dist_solve_windows(const mdb_dist* s) { return s ? (s->n + MDB_DSOLVE_WD - 1) / MDB_DSOLVE_WD : 0; }
extern "C" int64_t mdb_dist_solve_partial_elems(const mdb_dist* s) {
  return s ? (int64_t)s->ctx->num_sms * 2 * MDB_DSOLVE_WD * MDB_DSOLVE_KMAX : 0;
}

// band[g][c] = L[g, w(g) * WD + c] for this rank's columns (everything else is left as the caller zeroed it)
__global__ void __launch_bounds__(256) k_dsolve_fill_band(int64_t n, const double* __restrict__ S, int64_t ld, int rank,
                                                          int world, int64_t nblk_loc, double* __restrict__ band) {
  const int64_t g = blockIdx.x;
  const int64_t w = g / MDB_DSOLVE_WD;
  for (int c = threadIdx.x; c < MDB_DSOLVE_WD; c += 256) {
    const int64_t gc = w * MDB_DSOLVE_WD + c;
    const int64_t J = gc / MDB_NB;
    if (gc < n && (J % world) == rank) {
      const int64_t lc = (J / world) * MDB_NB + gc % MDB_NB;
      band[g * MDB_DSOLVE_WD + c] = S[g * ld + lc];
    }
  }
}

extern "C" int mdb_dist_solve_fill_band(mdb_dist* s, double* band) {
  if (!s || !band) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_CUDA(ctx, cudaMemsetAsync(band, 0, (size_t)s->n * MDB_DSOLVE_WD * 8, ctx->stream));
  k_dsolve_fill_band<<<(unsigned)s->n, 256, 0, ctx->stream>>>(s->n, s->S, s->ld, s->rank, s->world, s->nblk_loc, band);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// Inverse of the unit lower triangular 128 x 128 diagonal block J (rows / columns past n: identity).  One CTA per
// block, thread c computes column c of the inverse by forward substitution.  The inverse lives in shared memory
// ([128][129] doubles); T is read from the band in global memory (every thread walks the same row i at the same
// time, so the loads are L1-friendly).
__global__ void __launch_bounds__(MDB_NB) k_dsolve_invert(int64_t n, const double* __restrict__ band,
                                                          double* __restrict__ tinv, double* __restrict__ tinvT) {
  extern __shared__ double smem_d[];
  constexpr int LDS = MDB_NB + 1;
  double* X = smem_d;               // X[i * LDS + c]: column c of the inverse, owned by thread c
  const int64_t J = blockIdx.x, g0 = J * MDB_NB;
  const int c = threadIdx.x;
  const int cw = (int)(g0 % MDB_DSOLVE_WD);   // first column of the block inside its window
  for (int i = 0; i < MDB_NB; ++i) {
    const int64_t g = g0 + i;
    const double* Trow = band + g * MDB_DSOLVE_WD + cw;   // T[i][m] = Trow[m] for m < i (valid when g < n)
    double x;
    if (i < c) x = 0.0;
    else if (i == c) x = 1.0;
    else if (g >= n) x = 0.0;
    else {
      double a0 = 0, a1 = 0;
      int m = c;
      for (; m + 1 < i; m += 2) {
        a0 = fma(Trow[m], X[m * LDS + c], a0);
        a1 = fma(Trow[m + 1], X[(m + 1) * LDS + c], a1);
      }
      if (m < i) a0 = fma(Trow[m], X[m * LDS + c], a0);
      x = -(a0 + a1);
    }
    X[i * LDS + c] = x;
  }
  __syncthreads();
  double* out = tinv + J * MDB_NB * MDB_NB;
  double* outT = tinvT + J * MDB_NB * MDB_NB;
  for (int i = 0; i < MDB_NB; ++i) {
    out[i * MDB_NB + c] = X[i * LDS + c];
    outT[i * MDB_NB + c] = X[c * LDS + i];
  }
}

extern "C" int mdb_dist_solve_invert_band(mdb_dist* s, const double* band, double* tinv) {
  if (!s || !band || !tinv) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  static uint64_t configured = 0;
  const int smem = NB * (NB + 1) * 8;
  if (!mdb_device_marked(configured, ctx)) {
    MDB_CUDA(ctx, cudaFuncSetAttribute(k_dsolve_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mdb_mark_device(configured, ctx);
  }
  k_dsolve_invert<<<(unsigned)s->nblk_glob, NB, smem, ctx->stream>>>(s->n, band, tinv, tinv + s->nblk_glob * (int64_t)NB * NB);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// warp-level sum of K accumulators (fixed butterfly order: deterministic)
template <int K>
__device__ __forceinline__ void dsolve_warp_sum(double (&a)[K]) {
#pragma unroll
  for (int q = 0; q < K; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
  }
}

// The band solve of one window, one CTA of 1024 threads (32 warps), right-hand sides r (n x k, row-major) in place.
//   forward : r_w <- T_w^-1 (r_w - t_w)          (t = the all-reduced partial sums; null = nothing to subtract)
//   backward: r_w <- T_w^-T (r_w - s)            (s = the all-reduced column sums of the window, WD x k)
// Blocks of 128 inside the window in order (forward) / reverse order (backward): the diagonal block is applied through
// its pre-computed inverse, the coupling to the remaining blocks of the window is a rank-128 update of their entries.
template <int K>
__global__ void __launch_bounds__(1024) k_dsolve_band(int64_t n, int64_t w, const double* __restrict__ band,
                                                      const double* __restrict__ tinv, double* __restrict__ r,
                                                      const double* __restrict__ sub, int sub_is_window, int backward) {
  extern __shared__ double smem_d[];
  double* v = smem_d;                       // WD x K: the window's right-hand sides -> solution
  double* y = v + MDB_DSOLVE_WD * K;        // NB x K: the block being applied
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t g0 = w * MDB_DSOLVE_WD;
  const int rows = (int)((n - g0 < MDB_DSOLVE_WD) ? n - g0 : MDB_DSOLVE_WD);
  const int nblk = (rows + MDB_NB - 1) / MDB_NB;
  for (int idx = tid; idx < MDB_DSOLVE_WD * K; idx += 1024) {
    const int i = idx / K, q = idx % K;
    double x = 0;
    if (i < rows) {
      x = r[(g0 + i) * K + q];
      if (sub) x -= sub_is_window ? sub[(int64_t)i * K + q] : sub[(g0 + i) * K + q];
    }
    v[idx] = x;
  }
  __syncthreads();
  for (int step = 0; step < nblk; ++step) {
    const int J = backward ? nblk - 1 - step : step;
    const double* Ti = tinv + (g0 / MDB_NB + J) * (int64_t)MDB_NB * MDB_NB;
    // y = Tinv v_J (forward) or Tinv^T v_J (backward): warp per output row, lanes over the 128 terms
    for (int i = warp; i < MDB_NB; i += 32) {
      double a[K];
#pragma unroll
      for (int q = 0; q < K; ++q) a[q] = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = lane + 32 * e;
        const double t = Ti[i * MDB_NB + m];   // (the caller passes the transposed inverses for the backward sweep)
#pragma unroll
        for (int q = 0; q < K; ++q) a[q] = fma(t, v[(J * MDB_NB + m) * K + q], a[q]);
      }
      dsolve_warp_sum<K>(a);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < K; ++q) y[i * K + q] = a[q];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < MDB_NB * K; idx += 1024) v[J * MDB_NB * K + idx] = y[idx];
    // coupling: forward  v_i -= sum_c T[i, J cols c] y_c  for rows i of the later blocks   (warp per row, lanes over c)
    //           backward v_c -= sum_i T[J rows i, c] y_i  for columns c of the earlier blocks (thread per column)
    if (!backward) {
      for (int i = (J + 1) * MDB_NB + warp; i < rows; i += 32) {
        const double* Trow = band + (g0 + i) * MDB_DSOLVE_WD + J * MDB_NB;
        double a[K];
#pragma unroll
        for (int q = 0; q < K; ++q) a[q] = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = lane + 32 * e;
          const double t = Trow[c];
#pragma unroll
          for (int q = 0; q < K; ++q) a[q] = fma(t, y[c * K + q], a[q]);
        }
        dsolve_warp_sum<K>(a);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < K; ++q) v[i * K + q] -= a[q];
        }
      }
    } else {
      const int ncol = J * MDB_NB;
      const int jr = (rows - J * MDB_NB < MDB_NB) ? rows - J * MDB_NB : MDB_NB;
      for (int c = tid; c < ncol; c += 1024) {
        double a[K];
#pragma unroll
        for (int q = 0; q < K; ++q) a[q] = 0;
        const double* Tcol = band + (g0 + J * MDB_NB) * MDB_DSOLVE_WD + c;
        for (int i = 0; i < jr; ++i) {
          const double t = Tcol[(int64_t)i * MDB_DSOLVE_WD];
#pragma unroll
          for (int q = 0; q < K; ++q) a[q] = fma(t, y[i * K + q], a[q]);
        }
#pragma unroll
        for (int q = 0; q < K; ++q) v[c * K + q] -= a[q];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < rows * K; idx += 1024) r[g0 * K + idx] = v[idx];
}

// t[j, :] += sum_c L[j, lc0 + c] y[gcol(c), :]  for rows j in [row0, n): this rank's columns of one window, m of them
// (contiguous in the slab).  One warp per row, lanes over the columns (coalesced), y staged in shared memory.
template <int K>
__global__ void __launch_bounds__(256) k_dsolve_gemv(int64_t n, int64_t row0, const double* __restrict__ S, int64_t ld,
                                                     int64_t lc0, int m, int rank, int world,
                                                     const double* __restrict__ r, double* __restrict__ t) {
  extern __shared__ double ys[];   // m x K
  for (int idx = threadIdx.x; idx < m * K; idx += 256) {
    const int c = idx / K, q = idx % K;
    const int64_t lc = lc0 + c;
    const int64_t gc = ((lc / MDB_NB) * world + rank) * MDB_NB + lc % MDB_NB;
    ys[idx] = (gc < n) ? r[gc * K + q] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t j = row0 + (int64_t)blockIdx.x * 8 + warp; j < n; j += (int64_t)gridDim.x * 8) {
    const double* Lrow = S + j * ld + lc0;
    double a[K];
#pragma unroll
    for (int q = 0; q < K; ++q) a[q] = 0;
    for (int c = lane; c < m; c += 32) {
      const double l = Lrow[c];
#pragma unroll
      for (int q = 0; q < K; ++q) a[q] = fma(l, ys[c * K + q], a[q]);
    }
    dsolve_warp_sum<K>(a);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < K; ++q) t[j * K + q] += a[q];
    }
  }
}

// partial[chunk][c][q] = sum over the rows of the chunk of L[j, lc0 + c] x[j, q]; thread per column (coalesced rows)
template <int K>
__global__ void __launch_bounds__(256) k_dsolve_colsum(int64_t n, int64_t row0, int64_t rows_per_chunk,
                                                       const double* __restrict__ S, int64_t ld, int64_t lc0, int m,
                                                       const double* __restrict__ r, double* __restrict__ partial) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  const int64_t j0 = row0 + (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t j1 = (j0 + rows_per_chunk < n) ? j0 + rows_per_chunk : n;
  double a[K];
#pragma unroll
  for (int q = 0; q < K; ++q) a[q] = 0;
  if (c < m) {
    for (int64_t j = j0; j < j1; ++j) {
      const double l = S[j * ld + lc0 + c];
#pragma unroll
      for (int q = 0; q < K; ++q) a[q] = fma(l, r[j * K + q], a[q]);
    }
#pragma unroll
    for (int q = 0; q < K; ++q) partial[((int64_t)blockIdx.x * m + c) * K + q] = a[q];
  }
}

// xchg (WD x K, zero elsewhere) <- chunks summed in order, at the columns' positions inside the window
template <int K>
__global__ void __launch_bounds__(256) k_dsolve_colsum_reduce(int64_t n, int64_t w, int nchunks, int m, int64_t lc0,
                                                              int rank, int world, const double* __restrict__ partial,
                                                              double* __restrict__ xchg) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= m * K) return;
  const int c = idx / K, q = idx % K;
  double a = 0;
  for (int ch = 0; ch < nchunks; ++ch) a += partial[((int64_t)ch * m + c) * K + q];
  const int64_t lc = lc0 + c;
  const int64_t gc = ((lc / MDB_NB) * world + rank) * MDB_NB + lc % MDB_NB;
  if (gc < n) xchg[(gc - w * MDB_DSOLVE_WD) * K + q] = a;
}

// r[i, :] /= d[i]; a zero d raises *flag (the reference raises DecompositionSingularError before solving)
__global__ void k_dsolve_scale(int64_t n, int k, const double* __restrict__ d, double* __restrict__ r) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n * k) r[idx] = r[idx] / d[idx / k];
}

template <int K>
static int dsolve_configure(mdb_ctx* ctx) {
  static uint64_t configured = 0;
  if (mdb_device_marked(configured, ctx)) return MDB_OK;
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_dsolve_band<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (MDB_DSOLVE_WD + MDB_NB) * K * 8));
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_dsolve_gemv<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, MDB_DSOLVE_WD * K * 8));
  mdb_mark_device(configured, ctx);
  return MDB_OK;
}

// local columns of window w: [lc0, lc0 + m)
static inline void dsolve_window_cols(const mdb_dist* s, int64_t w, int64_t* lc0, int* m) {
  const int64_t J0 = w * MDB_DSOLVE_W, J1 = std::min<int64_t>(J0 + MDB_DSOLVE_W, s->nblk_glob);
  auto first_local = [&](int64_t J) { return std::max<int64_t>((J - s->rank + s->world - 1) / s->world, 0); };
  const int64_t lb0 = first_local(J0), lb1 = std::min<int64_t>(first_local(J1), s->nblk_loc);
  *lc0 = lb0 * NB;
  *m = (int)(std::max<int64_t>(lb1 - lb0, 0) * NB);
}

#define MDB_DSOLVE_DISPATCH(k, ...)                          \
  switch (k) {                                               \
    case 1: { constexpr int KK = 1; __VA_ARGS__; } break;    \
    case 2: { constexpr int KK = 2; __VA_ARGS__; } break;    \
    case 4: { constexpr int KK = 4; __VA_ARGS__; } break;    \
    case 8: { constexpr int KK = 8; __VA_ARGS__; } break;    \
    default: return mdb_fail(ctx, MDB_ERR_INVALID, "%s", "distributed solve: nrhs per pass must be 1, 2, 4 or 8"); \
  }

// forward step of window w.  r: n x k right-hand sides -> y (in place); t: n x k partial sums whose slice of window w
// has been all-reduced by the caller.
extern "C" int mdb_dist_solve_fwd_window(mdb_dist* s, int64_t w, const double* band, const double* tinv, double* r,
                                         double* t, int k) {
  if (!s || !band || !tinv || !r || !t) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, w >= 0 && w < mdb_dist_solve_windows(s), "mdb_dist_solve_fwd_window: bad window");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n, row0 = std::min<int64_t>((w + 1) * MDB_DSOLVE_WD, n);
  int64_t lc0; int m;
  dsolve_window_cols(s, w, &lc0, &m);
  MDB_DSOLVE_DISPATCH(k, {
    const int rcc = dsolve_configure<KK>(ctx);
    if (rcc) return rcc;
    k_dsolve_band<KK><<<1, 1024, (MDB_DSOLVE_WD + MDB_NB) * KK * 8, ctx->stream>>>(n, w, band, tinv, r, t, 0, 0);
    MDB_LAUNCHED(ctx);
    if (m > 0 && row0 < n) {
      const unsigned grid = (unsigned)std::min<int64_t>((n - row0 + 7) / 8, (int64_t)ctx->num_sms * 8);
      k_dsolve_gemv<KK><<<grid, 256, (size_t)m * KK * 8, ctx->stream>>>(n, row0, s->S, s->ld, lc0, m, s->rank, s->world, r, t);
      MDB_LAUNCHED(ctx);
    }
  });
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_dist_solve_scale(mdb_dist* s, double* r, int k) {
  if (!s || !r || k < 1) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t tot = s->n * k;
  k_dsolve_scale<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(s->n, k, s->d, r);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// backward, first half of window w: xchg (WD x k) <- zero, then this rank's columns of the window get their complete
// sums over the rows beyond the window.  The caller all-reduces xchg, then calls mdb_dist_solve_bwd_window.
extern "C" int mdb_dist_solve_bwd_colsum(mdb_dist* s, int64_t w, const double* r, double* xchg, double* partial, int k) {
  if (!s || !r || !xchg || !partial) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, w >= 0 && w < mdb_dist_solve_windows(s), "mdb_dist_solve_bwd_colsum: bad window");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n, row0 = std::min<int64_t>((w + 1) * MDB_DSOLVE_WD, n);
  int64_t lc0; int m;
  dsolve_window_cols(s, w, &lc0, &m);
  MDB_CUDA(ctx, cudaMemsetAsync(xchg, 0, (size_t)MDB_DSOLVE_WD * k * 8, ctx->stream));
  if (m > 0 && row0 < n) {
    const int64_t max_chunks = (int64_t)ctx->num_sms * 2;
    const int64_t rpc = std::max<int64_t>(64, (n - row0 + max_chunks - 1) / max_chunks);
    const int nchunks = (int)((n - row0 + rpc - 1) / rpc);
    MDB_DSOLVE_DISPATCH(k, {
      k_dsolve_colsum<KK><<<dim3((unsigned)nchunks, (unsigned)((m + 255) / 256)), 256, 0, ctx->stream>>>(
          n, row0, rpc, s->S, s->ld, lc0, m, r, partial);
      MDB_LAUNCHED(ctx);
      k_dsolve_colsum_reduce<KK><<<(unsigned)((m * KK + 255) / 256), 256, 0, ctx->stream>>>(n, w, nchunks, m, lc0, s->rank,
                                                                                          s->world, partial, xchg);
      MDB_LAUNCHED(ctx);
    });
  }
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_dist_solve_bwd_window(mdb_dist* s, int64_t w, const double* band, const double* tinv, double* r,
                                         const double* xchg, int k) {
  if (!s || !band || !tinv || !r || !xchg) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, w >= 0 && w < mdb_dist_solve_windows(s), "mdb_dist_solve_bwd_window: bad window");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_DSOLVE_DISPATCH(k, {
    const int rcc = dsolve_configure<KK>(ctx);
    if (rcc) return rcc;
    k_dsolve_band<KK><<<1, 1024, (MDB_DSOLVE_WD + MDB_NB) * KK * 8, ctx->stream>>>(
        s->n, w, band, tinv + s->nblk_glob * (int64_t)NB * NB, r, xchg, 1, 1);
    MDB_LAUNCHED(ctx);
  });
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}
