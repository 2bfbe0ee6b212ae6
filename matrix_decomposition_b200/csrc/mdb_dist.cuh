// mdb_dist.cuh -- per-rank device steps of the 1-D block-cyclic multi-GPU factorization.
// Included at the end of mdb_api.cu.  The loop, the buffers and the NCCL broadcast live in
// matrix_decomposition_b200/distributed.py; see include/mdb200.h for the contract.
#pragma once

struct mdb_dist {
  mdb_ctx* ctx;
  int64_t n, ld;       // ld = number of local columns (multiple of NB)
  int rank, world;
  int64_t nblk_glob, nblk_loc;
  double *S, *A;       // n x ld, owned by the caller (torch tensors)
  double* panel[2];    // broadcast buffers, owned by the caller
  double *OPX[2], *OPY[2];  // operands of one outer block, (n x ldp) each, double buffered by outer-block parity
  int64_t ldp;         // = (widest outer block) * NB
  double* vec;         // 7 x n: gamma, alpha, beta, rowmax, d, omega, delta (position order, full length)
  double *gamma, *alpha, *beta, *rowmax, *d, *omega, *delta;
  uint8_t* clamped;
  double *minB, *maxB;
  int* info;
  double* Ut;
  double* Vinv;
  int64_t* pcol_dev;   // scratch for the generator
  MdRuleParams prm;
  bool have_bounds;
};

static inline int64_t dist_local_blocks(int64_t n, int rank, int world) {
  const int64_t nblk = (n + NB - 1) / NB;
  return nblk > rank ? (nblk - rank + world - 1) / world : 0;
}

extern "C" int64_t mdb_dist_local_cols(int64_t n, int rank, int world) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world) return -1;
  return std::max<int64_t>(dist_local_blocks(n, rank, world), 1) * NB;
}

// header 5 x NB | vectors 3 x rows | X rows x NB   (Y = -X D is rebuilt by the receivers)
static inline int64_t dist_panel_elems(int64_t n, int64_t k) {
  const int64_t k1 = std::min<int64_t>((k + 1) * NB, n);
  const int64_t rows = n - k1;
  return 5 * NB + 3 * rows + rows * NB;
}

// number of panels in the outer block that starts at panel k (same rule as mdb_factor_run)
static inline int64_t dist_outer_width(const mdb_ctx* ctx, int64_t n, int64_t k) {
  const int64_t remaining = n - k * NB;
  int64_t a = 1;
  if (ctx->agg_fixed) a = ctx->agg_fixed;
  else if (remaining >= ctx->agg_thresh[2]) a = 8;
  else if (remaining >= ctx->agg_thresh[1]) a = 4;
  else if (remaining >= ctx->agg_thresh[0]) a = 2;
  const int64_t left = (n + NB - 1) / NB - k;
  return std::max<int64_t>(1, std::min<int64_t>(a, left));
}
extern "C" int64_t mdb_dist_outer_width(const mdb_dist* s, int64_t k) {
  return (s && k >= 0 && k < s->nblk_glob) ? std::min<int64_t>(dist_outer_width(s->ctx, s->n, k), s->ldp / NB) : 0;
}
extern "C" int64_t mdb_dist_panel_elems(const mdb_dist* s, int64_t k) { return s ? dist_panel_elems(s->n, k) : 0; }

extern "C" int mdb_dist_create(mdb_ctx* ctx, int64_t n, int rank, int world, double* S_loc, double* A_loc,
                               int64_t ld_loc, double* panel0, double* panel1, mdb_dist** out) {
  if (!ctx || !out) return MDB_ERR_INVALID;
  *out = nullptr;
  MDB_CHECK(ctx, n > 0 && world >= 1 && rank >= 0 && rank < world, "mdb_dist_create: bad n / rank / world");
  MDB_CHECK(ctx, S_loc && A_loc && panel0 && panel1, "mdb_dist_create: null buffer");
  MDB_CHECK(ctx, ld_loc == mdb_dist_local_cols(n, rank, world), "mdb_dist_create: ld_loc must be mdb_dist_local_cols()");
  ctx_reap(ctx);
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = configure_kernels(ctx);
  if (rc) return rc;
  mdb_dist* s = new (std::nothrow) mdb_dist();
  if (!s) return MDB_ERR_INVALID;
  s->ctx = ctx; s->n = n; s->ld = ld_loc; s->rank = rank; s->world = world;
  s->nblk_glob = (n + NB - 1) / NB;
  s->nblk_loc = dist_local_blocks(n, rank, world);
  s->S = S_loc; s->A = A_loc; s->panel[0] = panel0; s->panel[1] = panel1;
  s->vec = nullptr; s->clamped = nullptr; s->minB = s->maxB = nullptr; s->info = nullptr; s->Ut = nullptr; s->Vinv = nullptr;
  s->pcol_dev = nullptr; s->have_bounds = false;
  s->OPX[0] = s->OPX[1] = s->OPY[0] = s->OPY[1] = nullptr;
  s->ldp = dist_outer_width(ctx, n, 0) * NB;  // the first outer block is the widest
  cudaError_t e = dev_malloc(ctx, (void**)&s->vec, (size_t)7 * n * 8);
  for (int i = 0; i < 2; ++i) {
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->OPX[i], (size_t)n * s->ldp * 8);
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->OPY[i], (size_t)n * s->ldp * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->OPX[i], 0, (size_t)n * s->ldp * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->OPY[i], 0, (size_t)n * s->ldp * 8, ctx->stream);
  }
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->clamped, (size_t)n);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->info, sizeof(int));
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->Ut, (size_t)NB * NB * 8);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->Vinv, (size_t)MDB_VINV_DOUBLES * 8);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&s->pcol_dev, (size_t)ld_loc * 8);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->vec, 0, (size_t)7 * n * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->clamped, 0, (size_t)n, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->info, 0, sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->Ut, 0, (size_t)NB * NB * 8, ctx->stream);
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "mdb_dist_create: %s", cudaGetErrorString(e));
    mdb_dist_destroy(s);
    return MDB_ERR_CUDA;
  }
  s->gamma = s->vec; s->alpha = s->vec + n; s->beta = s->vec + 2 * n; s->rowmax = s->vec + 3 * n;
  s->d = s->vec + 4 * n; s->omega = s->vec + 5 * n; s->delta = s->vec + 6 * n;
  *out = s;
  return MDB_OK;
}

extern "C" int mdb_dist_destroy(mdb_dist* s) {
  if (!s) return MDB_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->vec); cudaFree(s->clamped);
  for (int i = 0; i < 2; ++i) { cudaFree(s->OPX[i]); cudaFree(s->OPY[i]); }
  cudaFree(s->minB); cudaFree(s->maxB); cudaFree(s->info); cudaFree(s->Ut); cudaFree(s->Vinv);
  cudaFree(s->pcol_dev);
  delete s;
  return MDB_OK;
}

// gamma of the locally owned columns from A_loc: gamma[gc] = A_loc[gc, lc]
__global__ void k_dist_extract_gamma(int64_t n, const double* __restrict__ A, int64_t ld, int64_t ncols, int rank,
                                     int world, double* __restrict__ gamma) {
  const int64_t lc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lc < ncols) {
    const int64_t gc = ((lc / MDB_NB) * world + rank) * MDB_NB + lc % MDB_NB;
    if (gc < n) gamma[gc] = A[gc * ld + lc];
  }
}

static int dist_after_matrix(mdb_dist* s) {
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaMemcpyAsync(s->S, s->A, (size_t)s->n * s->ld * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  k_dist_extract_gamma<<<(unsigned)((s->ld + 255) / 256), 256, 0, ctx->stream>>>(s->n, s->A, s->ld, s->ld, s->rank,
                                                                                s->world, s->gamma);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_dist_generate_f2(mdb_dist* s, uint64_t seed, double mu, const int64_t* p) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n;
  // original index of every local column (columns past n map to themselves; they are never read)
  std::vector<int64_t> pcol((size_t)s->ld);
  for (int64_t lc = 0; lc < s->ld; ++lc) {
    const int64_t gc = ((lc / NB) * s->world + s->rank) * NB + lc % NB;
    pcol[(size_t)lc] = (gc < n) ? (p ? p[gc] : gc) : gc;
  }
  int64_t* prow = nullptr;
  if (p) {
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&prow, (size_t)n * 8));
    MDB_CUDA(ctx, cudaMemcpyAsync(prow, p, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  MDB_CUDA(ctx, cudaMemcpyAsync(s->pcol_dev, pcol.data(), (size_t)s->ld * 8, cudaMemcpyHostToDevice, ctx->stream));
  int rc = generate_f2(ctx, n, s->ld, 0, prow, s->pcol_dev, seed, mu, s->A, s->ld);
  if (!rc) rc = dist_after_matrix(s);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(prow);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

extern "C" int mdb_dist_set_matrix(mdb_dist* s, const double* Ap, int64_t ld) {
  if (!s || !Ap) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, ld >= s->n, "mdb_dist_set_matrix: ld < n");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_CUDA(ctx, cudaMemsetAsync(s->A, 0, (size_t)s->n * s->ld * 8, ctx->stream));
  for (int64_t lb = 0; lb < s->nblk_loc; ++lb) {
    const int64_t gc0 = (lb * s->world + s->rank) * NB;
    const int64_t w = std::min<int64_t>(NB, s->n - gc0);
    int rc = copy2d_async(ctx, s->A + lb * NB, s->ld, Ap + gc0, ld, s->n, w, MDB_DEVICE, MDB_HOST);
    if (rc) return rc;
  }
  int rc = dist_after_matrix(s);
  if (rc) return rc;
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

// this rank's column slab from a host (or device) buffer that already has the slab layout: A_loc_src is n x ld_src
// with local column lc at A_loc_src[i * ld_src + lc]  (the reverse of reading S_loc after mdb_dist_finalize)
static int dist_set_local(mdb_dist* s, const double* A_loc_src, int64_t ld_src, int loc, bool lower_only) {
  if (!s || !A_loc_src) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, ld_src >= s->ld, "mdb_dist_set_local: ld_src < local columns");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = MDB_OK;
  if (!lower_only) {
    rc = copy2d_async(ctx, s->A, s->ld, A_loc_src, ld_src, s->n, s->ld, MDB_DEVICE, loc);
  } else {
    // Per column block only the rows from its own diagonal block downwards are ever read by the factorization (the
    // diagonal block with both its triangles, the rows below it; the part above belongs to finished rows of L): that
    // is all that crosses PCIe -- half of the slab.  The rest of A_loc is zero afterwards.
    MDB_CUDA(ctx, cudaMemsetAsync(s->A, 0, (size_t)s->n * s->ld * 8, ctx->stream));
    for (int64_t lb = 0; lb < s->nblk_loc && !rc; ++lb) {
      const int64_t gc0 = (lb * s->world + s->rank) * NB;
      const int64_t w = std::min<int64_t>(NB, s->n - gc0);
      rc = copy2d_async(ctx, s->A + gc0 * s->ld + lb * NB, s->ld, A_loc_src + gc0 * ld_src + lb * NB, ld_src, s->n - gc0, w,
                        MDB_DEVICE, loc);
    }
  }
  if (!rc) rc = dist_after_matrix(s);
  if (rc) return rc;
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

extern "C" int mdb_dist_set_local(mdb_dist* s, const double* A_loc_src, int64_t ld_src, int loc) {
  return dist_set_local(s, A_loc_src, ld_src, loc, false);
}
extern "C" int mdb_dist_set_local_lower(mdb_dist* s, const double* A_loc_src, int64_t ld_src, int loc) {
  return dist_set_local(s, A_loc_src, ld_src, loc, true);
}

extern "C" int mdb_dist_set_bounds(mdb_dist* s, const mdb_bounds* bnd, const int64_t* p) {
  if (!s || !bnd) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, bnd->rule == MDB_RULE_REIMER || bnd->rule == MDB_RULE_EXACT, "mdb_dist_set_bounds: unknown rule");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n;
  if (s->minB) { cudaFree(s->minB); s->minB = nullptr; }
  if (s->maxB) { cudaFree(s->maxB); s->maxB = nullptr; }
  double sc_minB = -INFINITY, sc_maxB = INFINITY;
  if (bnd->rule == MDB_RULE_REIMER) {
    if (bnd->min_diag_B) sc_minB = bnd->min_diag_B[0];
    if (bnd->max_diag_B) sc_maxB = bnd->max_diag_B[0];
    if (bnd->stride_B != 0) {
      MDB_CHECK(ctx, bnd->min_diag_B && bnd->max_diag_B, "vector bounds need both min_diag_B and max_diag_B");
      std::vector<double> mb((size_t)n), Mb((size_t)n);
      for (int64_t i = 0; i < n; ++i) {
        const int64_t pi = p ? p[i] : i;
        mb[(size_t)i] = bnd->min_diag_B[pi * bnd->stride_B];
        Mb[(size_t)i] = bnd->max_diag_B[pi * bnd->stride_B];
      }
      MDB_CUDA(ctx, dev_malloc(ctx, (void**)&s->minB, (size_t)n * 8));
      MDB_CUDA(ctx, dev_malloc(ctx, (void**)&s->maxB, (size_t)n * 8));
      MDB_CUDA(ctx, cudaMemcpyAsync(s->minB, mb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
      MDB_CUDA(ctx, cudaMemcpyAsync(s->maxB, Mb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
      MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  s->prm.rule = bnd->rule;
  s->prm.min_diag_D = bnd->min_diag_D;
  s->prm.max_diag_D = bnd->max_diag_D;
  s->prm.min_abs_value_D = bnd->min_abs_value_D;
  s->prm.min_abs_value_L = (bnd->rule == MDB_RULE_EXACT) ? 0.0 : bnd->min_abs_value_L;
  s->prm.min_diag_B = sc_minB;
  s->prm.max_diag_B = sc_maxB;
  MDB_CUDA(ctx, cudaMemsetAsync(s->alpha, 0, (size_t)3 * n * 8, ctx->stream));
  MDB_CUDA(ctx, cudaMemsetAsync(s->info, 0, sizeof(int), ctx->stream));
  ctx->gemm_flop[0] = ctx->gemm_flop[1] = 0;  // mdb_ctx_get_update_stats counts this rank's share from here on
  ctx->gemm_launches[0] = ctx->gemm_launches[1] = 0;
  s->have_bounds = true;
  return MDB_OK;
}

__global__ void k_dist_pack(int64_t rows, const double* __restrict__ OX, int64_t ldp, const double* __restrict__ alpha,
                            const double* __restrict__ beta, const double* __restrict__ rowmax, double* __restrict__ vecs,
                            double* __restrict__ MX);

static FactorDev dist_make_dev(mdb_dist* s, double* hdr) {
  FactorDev F;
  memset(&F, 0, sizeof(F));
  F.n = s->n; F.ldw = s->ld; F.W = s->S;
  F.gamma = s->gamma; F.alpha = s->alpha; F.beta = s->beta; F.rowmax = s->rowmax;
  F.d = s->d; F.omega = s->omega; F.delta = s->delta; F.clamped = s->clamped;
  F.minB = s->minB; F.maxB = s->maxB; F.info = s->info; F.Ut = s->Ut; F.hdr = hdr; F.Vinv = s->Vinv;
  F.prm = s->prm;
  return F;
}

extern "C" int mdb_dist_factor_panel(mdb_dist* s, int64_t k, int slot, int64_t k_outer, int oslot) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, s->have_bounds, "mdb_dist_factor_panel: call mdb_dist_set_bounds first");
  MDB_CHECK(ctx, k >= 0 && k < s->nblk_glob && (k % s->world) == s->rank, "mdb_dist_factor_panel: not the owner of panel k");
  MDB_CHECK(ctx, (slot == 0 || slot == 1) && (oslot == 0 || oslot == 1), "bad slot");
  MDB_CHECK(ctx, k_outer >= 0 && k_outer <= k && (k - k_outer + 1) * NB <= s->ldp, "mdb_dist_factor_panel: bad outer block");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n, k0 = k * NB, k1 = std::min<int64_t>(k0 + NB, n), rows = n - k1;
  const int kb = (int)(k1 - k0);
  const int64_t lc0 = (k / s->world) * NB;
  const int64_t ob = k_outer * NB;
  double* buf = s->panel[slot];
  double* hdr = buf;
  double* vecs = buf + 5 * NB;
  double* MX = vecs + 3 * rows;  // X part of the message
  FactorDev F = dist_make_dev(s, hdr);
  const int diag_smem = MDB_DIAG_SMEM_DOUBLES(NB) * 8;
  const int panel_smem = MDB_PANEL_SMEM_DOUBLES(NB) * 8;
  auto kdiag = (F.prm.rule == MD_RULE_EXACT) ? k_diag_block<NB, true> : k_diag_block<NB, false>;
  kdiag<<<1, MDB_DIAG_THREADS(NB), diag_smem, ctx->stream>>>(F, k0, kb, s->S + k0 * s->ld + lc0, s->ld, s->A + k0 * s->ld + lc0, s->ld);
  MDB_LAUNCHED(ctx);
  if (rows > 0) {
    dim3 grid((unsigned)((rows + PANEL_R - 1) / PANEL_R), 1, 1);
    // A[j, k0 + c] = A_loc[j * ld + lc0 + c]  ->  Asrc + j * a_rs + k0 * a_cs with a_rs = ld, a_cs = 1.
    // The owner's operands go straight into the outer-block buffers; X is then copied into the message.
    double* PX = s->OPX[oslot] + (k0 - ob);
    double* PY = s->OPY[oslot] + (k0 - ob);
    k_panel_gemm<NB, true><<<grid, MDB_PANEL_THREADS, panel_smem, ctx->stream>>>(F, k0, k1, rows, s->A + lc0 - k0, s->ld, 1,
                                                                     s->S + lc0, s->ld, PX, PY, ob, s->ldp);
    MDB_LAUNCHED(ctx);
    {  // message = [hdr (already there) | alpha, beta, rowmax of the rows below | X], packed by one kernel
      const unsigned grid = (unsigned)std::min<int64_t>((rows + 1) / 2, (int64_t)ctx->num_sms * 8);
      k_dist_pack<<<grid, 256, 0, ctx->stream>>>(rows, PX + (k1 - ob) * s->ldp, s->ldp, s->alpha + k1, s->beta + k1,
                                                 s->rowmax + k1, vecs, MX);
      MDB_LAUNCHED(ctx);
    }
  }
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// owner: operands X (rows x NB inside the outer-block buffer, leading dimension ldp) and the three state slices ->
// compact message
__global__ void __launch_bounds__(256) k_dist_pack(int64_t rows, const double* __restrict__ OX, int64_t ldp,
                                                   const double* __restrict__ alpha, const double* __restrict__ beta,
                                                   const double* __restrict__ rowmax, double* __restrict__ vecs,
                                                   double* __restrict__ MX) {
  const int c = threadIdx.x & (MDB_NB - 1);
  for (int64_t r = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); r < rows; r += (int64_t)gridDim.x * 2) {
    MX[r * MDB_NB + c] = OX[r * ldp + c];
    if (c == 0) { vecs[r] = alpha[r]; vecs[rows + r] = beta[r]; vecs[2 * rows + r] = rowmax[r]; }
  }
}

// header -> replicated d / omega / delta / clamped
__global__ void k_dist_absorb_hdr(int kb, int64_t k0, const double* __restrict__ hdr, double* __restrict__ d,
                                  double* __restrict__ omega, double* __restrict__ delta, uint8_t* __restrict__ clamped) {
  const int c = threadIdx.x;
  if (c < kb) {
    d[k0 + c] = hdr[c];
    omega[k0 + c] = hdr[MDB_NB + c];
    delta[k0 + c] = hdr[3 * MDB_NB + c];
    clamped[k0 + c] = (uint8_t)(hdr[4 * MDB_NB + c] != 0.0);
  }
}

// message X (rows x NB, compact) -> operand buffers: OX = X, OY = -X d (the same expression as k_panel, so the
// owner's and the receivers' operands are identical bit for bit).  HBM bound: 24 * rows * NB bytes.
// The same kernel takes the alpha / beta / rowmax slices and (its first CTA) the header: one launch per panel on the
// receiving side.
__global__ void __launch_bounds__(256) k_dist_absorb_x(int64_t rows, const double* __restrict__ X,
                                                       const double* __restrict__ hdr, double* __restrict__ OX,
                                                       double* __restrict__ OY, int64_t ldp,
                                                       const double* __restrict__ vecs, double* __restrict__ alpha,
                                                       double* __restrict__ beta, double* __restrict__ rowmax, int kb,
                                                       double* __restrict__ d, double* __restrict__ omega,
                                                       double* __restrict__ delta, uint8_t* __restrict__ clamped) {
  const int c = threadIdx.x & (MDB_NB - 1);
  const double dc = hdr[c];
  if (blockIdx.x == 0 && threadIdx.x < kb) {  // d / omega / delta / clamped of the panel's columns (pointers at column k0)
    d[threadIdx.x] = hdr[threadIdx.x];
    omega[threadIdx.x] = hdr[MDB_NB + threadIdx.x];
    delta[threadIdx.x] = hdr[3 * MDB_NB + threadIdx.x];
    clamped[threadIdx.x] = (uint8_t)(hdr[4 * MDB_NB + threadIdx.x] != 0.0);
  }
  for (int64_t r = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); r < rows; r += (int64_t)gridDim.x * 2) {
    const double x = X[r * MDB_NB + c];
    OX[r * ldp + c] = x;
    OY[r * ldp + c] = -x * dc;
    if (c == 0) { alpha[r] = vecs[r]; beta[r] = vecs[rows + r]; rowmax[r] = vecs[2 * rows + r]; }
  }
}

extern "C" int mdb_dist_absorb_panel(mdb_dist* s, int64_t k, int slot, int64_t k_outer, int oslot) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, k >= 0 && k < s->nblk_glob && (slot == 0 || slot == 1) && (oslot == 0 || oslot == 1),
            "mdb_dist_absorb_panel: bad panel / slot");
  MDB_CHECK(ctx, k_outer >= 0 && k_outer <= k && (k - k_outer + 1) * NB <= s->ldp, "mdb_dist_absorb_panel: bad outer block");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((k % s->world) == s->rank) return MDB_OK;  // the owner's arrays and operands are already in place
  const int64_t n = s->n, k0 = k * NB, k1 = std::min<int64_t>(k0 + NB, n), rows = n - k1, ob = k_outer * NB;
  double* buf = s->panel[slot];
  if (rows > 0) {
    double* vecs = buf + 5 * NB;
    const int64_t off = (k1 - ob) * s->ldp + (k0 - ob);
    const unsigned grid = (unsigned)std::min<int64_t>((rows + 1) / 2, (int64_t)ctx->num_sms * 8);
    k_dist_absorb_x<<<grid, 256, 0, ctx->stream>>>(rows, vecs + 3 * rows, buf, s->OPX[oslot] + off, s->OPY[oslot] + off, s->ldp,
                                                   vecs, s->alpha + k1, s->beta + k1, s->rowmax + k1, (int)(k1 - k0),
                                                   s->d + k0, s->omega + k0, s->delta + k0, s->clamped + k0);
    MDB_LAUNCHED(ctx);
  } else {
    k_dist_absorb_hdr<<<1, NB, 0, ctx->stream>>>((int)(k1 - k0), k0, buf, s->d, s->omega, s->delta, s->clamped);
    MDB_LAUNCHED(ctx);
  }
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_dist_apply_update(mdb_dist* s, int64_t k_outer, int oslot, int64_t k_first, int64_t k_count,
                                     int64_t blk_lo, int64_t blk_hi) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, oslot == 0 || oslot == 1, "mdb_dist_apply_update: bad slot");
  MDB_CHECK(ctx, k_outer >= 0 && k_first >= k_outer && k_count >= 1 && (k_first + k_count - k_outer) * NB <= s->ldp &&
                     k_first + k_count <= s->nblk_glob,
            "mdb_dist_apply_update: bad panel range");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (blk_hi < 0 || blk_hi > s->nblk_glob) blk_hi = s->nblk_glob;
  if (blk_lo < k_first + k_count) blk_lo = k_first + k_count;
  const int64_t n = s->n, ob = k_outer * NB, r0 = (k_first + k_count) * NB, rows = n - r0;
  if (rows <= 0 || blk_hi <= blk_lo) return MDB_OK;
  const int64_t tpb = NB / gemm::BN;  // N tiles per column block
  // local blocks lb with global index lb * world + rank in [blk_lo, blk_hi)
  auto first_local = [&](int64_t J) { return std::max<int64_t>((J - s->rank + s->world - 1) / s->world, 0); };
  const int64_t lb_begin = first_local(blk_lo), lb_end = std::min<int64_t>(first_local(blk_hi), s->nblk_loc);
  if (lb_end <= lb_begin) return MDB_OK;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const int64_t off = (r0 - ob) * s->ldp + (k_first - k_outer) * NB;
  g.A = s->OPX[oslot] + off; g.lda = s->ldp;
  g.B = s->OPY[oslot] + off; g.ldb = s->ldp;
  g.C = s->S + r0 * s->ld; g.ldc = s->ld;
  g.M = rows; g.N = s->nblk_loc * NB; g.K = (int)(k_count * NB);
  g.row_glob0 = r0; g.col_glob0 = 0; g.b_glob0 = r0; g.b_rows = rows;
  g.cyc_G = s->world; g.cyc_r = s->rank;  // one rank: plain mapping, global column = local column
  g.n_tile0 = lb_begin * tpb;
  g.n_tiles = (lb_end - lb_begin) * tpb;
  {  // this rank's algorithmic work: 2 K flop per strictly-lower element of its column blocks (bulk = the whole
     // outer block is contracted at once)
    double elems = 0;
    for (int64_t lb = lb_begin; lb < lb_end; ++lb) {
      const int64_t g0 = (lb * s->world + s->rank) * NB;         // first global column of the block
      const int64_t w = std::min<int64_t>(NB, n - g0);
      const double first = (double)(n - std::max<int64_t>(r0, g0 + 1));  // rows below the diagonal in column g0
      elems += w * first - 0.5 * (double)w * (double)(w - 1) * (g0 + 1 >= r0 ? 1.0 : 0.0);
    }
    const bool bulk = k_first == k_outer && k_count == mdb_dist_outer_width(s, k_outer);
    ctx->gemm_flop[bulk ? 0 : 1] += 2.0 * (double)g.K * elems;
    ctx->gemm_launches[bulk ? 0 : 1] += 1;
  }
  const int r = launch_gemm_tn_tma(ctx, g, rows, rows);
  if (r != -1) return r;
  return launch_gemm_tn<true, true, false>(ctx, g, 1);
}

static int dist_info_status(mdb_dist* s, int info) {
  if (info == 0) return MDB_OK;
  mdb_ctx* ctx = s->ctx;
  if (info > 0) {
    snprintf(ctx->err, sizeof(ctx->err), "distributed factorization: pivot %d is not positive", info - 1);
    return MDB_ERR_NOT_POSITIVE_DEFINITE;
  }
  snprintf(ctx->err, sizeof(ctx->err), "distributed factorization: device-side status %d; the factor is not valid", info);
  return MDB_ERR_STATE;
}

extern "C" int mdb_dist_finalize(mdb_dist* s, double* d, double* omega, double* delta, uint8_t* clamped) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = s->n;
  dim3 grid((unsigned)n, (unsigned)((s->ld + 1023) / 1024));
  k_finalize_cols<<<grid, 256, 0, ctx->stream>>>(n, s->S, s->ld, s->ld, s->rank, s->world, s->omega,
                                                 s->prm.min_abs_value_L);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  if (d) MDB_CUDA(ctx, cudaMemcpyAsync(d, s->d, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (omega) MDB_CUDA(ctx, cudaMemcpyAsync(omega, s->omega, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (delta) MDB_CUDA(ctx, cudaMemcpyAsync(delta, s->delta, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (clamped) MDB_CUDA(ctx, cudaMemcpyAsync(clamped, s->clamped, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  int info = 0;
  MDB_CUDA(ctx, cudaMemcpyAsync(&info, s->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return dist_info_status(s, info);
}

// The device-side status word of the current factorization: 0 = fine, > 0 = first non-positive pivot + 1
// (MDB_RULE_EXACT); negative values are reserved for device-side errors.  The factor of a run with info != 0 must
// not be used; mdb_dist_finalize returns the matching error status.  Every rank should exchange the value
// (distributed.py all-reduces it): a failing pivot is only seen by the rank that owns its panel.
extern "C" int mdb_dist_info(mdb_dist* s, int* info_out) {
  if (!s || !info_out) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_CUDA(ctx, cudaMemcpyAsync(info_out, s->info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}
