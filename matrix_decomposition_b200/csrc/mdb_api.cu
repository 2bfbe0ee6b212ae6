// mdb_api.cu -- the C ABI of libmdb200.so (see include/mdb200.h).  Host orchestration of the blocked
// right-looking modified LDL^T, the permutation gather, and solve / multiply with a resident factor.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <math.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <new>

#include "kernels_factor.cuh"
#include "kernels_gemm.cuh"
#include "kernels_gemm_tma.cuh"
#include "kernels_solve.cuh"
#include "kernels_dynamic.cuh"
#include "kernels_gmw.cuh"

#define NB MDB_NB
#define PANEL_R MDB_PANEL_RT

enum { ST_EMPTY = 0, ST_MATRIX = 1, ST_FACTORED = 2, ST_FINAL = 3 };

struct mdb_factor {
  mdb_ctx* ctx;
  int64_t n, ld, batch;
  double* W;         // batch x n x ld
  size_t W_cap;      // bytes behind W (for the context's buffer cache)
  double* vec;       // 7 x batch x n: gamma, alpha, beta, rowmax, d, omega, delta
  double *gamma, *alpha, *beta, *rowmax, *d, *omega, *delta;
  uint8_t* clamped;  // batch x n
  double *minB, *maxB;
  int64_t* p_dev;    // batch x n (null = identity)
  int* info;         // batch
  double *Ut, *hdr;  // batch x NB*NB, batch x 5*NB
  double* Vinv;      // batch x MDB_VINV_DOUBLES
  double *PX, *PY;   // batch x panel_rows x (agg * NB): operands of one outer block
  double *PX2, *PY2; // second pair for lookahead
  int64_t panel_rows;
  int agg;           // panels per outer block (1 for batched small matrices)
  int state, out_type, rule;
  double mavL;
  bool prepared;
  double *Linv, *LinvT, *Ltri, *Utri;
  // right-hand-side workspaces of solve / multiply (grow-only, so that a solve does not pay cudaMalloc / cudaFree)
  double* ws[3]; size_t ws_cap[3];
  // streamed output (one-shot calls with a host L): rows of L are finalised and copied out per outer block
  double* out_host; int64_t out_ld; int out_stream_type;
  double* out_host_dev;   // device-side address of out_host when the pinned memory is mapped (null: not addressable)
  // persistent triangular sweep (<= 8 right-hand sides): task list, flags, partial sums
  int2* trsv_tasks; int trsv_ntasks; unsigned int* trsv_sync; double* trsv_partial; int trsv_partial_rhs;
  std::vector<int64_t> p_host;
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
// Wait for the background release of an earlier one-shot call (before anything that allocates device memory).
static void ctx_reap(mdb_ctx* ctx) {
  if (ctx && ctx->reaper) {
    ctx->reaper->join();
    delete ctx->reaper;
    ctx->reaper = nullptr;
  }
}

// ---- device memory ----
// The two largest buffers of a factorization (the working matrix and the upload staging area, 34 GB each at
// n = 65536) are kept in the context when they are released and handed out again to the next call that needs the
// same amount: cudaMalloc + cudaFree of that size cost ~0.1 s per call (and still 10-20 ms at 74 MB, so everything
// from 1 MB on goes through the cache).  mdb_ctx_trim() gives them back to the driver;
// any allocation of the library that fails trims the cache and tries once more.  MDB200_CACHE=0 disables the cache.
static void factor_release(mdb_factor* f);
static void ctx_trim(mdb_ctx* ctx) {
  if (ctx->parked) { factor_release(ctx->parked); ctx->parked = nullptr; }
  if (ctx->tail_flags) { cudaFree(ctx->tail_flags); ctx->tail_flags = nullptr; ctx->tail_flags_cap = 0; }
  for (int i = 0; i < 3; ++i) {
    if (ctx->big_ptr[i]) cudaFree(ctx->big_ptr[i]);
    ctx->big_ptr[i] = nullptr;
    ctx->big_bytes[i] = 0;
  }
}

static cudaError_t dev_malloc(mdb_ctx* ctx, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation && (ctx->big_ptr[0] || ctx->big_ptr[1] || ctx->big_ptr[2])) {
    cudaGetLastError();
    ctx_trim(ctx);
    e = cudaMalloc(p, bytes);
  }
  return e;
}

// *cap receives the size to pass to big_free (the cached buffer may be a little larger than asked for)
static cudaError_t big_malloc(mdb_ctx* ctx, void** p, size_t bytes, size_t* cap) {
  for (int i = 0; i < 3; ++i)
    if (ctx->big_ptr[i] && ctx->big_bytes[i] >= bytes && ctx->big_bytes[i] - bytes <= bytes / 8) {
      *p = ctx->big_ptr[i];
      *cap = ctx->big_bytes[i];
      ctx->big_ptr[i] = nullptr;
      ctx->big_bytes[i] = 0;
      return cudaSuccess;
    }
  *cap = bytes;
  return dev_malloc(ctx, p, bytes);
}

// Caller's thread only (never the reaper), and only after the work that used the buffer has completed.
static void big_free(mdb_ctx* ctx, void* p, size_t cap) {
  if (!p) return;
  if (ctx->use_cache && cap >= ((size_t)1 << 20)) {   // (from n ~ 360: cudaMalloc + cudaFree cost 2 ms per call at 8 MB, 10-20 ms at 74 MB)
    int slot = 0;   // a free slot, else the smallest buffer
    for (int i = 0; i < 3; ++i) {
      if (!ctx->big_ptr[i]) { slot = i; break; }
      if (ctx->big_bytes[i] < ctx->big_bytes[slot]) slot = i;
    }
    if (ctx->big_ptr[slot] && ctx->big_bytes[slot] > cap) {  // every slot holds a larger buffer: keep those
      cudaFree(p);
      return;
    }
    if (ctx->big_ptr[slot]) cudaFree(ctx->big_ptr[slot]);
    ctx->big_ptr[slot] = p;
    ctx->big_bytes[slot] = cap;
    return;
  }
  cudaFree(p);
}

extern "C" int mdb_ctx_trim(mdb_ctx* ctx) {
  if (!ctx) return MDB_ERR_INVALID;
  ctx_reap(ctx);
  cudaSetDevice(ctx->device);
  ctx_trim(ctx);
  return MDB_OK;
}

extern "C" int mdb_version(void) { return 100; }

extern "C" int mdb_ctx_create(int device, mdb_ctx** out) {
  if (!out) return MDB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return MDB_ERR_NO_DEVICE;
  mdb_ctx* ctx = new (std::nothrow) mdb_ctx();
  if (!ctx) return MDB_ERR_INVALID;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  *out = ctx;  // also on failure: the caller reads mdb_last_error() and then destroys the context
  MDB_CUDA(ctx, cudaSetDevice(device));
  {
    int lo_pri = 0, hi_pri = 0;
    MDB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    MDB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, hi_pri));
    MDB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, lo_pri));
    MDB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi_pri));
    MDB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      MDB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_panel[i], cudaEventDisableTiming));
      MDB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_upd[i], cudaEventDisableTiming));
    }
    const char* la = getenv("MDB200_LOOKAHEAD");
    ctx->lookahead = !(la && la[0] == '0');
    const char* tm = getenv("MDB200_TMA");
    ctx->use_tma = !(tm && tm[0] == '0');
    const char* tv = getenv("MDB200_TRSV");
    ctx->use_trsv = !(tv && tv[0] == '0');
    const char* cch = getenv("MDB200_CACHE");
    ctx->use_cache = !(cch && cch[0] == '0');
    const char* cgv = getenv("MDB200_CLUSTER_GATHER");
    ctx->use_cluster_gather = !(cgv && cgv[0] == '0');
    // outer block width of the factorization: MDB200_AGG = 1, 2, 4, 8 panels (fixed) or unset = by remaining size,
    // thresholds MDB200_AGG_THRESH = "t2,t4,t8" (remaining rows from which 2, 4, 8 panels are aggregated)
    const char* ag = getenv("MDB200_AGG");
    ctx->agg_fixed = ag ? atoi(ag) : 0;
    if (ctx->agg_fixed != 1 && ctx->agg_fixed != 2 && ctx->agg_fixed != 4 && ctx->agg_fixed != 8) ctx->agg_fixed = 0;
    ctx->agg_thresh[0] = 4096; ctx->agg_thresh[1] = 8192; ctx->agg_thresh[2] = 32768;
    const char* at = getenv("MDB200_AGG_THRESH");
    if (at) {
      long long a = 0, b = 0, c = 0;
      if (sscanf(at, "%lld,%lld,%lld", &a, &b, &c) == 3 && a > 0 && b >= a && c >= b) {
        ctx->agg_thresh[0] = a; ctx->agg_thresh[1] = b; ctx->agg_thresh[2] = c;
      }
    }
  }
  ctx->stream = ctx->own_stream;
  MDB_CUDA(ctx, cudaEventCreate(&ctx->ev[0]));
  MDB_CUDA(ctx, cudaEventCreate(&ctx->ev[1]));
  cudaDeviceProp prop;
  MDB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    snprintf(ctx->err, sizeof(ctx->err), "device %d is sm_%d%d; libmdb200 is built for sm_100a only", device,
             prop.major, prop.minor);
    return MDB_ERR_NO_DEVICE;
  }
  return MDB_OK;
}

extern "C" int mdb_ctx_destroy(mdb_ctx* ctx) {
  if (!ctx) return MDB_OK;
  ctx_reap(ctx);
  cudaSetDevice(ctx->device);
  ctx_trim(ctx);
  for (int i = 0; i < 3; ++i) {
    if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_panel[i]) cudaEventDestroy(ctx->ev_panel[i]);
    if (ctx->ev_upd[i]) cudaEventDestroy(ctx->ev_upd[i]);
  }
  if (ctx->ev[0]) cudaEventDestroy(ctx->ev[0]);
  if (ctx->ev[1]) cudaEventDestroy(ctx->ev[1]);
  delete ctx;
  return MDB_OK;
}

extern "C" int mdb_ctx_set_stream(mdb_ctx* ctx, void* s) {
  if (!ctx) return MDB_ERR_INVALID;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return MDB_OK;
}

extern "C" int mdb_ctx_synchronize(mdb_ctx* ctx) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

extern "C" const char* mdb_last_error(const mdb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t mdb_launch_count(const mdb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int mdb_ctx_set_profiling(mdb_ctx* ctx, int on) {
  if (!ctx) return MDB_ERR_INVALID;
  ctx->profiling = on;
  return MDB_OK;
}
extern "C" int mdb_ctx_last_input_nonfinite(const mdb_ctx* ctx) { return ctx ? ctx->last_input_nonfinite : 0; }
extern "C" int mdb_ctx_get_update_stats(const mdb_ctx* ctx, double out[4]) {
  if (!ctx || !out) return MDB_ERR_INVALID;
  out[0] = ctx->gemm_flop[0]; out[1] = (double)ctx->gemm_launches[0];
  out[2] = ctx->gemm_flop[1]; out[3] = (double)ctx->gemm_launches[1];
  return MDB_OK;
}
extern "C" int mdb_ctx_set_blocking(mdb_ctx* ctx, int fixed, int64_t t2, int64_t t4, int64_t t8) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, fixed == 0 || fixed == 1 || fixed == 2 || fixed == 4 || fixed == 8, "mdb_ctx_set_blocking: fixed must be 0, 1, 2, 4 or 8");
  MDB_CHECK(ctx, fixed != 0 || (t2 > 0 && t4 >= t2 && t8 >= t4), "mdb_ctx_set_blocking: need 0 < t2 <= t4 <= t8");
  ctx->agg_fixed = fixed;
  if (fixed == 0) { ctx->agg_thresh[0] = t2; ctx->agg_thresh[1] = t4; ctx->agg_thresh[2] = t8; }
  return MDB_OK;
}
extern "C" int mdb_ctx_get_phase_ms(const mdb_ctx* ctx, double ms[4]) {
  if (!ctx || !ms) return MDB_ERR_INVALID;
  for (int i = 0; i < 4; ++i) ms[i] = ctx->phase_ms[i];
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
extern "C" int mdb_malloc(mdb_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return MDB_ERR_INVALID;
  ctx_reap(ctx);
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_CUDA(ctx, dev_malloc(ctx, dptr, bytes ? bytes : 8));
  return MDB_OK;
}
extern "C" int mdb_free(mdb_ctx* ctx, void* dptr) {
  if (!ctx) return MDB_ERR_INVALID;
  if (dptr) MDB_CUDA(ctx, cudaFree(dptr));
  return MDB_OK;
}
extern "C" int mdb_malloc_host(mdb_ctx* ctx, size_t bytes, void** hptr) {
  if (!ctx || !hptr) return MDB_ERR_INVALID;
  MDB_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 8));
  return MDB_OK;
}
extern "C" int mdb_free_host(mdb_ctx* ctx, void* hptr) {
  if (!ctx) return MDB_ERR_INVALID;
  if (hptr) MDB_CUDA(ctx, cudaFreeHost(hptr));
  return MDB_OK;
}
static cudaMemcpyKind copy_kind(int dst_loc, int src_loc) {
  if (dst_loc == MDB_DEVICE && src_loc == MDB_HOST) return cudaMemcpyHostToDevice;
  if (dst_loc == MDB_HOST && src_loc == MDB_DEVICE) return cudaMemcpyDeviceToHost;
  if (dst_loc == MDB_DEVICE && src_loc == MDB_DEVICE) return cudaMemcpyDeviceToDevice;
  return cudaMemcpyHostToHost;
}
extern "C" int mdb_memcpy(mdb_ctx* ctx, void* dst, const void* src, size_t bytes, int dst_loc, int src_loc) {
  if (!ctx) return MDB_ERR_INVALID;
  if (!bytes) return MDB_OK;
  MDB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, copy_kind(dst_loc, src_loc), ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}
// Large copies between PAGEABLE host memory and the device: the driver stages those through a single thread
// (~10 GB/s).  Here the host side of the staging is done by several threads into / out of three pinned buffers
// while the DMA of the neighbouring chunks is in flight.  Returns -1 when the plain path should be used.
static bool host_is_pageable(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return attr.type == cudaMemoryTypeUnregistered;
}

// One row, source and destination anywhere: unless MDB200_COPY_NT=0 the destination is written with streaming stores
// (no read-for-ownership of the destination lines; the staging buffers are read next by the DMA engine, not the CPU).
static inline void row_copy(char* dst, const char* src, size_t bytes, bool nt) {
#if defined(__x86_64__)
  if (nt && bytes >= 256) {
    size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head) { memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
    const size_t body = bytes & ~(size_t)63;
    for (size_t i = 0; i < body; i += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
      _mm_stream_si128((__m128i*)(dst + i), a); _mm_stream_si128((__m128i*)(dst + i + 16), b);
      _mm_stream_si128((__m128i*)(dst + i + 32), c); _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    if (bytes > body) memcpy(dst + body, src + body, bytes - body);
    return;
  }
#endif
  memcpy(dst, src, bytes);
}

static void parallel_rows_copy(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes,
                               int64_t rows) {
  // measured on the 16-vCPU B200 host, 8.6 GB pageable -> device (tools/staging_bench.py, profiles/r02_staging_bench.json):
  // 8 threads + memcpy 37 GB/s, 12 threads + streaming stores 50.5 GB/s (the pinned-memory rate is 52 - 55 GB/s)
  const char* te = getenv("MDB200_COPY_THREADS");
  const int64_t hw = std::max<int64_t>(1, (int64_t)std::thread::hardware_concurrency());
  const int64_t cap = te ? std::max<int64_t>(1, atoll(te)) : std::min<int64_t>(12, std::max<int64_t>(1, hw * 3 / 4));
  const char* ne = getenv("MDB200_COPY_NT");
  const bool nt = !(ne && ne[0] == '0');
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(cap, hw), rows));
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t)
    th.emplace_back([=]() {
      for (int64_t r = rows * t / T; r < rows * (t + 1) / T; ++r) row_copy(dst + r * dst_pitch, src + r * src_pitch, row_bytes, nt);
#if defined(__x86_64__)
      if (nt) _mm_sfence();
#endif
    });
  for (int64_t r = 0; r < rows / T; ++r) row_copy(dst + r * dst_pitch, src + r * src_pitch, row_bytes, nt);
#if defined(__x86_64__)
  if (nt) _mm_sfence();
#endif
  for (auto& x : th) x.join();
}

static int staged_copy2d(mdb_ctx* ctx, void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows,
                         int64_t cols, int dst_loc, int src_loc) {
  static const bool enabled = !(getenv("MDB200_STAGED") && getenv("MDB200_STAGED")[0] == '0');
  if (!enabled) return -1;
  const bool h2d = (src_loc == MDB_HOST && dst_loc == MDB_DEVICE), d2h = (src_loc == MDB_DEVICE && dst_loc == MDB_HOST);
  const size_t row_bytes = (size_t)cols * 8;
  const size_t total = (size_t)rows * row_bytes;
  if (!(h2d || d2h) || total < ((size_t)4 << 20) || row_bytes > ((size_t)8 << 20)) return -1;
  if (!host_is_pageable(h2d ? src : dst)) return -1;
  const size_t SB = (size_t)64 << 20;
  if (!ctx->stage[0]) {
    for (int i = 0; i < 3; ++i) {
      if (cudaMallocHost(&ctx->stage[i], SB) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        for (int k = 0; k <= i; ++k) { if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]); ctx->stage[k] = nullptr; }
        return -1;
      }
    }
    ctx->stage_bytes = SB;
  }
  // chunks of 1 ... 64 MB: at least ~4-8 of them, so that staging and DMA overlap also for a few tens of MB
  const size_t chunk_bytes = std::min<size_t>(SB, std::max<size_t>((size_t)1 << 20, total / 8));
  const int64_t R = std::max<int64_t>(1, (int64_t)(chunk_bytes / row_bytes));   // rows per chunk
  const int64_t nchunks = (rows + R - 1) / R;
  if (h2d) {
    // The buffers rotate across calls and nothing is drained at the end: once a chunk sits in a staging buffer the
    // caller's memory is no longer needed, and the copy is ordered on ctx->stream like any other asynchronous copy
    // (mdb_factor_set_matrix uploads 32 row blocks at n = 65536; a drain per block cost ~0.1 s of bubbles).
    for (int64_t c = 0; c < nchunks; ++c) {
      const int b = ctx->stage_next;
      ctx->stage_next = (b + 1) % 3;
      const int64_t r0 = c * R, nr = std::min<int64_t>(R, rows - r0);
      if (ctx->stage_busy[b]) MDB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[b]));   // the DMA that read this buffer is done
      parallel_rows_copy((char*)ctx->stage[b], row_bytes, (const char*)src + (size_t)r0 * ld_src * 8, (size_t)ld_src * 8,
                         row_bytes, nr);
      MDB_CUDA(ctx, cudaMemcpy2DAsync((char*)dst + (size_t)r0 * ld_dst * 8, (size_t)ld_dst * 8, ctx->stage[b], row_bytes,
                                      row_bytes, (size_t)nr, cudaMemcpyHostToDevice, ctx->stream));
      MDB_CUDA(ctx, cudaEventRecord(ctx->stage_ev[b], ctx->stream));
      ctx->stage_busy[b] = true;
    }
  } else {
    for (int b = 0; b < 3; ++b)
      if (ctx->stage_busy[b]) { MDB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[b])); ctx->stage_busy[b] = false; }
    for (int64_t c = 0; c < nchunks + 2; ++c) {
      if (c < nchunks) {
        const int b = (int)(c % 3);
        const int64_t r0 = c * R, nr = std::min<int64_t>(R, rows - r0);
        MDB_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage[b], row_bytes, (const char*)src + (size_t)r0 * ld_src * 8,
                                        (size_t)ld_src * 8, row_bytes, (size_t)nr, cudaMemcpyDeviceToHost, ctx->stream));
        MDB_CUDA(ctx, cudaEventRecord(ctx->stage_ev[b], ctx->stream));
      }
      if (c >= 2) {
        const int64_t cc = c - 2;
        const int b = (int)(cc % 3);
        const int64_t r0 = cc * R, nr = std::min<int64_t>(R, rows - r0);
        MDB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[b]));
        parallel_rows_copy((char*)dst + (size_t)r0 * ld_dst * 8, (size_t)ld_dst * 8, (const char*)ctx->stage[b], row_bytes,
                           row_bytes, nr);
      }
    }
  }
  return MDB_OK;
}

static int copy2d_async(mdb_ctx* ctx, void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows,
                        int64_t cols, int dst_loc, int src_loc) {
  if (rows <= 0 || cols <= 0) return MDB_OK;
  {
    const int r = staged_copy2d(ctx, dst, ld_dst, src, ld_src, rows, cols, dst_loc, src_loc);
    if (r != -1) return r;
  }
  MDB_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)ld_dst * 8, src, (size_t)ld_src * 8, (size_t)cols * 8, (size_t)rows,
                                  copy_kind(dst_loc, src_loc), ctx->stream));
  return MDB_OK;
}
extern "C" int mdb_memcpy2d(mdb_ctx* ctx, void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t rows,
                            int64_t cols, int dst_loc, int src_loc) {
  if (!ctx) return MDB_ERR_INVALID;
  int rc = copy2d_async(ctx, dst, ld_dst, src, ld_src, rows, cols, dst_loc, src_loc);
  if (rc) return rc;
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// factor object
// ------------------------------------------------------------------------------------------------
// W goes back to the context's cache: call on the caller's thread, after the factor's work has completed, before
// factor_release (which may then run on the reaper thread).
static void factor_detach_W(mdb_factor* f) {
  big_free(f->ctx, f->W, f->W_cap);
  f->W = nullptr;
  f->W_cap = 0;
}

// Keep the workspace of a finished one-shot factor (W already detached) for the next call of the same shape: what is
// allocated on demand goes, the fixed-size buffers stay, the state is that of a fresh factor.
static void factor_park(mdb_factor* f) {
  mdb_ctx* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  void* lazy[] = {f->minB, f->maxB, f->p_dev, f->Linv, f->LinvT, f->Ltri, f->Utri, f->trsv_tasks, f->trsv_sync,
                  f->trsv_partial, f->ws[0], f->ws[1], f->ws[2]};
  for (void* q : lazy) cudaFree(q);
  f->minB = f->maxB = nullptr; f->p_dev = nullptr;
  f->Linv = f->LinvT = f->Ltri = f->Utri = nullptr;
  f->trsv_tasks = nullptr; f->trsv_ntasks = 0; f->trsv_sync = nullptr; f->trsv_partial = nullptr; f->trsv_partial_rhs = 0;
  for (int i = 0; i < 3; ++i) { f->ws[i] = nullptr; f->ws_cap[i] = 0; }
  f->state = ST_EMPTY; f->out_type = MDB_TYPE_LDL; f->rule = MDB_RULE_REIMER; f->mavL = 0; f->prepared = false;
  f->out_host = nullptr; f->out_host_dev = nullptr; f->out_ld = 0; f->out_stream_type = MDB_TYPE_LDL;
  f->p_host.clear();
  if (ctx->parked) factor_release(ctx->parked);
  ctx->parked = f;
}

static void factor_release(mdb_factor* f) {
  if (!f) return;
  cudaSetDevice(f->ctx->device);
  cudaFree(f->W); cudaFree(f->vec); cudaFree(f->clamped); cudaFree(f->minB); cudaFree(f->maxB);
  cudaFree(f->p_dev); cudaFree(f->info); cudaFree(f->Ut); cudaFree(f->hdr); cudaFree(f->Vinv); cudaFree(f->PX); cudaFree(f->PY); cudaFree(f->PX2); cudaFree(f->PY2);
  cudaFree(f->Linv); cudaFree(f->LinvT); cudaFree(f->Ltri); cudaFree(f->Utri);
  cudaFree(f->trsv_tasks); cudaFree(f->trsv_sync); cudaFree(f->trsv_partial);
  for (int i = 0; i < 3; ++i) cudaFree(f->ws[i]);
  delete f;
}

extern "C" int mdb_factor_create(mdb_ctx* ctx, int64_t n, int64_t batch, mdb_factor** out) {
  if (!ctx || !out) return MDB_ERR_INVALID;
  *out = nullptr;
  MDB_CHECK(ctx, n >= 0 && batch >= 1, "mdb_factor_create: n >= 0 and batch >= 1 required");
  ctx_reap(ctx);
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int agg_new = batch != 1 ? 1 : ctx->agg_fixed ? ctx->agg_fixed
                      : std::max<int64_t>(n, 1) >= ctx->agg_thresh[2] ? 8 : std::max<int64_t>(n, 1) >= ctx->agg_thresh[1] ? 4
                      : std::max<int64_t>(n, 1) >= ctx->agg_thresh[0] ? 2 : 1;
  if (ctx->parked) {
    // the previous one-shot call left its workspace behind: same shape -> take it over (saves ~10 cudaMalloc /
    // cudaFree pairs, 2 GB at n = 65536, per call); another shape -> give it back first
    mdb_factor* f = ctx->parked;
    ctx->parked = nullptr;
    if (f->n == n && f->batch == batch && f->agg == agg_new) {
      const int64_t nn = std::max<int64_t>(n, 1), bn = batch * nn;
      const size_t panel_bytes = (size_t)batch * f->panel_rows * f->agg * NB * 8;
      cudaError_t e = big_malloc(ctx, (void**)&f->W, (size_t)batch * nn * f->ld * 8, &f->W_cap);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->W, 0, (size_t)batch * nn * f->ld * 8, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->vec, 0, (size_t)7 * bn * 8, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->clamped, 0, (size_t)bn, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->info, 0, (size_t)batch * sizeof(int), ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->Ut, 0, (size_t)batch * NB * NB * 8, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->PX, 0, panel_bytes, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->PY, 0, panel_bytes, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->PX2, 0, panel_bytes, ctx->stream);
      if (e == cudaSuccess) e = cudaMemsetAsync(f->PY2, 0, panel_bytes, ctx->stream);
      if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "mdb_factor_create: %s", cudaGetErrorString(e));
        factor_release(f);
        return MDB_ERR_CUDA;
      }
      *out = f;
      return MDB_OK;
    }
    factor_release(f);
  }
  mdb_factor* f = new (std::nothrow) mdb_factor();
  if (!f) return mdb_fail(ctx, MDB_ERR_INVALID, "%s", "out of host memory");
  f->ctx = ctx; f->n = n; f->batch = batch;
  f->ld = std::max<int64_t>(mdb_round_up(n, NB), NB);
  f->W = nullptr; f->vec = nullptr; f->clamped = nullptr; f->minB = f->maxB = nullptr; f->p_dev = nullptr;
  f->info = nullptr; f->Vinv = nullptr; f->Ut = f->hdr = f->PX = f->PY = f->PX2 = f->PY2 = nullptr; f->Linv = f->LinvT = f->Ltri = f->Utri = nullptr;
  f->state = ST_EMPTY; f->out_type = MDB_TYPE_LDL; f->rule = MDB_RULE_REIMER; f->mavL = 0; f->prepared = false;
  f->out_host = nullptr; f->out_host_dev = nullptr; f->out_ld = 0; f->out_stream_type = MDB_TYPE_LDL;
  for (int i = 0; i < 3; ++i) { f->ws[i] = nullptr; f->ws_cap[i] = 0; }
  f->trsv_tasks = nullptr; f->trsv_ntasks = 0; f->trsv_sync = nullptr; f->trsv_partial = nullptr; f->trsv_partial_rhs = 0;
  const int64_t nn = std::max<int64_t>(n, 1);
  f->panel_rows = std::max<int64_t>(mdb_round_up(nn, NB), NB);
  // widest outer block this factor can use (operand buffer width): 1 for batched small matrices
  f->agg = 1;
  if (batch == 1) {
    if (ctx->agg_fixed) f->agg = ctx->agg_fixed;
    else f->agg = nn >= ctx->agg_thresh[2] ? 8 : nn >= ctx->agg_thresh[1] ? 4 : nn >= ctx->agg_thresh[0] ? 2 : 1;
  }
  const size_t panel_bytes = (size_t)batch * f->panel_rows * f->agg * NB * 8;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = dev_malloc(ctx, p, bytes ? bytes : 8); };
  f->W_cap = 0;
  e = big_malloc(ctx, (void**)&f->W, (size_t)batch * nn * f->ld * 8, &f->W_cap);
  A((void**)&f->vec, (size_t)7 * batch * nn * 8);
  A((void**)&f->clamped, (size_t)batch * nn);
  A((void**)&f->info, (size_t)(batch + 1) * sizeof(int));
  A((void**)&f->Ut, (size_t)batch * NB * NB * 8);
  A((void**)&f->hdr, (size_t)batch * 5 * NB * 8);
  A((void**)&f->Vinv, (size_t)batch * MDB_VINV_DOUBLES * 8);
  A((void**)&f->PX, panel_bytes);
  A((void**)&f->PY, panel_bytes);
  A((void**)&f->PX2, panel_bytes);
  A((void**)&f->PY2, panel_bytes);
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "mdb_factor_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    factor_release(f);
    return MDB_ERR_CUDA;
  }
  const int64_t bn = batch * nn;
  f->gamma = f->vec; f->alpha = f->vec + bn; f->beta = f->vec + 2 * bn; f->rowmax = f->vec + 3 * bn;
  f->d = f->vec + 4 * bn; f->omega = f->vec + 5 * bn; f->delta = f->vec + 6 * bn;
  // zero everything once: the padding columns of W (ld > n) must stay zero for the solve kernels
  if (e == cudaSuccess) e = cudaMemsetAsync(f->W, 0, (size_t)batch * nn * f->ld * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->vec, 0, (size_t)7 * bn * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->clamped, 0, (size_t)bn, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->info, 0, (size_t)batch * sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->Ut, 0, (size_t)batch * NB * NB * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->PX, 0, panel_bytes, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->PY, 0, panel_bytes, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->PX2, 0, panel_bytes, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->PY2, 0, panel_bytes, ctx->stream);
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "mdb_factor_create: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    factor_release(f);
    return MDB_ERR_CUDA;
  }
  *out = f;
  return MDB_OK;
}

extern "C" int mdb_factor_destroy(mdb_factor* f) {
  if (f) {
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    factor_detach_W(f);
    if (f->ctx->use_cache && f->n >= 256) factor_park(f);   // the next factor of this shape takes the workspace over
    else factor_release(f);
  }
  return MDB_OK;
}

static int factor_set_p(mdb_factor* f, const int64_t* p) {
  mdb_ctx* ctx = f->ctx;
  const int64_t bn = f->batch * f->n;
  if (!p) {
    f->p_host.clear();
    if (f->p_dev) { cudaFree(f->p_dev); f->p_dev = nullptr; }
    return MDB_OK;
  }
  for (int64_t b = 0; b < f->batch; ++b) {  // validate: a permutation of 0..n-1
    std::vector<uint8_t> seen((size_t)f->n, 0);
    for (int64_t i = 0; i < f->n; ++i) {
      const int64_t v = p[b * f->n + i];
      if (v < 0 || v >= f->n || seen[(size_t)v]) return mdb_fail(ctx, MDB_ERR_INVALID, "%s", "p is not a permutation");
      seen[(size_t)v] = 1;
    }
  }
  f->p_host.assign(p, p + bn);
  if (!f->p_dev) MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->p_dev, (size_t)std::max<int64_t>(bn, 1) * 8));
  MDB_CUDA(ctx, cudaMemcpyAsync(f->p_dev, p, (size_t)bn * 8, cudaMemcpyHostToDevice, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

static int gather_sym(mdb_ctx* ctx, int64_t n, int64_t batch, const double* A_dev, int64_t lda, int64_t sA,
                      const int64_t* p_dev, double* W, int64_t ldw, int64_t sW, double* gamma, int* nonfinite = nullptr,
                      int upper_only = 0, int known_symmetric = 0) {
  // known_symmetric: A[i, j] == A[j, i] bit for bit (the mirrored upload of mdb_factor_set_matrix); the second pass of
  // the large-matrix path then reads only half of its input.  permute.symmetric (dense/permute.py:26) is defined for any
  // square matrix and never sets it.
  if (n == 0) return MDB_OK;
  // large single matrices: two passes of "gather rows, transpose" through a temporary n x n buffer (k_gather_rows_transpose)
  if (!upper_only && ctx->use_cluster_gather && batch == 1 && n > 8192) {
    double* B = nullptr;
    size_t cap = 0;
    if (big_malloc(ctx, (void**)&B, (size_t)n * n * 8, &cap) == cudaSuccess) {
      const unsigned t = (unsigned)((n + MDB_TT - 1) / MDB_TT);
      k_gather_rows_transpose<<<dim3(t, t), MDB_TT_THREADS, 0, ctx->stream>>>(n, A_dev, lda, p_dev, B, n, 0, nullptr, nonfinite);
      k_gather_rows_transpose<<<dim3(t, t), MDB_TT_THREADS, 0, ctx->stream>>>(n, B, n, p_dev, W, ldw, known_symmetric, gamma, nullptr);
      MDB_LAUNCHED(ctx);
      MDB_LAUNCHED(ctx);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // B goes back to the cache only when nothing reads it
      big_free(ctx, B, cap);
      MDB_CUDA(ctx, e);
      return MDB_OK;
    }
    cudaGetLastError();   // no room for the temporary: one pass with the source row in distributed shared memory
  }
  // cluster kernel: source row staged in (distributed) shared memory, row read once
  const int shift = (n <= 8 * 8192) ? 13 : 14;  // chunk of 8192 doubles (64 KB) or 16384 (128 KB)
  const int64_t chunk = (int64_t)1 << shift;
  int C = 1;
  while (C < 8 && (int64_t)C * chunk < n) C *= 2;
  if (!upper_only && ctx->use_cluster_gather && (int64_t)C * chunk >= n && n * C <= 0x7fffffff) {
    static uint64_t configured = 0;
    if (!mdb_device_marked(configured, ctx)) {
      MDB_CUDA(ctx, cudaFuncSetAttribute(k_gather_sym_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
      mdb_mark_device(configured, ctx);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(n * C), 1, (unsigned)batch);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = (size_t)std::min<int64_t>(chunk, n) * 8;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int64_t sP = n, sVec = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_gather_sym_cluster, n, A_dev, lda, sA, p_dev, sP, W, ldw, sW, gamma, sVec,
                                       nonfinite, shift);
    if (e == cudaSuccess) {
      MDB_LAUNCHED(ctx);
      return MDB_OK;
    }
    cudaGetLastError();  // fall through to the plain gather
  }
  dim3 grid((unsigned)n, (unsigned)((n + 1023) / 1024), (unsigned)batch);
  k_gather_sym<<<grid, 256, 0, ctx->stream>>>(n, A_dev, lda, sA, p_dev, n, W, ldw, sW, gamma, n, nonfinite, upper_only);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_factor_set_matrix(mdb_factor* f, const double* A, int64_t lda, const int64_t* p, int a_loc) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  ctx->last_input_nonfinite = 0;   // describes THIS call from here on, also when it returns early
  MDB_CHECK(ctx, A || f->n == 0, "mdb_factor_set_matrix: A is null");
  MDB_CHECK(ctx, lda >= f->n, "mdb_factor_set_matrix: lda < n");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = factor_set_p(f, p);
  if (rc) return rc;
  const int64_t n = f->n;
  if (n == 0) { f->state = ST_MATRIX; return MDB_OK; }
  const double* A_dev = A;
  double* staging = nullptr;
  size_t staging_cap = 0;
  int64_t lds = lda;
  int upper_only = 0, mirrored = 0;
  if (a_loc == MDB_HOST) {
    MDB_CUDA(ctx, big_malloc(ctx, (void**)&staging, (size_t)f->batch * n * n * 8, &staging_cap));
    if (f->batch == 1 && n >= 4096) {
      // A is symmetric: only its upper triangle crosses PCIe (row blocks [r0, r0 + RB) x columns [r0, n)),
      // the gather mirrors it.  Half the bytes of the full copy.
      const int64_t RB = 2048;
      for (int64_t r0 = 0; r0 < n && !rc; r0 += RB)
        rc = copy2d_async(ctx, staging + r0 * n + r0, n, A + r0 * lda + r0, lda, std::min<int64_t>(RB, n - r0), n - r0,
                          MDB_DEVICE, MDB_HOST);
      if (!rc) {
        const unsigned t32 = (unsigned)((n + 31) / 32);
        k_mirror_upper_to_lower<<<dim3(t32, t32), 256, 0, ctx->stream>>>(n, staging, n);
        MDB_LAUNCHED(ctx);
        mirrored = 1;
      }
    } else {
      rc = copy2d_async(ctx, staging, n, A, lda, f->batch * n, n, MDB_DEVICE, MDB_HOST);  // batch stride = n*lda
      upper_only = 1;   // a host matrix is read through its upper triangle at EVERY size (the large ones upload nothing else)
    }
    if (rc) { cudaStreamSynchronize(ctx->stream); big_free(ctx, staging, staging_cap); return rc; }
    A_dev = staging;
    lds = n;
  }
  // the finite check of the input rides on the gather: flag = the int after the per-matrix info entries
  int* flag = f->info + f->batch;
  MDB_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
  rc = gather_sym(ctx, n, f->batch, A_dev, lds, n * lds, f->p_dev, f->W, f->ld, n * f->ld, f->gamma, flag, upper_only, mirrored);
  int bad = 0;
  cudaError_t ef = cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (ef == cudaSuccess) ef = cudaStreamSynchronize(ctx->stream);
  if (staging) {
    if (ef != cudaSuccess) cudaStreamSynchronize(ctx->stream);
    big_free(ctx, staging, staging_cap);   // the stream has been synchronised: nothing reads it any more
  }
  if (rc) return rc;
  MDB_CUDA(ctx, ef);
  ctx->last_input_nonfinite = bad;
  f->state = ST_MATRIX;
  f->prepared = false;
  return MDB_OK;
}

extern "C" int mdb_generate_diag_f2(mdb_ctx* ctx, int64_t n, uint64_t seed, double mu, double* diag_host) {
  if (!ctx || !diag_host) return MDB_ERR_INVALID;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  double* dd = nullptr;
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&dd, (size_t)std::max<int64_t>(n, 1) * 8));
  if (n > 0) {
    k_generate_f2_diag<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, seed, mu, dd);
    MDB_LAUNCHED(ctx);
  }
  cudaError_t e = cudaMemcpyAsync(diag_host, dd, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dd);
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

static int generate_f2(mdb_ctx* ctx, int64_t rows, int64_t cols, int64_t row0, const int64_t* prow,
                       const int64_t* pcol, uint64_t seed, double mu, double* out, int64_t ld) {
  if (rows <= 0 || cols <= 0) return MDB_OK;
  dim3 grid((unsigned)rows, (unsigned)((cols + 1023) / 1024), 1);
  k_generate_f2<<<grid, 256, 0, ctx->stream>>>(rows, cols, row0, prow, pcol, seed, mu, out, ld);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

extern "C" int mdb_generate_f2(mdb_ctx* ctx, int64_t n, uint64_t seed, double mu, double* A_dev, int64_t lda) {
  if (!ctx || !A_dev) return MDB_ERR_INVALID;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = generate_f2(ctx, n, n, 0, nullptr, nullptr, seed, mu, A_dev, lda);
  if (rc) return rc;
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MDB_OK;
}

extern "C" int mdb_factor_generate_f2(mdb_factor* f, uint64_t seed, double mu, const int64_t* p) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, f->batch == 1, "mdb_factor_generate_f2: batch must be 1");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = factor_set_p(f, p);
  if (rc) return rc;
  rc = generate_f2(ctx, f->n, f->n, 0, f->p_dev, f->p_dev, seed, mu, f->W, f->ld);
  if (rc) return rc;
  if (f->n > 0) {
    k_extract_diag<<<(unsigned)((f->n + 255) / 256), 256, 0, ctx->stream>>>(f->n, f->W, f->ld, f->gamma);
    MDB_LAUNCHED(ctx);
    MDB_CUDA(ctx, cudaGetLastError());
  }
  f->state = ST_MATRIX;
  f->prepared = false;
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// the blocked factorization
// ------------------------------------------------------------------------------------------------
static FactorDev make_dev(mdb_factor* f) {
  FactorDev F;
  memset(&F, 0, sizeof(F));
  F.n = f->n; F.ldw = f->ld; F.W = f->W;
  F.gamma = f->gamma; F.alpha = f->alpha; F.beta = f->beta; F.rowmax = f->rowmax;
  F.d = f->d; F.omega = f->omega; F.delta = f->delta; F.clamped = f->clamped;
  F.minB = f->minB; F.maxB = f->maxB; F.info = f->info; F.Ut = f->Ut; F.hdr = f->hdr; F.Vinv = f->Vinv;
  F.sW = f->n * f->ld; F.sVec = f->n; F.sUt = NB * NB; F.sHdr = 5 * NB; F.sV = MDB_VINV_DOUBLES; F.sPanel = f->panel_rows * f->agg * NB;
  return F;
}

static int configure_kernels(mdb_ctx* ctx) {
  static uint64_t done = 0;
  if (mdb_device_marked(done, ctx)) return MDB_OK;
  const int diag_smem = MDB_DIAG_SMEM_DOUBLES(NB) * 8;
  const int panel_smem = MDB_PANEL_SMEM_DOUBLES(NB) * 8;
  const int prep_smem = (NB * (NB + 1) / 2 + NB * (NB + 1)) * 8;
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_diag_block<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_diag_block<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_panel_gemm<NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem));
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_panel_gemm<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem));
  MDB_CUDA(ctx, cudaFuncSetAttribute(k_prepare_blocks<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, prep_smem));
  mdb_mark_device(done, ctx);
  return MDB_OK;
}

struct PhaseTimer {
  mdb_ctx* ctx;
  int phase;
  PhaseTimer(mdb_ctx* c, int ph) : ctx(c), phase(ph) {
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], ctx->stream);
  }
  ~PhaseTimer() {
    if (ctx->profiling) {
      cudaEventRecord(ctx->ev[1], ctx->stream);
      cudaEventSynchronize(ctx->ev[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
      ctx->phase_ms[phase] += ms;
    }
  }
};

extern "C" int mdb_factor_run(mdb_factor* f, const mdb_bounds* bnd) {
  if (!f || !bnd) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, f->state == ST_MATRIX, "mdb_factor_run: set a matrix first (state)");
  MDB_CHECK(ctx, bnd->rule == MDB_RULE_REIMER || bnd->rule == MDB_RULE_EXACT, "mdb_factor_run: unknown rule");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = configure_kernels(ctx);
  if (rc) return rc;
  const int64_t n = f->n, batch = f->batch, bn = batch * n;
  // per-index B bounds -> position order (Reimer.py:480-487 index them by p[j])
  if (f->minB) { cudaFree(f->minB); f->minB = nullptr; }
  if (f->maxB) { cudaFree(f->maxB); f->maxB = nullptr; }
  double sc_minB = -INFINITY, sc_maxB = INFINITY;
  if (bnd->rule == MDB_RULE_REIMER) {
    if (bnd->min_diag_B) sc_minB = bnd->min_diag_B[0];
    if (bnd->max_diag_B) sc_maxB = bnd->max_diag_B[0];
    if (bnd->stride_B != 0 && n > 0) {
      MDB_CHECK(ctx, bnd->min_diag_B && bnd->max_diag_B, "mdb_factor_run: vector bounds need both min_diag_B and max_diag_B");
      std::vector<double> mb((size_t)bn), Mb((size_t)bn);
      for (int64_t b = 0; b < batch; ++b)
        for (int64_t i = 0; i < n; ++i) {
          const int64_t pi = f->p_host.empty() ? i : f->p_host[(size_t)(b * n + i)];
          mb[(size_t)(b * n + i)] = bnd->min_diag_B[pi * bnd->stride_B];
          Mb[(size_t)(b * n + i)] = bnd->max_diag_B[pi * bnd->stride_B];
        }
      MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->minB, (size_t)bn * 8));
      MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->maxB, (size_t)bn * 8));
      MDB_CUDA(ctx, cudaMemcpyAsync(f->minB, mb.data(), (size_t)bn * 8, cudaMemcpyHostToDevice, ctx->stream));
      MDB_CUDA(ctx, cudaMemcpyAsync(f->maxB, Mb.data(), (size_t)bn * 8, cudaMemcpyHostToDevice, ctx->stream));
      MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  FactorDev F = make_dev(f);
  F.prm.rule = bnd->rule;
  F.prm.min_diag_D = bnd->min_diag_D;
  F.prm.max_diag_D = bnd->max_diag_D;
  F.prm.min_abs_value_D = bnd->min_abs_value_D;
  // plain LDL^T / Cholesky (MDB_RULE_EXACT) must not drop tiny entries
  F.prm.min_abs_value_L = (bnd->rule == MDB_RULE_EXACT) ? 0.0 : bnd->min_abs_value_L;
  F.prm.min_diag_B = sc_minB;
  F.prm.max_diag_B = sc_maxB;
  f->rule = bnd->rule;
  f->mavL = F.prm.min_abs_value_L;
  for (int i = 0; i < 4; ++i) ctx->phase_ms[i] = 0;
  ctx->gemm_flop[0] = ctx->gemm_flop[1] = 0;
  ctx->gemm_launches[0] = ctx->gemm_launches[1] = 0;

  if (bn > 0) {
    MDB_CUDA(ctx, cudaMemsetAsync(f->alpha, 0, (size_t)3 * bn * 8, ctx->stream));  // alpha, beta, rowmax contiguous
    MDB_CUDA(ctx, cudaMemsetAsync(f->info, 0, (size_t)batch * sizeof(int), ctx->stream));
  }
  const int diag_smem = MDB_DIAG_SMEM_DOUBLES(NB) * 8;
  const int panel_smem = MDB_PANEL_SMEM_DOUBLES(NB) * 8;
  // Two-level blocking.  Panels are NB = 128 columns wide (diagonal block + row-parallel panel kernel); `agg`
  // consecutive panels form an outer block of KB = agg * NB columns whose operands X, Y sit side by side in one
  // (rows x KB) buffer.  After panel j only the remaining columns of its outer block are updated (K = 128,
  // "inner" update); the rest of the trailing matrix gets one update per outer block with K = KB ("bulk"), which
  // is where n^3/3 of the flops go: the C tile is read and written once per KB instead of once per NB, and the
  // DMMA pipeline of a tile runs 4 x longer between its prologue and epilogue.
  // Lookahead-1 (default): while the side (low priority) stream runs the bulk update (b) of outer block O-1 on
  // the columns beyond outer block O, the main stream factors outer block O; then (a) = the bulk update of O on
  // the columns of outer block O+1 goes to the main stream and (b) to the side stream, concurrently (they touch
  // disjoint columns).  Operand buffers are double buffered by outer-block parity.  With profiling on,
  // everything runs serialised on the main stream so that the per-phase timers mean something.
  // The outer block width shrinks with the remaining size (agg_for): a wide block makes the bulk kernel more
  // efficient but lengthens the sequential work (agg diagonal blocks + panels) that one bulk update has to hide.
  const int64_t LDP = (int64_t)f->agg * NB;  // leading dimension of the operand buffers (widest outer block)
  auto agg_for = [&](int64_t remaining) -> int64_t {
    int a = 1;
    if (ctx->agg_fixed) a = ctx->agg_fixed;
    else if (remaining >= ctx->agg_thresh[2]) a = 8;
    else if (remaining >= ctx->agg_thresh[1]) a = 4;
    else if (remaining >= ctx->agg_thresh[0]) a = 2;
    return std::min<int64_t>(a, f->agg);
  };
  const bool lookahead = !ctx->profiling && ctx->lookahead && n > 2 * NB;
  cudaStream_t main_s = ctx->stream, side_s = ctx->side_stream;
  auto trailing = [&](const GemmArgs& ga, cudaStream_t st) -> int {  // TMA-fed kernel when it applies, else cp.async
    cudaStream_t saved = ctx->stream;
    ctx->stream = st;
    int r = -1;
    if (batch == 1) r = launch_gemm_tn_tma(ctx, ga, ga.M, ga.b_rows);
    if (r == -1) r = launch_gemm_tn<true, true, false>(ctx, ga, batch);
    ctx->stream = saved;
    return r;
  };
  // One-shot call with L streaming to pinned host memory: the last rows of L become final faster than PCIe drains them
  // (at n = 65536 the last 11 k rows, 5.9 GB, within the last 30 ms: a tail of 0.10 s).  Most of such a row is known long
  // before -- Lu[j, m] is final when the outer block of column m is done -- and equals L[j, m] unless omega_j != 1 or the
  // threshold removes the entry.  So the columns left of `r_tail` of the rows from r_tail on (the last quarter, at an
  // outer-block boundary) are shipped outer block by outer block as they are finished, the final copy of those rows only
  // carries the columns from r_tail on, k_finalize flags the rows whose early part turned out different, and those are
  // sent again at the end.  LDL / LDL_compressed only (LL scales columns as well).
  int64_t r_tail = n;
  uint8_t* tail_changed = nullptr;
  {
    const char* e = getenv("MDB200_TAIL_MIN");
    const int64_t tail_min = e ? atoll(e) : 16384;
    if (f->out_host && batch == 1 && n >= tail_min && f->out_stream_type != MDB_TYPE_LL) {
      for (int64_t ob = 0; ob < n;) {
        if (ob >= n - n / 4) { r_tail = ob; break; }
        ob = std::min<int64_t>(ob + agg_for(n - ob) * NB, n);
      }
      if (r_tail < n) {
        // (context-owned, grow-only: a stream-ordered allocation here cost 0.1 - 0.5 s per call when its pool was
        // trimmed at the next synchronisation, with ~100 GB of other allocations alive)
        if (ctx->tail_flags_cap < (size_t)n) {
          if (ctx->tail_flags) { MDB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream)); cudaFree(ctx->tail_flags); }
          ctx->tail_flags = nullptr; ctx->tail_flags_cap = 0;
          MDB_CUDA(ctx, dev_malloc(ctx, (void**)&ctx->tail_flags, (size_t)n));
          ctx->tail_flags_cap = (size_t)n;
        }
        tail_changed = ctx->tail_flags;
        MDB_CUDA(ctx, cudaMemsetAsync(tail_changed, 0, (size_t)n, ctx->copy_stream));
      }
    }
  }
  int64_t oidx = 0;
  for (int64_t ob = 0, KB = 0; ob < n; ob += KB, ++oidx) {
    KB = agg_for(n - ob) * NB;
    const int64_t oe = std::min<int64_t>(ob + KB, n);
    const int buf = lookahead ? (int)(oidx & 1) : 0;
    double* PX = buf ? f->PX2 : f->PX;
    double* PY = buf ? f->PY2 : f->PY;
    for (int64_t k0 = ob; k0 < oe; k0 += NB) {
      const int kb = (int)std::min<int64_t>(NB, n - k0);
      const int64_t k1 = k0 + kb;
      {
        PhaseTimer t(ctx, 0);
        auto kdiag = (bnd->rule == MDB_RULE_EXACT) ? k_diag_block<NB, true> : k_diag_block<NB, false>;
        kdiag<<<dim3(1, 1, (unsigned)batch), MDB_DIAG_THREADS(NB), diag_smem, main_s>>>(F, k0, kb, f->W + k0 * f->ld + k0, f->ld,
                                                                                        f->W + k0 * f->ld + k0, f->ld);
        MDB_LAUNCHED(ctx);
      }
      if (k1 >= n) break;
      const int64_t rows = n - k1;
      {
        PhaseTimer t(ctx, 1);
        dim3 grid((unsigned)((rows + PANEL_R - 1) / PANEL_R), 1, (unsigned)batch);
        k_panel_gemm<NB, false><<<grid, MDB_PANEL_THREADS, panel_smem, main_s>>>(F, k0, k1, rows, f->W, 1, f->ld, f->W + k0, f->ld,
                                                                    PX + (k0 - ob), PY + (k0 - ob), ob, LDP);
        MDB_LAUNCHED(ctx);
      }
      if (k1 < oe) {  // inner update: columns [k1, oe) of the outer block, K = NB
        PhaseTimer t(ctx, 3);
        GemmArgs gi;
        memset(&gi, 0, sizeof(gi));
        gi.A = PX + (k1 - ob) * LDP + (k0 - ob); gi.lda = LDP; gi.sA = F.sPanel;
        gi.B = PY + (k1 - ob) * LDP + (k0 - ob); gi.ldb = LDP; gi.sB = F.sPanel;
        gi.C = f->W + k1 * f->ld + k1; gi.ldc = f->ld; gi.sC = F.sW;
        gi.M = rows; gi.N = oe - k1; gi.K = NB;
        gi.row_glob0 = k1; gi.col_glob0 = k1; gi.b_glob0 = k1; gi.b_rows = rows;
        rc = trailing(gi, main_s);
        if (rc) return rc;
        {  // strictly-lower elements of the rows x (oe - k1) region, 2 K flop each
          const double M = (double)rows, N = (double)(oe - k1);
          ctx->gemm_flop[1] += 2.0 * NB * (N * M - 0.5 * N * (N + 1)) * (double)batch;
          ctx->gemm_launches[1] += 1;
        }
      }
    }
    if (f->out_host) {
      // rows [ob, oe) of L are final (nothing reads them or the A entries above their diagonal again): scale /
      // threshold them in place and send them home while the factorization goes on
      MDB_CUDA(ctx, cudaEventRecord(ctx->ev_copy, main_s));
      MDB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
      if (oe <= r_tail && r_tail < n) {
        // columns [ob, oe) of the late rows: final as Lu from now on (before k_finalize touches anything below)
        MDB_CUDA(ctx, cudaMemcpy2DAsync(f->out_host + r_tail * f->out_ld + ob, (size_t)f->out_ld * 8,
                                        f->W + r_tail * f->ld + ob, (size_t)f->ld * 8, (size_t)(oe - ob) * 8,
                                        (size_t)(n - r_tail), cudaMemcpyDeviceToHost, ctx->copy_stream));
      }
      dim3 grid((unsigned)(oe - ob), (unsigned)((n + 1023) / 1024), 1);
      const bool late = ob >= r_tail;
      k_finalize<<<grid, 256, 0, ctx->copy_stream>>>(n, f->W, f->ld, F.sW, f->d, f->omega, n, f->mavL, f->out_stream_type,
                                                     f->W, f->ld, F.sW, ob, late ? tail_changed : nullptr, r_tail);
      MDB_LAUNCHED(ctx);
      if (late && tail_changed && f->out_host_dev) {
        // rows whose early-shipped left part is stale go again at once, written by the device into the pinned matrix
        k_resend_rows<<<MDB_RESEND_CTAS, 256, 0, ctx->copy_stream>>>(tail_changed, f->W, f->ld, f->out_host_dev, f->out_ld, ob,
                                                                     oe - ob, r_tail);
        MDB_LAUNCHED(ctx);
      }
      const int64_t c_first = late ? r_tail : 0;   // the late rows only send what has not been shipped yet
      MDB_CUDA(ctx, cudaMemcpy2DAsync(f->out_host + ob * f->out_ld + c_first, (size_t)f->out_ld * 8,
                                      f->W + ob * f->ld + c_first, (size_t)f->ld * 8, (size_t)(n - c_first) * 8,
                                      (size_t)(oe - ob), cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    if (oe >= n) break;
    // bulk update of everything beyond the outer block, K = KB
    const int64_t rows = n - oe;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = PX + (oe - ob) * LDP; g.lda = LDP; g.sA = F.sPanel;
    g.B = PY + (oe - ob) * LDP; g.ldb = LDP; g.sB = F.sPanel;
    g.C = f->W + oe * f->ld + oe; g.ldc = f->ld; g.sC = F.sW;
    g.M = rows; g.N = rows; g.K = (int)(oe - ob);
    g.row_glob0 = oe; g.col_glob0 = oe; g.b_glob0 = oe; g.b_rows = rows;
    ctx->gemm_flop[0] += (double)g.K * (double)rows * (double)(rows - 1) * (double)batch;  // 2 K flop per element i > j
    ctx->gemm_launches[0] += 1;
    if (!lookahead) {
      PhaseTimer t(ctx, 2);
      rc = trailing(g, main_s);
      if (rc) return rc;
    } else {
      MDB_CUDA(ctx, cudaEventRecord(ctx->ev_panel[oidx & 1], main_s));  // operands of outer block O complete
      // (b) of O-1 must have reached the columns of outer block O+1 (and released the other operand buffer)
      if (oidx >= 1) MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_upd[(oidx - 1) & 1], 0));
      const int64_t KBn = agg_for(rows) * NB;  // width of the next outer block
      GemmArgs ga = g;
      ga.n_tile0 = 0; ga.n_tiles = KBn / gemm::BN;
      rc = trailing(ga, main_s);  // (a) columns of the next outer block, main stream
      if (rc) return rc;
      if (rows > KBn) {
        MDB_CUDA(ctx, cudaStreamWaitEvent(side_s, ctx->ev_panel[oidx & 1], 0));
        GemmArgs gb = g;
        gb.n_tile0 = KBn / gemm::BN; gb.n_tiles = 0; gb.m_tile0 = KBn / gemm::BM;
        rc = trailing(gb, side_s);  // (b) the rest, side stream, concurrently with (a)
        if (rc) return rc;
      }
      MDB_CUDA(ctx, cudaEventRecord(ctx->ev_upd[oidx & 1], side_s));
    }
  }
  if (lookahead && oidx >= 1) {
    MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_upd[(oidx - 1) & 1], 0));
    if (oidx >= 2) MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_upd[(oidx - 2) & 1], 0));
  }
  MDB_CUDA(ctx, cudaGetLastError());
  f->state = ST_FACTORED;
  f->prepared = false;
  static const bool run_timing = getenv("MDB200_TIMING") != nullptr;
  auto wall = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double tw0 = wall();
  double tw1 = tw0, tw2 = tw0;
  int64_t resent = 0;
  if (run_timing) { cudaStreamSynchronize(main_s); tw1 = wall(); }
  if (f->out_host && tail_changed && (!f->out_host_dev || run_timing)) {
    // late rows whose early-shipped columns are not what L holds (omega != 1, thresholded entries): send them again --
    // only when the device cannot address the host matrix itself (else k_resend_rows has done it on the way; with
    // MDB200_TIMING the flags are still read for the count)
    std::vector<uint8_t> ch((size_t)(n - r_tail));
    cudaError_t e = cudaMemcpyAsync(ch.data(), tail_changed + r_tail, (size_t)(n - r_tail), cudaMemcpyDeviceToHost, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->copy_stream);
    for (int64_t j = r_tail; j < n && e == cudaSuccess;) {
      if (!ch[(size_t)(j - r_tail)]) { ++j; continue; }
      int64_t j1 = j;
      while (j1 < n && ch[(size_t)(j1 - r_tail)]) ++j1;
      if (!f->out_host_dev)
        e = cudaMemcpy2DAsync(f->out_host + j * f->out_ld, (size_t)f->out_ld * 8, f->W + j * f->ld, (size_t)f->ld * 8,
                              (size_t)r_tail * 8, (size_t)(j1 - j), cudaMemcpyDeviceToHost, ctx->copy_stream);
      resent += j1 - j;
      j = j1;
    }
    if (run_timing) tw2 = wall();
    MDB_CUDA(ctx, e);
  }
  if (f->out_host) {  // every row has been finalised and copied by the copy stream
    MDB_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_copy, 0));
    f->state = ST_FINAL;
    f->out_type = f->out_stream_type;
    f->out_host = nullptr;
    f->out_host_dev = nullptr;
  }
  if (run_timing) {
    cudaStreamSynchronize(main_s);
    if (tail_changed)
      fprintf(stderr, "[mdb200] run n=%lld: factorization done %.3f s after the last launch was issued, copy stream idle "
                      "%.3f s later, remaining copies %.3f s (r_tail %lld, rows sent again %lld)\n", (long long)n,
              tw1 - tw0, tw2 - tw1, wall() - tw2, (long long)r_tail, (long long)resent);
    else
      fprintf(stderr, "[mdb200] run n=%lld: factorization done %.3f s after the last launch was issued, download %.3f s "
                      "later\n", (long long)n, tw1 - tw0, wall() - tw1);
  }
  return MDB_OK;
}

static int factor_finalize(mdb_factor* f, int out_type) {
  mdb_ctx* ctx = f->ctx;
  if (f->state == ST_FINAL) {
    if (f->out_type != out_type) return mdb_fail(ctx, MDB_ERR_STATE, "%s", "factor already finalised with another type");
    return MDB_OK;
  }
  if (f->state != ST_FACTORED) return mdb_fail(ctx, MDB_ERR_STATE, "%s", "factor not computed yet");
  if (f->n > 0) {
    dim3 grid((unsigned)f->n, (unsigned)((f->n + 1023) / 1024), (unsigned)f->batch);
    k_finalize<<<grid, 256, 0, ctx->stream>>>(f->n, f->W, f->ld, f->n * f->ld, f->d, f->omega, f->n, f->mavL, out_type,
                                              f->W, f->ld, f->n * f->ld, 0);
    MDB_LAUNCHED(ctx);
    MDB_CUDA(ctx, cudaGetLastError());
  }
  f->state = ST_FINAL;
  f->out_type = out_type;
  return MDB_OK;
}

extern "C" int mdb_factor_get(mdb_factor* f, int out_type, double* L, int64_t ldl, int l_loc, double* d, double* omega,
                              double* delta, uint8_t* clamped, int64_t* info) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = MDB_OK;
  if (out_type != MDB_TYPE_NONE) rc = factor_finalize(f, out_type);   // MDB_TYPE_NONE: vectors only, W stays un-finalised
  else if (L) return mdb_fail(ctx, MDB_ERR_INVALID, "%s", "mdb_factor_get: MDB_TYPE_NONE cannot return L");
  if (rc) return rc;
  const int64_t n = f->n, batch = f->batch, bn = batch * n;
  if (L && n > 0) {
    MDB_CHECK(ctx, ldl >= n, "mdb_factor_get: ldl < n");
    if (f->prepared) {  // a solve mirrored L into the upper triangle: undo that (the next solve mirrors it again)
      dim3 grid((unsigned)n, (unsigned)((n + 1023) / 1024));
      k_zero_upper<<<grid, 256, 0, ctx->stream>>>(n, f->W, f->ld);
      MDB_LAUNCHED(ctx);
      f->prepared = false;
    }
    // rows of consecutive batch members are contiguous in both layouts when viewed as (batch*n) x n
    rc = copy2d_async(ctx, L, ldl, f->W, f->ld, bn, n, l_loc, MDB_DEVICE);
    if (rc) return rc;
  }
  if (d && bn) MDB_CUDA(ctx, cudaMemcpyAsync(d, f->d, (size_t)bn * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (clamped && bn) MDB_CUDA(ctx, cudaMemcpyAsync(clamped, f->clamped, (size_t)bn, cudaMemcpyDeviceToHost, ctx->stream));
  if ((omega || delta) && bn) {
    // position -> original index (Reimer.py:538,545); alpha / beta are free scratch after the run
    for (int64_t b = 0; b < batch; ++b) {
      const int64_t* pb = f->p_dev ? f->p_dev + b * n : nullptr;
      k_scatter_by_p<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, pb, f->omega + b * n, f->alpha + b * n);
      k_scatter_by_p<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, pb, f->delta + b * n, f->beta + b * n);
      ctx->launches += 2;
    }
    if (omega) MDB_CUDA(ctx, cudaMemcpyAsync(omega, f->alpha, (size_t)bn * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (delta) MDB_CUDA(ctx, cudaMemcpyAsync(delta, f->beta, (size_t)bn * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  std::vector<int> inf((size_t)batch, 0);
  if (info) MDB_CUDA(ctx, cudaMemcpyAsync(inf.data(), f->info, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (info)
    for (int64_t b = 0; b < batch; ++b) info[b] = inf[(size_t)b];
  return MDB_OK;
}

// B of positive_semidefinite_matrix from an un-finalised factor (Reimer.py:899-992)
extern "C" int mdb_factor_psd_matrix(mdb_factor* f, const double* min_diag_B, const double* max_diag_B, int64_t stride_B,
                                     double* B, int64_t ldb, int loc) {
  if (!f || !B) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, f->state == ST_FACTORED, "mdb_factor_psd_matrix: needs a factorised, not yet finalised factor");
  MDB_CHECK(ctx, f->batch == 1 && f->rule == MDB_RULE_REIMER, "mdb_factor_psd_matrix: one approximate factorization");
  MDB_CHECK(ctx, ldb >= f->n, "mdb_factor_psd_matrix: ldb < n");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = f->n;
  if (n == 0) return MDB_OK;
  double *mB = nullptr, *MB = nullptr, *Bd = (loc == MDB_DEVICE) ? B : nullptr;
  int64_t* q = nullptr;
  const int64_t nb = stride_B != 0 ? n : 1, ldd = (loc == MDB_DEVICE) ? ldb : n;
  cudaError_t e = cudaSuccess;
  if (min_diag_B) {
    e = dev_malloc(ctx, (void**)&mB, (size_t)nb * 8);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(mB, 8, min_diag_B, (size_t)std::max<int64_t>(stride_B, 1) * 8, 8, (size_t)nb, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess && max_diag_B) {
    e = dev_malloc(ctx, (void**)&MB, (size_t)nb * 8);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(MB, 8, max_diag_B, (size_t)std::max<int64_t>(stride_B, 1) * 8, 8, (size_t)nb, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess && f->p_dev) {
    e = dev_malloc(ctx, (void**)&q, (size_t)n * 8);
    if (e == cudaSuccess) {
      k_invert_permutation<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, f->p_dev, q);
      MDB_LAUNCHED(ctx);
    }
  }
  if (e == cudaSuccess && !Bd) e = dev_malloc(ctx, (void**)&Bd, (size_t)n * n * 8);
  int rc = MDB_OK;
  if (e == cudaSuccess) {
    dim3 grid((unsigned)n, (unsigned)((n + 1023) / 1024));
    k_psd_matrix<<<grid, 256, 0, ctx->stream>>>(n, f->W, f->ld, f->gamma, f->delta, f->omega, f->d, q, mB, MB,
                                                stride_B != 0 ? 1 : 0, f->mavL, Bd, ldd);
    MDB_LAUNCHED(ctx);
    e = cudaGetLastError();
    if (e == cudaSuccess && loc != MDB_DEVICE) rc = copy2d_async(ctx, B, ldb, Bd, ldd, n, n, MDB_HOST, MDB_DEVICE);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(mB); cudaFree(MB); cudaFree(q);
  if (loc != MDB_DEVICE) cudaFree(Bd);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// one-shot entry points
// ------------------------------------------------------------------------------------------------
static int one_shot(mdb_ctx* ctx, int64_t batch, int64_t n, const double* A, int64_t lda, const int64_t* p,
                    const mdb_bounds* bounds, int out_type, double* L, int64_t ldl, double* d, double* omega,
                    double* delta, uint8_t* clamped, int64_t* info, int loc, mdb_factor** factor_out) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, n >= 0, "n < 0");
  MDB_CHECK(ctx, out_type == MDB_TYPE_LDL || out_type == MDB_TYPE_LDL_COMPRESSED || out_type == MDB_TYPE_LL ||
                     (out_type == MDB_TYPE_NONE && factor_out && !L),
            "unknown decomposition type");
  mdb_factor* f = nullptr;
  // MDB200_TIMING=1: wall-clock of the stages of a one-shot call on stderr (debugging aid)
  static const bool timing = getenv("MDB200_TIMING") != nullptr;
  auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t0 = now();
  int rc = mdb_factor_create(ctx, n, batch, &f);
  if (rc) return rc;
  if (timing) cudaStreamSynchronize(ctx->stream);
  const double t1 = now();
  rc = mdb_factor_set_matrix(f, A, lda, p, loc);
  if (timing) cudaStreamSynchronize(ctx->stream);
  const double t2 = now();
  bool streamed = false;
  // Streaming L out needs PINNED host memory: a device -> pageable copy blocks the calling thread until the data is
  // there, which would stall the launch loop of the factorization behind every row block.
  bool l_pinned = false;
  double* l_dev = nullptr;
  if (!rc && loc == MDB_HOST && L) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, L) == cudaSuccess) {
      l_pinned = (attr.type == cudaMemoryTypeHost);
      if (l_pinned && !getenv("MDB200_NO_ZEROCOPY")) l_dev = (double*)attr.devicePointer;   // (the switch is for the tests)
    } else {
      cudaGetLastError();
    }
  }
  if (!rc && loc == MDB_HOST && L && l_pinned && batch == 1 && n >= 4096 && ldl >= n) {
    f->out_host = L; f->out_ld = ldl; f->out_stream_type = out_type;  // L leaves the device row block by row block
    f->out_host_dev = l_dev;
    streamed = true;
  }
  if (!rc) rc = mdb_factor_run(f, bounds);
  if (timing) cudaStreamSynchronize(ctx->stream);
  const double t3 = now();
  if (!rc) rc = mdb_factor_get(f, out_type, streamed ? nullptr : L, ldl, loc, d, omega, delta, clamped, info);
  const double t4 = now();
  if (rc || !factor_out) {
    // the results are complete; the tens of GB of device buffers are released in the background
    cudaStreamSynchronize(ctx->stream);
    ctx_reap(ctx);
    factor_detach_W(f);
    mdb_factor* dead = f;
    if (!rc && ctx->use_cache) {
      factor_park(dead);
    } else {
      ctx->reaper = new (std::nothrow) std::thread([dead]() { factor_release(dead); });
      if (!ctx->reaper) factor_release(dead);
    }
    f = nullptr;
  }
  if (timing)
    fprintf(stderr, "[mdb200] one-shot n=%lld: create %.3f s, set_matrix %.3f s, run %.3f s, get %.3f s, destroy %.3f s\n",
            (long long)n, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now() - t4);
  if (factor_out) *factor_out = f;
  return rc;
}

extern "C" int mdb_approximate_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, const int64_t* p,
                                   const mdb_bounds* bounds, int out_type, double* L, int64_t ldl, double* d,
                                   double* omega, double* delta, uint8_t* clamped, int loc, mdb_factor** factor_out) {
  if (!ctx || !bounds) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, bounds->rule == MDB_RULE_REIMER, "mdb_approximate_f64: bounds->rule must be MDB_RULE_REIMER");
  return one_shot(ctx, 1, n, A, lda, p, bounds, out_type, L, ldl, d, omega, delta, clamped, nullptr, loc, factor_out);
}

extern "C" int mdb_approximate_dynamic_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, int method,
                                           const mdb_bounds* bnd, int out_type, double* L, int64_t ldl, int64_t* p_out,
                                           double* d, double* omega, double* delta, uint8_t* clamped, int loc,
                                           mdb_factor** factor_out) {
  if (!ctx || !bnd) return MDB_ERR_INVALID;
  if (factor_out) *factor_out = nullptr;
  MDB_CHECK(ctx, method == 1 || method == 2, "mdb_approximate_dynamic_f64: method must be 1 or 2");
  MDB_CHECK(ctx, bnd->rule == MDB_RULE_REIMER, "mdb_approximate_dynamic_f64: bounds->rule must be MDB_RULE_REIMER");
  MDB_CHECK(ctx, out_type == MDB_TYPE_LDL || out_type == MDB_TYPE_LDL_COMPRESSED || out_type == MDB_TYPE_LL ||
                     (out_type == MDB_TYPE_NONE && factor_out && !L),
            "unknown decomposition type");
  mdb_factor* f = nullptr;
  int rc = mdb_factor_create(ctx, n, 1, &f);
  if (rc) return rc;
  rc = mdb_factor_set_matrix(f, A, lda, nullptr, loc);
  double *minB = nullptr, *maxB = nullptr;
  DynCand* cand = nullptr;
  DynSel* sel = nullptr;
  std::vector<int64_t> pfin((size_t)std::max<int64_t>(n, 1));
  if (!rc && n > 0) {
    // per-index B bounds in position order (the start permutation is the identity, Reimer.py:443)
    std::vector<double> mb((size_t)n), Mb((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      mb[(size_t)i] = bnd->min_diag_B ? bnd->min_diag_B[i * bnd->stride_B] : -INFINITY;
      Mb[(size_t)i] = bnd->max_diag_B ? bnd->max_diag_B[i * bnd->stride_B] : INFINITY;
    }
    const int nb_max = (int)((n + MDB_DYN_ROWS - 1) / MDB_DYN_ROWS) + (int)((n + 255) / 256) + 2;   // candidates: one per block of k_dyn_select1 / k_dyn_column
    cudaError_t e = dev_malloc(ctx, (void**)&minB, (size_t)n * 8);
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&maxB, (size_t)n * 8);
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&cand, (size_t)nb_max * sizeof(DynCand));
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&sel, 2 * sizeof(DynSel) + 32);
    if (e == cudaSuccess && !f->p_dev) e = dev_malloc(ctx, (void**)&f->p_dev, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(minB, mb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(maxB, Mb.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->alpha, 0, (size_t)3 * n * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "mdb_approximate_dynamic_f64: %s", cudaGetErrorString(e));
    if (!rc) {
      k_fill_iota<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, f->p_dev);
      unsigned* counter = (unsigned*)(sel + 2);   // blocks of k_dyn_column that have finished
      cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream);
      FactorDev F = make_dev(f);
      F.minB = minB; F.maxB = maxB;
      F.prm.rule = MD_RULE_REIMER;
      F.prm.min_diag_D = bnd->min_diag_D; F.prm.max_diag_D = bnd->max_diag_D;
      F.prm.min_abs_value_D = bnd->min_abs_value_D; F.prm.min_abs_value_L = bnd->min_abs_value_L;
      F.prm.min_diag_B = -INFINITY; F.prm.max_diag_B = INFINITY;
      f->rule = MDB_RULE_REIMER;
      f->mavL = bnd->min_abs_value_L;
      // panels of NB columns: pending updates are applied on the fly inside a panel (k_dyn_column), the trailing
      // matrix gets one K = NB tensor-core update per panel
      const int64_t LDP = (int64_t)f->agg * NB;
      // Every full panel but the last replays ONE captured CUDA graph of 128 column kernel nodes: the kernel reads the
      // column index from the decision record of its parity (written by the step before), so the nodes are the same
      // for every panel and the ~4 us gaps of dependent stream launches shrink to the graph's node-to-node latency.
      cudaGraphExec_t gexec = nullptr;
      const bool use_graph = n >= 3 * NB && !getenv("MDB200_DYN_NOGRAPH");
      const unsigned gswap = (unsigned)((n + 255) / 256);   // blocks for the positions left of the pivot (k_dyn_column)
      if (use_graph) {
        cudaGraph_t graph = nullptr;
        cudaError_t ge = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        if (ge == cudaSuccess) {
          const unsigned gcol = (unsigned)((n - 1 + MDB_DYN_ROWS - 1) / MDB_DYN_ROWS);   // rows below the very first column
          for (int c = 0; c < NB; ++c)
            k_dyn_column<<<gcol + gswap, 256, 0, ctx->stream>>>(F, method, sel, c & 1, (int)gcol, f->PX, f->PY, LDP, cand,
                                                                counter, minB, maxB, f->p_dev);
          ge = cudaStreamEndCapture(ctx->stream, &graph);
        }
        if (ge == cudaSuccess) ge = cudaGraphInstantiate(&gexec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ge != cudaSuccess) { cudaGetLastError(); gexec = nullptr; }
      }
      for (int64_t pb = 0; pb < n && !rc; pb += NB) {
        const int64_t pe = std::min<int64_t>(pb + NB, n);
        if (pb == 0) {  // the first column is selected here, later ones by the column kernel of the step before
          const int nblk = (int)((n + 255) / 256);
          k_dyn_select1<<<nblk, 256, 0, ctx->stream>>>(F, method, 0, cand);
          k_dyn_select2<<<1, 256, 0, ctx->stream>>>(F, method, 0, nblk, cand, minB, maxB, f->p_dev, sel);
          ctx->launches += 2;
        }
        if (gexec && pe < n && pe - pb == NB) {
          cudaGraphLaunch(gexec, ctx->stream);
          ctx->launches += NB;
        } else {
          for (int64_t i = pb; i < pe; ++i) {
            const int64_t rem = n - i;
            if (rem > 1) {   // (the last index is its own pivot: nothing to swap, nothing below)
              const unsigned gcol = (unsigned)((rem - 1 + MDB_DYN_ROWS - 1) / MDB_DYN_ROWS);
              k_dyn_column<<<gcol + (unsigned)((i + 255) / 256), 256, 0, ctx->stream>>>(
                  F, method, sel, (int)(i & 1), (int)gcol, f->PX, f->PY, LDP, cand, counter, minB, maxB, f->p_dev);
              ctx->launches += 1;
            }
          }
        }
        if (pe < n) {
          GemmArgs g;
          memset(&g, 0, sizeof(g));
          const int64_t rows = n - pe;
          g.A = f->PX + pe * LDP; g.lda = LDP;
          g.B = f->PY + pe * LDP; g.ldb = LDP;
          g.C = f->W + pe * f->ld + pe; g.ldc = f->ld;
          g.M = rows; g.N = rows; g.K = NB;
          g.row_glob0 = pe; g.col_glob0 = pe; g.b_glob0 = pe; g.b_rows = rows;
          int r = launch_gemm_tn_tma(ctx, g, rows, rows);
          if (r == -1) r = launch_gemm_tn<true, true, false>(ctx, g, 1);
          rc = r;
        }
      }
      cudaError_t e2 = cudaGetLastError();
      if (gexec) { cudaStreamSynchronize(ctx->stream); cudaGraphExecDestroy(gexec); }
      if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(pfin.data(), f->p_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
      if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
      if (e2 != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "mdb_approximate_dynamic_f64: %s", cudaGetErrorString(e2));
      f->p_host.assign(pfin.begin(), pfin.begin() + n);
      f->state = ST_FACTORED;
    }
  } else if (!rc) {
    f->state = ST_FACTORED;
  }
  cudaFree(minB); cudaFree(maxB); cudaFree(cand); cudaFree(sel);
  if (!rc) rc = mdb_factor_get(f, out_type, L, ldl, loc, d, omega, delta, clamped, nullptr);
  if (!rc && p_out)
    for (int64_t i = 0; i < n; ++i) p_out[i] = pfin[(size_t)i];
  if (rc || !factor_out) {
    mdb_factor_destroy(f);
    f = nullptr;
  }
  if (factor_out) *factor_out = f;
  return rc;
}

// GMW family (GMW_SE.py:6-35, :67-152): d and delta of the modified LDL^T without pivoting; B = A + diag(delta) and
// the optional diagonal rescaling (:43-62) are O(n^2) and stay with the caller.
static int gmw_family_run(mdb_ctx* ctx, int family, int64_t n, const double* A, int64_t lda, int loc, int use_two_phases,
                          int nondecreasing, int use_abs, double mu, double min_diag_D, double* d_out, double* delta_out);

extern "C" int mdb_gmw_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, int loc, int use_two_phases,
                           int nondecreasing, int use_abs, double mu, double min_diag_D, double* d_out,
                           double* delta_out) {
  return gmw_family_run(ctx, 0, n, A, lda, loc, use_two_phases, nondecreasing, use_abs, mu, min_diag_D, d_out, delta_out);
}

// SE family (GMW_SE.py:155-194, :374-558): the same driver with the choose_d of _SE
extern "C" int mdb_se_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, int loc, int use_two_phases,
                          int nondecreasing, int use_abs, double mu, double min_diag_D, double* d_out,
                          double* delta_out) {
  return gmw_family_run(ctx, 1, n, A, lda, loc, use_two_phases, nondecreasing, use_abs, mu, min_diag_D, d_out, delta_out);
}

static int gmw_family_run(mdb_ctx* ctx, int family, int64_t n, const double* A, int64_t lda, int loc, int use_two_phases,
                          int nondecreasing, int use_abs, double mu, double min_diag_D, double* d_out, double* delta_out) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, n >= 0 && (n == 0 || (A && lda >= n)), "mdb_gmw_f64: bad matrix");
  MDB_CHECK(ctx, d_out && delta_out, "mdb_gmw_f64: null output");
  MDB_CHECK(ctx, min_diag_D > 0, "mdb_gmw_f64: min_diag_D must be positive (GMW_SE.py:81)");
  if (n == 0) return MDB_OK;
  ctx_reap(ctx);
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  double *W = nullptr, *cvec = nullptr, *dd = nullptr;
  GmwState* st = nullptr;
  cudaError_t e = dev_malloc(ctx, (void**)&W, (size_t)n * n * 8);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&cvec, (size_t)n * 8);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&dd, (size_t)2 * n * 8);
  if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&st, sizeof(GmwState));
  int rc = MDB_OK;
  if (e == cudaSuccess) rc = copy2d_async(ctx, W, n, A, lda, n, n, MDB_DEVICE, loc);
  if (e == cudaSuccess && !rc) {
    GmwState h;
    memset(&h, 0, sizeof(h));
    h.phase = use_two_phases ? 10 : 15;                       // :84
    h.min_diag_D = min_diag_D;
    h.use_abs = use_abs; h.nondecreasing = nondecreasing;
    h.family = family;
    {
      const double eps = 2.220446049250313e-16;
      h.tau = pow(eps, isnan(mu) ? 1.0 / 3.0 : 2.0 / 3.0);            // GMW_SE.py:158-161
      h.eps13 = pow(eps, 1.0 / 3.0);
    }
    if (!isnan(mu)) {                                         // :87-90: needs max |diag A|, a host-side pass over n values
      std::vector<double> diag((size_t)n);
      e = cudaMemcpy2DAsync(diag.data(), 8, W, (size_t)(n + 1) * 8, 8, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      double mx = 0;
      for (int64_t i = 0; i < n; ++i) mx = std::max(mx, fabs(diag[(size_t)i]));
      h.min_d_next = -mu * mx;
      h.min_d_factor = -mu;
    } else {                                                  // :91-93
      h.min_d_next = min_diag_D;
      h.min_d_factor = -INFINITY;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // `h` goes out of scope
    for (int64_t k = 0; k < n && e == cudaSuccess; ++k) {
      const int64_t m = n - k - 1;
      k_gmw_reduce<<<1, 1024, 0, ctx->stream>>>(n, W, n, k, cvec, st);   // + the phase test
      k_gmw_init_decide<<<(unsigned)std::max<int64_t>(m, 1), 256, 0, ctx->stream>>>(n, W, n, k, st, dd, dd + n);
      ctx->launches += 2;
      if (m > 0) {
        const unsigned t32 = (unsigned)((m + 31) / 32);
        k_gmw_rank1<<<dim3(t32, t32), 256, 0, ctx->stream>>>(n, W, n, k, cvec, st);
        ctx->launches += 1;
      }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_out, dd, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(delta_out, dd + n, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(W); cudaFree(cvec); cudaFree(dd); cudaFree(st);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

extern "C" int mdb_approximate_batched_f64(mdb_ctx* ctx, int64_t batch, int64_t n, const double* A, int64_t lda,
                                           const int64_t* p, const mdb_bounds* bounds, double* L, int64_t ldl,
                                           double* d, double* omega, double* delta, uint8_t* clamped, int loc) {
  if (!ctx || !bounds) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, batch >= 1, "batch < 1");
  MDB_CHECK(ctx, lda == n && (ldl == n || !L), "batched mode expects densely packed matrices (lda == ldl == n)");
  return one_shot(ctx, batch, n, A, lda, p, bounds, MDB_TYPE_LDL, L, ldl, d, omega, delta, clamped, nullptr, loc,
                  nullptr);
}

extern "C" int mdb_decompose_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, const int64_t* p, int out_type,
                                 double* L, int64_t ldl, double* d, int64_t* info, int loc, mdb_factor** factor_out) {
  if (!ctx) return MDB_ERR_INVALID;
  mdb_bounds b;
  memset(&b, 0, sizeof(b));
  b.rule = MDB_RULE_EXACT;
  b.min_diag_D = 0; b.max_diag_D = INFINITY; b.min_abs_value_D = 0; b.min_abs_value_L = 0;
  int64_t inf = 0;
  mdb_factor* f = nullptr;
  int rc = one_shot(ctx, 1, n, A, lda, p, &b, out_type, L, ldl, d, nullptr, nullptr, nullptr, &inf, loc, &f);
  if (info) *info = inf;
  if (!rc && inf > 0) {
    rc = MDB_ERR_NOT_POSITIVE_DEFINITE;
    snprintf(ctx->err, sizeof(ctx->err), "leading minor of order %lld is not positive definite", (long long)inf);
  }
  if (rc || !factor_out) {
    if (f) mdb_factor_destroy(f);
    f = nullptr;
  }
  if (factor_out) *factor_out = f;
  return rc;
}

extern "C" int mdb_permute_sym_f64(mdb_ctx* ctx, int64_t n, const double* A, int64_t lda, const int64_t* p, double* B,
                                   int64_t ldb, int loc) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, n >= 0 && lda >= n && ldb >= n, "mdb_permute_sym_f64: bad dimensions");
  if (n == 0) return MDB_OK;
  MDB_CHECK(ctx, A && B, "mdb_permute_sym_f64: null matrix");
  ctx_reap(ctx);
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t* p_dev = nullptr;
  double *A_dev = nullptr, *B_dev = nullptr;
  int rc = MDB_OK;
  cudaError_t e = cudaSuccess;
  if (p) {
    std::vector<uint8_t> seen((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {
      if (p[i] < 0 || p[i] >= n || seen[(size_t)p[i]]) return mdb_fail(ctx, MDB_ERR_INVALID, "%s", "p is not a permutation");
      seen[(size_t)p[i]] = 1;
    }
    e = dev_malloc(ctx, (void**)&p_dev, (size_t)n * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p_dev, p, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
  }
  const double* src = A;
  double* dst = B;
  int64_t lsrc = lda, ldst = ldb;
  if (e == cudaSuccess && loc == MDB_HOST) {
    e = dev_malloc(ctx, (void**)&A_dev, (size_t)n * n * 8);
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&B_dev, (size_t)n * n * 8);
    if (e == cudaSuccess) rc = copy2d_async(ctx, A_dev, n, A, lda, n, n, MDB_DEVICE, MDB_HOST);
    src = A_dev; dst = B_dev; lsrc = n; ldst = n;
  }
  if (e == cudaSuccess && !rc) rc = gather_sym(ctx, n, 1, src, lsrc, 0, p_dev, dst, ldst, 0, nullptr);
  if (e == cudaSuccess && !rc && loc == MDB_HOST) rc = copy2d_async(ctx, B, ldb, B_dev, n, n, n, MDB_HOST, MDB_DEVICE);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(p_dev); cudaFree(A_dev); cudaFree(B_dev);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// upload, solve, multiply
// ------------------------------------------------------------------------------------------------
extern "C" int mdb_factor_upload(mdb_ctx* ctx, int64_t n, int type, const double* L, int64_t ldl, const double* d,
                                 const int64_t* p, mdb_factor** out) {
  if (!ctx || !out) return MDB_ERR_INVALID;
  *out = nullptr;
  MDB_CHECK(ctx, type == MDB_TYPE_LDL || type == MDB_TYPE_LL, "mdb_factor_upload: type must be LDL or LL");
  MDB_CHECK(ctx, (L || n == 0) && ldl >= n, "mdb_factor_upload: bad L");
  MDB_CHECK(ctx, type == MDB_TYPE_LL || d || n == 0, "mdb_factor_upload: LDL needs d");
  mdb_factor* f = nullptr;
  int rc = mdb_factor_create(ctx, n, 1, &f);
  if (rc) return rc;
  rc = factor_set_p(f, p);
  if (!rc && n > 0) rc = copy2d_async(ctx, f->W, f->ld, L, ldl, n, n, MDB_DEVICE, MDB_HOST);
  if (!rc && n > 0 && d) {
    cudaError_t e = cudaMemcpyAsync(f->d, d, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "%s", cudaGetErrorString(e));
  }
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "%s", cudaGetErrorString(e));
  }
  if (rc) { mdb_factor_destroy(f); return rc; }
  f->state = ST_FINAL;
  f->out_type = type;
  *out = f;
  return MDB_OK;
}

static int factor_prepare(mdb_factor* f) {
  mdb_ctx* ctx = f->ctx;
  if (f->prepared) return MDB_OK;
  if (f->state != ST_FINAL) return mdb_fail(ctx, MDB_ERR_STATE, "%s", "solve/multiply need a finalised factor");
  if (f->batch != 1) return mdb_fail(ctx, MDB_ERR_STATE, "%s", "solve/multiply are not available for batched factors");
  int rc = configure_kernels(ctx);
  if (rc) return rc;
  const int64_t n = f->n;
  if (n == 0) { f->prepared = true; return MDB_OK; }
  const int64_t T = (n + NB - 1) / NB;
  const size_t bytes = (size_t)T * NB * NB * 8;
  if (!f->Linv) {
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->Linv, bytes));
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->LinvT, bytes));
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->Ltri, bytes));
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->Utri, bytes));
  }
  const int64_t t32 = (n + 31) / 32;
  k_symmetrize<<<dim3((unsigned)t32, (unsigned)t32), 256, 0, ctx->stream>>>(n, f->W, f->ld);
  MDB_LAUNCHED(ctx);
  const int prep_smem = (NB * (NB + 1) / 2 + NB * (NB + 1)) * 8;
  k_prepare_blocks<NB><<<(unsigned)T, NB, prep_smem, ctx->stream>>>(n, f->W, f->ld, f->out_type != MDB_TYPE_LL ? 1 : 0,
                                                                   f->Linv, f->LinvT, f->Ltri, f->Utri);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  f->prepared = true;
  return MDB_OK;
}

// C[0:M, 0:N] = beta C + sign * A (M x K) B^T (N x K)
static int product_tn(mdb_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
                      int64_t ldb, double* C, int64_t ldc, bool beta1, bool neg) {
  if (M <= 0 || N <= 0 || K <= 0) return MDB_OK;
  if (M <= 8 && (K % 128) == 0) {
    k_gemv_tn<8><<<(unsigned)((N + 7) / 8), 256, 0, ctx->stream>>>((int)M, N, K, A, lda, B, ldb, C, ldc, beta1 ? 1.0 : 0.0,
                                                                  neg ? -1.0 : 1.0);
    MDB_LAUNCHED(ctx);
    MDB_CUDA(ctx, cudaGetLastError());
    return MDB_OK;
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = (int)K;
  if (beta1 || !neg) {
    // TMA-fed tile kernel on the full rectangle: a row origin beyond every column makes each element "below the diagonal"
    g.row_glob0 = (int64_t)1 << 40; g.b_rows = N;
    const int r = launch_gemm_tn_tma(ctx, g, M, N, (neg ? 1 : 0) | (beta1 ? 0 : 2));
    if (r != -1) return r;
    g.row_glob0 = 0; g.b_rows = 0;
  }
  if (beta1 && neg) return launch_gemm_tn<false, true, true>(ctx, g, 1);
  if (beta1) return launch_gemm_tn<false, true, false>(ctx, g, 1);
  return launch_gemm_tn<false, false, false>(ctx, g, 1);  // neg without beta is never needed
}

// One persistent launch per triangular sweep (kernels_solve.cuh, k_trsv_persistent).
static int trsv_prepare(mdb_factor* f, int nrhs) {
  mdb_ctx* ctx = f->ctx;
  const int T = (int)((f->n + NB - 1) / NB);
  const int NC = (T + trsv::CH - 1) / trsv::CH;
  if (!f->trsv_tasks) {
    // claim order: diag(t), then every partial task whose last needed block is t (so that a task only ever
    // waits for tasks claimed before it)
    std::vector<int2> tasks;
    for (int t = 0; t < T; ++t) {  // (the diag tasks themselves run on dedicated CTAs)
      if ((t + 1) % trsv::CH == 0) {
        for (int I = t + 2; I < T; ++I) tasks.push_back(make_int2(I, (t + 1) / trsv::CH - 1));
      } else if (t + 2 < T) {
        tasks.push_back(make_int2(t + 2, t / trsv::CH));
      }
    }
    f->trsv_ntasks = (int)tasks.size();
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->trsv_tasks, std::max<size_t>(tasks.size(), 1) * sizeof(int2)));
    MDB_CUDA(ctx, cudaMemcpyAsync(f->trsv_tasks, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `tasks` goes out of scope
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->trsv_sync, 2 * sizeof(unsigned int)));
  }
  if (f->trsv_partial_rhs < nrhs) {
    if (f->trsv_partial) cudaFree(f->trsv_partial);
    f->trsv_partial = nullptr;
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->trsv_partial, (size_t)T * NC * nrhs * NB * 8));
    f->trsv_partial_rhs = nrhs;
  }
  static uint64_t configured = 0;
  if (!mdb_device_marked(configured, ctx)) {
    const int smem = (NB * NB + 2 * 8 * NB) * 8;
    MDB_CUDA(ctx, cudaFuncSetAttribute(k_trsv_persistent<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MDB_CUDA(ctx, cudaFuncSetAttribute(k_trsv_persistent<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mdb_mark_device(configured, ctx);
  }
  return MDB_OK;
}

static int trsv_sweep(mdb_factor* f, int nrhs, bool reverse, const double* in, double* out, int64_t ldt,
                      const double* scale_d) {
  mdb_ctx* ctx = f->ctx;
  const int T = (int)((f->n + NB - 1) / NB);
  TrsvArgs a;
  memset(&a, 0, sizeof(a));
  a.W = f->W; a.ldw = f->ld; a.n = f->n; a.T = T;
  a.DinvT = reverse ? f->Linv : f->LinvT;
  a.in = in; a.out = out; a.ldt = ldt; a.nrhs = nrhs; a.scale_d = scale_d;
  a.partial = f->trsv_partial; a.NC = (T + trsv::CH - 1) / trsv::CH;
  a.next_task = f->trsv_sync;
  a.err = (int*)(f->trsv_sync + 1);
  a.tasks = f->trsv_tasks; a.ntasks = f->trsv_ntasks; a.reverse = reverse ? 1 : 0;
  MDB_CUDA(ctx, cudaMemsetAsync(f->trsv_sync, 0, 2 * sizeof(unsigned int), ctx->stream));
  {  // the data is its own "ready" flag: output and partial sums start out as the sentinel
    const int64_t n_out = (int64_t)nrhs * ldt, n_part = (int64_t)T * a.NC * nrhs * NB;
    k_fill_sentinel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(n_out, out);
    k_fill_sentinel<<<(unsigned)((n_part + 255) / 256), 256, 0, ctx->stream>>>(n_part, f->trsv_partial);
    ctx->launches += 2;
  }
  { const char* e = getenv("MDB200_TRSV_ND"); a.nd = e ? atoi(e) : 16; if (a.nd < 1) a.nd = 1; if (a.nd > 64) a.nd = 64; }
  const int grid = std::max(a.nd + 1, std::min(ctx->num_sms, f->trsv_ntasks + a.nd));
  const int smem = (NB * NB + 2 * 8 * NB) * 8;
  if (nrhs == 1) k_trsv_persistent<1><<<grid, trsv::THREADS, smem, ctx->stream>>>(a);
  else k_trsv_persistent<8><<<grid, trsv::THREADS, smem, ctx->stream>>>(a);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

static int trsv_check(mdb_factor* f) {
  mdb_ctx* ctx = f->ctx;
  int err = 0;
  MDB_CUDA(ctx, cudaMemcpyAsync(&err, f->trsv_sync + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (err) return mdb_fail(ctx, MDB_ERR_CUDA, "%s", "k_trsv_persistent: a dependency wait ran into its spin limit "
                                                    "(a right-hand side holding the kernel's sentinel NaN would do that)");
  return MDB_OK;
}

static int ws_get(mdb_factor* f, int i, size_t bytes, double** out) {
  mdb_ctx* ctx = f->ctx;
  if (f->ws_cap[i] < bytes) {
    if (f->ws[i]) cudaFree(f->ws[i]);
    f->ws[i] = nullptr; f->ws_cap[i] = 0;
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&f->ws[i], bytes));
    f->ws_cap[i] = bytes;
  }
  *out = f->ws[i];
  return MDB_OK;
}

// Bt / Yt / stage are workspaces owned by the factor (not to be freed by the caller)
static int rhs_in(mdb_factor* f, const double* B, int64_t ldb, int64_t nrhs, int loc, double** Bt, double** Yt,
                  int64_t* ldt_out, double** stage) {
  mdb_ctx* ctx = f->ctx;
  const int64_t n = f->n, ldt = mdb_round_up(n, NB);
  *ldt_out = ldt;
  int rcw = ws_get(f, 0, (size_t)nrhs * ldt * 8, Bt);
  if (!rcw) rcw = ws_get(f, 1, (size_t)nrhs * ldt * 8, Yt);
  if (rcw) return rcw;
  MDB_CUDA(ctx, cudaMemsetAsync(*Bt, 0, (size_t)nrhs * ldt * 8, ctx->stream));
  MDB_CUDA(ctx, cudaMemsetAsync(*Yt, 0, (size_t)nrhs * ldt * 8, ctx->stream));
  const double* src = B;
  int64_t lsrc = ldb;
  *stage = nullptr;
  if (loc == MDB_HOST) {
    rcw = ws_get(f, 2, (size_t)n * nrhs * 8, stage);
    if (rcw) return rcw;
    int rc = copy2d_async(ctx, *stage, nrhs, B, ldb, n, nrhs, MDB_DEVICE, MDB_HOST);
    if (rc) return rc;
    src = *stage;
    lsrc = nrhs;
  }
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nrhs + 31) / 32));
  k_rhs_gather_t<<<grid, 256, 0, ctx->stream>>>(n, nrhs, src, lsrc, f->p_dev, *Bt, ldt);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

static int rhs_out(mdb_factor* f, const double* Xt, int64_t ldt, int64_t nrhs, double* X, int64_t ldx, int loc,
                   double* stage) {
  mdb_ctx* ctx = f->ctx;
  const int64_t n = f->n;
  double* dst = X;
  int64_t ldst = ldx;
  if (loc == MDB_HOST) { dst = stage; ldst = nrhs; }
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nrhs + 31) / 32));
  k_rhs_scatter_t<<<grid, 256, 0, ctx->stream>>>(n, nrhs, Xt, ldt, f->p_dev, dst, ldst);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  if (loc == MDB_HOST) {
    int rc = copy2d_async(ctx, X, ldx, stage, nrhs, n, nrhs, MDB_HOST, MDB_DEVICE);
    if (rc) return rc;
  }
  return MDB_OK;
}

extern "C" int mdb_factor_solve(mdb_factor* f, const double* B, int64_t ldb, int64_t nrhs, double* X, int64_t ldx,
                                int loc) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, nrhs >= 1 && ldb >= nrhs && ldx >= nrhs, "mdb_factor_solve: bad right-hand side shape");
  MDB_CHECK(ctx, f->out_type == MDB_TYPE_LDL || f->out_type == MDB_TYPE_LL || f->out_type == MDB_TYPE_LDL_COMPRESSED,
            "mdb_factor_solve: bad factor type");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = factor_prepare(f);
  if (rc) return rc;
  const int64_t n = f->n;
  if (n == 0) return MDB_OK;
  double *Bt = nullptr, *Yt = nullptr, *stage = nullptr;
  int64_t ldt = 0;
  rc = rhs_in(f, B, ldb, nrhs, loc, &Bt, &Yt, &ldt, &stage);
  const int64_t T = (n + NB - 1) / NB;
  if (!rc && nrhs <= 8 && ctx->use_trsv && T >= 2) {
    // bandwidth-bound case: one persistent launch per sweep, L streamed once
    rc = trsv_prepare(f, (int)nrhs);
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], ctx->stream);
    if (!rc) rc = trsv_sweep(f, (int)nrhs, false, Bt, Yt, ldt, nullptr);                                   // L z = b
    if (!rc) rc = trsv_sweep(f, (int)nrhs, true, Yt, Bt, ldt, f->out_type != MDB_TYPE_LL ? f->d : nullptr);  // L^T x = z / d
    if (ctx->profiling) {  // device time of the two sweeps -> mdb_ctx_get_phase_ms()[0]
      cudaEventRecord(ctx->ev[1], ctx->stream);
      cudaEventSynchronize(ctx->ev[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
      ctx->phase_ms[0] = ms;
    }
    if (!rc) rc = rhs_out(f, Bt, ldt, nrhs, X, ldx, loc, stage);
    if (!rc) rc = trsv_check(f);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    MDB_CUDA(ctx, e2);
    return MDB_OK;
  }
  if (ctx->profiling) cudaEventRecord(ctx->ev[0], ctx->stream);
  // Many right-hand sides: blocked sweeps on the DMMA tile kernel, two-level like the factorization.  Inside a
  // super block of SB diagonal blocks the updates have K = 128; everything beyond it gets ONE update with
  // K = 128 SB, so the right-hand sides are read and written once per super block.
  const int64_t SB = 4;
  // Lookahead, as in the factorization: the update of super block S is split into (a) the rows of the next super
  // block, on the main stream, and (b) everything beyond, on the side stream; the dependent chain of small products
  // of the next super block then runs while (b) keeps the SMs busy.  Every row block still receives its updates in
  // the same order (b_0 ... b_{r-2}, a_{r-1}), so the results do not depend on the overlap.
  const bool overlap = ctx->lookahead;  // the profiling events bracket both sweeps on the main stream, which joins the side stream
  cudaStream_t main_s = ctx->stream, side_s = overlap ? ctx->side_stream : ctx->stream;
  auto on_side = [&](auto&& fn) -> int {
    ctx->stream = side_s;
    const int r = fn();
    ctx->stream = main_s;
    return r;
  };
  int64_t si = 0;          // super blocks processed so far (both sweeps): event slots alternate
  bool side_busy = false;  // a (b) update is in flight on the side stream
  auto split_update = [&](int64_t rows_a, int64_t rows_b, auto&& upd_a, auto&& upd_b) -> int {
    // main: (a) after the previous (b) has finished (both touch the rows of the next super block)
    int r = MDB_OK;
    if (overlap) {
      MDB_CUDA(ctx, cudaEventRecord(ctx->ev_panel[si & 1], main_s));  // this super block's solution is complete
      if (side_busy) MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_upd[(si - 1) & 1], 0));
    }
    if (rows_a > 0) r = upd_a();
    if (!r && rows_b > 0) {
      if (overlap) MDB_CUDA(ctx, cudaStreamWaitEvent(side_s, ctx->ev_panel[si & 1], 0));
      r = on_side(upd_b);
      if (overlap) {
        MDB_CUDA(ctx, cudaEventRecord(ctx->ev_upd[si & 1], side_s));
        side_busy = true;
      }
    } else if (overlap) {
      side_busy = false;
    }
    ++si;
    return r;
  };
  auto join_side = [&]() -> int {
    if (overlap && side_busy) MDB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_upd[(si - 1) & 1], 0));
    side_busy = false;
    return MDB_OK;
  };
  // forward sweep  L z = b  (decompositions.py:987 / :1349)
  for (int64_t S0 = 0; S0 < T && !rc; S0 += SB) {
    const int64_t S1 = std::min<int64_t>(S0 + SB, T);
    const int64_t s0 = S0 * NB, s1 = std::min<int64_t>(S1 * NB, n);
    for (int64_t J = S0; J < S1 && !rc; ++J) {
      const int64_t j0 = J * NB, j1 = std::min<int64_t>(j0 + NB, n);
      rc = product_tn(ctx, nrhs, NB, NB, Bt + j0, ldt, f->Linv + J * NB * NB, NB, Yt + j0, ldt, false, false);
      if (!rc && j1 < s1)
        rc = product_tn(ctx, nrhs, s1 - j1, NB, Yt + j0, ldt, f->W + j1 * f->ld + j0, f->ld, Bt + j1, ldt, true, true);
    }
    if (!rc && s1 < n) {
      const int64_t K = (S1 - S0) * NB;
      const int64_t s2 = std::min<int64_t>(s1 + SB * NB, n);   // end of the next super block
      rc = split_update(
          s2 - s1, n - s2,
          [&]() { return product_tn(ctx, nrhs, s2 - s1, K, Yt + s0, ldt, f->W + s1 * f->ld + s0, f->ld, Bt + s1, ldt, true, true); },
          [&]() { return product_tn(ctx, nrhs, n - s2, K, Yt + s0, ldt, f->W + s2 * f->ld + s0, f->ld, Bt + s2, ldt, true, true); });
    }
  }
  if (!rc) rc = join_side();
  if (!rc && f->out_type != MDB_TYPE_LL) {  // x / d  (decompositions.py:988)
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nrhs);
    k_scale_by_d<<<grid, 256, 0, ctx->stream>>>(n, nrhs, Yt, ldt, f->d, 0);
    MDB_LAUNCHED(ctx);
  }
  // backward sweep  L^T x = z  (decompositions.py:989 / :1350): rows of U = L^T from the mirrored upper triangle
  for (int64_t S1 = T; S1 > 0 && !rc; S1 -= std::min<int64_t>(SB, S1)) {
    const int64_t S0 = S1 > SB ? S1 - SB : 0;
    const int64_t s0 = S0 * NB;
    for (int64_t J = S1 - 1; J >= S0 && !rc; --J) {
      const int64_t j0 = J * NB;
      rc = product_tn(ctx, nrhs, NB, NB, Yt + j0, ldt, f->LinvT + J * NB * NB, NB, Bt + j0, ldt, false, false);
      if (!rc && j0 > s0)
        rc = product_tn(ctx, nrhs, j0 - s0, NB, Bt + j0, ldt, f->W + s0 * f->ld + j0, f->ld, Yt + s0, ldt, true, true);
    }
    if (!rc && s0 > 0) {
      const int64_t K = (S1 - S0) * NB;
      const int64_t sp = s0 > SB * NB ? s0 - SB * NB : 0;      // start of the previous super block
      rc = split_update(
          s0 - sp, sp,
          [&]() { return product_tn(ctx, nrhs, s0 - sp, K, Bt + s0, ldt, f->W + sp * f->ld + s0, f->ld, Yt + sp, ldt, true, true); },
          [&]() { return product_tn(ctx, nrhs, sp, K, Bt + s0, ldt, f->W + s0, f->ld, Yt, ldt, true, true); });
    }
  }
  if (!rc) rc = join_side();
  if (ctx->profiling) {  // device time of the two sweeps -> mdb_ctx_get_phase_ms()[0]
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaEventSynchronize(ctx->ev[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->phase_ms[0] = ms;
  }
  if (!rc) rc = rhs_out(f, Bt, ldt, nrhs, X, ldx, loc, stage);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

extern "C" int mdb_factor_multiply(mdb_factor* f, const double* B, int64_t ldb, int64_t nrhs, double* X, int64_t ldx,
                                   int loc) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, nrhs >= 1 && ldb >= nrhs && ldx >= nrhs, "mdb_factor_multiply: bad right-hand side shape");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = factor_prepare(f);
  if (rc) return rc;
  const int64_t n = f->n;
  if (n == 0) return MDB_OK;
  double *Xt = nullptr, *Zt = nullptr, *stage = nullptr;
  int64_t ldt = 0;
  rc = rhs_in(f, B, ldb, nrhs, loc, &Xt, &Zt, &ldt, &stage);
  const int64_t T = (n + NB - 1) / NB;
  // z = L^T x  (decompositions.py:964 / :1327): block row J of U = L^T
  for (int64_t J = 0; J < T && !rc; ++J) {
    const int64_t j0 = J * NB, j1 = j0 + NB, nb_rows = std::min<int64_t>(NB, n - j0);
    rc = product_tn(ctx, nrhs, nb_rows, NB, Xt + j0, ldt, f->Utri + J * NB * NB, NB, Zt + j0, ldt, false, false);
    if (!rc && j1 < ldt)
      rc = product_tn(ctx, nrhs, nb_rows, ldt - j1, Xt + j1, ldt, f->W + j0 * f->ld + j1, f->ld, Zt + j0, ldt, true, false);
  }
  if (!rc && f->out_type != MDB_TYPE_LL) {  // d * z  (decompositions.py:965)
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)nrhs);
    k_scale_by_d<<<grid, 256, 0, ctx->stream>>>(n, nrhs, Zt, ldt, f->d, 1);
    MDB_LAUNCHED(ctx);
  }
  // y = L z  (decompositions.py:966 / :1328), result back into Xt
  for (int64_t J = 0; J < T && !rc; ++J) {
    const int64_t j0 = J * NB, nb_rows = std::min<int64_t>(NB, n - j0);
    rc = product_tn(ctx, nrhs, nb_rows, NB, Zt + j0, ldt, f->Ltri + J * NB * NB, NB, Xt + j0, ldt, false, false);
    if (!rc && j0 > 0)
      rc = product_tn(ctx, nrhs, nb_rows, j0, Zt, ldt, f->W + j0 * f->ld, f->ld, Xt + j0, ldt, true, false);
  }
  if (!rc) rc = rhs_out(f, Xt, ldt, nrhs, X, ldx, loc, stage);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

// B = P^T L D L^T P (LDL types) / P^T L L^T P (LL): the matrix the decomposition stands for (decompositions.py:826-830,
// :1202-1205: `L @ diag(d) @ L.H` and the inverse permutation).  Block column J of the lower triangle is
//   C[j0:n, J] = L[j0:n, 0:j1] (L[J, 0:j1] D)^T
// two tile-kernel products per block column (the diagonal tile with the clean diagonal block, the rows below straight
// from W): n^3 / 3 multiply-adds in total, a quarter of multiplying the factor with an identity matrix, and nothing
// but the result crosses PCIe.  The upper triangle is the mirror image; the permutation is undone by the gather kernel.
extern "C" int mdb_factor_composed(mdb_factor* f, double* B, int64_t ldb, int loc) {
  if (!f) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, B && ldb >= f->n, "mdb_factor_composed: bad output");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = factor_prepare(f);
  if (rc) return rc;
  const int64_t n = f->n;
  if (n == 0) return MDB_OK;
  const int64_t T = (n + NB - 1) / NB, ldt = f->panel_rows;
  double *C = nullptr, *O = nullptr;
  int64_t* q = nullptr;
  size_t capC = 0, capO = 0;
  cudaError_t e = big_malloc(ctx, (void**)&C, (size_t)n * n * 8, &capC);
  const bool permuted = f->p_dev != nullptr;
  if (e == cudaSuccess && permuted) e = big_malloc(ctx, (void**)&O, (size_t)n * n * 8, &capO);
  if (e == cudaSuccess && permuted) e = dev_malloc(ctx, (void**)&q, (size_t)n * 8);
  if (e != cudaSuccess) {
    big_free(ctx, C, capC); big_free(ctx, O, capO); cudaFree(q);
    return mdb_fail(ctx, MDB_ERR_CUDA, "mdb_factor_composed: %s", cudaGetErrorString(e));
  }
  const int use_d = f->out_type != MDB_TYPE_LL ? 1 : 0;
  for (int64_t J = 0; J < T && !rc; ++J) {
    const int64_t j0 = J * NB, j1 = j0 + NB, nb = std::min<int64_t>(NB, n - j0);
    k_composed_operands<<<dim3((unsigned)((j1 + 255) / 256), NB), 256, 0, ctx->stream>>>(
        n, f->W, f->ld, f->Ltri + J * NB * NB, f->d, use_d, j0, NB, f->PX, f->PY, ldt);
    MDB_LAUNCHED(ctx);
    rc = product_tn(ctx, nb, nb, j1, f->PX, ldt, f->PY, ldt, C + j0 * n + j0, n, false, false);
    if (!rc && j1 < n)
      rc = product_tn(ctx, n - j1, nb, j1, f->W + j1 * f->ld, f->ld, f->PY, ldt, C + j1 * n + j0, n, false, false);
  }
  if (!rc) {
    const int64_t t32 = (n + 31) / 32;
    k_symmetrize<<<dim3((unsigned)t32, (unsigned)t32), 256, 0, ctx->stream>>>(n, C, n);
    MDB_LAUNCHED(ctx);
    const double* res = C;
    if (permuted) {   // B[p[i], p[j]] = C[i, j]
      k_invert_permutation<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, f->p_dev, q);
      MDB_LAUNCHED(ctx);
      rc = gather_sym(ctx, n, 1, C, n, 0, q, O, n, 0, nullptr);
      res = O;
    }
    if (!rc) rc = copy2d_async(ctx, B, ldb, res, n, n, n, loc, MDB_DEVICE);
  }
  e = cudaStreamSynchronize(ctx->stream);
  big_free(ctx, C, capC); big_free(ctx, O, capO); cudaFree(q);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

extern "C" int mdb_factor_identity_residual(mdb_factor* f, int nprobe, uint64_t seed, double* residual) {
  if (!f || !residual) return MDB_ERR_INVALID;
  mdb_ctx* ctx = f->ctx;
  MDB_CHECK(ctx, f->state == ST_FACTORED && f->batch == 1, "identity check runs after mdb_factor_run and before mdb_factor_get");
  MDB_CHECK(ctx, f->rule == MDB_RULE_REIMER || f->rule == MDB_RULE_EXACT, "bad rule");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = f->n;
  *residual = 0;
  if (n == 0) return MDB_OK;
  // Right-hand side of the identity from the un-finalised W (upper triangle = A); then finalise a COPY-free
  // way is not possible, so the left-hand side L D L^T v is evaluated here with two triangular sweeps over
  // the unscaled rows: L[j, m] = thresh(omega_j Lu[j, m]).  To keep this check simple and independent of the
  // solve path it finalises the factor (LDL) afterwards and uses mdb_factor_multiply.
  double *v = nullptr, *yb = nullptr, *ya = nullptr, *lhs = nullptr;
  unsigned long long* mx = nullptr;
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&v, (size_t)nprobe * n * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&yb, (size_t)nprobe * n * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&ya, (size_t)nprobe * n * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&lhs, (size_t)nprobe * n * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&mx, 16));
  MDB_CUDA(ctx, cudaMemsetAsync(yb, 0, (size_t)nprobe * n * 8, ctx->stream));
  MDB_CUDA(ctx, cudaMemsetAsync(ya, 0, (size_t)nprobe * n * 8, ctx->stream));
  MDB_CUDA(ctx, cudaMemsetAsync(mx, 0, 16, ctx->stream));
  for (int q = 0; q < nprobe; ++q) {
    k_fill_random<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, seed + 7919ull * q, v + (size_t)q * n);
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 1023) / 1024));
    k_identity_rhs<<<grid, 256, 0, ctx->stream>>>(n, f->W, f->ld, f->gamma, f->delta, f->omega, v + (size_t)q * n,
                                                  yb + (size_t)q * n, ya + (size_t)q * n);
    ctx->launches += 2;
  }
  MDB_CUDA(ctx, cudaGetLastError());
  int rc = factor_finalize(f, MDB_TYPE_LDL);
  // vectors here are in POSITION order already; multiply expects original order and applies p, so bypass p
  int64_t* saved_p = f->p_dev;
  f->p_dev = nullptr;
  // probes as an n x nprobe row-major matrix: transpose the nprobe x n layout through the solve helpers
  double* vcol = nullptr;
  double* lcol = nullptr;
  if (!rc) {
    cudaError_t e = dev_malloc(ctx, (void**)&vcol, (size_t)nprobe * n * 8);
    if (e == cudaSuccess) e = dev_malloc(ctx, (void**)&lcol, (size_t)nprobe * n * 8);
    if (e != cudaSuccess) rc = mdb_fail(ctx, MDB_ERR_CUDA, "%s", cudaGetErrorString(e));
  }
  if (!rc) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nprobe + 31) / 32));
    k_rhs_scatter_t<<<grid, 256, 0, ctx->stream>>>(n, nprobe, v, n, nullptr, vcol, nprobe);  // vcol[i, q] = v[q, i]
    MDB_LAUNCHED(ctx);
    rc = mdb_factor_multiply(f, vcol, nprobe, nprobe, lcol, nprobe, MDB_DEVICE);
  }
  if (!rc) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nprobe + 31) / 32));
    k_rhs_gather_t<<<grid, 256, 0, ctx->stream>>>(n, nprobe, lcol, nprobe, nullptr, lhs, n);  // lhs[q, i] = lcol[i, q]
    MDB_LAUNCHED(ctx);
    const int64_t tot = (int64_t)nprobe * n;
    k_max_abs_diff<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(tot, lhs, yb, ya, mx);
    MDB_LAUNCHED(ctx);
  }
  f->p_dev = saved_p;
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, mx, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(v); cudaFree(yb); cudaFree(ya); cudaFree(lhs); cudaFree(mx); cudaFree(vcol); cudaFree(lcol);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  double num, den;
  memcpy(&num, &h[0], 8);
  memcpy(&den, &h[1], 8);
  *residual = den > 0 ? num / den : num;
  return MDB_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection / unit-test hooks
// ------------------------------------------------------------------------------------------------
extern "C" void* mdb_factor_device_ptr(mdb_factor* f, int which) {
  if (!f) return nullptr;
  switch (which) {
    case 0: return f->W;
    case 1: return f->gamma;
    case 2: return f->alpha;
    case 3: return f->beta;
    case 4: return f->rowmax;
    case 5: return f->d;
    case 6: return f->omega;
    case 7: return f->delta;
    default: return nullptr;
  }
}
extern "C" int64_t mdb_factor_ld(const mdb_factor* f) { return f ? f->ld : 0; }

extern "C" int mdb_debug_gemm_tn(mdb_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda,
                                 const double* B, int64_t ldb, double* C, int64_t ldc, int lower, int beta1, int neg,
                                 int64_t row_glob0, int64_t col_glob0, int reps, double* ms) {
  if (!ctx) return MDB_ERR_INVALID;
  MDB_CHECK(ctx, M > 0 && N > 0 && K > 0 && K % 16 == 0, "mdb_debug_gemm_tn: bad shape");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&dA, (size_t)M * K * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&dB, (size_t)N * K * 8));
  MDB_CUDA(ctx, dev_malloc(ctx, (void**)&dC, (size_t)M * N * 8));
  int rc = copy2d_async(ctx, dA, K, A, lda, M, K, MDB_DEVICE, MDB_HOST);
  if (!rc) rc = copy2d_async(ctx, dB, K, B, ldb, N, K, MDB_DEVICE, MDB_HOST);
  if (!rc) rc = copy2d_async(ctx, dC, N, C, ldc, M, N, MDB_DEVICE, MDB_HOST);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = dA; g.lda = K; g.B = dB; g.ldb = K; g.C = dC; g.ldc = N; g.M = M; g.N = N; g.K = (int)K;
  g.row_glob0 = row_glob0; g.col_glob0 = col_glob0;
  if (reps < 1) reps = 1;
  cudaEventRecord(ctx->ev[0], ctx->stream);
  for (int r = 0; r < reps && !rc; ++r) {
    if (lower) {
      rc = launch_gemm_tn_tma(ctx, g, M, N);
      if (rc == -1) rc = launch_gemm_tn<true, true, false>(ctx, g, 1);
    }
    else rc = product_tn(ctx, M, N, K, dA, K, dB, K, dC, N, beta1 != 0, neg != 0);   // the solve / multiply products
  }
  cudaEventRecord(ctx->ev[1], ctx->stream);
  if (!rc) rc = copy2d_async(ctx, C, ldc, dC, N, M, N, MDB_HOST, MDB_DEVICE);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  float t = 0;
  cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
  if (ms) *ms = (double)t / reps;
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  if (rc) return rc;
  MDB_CUDA(ctx, e);
  return MDB_OK;
}

#include "mdb_dist.cuh"
#include "mdb_dist_solve.cuh"
