// mdb_peer.cuh -- panel exchange of the multi-GPU factorization as peer-memory stores (NVLink / NVSwitch) fused into
// the owner's pack step, instead of a NCCL broadcast followed by an unpack kernel on every receiver.
// Included at the end of mdb_api.cu, after mdb_dist.cuh.
//
// EXPERIMENTAL, off by default (distributed.py enables it with MDB200_P2P=1): one run on 2 B200 at the end of round 1
// gave results bit-identical to the single-GPU path (tools/dist_check.py); not yet measured at scale, so the NCCL path
// stays the default.
//
// Every rank owns one "symmetric" buffer of mdb_dist_peer_buffer_elems() doubles, mapped into all peers
// (torch.distributed._symmetric_memory does the mapping; the library only sees the base pointers):
//
//     [ OPX0 | OPY0 | OPX1 | OPY1 | vec (7 n) | clamped (n bytes, padded) | arrive (one u64 per panel) | released (one u64 per rank) ]
//
// Owner of panel k, after k_diag_block / k_panel:  k_dist_push  reads its own X (already in its OPX), the panel
// header and its alpha / beta / rowmax slices and stores X, Y = -X d, the slices and d / omega / delta / clamped
// straight into every peer's buffers -- the receivers need no unpack pass -- then, after a system-scope fence, the last
// CTA to finish raises arrive[k] on every peer.  A receiver's stream runs k_dist_wait(arrive[k]) in front of the
// kernels that consume panel k.
//
// Flow control (one-sided stores need it, a collective does not): the operand buffers are double buffered by outer-block
// parity, so panel data of outer block O may only land on a peer once that peer has finished every kernel that read the
// operands of outer block O - 2.  Each rank announces that with k_dist_release (released[rank] on every peer := number
// of outer blocks released, after its bulk update and its main-stream part (a) of that outer block); k_dist_push waits
// for released[q] >= O - 1 of every peer q before its first store.
//
// All waits are bounded (about two seconds); a timeout sets info = -(1000 + code) and lets the kernel proceed, so
// a protocol error ends in an error status instead of a hung GPU.  Values are tagged with the run's epoch, so nothing
// has to be cleared between factorizations (a barrier between mdb_dist_set_bounds and the first panel is still needed,
// see distributed.py).
#pragma once

struct mdb_peer_layout {
  int64_t opx[2], opy[2], vec, clamped, arrive, released, total;  // offsets in doubles
};

static mdb_peer_layout peer_layout(const mdb_dist* s) {
  mdb_peer_layout L;
  const int64_t op = s->n * s->ldp;
  int64_t o = 0;
  for (int i = 0; i < 2; ++i) { L.opx[i] = o; o += op; L.opy[i] = o; o += op; }
  L.vec = o; o += 7 * s->n;
  L.clamped = o; o += (s->n + 7) / 8;
  L.arrive = o; o += s->nblk_glob;
  L.released = o; o += s->world;
  L.total = mdb_round_up(o, 2);
  return L;
}

extern "C" int64_t mdb_dist_peer_buffer_elems(const mdb_dist* s) { return s ? peer_layout(s).total : 0; }

#define MDB_PEER_MAX 16

struct PeerPtrs {
  double* base[MDB_PEER_MAX];  // symmetric buffer of every rank as mapped into this process (base[rank] = own)
  int world, rank;
};

__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// true when *p >= want (same epoch in the upper 32 bits) within the spin budget
__device__ __forceinline__ bool peer_spin_ge(const unsigned long long* p, unsigned long long want) {
  for (long long it = 0; it < 4000000; ++it) {  // ~ 2 s with the sleep below
    if (peer_ld_acquire(p) >= want) return true;
    __nanosleep(500);
  }
  return false;
}

// Owner: X (rows x NB, inside its own operand buffer) and the state slices -> every peer.
__global__ void __launch_bounds__(256) k_dist_push(int64_t rows, const double* __restrict__ OX, int64_t ldp,
                                                   const double* __restrict__ hdr, const double* __restrict__ alpha,
                                                   const double* __restrict__ beta, const double* __restrict__ rowmax,
                                                   PeerPtrs P, int64_t off_x, int64_t off_y, int64_t off_alpha,
                                                   int64_t n, int64_t k0, int kb, int64_t off_vec, int64_t off_clamped,
                                                   int64_t off_arrive_k, int64_t off_released,
                                                   unsigned long long need_released, unsigned long long arrive_value,
                                                   unsigned int* counter, int* info) {
  __shared__ bool s_last;
  // flow control: every peer has released the operand buffer this panel goes into
  if (threadIdx.x < P.world && threadIdx.x != P.rank && need_released != 0) {
    const unsigned long long* rel = (const unsigned long long*)(P.base[P.rank] + off_released) + threadIdx.x;
    if (!peer_spin_ge(rel, need_released)) atomicCAS(info, 0, -1001);
  }
  __syncthreads();
  const int c = threadIdx.x & (MDB_NB - 1);
  const double dc = hdr[c];
  if (blockIdx.x == 0 && threadIdx.x < kb) {  // d / omega / delta / clamped of the panel's columns
    const double dv = hdr[threadIdx.x], ov = hdr[MDB_NB + threadIdx.x], ev = hdr[3 * MDB_NB + threadIdx.x];
    const uint8_t cv = (uint8_t)(hdr[4 * MDB_NB + threadIdx.x] != 0.0);
    for (int q = 0; q < P.world; ++q) {   // 4 x 128 values per panel
      if (q == P.rank) continue;
      double* v = P.base[q] + off_vec;
      v[4 * n + k0 + threadIdx.x] = dv;
      v[5 * n + k0 + threadIdx.x] = ov;
      v[6 * n + k0 + threadIdx.x] = ev;
      ((uint8_t*)(P.base[q] + off_clamped))[k0 + threadIdx.x] = cv;
    }
  }
  for (int64_t r = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); r < rows; r += (int64_t)gridDim.x * 2) {
    const double x = OX[r * ldp + c];
    const double y = -x * dc;
    for (int q = 0; q < P.world; ++q) {
      if (q == P.rank) continue;
      P.base[q][off_x + r * ldp + c] = x;
      P.base[q][off_y + r * ldp + c] = y;
    }
    if (c == 0) {
      const double a = alpha[r], b = beta[r], m = rowmax[r];
      for (int q = 0; q < P.world; ++q) {
        if (q == P.rank) continue;
        double* v = P.base[q] + off_alpha + r;   // alpha at vec + n, beta at vec + 2 n, rowmax at vec + 3 n
        v[0] = a; v[n] = b; v[2 * n] = m;
      }
    }
  }
  // all stores of this CTA are visible system-wide before it is counted; the last CTA raises the flags
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < P.world) {
    if (threadIdx.x == 0) *counter = 0;
    peer_st_release((unsigned long long*)(P.base[threadIdx.x] + off_arrive_k), arrive_value);
  }
}

__global__ void k_dist_wait(const unsigned long long* flag, unsigned long long value, int* info) {
  if (threadIdx.x == 0 && !peer_spin_ge(flag, value)) atomicCAS(info, 0, -1002);
  __threadfence_system();
}

__global__ void k_dist_release(PeerPtrs P, int64_t off_released, unsigned long long value) {
  __threadfence_system();
  if (threadIdx.x < P.world)
    peer_st_release((unsigned long long*)(P.base[threadIdx.x] + off_released) + P.rank, value);
}

static PeerPtrs peer_ptrs(const mdb_dist* s) {
  PeerPtrs P;
  memset(&P, 0, sizeof(P));
  P.world = s->world; P.rank = s->rank;
  for (int q = 0; q < s->world; ++q) P.base[q] = s->peer_base[q];
  return P;
}

// peer_base: host array of `world` device pointers, the symmetric buffer of every rank as mapped into this process
// (peer_base[rank] = this rank's own).  The buffers must be zero-filled and all ranks must have attached before the
// first mdb_dist_set_bounds.  From here on the operand buffers and the state vectors live in the symmetric buffer.
// Stores are unicast (one copy per peer on the owner's egress links).  An NVLS multicast variant would have to use
// multimem.st with its own fence (plain stores to a multicast address are undefined); it is not implemented.
extern "C" int mdb_dist_attach_peers(mdb_dist* s, double* const* peer_base) {
  if (!s || !peer_base) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, s->world <= MDB_PEER_MAX, "mdb_dist_attach_peers: more ranks than MDB_PEER_MAX");
  MDB_CHECK(ctx, !s->have_bounds, "mdb_dist_attach_peers: attach before mdb_dist_set_bounds");
  for (int q = 0; q < s->world; ++q) MDB_CHECK(ctx, peer_base[q] != nullptr, "mdb_dist_attach_peers: null peer pointer");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const mdb_peer_layout L = peer_layout(s);
  double* own = peer_base[s->rank];
  // gamma was extracted into the old vec by set_matrix / generate: carry the vectors over
  MDB_CUDA(ctx, cudaMemcpyAsync(own + L.vec, s->vec, (size_t)7 * s->n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 2; ++i) { cudaFree(s->OPX[i]); cudaFree(s->OPY[i]); }
  cudaFree(s->vec); cudaFree(s->clamped);
  for (int i = 0; i < 2; ++i) { s->OPX[i] = own + L.opx[i]; s->OPY[i] = own + L.opy[i]; }
  s->vec = own + L.vec;
  s->clamped = (uint8_t*)(own + L.clamped);
  const int64_t n = s->n;
  s->gamma = s->vec; s->alpha = s->vec + n; s->beta = s->vec + 2 * n; s->rowmax = s->vec + 3 * n;
  s->d = s->vec + 4 * n; s->omega = s->vec + 5 * n; s->delta = s->vec + 6 * n;
  for (int q = 0; q < s->world; ++q) s->peer_base[q] = peer_base[q];
  if (!s->push_counter) {
    MDB_CUDA(ctx, dev_malloc(ctx, (void**)&s->push_counter, sizeof(unsigned int)));
    MDB_CUDA(ctx, cudaMemsetAsync(s->push_counter, 0, sizeof(unsigned int), ctx->stream));
    MDB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  s->p2p = true;
  return MDB_OK;
}

// Owner of panel k (after mdb_dist_factor_panel on the same stream): stores the panel into every peer and raises
// arrive[k] everywhere.  outer_index = position of the panel's outer block in the sequence of outer blocks.
extern "C" int mdb_dist_push_panel(mdb_dist* s, int64_t k, int slot, int64_t k_outer, int oslot, int64_t outer_index) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, s->p2p, "mdb_dist_push_panel: mdb_dist_attach_peers first");
  MDB_CHECK(ctx, k >= 0 && k < s->nblk_glob && (k % s->world) == s->rank, "mdb_dist_push_panel: not the owner of panel k");
  MDB_CHECK(ctx, (slot == 0 || slot == 1) && (oslot == 0 || oslot == 1) && outer_index >= 0, "bad slot");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const mdb_peer_layout L = peer_layout(s);
  const int64_t n = s->n, k0 = k * NB, k1 = std::min<int64_t>(k0 + NB, n), rows = n - k1, ob = k_outer * NB;
  const int64_t off = (k1 - ob) * s->ldp + (k0 - ob);
  const unsigned long long tag = (unsigned long long)s->epoch << 32;
  const unsigned long long need = outer_index >= 2 ? (tag | (unsigned long long)(outer_index - 1)) : 0ull;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 1) / 2, (int64_t)ctx->num_sms * 4));
  k_dist_push<<<grid, 256, 0, ctx->stream>>>(rows, s->OPX[oslot] + off, s->ldp, s->panel[slot], s->alpha + k1, s->beta + k1,
                                             s->rowmax + k1, peer_ptrs(s), L.opx[oslot] + off, L.opy[oslot] + off,
                                             L.vec + n + k1, n, k0, (int)(k1 - k0), L.vec, L.clamped, L.arrive + k,
                                             L.released, need, tag | 1ull, s->push_counter, s->info);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// Everybody, on the stream that consumes panel k: wait until its owner's stores have landed here.
extern "C" int mdb_dist_wait_panel(mdb_dist* s, int64_t k) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, s->p2p && k >= 0 && k < s->nblk_glob, "mdb_dist_wait_panel: bad panel / not attached");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((k % s->world) == s->rank) return MDB_OK;  // the owner's own data is in place in stream order
  const mdb_peer_layout L = peer_layout(s);
  const unsigned long long* flag = (const unsigned long long*)(s->peer_base[s->rank] + L.arrive) + k;
  k_dist_wait<<<1, 32, 0, ctx->stream>>>(flag, ((unsigned long long)s->epoch << 32) | 1ull, s->info);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}

// On a stream that is ordered after every kernel of this rank that read the operands of the first `count` outer blocks.
extern "C" int mdb_dist_release_outer(mdb_dist* s, int64_t count) {
  if (!s) return MDB_ERR_INVALID;
  mdb_ctx* ctx = s->ctx;
  MDB_CHECK(ctx, s->p2p && count >= 0, "mdb_dist_release_outer: not attached");
  MDB_CUDA(ctx, cudaSetDevice(ctx->device));
  const mdb_peer_layout L = peer_layout(s);
  k_dist_release<<<1, 32, 0, ctx->stream>>>(peer_ptrs(s), L.released, ((unsigned long long)s->epoch << 32) | (unsigned long long)count);
  MDB_LAUNCHED(ctx);
  MDB_CUDA(ctx, cudaGetLastError());
  return MDB_OK;
}
