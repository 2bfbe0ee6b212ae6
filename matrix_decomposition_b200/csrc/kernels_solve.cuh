// kernels_solve.cuh -- device kernels behind LDL / LL solve and multiply with a resident factor
// (decompositions.py:961-991, :1324-1352; dense/util.py:24-29).
//
// Storage: after finalisation W holds L (lower, with its diagonal); k_symmetrize mirrors it into the
// upper triangle so that BOTH sweeps stream row-contiguous data: the forward sweep reads rows of L
// (lower triangle), the backward sweep reads rows of U = L^T (upper triangle).  Each sweep therefore
// reads n^2/2 doubles once: 8 n^2 bytes per solve in total, the algorithmic minimum.
// Right-hand sides are kept transposed on the device (Bt: nrhs x n, one contiguous row per right-hand
// side), which makes every block update a K-major "TN" product: Bt[:, I] -= Y[:, J] L[I, J]^T.
// Diagonal blocks are applied through their explicit inverses (computed once per factor), so a block
// step is two products and no sequential substitution.
#pragma once
#include "mdb_common.cuh"

// W[i, j] = W[j, i] for i < j  (tile transpose through shared memory, coalesced both ways)
__global__ void __launch_bounds__(256) k_symmetrize(int64_t n, double* __restrict__ W, int64_t ldw) {
  __shared__ double tile[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;  // tile (bi, bj) of the LOWER triangle, bj <= bi
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? W[i * ldw + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // destination element (row = bj*32 + r, col = bi*32 + tx) = source (bi*32 + tx, bj*32 + r)
    const int64_t i = bj * 32 + r, j = bi * 32 + tx;
    if (i < n && j < n && i < j) W[i * ldw + j] = tile[tx][r];
  }
}

// undo k_symmetrize: W[i, j] = 0 for i < j (before the factor is copied out again)
__global__ void __launch_bounds__(256) k_zero_upper(int64_t n, double* __restrict__ W, int64_t ldw) {
  const int64_t i = blockIdx.x;
  const int64_t j0 = (int64_t)blockIdx.y * 1024 + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t j = j0 + 256 * e;
    if (j < n && j > i) W[i * ldw + j] = 0.0;
  }
}

// Per diagonal block J (NB x NB, lower triangular with diagonal; unit_diag forces a unit diagonal):
//   Linv[J]  = inverse (lower),  LinvT[J] = its transpose (= inverse of the upper block U_JJ = L_JJ^T)
//   Ltri[J]  = the lower triangle incl. diagonal, zeros above;  Utri[J] = its transpose
// all row-major NB x NB.  Blocks past n are padded with the identity.  One CTA per block; thread c
// owns column c of the inverse (forward substitution against e_c).  The block is kept packed
// (r (r + 1) / 2 + c) so that block + inverse fit in shared memory.
template <int NB>
__global__ void __launch_bounds__(NB) k_prepare_blocks(int64_t n, const double* __restrict__ W, int64_t ldw,
                                                       int unit_diag, double* __restrict__ Linv,
                                                       double* __restrict__ LinvT, double* __restrict__ Ltri,
                                                       double* __restrict__ Utri) {
  extern __shared__ double smem_d[];
  constexpr int LD = NB + 1;
  double* Lp = smem_d;                      // packed lower triangle, NB (NB + 1) / 2
  double* Xs = Lp + NB * (NB + 1) / 2;      // Xs[r * LD + c] = inverse
  const int64_t J = blockIdx.x, j0 = J * NB;
  const int tid = threadIdx.x;
  for (int r = 0; r < NB; ++r) {
    if (tid <= r) {
      const int64_t i = j0 + r, j = j0 + tid;
      double v = (r == tid) ? 1.0 : 0.0;
      if (i < n && j < n && !(unit_diag && r == tid)) v = W[i * ldw + j];
      Lp[r * (r + 1) / 2 + tid] = v;
    }
  }
  __syncthreads();
  {
    const int c = tid;
    for (int r = 0; r < NB; ++r) {
      if (r < c) {
        Xs[r * LD + c] = 0.0;
        continue;
      }
      double s = (r == c) ? 1.0 : 0.0;
      const double* Lr = Lp + r * (r + 1) / 2;
      for (int m = c; m < r; ++m) s -= Lr[m] * Xs[m * LD + c];
      Xs[r * LD + c] = s / Lr[r];
    }
  }
  __syncthreads();
  double* oLinv = Linv + J * NB * NB;
  double* oLinvT = LinvT + J * NB * NB;
  double* oLtri = Ltri + J * NB * NB;
  double* oUtri = Utri + J * NB * NB;
  for (int r = 0; r < NB; ++r) {
    const double l_rt = (tid <= r) ? Lp[r * (r + 1) / 2 + tid] : 0.0;   // L[r][tid]
    const double l_tr = (r <= tid) ? Lp[tid * (tid + 1) / 2 + r] : 0.0;  // L[tid][r]
    oLinv[r * NB + tid] = Xs[r * LD + tid];
    oLinvT[r * NB + tid] = Xs[tid * LD + r];
    oLtri[r * NB + tid] = l_rt;
    oUtri[r * NB + tid] = l_tr;
  }
}

// Operands of block column J of the composed matrix (mdb_factor_composed): rows [j0, j0 + NB) of L over the columns
// [0, j0 + NB), clean (the diagonal block comes from Ltri[J]: zeros above the diagonal, identity padding past n; what
// lies left of it in W is the strictly lower part anyway), once as they are (AJ) and once scaled by d (BJ = AJ D).
// Row pitch ldt.  grid = (ceil(K / 256), NB).
__global__ void __launch_bounds__(256) k_composed_operands(int64_t n, const double* __restrict__ W, int64_t ldw,
                                                           const double* __restrict__ LtriJ, const double* __restrict__ d,
                                                           int use_d, int64_t j0, int nb, double* __restrict__ AJ,
                                                           double* __restrict__ BJ, int64_t ldt) {
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= j0 + nb) return;
  double a;
  if (k < j0) a = (j0 + c < n) ? W[(j0 + c) * ldw + k] : 0.0;
  else a = LtriJ[c * nb + (k - j0)];
  const double dk = use_d ? (k < n ? d[k] : 0.0) : 1.0;
  AJ[c * ldt + k] = a;
  BJ[c * ldt + k] = a * dk;
}

// Small-M "TN" product for up to 8 right-hand sides (bandwidth bound, one warp per B row):
//   C[r, i] = beta C[r, i] + sign * sum_{k < K} A[r, k] B[i, k],  r < M <= 8, i < N, K multiple of 128.
// A rows live in registers (each lane holds 4 consecutive doubles per 128-wide chunk), B rows are
// streamed once with 32-byte per-lane loads: 8 N K bytes of traffic, the algorithmic minimum.
template <int MMAX>
__global__ void __launch_bounds__(256) k_gemv_tn(int M, int64_t N, int64_t K, const double* __restrict__ A, int64_t lda,
                                                 const double* __restrict__ B, int64_t ldb, double* __restrict__ C,
                                                 int64_t ldc, double beta, double sign) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= N) return;
  double acc[MMAX];
#pragma unroll
  for (int r = 0; r < MMAX; ++r) acc[r] = 0.0;
  const double* Brow = B + i * ldb;
  for (int64_t k0 = 0; k0 < K; k0 += 128) {
    const double2* bp = reinterpret_cast<const double2*>(Brow + k0 + lane * 4);
    const double2 b0 = bp[0], b1 = bp[1];
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      if (r < M) {
        const double2* ap = reinterpret_cast<const double2*>(A + r * lda + k0 + lane * 4);
        const double2 a0 = __ldg(ap), a1 = __ldg(ap + 1);
        acc[r] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < MMAX; ++r) {
    double v = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && r < M) {
      double* c = C + r * ldc + i;
      *c = (beta != 0.0 ? beta * (*c) : 0.0) + sign * v;
    }
  }
}

// Bt[r, i] = B[p[i] * ldb + r]   (gather + transpose of the right-hand sides: x = x[p], decompositions.py:986)
__global__ void __launch_bounds__(256) k_rhs_gather_t(int64_t n, int64_t nrhs, const double* __restrict__ B,
                                                      int64_t ldb, const int64_t* __restrict__ p,
                                                      double* __restrict__ Bt, int64_t ldbt) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int a = ty; a < 32; a += 8) {  // rows i0 + a of B (after permutation), columns r0 + tx
    const int64_t i = i0 + a, r = r0 + tx;
    tile[a][tx] = (i < n && r < nrhs) ? B[(p ? p[i] : i) * ldb + r] : 0.0;
  }
  __syncthreads();
  for (int a = ty; a < 32; a += 8) {
    const int64_t r = r0 + a, i = i0 + tx;
    if (r < nrhs && i < n) Bt[r * ldbt + i] = tile[tx][a];
  }
}

// X[p[i] * ldx + r] = Xt[r, i]   (x = x[p_inverse], decompositions.py:990)
__global__ void __launch_bounds__(256) k_rhs_scatter_t(int64_t n, int64_t nrhs, const double* __restrict__ Xt,
                                                       int64_t ldxt, const int64_t* __restrict__ p,
                                                       double* __restrict__ X, int64_t ldx) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int a = ty; a < 32; a += 8) {
    const int64_t r = r0 + a, i = i0 + tx;
    tile[a][tx] = (r < nrhs && i < n) ? Xt[r * ldxt + i] : 0.0;
  }
  __syncthreads();
  for (int a = ty; a < 32; a += 8) {
    const int64_t i = i0 + a, r = r0 + tx;
    if (i < n && r < nrhs) X[(p ? p[i] : i) * ldx + r] = tile[tx][a];
  }
}

// Yt[r, i] = Yt[r, i] / d[i]  (mode 0, decompositions.py:988)  or  * d[i]  (mode 1, :965)
__global__ void __launch_bounds__(256) k_scale_by_d(int64_t n, int64_t nrhs, double* __restrict__ Yt, int64_t ld,
                                                    const double* __restrict__ d, int mode) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (i < n && r < nrhs) {
    const double v = Yt[r * ld + i];
    Yt[r * ld + i] = mode == 0 ? v / d[i] : v * d[i];
  }
}

// y_i += (A o Omega + diag(gamma + delta)) v evaluated from the UPPER triangle of the un-finalised W
// (which still holds the permuted A): position order, Omega[i, j] = omega[max(i, j)]  (Reimer.py:947-957).
// One CTA per slab of 32 rows; the symmetric contribution of element (i, j), j > i, goes to y_i (row
// dot, warp reduce) and to y_j (per-thread column accumulators, one atomic per column per slab).
__global__ void __launch_bounds__(256) k_identity_rhs(int64_t n, const double* __restrict__ W, int64_t ldw,
                                                      const double* __restrict__ gamma, const double* __restrict__ delta,
                                                      const double* __restrict__ omega, const double* __restrict__ v,
                                                      double* __restrict__ y_b, double* __restrict__ y_a) {
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int64_t jc0 = (int64_t)blockIdx.y * 1024;  // column chunk
  const int tid = threadIdx.x;
  if (jc0 + 1024 <= i0) return;  // every column of this chunk is left of the slab: nothing above the diagonal
  double colacc_b[4] = {0, 0, 0, 0}, colacc_a[4] = {0, 0, 0, 0};
  for (int r = 0; r < 32; ++r) {
    const int64_t i = i0 + r;
    if (i >= n) break;
    const double vi = v[i];
    double rowsum_b = 0, rowsum_a = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t j = jc0 + tid + 256 * e;
      if (j < n && j > i) {
        const double a = W[i * ldw + j];
        const double wj = omega[j];
        rowsum_b += a * wj * v[j];
        rowsum_a += a * v[j];
        colacc_b[e] += a * wj * vi;
        colacc_a[e] += a * vi;
      }
    }
    // block reduce of the two row sums
    double s1 = rowsum_b, s2 = rowsum_a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&y_b[i], s1);
      atomicAdd(&y_a[i], s2);
    }
    if (tid == 0 && i >= jc0 && i < jc0 + 1024) {  // the chunk that holds the diagonal element (never skipped)
      atomicAdd(&y_b[i], (gamma[i] + delta[i]) * vi);
      atomicAdd(&y_a[i], gamma[i] * vi);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t j = jc0 + tid + 256 * e;
    if (j < n && (colacc_b[e] != 0.0 || colacc_a[e] != 0.0)) {
      atomicAdd(&y_b[j], colacc_b[e]);
      atomicAdd(&y_a[j], colacc_a[e]);
    }
  }
}

__global__ void k_fill_random(int64_t n, uint64_t seed, double* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint64_t h = mdb_splitmix64(seed ^ mdb_splitmix64((uint64_t)i));
    v[i] = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
  }
}

// out[0] = max_i |a_i - b_i|, out[1] = max_i |c_i|   (non-negative doubles compare like their bit patterns)
__global__ void k_max_abs_diff(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                               const double* __restrict__ c, unsigned long long* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double e = fabs(a[i] - b[i]);
    const double m = fabs(c[i]);
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(isnan(e) ? INFINITY : e));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(m));
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent triangular sweep for up to 8 right-hand sides (the bandwidth-bound solve, decompositions.py:987-989):
//   forward  (reverse = 0):  z_I = Linv_I  (b_I - sum_{J < I} L[I, J] z_J)        L  = lower triangle of W
//   backward (reverse = 1):  x_I = LinvT_I (y_I - sum_{J > I} U[I, J] x_J)        U  = L^T = mirrored upper triangle
// in ONE launch per sweep instead of two launches per 128-row block.  L is streamed exactly once (8 n^2 bytes per
// solve, both sweeps): the work is cut into tasks that CTAs claim from a global counter in dependency order,
//   partial (I, c) : p[I][c] = sum over the column blocks J of chunk c (8 blocks, J < I - 1) of L[I, J] z_J
//   diag I         : z_I = Linv_I (b_I - sum_c p[I][c] - L[I, I-1] z_{I-1})
// Each block L[I, J] (128 x 128) is loaded into REGISTERS before the task waits for z_J, so after z_J arrives only
// the arithmetic is left; the diag task also holds Linv_I (transposed) in shared memory ahead of time.
// Synchronisation is on the DATA: the output vector and the partial sums are pre-filled with a sentinel (a NaN with
// a private payload) and a consumer polls the very doubles it needs until they are no longer the sentinel -- no
// flags, no fences, one L2 round trip per step of the dependency chain z_0 -> z_1 -> ...  (a right-hand side that
// contains exactly that NaN pattern runs into the bounded spin and is reported as an error).  A task only waits
// for tasks earlier in the claim order, so resident CTAs always make progress (no co-residency assumption).
// Partial sums are combined in a fixed order: results are deterministic.  In reversed mode block index I stands
// for T - 1 - I.
// ------------------------------------------------------------------------------------------------
struct TrsvArgs {
  const double* W; int64_t ldw; int64_t n; int T;
  const double* DinvT;    // per block: TRANSPOSE of the inverse diagonal block to apply (forward: LinvT, backward: Linv)
  const double* in;       // nrhs x ldt
  double* out;            // nrhs x ldt, pre-filled with the sentinel
  int64_t ldt; int nrhs;
  const double* scale_d;  // backward sweep of an LDL factor: the input is divided by d first (decompositions.py:988)
  double* partial;        // [T][NC][nrhs][128], pre-filled with the sentinel
  int NC;
  const int2* tasks;      // partial tasks (I, c) in claim order
  int ntasks;
  int nd;                 // CTAs dedicated to the diag tasks
  unsigned int* next_task;
  int reverse;
  int* err;
};

namespace trsv {
constexpr int NB = MDB_NB, CH = 8, THREADS = 256;
constexpr unsigned int SPIN_LIMIT = 1u << 26;
constexpr unsigned long long SENTINEL = 0x7ff8dead5eed0001ull;

// polls one double until the producer has overwritten the sentinel
__device__ __forceinline__ double poll(const double* p, int* err, bool* ok) {
  unsigned long long bits;
  unsigned int it = 0;
  for (;;) {
    asm volatile("ld.volatile.global.u64 %0, [%1];\n" : "=l"(bits) : "l"(p) : "memory");
    if (bits != SENTINEL) break;
    if (++it > SPIN_LIMIT) { atomicExch(err, 1); *ok = false; break; }
    if (it > 256) __nanosleep(32);
  }
  return __longlong_as_double((long long)bits);
}
__device__ __forceinline__ void publish(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}
}  // namespace trsv

__global__ void k_fill_sentinel(int64_t n, double* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<unsigned long long*>(x)[i] = trsv::SENTINEL;
}

template <int MR>
__global__ void __launch_bounds__(trsv::THREADS, 1) k_trsv_persistent(TrsvArgs a) {
  using namespace trsv;
  extern __shared__ double smem_d[];
  double* sD = smem_d;              // NB x NB: DinvT of the current diag task, sD[c * NB + r]
  double* zs = sD + NB * NB;        // MR x NB: z of the block being applied
  double* vs = zs + MR * NB;        // MR x NB: right-hand side of the diag task
  __shared__ int s_task;
  __shared__ int s_fail;
  const int tid = threadIdx.x, r = tid >> 1, h = tid & 1;
  const int nrhs = a.nrhs;
  const int T = a.T;
  if (tid == 0) s_fail = 0;

  // element (row block I, column block J) of the triangle being swept, as global (row, col) of W
  auto blk_row = [&](int I) { return (int64_t)(a.reverse ? T - 1 - I : I) * NB; };

  // The first ND CTAs own the dependency chain: CTA b runs the diag tasks I = b, b + ND, ... and nothing else, so
  // a diag task never queues behind a long partial task and the next one is already prefetching (its L block and
  // inverse diagonal block) while the current one computes.  All other CTAs claim partial tasks in order.
  const bool diag_cta = (int)blockIdx.x < a.nd;
  int my_diag = (int)blockIdx.x;
  for (;;) {
    int I, c;
    if (diag_cta) {
      __syncthreads();
      if (my_diag >= T || s_fail != 0) break;
      I = my_diag; c = -1;
      my_diag += a.nd;
    } else {
      if (tid == 0) s_task = (int)atomicAdd(a.next_task, 1u);
      __syncthreads();
      const int t = s_task;
      const bool failed = s_fail != 0;
      __syncthreads();
      if (t >= a.ntasks || failed) break;
      I = a.tasks[t].x; c = a.tasks[t].y;
    }
    const int64_t row0 = blk_row(I);
    const int64_t grow = row0 + r;
    const bool rowok = grow < a.n;
    const double* Wrow = a.W + (rowok ? grow : 0) * a.ldw;
    double blk[64];
    double acc[MR];

    // registers <- block (I, J): thread (r, h) holds the 16-byte chunks q = 2 i + h of row r
    auto load_block = [&](int J) {
      const double* src = Wrow + blk_row(J);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double2 v = make_double2(0.0, 0.0);
        if (rowok) v = *reinterpret_cast<const double2*>(src + 4 * i + 2 * h);
        blk[2 * i] = v.x;
        blk[2 * i + 1] = v.y;
      }
    };
    // zs <- z_J (all right-hand sides), polling the data itself
    auto load_z = [&](int J) {
      const int64_t c0 = blk_row(J);
      bool ok = true;
      for (int e = tid; e < nrhs * NB; e += THREADS)
        zs[e] = poll(a.out + (int64_t)(e / NB) * a.ldt + c0 + (e % NB), a.err, &ok);
      if (!ok) s_fail = 1;
    };
    auto apply_block = [&]() {  // acc[m] += sum over this thread's chunks
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        if (m < nrhs) {
          double s0 = 0, s1 = 0;
          const double* z = zs + m * NB + 2 * h;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const double2 zz = *reinterpret_cast<const double2*>(z + 4 * i);
            s0 = fma(blk[2 * i], zz.x, s0);
            s1 = fma(blk[2 * i + 1], zz.y, s1);
          }
          acc[m] += s0 + s1;
        }
      }
    };
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[m] = 0.0;

    if (c >= 0) {
      // ---- partial task: column blocks [c CH, min(c CH + CH, I - 1)) ----
      const int J0 = c * CH, J1 = min(J0 + CH, I - 1);
      for (int J = J0; J < J1; ++J) {
        load_block(J);
        load_z(J);
        __syncthreads();
        apply_block();
        __syncthreads();
      }
      double* out = a.partial + ((int64_t)I * a.NC + c) * nrhs * NB;
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        if (m < nrhs) {
          const double v = acc[m] + __shfl_xor_sync(0xffffffffu, acc[m], 1);
          if (h == 0) publish(out + m * NB + r, v);
        }
      }
    } else {
      // ---- diag task ----
      const int nparts = (I - 1 > 0) ? (I - 1 + CH - 1) / CH : 0;
      {  // the inverse diagonal block (already transposed in memory) -> shared, 16-byte copies
        const double* src = a.DinvT + (int64_t)(a.reverse ? T - 1 - I : I) * NB * NB;
        for (int e = tid; e < NB * NB / 2; e += THREADS)
          reinterpret_cast<double2*>(sD)[e] = reinterpret_cast<const double2*>(src)[e];
      }
      if (I >= 1) load_block(I - 1);
      // vs = b_I - sum_c partial[I][c]   (fixed order; each partial is polled until it has been produced)
      {
        bool ok = true;
        for (int e = tid; e < nrhs * NB; e += THREADS) {
          const int m = e / NB, rr = e % NB;
          const int64_t gi = row0 + rr;
          double v = a.in[(int64_t)m * a.ldt + gi];
          if (a.scale_d) v = (gi < a.n) ? v / a.scale_d[gi] : 0.0;
          const double* ps = a.partial + (int64_t)I * a.NC * nrhs * NB + m * NB + rr;
          for (int cc = 0; cc < nparts; ++cc) v -= poll(ps + (int64_t)cc * nrhs * NB, a.err, &ok);
          vs[e] = v;
        }
        if (!ok) s_fail = 1;
      }
      if (I >= 1) {
        load_z(I - 1);
        __syncthreads();
        apply_block();
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          if (m < nrhs) {
            const double v = acc[m] + __shfl_xor_sync(0xffffffffu, acc[m], 1);
            if (h == 0) vs[m * NB + r] -= v;
          }
        }
      }
      __syncthreads();
      // z_I = Dinv v: thread (r, h) sums the columns cc = 4 i + 2 h, 4 i + 2 h + 1
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        if (m < nrhs) {
          double s0 = 0, s1 = 0;
          const double* v = vs + m * NB + 2 * h;
          const double* dcol = sD + (2 * h) * NB + r;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const double2 vv = *reinterpret_cast<const double2*>(v + 4 * i);
            s0 = fma(dcol[(4 * i) * NB], vv.x, s0);
            s1 = fma(dcol[(4 * i + 1) * NB], vv.y, s1);
          }
          double z = s0 + s1;
          z += __shfl_xor_sync(0xffffffffu, z, 1);
          if (h == 0) publish(a.out + (int64_t)m * a.ldt + row0 + r, z);
        }
      }
    }
  }
}
