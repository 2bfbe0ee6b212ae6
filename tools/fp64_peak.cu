// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) register-resident
// throughput.  The result is the denominator of the "fraction of FP64 tensor-core peak" metric
// (MEASURED_PEAKS.json holds only HBM and bf16 peaks).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
// Run on the GPU box: tools/fp64_peak [seconds_per_test]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
               "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

constexpr int ILP = 8;

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double s) {
  double acc[ILP];
  double a = 1.0 + 1e-9 * threadIdx.x, b = s;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += acc[i];
  if (r == 12345.678) out[0] = r;
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double s) {
  double c[ILP][2];
  double a = 1e-3 * (threadIdx.x & 31), b = s;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1];
  if (r == 12345.678) out[0] = r;
}

template <int SHAPE>  // 4, 8, 16 = K of m16n8kK
__global__ void __launch_bounds__(256) k_dmma168(double* out, int iters, double s) {
  double c[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1e-3 * ((threadIdx.x & 31) + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = s + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 3 * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (SHAPE == 4) { double a2[2] = {a[0], a[1]}; dmma1684(c[i], a2, b[0]); }
        if (SHAPE == 8) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; dmma1688(c[i], a4, b2); }
        if (SHAPE == 16) dmma16816(c[i], a, b);
      }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (r == 12345.678) out[0] = r;
}

// DMMA m16n8k4 and DFMA interleaved: do the two pipes add up?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double s) {
  double c[ILP][4];
  double f[ILP];
  double a[2] = {1e-3 * (threadIdx.x & 31), 2e-3}, b = s;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 3 * i; f[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        dmma1684(c[i], a, b);
        f[i] = fma(f[i], a[0], b);
      }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3] + f[i];
  if (r == 12345.678) out[0] = r;
}

template <typename F>
static void run(const char* name, F launch, double flop_per_thread_iter, int threads, int blocks, double seconds) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int iters = 2000;
  launch(iters);  // warm-up
  CK(cudaDeviceSynchronize());
  // calibrate
  CK(cudaEventRecord(e0)); launch(iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  int reps = (int)(seconds * 1000.0 / ms) + 1;
  double best = 0, tot_ms = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = flop_per_thread_iter * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    tot_ms += ms;
  }
  double sustained = flop_per_thread_iter * iters * (double)threads * blocks * reps / (tot_ms * 1e-3) / 1e12;
  printf("{\"test\": \"%s\", \"threads\": %d, \"blocks\": %d, \"reps\": %d, \"best_tflops\": %.3f, \"sustained_tflops\": %.3f}\n",
         name, threads, blocks, reps, best, sustained);
  fflush(stdout);
}

int main(int argc, char** argv) {
  double seconds = argc > 1 ? atof(argv[1]) : 1.0;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount,
         prop.major, prop.minor, prop.clockRate);
  double* out; CK(cudaMalloc(&out, 8));
  const int T = 256;
  for (int bps : {1, 2, 4, 8}) {
    int blocks = prop.multiProcessorCount * bps;
    char nm[64];
    snprintf(nm, 64, "dfma_x%d", bps);
    run(nm, [&](int it) { k_dfma<<<blocks, T>>>(out, it, 1e-300); }, 2.0 * ILP * 4, T, blocks, seconds);
    // per warp instruction flop / 32 threads
    snprintf(nm, 64, "dmma_m8n8k4_x%d", bps);
    run(nm, [&](int it) { k_dmma884<<<blocks, T>>>(out, it, 1e-300); }, 2.0 * 8 * 8 * 4 / 32 * ILP * 4, T, blocks, seconds);
    snprintf(nm, 64, "dmma_m16n8k4_x%d", bps);
    run(nm, [&](int it) { k_dmma168<4><<<blocks, T>>>(out, it, 1e-300); }, 2.0 * 16 * 8 * 4 / 32 * ILP * 4, T, blocks, seconds);
    snprintf(nm, 64, "dmma_m16n8k8_x%d", bps);
    run(nm, [&](int it) { k_dmma168<8><<<blocks, T>>>(out, it, 1e-300); }, 2.0 * 16 * 8 * 8 / 32 * ILP * 4, T, blocks, seconds);
    snprintf(nm, 64, "dmma_m16n8k16_x%d", bps);
    run(nm, [&](int it) { k_dmma168<16><<<blocks, T>>>(out, it, 1e-300); }, 2.0 * 16 * 8 * 16 / 32 * ILP * 4, T, blocks, seconds);
    snprintf(nm, 64, "mixed_m16n8k4_dfma_x%d", bps);
    run(nm, [&](int it) { k_mixed<<<blocks, T>>>(out, it, 1e-300); }, (2.0 * 16 * 8 * 4 / 32 + 2.0) * ILP * 4, T, blocks, seconds);
  }
  // long sustained run of the best shape for the clock record
  {
    int blocks = prop.multiProcessorCount * 4;
    run("dmma_m16n8k16_x4_long", [&](int it) { k_dmma168<16><<<blocks, T>>>(out, it, 1e-300); },
        2.0 * 16 * 8 * 16 / 32 * ILP * 4, T, blocks, 4.0);
    run("dfma_x4_long", [&](int it) { k_dfma<<<blocks, T>>>(out, it, 1e-300); }, 2.0 * ILP * 4, T, blocks, 4.0);
  }
  return 0;
}
