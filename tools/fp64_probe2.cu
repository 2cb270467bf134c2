// FP64 pipe micro-probe for B200 (debug tool): DFMA / DMMA.8x8x4 latency and throughput per SM as a function of the number
// of resident warps, MUFU.RSQ64H chain, LDS broadcast / distinct loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe2 tools/fp64_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void probe(double* out, long long* cyc, double seed, int mode) {
  __shared__ double sd[2048];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += blockDim.x) sd[i] = seed + i * 1e-9;
  __syncthreads();
  double y = seed * 0.5 + 1.0;
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  double r = 0;
  long long t0 = 0, t1 = 0;
  if (mode == 0) {   // DFMA throughput, 8 independent chains per thread
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
      a0 = fma(a0, 1.0000001, y); a1 = fma(a1, 1.0000001, y); a2 = fma(a2, 1.0000001, y); a3 = fma(a3, 1.0000001, y);
      a4 = fma(a4, 1.0000001, y); a5 = fma(a5, 1.0000001, y); a6 = fma(a6, 1.0000001, y); a7 = fma(a7, 1.0000001, y);
    }
    __syncthreads();
    t1 = clock64();
  } else if (mode == 1) {   // DMMA throughput, 4 independent accumulator tiles per warp
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
      dmma(a0, a1, y, seed); dmma(a2, a3, y, seed); dmma(a4, a5, y, seed); dmma(a6, a7, y, seed);
    }
    __syncthreads();
    t1 = clock64();
  } else if (mode == 2) {   // dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) { a0 = fma(a0, 1.0000001, y); a0 = fma(a0, 0.9999999, y); a0 = fma(a0, 1.0000001, y); a0 = fma(a0, 0.9999999, y); }
    t1 = clock64();
  } else if (mode == 3) {   // dependent DMMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) { dmma(a0, a1, y, seed); dmma(a0, a1, y, seed); dmma(a0, a1, y, seed); dmma(a0, a1, y, seed); }
    t1 = clock64();
  } else if (mode == 4) {   // rsqrt.approx + third-order correction chain
    double d = seed + 2.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
      double yy;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yy) : "d"(d));
      const double e = fma(d, -(yy * yy), 1.0);
      d = fma(fma(e, 0.375, 0.5), yy * e, yy) + 1.5;
    }
    t1 = clock64();
    a0 = d;
  } else if (mode == 5) {   // LDS.64 broadcast loads + DFMA (8 per iteration)
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
      const double* p = sd + ((i * 8) & 1023);
      a0 = fma(p[0], y, a0); a1 = fma(p[1], y, a1); a2 = fma(p[2], y, a2); a3 = fma(p[3], y, a3);
      a4 = fma(p[4], y, a4); a5 = fma(p[5], y, a5); a6 = fma(p[6], y, a6); a7 = fma(p[7], y, a7);
    }
    __syncthreads();
    t1 = clock64();
  } else if (mode == 6) {   // LDS.64 distinct addresses per lane + DFMA
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 128; i++) {
      const double* p = sd + ((i * 8 + (tid & 31) * 33) & 1023);
      a0 = fma(p[0], y, a0); a1 = fma(p[64], y, a1); a2 = fma(p[128], y, a2); a3 = fma(p[192], y, a3);
      a4 = fma(p[256], y, a4); a5 = fma(p[320], y, a5); a6 = fma(p[384], y, a6); a7 = fma(p[448], y, a7);
    }
    __syncthreads();
    t1 = clock64();
  }
  if (tid == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + r;
}

int main() {
  double* out; cudaMalloc(&out, 1024 * 8 * 4);
  long long* cyc; cudaMallocManaged(&cyc, 64);
  const char* names[] = {"DFMA tput (8 chains)", "DMMA tput (4 tiles)", "DFMA dependent", "DMMA dependent", "rsqrt chain", "LDS bcast+DFMA", "LDS distinct+DFMA"};
  for (int mode = 0; mode < 7; mode++) {
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
      if ((mode == 2 || mode == 3 || mode == 4) && threads > 32) continue;
      for (int rep = 0; rep < 2; rep++) { probe<<<1, threads>>>(out, cyc, 1.0, mode); cudaDeviceSynchronize(); }
      const double c = (double)cyc[0];
      if (mode == 0 || mode == 5 || mode == 6) printf("%-22s threads=%4d: %8.0f cycles, %.2f FMA/clk/SM, %.2f cycles per warp-instruction per SM\n", names[mode], threads, c, 128.0 * 8 * threads / c, c / (128.0 * 8 * threads / 32));
      else if (mode == 1) printf("%-22s threads=%4d: %8.0f cycles, %.2f FMA/clk/SM, %.2f cycles per DMMA per SM\n", names[mode], threads, c, 128.0 * 4 * 256 * (threads / 32) / c, c / (128.0 * 4 * threads / 32));
      else printf("%-22s: %.1f cycles per op\n", names[mode], c / (128.0 * (mode == 4 ? 1 : 4)));
    }
  }
  return 0;
}
