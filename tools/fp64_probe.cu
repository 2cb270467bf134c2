// Micro-probe of the latencies the window-BA kernel is bound by (debug tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void probe(double* out, long long* cyc, const int* chase_g, double seed) {
  __shared__ int chase_s[1024];
  __shared__ double sd[256];
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) chase_s[i] = (i * 37 + 11) & 1023;
  sd[tid & 255] = seed + tid;
  __syncthreads();
  long long t0, t1;
  double x = seed, y = seed * 0.5 + 1.0;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
    x = fma(x, 1.0000001, y); x = fma(x, 0.9999999, y); x = fma(x, 1.0000001, y); x = fma(x, 0.9999999, y);
  }
  t1 = clock64();
  if (tid == 0) cyc[0] = (t1 - t0) / 256;
  // 2. DFMA throughput: 8 independent chains per thread, all threads
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
    a0 = fma(a0, 1.0000001, y); a1 = fma(a1, 1.0000001, y); a2 = fma(a2, 1.0000001, y); a3 = fma(a3, 1.0000001, y);
    a4 = fma(a4, 1.0000001, y); a5 = fma(a5, 1.0000001, y); a6 = fma(a6, 1.0000001, y); a7 = fma(a7, 1.0000001, y);
  }
  __syncthreads();
  t1 = clock64();
  if (tid == 0) cyc[1] = (t1 - t0);  // 512 DFMA per thread, blockDim threads
  x = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  // 3. dependent rsqrt chain
  double r = fabs(x) + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) r = rsqrt(r) + 1.5;
  t1 = clock64();
  if (tid == 0) cyc[2] = (t1 - t0) / 64;
  // 4. dependent division chain
  double d = r + 1.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) d = 1.0 / d + 1.25;
  t1 = clock64();
  if (tid == 0) cyc[3] = (t1 - t0) / 64;
  // 4b. sqrt chain
  double q = d + 3.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) q = sqrt(q) + 2.5;
  t1 = clock64();
  if (tid == 0) cyc[9] = (t1 - t0) / 64;
  // 5. __syncthreads
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) __syncthreads();
  t1 = clock64();
  if (tid == 0) cyc[4] = (t1 - t0) / 64;
  // 6. cluster.sync
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) cl.sync();
  t1 = clock64();
  if (tid == 0) cyc[5] = (t1 - t0) / 64;
  // 7. shuffle chain (double = 2 shuffles)
  double sv = d;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) sv += __shfl_xor_sync(0xffffffffu, sv, 1 + (i & 15));
  t1 = clock64();
  if (tid == 0) cyc[6] = (t1 - t0) / 64;
  // 8. LDS pointer chase
  int p = tid & 1023;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) p = chase_s[p];
  t1 = clock64();
  if (tid == 0) cyc[7] = (t1 - t0) / 64;
  // 8b. LDS.64 + DFMA dependent
  double lv = sv;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) lv = fma(sd[(tid + i) & 255], 1.0000001, lv);
  t1 = clock64();
  if (tid == 0) cyc[10] = (t1 - t0) / 64;
  // 9. global (L2) pointer chase, written by the host -> L2 after first touch
  int g = (tid * 64 + blockIdx.x * 7) & ((1 << 18) - 1);
  for (int i = 0; i < 8; i++) g = __ldcg(chase_g + g);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) g = __ldcg(chase_g + g);
  t1 = clock64();
  if (tid == 0) cyc[8] = (t1 - t0) / 64;
  out[blockIdx.x * blockDim.x + tid] = x + r + d + sv + p + g + q + lv;
}

int main() {
  const int N = 1 << 18;
  int* h = new int[N];
  for (int i = 0; i < N; i++) h[i] = (int)(((long long)i * 40503 + 12345) & (N - 1));
  int* dch; cudaMalloc(&dch, N * 4); cudaMemcpy(dch, h, N * 4, cudaMemcpyHostToDevice);
  double* out; cudaMalloc(&out, 16 * 256 * 8);
  long long* cyc; cudaMallocManaged(&cyc, 16 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cluster : {16, 8, 1}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
      cudaError_t e = cudaLaunchKernelEx(&cfg, probe, out, cyc, (const int*)dch, 1.0);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cluster %d: %s\n", cluster, cudaGetErrorString(e)); break; }
    }
    printf("cluster=%d cycles: dfma_lat=%lld dfma_8x64x256thr_total=%lld (=> %.2f DFMA/clk/SM) rsqrt=%lld div=%lld sqrt=%lld syncthreads=%lld cluster_sync=%lld shfl64=%lld lds_chase=%lld lds64+dfma=%lld l2_chase=%lld\n",
           cluster, cyc[0], cyc[1], 512.0 * 256 / (double)cyc[1], cyc[2], cyc[3], cyc[9], cyc[4], cyc[5], cyc[6], cyc[7], cyc[10], cyc[8]);
  }
  return 0;
}
