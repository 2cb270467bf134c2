// Stand-alone correctness + timing harness for ba_chol.h (debug tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/chol_probe2 tools/chol_probe2.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vido-slam_b200/csrc/ba_chol.h"

__global__ void __launch_bounds__(BC_THREADS, 1) probe_kernel(const double* A, const double* b, double* x, int n, long long* cyc, int reps, int* bad) {
  extern __shared__ double sm[];
  __shared__ int s_bad;
  __shared__ long long tk[16 + 8 * 18];
  CholSm cs;
  bc_carve(sm, n, cs);
  const int tid = threadIdx.x;
  if (tid < 16 + 8 * 18) tk[tid] = 0;
  long long tot = 0, tback = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = tid; i < (cs.np + 1) * cs.ld; i += blockDim.x) cs.S[i] = 0;
    __syncthreads();
    for (int i = tid; i < cs.np * cs.np; i += blockDim.x) {
      const int r = i / cs.np, c = i - r * cs.np;
      if (c <= r) cs.S[r * cs.ld + c] = (r < n && c < n) ? A[r * n + c] : (r == c ? 1.0 : 0.0);
    }
    for (int i = tid; i < cs.np; i += blockDim.x) cs.S[cs.np * cs.ld + i] = i < n ? b[i] : 0.0;
    __syncthreads();
    const long long t0 = clock64();
    bc_factor(cs, &s_bad, tk);
    __syncthreads();
    const long long t1 = clock64();
    if (!s_bad) bc_backsolve(cs);
    const long long t2b = clock64();
    __syncthreads();
    const long long t2 = clock64();
    tot += t2 - t0; tback += t2 - t1;
    if (tid == BC_CHAIN_WARP * 32) tk[8] += t2b - t1;
  }
  for (int i = tid; i < n; i += blockDim.x) x[i] = cs.S[cs.np * cs.ld + i];
  __syncthreads();
  if (tid == 0) { cyc[0] = tot / reps; cyc[1] = tback / reps; for (int k = 0; k < 9; k++) cyc[2 + k] = tk[k] / reps;
    for (int k = 16; k < 16 + 8 * 18; k++) cyc[k] = tk[k]; *bad = s_bad; }
}

int main() {
  for (int W : {20, 24, 5, 1, 12}) {
    const int n = 6 * W;
    std::vector<double> A((size_t)n * n), b(n), G((size_t)n * n);
    srand(7 + W);
    for (auto& g : G) g = (rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += G[(size_t)i * n + k] * G[(size_t)j * n + k];
        A[(size_t)i * n + j] = s + (i == j ? 1e-3 * n : 0.0);
      }
    for (int i = 0; i < n; i++) b[i] = rand() / (double)RAND_MAX - 0.5;
    // host reference
    std::vector<double> L(A), y(b), xr(n);
    for (int j = 0; j < n; j++) {
      for (int k = 0; k < j; k++) L[(size_t)j * n + j] -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
      L[(size_t)j * n + j] = sqrt(L[(size_t)j * n + j]);
      for (int i = j + 1; i < n; i++) {
        double s = L[(size_t)i * n + j];
        for (int k = 0; k < j; k++) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
        L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
      }
    }
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * y[k]; y[i] = s / L[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * xr[k]; xr[i] = s / L[(size_t)i * n + i]; }
    double *dA, *db, *dx; long long* dc; int* dbad;
    cudaMalloc(&dA, sizeof(double) * n * n); cudaMalloc(&db, sizeof(double) * n); cudaMalloc(&dx, sizeof(double) * n);
    cudaMalloc(&dc, sizeof(long long) * 160); cudaMalloc(&dbad, sizeof(int));
    cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), sizeof(double) * n, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(double) * bc_smem_doubles(n);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, BC_THREADS, smem>>>(dA, db, dx, n, dc, 20, dbad);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> x(n); long long c[160]; int bad = 0;
    cudaMemcpy(x.data(), dx, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, dc, sizeof c, cudaMemcpyDeviceToHost);
    cudaMemcpy(&bad, dbad, sizeof bad, cudaMemcpyDeviceToHost);
    double err = 0, nx = 0;
    for (int i = 0; i < n; i++) { err = fmax(err, fabs(x[i] - xr[i])); nx = fmax(nx, fabs(xr[i])); }
    printf("W=%d n=%d smem=%zu: %s bad=%d rel.err=%.3e | cycles total=%lld backsolve=%lld diag0=%lld chain_work=%lld chain_wait=%lld bulk_panel=%lld bulk_trail=%lld bulk_wait=%lld | chain: trsm=%lld diag_update=%lld backsolve(chain warp)=%lld\n",
           W, n, smem, cudaGetErrorString(e), bad, err / nx, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10]);
    if (W == 20) for (int jb = 0; jb < 15; jb++) printf("   step %2d: chain work=%lld wait=%lld | bulk panel=%lld (rows done at %lld, prefetch at %lld) panel-barrier=%lld trailing(incl. barrier)=%lld step-barrier=%lld\n", jb, c[16 + 8 * jb], c[17 + 8 * jb], c[18 + 8 * jb], c[22 + 8 * jb], c[23 + 8 * jb], c[19 + 8 * jb], c[20 + 8 * jb], c[21 + 8 * jb]);
    cudaFree(dA); cudaFree(db); cudaFree(dx); cudaFree(dc); cudaFree(dbad);
  }
  return 0;
}
