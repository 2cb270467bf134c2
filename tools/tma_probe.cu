// tma_probe.cu -- standalone probe of cp.async.bulk.tensor variants on u8 images (debug tool, not product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const CUtensorMap* map, int x0, int y0, int z0, int bytes, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t tile[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                   "r"(smem_u32(tile)), "l"(map), "r"(x0), "r"(y0), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                   "r"(smem_u32(tile)), "l"(map), "r"(x0), "r"(y0), "r"(z0), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0; int spins = 0;
  while (!done && spins < (1 << 22)) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    spins++;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? tile[i] : 0xEE;
}

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  PFN_encodeTiled encode = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  const int W = 640, H = 480, B = 2, pitch = 640;
  size_t fs = (size_t)pitch * H;
  std::vector<uint8_t> img(fs * B);
  for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 7 + (i / pitch) * 13) & 0xff);
  uint8_t *d_img, *d_out; CUtensorMap* d_map;
  cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 1 << 16); cudaMalloc(&d_map, sizeof(CUtensorMap));
  struct Case { int rank, bw, bh, x0, y0, z0; const char* name; };
  Case cases[] = {{2, 64, 32, 0, 0, 0, "2d box64 aligned"}, {2, 48, 38, 16, 16, 0, "2d box48 x16"}, {2, 48, 38, 47, 16, 0, "2d box48 x47"},
                  {3, 64, 32, 0, 0, 1, "3d box64 aligned z1"}, {3, 48, 38, 47, 16, 1, "3d box48 x47 z1"}, {3, 48, 38, 202, 16, 0, "3d box48 x202"},
                  {2, 48, 38, 620, 460, 0, "2d box48 oob corner"}, {2, 16, 8, 5, 5, 0, "2d box16x8 x5"}, {2, 32, 38, 47, 16, 0, "2d box32 x47"}};
  int ci = -1;
  for (auto& c : cases) {
    ci++;
    if (only >= 0 && ci != only) continue;
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
    cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.rank, d_img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemcpy(d_map, &m, sizeof m, cudaMemcpyHostToDevice);
    int bytes = c.bw * c.bh;
    cudaMemset(d_out, 0, 1 << 16);
    if (c.rank == 2) probe<2><<<1, 128, bytes + 128>>>(d_map, c.x0, c.y0, c.z0, bytes, d_out);
    else probe<3><<<1, 128, bytes + 128>>>(d_map, c.x0, c.y0, c.z0, bytes, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint8_t> out(bytes);
    cudaMemcpy(out.data(), d_out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < c.bh; y++)
      for (int x = 0; x < c.bw; x++) {
        int gx = c.x0 + x, gy = c.y0 + y;
        uint8_t ref = (gx < W && gy < H) ? img[c.z0 * fs + (size_t)gy * pitch + gx] : 0;
        if (out[y * c.bw + x] != ref) bad++;
      }
    printf("%-24s encode=%d launch=%s mismatches=%d first=%02x\n", c.name, (int)r, cudaGetErrorString(e), bad, out[0]);
    if (e != cudaSuccess) { printf("sticky error, stopping\n"); break; }
  }
  return 0;
}
