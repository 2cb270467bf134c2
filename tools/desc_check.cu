// desc_check.cu -- stand-alone GPU check + timing of the descriptor stage through the C-ABI (no Python start-up: the whole run is
// a few seconds).  Compares vido_orb_extract_describe / vido_orb_get_blurred_level / vido_hamming_match(_dev) with the oracle
// (oracle/liboracle.so -- test infrastructure, used here as the checker only) on synthetic frames, then times the device-resident
// chain describe -> match at batch 16 with CUDA events on the context stream.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/desc_check tools/desc_check.cu \
//        -Lvido-slam_b200 -lvido_b200 -Loracle -loracle -Xlinker -rpath,'$ORIGIN/../vido-slam_b200' -Xlinker -rpath,'$ORIGIN/../oracle'
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/vido_b200.h"
#include "../oracle/vido_oracle.h"

extern "C" void* vido_stream(vido_ctx* ctx);

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

// blocks of random grey level with a little noise: plenty of FAST corners on every pyramid level
static void make_frame(uint8_t* img, int W, int H, int shift) {
  std::vector<uint8_t> blocks(((W + 200) / 9 + 2) * (H / 7 + 2));
  rng_state = 777;
  for (auto& b : blocks) b = (uint8_t)(rnd() % 200 + 20);
  const int bw = (W + 200) / 9 + 2;
  rng_state = 4242 + shift;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) img[(size_t)y * W + x] = (uint8_t)(blocks[(y / 7) * bw + (x + 3 * shift) / 9] + rnd() % 7);
}

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("\n"); fails++; } } while (0)
#define VCALL(call) do { int rc_ = (call); if (rc_ != VIDO_OK) { printf("FAIL: %s -> %d (%s)\n", #call, rc_, vido_last_error(ctx)); return 1; } } while (0)

static int check_size(int W, int H, int nframes, int max_batch) {
  vido_config cfg;
  vido_default_config(&cfg);
  cfg.width = W; cfg.height = H; cfg.max_batch = max_batch;
  vido_ctx* ctx = vido_create(&cfg);
  if (!ctx) { printf("FAIL: vido_create: %s\n", vido_last_error(nullptr)); return 1; }
  const int cap = cfg.nfeatures + 64;
  std::vector<uint8_t> imgs((size_t)nframes * W * H);
  for (int f = 0; f < nframes; f++) make_frame(imgs.data() + (size_t)f * W * H, W, H, f);
  std::vector<vido_keypoint> kp((size_t)nframes * cap);
  std::vector<int32_t> n(nframes);
  std::vector<uint8_t> desc((size_t)nframes * cap * 32, 0xEE);
  VCALL(vido_orb_extract_describe(ctx, imgs.data(), nframes, (size_t)W * H, W, kp.data(), cap, n.data(), desc.data()));
  vo_orb_params p = {cfg.nfeatures, cfg.scale_factor, cfg.nlevels, cfg.ini_th_fast, cfg.min_th_fast};
  std::vector<std::vector<uint8_t>> odesc(nframes);
  std::vector<int> on(nframes);
  for (int f = 0; f < nframes; f++) {
    std::vector<vo_keypoint> okp(20000);
    odesc[f].resize(20000 * 32);
    on[f] = vo_orb_extract_describe(imgs.data() + (size_t)f * W * H, W, H, W, &p, okp.data(), 20000, odesc[f].data());
    CHECK(on[f] == n[f], "%dx%d frame %d: %d key points, oracle %d", W, H, f, n[f], on[f]);
    if (on[f] != n[f]) continue;
    CHECK(memcmp(okp.data(), kp.data() + (size_t)f * cap, sizeof(vido_keypoint) * n[f]) == 0, "%dx%d frame %d: key points differ", W, H, f);
    int bad = 0;
    for (int k = 0; k < n[f]; k++) bad += memcmp(odesc[f].data() + (size_t)k * 32, desc.data() + ((size_t)f * cap + k) * 32, 32) != 0;
    CHECK(bad == 0, "%dx%d frame %d: %d of %d descriptors differ", W, H, f, bad, n[f]);
    printf("%dx%d frame %d: %d key points, descriptors %s\n", W, H, f, n[f], bad ? "DIFFER" : "identical");
  }
  // blurred levels of the last batch's slot 0 = frame ((nframes - 1) / max_batch) * max_batch
  {
    const int f = ((nframes - 1) / max_batch) * max_batch;
    int32_t lw[8], lh[8];
    vido_orb_level_info(ctx, lw, lh, nullptr, nullptr);
    size_t total = 0;
    for (int l = 0; l < cfg.nlevels; l++) total += (size_t)lw[l] * lh[l];
    std::vector<uint8_t> pyr(total);
    int64_t offs[9];
    vo_orb_pyramid(imgs.data() + (size_t)f * W * H, W, H, W, &p, pyr.data(), offs);
    for (int l = 0; l < cfg.nlevels; l++) {
      std::vector<uint8_t> want((size_t)lw[l] * lh[l]), got((size_t)lw[l] * lh[l]);
      vo_gauss7_u8(pyr.data() + offs[l], lw[l], lh[l], lw[l], want.data(), lw[l]);
      VCALL(vido_orb_get_blurred_level(ctx, 0, l, got.data()));
      CHECK(want == got, "%dx%d blurred level %d differs", W, H, l);
    }
  }
  // matcher, host pointers: frame 0 against frame 1, and degenerate train sets
  if (nframes >= 2 && on[0] == n[0] && on[1] == n[1]) {
    std::vector<int32_t> bi(n[0]), bd(n[0]), sd(n[0]), obi(n[0]), obd(n[0]), osd(n[0]);
    for (int nt : {n[1], 0, 1, 5, 9}) {
      VCALL(vido_hamming_match(ctx, desc.data(), n[0], desc.data() + (size_t)cap * 32, nt, bi.data(), bd.data(), sd.data()));
      vo_hamming_match(odesc[0].data(), n[0], odesc[1].data(), nt, obi.data(), obd.data(), osd.data());
      CHECK(bi == obi && bd == obd && sd == osd, "%dx%d hamming match (nt = %d) differs", W, H, nt);
    }
    printf("%dx%d matcher (host pointers): checked\n", W, H);
  }
  vido_destroy(ctx);
  return 0;
}

static int time_chain(int W, int H, int B, int iters) {
  vido_config cfg;
  vido_default_config(&cfg);
  cfg.width = W; cfg.height = H; cfg.max_batch = B;
  vido_ctx* ctx = vido_create(&cfg);
  if (!ctx) { printf("FAIL: vido_create: %s\n", vido_last_error(nullptr)); return 1; }
  const int cap = cfg.nfeatures + 64;
  std::vector<uint8_t> imgs((size_t)B * W * H);
  for (int f = 0; f < B; f++) make_frame(imgs.data() + (size_t)f * W * H, W, H, f);
  uint8_t *d_gray, *d_desc; vido_keypoint* d_kp; int32_t *d_n, *d_out;
  cudaMalloc(&d_gray, imgs.size()); cudaMalloc(&d_kp, sizeof(vido_keypoint) * B * cap); cudaMalloc(&d_n, 4 * B);
  cudaMalloc(&d_desc, (size_t)B * cap * 32); cudaMalloc(&d_out, (size_t)3 * B * cap * 4);
  cudaMemcpy(d_gray, imgs.data(), imgs.size(), cudaMemcpyHostToDevice);
  cudaStream_t st = (cudaStream_t)vido_stream(ctx);
  cudaEvent_t e[4];
  for (auto& ev : e) cudaEventCreate(&ev);
  float ms_ext = 0, ms_desc = 0, ms_match = 0;
  for (int it = -3; it < iters; it++) {
    cudaEventRecord(e[0], st);
    VCALL(vido_orb_extract_dev(ctx, d_gray, B, (size_t)W * H, W, d_kp, cap, d_n, 0));
    cudaEventRecord(e[1], st);
    VCALL(vido_orb_describe_dev(ctx, d_kp, d_n, B, cap, d_desc, 0));
    cudaEventRecord(e[2], st);
    VCALL(vido_hamming_match_dev(ctx, d_desc, (size_t)cap * 32, d_n, d_desc + (size_t)cap * 32, (size_t)cap * 32, d_n + 1, B - 1, cap, d_out,
                                 d_out + (size_t)B * cap, d_out + (size_t)2 * B * cap, 0));
    cudaEventRecord(e[3], st);
    cudaStreamSynchronize(st);
    if (it >= 0) {
      float a, b, c;
      cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[2], e[3]);
      ms_ext += a; ms_desc += b; ms_match += c;
    }
  }
  // the smoothing alone: the same call with all key-point counts zero (the descriptor threads leave at their first test)
  float ms_blur = 0;
  {
    int32_t* d_zero;
    cudaMalloc(&d_zero, 4 * B); cudaMemset(d_zero, 0, 4 * B);
    for (int it = -3; it < iters; it++) {
      cudaEventRecord(e[0], st);
      VCALL(vido_orb_describe_dev(ctx, d_kp, d_zero, B, cap, d_desc, 0));
      cudaEventRecord(e[1], st);
      cudaStreamSynchronize(st);
      float a; cudaEventElapsedTime(&a, e[0], e[1]);
      if (it >= 0) ms_blur += a;
    }
    cudaFree(d_zero);
    VCALL(vido_orb_describe_dev(ctx, d_kp, d_n, B, cap, d_desc, 1));   // descriptors back in place for the comparison below
  }
  cudaError_t err = cudaGetLastError();
  CHECK(err == cudaSuccess, "CUDA error after the timed chain: %s", cudaGetErrorString(err));
  std::vector<int32_t> n(B);
  cudaMemcpy(n.data(), d_n, 4 * B, cudaMemcpyDeviceToHost);
  // device-pointer results of pair 0 against the oracle
  {
    std::vector<uint8_t> desc((size_t)2 * cap * 32);
    cudaMemcpy(desc.data(), d_desc, desc.size(), cudaMemcpyDeviceToHost);
    std::vector<int32_t> bi(n[0]), bd(n[0]), sd(n[0]), obi(n[0]), obd(n[0]), osd(n[0]);
    cudaMemcpy(bi.data(), d_out, 4 * n[0], cudaMemcpyDeviceToHost);
    cudaMemcpy(bd.data(), d_out + (size_t)B * cap, 4 * n[0], cudaMemcpyDeviceToHost);
    cudaMemcpy(sd.data(), d_out + (size_t)2 * B * cap, 4 * n[0], cudaMemcpyDeviceToHost);
    vo_hamming_match(desc.data(), n[0], desc.data() + (size_t)cap * 32, n[1], obi.data(), obd.data(), osd.data());
    CHECK(bi == obi && bd == obd && sd == osd, "device-pointer matcher differs from the oracle");
  }
  long long pyr_px = 0;
  { int32_t lw[8], lh[8]; vido_orb_level_info(ctx, lw, lh, nullptr, nullptr); for (int l = 0; l < cfg.nlevels; l++) pyr_px += (long long)lw[l] * lh[l]; }
  printf("TIMING %dx%d batch %d, %d iterations, %d key points in frame 0:\n", W, H, B, iters, n[0]);
  printf("  extraction            %.3f ms per batch (%.1f us per frame)\n", ms_ext / iters, 1e3 * ms_ext / iters / B);
  printf("  blur + rBRIEF         %.3f ms per batch (%.1f us per frame); blur traffic %.2f MB per frame (read + write)\n", ms_desc / iters,
         1e3 * ms_desc / iters / B, 2e-6 * pyr_px);
  printf("  blur alone (+ an empty descriptor launch)  %.3f ms per batch: %.1f MB read + written = %.0f GB/s\n", ms_blur / iters,
         2e-6 * pyr_px * B, 2e-9 * pyr_px * B / (ms_blur / iters * 1e-3));
  printf("  Hamming match         %.3f ms per %d pairs (%.1f us per pair of ~%d x %d descriptors)\n", ms_match / iters, B - 1,
         1e3 * ms_match / iters / (B - 1), n[0], n[1]);
  cudaFree(d_gray); cudaFree(d_kp); cudaFree(d_n); cudaFree(d_desc); cudaFree(d_out);
  vido_destroy(ctx);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 2 && !strcmp(argv[1], "time")) {   // timing only: desc_check time <batch>
    if (time_chain(1242, 375, atoi(argv[2]), 20)) return 1;
    printf(fails ? "DESC_CHECK FAILED (%d)\n" : "DESC_CHECK PASSED\n", fails);
    return fails ? 2 : 0;
  }
#ifdef DRY_RUN   // no device: only show that the synthetic frames give the oracle something to describe
  for (int f = 0; f < 2; f++) {
    const int W = f ? 1242 : 333, H = f ? 375 : 211;
    std::vector<uint8_t> img((size_t)W * H), desc(20000 * 32);
    std::vector<vo_keypoint> okp(20000);
    make_frame(img.data(), W, H, 1);
    vo_orb_params p = {2500, 1.2f, 8, 20, 7};
    printf("%dx%d: oracle finds %d key points\n", W, H, vo_orb_extract_describe(img.data(), W, H, W, &p, okp.data(), 20000, desc.data()));
  }
  return 0;
#endif
  if (check_size(333, 211, 3, 2)) return 1;
  if (check_size(1242, 375, 2, 2)) return 1;
  if (time_chain(1242, 375, 16, 20)) return 1;
  printf(fails ? "DESC_CHECK FAILED (%d)\n" : "DESC_CHECK PASSED\n", fails);
  return fails ? 2 : 0;
}
