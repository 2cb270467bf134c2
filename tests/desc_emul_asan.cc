// desc_emul_asan.cc -- TEST INFRASTRUCTURE: the per-thread bodies of the descriptor kernels (vido-slam_b200/csrc/desc_device.h)
// walked over their launch grids under AddressSanitizer, on heap buffers of exactly the size the product allocates (orb_setup's
// pyramid layout: pitch = width rounded up to 64, frame stride rounded up to 256, levels back to back; descriptor / key-point /
// matcher arrays without slack).  Any read or write outside those allocations -- which a GPU would not necessarily report -- aborts
// the run.  Also asserts the 4-byte alignment of every vector access (compile-time switch of this file only).
// Built and run by tests/test_desc_emul.py:  g++ -O1 -g -fsanitize=address -ffp-contract=off -DVIDO_EMUL_CHECK_ALIGN tests/desc_emul_asan.cc
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/vido_orb_pattern.h"
#include "../vido-slam_b200/csrc/desc_device.h"

static uint32_t rs = 99;
static uint32_t rnd() { rs = rs * 1664525u + 1013904223u; return rs >> 8; }

struct Layout {
  int n = 0, w[8], h[8], pitch[8];
  long long base[8], fs[8], bytes = 0;
  float scale[8];
};

static Layout make_layout(int W, int H, int nlevels, int nframes) {
  Layout L;
  L.n = nlevels;
  float s = 1.f;
  for (int l = 0; l < nlevels; l++) {
    L.scale[l] = s;
    L.w[l] = (int)(W / s + 0.5f); L.h[l] = (int)(H / s + 0.5f);
    L.pitch[l] = (L.w[l] + 63) / 64 * 64;
    L.fs[l] = ((long long)L.pitch[l] * L.h[l] + 255) / 256 * 256;
    L.base[l] = L.bytes;
    L.bytes += L.fs[l] * nframes;
    s *= 1.2f;
  }
  return L;
}

static int run_size(int W, int H, int nlevels, int nframes) {
  Layout L = make_layout(W, H, nlevels, nframes);
  uint8_t* pyr = (uint8_t*)malloc(L.bytes);
  uint8_t* blur = (uint8_t*)malloc(L.bytes);
  for (long long i = 0; i < L.bytes; i++) pyr[i] = (uint8_t)rnd();
  memset(blur, 0, L.bytes);
  BlurParams B;
  memset(&B, 0, sizeof B);
  DescParams D;
  memset(&D, 0, sizeof D);
  for (int l = 0; l < L.n; l++) {
    blur_params_add_level(B, l, L.w[l], L.h[l], L.pitch[l], L.base[l], L.fs[l]);
    desc_params_add_level(D, l, L.w[l], L.h[l], L.pitch[l], L.base[l], L.fs[l], L.scale[l]);
  }
  B.nframes = nframes;
  for (int z = 0; z < nframes; z++)
    for (unsigned b = 0; b < blur_grid_x(B); b++)
      for (int t = 0; t < BLUR_THREADS; t++) blur7_thread((int)(b * BLUR_THREADS + t), z, B, pyr, blur);
  {   // variant 2 of the smoothing kernel into a second buffer: same bytes
    uint8_t* blur2 = (uint8_t*)malloc(L.bytes);
    memset(blur2, 0, L.bytes);
    BlurParams B2;
    memset(&B2, 0, sizeof B2);
    for (int l = 0; l < L.n; l++) blur2_params_add_level(B2, l, L.w[l], L.h[l], L.pitch[l], L.base[l], L.fs[l]);
    B2.nframes = nframes;
    for (int z = 0; z < nframes; z++)
      for (unsigned b = 0; b < blur_grid_x(B2); b++)
        for (int t = 0; t < BLUR_THREADS; t++) blur7_thread_v2((int)(b * BLUR_THREADS + t), z, B2, pyr, blur2);
    if (memcmp(blur, blur2, L.bytes) != 0) { printf("variant 2 differs at %dx%d\n", W, H); abort(); }
    free(blur2);
  }
  // key points everywhere, including the border and outside the image (foreign key points must not fault)
  const int cap = 257;
  DescKeyPoint* kp = (DescKeyPoint*)malloc(sizeof(DescKeyPoint) * cap * nframes);
  int32_t* nkp = (int32_t*)malloc(sizeof(int32_t) * (size_t)nframes);
  uint8_t* desc = (uint8_t*)malloc((size_t)cap * nframes * 32);
  int8_t* pat = (int8_t*)malloc(sizeof vido_orb_pattern_31);
  memcpy(pat, vido_orb_pattern_31, sizeof vido_orb_pattern_31);
  for (int f = 0; f < nframes; f++) {
    nkp[f] = cap - f;
    for (int k = 0; k < cap; k++) {
      DescKeyPoint& q = kp[f * cap + k];
      q.octave = (int)(rnd() % (unsigned)(L.n + 2)) - 1;   // also out-of-range octaves
      const int l = q.octave < 0 || q.octave >= L.n ? 0 : q.octave;
      const float lx = (float)((int)(rnd() % (unsigned)(L.w[l] + 8)) - 4), ly = (float)((int)(rnd() % (unsigned)(L.h[l] + 8)) - 4);
      q.x = l ? lx * L.scale[l] : lx; q.y = l ? ly * L.scale[l] : ly;
      q.angle = (float)(rnd() % 36000) / 100.f; q.size = 31.f; q.response = 1.f;
    }
  }
  D.nframes = nframes; D.cap_per_frame = cap;
  for (int z = 0; z < nframes; z++)
    for (unsigned b = 0; b < rbrief_grid_x(cap); b++)
      for (int t = 0; t < RBRIEF_THREADS; t++) rbrief_thread((int)(b * RBRIEF_THREADS + t), z, D, blur, kp, nkp, pat, desc);
  // matcher: frame f against frame f + 1 on the same arrays, exact-size outputs
  if (nframes > 1) {
    const int np = nframes - 1;
    HamParams P;
    P.npairs = np; P.qcap = cap; P.q_stride = (long long)cap * 32; P.t_stride = (long long)cap * 32;
    int32_t* part = (int32_t*)malloc(hamming_part_bytes(np, cap));
    int32_t* out = (int32_t*)malloc((size_t)3 * np * cap * 4);
    for (int z = 0; z < np; z++)
      for (int y = 0; y < HAM_CHUNKS; y++)
        for (unsigned b = 0; b < hamming_grid_x(cap); b++)
          for (int t = 0; t < HAM_THREADS; t++)
            hamming_partial_thread((int)(b * HAM_THREADS + t), y, z, P, desc, desc + (size_t)cap * 32, nkp, nkp + 1, part);
    for (int z = 0; z < np; z++)
      for (unsigned b = 0; b < hamming_grid_x(cap); b++)
        for (int t = 0; t < HAM_THREADS; t++)
          hamming_merge_thread((int)(b * HAM_THREADS + t), z, P, nkp, part, out, out + (size_t)np * cap, out + (size_t)2 * np * cap);
    free(part); free(out);
  }
  free(pyr); free(blur); free(kp); free(nkp); free(desc); free(pat);
  return 0;
}

int main() {
  const int sizes[][2] = {{1242, 375}, {333, 211}, {640, 480}, {64, 64}, {128, 40}, {65, 33}, {67, 35}, {1024, 64}, {40, 40}};
  for (auto& s : sizes) {
    const int levels = (s[0] >= 300 && s[1] >= 200) ? 8 : 3;
    run_size(s[0], s[1], levels, 3);
    printf("%dx%d: %d levels, 3 frames walked\n", s[0], s[1], levels);
  }
  printf("ASAN_WALK_OK\n");
  return 0;
}
