// desc_emul.cc -- TEST INFRASTRUCTURE: runs the per-thread bodies of the descriptor kernels (vido-slam_b200/csrc/desc_device.h,
// the very source the __global__ wrappers of desc_kernels.cu call) on the CPU, one simulated thread after the other over the
// same launch grids (grid sizes come from the shared helpers; threads beyond the work list are simulated too, so the guards are
// exercised).  The bodies use no shared memory, barriers or atomics and no thread reads what another thread of the same launch
// writes, so sequential execution is a valid schedule.  Built by tests/test_desc_emul.py with g++ -ffp-contract=off.
#include "../vido-slam_b200/csrc/desc_device.h"

extern "C" {

// pyramid geometry as in orb_setup (vido-slam_b200/csrc/orb_kernels.cu): level l at base[l] + frame * frame_stride[l], rows `pitch[l]` apart
void emul_blur_v(const uint8_t* pyr, uint8_t* out, int nlevels, const int* w, const int* h, const int* pitch, const long long* base,
                 const long long* frame_stride, int nframes, int version) {
  BlurParams P;
  memset(&P, 0, sizeof P);
  for (int l = 0; l < nlevels; l++) {
    if (version == 2) blur2_params_add_level(P, l, w[l], h[l], pitch[l], base[l], frame_stride[l]);
    else blur_params_add_level(P, l, w[l], h[l], pitch[l], base[l], frame_stride[l]);
  }
  P.nframes = nframes;
  const unsigned gx = blur_grid_x(P);
  for (int z = 0; z < nframes; z++)
    for (unsigned b = 0; b < gx; b++)
      for (int t = 0; t < BLUR_THREADS; t++) {
        if (version == 2) blur7_thread_v2((int)(b * BLUR_THREADS + t), z, P, pyr, out);
        else blur7_thread((int)(b * BLUR_THREADS + t), z, P, pyr, out);
      }
}

void emul_blur(const uint8_t* pyr, uint8_t* out, int nlevels, const int* w, const int* h, const int* pitch, const long long* base,
               const long long* frame_stride, int nframes) {
  emul_blur_v(pyr, out, nlevels, w, h, pitch, base, frame_stride, nframes, 1);
}

void emul_rbrief(const uint8_t* blurred, int nlevels, const int* w, const int* h, const int* pitch, const long long* base,
                 const long long* frame_stride, const float* scale, const DescKeyPoint* kps, const int32_t* nkp, int nframes,
                 int cap_per_frame, const int8_t* pattern, uint8_t* desc) {
  DescParams P;
  memset(&P, 0, sizeof P);
  for (int l = 0; l < nlevels; l++) desc_params_add_level(P, l, w[l], h[l], pitch[l], base[l], frame_stride[l], scale[l]);
  P.nframes = nframes;
  P.cap_per_frame = cap_per_frame;
  const unsigned gx = rbrief_grid_x(cap_per_frame);
  for (int z = 0; z < nframes; z++)
    for (unsigned b = 0; b < gx; b++)
      for (int t = 0; t < RBRIEF_THREADS; t++) rbrief_thread((int)(b * RBRIEF_THREADS + t), z, P, blurred, kps, nkp, pattern, desc);
}

void emul_hamming(const uint8_t* q, long long q_stride, const int32_t* nq, const uint8_t* t, long long t_stride, const int32_t* nt,
                  int npairs, int qcap, int32_t* part, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
  HamParams P;
  P.npairs = npairs; P.qcap = qcap; P.q_stride = q_stride; P.t_stride = t_stride;
  const unsigned gx = hamming_grid_x(qcap);
  for (int z = 0; z < npairs; z++)
    for (int y = 0; y < HAM_CHUNKS; y++)
      for (unsigned b = 0; b < gx; b++)
        for (int th = 0; th < HAM_THREADS; th++) hamming_partial_thread((int)(b * HAM_THREADS + th), y, z, P, q, t, nq, nt, part);
  for (int z = 0; z < npairs; z++)
    for (unsigned b = 0; b < gx; b++)
      for (int th = 0; th < HAM_THREADS; th++)
        hamming_merge_thread((int)(b * HAM_THREADS + th), z, P, nq, part, best_idx, best_dist, second_dist);
}

long long emul_hamming_part_bytes(int npairs, int qcap) { return (long long)hamming_part_bytes(npairs, qcap); }

}  // extern "C"
