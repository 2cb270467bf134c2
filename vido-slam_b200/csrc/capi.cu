// capi.cu -- C-ABI entry points of libvido_b200.so (declared in include/vido_b200.h).
#include <cstring>
#include <new>

#include "ctx.h"

static thread_local std::string g_create_err;

extern "C" {

int vido_version(void) { return 100; }

void vido_default_config(vido_config* c) {
  // src/config/kitti_config.yaml of the reference, with the full-resolution KITTI camera the 1242x375
  // configurations of BASELINE.json use
  memset(c, 0, sizeof *c);
  c->width = 1242; c->height = 375;
  c->fx = 718.856f; c->fy = 718.856f; c->cx = 607.1928f; c->cy = 185.2157f; c->bf = 386.1448f;
  c->choose_data = 2;
  c->depth_map_factor = 256.f;
  c->th_depth_bg = 5000.f; c->th_depth_obj = 25.f;
  c->max_track_bg = 1000; c->max_track_obj = 500;
  c->sf_mg_thres = 0.12f; c->sf_ds_thres = 0.3f;
  c->b_joint = 1;
  c->window_size = 20;
  c->nfeatures = 2500; c->scale_factor = 1.2f; c->nlevels = 8; c->ini_th_fast = 20; c->min_th_fast = 7;
  c->rgb = 0;
  c->max_batch = 8;
  c->device = 0;
}

const char* vido_last_error(vido_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

vido_ctx* vido_create(const vido_config* cfg) {
  if (!cfg) { g_create_err = "null config"; return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // no CPU fallback by design: the product path is CUDA only
    g_create_err = std::string("no CUDA device available: ") + cudaGetErrorString(e);
    return nullptr;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "bad device ordinal"; return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major < 10) {
    g_create_err = "libvido_b200 is built for sm_100a (Blackwell) only; found " + std::string(prop.name);
    return nullptr;
  }
  vido_ctx* ctx = new (std::nothrow) vido_ctx();
  if (!ctx) { g_create_err = "out of memory"; return nullptr; }
  ctx->cfg = *cfg;
  if (ctx->cfg.max_batch < 1) ctx->cfg.max_batch = 1;
  if (cfg->window_size < 1 || cfg->window_size > 24) {   // the window solver keeps the reduced system in shared memory (BA_MAX_W)
    g_create_err = "window_size must be in 1..24";
    delete ctx;
    return nullptr;
  }
  ctx->device = cfg->device;
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaSetDevice(cfg->device) != cudaSuccess || vido_create_stream(&ctx->stream, true) != cudaSuccess) {
    g_create_err = "cudaSetDevice/cudaStreamCreate failed";
    delete ctx;
    return nullptr;
  }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  int rc = orb_setup(ctx);
  if (rc == VIDO_OK) rc = ba_setup(ctx, 24, 16384, 131072);
  if (rc == VIDO_OK) rc = po_setup(ctx, 8192, 16);
  if (rc == VIDO_OK) rc = pnp_setup(ctx, 8192, 2048);
  if (rc == VIDO_OK) rc = trk_setup(ctx);
  if (rc == VIDO_OK) rc = chain_setup(ctx, ctx->cfg.max_batch);
  if (rc != VIDO_OK) {
    g_create_err = ctx->err;
    chain_teardown(ctx);
    trk_teardown(ctx);
    pnp_teardown(ctx);
    po_teardown(ctx);
    ba_teardown(ctx);
    orb_teardown(ctx);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return nullptr;
  }
  return ctx;
}

void vido_destroy(vido_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  chain_teardown(ctx);
  trk_teardown(ctx);
  pnp_teardown(ctx);
  po_teardown(ctx);
  ba_teardown(ctx);
  desc_teardown(ctx);
  orb_teardown(ctx);
  if (ctx->um_ws) cudaFree(ctx->um_ws);
  for (int k = 0; k < 3; k++) if (ctx->raw_stage[k]) cudaFree(ctx->raw_stage[k]);
  for (int k = 0; k < 4; k++) if (ctx->raw_dev[k]) cudaFree(ctx->raw_dev[k]);
  if (ctx->fba_arena) cudaFree(ctx->fba_arena);
  for (int k = 0; k < 3; k++) if (ctx->scratch[k]) cudaFree(ctx->scratch[k]);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int64_t vido_kernel_launches(vido_ctx* ctx) { return ctx ? ctx->launches.load() : (int64_t)0; }
void* vido_stream(vido_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vido_sync(vido_ctx* ctx) {
  if (!ctx) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  { int rc = trk_drain(ctx); if (rc) return rc; }   // window solves a vido_track_frames call without statistics left queued
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

int vido_orb_level_info(vido_ctx* ctx, int32_t* w, int32_t* h, int32_t* quota, float* scale) {
  if (!ctx) return VIDO_ERR_ARG;
  for (int l = 0; l < ctx->nlevels; l++) {
    if (w) w[l] = ctx->lv[l].w;
    if (h) h[l] = ctx->lv[l].h;
    if (quota) quota[l] = ctx->lv[l].quota;
    if (scale) scale[l] = ctx->lv[l].scale;
  }
  return ctx->nlevels;
}

static int check_dev_err(vido_ctx* ctx) {
  int32_t flag = 0;
  VIDO_CUDA(cudaMemcpyAsync(&flag, ctx->d_err, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream));
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag) {
    char buf[128];
    snprintf(buf, sizeof buf, "ORB front-end capacity flag 0x%x (1: >65535 candidates/level, 2-16: quad-tree arrays, 32: output cap)", flag);
    ctx->err = buf;
    cudaMemsetAsync(ctx->d_err, 0, sizeof flag, ctx->stream);
    return VIDO_ERR_CAPACITY;
  }
  return VIDO_OK;
}

int vido_orb_extract_dev(vido_ctx* ctx, const uint8_t* d_gray, int nframes, size_t frame_stride, int stride,
                         vido_keypoint* d_out, int cap_per_frame, int32_t* d_n_out, int sync) {
  if (!ctx || !d_gray || !d_out || !d_n_out || cap_per_frame < 1 || stride < ctx->cfg.width) {
    if (ctx) ctx->err = "vido_orb_extract_dev: bad argument";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  trk_quiesce(ctx);
  int rc = orb_run(ctx, d_gray, nframes, frame_stride, stride, d_out, cap_per_frame, d_n_out);
  if (rc != VIDO_OK) return rc;
  if (sync) return check_dev_err(ctx);
  return VIDO_OK;
}

int vido_orb_extract(vido_ctx* ctx, const uint8_t* gray, int nframes, size_t frame_stride, int stride, vido_keypoint* out,
                     int cap_per_frame, int32_t* n_out) {
  if (!ctx || !gray || !out || !n_out || cap_per_frame < 1 || stride < ctx->cfg.width) {
    if (ctx) ctx->err = "vido_orb_extract: bad argument";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  trk_quiesce(ctx);
  const vido_config& c = ctx->cfg;
  int done = 0;
  while (done < nframes) {
    const int B = std::min(c.max_batch, nframes - done);
    for (int b = 0; b < B; b++)
      VIDO_CUDA(cudaMemcpy2DAsync(ctx->d_in + (size_t)b * ctx->in_pitch * c.height, ctx->in_pitch,
                                  gray + (size_t)(done + b) * frame_stride, stride, c.width, c.height,
                                  cudaMemcpyHostToDevice, ctx->stream));
    int rc = orb_run(ctx, ctx->d_in, B, (size_t)ctx->in_pitch * c.height, ctx->in_pitch, ctx->d_kp, ctx->kp_cap, ctx->d_nkp);
    if (rc != VIDO_OK) return rc;
    VIDO_CUDA(cudaMemcpyAsync(n_out + done, ctx->d_nkp, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_dev_err(ctx);  // synchronises
    if (rc != VIDO_OK) return rc;
    for (int b = 0; b < B; b++) {
      int n = n_out[done + b];
      if (n > cap_per_frame) { ctx->err = "vido_orb_extract: cap_per_frame too small"; return VIDO_ERR_CAPACITY; }
      VIDO_CUDA(cudaMemcpyAsync(out + (size_t)(done + b) * cap_per_frame, ctx->d_kp + (size_t)b * ctx->kp_cap,
                                sizeof(vido_keypoint) * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
    done += B;
  }
  return VIDO_OK;
}

/* ---- descriptor stage (desc_kernels.cu) ---- */
int vido_orb_describe_dev(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int nframes, int cap_per_frame,
                          uint8_t* d_desc, int sync) {
  if (!ctx || !d_kps || !d_nkp || !d_desc || cap_per_frame < 1) {
    if (ctx) ctx->err = "vido_orb_describe_dev: bad argument";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  trk_quiesce(ctx);
  int rc = desc_run(ctx, d_kps, d_nkp, nframes, cap_per_frame, d_desc);
  if (rc != VIDO_OK) return rc;
  if (sync) VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

int vido_orb_extract_describe(vido_ctx* ctx, const uint8_t* gray, int nframes, size_t frame_stride, int stride, vido_keypoint* out,
                              int cap_per_frame, int32_t* n_out, uint8_t* desc) {
  if (!ctx || !gray || !out || !n_out || !desc || cap_per_frame < 1 || stride < ctx->cfg.width) {
    if (ctx) ctx->err = "vido_orb_extract_describe: bad argument";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  trk_quiesce(ctx);
  const vido_config& c = ctx->cfg;
  uint8_t* d_desc = desc_staging(ctx);
  if (!d_desc) return VIDO_ERR_CUDA;
  int done = 0;
  while (done < nframes) {
    const int B = std::min(c.max_batch, nframes - done);
    for (int b = 0; b < B; b++)
      VIDO_CUDA(cudaMemcpy2DAsync(ctx->d_in + (size_t)b * ctx->in_pitch * c.height, ctx->in_pitch,
                                  gray + (size_t)(done + b) * frame_stride, stride, c.width, c.height,
                                  cudaMemcpyHostToDevice, ctx->stream));
    int rc = orb_run(ctx, ctx->d_in, B, (size_t)ctx->in_pitch * c.height, ctx->in_pitch, ctx->d_kp, ctx->kp_cap, ctx->d_nkp);
    if (rc != VIDO_OK) return rc;
    rc = desc_run(ctx, ctx->d_kp, ctx->d_nkp, B, ctx->kp_cap, d_desc);
    if (rc != VIDO_OK) return rc;
    VIDO_CUDA(cudaMemcpyAsync(n_out + done, ctx->d_nkp, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_dev_err(ctx);  // synchronises
    if (rc != VIDO_OK) return rc;
    for (int b = 0; b < B; b++) {
      const int n = n_out[done + b];
      if (n > cap_per_frame) { ctx->err = "vido_orb_extract_describe: cap_per_frame too small"; return VIDO_ERR_CAPACITY; }
      VIDO_CUDA(cudaMemcpyAsync(out + (size_t)(done + b) * cap_per_frame, ctx->d_kp + (size_t)b * ctx->kp_cap,
                                sizeof(vido_keypoint) * n, cudaMemcpyDeviceToHost, ctx->stream));
      VIDO_CUDA(cudaMemcpyAsync(desc + (size_t)(done + b) * cap_per_frame * 32, d_desc + (size_t)b * ctx->kp_cap * 32, (size_t)n * 32,
                                cudaMemcpyDeviceToHost, ctx->stream));
    }
    VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
    done += B;
  }
  return VIDO_OK;
}

int vido_orb_get_blurred_level(vido_ctx* ctx, int frame, int level, uint8_t* out) {
  if (!ctx || !out || level < 0 || level >= ctx->nlevels || frame < 0 || frame >= ctx->cfg.max_batch) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return desc_get_blurred_level(ctx, frame, level, out);
}

int vido_hamming_match(vido_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* best_idx, int32_t* best_dist,
                       int32_t* second_dist) {
  if (!ctx || nq < 0 || nt < 0 || (nq && (!query || !best_idx || !best_dist || !second_dist)) || (nt && !train)) {
    if (ctx) ctx->err = "vido_hamming_match: bad argument";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  return desc_match_host(ctx, query, nq, train, nt, best_idx, best_dist, second_dist);
}

int vido_hamming_match_dev(vido_ctx* ctx, const uint8_t* d_query, size_t query_stride, const int32_t* d_nq, const uint8_t* d_train,
                           size_t train_stride, const int32_t* d_nt, int npairs, int qcap, int32_t* d_best_idx, int32_t* d_best_dist,
                           int32_t* d_second_dist, int sync) {
  if (!ctx || !d_query || !d_nq || !d_train || !d_nt || npairs < 1 || qcap < 1 || !d_best_idx || !d_best_dist || !d_second_dist ||
      ((size_t)d_query & 3) || ((size_t)d_train & 3) || (query_stride & 3) || (train_stride & 3)) {
    if (ctx) ctx->err = "vido_hamming_match_dev: bad argument (descriptor arrays must be 4-byte aligned)";
    return VIDO_ERR_ARG;
  }
  cudaSetDevice(ctx->device);
  int rc = desc_match_device_api(ctx, d_query, query_stride, d_nq, d_train, train_stride, d_nt, npairs, qcap, d_best_idx, d_best_dist,
                                 d_second_dist);
  if (rc != VIDO_OK) return rc;
  if (sync) VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

int vido_bgr_to_gray_dev(vido_ctx* ctx, const uint8_t* d_bgr, int nframes, size_t frame_stride, int stride, uint8_t* d_gray,
                         size_t gray_frame_stride, int gray_stride) {
  if (!ctx || !d_bgr || !d_gray) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return orb_bgr_to_gray(ctx, d_bgr, nframes, frame_stride, stride, d_gray, gray_frame_stride, gray_stride);
}

int vido_orb_get_level(vido_ctx* ctx, int frame, int level, uint8_t* out) {
  if (!ctx || !out || level < 0 || level >= ctx->nlevels || frame < 0 || frame >= ctx->cfg.max_batch) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  const OrbLevel& L = ctx->lv[level];
  VIDO_CUDA(cudaMemcpy2DAsync(out, L.w, ctx->d_pyr + L.base + (size_t)frame * L.frame_stride, L.pitch, L.w, L.h,
                              cudaMemcpyDeviceToHost, ctx->stream));
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

int vido_orb_get_candidates(vido_ctx* ctx, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores, int cap,
                            int32_t* n) {
  if (!ctx || level < 0 || level >= ctx->nlevels || frame < 0 || frame >= ctx->cfg.max_batch || !n) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  const OrbLevel& L = ctx->lv[level];
  std::vector<int32_t> cnt(std::max(L.ncells, 1));
  std::vector<uint32_t> slots(ctx->slots_per_frame);
  VIDO_CUDA(cudaMemcpyAsync(cnt.data(), ctx->d_cell_count + (size_t)frame * ctx->cells_per_frame + L.cell_begin,
                            sizeof(int32_t) * L.ncells, cudaMemcpyDeviceToHost, ctx->stream));
  VIDO_CUDA(cudaMemcpyAsync(slots.data(), ctx->d_slots + (size_t)frame * ctx->slots_per_frame,
                            sizeof(uint32_t) * ctx->slots_per_frame, cudaMemcpyDeviceToHost, ctx->stream));
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  int k = 0;
  for (int i = 0; i < L.ncells; i++) {
    const OrbCell& cell = ctx->cells[L.cell_begin + i];
    for (int j = 0; j < cnt[i]; j++, k++) {
      if (k < cap) {
        uint32_t v = slots[cell.slot_base + j];
        xs[k] = (int)(v >> 20);
        ys[k] = (int)((v >> 8) & 0xfff);
        scores[k] = (int)(v & 0xff);
      }
    }
  }
  *n = k;
  return VIDO_OK;
}

void vido_ba_default_params(vido_ba_problem* p) {
  // hard-coded constants of Optimizer::PartialBatchOptimization (src/Optimizer.cc:183,192-216,806)
  p->max_iterations = 100;
  p->sigma2_cam = 0.0001f;
  p->sigma2_3d = 16.f;
  p->huber_cam = 0.01f;
  p->huber_3d = 0.01f;
  p->gain_threshold = 1e-3f;
  p->fix_first = 0;
}

int vido_ba_partial(vido_ctx* ctx, vido_ba_problem* p, vido_lm_stats* stats) {
  if (!ctx || !p) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  { int rc = trk_drain(ctx); if (rc) return rc; }   // the solver workspace is shared with the tracker's queued window solves
  return ba_partial_host(ctx, p, stats);
}

void vido_poseopt_default_params(vido_poseopt_problem* p) {
  p->info_flow = 0.1f; p->info_prior = 0.3f; p->rp_thres = 0.04f; p->chi2_th = 5.991f;
  p->rounds = 4; p->its = 100;
}

int vido_pose_opt_flow2(vido_ctx* ctx, vido_poseopt_problem* problems, int nproblems, vido_lm_stats* stats) {
  if (!ctx || !problems) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return po_flow2_host(ctx, problems, nproblems, stats);
}

void vido_pnp_default_params(vido_pnp_problem* p) { p->iters = 500; p->reproj_err = 0.4f; p->confidence = 0.98f; }

int vido_init_model(vido_ctx* ctx, vido_pnp_problem* p) {
  if (!ctx || !p) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return pnp_init_model_host(ctx, p);
}

int vido_update_mask_dev(vido_ctx* ctx, const int32_t* sem_label, const float* corres_xy, int n, const int32_t* d_mask_last,
                         const float* d_flow_last, int32_t* d_mask_cur, int32_t* uniq_out, int32_t* recovered, int cap) {
  if (!ctx || n < 0 || (n > 0 && (!sem_label || !corres_xy || !d_mask_last || !d_flow_last || !d_mask_cur))) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return assoc_update_mask(ctx, sem_label, corres_xy, n, d_mask_last, d_flow_last, d_mask_cur, uniq_out, recovered, cap);
}

int vido_depth_prep_dev(vido_ctx* ctx, float* d_depth, int nframes, size_t frame_stride_elems, int stride_elems) {
  if (!ctx || !d_depth) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return assoc_depth_prep(ctx, d_depth, nframes, frame_stride_elems, stride_elems);
}

int vido_frame_associate_dev(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int kp_cap, const float* d_depth,
                             const float* d_flow, const int32_t* d_mask, int nframes, int raw_depth, int32_t* d_idx,
                             float* d_corres_xy, float* d_flow_xy, float* d_depth_out, int32_t* d_n, int out_cap) {
  if (!ctx || !d_kps || !d_nkp || !d_depth || !d_flow || !d_mask) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return assoc_frame_associate(ctx, d_kps, d_nkp, kp_cap, d_depth, d_flow, d_mask, nframes, raw_depth, d_idx, d_corres_xy,
                               d_flow_xy, d_depth_out, d_n, out_cap);
}

int vido_frame_sample_objects_dev(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int nframes,
                                  int raw_depth, float* d_keys_xy, float* d_corres_xy, float* d_flow_xy, float* d_depth_out,
                                  int32_t* d_label, int32_t* d_n, int out_cap) {
  if (!ctx || !d_depth || !d_flow || !d_mask) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return assoc_sample_objects(ctx, d_depth, d_flow, d_mask, nframes, raw_depth, d_keys_xy, d_corres_xy, d_flow_xy, d_depth_out,
                              d_label, d_n, out_cap);
}

int vido_gather_dev(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int frame, int raw_depth,
                    const float* d_xy, int n, int32_t* d_mask_out, float* d_depth_out, float* d_flow_out) {
  if (!ctx) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return assoc_gather(ctx, d_depth, d_flow, d_mask, frame, raw_depth, d_xy, n, d_mask_out, d_depth_out, d_flow_out);
}

int vido_set_depth_scale(vido_ctx* ctx, float mscale) {
  if (!ctx) return VIDO_ERR_ARG;
  ctx->mscale = mscale;
  return VIDO_OK;
}

int vido_track_frames(vido_ctx* ctx, const vido_frame_inputs* frames, int nframes, float* Tcw_out, vido_track_stats* stats) {
  if (!ctx || !frames || !Tcw_out || nframes < 1) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_track_chunk(ctx, frames, nframes, Tcw_out, stats);
}
int vido_track_reset(vido_ctx* ctx) {
  if (!ctx) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_reset(ctx);
}

int vido_track_prefetch(vido_ctx* ctx, const vido_frame_inputs* frames, int nframes) {
  if (!ctx || (!frames && nframes > 0)) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_prefetch(ctx, frames, nframes);
}
int vido_map_num_frames(vido_ctx* ctx) { return ctx ? trk_num_frames(ctx) : VIDO_ERR_ARG; }
int vido_map_get_poses(vido_ctx* ctx, float* poses, int cap) { return (ctx && poses) ? trk_get_map_poses(ctx, poses, cap) : VIDO_ERR_ARG; }
int vido_map_get_static(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap) {
  return ctx ? trk_get_static(ctx, frame, xy, depth, p3, asso, cap) : VIDO_ERR_ARG;
}

int vido_map_get_dynamic(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap) {
  return ctx ? trk_get_dynamic(ctx, frame, xy, depth, p3, asso, label, cap) : VIDO_ERR_ARG;
}
int vido_map_get_objects(vido_ctx* ctx, int frame, int32_t* label, int32_t* sem_label, float* motion, float* centre, int cap) {
  return ctx ? trk_get_objects(ctx, frame, label, sem_label, motion, centre, cap) : VIDO_ERR_ARG;
}
int vido_map_get_dyn_tracks(vido_ctx* ctx, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap) {
  return ctx ? trk_get_dyn_tracks(ctx, len, obj_id, first_frame, first_feat, cap) : VIDO_ERR_ARG;
}

void vido_fba_default_params(vido_fba_problem* p) { if (p) vido_fba_default_params_impl(p); }
int vido_ba_full(vido_ctx* ctx, vido_fba_problem* p, vido_lm_stats* stats) {
  if (!ctx || !p) return VIDO_ERR_ARG;
  if (p->n_poses < 0 || p->n_motions < 0 || p->n_points < 0 || p->n_obs < 0 || p->n_e6 < 0 || p->n_tern < 0) { ctx->err = "negative size"; return VIDO_ERR_ARG; }
  cudaSetDevice(ctx->device);
  return fba_solve_host(ctx, p, stats);
}
int vido_convert_raw(vido_ctx* ctx, const uint8_t* bayer, const uint16_t* depth16, const uint8_t* mask8, int nframes, uint8_t* d_bgr,
                     float* d_depth, int32_t* d_mask) {
  if (!ctx || nframes < 1) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return input_convert_raw(ctx, bayer, depth16, mask8, nframes, d_bgr, d_depth, d_mask);
}
int vido_track_raw_frames(vido_ctx* ctx, const vido_raw_inputs* frames, int nframes, float* Tcw_out, vido_track_stats* stats) {
  if (!ctx || !frames || !Tcw_out || nframes < 1) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  const size_t px = (size_t)ctx->cfg.width * ctx->cfg.height;
  const int cap = std::max(1, ctx->cfg.max_batch);
  if (ctx->raw_dev_frames < cap) {
    for (int k = 0; k < 4; k++) { if (ctx->raw_dev[k]) cudaFree(ctx->raw_dev[k]); ctx->raw_dev[k] = nullptr; }
    ctx->raw_dev_frames = 0;
    const size_t bytes[4] = {3 * px, 4 * px, 8 * px, 4 * px};
    for (int k = 0; k < 4; k++) VIDO_CUDA(cudaMalloc(&ctx->raw_dev[k], bytes[k] * cap));
    ctx->raw_dev_frames = cap;
  }
  uint8_t* d_bgr = (uint8_t*)ctx->raw_dev[0]; float* d_depth = (float*)ctx->raw_dev[1];
  float* d_flow = (float*)ctx->raw_dev[2]; int32_t* d_mask = (int32_t*)ctx->raw_dev[3];
  std::vector<vido_frame_inputs> in(cap);
  for (int done = 0; done < nframes; done += cap) {
    const int B = std::min(cap, nframes - done);
    for (int b = 0; b < B; b++) {
      const vido_raw_inputs& r = frames[done + b];
      if (!r.bayer || !r.depth16 || !r.flow || !r.mask8) { ctx->err = "vido_track_raw_frames: a frame has a NULL input"; return VIDO_ERR_ARG; }
      int rc = input_convert_raw(ctx, r.bayer, r.depth16, r.mask8, 1, d_bgr + 3 * px * b, d_depth + px * b, d_mask + px * b);
      if (rc) return rc;
      VIDO_CUDA(cudaMemcpyAsync(d_flow + 2 * px * b, r.flow, 8 * px, cudaMemcpyHostToDevice, ctx->stream));
      // (one frame at a time: the staging buffers of vido_convert_raw hold a single frame here)
      vido_frame_inputs& f = in[b];
      memset(&f, 0, sizeof f);
      f.image = d_bgr + 3 * px * b; f.channels = 3; f.on_device = 1;
      f.depth = d_depth + px * b; f.flow = d_flow + 2 * px * b; f.mask = d_mask + px * b;
      f.timestamp = r.timestamp;
    }
    VIDO_CUDA(cudaStreamSynchronize(ctx->stream));   // the front-end reads the converted frames on its own stream
    int rc = trk_track_chunk(ctx, in.data(), B, Tcw_out + 16 * (size_t)done, stats ? stats + done : nullptr);
    if (rc < 0) return rc;
  }
  return VIDO_OK;
}
int vido_fba_save_g2o(const vido_fba_problem* p, const char* path, int precision) {
  if (!p || !path) return VIDO_ERR_ARG;
  if (p->n_poses < 0 || p->n_motions < 0 || p->n_points < 0 || p->n_obs < 0 || p->n_e6 < 0 || p->n_tern < 0) return VIDO_ERR_ARG;
  return fba_save_g2o(p, path, precision);
}
int vido_full_batch(vido_ctx* ctx, vido_lm_stats* stats, int32_t* sizes) {
  if (!ctx) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_full_batch(ctx, stats, sizes);
}
int vido_map_get_poses_rf(vido_ctx* ctx, float* poses, int cap) { return (ctx && poses) ? trk_get_map_poses_rf(ctx, poses, cap) : VIDO_ERR_ARG; }
int vido_map_get_objects_rf(vido_ctx* ctx, int frame, float* motion, int cap) { return ctx ? trk_get_objects_rf(ctx, frame, motion, cap) : VIDO_ERR_ARG; }
int vido_map_export_full_graph(vido_ctx* ctx, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j, int32_t* e6_kind,
                               float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind, float* obs_xyz,
                               int32_t* tern_p1, int32_t* tern_p2, int32_t* tern_h) {
  if (!ctx || !sizes) return VIDO_ERR_ARG;
  return trk_export_full_graph(ctx, sizes, se3, points, e6_i, e6_j, e6_kind, e6_meas, obs_se3, obs_point, obs_kind, obs_xyz, tern_p1,
                               tern_p2, tern_h);
}

void vido_projopt_default_params(vido_projopt_problem* p, int kind) { if (p) projopt_default_params(p, kind); }
int vido_pose_opt_proj(vido_ctx* ctx, vido_projopt_problem* problems, int nproblems, vido_lm_stats* stats) {
  if (!ctx || !problems || nproblems < 0) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return projopt_host(ctx, problems, nproblems, stats);
}

void vido_inertial_default_params(vido_inertial_problem* p) { if (p) inertial_default_params(p); }
int vido_inertial_opt(vido_ctx* ctx, vido_inertial_problem* p, vido_lm_stats* stats) {
  if (!ctx || !p) return VIDO_ERR_ARG;
  if (p->n_frames < 0 || (p->n_frames >= 2 && (!p->Rwb || !p->twb || !p->velocity || !p->preint || !p->bias_lin))) { ctx->err = "inertial: null input"; return VIDO_ERR_ARG; }
  cudaSetDevice(ctx->device);
  return inertial_opt_host(ctx, p, stats);
}

int vido_track_set_imu(vido_ctx* ctx, const float* Tbc, const float* noise) { return (ctx && ((Tbc && noise) || (!Tbc && !noise))) ? trk_set_imu(ctx, Tbc, noise) : VIDO_ERR_ARG; }
int vido_track_grab_imu(vido_ctx* ctx, const vido_imu_sample* samples, int n, int frames_ahead) {
  if (!ctx || n < 0 || (n > 0 && !samples) || frames_ahead < 0) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_grab_imu(ctx, samples, n, frames_ahead);
}
int vido_track_get_imu_state(vido_ctx* ctx, vido_imu_state* out) { return (ctx && out) ? trk_get_imu_state(ctx, out) : VIDO_ERR_ARG; }
int vido_map_get_imu_frames(vido_ctx* ctx, float* Tcw, float* vel, float* bias, int cap) { return ctx ? trk_get_imu_frames(ctx, Tcw, vel, bias, cap) : VIDO_ERR_ARG; }
int vido_map_apply_scaled_rotation(vido_ctx* ctx, const float* R, float s) {
  if (!ctx || !R) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_apply_scaled_rotation(ctx, R, s);
}

int vido_metric_error(vido_ctx* ctx, const float* cam_pose_gt, int n_gt, int refined, const float* obj_pose_pre,
                      const float* obj_motion_gt, int n_obj, vido_metric* out, float* per_item) {
  if (!ctx || !out || !cam_pose_gt || n_gt < 0 || n_obj < 0 || (n_obj > 0 && (!obj_pose_pre || !obj_motion_gt))) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return trk_metric_error(ctx, cam_pose_gt, n_gt, refined, obj_pose_pre, obj_motion_gt, n_obj, out, per_item);
}

int vido_get_kernel_times(vido_ctx* ctx, double* ms, int64_t* launches, double* ba_alg_bytes) {
  if (!ctx) return VIDO_ERR_ARG;
  for (int k = 0; k < 4; k++) { if (ms) ms[k] = ctx->t_ms[k]; if (launches) launches[k] = ctx->t_n[k]; }
  if (ba_alg_bytes) *ba_alg_bytes = ctx->ba_alg_bytes;
  return VIDO_OK;
}

int vido_imu_preintegrate(vido_ctx* ctx, const vido_imu_sample* samples, int n, const double* t_prev, const double* t_cur,
                          int njobs, const float* bias, const float* noise, vido_imu_preint* out) {
  if (!ctx || !t_prev || !t_cur || !bias || !noise || !out || (n > 0 && !samples)) return VIDO_ERR_ARG;
  cudaSetDevice(ctx->device);
  return imu_preintegrate_host(ctx, samples, n, t_prev, t_cur, njobs, bias, noise, out);
}

}  // extern "C"
