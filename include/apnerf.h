/*
 * apnerf.h -- C-ABI of libapnerf.so: the B200 (sm_100a) render + score hot path of
 * grasp-lyrl/Active-Perception-using-Neural-Radiance-Fields.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; buffers are
 *     contiguous, caller-owned and caller-sized (nothing is allocated behind the boundary
 *     except by the apnerf_*_create calls, which say so);
 *   - bool tensors are 1 byte per element (torch.bool), indices are int64, values float32,
 *     exactly as in the reference's pybind module (perception/nerfacc/nerfacc/cuda/csrc/nerfacc.cpp:100-129);
 *   - `stream` is a cudaStream_t (0 = legacy default stream); calls only enqueue work;
 *   - return value 0 = ok, otherwise a cudaError_t; apnerf_last_error() describes it.  The
 *     Python host maps non-zero to RuntimeError, as TORCH_CHECK does in the reference
 *     (csrc/scan.cu:18-26, csrc/grid.cu:345).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef APNERF_H
#define APNERF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* apnerf_last_error(void);
int apnerf_abi_version(void);

/* ---- kernel (1): ray / AABB slab test and occupancy-grid traversal ------------------- */

/* nerfacc_cuda.ray_aabb_intersect -- csrc/grid.cu:477-519 (kernel :284-313).
 * t_mins/t_maxs [n_rays, n_aabbs] f32, hits [n_rays, n_aabbs] bool. */
int apnerf_ray_aabb_intersect(int n_rays, const float* rays_o, const float* rays_d, int n_aabbs,
                              const float* aabbs, float near_plane, float far_plane, float miss_value,
                              float* t_mins, float* t_maxs, uint8_t* hits, void* stream);

/* One launch of the reference's traverse_grids_kernel -- csrc/grid.cu:68-282, launched from
 * :320-474.  first_pass != 0 only counts (chunk_cnts are written); otherwise segments are
 * written at chunk_starts[ray].  The iv_* / sm_* groups are the fields of the two
 * RaySegmentsSpec outputs (csrc/include/data_spec.hpp:6-14); a NULL *_chunk_cnts disables that
 * group, rays_mask / t_indices / terminate_planes may be NULL (t_indices NULL = identity
 * order, the single-grid case of perception/models/utils.py:886-892). */
int apnerf_traverse_grids(int n_rays, const float* rays_o, const float* rays_d, const uint8_t* rays_mask,
                          int n_grids, int rx, int ry, int rz, const uint8_t* binaries, const float* aabbs,
                          const uint8_t* hits, const float* t_sorted, const int64_t* t_indices,
                          const float* near_planes, const float* far_planes, float step_size,
                          float cone_angle, int traverse_steps_limit, int first_pass,
                          float* iv_vals, int64_t* iv_ray_indices, uint8_t* iv_is_left, uint8_t* iv_is_right,
                          const int64_t* iv_chunk_starts, int64_t* iv_chunk_cnts,
                          float* sm_vals, int64_t* sm_ray_indices, uint8_t* sm_is_valid,
                          const int64_t* sm_chunk_starts, int64_t* sm_chunk_cnts,
                          float* terminate_planes, void* stream);

/* The samples of OccGridEstimator.sampling (estimators/occ_grid.py:117-131: traverse_grids, then
 * t_starts = intervals.vals[is_left], t_ends = intervals.vals[is_right], ray_indices = samples.ray_indices) written
 * directly as packed triples.  Two calls: chunk_starts == NULL counts (chunk_cnts [n_rays] written); then, with
 * chunk_starts = exclusive scan of the counts, fills ray_indices / t_starts / t_ends.  Same kernel arithmetic as
 * apnerf_traverse_grids (bit-identical values), without the interval edges and masks. */
int apnerf_sample_rays(int n_rays, const float* rays_o, const float* rays_d, int n_grids, int rx, int ry, int rz,
                       const uint8_t* binaries, const float* aabbs, const uint8_t* hits, const float* t_sorted,
                       const int64_t* t_indices, const float* near_planes, const float* far_planes,
                       float step_size, float cone_angle, int traverse_steps_limit,
                       const int64_t* chunk_starts, int64_t* chunk_cnts, int64_t* ray_indices, float* t_starts,
                       float* t_ends, void* stream);

/* chunk_cnts -> chunk_starts (+ device-side total): RaySegmentsSpec::memalloc_data_from_chunk /
 * compute_chunk_start -- csrc/include/data_spec.hpp:86-106.  scratch: apnerf_scan_scratch_elems(n)
 * int64 elements. */
int apnerf_exclusive_scan_i64(long long n, const int64_t* in, int64_t* out, int64_t* total,
                              int64_t* scratch, void* stream);
long long apnerf_scan_scratch_elems(long long n);

/* ---- kernel (4): packed scans, transmittance weights, accumulation ------------------- */

/* nerfacc_cuda.inclusive_sum / exclusive_sum -- csrc/scan.cu:9-125 (backward = reverse scan). */
int apnerf_packed_sum(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                      long long n_edges, const float* inputs, float* outputs, int inclusive,
                      int normalize, int backward, void* stream);

/* render_weight_from_density / render_transmittance_from_density -- nerfacc/volrend.py:212-267,
 * 315-365, fused (sigma*dt, exclusive sum, exp, alpha, weight).  prefix_trans and any output
 * may be NULL. */
int apnerf_weights_from_density(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                                long long n_samples, const float* t_starts, const float* t_ends,
                                const float* sigmas, const float* prefix_trans, float* weights,
                                float* trans, float* alphas, void* stream);
/* autograd of the above (reference: _ExclusiveSum.backward, nerfacc/scan.py:206-229, plus the
 * ATen elementwise graph of volrend.py:259-267). */
int apnerf_weights_from_density_bwd(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                                    long long n_samples, const float* t_starts, const float* t_ends,
                                    const float* sigmas, const float* prefix_trans,
                                    const float* g_weights, const float* g_trans, const float* g_alphas,
                                    float* g_sigmas, float* g_prefix, void* stream);

/* accumulate_along_rays / accumulate_along_rays_ -- nerfacc/volrend.py:486-576:
 * outputs[ray_indices[i], :] += weights[i] * values[i, :] (values NULL: D == 1, += weights). */
int apnerf_accumulate_along_rays(long long n_samples, int D, const float* weights, const float* values,
                                 const int64_t* ray_indices, float* outputs, void* stream);
int apnerf_accumulate_along_rays_bwd(long long n_samples, int D, const float* weights,
                                     const float* values, const int64_t* ray_indices,
                                     const float* g_outputs, float* g_weights, float* g_values,
                                     void* stream);

/* pack_info -- nerfacc/pack.py:10-49.  packed_info [n_rays, 2] int64 (start, count).
 * scratch: 2 * n_rays + apnerf_scan_scratch_elems(n_rays) int64 elements. */
int apnerf_pack_info(long long n_samples, const int64_t* ray_indices, int n_rays, int64_t* packed_info,
                     int64_t* scratch, void* stream);

/* ---- kernels (2) + (3): hash-grid encode and the fused field (hash grid + MLPs) -------- */

/* Stand-alone multiresolution hash-grid encode (tcnn HashGrid part of
 * tcnn.NetworkWithInputEncoding, perception/models/radiance_fields/ngp.py:123-133), used for the
 * bit-exact cell-index parity check.  x01 [n,3] f32 in aabb-normalised coordinates;
 * meta_host: HOST array [n_levels][5] u32 {scale (f32 bits), resolution, entries, first entry,
 * hashed}; table fp16 [entries,4]; out_enc fp16 [n, n_levels*4] (NULL ok); out_idx u32
 * [n, n_levels, 8] table-entry indices (NULL ok). */
int apnerf_hashgrid_encode(long long n, const float* x01, int n_levels, const uint32_t* meta_host,
                           const void* table, void* out_enc, uint32_t* out_idx, void* stream);

/* NGPRadianceField.query_density / forward -- perception/models/radiance_fields/ngp.py:171-238
 * (tcnn modules :107-169), one fused kernel.  Sample points are either positions(+directions)
 * [n,3] or ray samples (ray_idx i32 [n], t_starts/t_ends [n], rays_o/rays_d [n_rays,3]) whose
 * midpoints are formed in-kernel exactly as perception/models/utils.py:833-836.  n_dev
 * (device int32, may be NULL) overrides n for device-driven loops; max_tiles then bounds the
 * grid.  aabb_host: HOST float[6].  weights: fp16 blob of apnerf_field_weight_bytes() bytes in
 * UMMA K-major layout (see csrc/field.cuh).  Outputs: density [n]; rgb(i,c) at
 * rgb[c*rgb_ch + i*rgb_row]; sem(i,c) likewise for c < n_sem; feat fp16 [n,15] (NULL ok).
 * packed (NULL ok): instead of density/rgb/sem, write one 80-byte row of 40 fp16 per sample
 * {density logit (-inf outside the aabb), 3 rgb logits, sigma = exp(logit - 1) as fp32, 2 pad, 32
 * semantic logits} -- the raw
 * fp16 network outputs the fused renderer's compositing kernel activates itself. */
int apnerf_field_forward(long long n, const int* n_dev, const float* positions, const float* directions,
                         const int* ray_idx, const float* t_starts, const float* t_ends,
                         const float* rays_o, const float* rays_d, const float* aabb_host,
                         int n_levels, const uint32_t* meta_host, const void* table,
                         const void* weights, float* density, float* rgb, long long rgb_row,
                         long long rgb_ch, float* sem, long long sem_row, long long sem_ch, int n_sem,
                         void* feat, void* packed, int density_only, long long max_tiles, void* stream);
int apnerf_field_weight_bytes(void);

/* Training side of the fused MLPs (tcnn FullyFusedMLP forward/backward behind loss.backward(),
 * scripts/pipeline.py:518; modules ngp.py:123-169).
 *   apnerf_field_forward_train: the kernel of apnerf_field_forward with RAW outputs (fp16 network outputs upcast:
 *       density logit [n], rgb logits [n,3], semantic logits [n,n_sem]) and the activations the backward
 *       needs saved as row-major fp16: enc [n,64], h1 [n,128], h2 [n,128], xh [n,32] (16 SH | 15 geo | 1.0),
 *       xs [n,16] (15 geo | 1.0), hh1, hh2, hs1, hs2 [n,64]; rows are save_stride fp16 elements apart, so the
 *       nine matrices can be column slices of one [n, 624] matrix (then all weight gradients are ONE GEMM).
 *   apnerf_field_backward: d_dens [n], d_rgb [n,3], d_sem [n,n_sem] (fp32, NULL ok) -> the chain of dX = dY.W
 *       products on the tensor cores (csrc/field_bwd_kernel.cuh): g_* = gradients w.r.t. every layer's
 *       pre-activation output x loss_scale in fp16 (g_base [n,16] for the base output; g_out_h [n,16] /
 *       g_out_s [n,32] are the padded incoming gradients; rows g_stride elements apart), d_enc [n,64] fp32
 *       unscaled (feeds apnerf_hashgrid_encode_bwd).  weights_t: the blob of apnerf_field_forward with every
 *       matrix transposed ([in,out] in the same UMMA layout, same offsets).  Weight gradients are the plain
 *       GEMMs dW = g^T . x the caller runs with a library (as tcnn does). */
int apnerf_field_forward_train(long long n, const float* positions, const float* directions,
                               const float* aabb_host, int n_levels, const uint32_t* meta_host,
                               const void* table, const void* weights, float* dens_logit,
                               float* rgb_logit, float* sem_logit, int n_sem, long long save_stride,
                               void* save_enc,
                               void* save_h1, void* save_h2, void* save_xh, void* save_xs,
                               void* save_hh1, void* save_hh2, void* save_hs1, void* save_hs2,
                               void* stream);
int apnerf_field_backward(long long n, const float* d_dens, const float* d_rgb, const float* d_sem,
                          int n_sem, long long act_stride, const void* h1, const void* h2,
                          const void* hh1, const void* hh2, const void* hs1, const void* hs2,
                          const void* weights_t, float loss_scale, long long g_stride, void* g_out_h,
                          void* g_out_s, void* g_hh2, void* g_hs2, void* g_hh1, void* g_hs1, void* g_base,
                          void* g_h2, void* g_h1, float* d_enc, void* stream);

/* Weight gradients of the three MLPs, dW = dY^T . X, as tcgen05 split-K products with all nine accumulators in TMEM
 * (csrc/field_wgrad_kernel.cuh) -- the other half of tcnn's FullyFusedMLP backward behind loss.backward()
 * (scripts/pipeline.py:518; ngp.py:123-169).  G [n_pad, 576] / X [n_pad, 624]: the fp16 matrices written by
 * apnerf_field_backward (g_stride 576) and apnerf_field_forward_train (save_stride 624), rows >= n zero, n_pad a
 * multiple of 32.  d_base / d_head / d_sem (fp32, += ; d_sem may be NULL): flat gradients of [W1|W2|W3],
 * [WH1|WH2|WH3], [WS1|WS2|WS3] in tcnn's parameter order; sem_out_rows = rows of WS3 (16 or 32). */
int apnerf_field_wgrad(long long n, const void* G, const void* X, float loss_scale, float* d_base, float* d_head,
                       float* d_sem, int sem_out_rows, void* stream);

/* OccGridEstimator._update -- perception/nerfacc/nerfacc/estimators/occ_grid.py:377-437, the per-level body
 * fused into the field kernel: x = level_aabb_lo + ((grid_coords(cell) + jitter) / res) * extent; occ =
 * query_density(x) * occ_scale (the pipeline's occ_eval_fn, scripts/pipeline.py:470-475); occs_new[cell] =
 * maximum(occs_old[cell] * ema_decay, occ), keeping occs_old[cell] where that is NaN (:405, :430-434).
 * cell_ids i64 [n] (cells may repeat: occs_old must be a snapshot, the survivor is unspecified like the
 * reference's indexed assignment), jitter f32 [n,3], level_aabb_host HOST float[6]. */
int apnerf_occ_update(long long n, const long long* cell_ids, const float* jitter,
                      const float* level_aabb_host, int rx, int ry, int rz, const float* occs_old,
                      float* occs_new, float occ_scale, float ema_decay, const float* aabb_host,
                      int n_levels, const uint32_t* meta_host, const void* table, const void* weights,
                      void* stream);

/* Training side of the hash grid (tcnn HashGrid backward, reached from loss.backward() at
 * scripts/pipeline.py:518): d_table [entries,4] fp32 += sum over samples/corners of
 * w_corner * d_enc [n, n_levels*4] fp32 (vector atomics). */
int apnerf_hashgrid_encode_bwd(long long n, const float* x01, int n_levels, const uint32_t* meta_host,
                               const float* d_enc, float* d_table, void* stream);
/* tcnn SphericalHarmonics degree 4 (ngp.py:108-121): dirs [n,3] f32 unit vectors -> fp16 [n,16]. */
int apnerf_sh4(long long n, const float* dirs, void* out, void* stream);


/* ---- the device-driven test-mode renderer + scorer (kernels 1, 4, 5 fused) -------------
 * One "call" = one view through one ensemble member (rays_per_call rays); a batch of calls
 * advances in lock step with the reference's per-call marching schedule
 * (perception/models/utils.py:896-1009) decided entirely on the device.
 * Per-ray state: float [n_state = 9 + n_sem][n_rays] structure-of-arrays
 *   (0-2 rgb, 3 opacity, 4 depth, 5-7 rgb_var, 8 depth_var, 9.. semantic logits).
 * counters: int[8] = {live rays now, live rays collected for the next iteration, samples (rows)
 *                     emitted this iteration, iterations that had work, compaction ticket,
 *                     sample-list overflow flag, real samples this iteration, compaction
 *                     generation}. */

/* Dataset.generate_image_rays (+ the rounded-linspace subsample) --
 * perception/data_proc/habitat_to_data.py:274-301, 461-467.  c2w [n_views,3,4] f32;
 * keep_idx int32 [n_keep] pixel indices (NULL = all width*height pixels, n_keep must match). */
int apnerf_generate_rays(int n_views, const float* c2w, int width, int height, float focal, int n_keep,
                         const int* keep_idx, float* rays_o, float* rays_d, void* stream);

/* utils.py:860-892: zero the state, ray/aabb intersection, every ray live. */
int apnerf_render_init(int n_rays, int rays_per_call, const float* rays_o, const float* rays_d, int rx,
                       int ry, int rz, const uint8_t* binaries, const float* aabbs, float near_plane,
                       int n_state, float* state, float* t_min, float* t_max, uint8_t* hit, float* near,
                       int* alive, int* n_alive_acc, int* iter_samples, int* total_samples, int n_calls,
                       int* counters, uint32_t* occ_bits, void* stream);
/* utils.py:896-903: per call n = max(min(R // n_alive, 64), min_samples), iter_samples += n.
 * call_rows (nullable) int32 [n_calls]: += n * n_alive per call, i.e. the sample rows the call sends through the
 * field (an upper bound: a ray's last iteration may emit fewer than n) -- the scheduler's per-view cost. */
int apnerf_render_schedule(int n_calls, int rays_per_call, int max_samples, int min_samples,
                           int* n_alive_acc, int* n_samp, int* iter_samples, int* counters, int* call_rows,
                           void* stream);
/* utils.py:906-929: limited traversal of every live ray from its last terminate plane;
 * emits the compact sample list (s_ray, s_ts, s_te), counters[2] and per-entry (base, count), plus
 * s_x [rows] float4: every sample's midpoint o + d (t_s + t_e) / 2 (utils.py:833-836) normalised by the
 * FIELD's aabb (ngp.py:175-176; field_aabb_host: 6 floats on the host) for apnerf_field_forward_rows. */
int apnerf_render_march(int max_live, int rays_per_call, const int* alive, const int* n_samp,
                        const float* rays_o, const float* rays_d, int rx, int ry, int rz,
                        const uint8_t* binaries, const float* aabbs, const float* t_min, const float* t_max,
                        const uint8_t* hit, float* near, float far_plane, float step_size, float cone_angle,
                        int* entry_base, int* entry_cnt, int* s_ray, float* s_ts, float* s_te,
                        const float* field_aabb_host, void* s_x, int* counters, const uint32_t* occ_bits,
                        void* stream);
/* utils.py:937-1009: weights with prefix transmittance, alpha_thre filter, accumulation, variance
 * terms, next ray mask, live-list compaction.  rows: the field kernel's packed fp16 rows
 * [s][40]; density = exp(logit - 1) (ngp.py:79), rgb = sigmoid(logit) (ngp.py:211-212).
 * ray_counts (may be NULL): int32 [2][n_rays], += per ray the samples evaluated / composited after the
 * alpha_thre filter -- the discrete decisions of a render, which the parity tests compare with the oracle's. */
int apnerf_render_composite(int max_live, int n_rays, int rays_per_call, int n_sem,
                            const int* alive, const int* entry_base, const int* entry_cnt,
                            const float* s_ts, const float* s_te, const void* rows, float* state, float alpha_thre, float opc_thre,
                            const int* n_samp, const int* iter_samples, int max_samples, int* alive_next,
                            int* n_alive_acc, int* total_samples, int* counters, int probabilistic,
                            int* ray_counts, void* stream);

/* The renderer's field query (kernels 2 + 3) on the marcher's sample rows: s_ray [rows] (ray id, -1 = padding row),
 * s_x [rows] float4 (aabb-normalised sample point written by apnerf_render_march), rays_d [n_rays,3] -> packed:
 * one 80-byte row of raw fp16 network outputs per sample (layout under apnerf_field_forward) for
 * apnerf_render_composite.  *n_rows_dev rows, clamped to max_tiles * 128.  Replaces, per marching iteration,
 * radiance_field(positions, t_dirs) at perception/models/utils.py:931-935 (ngp.py:222-238 + tcnn). */
int apnerf_field_forward_rows(const int* n_rows_dev, long long max_tiles, const int* s_ray, const void* s_x,
                              const float* rays_d, const float* aabb_host, int n_levels,
                              const uint32_t* meta_host, const void* table, const void* weights, void* packed,
                              void* stream);

/* Fused form of one marching iteration (kernels 1 + 2 + 3 + 4 in three launches): the compositor
 * of utils.py:937-1009 runs inside the field kernel's epilogue, so per-sample network outputs
 * never leave the SM.
 *   apnerf_render_march_tiles : like apnerf_render_march, but every ray's samples are placed inside
 *       one 128-row tile (s_ray = -1 marks padding rows, s_cnt = #samples at a ray's first row);
 *       counters[2] = rows reserved (the field kernel clamps it to s_cap); counters[5] is set if
 *       the s_cap rows did not suffice; keep_flag[ray] cleared for every live ray.
 *   apnerf_field_forward_fused: hash grid + MLPs + compositing into `state`; sets keep_flag.
 *   apnerf_render_compact     : ordered compaction of the live list by keep_flag (list stays sorted
 *       by ray id), per-call live counts; chain: ceil(n_rays/2048)+1 uint64 scratch (never needs
 *       clearing: entries carry the generation counters[7] that the schedule kernel bumps). */
int apnerf_render_march_tiles(int max_live, int rays_per_call, const int* alive, const int* n_samp,
                              const float* rays_o, const float* rays_d, int rx, int ry, int rz,
                              const uint8_t* binaries, const float* aabbs, const float* t_min,
                              const float* t_max, const uint8_t* hit, float* near, float far_plane,
                              float step_size, float cone_angle, int* s_ray, uint8_t* s_cnt, float* s_ts,
                              float* s_te, const float* field_aabb_host, void* s_x, uint8_t* keep_flag, int s_cap,
                              int* counters, const uint32_t* occ_bits, void* stream);
int apnerf_field_forward_fused(const int* n_rows_dev, long long max_tiles, const int* s_ray,
                               const uint8_t* s_cnt, const float* s_ts, const float* s_te, const void* s_x,
                               const float* rays_d, const float* aabb_host,
                               int n_levels, const uint32_t* meta_host, const void* table,
                               const void* weights, int n_sem, float* state, int n_rays_total,
                               int rays_per_call, float alpha_thre, float opc_thre, const int* n_samp,
                               const int* iter_samples, int max_samples, uint8_t* keep_flag,
                               int* total_samples, int probabilistic, int* ray_counts, void* stream);
int apnerf_render_compact(int max_live, int rays_per_call, const int* alive, const uint8_t* keep_flag,
                          int* alive_next, int* n_alive_acc, void* chain, int* counters, void* stream);
/* utils.py:1012-1032: background, depth normalisation, [n_rays, D] outputs (any may be NULL). */
int apnerf_render_finalize(int n_rays, int n_sem, const float* state, float bkgd_r, float bkgd_g,
                           float bkgd_b, float* rgb, float* rgb_var, float* opacity, float* depth,
                           float* depth_var, float* sem, void* stream);
/* ActiveNeRFMapper.probablistic_uncertainty arithmetic -- scripts/pipeline.py:727-781.
 * state0..3: per-member state arrays (unused ones NULL); view_traj int32 [n_views] maps a view to
 * its trajectory (-1 = skip); sums double [n_traj][4] += per-pixel (rgb, depth, sem, occ) terms. */
int apnerf_score_views(int n_members, const float* state0, const float* state1, const float* state2,
                       const float* state3, int n_rays, int rays_per_view, int n_sem,
                       const int* view_traj, int n_traj, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APNERF_H */
