// C-ABI entry points of the radiance field (kernels 2 + 3): the fused hash-grid + MLP forward
// and a stand-alone hash-grid encode used for the bit-exact cell-index parity tests.
// Replaces tcnn.NetworkWithInputEncoding / tcnn.Network / tcnn.Encoding as used by
// NGPRadianceField (perception/models/radiance_fields/ngp.py:107-238).
#include "field_kernel.cuh"
#include "field_bwd_kernel.cuh"
#include "field_wgrad_kernel.cuh"

namespace apnerf {

__global__ void __launch_bounds__(256) hashgrid_encode_kernel(long long n, const float* __restrict__ x01,
                                                              const uint2* __restrict__ table, HashGridMeta meta,
                                                              __half* __restrict__ out_enc,
                                                              uint32_t* __restrict__ out_idx) {
  const long long total = n * meta.n_levels;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)blockDim.x * gridDim.x) {
    const long long s = t / meta.n_levels;
    const int l = (int)(t % meta.n_levels);
    const float x[3] = {x01[3 * s], x01[3 * s + 1], x01[3 * s + 2]};
    if (out_enc) {
      const uint2 f = encode_level(meta, l, x, table);
      *reinterpret_cast<uint2*>(out_enc + s * (meta.n_levels * FEATS) + l * FEATS) = f;
    }
    if (out_idx) {
      uint32_t cell[3];
      float w[3];
      level_cell(meta, l, x, cell, w);
      for (int c = 0; c < 8; ++c) out_idx[(s * meta.n_levels + l) * 8 + c] = corner_index(meta, l, cell, c);
    }
  }
}


// Backward of the hash-grid interpolation w.r.t. the table: the encoding is linear in the table,
// d table[corner] += w_corner * d enc[sample, level, :].  One thread per (sample, level); one
// 16-byte vector atomic per corner into the fp32 gradient of the flat parameter vector
// (tcnn accumulates grid gradients with atomics as well).
__global__ void __launch_bounds__(256) hashgrid_encode_bwd_kernel(long long n, const float* __restrict__ x01,
                                                                  HashGridMeta meta,
                                                                  const float* __restrict__ d_enc,  // [n, L*4]
                                                                  float* __restrict__ d_table) {    // [entries, 4]
  const long long total = n * meta.n_levels;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)blockDim.x * gridDim.x) {
    const long long s = t / meta.n_levels;
    const int l = (int)(t % meta.n_levels);
    const float x[3] = {x01[3 * s], x01[3 * s + 1], x01[3 * s + 2]};
    const float4 g = *reinterpret_cast<const float4*>(d_enc + s * (meta.n_levels * FEATS) + l * FEATS);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    uint32_t cell[3];
    float w[3];
    level_cell(meta, l, x, cell, w);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float wt = corner_weight(w, c);
      float4* dst = reinterpret_cast<float4*>(d_table) + corner_index(meta, l, cell, c);
      atomicAdd(dst, make_float4(wt * g.x, wt * g.y, wt * g.z, wt * g.w));
    }
  }
}

__global__ void __launch_bounds__(256) sh4_kernel(long long n, const float* __restrict__ dirs, __half* __restrict__ out) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n; s += (long long)blockDim.x * gridDim.x) {
    const float d[3] = {dirs[3 * s], dirs[3 * s + 1], dirs[3 * s + 2]};
    float o[16];
    sh4(d, o);
    __align__(16) __half h[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) h[i] = __float2half_rn(o[i]);
    reinterpret_cast<uint4*>(out + s * 16)[0] = reinterpret_cast<const uint4*>(h)[0];
    reinterpret_cast<uint4*>(out + s * 16)[1] = reinterpret_cast<const uint4*>(h)[1];
  }
}

static int fill_meta(HashGridMeta& m, int n_levels, const uint32_t* meta_host) {
  if (n_levels < 1 || n_levels > MAX_LEVELS) return 1;
  m.n_levels = n_levels;
  for (int l = 0; l < MAX_LEVELS; ++l) {
    if (l < n_levels) {
      uint32_t bits = meta_host[5 * l + 0];
      float sc;
      memcpy(&sc, &bits, 4);
      m.scale[l] = sc;
      m.res[l] = meta_host[5 * l + 1];
      m.size[l] = meta_host[5 * l + 2];
      m.offset[l] = meta_host[5 * l + 3];
      m.hashed[l] = meta_host[5 * l + 4];
      if (m.hashed[l] && (m.size[l] & (m.size[l] - 1))) return 2;
    } else {
      m.scale[l] = 0.f, m.res[l] = 1, m.size[l] = 1, m.offset[l] = 0, m.hashed[l] = 0;
    }
  }
  return 0;
}

// Opt-in to > 48 KB of dynamic shared memory once per (kernel instantiation, device): function attributes are
// per device, and a process may drive several GPUs.
template <int MODE>
static int launch_field(const FieldIO& io_in, const HashGridMeta& m, const FieldConst& fc, int grid, int smem,
                        cudaStream_t stream, const char* name) {
  // APNERF_FIELD_PAIR_X=1: x-neighbour pairs through one 16-byte load (field.cuh: ldg_entry_pair).  An experiment
  // switch, off by default: measured 14 % SLOWER (4.46 vs 5.17 G rows/s, profiles/r02_field_kernel.md).
  static const int pair_x = getenv("APNERF_FIELD_PAIR_X") ? atoi(getenv("APNERF_FIELD_PAIR_X")) : 0;
  FieldIO io = io_in;
  io.pair_x = pair_x;
  static unsigned long long attr_done = 0ull;  // bit d: set on device d
  int dev = 0;
  APNERF_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_done >> dev) & 1ull)) {
    APNERF_CUDA(cudaFuncSetAttribute(field_forward_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_SMEM));
    if (dev < 64) attr_done |= 1ull << dev;
  }
  field_forward_kernel<MODE><<<grid, FIELD_THREADS, smem, stream>>>(io, m, fc);
  APNERF_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace apnerf

using namespace apnerf;

// meta_host: HOST array [n_levels][5] u32 = {scale (f32 bits), resolution, size, offset, hashed}.
APNERF_API int apnerf_hashgrid_encode(long long n, const float* x01, int n_levels, const uint32_t* meta_host,
                                      const void* table, void* out_enc, uint32_t* out_idx, void* stream) {
  if (n == 0) return 0;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "hashgrid_encode: bad level table");
  hashgrid_encode_kernel<<<grid_for(n * n_levels, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n, x01, (const uint2*)table, m, (__half*)out_enc, out_idx);
  APNERF_CHECK_LAUNCH("hashgrid_encode_kernel");
  return 0;
}


// d_table [entries, 4] fp32 += d enc / d table ^T * d_enc  (d_enc fp32 [n, n_levels*4]).
APNERF_API int apnerf_hashgrid_encode_bwd(long long n, const float* x01, int n_levels, const uint32_t* meta_host,
                                          const float* d_enc, float* d_table, void* stream) {
  if (n == 0) return 0;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "hashgrid_encode_bwd: bad level table");
  hashgrid_encode_bwd_kernel<<<grid_for(n * n_levels, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, x01, m, d_enc,
                                                                                               d_table);
  APNERF_CHECK_LAUNCH("hashgrid_encode_bwd_kernel");
  return 0;
}

// SH degree 4 of unit directions, fp16 [n, 16] (tcnn SphericalHarmonics as used at ngp.py:108-121).
APNERF_API int apnerf_sh4(long long n, const float* dirs, void* out, void* stream) {
  if (n == 0) return 0;
  sh4_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(n, dirs, (__half*)out);
  APNERF_CHECK_LAUNCH("sh4_kernel");
  return 0;
}

// Fused field query.  Exactly one of {positions(+directions)} / {ray_idx, t_starts, t_ends,
// rays_o, rays_d} describes the sample points.  n_dev (optional, device int32) overrides n.
// Outputs: density [n]; rgb / sem element (i, c) at base[c * ch_stride + i * row_stride]; or, when
// `packed` is given, one 80-byte row of raw fp16 network outputs per sample (fused renderer).
APNERF_API int apnerf_field_forward(long long n, const int* n_dev, const float* positions, const float* directions,
                                    const int* ray_idx, const float* t_starts, const float* t_ends,
                                    const float* rays_o, const float* rays_d, const float* aabb_host,
                                    int n_levels, const uint32_t* meta_host, const void* table,
                                    const void* weights, float* density, float* rgb, long long rgb_row,
                                    long long rgb_ch, float* sem, long long sem_row, long long sem_ch, int n_sem,
                                    void* feat, void* packed, int density_only, long long max_tiles, void* stream) {
  if (n == 0 && n_dev == nullptr) return 0;
  APNERF_REQUIRE(positions != nullptr || ray_idx != nullptr, "field_forward: no sample points given");
  APNERF_REQUIRE(density_only || positions == nullptr || directions != nullptr, "field_forward: directions missing");
  APNERF_REQUIRE(n_sem >= 0 && n_sem <= SEM_OUT, "field_forward: at most 32 semantic classes");
  FieldIO io;
  io.n = n, io.n_dev = n_dev, io.positions = positions, io.directions = directions, io.ray_idx = ray_idx;
  io.t_starts = t_starts, io.t_ends = t_ends, io.rays_o = rays_o, io.rays_d = rays_d, io.x01 = nullptr;
  io.table = (const uint2*)table, io.weights = (const uint4*)weights;
  io.density = density, io.rgb = rgb, io.rgb_row = rgb_row, io.rgb_ch = rgb_ch;
  io.sem = sem, io.sem_row = sem_row, io.sem_ch = sem_ch, io.feat = (__half*)feat, io.n_sem = sem ? n_sem : 0;
  io.density_only = density_only;
  io.packed = (uint4*)packed;
  io.state = nullptr, io.n_rays_total = 0, io.rays_per_call = 1, io.alpha_thre = 0.f, io.opc_thre = 0.f;
  io.n_samp = nullptr, io.iter_samples = nullptr, io.max_samples = 0, io.s_cnt = nullptr, io.keep_flag = nullptr;
  io.total_samples = nullptr, io.probabilistic = 0, io.ray_counts = nullptr;
  io.cell_ids = nullptr, io.jitter = nullptr, io.occs_old = nullptr, io.occs_new = nullptr;
  io.save_enc = io.save_h1 = io.save_h2 = io.save_xh = io.save_xs = nullptr;
  io.save_hh1 = io.save_hh2 = io.save_hs1 = io.save_hs2 = nullptr;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "field_forward: bad level table");
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = aabb_host[i];
  long long tiles = n_dev ? max_tiles : (n + TILE_M - 1) / TILE_M;
  if (tiles < 1) tiles = 1;
  const int sms = apnerf_num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  return launch_field<0>(io, m, fc, grid, FIELD_SMEM_MIN, (cudaStream_t)stream, "field_forward_kernel");
}

APNERF_API int apnerf_field_weight_bytes(void) { return W_BYTES; }

// Training forward: the same kernel with raw outputs (fp16 logits upcast) and every activation the backward
// needs written out as row-major fp16.  save: [enc 64 | h1 128 | h2 128 | xh 32 | xs 16 | hh1 64 | hh2 64 |
// hs1 64 | hs2 64] = nine pointers.
APNERF_API int apnerf_field_forward_train(long long n, const float* positions, const float* directions,
                                          const float* aabb_host, int n_levels, const uint32_t* meta_host,
                                          const void* table, const void* weights, float* dens_logit,
                                          float* rgb_logit, float* sem_logit, int n_sem, long long save_stride,
                                          void* save_enc,
                                          void* save_h1, void* save_h2, void* save_xh, void* save_xs,
                                          void* save_hh1, void* save_hh2, void* save_hs1, void* save_hs2,
                                          void* stream) {
  if (n == 0) return 0;
  APNERF_REQUIRE(positions && directions && dens_logit && rgb_logit, "field_forward_train: null buffer");
  APNERF_REQUIRE(save_enc && save_h1 && save_h2 && save_xh && save_xs && save_hh1 && save_hh2 && save_hs1 && save_hs2,
                 "field_forward_train: all nine activation buffers are required");
  APNERF_REQUIRE(n_sem >= 0 && n_sem <= SEM_OUT, "field_forward_train: at most 32 semantic classes");
  FieldIO io;
  memset(&io, 0, sizeof(io));
  io.n = n, io.positions = positions, io.directions = directions;
  io.table = (const uint2*)table, io.weights = (const uint4*)weights;
  io.density = dens_logit, io.rgb = rgb_logit, io.rgb_row = 3, io.rgb_ch = 1;
  io.sem = sem_logit, io.sem_row = n_sem, io.sem_ch = 1, io.n_sem = sem_logit ? n_sem : 0;
  io.rays_per_call = 1;
  APNERF_REQUIRE(save_stride >= 128 && save_stride % 8 == 0, "field_forward_train: save_stride must be a multiple of 8, >= 128");
  io.save_stride = save_stride;
  io.save_enc = (__half*)save_enc, io.save_h1 = (__half*)save_h1, io.save_h2 = (__half*)save_h2;
  io.save_xh = (__half*)save_xh, io.save_xs = (__half*)save_xs, io.save_hh1 = (__half*)save_hh1;
  io.save_hh2 = (__half*)save_hh2, io.save_hs1 = (__half*)save_hs1, io.save_hs2 = (__half*)save_hs2;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "field_forward_train: bad level table");
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = aabb_host[i];
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const int sms = apnerf_num_sms();
  return launch_field<1>(io, m, fc, (int)(tiles < sms ? tiles : sms), FIELD_SMEM_MIN, (cudaStream_t)stream,
                         "field_forward_kernel(train)");
}

// Backward of the three MLPs (csrc/field_bwd_kernel.cuh).  weights_t: the blob of apnerf_field_forward with
// every matrix transposed.  Outputs: g_* = activation gradients x loss_scale (fp16), d_enc fp32 unscaled.
APNERF_API int apnerf_field_backward(long long n, const float* d_dens, const float* d_rgb, const float* d_sem,
                                     int n_sem, long long act_stride, const void* h1, const void* h2,
                                     const void* hh1, const void* hh2, const void* hs1, const void* hs2,
                                     const void* weights_t, float loss_scale, long long g_stride, void* g_out_h,
                                     void* g_out_s, void* g_hh2, void* g_hs2, void* g_hh1, void* g_hs1, void* g_base,
                                     void* g_h2, void* g_h1, float* d_enc, void* stream) {
  if (n == 0) return 0;
  APNERF_REQUIRE(d_dens && d_rgb && h1 && h2 && hh1 && hh2 && hs1 && hs2 && weights_t, "field_backward: null input");
  APNERF_REQUIRE(g_out_h && g_out_s && g_hh2 && g_hs2 && g_hh1 && g_hs1 && g_base && g_h2 && g_h1 && d_enc,
                 "field_backward: null output");
  APNERF_REQUIRE(act_stride >= 128 && act_stride % 8 == 0 && g_stride >= 128 && g_stride % 8 == 0,
                 "field_backward: row strides must be multiples of 8, >= 128");
  APNERF_REQUIRE(n_sem >= 0 && n_sem <= SEM_OUT && loss_scale > 0.f, "field_backward: bad n_sem / loss_scale");
  static unsigned long long attr_done = 0ull;  // bit d: opted in on device d (function attributes are per device)
  int dev = 0;
  APNERF_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_done >> dev) & 1ull)) {
    APNERF_CUDA(cudaFuncSetAttribute(field_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    if (dev < 64) attr_done |= 1ull << dev;
  }
  FieldBwdIO io;
  io.n = n, io.d_dens = d_dens, io.d_rgb = d_rgb, io.d_sem = d_sem, io.n_sem = d_sem ? n_sem : 0;
  io.h1 = (const __half*)h1, io.h2 = (const __half*)h2, io.hh1 = (const __half*)hh1, io.hh2 = (const __half*)hh2;
  io.hs1 = (const __half*)hs1, io.hs2 = (const __half*)hs2, io.weights_t = (const uint4*)weights_t;
  io.loss_scale = loss_scale, io.act_stride = act_stride, io.g_stride = g_stride;
  io.g_out_h = (__half*)g_out_h, io.g_out_s = (__half*)g_out_s;
  io.g_hh2 = (__half*)g_hh2, io.g_hs2 = (__half*)g_hs2, io.g_hh1 = (__half*)g_hh1, io.g_hs1 = (__half*)g_hs1;
  io.g_base = (__half*)g_base, io.g_h2 = (__half*)g_h2, io.g_h1 = (__half*)g_h1, io.d_enc = d_enc;
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const long long cap = 2LL * apnerf_num_sms();  // two 112 KB CTAs fit one SM: one's epilogue hides the other's MMAs
  field_backward_kernel<<<(int)(tiles < cap ? tiles : cap), BWD_THREADS, BWD_SMEM, (cudaStream_t)stream>>>(io);
  APNERF_CHECK_LAUNCH("field_backward_kernel");
  return 0;
}

// Weight gradients of the three MLPs (csrc/field_wgrad_kernel.cuh): d_* += G^T . X / loss_scale for the nine layers.
// G [n_pad, 576], X [n_pad, 624] fp16 row-major with rows >= n zero and n_pad a multiple of 32; d_base / d_head /
// d_sem: the flat fp32 gradients of [W1|W2|W3], [WH1|WH2|WH3], [WS1|WS2|WS3] (d_sem may be NULL).
APNERF_API int apnerf_field_wgrad(long long n, const void* G, const void* X, float loss_scale, float* d_base,
                                  float* d_head, float* d_sem, int sem_out_rows, void* stream) {
  if (n == 0) return 0;
  APNERF_REQUIRE(G && X && d_base && d_head && loss_scale > 0.f, "field_wgrad: null buffer / bad loss scale");
  APNERF_REQUIRE(sem_out_rows == 16 || sem_out_rows == 32 || d_sem == nullptr, "field_wgrad: sem_out_rows must be 16 or 32");
  static unsigned long long attr_done = 0ull;
  int dev = 0;
  APNERF_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_done >> dev) & 1ull)) {
    APNERF_CUDA(cudaFuncSetAttribute(field_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    if (dev < 64) attr_done |= 1ull << dev;
  }
  WgradIO io;
  io.n = n, io.G = (const __half*)G, io.X = (const __half*)X, io.scale = 1.0f / loss_scale;
  io.d_base = d_base, io.d_head = d_head, io.d_sem = d_sem, io.sem_out_rows = sem_out_rows;
  const long long stages = (n + WG_STAGE_ROWS - 1) / WG_STAGE_ROWS;
  const int sms = apnerf_num_sms();
  field_wgrad_kernel<<<(int)(stages < sms ? stages : sms), WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(io);
  APNERF_CHECK_LAUNCH("field_wgrad_kernel");
  return 0;
}

// OccGridEstimator._update for one grid level in one launch: jittered cell -> density -> EMA-max.
APNERF_API int apnerf_occ_update(long long n, const long long* cell_ids, const float* jitter,
                                 const float* level_aabb_host, int rx, int ry, int rz, const float* occs_old,
                                 float* occs_new, float occ_scale, float ema_decay, const float* aabb_host,
                                 int n_levels, const uint32_t* meta_host, const void* table, const void* weights,
                                 void* stream) {
  if (n == 0) return 0;
  APNERF_REQUIRE(cell_ids && jitter && occs_old && occs_new, "occ_update: null buffer");
  APNERF_REQUIRE(occs_old != occs_new, "occ_update: occs_old must be a snapshot (cells may repeat)");
  FieldIO io;
  memset(&io, 0, sizeof(io));
  io.n = n, io.density_only = 1;
  io.table = (const uint2*)table, io.weights = (const uint4*)weights;
  io.cell_ids = cell_ids, io.jitter = jitter, io.occs_old = occs_old, io.occs_new = occs_new;
  io.occ_scale = occ_scale, io.ema_decay = ema_decay;
  io.cell_res[0] = rx, io.cell_res[1] = ry, io.cell_res[2] = rz;
  for (int a = 0; a < 3; ++a) {
    io.cell_lo[a] = level_aabb_host[a];
    io.cell_ext[a] = level_aabb_host[3 + a] - level_aabb_host[a];  // fp32, as aabbs[lvl, 3:] - aabbs[lvl, :3]
  }
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "occ_update: bad level table");
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = aabb_host[i];
  const long long tiles = (n + TILE_M - 1) / TILE_M;
  const int sms = apnerf_num_sms();
  return launch_field<0>(io, m, fc, (int)(tiles < sms ? tiles : sms), FIELD_SMEM_MIN, (cudaStream_t)stream,
                         "field_forward_kernel(occ_update)");
}

// Field query of the device-driven renderer on the marcher's sample rows (s_ray: ray id, -1 = padding; s_x: the
// sample's aabb-normalised point, written by apnerf_render_march*): one 80-byte row of raw fp16 network outputs per
// sample for apnerf_render_composite.  *n_rows_dev rows (clamped to max_tiles * 128).
APNERF_API int apnerf_field_forward_rows(const int* n_rows_dev, long long max_tiles, const int* s_ray, const void* s_x,
                                         const float* rays_d, const float* aabb_host, int n_levels,
                                         const uint32_t* meta_host, const void* table, const void* weights,
                                         void* packed, void* stream) {
  APNERF_REQUIRE(n_rows_dev && s_ray && s_x && rays_d && packed, "field_forward_rows: null buffer");
  FieldIO io;
  memset(&io, 0, sizeof(io));
  io.n = max_tiles * TILE_M;  // capacity of the row buffers
  io.n_dev = n_rows_dev, io.ray_idx = s_ray, io.x01 = (const float4*)s_x, io.rays_d = rays_d;
  io.table = (const uint2*)table, io.weights = (const uint4*)weights, io.packed = (uint4*)packed;
  io.rays_per_call = 1;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "field_forward_rows: bad level table");
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = aabb_host[i];
  const int sms = apnerf_num_sms();
  const int grid = (int)(max_tiles < 1 ? 1 : (max_tiles < sms ? max_tiles : sms));
  // Coarse-level staging (north star (2)): measured on B200, staging level 0 in shared memory is SLOWER than leaving
  // it to L1 (profiles/r02_field_kernel.md) -- L1 and shared memory are the same SRAM behind the same pipe, a level-0
  // gather of 32 neighbouring samples is already a broadcast of one or two lines, and the 32 KB come out of the L1
  // that caches the fine levels.  Kept as an experiment switch.
  static const bool stage = getenv("APNERF_FIELD_STAGE_L0") && atoi(getenv("APNERF_FIELD_STAGE_L0")) != 0;
  io.stage_level0 = stage ? 1 : 0;
  return launch_field<3>(io, m, fc, grid, FIELD_SMEM_MIN + (stage ? FIELD_L0_BYTES : 0), (cudaStream_t)stream,
                         "field_forward_kernel(rows)");
}

// The same with the compositor fused into the epilogue: sample rows (s_ray, s_cnt, s_ts, s_te, s_x; *n_rows_dev
// rows, a multiple of 128, rays never straddle a tile) -> per-ray state update + keep flags.  Replaces
// apnerf_field_forward_rows + apnerf_render_composite.
APNERF_API int apnerf_field_forward_fused(const int* n_rows_dev, long long max_tiles, const int* s_ray,
                                          const uint8_t* s_cnt, const float* s_ts, const float* s_te, const void* s_x,
                                          const float* rays_d, const float* aabb_host,
                                          int n_levels, const uint32_t* meta_host, const void* table,
                                          const void* weights, int n_sem, float* state, int n_rays_total,
                                          int rays_per_call, float alpha_thre, float opc_thre, const int* n_samp,
                                          const int* iter_samples, int max_samples, uint8_t* keep_flag,
                                          int* total_samples, int probabilistic, int* ray_counts, void* stream) {
  APNERF_REQUIRE(n_sem >= 0 && n_sem <= SEM_OUT, "field_forward_fused: at most 32 semantic classes");
  APNERF_REQUIRE(n_rows_dev && s_ray && s_cnt && s_ts && s_te && s_x && rays_d && state, "field_forward_fused: null buffer");
  FieldIO io;
  memset(&io, 0, sizeof(io));
  io.n = max_tiles * TILE_M;  // capacity of the row buffers
  io.n_dev = n_rows_dev, io.ray_idx = s_ray, io.t_starts = s_ts, io.t_ends = s_te, io.rays_d = rays_d;
  io.x01 = (const float4*)s_x;
  io.table = (const uint2*)table, io.weights = (const uint4*)weights, io.n_sem = n_sem;
  io.state = state, io.n_rays_total = n_rays_total, io.rays_per_call = rays_per_call, io.alpha_thre = alpha_thre;
  io.opc_thre = opc_thre, io.n_samp = n_samp, io.iter_samples = iter_samples, io.max_samples = max_samples;
  io.s_cnt = s_cnt, io.keep_flag = keep_flag, io.total_samples = total_samples, io.probabilistic = probabilistic;
  io.ray_counts = ray_counts;
  HashGridMeta m;
  APNERF_REQUIRE(fill_meta(m, n_levels, meta_host) == 0, "field_forward_fused: bad level table");
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = aabb_host[i];
  const int sms = apnerf_num_sms();
  const int grid = (int)(max_tiles < 1 ? 1 : (max_tiles < sms ? max_tiles : sms));
  return launch_field<2>(io, m, fc, grid, FIELD_SMEM, (cudaStream_t)stream, "field_forward_kernel(fused)");
}
