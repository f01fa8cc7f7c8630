/*
 * apnerf_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar loops; threaded by range from Python) of the reference's
 * render hot path.  It is the checker the CUDA kernels are compared against and the
 * "port" CPU baseline bench.py times; the product never links, imports or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Floating-point contraction is spelled out explicitly: this file must
 * be compiled with -ffp-contract=off, and fmaf() appears exactly where nvcc/ptxas fuse a
 * multiply-add in the reference build (re-checked on the GPU against
 * oracle/_ref, the reference's own kernels, by tests/test_traverse_gpu.py).
 *
 * Parity status:
 *   - ray_aabb / traverse_grids: pinned against the reference CUDA kernels (oracle/_ref)
 *     on the GPU box, bit-exact.
 *   - hash-grid / SH / MLP:  "parity unpinned" -- tiny-cuda-nn is an external, unpinned
 *     dependency absent from /root/reference (perception/models/requirements.txt:1); this
 *     restates its published algorithm.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* Every entry point works on an index range [i0, i1) so the Python wrapper can spread
 * ranges over host threads (ctypes drops the GIL); this image's gcc has no libgomp. */

/* cvt.rzi.s32.f32 semantics (saturating, NaN -> 0); utils_math.cuh:177-180 int(float). */
static inline int32_t f2i_rz(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}
static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

/* ---------------------------------------------------------------------------------
 * ray_aabb_intersect : nerfacc/cuda/csrc/include/utils_grid.cuh:10-55 (slab test) and
 * the kernel wrapper csrc/grid.cu:284-313.  inv_dir = 1.0f/dir (rcp.rn,
 * data_spec_packed.cuh:49).
 * --------------------------------------------------------------------------------- */
static int slab(const float *o, const float *d, float rtmin, float rtmax, const float *ab,
                float *tmin_out, float *tmax_out) {
  float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
  float tmin, tmax, a, b;
  if (inv[0] >= 0) { tmin = (ab[0] - o[0]) * inv[0]; tmax = (ab[3] - o[0]) * inv[0]; }
  else             { tmin = (ab[3] - o[0]) * inv[0]; tmax = (ab[0] - o[0]) * inv[0]; }
  if (inv[1] >= 0) { a = (ab[1] - o[1]) * inv[1]; b = (ab[4] - o[1]) * inv[1]; }
  else             { a = (ab[4] - o[1]) * inv[1]; b = (ab[1] - o[1]) * inv[1]; }
  if (tmin > b || a > tmax) return 0;
  if (a > tmin) tmin = a;
  if (b < tmax) tmax = b;
  if (inv[2] >= 0) { a = (ab[2] - o[2]) * inv[2]; b = (ab[5] - o[2]) * inv[2]; }
  else             { a = (ab[5] - o[2]) * inv[2]; b = (ab[2] - o[2]) * inv[2]; }
  if (tmin > b || a > tmax) return 0;
  if (a > tmin) tmin = a;
  if (b < tmax) tmax = b;
  if (tmax <= 0) return 0;
  *tmin_out = fmaxf(tmin, rtmin);
  *tmax_out = fminf(tmax, rtmax);
  return 1;
}

API void apo_ray_aabb_intersect(int32_t r0, int32_t r1, const float *rays_o, const float *rays_d,
                                float near, float far, int32_t n_aabbs, const float *aabbs,
                                float miss, float *t_mins, float *t_maxs, uint8_t *hits) {
  for (int64_t tid = (int64_t)r0 * n_aabbs; tid < (int64_t)r1 * n_aabbs; ++tid) {
    int32_t r = (int32_t)(tid / n_aabbs), b = (int32_t)(tid % n_aabbs);
    float t0, t1;
    int hit = slab(rays_o + 3 * r, rays_d + 3 * r, near, far, aabbs + 6 * b, &t0, &t1);
    t_mins[tid] = hit ? t0 : miss;
    t_maxs[tid] = hit ? t1 : miss;
    hits[tid] = (uint8_t)hit;
  }
}

/* ---------------------------------------------------------------------------------
 * traverse_grids_kernel : nerfacc/cuda/csrc/grid.cu:68-282
 *   _calc_dt            : grid.cu:23-28   (clamp = fmaxf(a, fminf(f, b)), utils_math.cuh:1167)
 *   setup_traversal     : include/utils_grid.cuh:58-114
 *   single_traversal    : include/utils_grid.cuh:116-142
 * One call = one kernel launch (first_pass = count only; else fill).
 * Segment pointers may be NULL exactly like PackedRaySegmentsSpec fields.
 * --------------------------------------------------------------------------------- */
static inline float calc_dt(float t, float cone, float dt_min, float dt_max) {
  return fmaxf(dt_min, fminf(t * cone, dt_max));
}

API void apo_traverse_grids(
    int32_t r0, int32_t r1, const float *rays_o, const float *rays_d, const uint8_t *rays_mask,
    int32_t n_grids, const int32_t *res, const uint8_t *binaries, const float *aabbs,
    const uint8_t *hits, const float *t_sorted, const int64_t *t_indices,
    const float *near_planes, const float *far_planes, float step_size, float cone_angle,
    int32_t limit, int32_t first_pass,
    /* intervals */ float *iv_vals, int64_t *iv_ray, uint8_t *iv_left, uint8_t *iv_right,
    const int64_t *iv_starts, int64_t *iv_cnts,
    /* samples */ float *sm_vals, int64_t *sm_ray, uint8_t *sm_valid, const int64_t *sm_starts,
    int64_t *sm_cnts,
    float *terminate_planes) {
  const float eps = 1e-6f;
  const int32_t rx = res[0], ry = res[1], rz = res[2];
  for (int32_t tid = r0; tid < r1; ++tid) {
    if (rays_mask && !rays_mask[tid]) continue;                       /* grid.cu:100 */
    if (iv_cnts && !first_pass && iv_cnts[tid] == 0) continue;        /* :103-106 */
    if (sm_cnts && !first_pass && sm_cnts[tid] == 0) continue;
    int64_t chunk_start = 0, chunk_start_bin = 0;
    if (!first_pass) {
      if (iv_cnts) chunk_start = iv_starts[tid];
      if (sm_cnts) chunk_start_bin = sm_starts[tid];
    }
    const float near_plane = near_planes[tid], far_plane = far_planes[tid];
    const float *o = rays_o + 3 * tid, *d = rays_d + 3 * tid;
    const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    const int32_t base_hits = tid * n_grids, base_t = tid * n_grids * 2;

    int64_t n_intervals = 0, n_samples = 0;
    float t_last = near_plane;
    int continuous = 0;
    for (int32_t i = base_t; i < base_t + n_grids * 2 - 1; ++i) {     /* :129 */
      int is_entering = t_indices[i] < n_grids;
      int64_t level = t_indices[i] % n_grids;
      if (!hits[base_hits + level]) continue;
      if (!is_entering) {
        int next_is_entering = t_indices[i + 1] < n_grids;
        if (next_is_entering) continue;
        level = t_indices[i + 1] % n_grids;
        if (!hits[base_hits + level]) continue;
      }
      float this_tmin = fmaxf(t_sorted[i], near_plane);
      float this_tmax = fminf(t_sorted[i + 1], far_plane);
      if (this_tmin >= this_tmax) continue;

      if (!continuous) {                                              /* :153-163 */
        if (step_size <= 0.0f) {
          t_last = this_tmin;
        } else {
          float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
          while (1) {
            if (fmaf(dt, 0.5f, t_last) >= this_tmin) break;
            t_last += dt;
          }
        }
      }

      /* setup_traversal (utils_grid.cuh:58-114) */
      const float *ab = aabbs + level * 6;
      float voxel[3], ray_start[3], ray_end[3], tdist[3], delta[3];
      int32_t cur[3], fin[3], stp[3], ovf[3];
      const int32_t resv[3] = {rx, ry, rz};
      const float ts = this_tmin + eps, te = this_tmax - eps;
      for (int a = 0; a < 3; ++a) {
        float fres = (float)resv[a];
        voxel[a] = (ab[3 + a] - ab[a]) / fres;
        ray_start[a] = fmaf(d[a], ts, o[a]);                          /* FFMA in the ref SASS */
        ray_end[a] = fmaf(d[a], te, o[a]);
        float ext = ab[3 + a] - ab[a];
        cur[a] = clampi(f2i_rz(((ray_start[a] - ab[a]) / ext) * fres), 0, resv[a] - 1);
        fin[a] = clampi(f2i_rz(((ray_end[a] - ab[a]) / ext) * fres), 0, resv[a] - 1);
        int32_t start_index = cur[a] + (d[a] > 0 ? 1 : 0);
        /* ((aabb.min + ((float(idx) * voxel) - ray_start)) * inv_dir) + tmin :
           inner mul-sub fused (ptxas), FADD, outer mul-add fused (nvcc). */
        float inner = fmaf((float)start_index, voxel[a], -ray_start[a]);
        float tm = fmaf(ab[a] + inner, inv[a], this_tmin);
        tdist[a] = (d[a] == 0.0f) ? this_tmax : tm;
        float sf = (d[a] == 0.0f) ? 0.0f : (d[a] > 0.0f ? 1.0f : -1.0f);
        stp[a] = (int32_t)sf;
        float dtemp = (voxel[a] * inv[a]) * sf;
        delta[a] = (d[a] == 0.0f) ? this_tmax : dtemp;
        ovf[a] = fin[a] + stp[a];
      }

      while (limit <= 0 || n_samples < limit) {                       /* :184 */
        float t_traverse = fminf(tdist[0], fminf(tdist[1], tdist[2]));
        t_traverse = fminf(t_traverse, this_tmax);
        int64_t cell_id = (int64_t)(cur[0] * ry * rz + cur[1] * rz + cur[2]) +
                          level * (int64_t)rx * ry * rz;
        if (!binaries[cell_id]) {
          if (step_size <= 0.0f) {
            t_last = t_traverse;
          } else {
            float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
            while (1) {
              if (fmaf(dt, 0.5f, t_last) >= t_traverse) break;
              t_last += dt;
            }
          }
          continuous = 0;
        } else {
          while (limit <= 0 || n_samples < limit) {                   /* :208 */
            float t_next;
            if (step_size <= 0.0f) {
              t_next = t_traverse;
            } else {
              float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
              if (fmaf(dt, 0.5f, t_last) >= t_traverse) break;
              t_next = t_last + dt;
            }
            if (iv_cnts) {                                            /* :219-246 */
              if (!continuous) {
                if (!first_pass) {
                  int64_t idx = chunk_start + n_intervals;
                  iv_vals[idx] = t_last; iv_ray[idx] = tid; iv_left[idx] = 1;
                }
                n_intervals++;
                if (!first_pass) {
                  int64_t idx = chunk_start + n_intervals;
                  iv_vals[idx] = t_next; iv_ray[idx] = tid; iv_right[idx] = 1;
                }
                n_intervals++;
              } else {
                if (!first_pass) {
                  int64_t idx = chunk_start + n_intervals;
                  iv_vals[idx] = t_next; iv_ray[idx] = tid;
                  iv_left[idx - 1] = 1; iv_right[idx] = 1;
                }
                n_intervals++;
              }
            }
            if (sm_cnts) {                                            /* :249-256 */
              if (!first_pass) {
                int64_t idx = chunk_start_bin + n_samples;
                sm_vals[idx] = (t_next + t_last) * 0.5f;
                sm_ray[idx] = tid; sm_valid[idx] = 1;
              }
            }
            n_samples++;
            continuous = 1;
            t_last = t_next;
            if (t_next >= t_traverse) break;
          }
        }
        /* single_traversal (utils_grid.cuh:116-142) */
        int a;
        if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) a = 0;
        else if (tdist[1] < tdist[2]) a = 1;
        else a = 2;
        cur[a] += stp[a];
        tdist[a] += delta[a];
        if (cur[a] == ovf[a]) break;
      }
    }
    if (terminate_planes) terminate_planes[tid] = t_last;             /* :274-280 */
    if (iv_cnts) iv_cnts[tid] = n_intervals;
    if (sm_cnts) sm_cnts[tid] = n_samples;
  }
}

/* ---------------------------------------------------------------------------------
 * Multiresolution hash-grid encoding (tiny-cuda-nn GridEncoding, HashGrid / Linear
 * interpolation / no smoothstep), as instantiated at
 * perception/models/radiance_fields/ngp.py:123-133.  "parity unpinned" (see header).
 *   meta[l] = {scale (float bits), resolution, level_size (entries), offset (entries)}
 *   table   : fp16 bit patterns, [total_entries][F=4]
 *   out_enc : fp16 bit patterns, [N][16*4] level-major;  out_idx: [N][L][8] uint32 (optional)
 * --------------------------------------------------------------------------------- */
static inline float half_to_float(uint16_t h) {
  uint32_t s = (uint32_t)(h & 0x8000) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ff, bits;
  if (e == 0) {
    if (m == 0) bits = s;
    else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } m &= 0x3ff;
           bits = s | ((uint32_t)(113 - sh) << 23) | (m << 13); }
  } else if (e == 31) bits = s | 0x7f800000u | (m << 13);
  else bits = s | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &bits, 4); return f;
}
static inline uint16_t float_to_half(float f) { /* round-to-nearest-even, cvt.rn.f16.f32 */
  uint32_t x; memcpy(&x, &f, 4);
  uint32_t s = (x >> 16) & 0x8000; int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t m = x & 0x7fffff;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00 | (m ? 0x200 : 0));
  if (e >= 31) return (uint16_t)(s | 0x7c00);
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000; int sh = 14 - e; uint32_t hm = m >> sh, rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (hm & 1))) hm++;
    return (uint16_t)(s | hm);
  }
  uint32_t hm = m >> 13, rem = m & 0x1fff; uint16_t h = (uint16_t)(s | ((uint32_t)e << 10) | hm);
  if (rem > 0x1000 || (rem == 0x1000 && (hm & 1))) h++;
  return h;
}

API void apo_half_to_float(int64_t n, const uint16_t *h, float *f) {
  for (int64_t i = 0; i < n; ++i) f[i] = half_to_float(h[i]);
}
API void apo_float_to_half(int64_t n, const float *f, uint16_t *h) {
  for (int64_t i = 0; i < n; ++i) h[i] = float_to_half(f[i]);
}

API void apo_hashgrid_encode(int64_t s0, int64_t s1, const float *x /* [n][3] in (0,1) */, int32_t n_levels,
                             const uint32_t *meta /* [L][4] */, const uint16_t *table,
                             uint16_t *out_enc, uint32_t *out_idx) {
  for (int64_t s = s0; s < s1; ++s) {
    for (int32_t l = 0; l < n_levels; ++l) {
      float scale; memcpy(&scale, &meta[4 * l + 0], 4);
      const uint32_t resolution = meta[4 * l + 1], size = meta[4 * l + 2], offset = meta[4 * l + 3];
      uint32_t cell[3]; float w[3];
      for (int a = 0; a < 3; ++a) {
        float pos = fmaf(scale, x[3 * s + a], 0.5f);
        float fl = floorf(pos);
        cell[a] = (uint32_t)f2i_rz(fl);
        w[a] = pos - fl;
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (uint32_t c = 0; c < 8; ++c) {
        uint32_t g[3]; float wt = 1.0f;
        for (int a = 0; a < 3; ++a) {
          if (c & (1u << a)) { g[a] = cell[a] + 1; wt = wt * w[a]; }
          else               { g[a] = cell[a];     wt = wt * (1.0f - w[a]); }
        }
        uint32_t stride = 1, index = 0;
        for (int a = 0; a < 3 && stride <= size; ++a) { index += g[a] * stride; stride *= resolution; }
        if (size < stride) index = g[0] ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
        index = index % size;
        if (out_idx) out_idx[(s * n_levels + l) * 8 + c] = offset + index;
        const uint16_t *e = table + ((int64_t)(offset + index)) * 4;
        for (int f = 0; f < 4; ++f) acc[f] = fmaf(wt, half_to_float(e[f]), acc[f]);
      }
      for (int f = 0; f < 4; ++f) out_enc[s * (n_levels * 4) + l * 4 + f] = float_to_half(acc[f]);
    }
  }
}

/* ---------------------------------------------------------------------------------
 * Packed exclusive sum (nerfacc/scan.py:57-97 -> csrc/scan.cu:68-125): per-ray exclusive
 * prefix sum over a flattened array.  The reference uses a 32-wide Blelloch tree per tile
 * (include/utils_scan.cuh:146-263); this restatement sums sequentially in fp32, which is
 * what the reference's own test compares against (tests/test_scan.py:38-64, atol 3e-4).
 * --------------------------------------------------------------------------------- */
API void apo_exclusive_sum(int32_t r0, int32_t r1, const int64_t *starts, const int64_t *cnts,
                           const float *in, float *out, int32_t backward) {
  for (int32_t r = r0; r < r1; ++r) {
    int64_t s = starts[r], c = cnts[r];
    float acc = 0.f;
    if (!backward) for (int64_t i = 0; i < c; ++i) { out[s + i] = acc; acc += in[s + i]; }
    else for (int64_t i = c - 1; i >= 0; --i) { out[s + i] = acc; acc += in[s + i]; }
  }
}

/* ---------------------------------------------------------------------------------
 * Spherical harmonics degree 4 (tiny-cuda-nn SphericalHarmonics encoding as configured at
 * perception/models/radiance_fields/ngp.py:108-121).  Input is the reference's own
 * (d + 1) / 2 (ngp.py:205); tcnn maps it back with 2u - 1.  fp32 math, fp16 output.
 * "parity unpinned" (tcnn absent); every product/sum below is individually rounded.
 * --------------------------------------------------------------------------------- */
API void apo_sh4(int64_t s0, int64_t s1, const float *dirs /* [n][3], unit vectors */,
                 uint16_t *out /* [n][16] fp16 bits */) {
  for (int64_t s = s0; s < s1; ++s) {
    float u0 = (dirs[3 * s + 0] + 1.0f) / 2.0f, u1 = (dirs[3 * s + 1] + 1.0f) / 2.0f,
          u2 = (dirs[3 * s + 2] + 1.0f) / 2.0f;
    float x = u0 * 2.0f - 1.0f, y = u1 * 2.0f - 1.0f, z = u2 * 2.0f - 1.0f;
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    float o[16];
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = (0.59004358992664352f * y) * (-3.0f * x2 + y2);
    o[10] = (2.8906114426405538f * xy) * z;
    o[11] = (0.45704579946446572f * y) * (1.0f - 5.0f * z2);
    o[12] = (0.3731763325901154f * z) * (5.0f * z2 - 3.0f);
    o[13] = (0.45704579946446572f * x) * (1.0f - 5.0f * z2);
    o[14] = (1.4453057213202769f * z) * (x2 - y2);
    o[15] = (0.59004358992664352f * x) * (-x2 + 3.0f * y2);
    for (int i = 0; i < 16; ++i) out[16 * s + i] = float_to_half(o[i]);
  }
}
