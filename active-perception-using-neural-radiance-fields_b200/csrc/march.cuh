// Occupancy-grid ray marcher (kernel 1).  One thread walks one ray; a warp holds 32
// neighbouring rays of the same image so the DDA stays coherent.
//
// Replaces the reference's traverse_grids_kernel (perception/nerfacc/nerfacc/cuda/csrc/grid.cu:68-282,
// helpers include/utils_grid.cuh:58-142).  The march is a serial fp32 recurrence, so bit-exact
// parity fixes every rounding: each operation below is an explicit round-to-nearest intrinsic
// (never contracted by the compiler), and fused multiply-adds appear exactly where the
// reference build has an FFMA.
#pragma once
#include "common.cuh"

namespace apnerf {

struct GridView {
  const uint8_t* binaries;  // [n_grids, rx, ry, rz] bool
  const float* aabbs;       // [n_grids, 6]
  int n_grids;
  int rx, ry, rz;
  float skip_min_steps;  // use the closed-form skip only when more than this many steps are expected
};

__device__ __forceinline__ float calc_dt(float t, float cone, float dt_min) {
  // grid.cu:23-28 : clamp(t * cone, dt_min, 1e10) = fmaxf(dt_min, fminf(t * cone, 1e10))
  return fmaxf(dt_min, fminf(__fmul_rn(t, cone), 1e10f));
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

// Slab test, include/utils_grid.cuh:10-55.  inv = 1/d (rcp.rn).
__device__ __forceinline__ bool ray_aabb(const float o[3], const float inv[3], float rtmin, float rtmax,
                                         const float* __restrict__ ab, float& tmin, float& tmax) {
  float a, b;
  if (inv[0] >= 0) { tmin = __fmul_rn(__fsub_rn(ab[0], o[0]), inv[0]); tmax = __fmul_rn(__fsub_rn(ab[3], o[0]), inv[0]); }
  else             { tmin = __fmul_rn(__fsub_rn(ab[3], o[0]), inv[0]); tmax = __fmul_rn(__fsub_rn(ab[0], o[0]), inv[0]); }
  if (inv[1] >= 0) { a = __fmul_rn(__fsub_rn(ab[1], o[1]), inv[1]); b = __fmul_rn(__fsub_rn(ab[4], o[1]), inv[1]); }
  else             { a = __fmul_rn(__fsub_rn(ab[4], o[1]), inv[1]); b = __fmul_rn(__fsub_rn(ab[1], o[1]), inv[1]); }
  if (tmin > b || a > tmax) return false;
  if (a > tmin) tmin = a;
  if (b < tmax) tmax = b;
  if (inv[2] >= 0) { a = __fmul_rn(__fsub_rn(ab[2], o[2]), inv[2]); b = __fmul_rn(__fsub_rn(ab[5], o[2]), inv[2]); }
  else             { a = __fmul_rn(__fsub_rn(ab[5], o[2]), inv[2]); b = __fmul_rn(__fsub_rn(ab[2], o[2]), inv[2]); }
  if (tmin > b || a > tmax) return false;
  if (a > tmin) tmin = a;
  if (b < tmax) tmax = b;
  if (tmax <= 0) return false;
  tmin = fmaxf(tmin, rtmin);
  tmax = fminf(tmax, rtmax);
  return true;
}


// Exact closed form of the reference's empty-space skip loop (grid.cu:157-161, 199-203)
//     while (fma(dt, 0.5, t) < target) t = t + dt;          // every add rounded to nearest
// Inside one binade a float's bit pattern is its mantissa counter, and adding the same dt always
// advances it by the same integer `inc` (dt / ulp rounded to nearest; the rounding direction depends
// only on dt's low bits, not on t), so k steps are `bits(t) + k * inc`.  k is estimated in floating
// point and then corrected with the loop's own predicate, so the result is bit-identical to running
// the loop (checked against the reference kernels in tests/test_gpu_reference.py).  Ties (dt's low
// bits exactly half an ulp), binade crossings and tiny t fall back to single exact steps.
__device__ __forceinline__ float skip_to(float t, float dt, float target, float min_steps) {
  if (!(target - t > min_steps * dt)) {  // short skip: the plain loop is cheaper than the closed form
    while (__fmaf_rn(dt, 0.5f, t) < target) t = __fadd_rn(t, dt);
    return t;
  }
  const float h = dt * 0.5f;  // exact (dt >= step_size is never subnormal here)
  while (__fmaf_rn(dt, 0.5f, t) < target) {
    const int tb = __float_as_int(t), db = __float_as_int(dt);
    const int e = (tb >> 23) & 0xff, ed = (db >> 23) & 0xff;
    const int s = e - ed;
    bool jumped = false;
    if (t > 0.0f && s >= 1 && s <= 23 && e > 24 && e < 0xfe) {
      const unsigned md = ((unsigned)db & 0x7fffffu) | 0x800000u;
      const unsigned q = md >> s, rem = md & ((1u << s) - 1u), half = 1u << (s - 1);
      const unsigned inc = q + (rem > half ? 1u : 0u);
      if (rem != half && inc > 0u) {
        const unsigned room = 0x7fffffu - ((unsigned)tb & 0x7fffffu);
        const int kmax = (int)(room / inc);
        if (kmax >= 2) {
          const float ulp = __int_as_float((e - 23) << 23);
          const float kf = (target - h - t) / ((float)inc * ulp);
          int k = kf >= (float)kmax ? kmax : (kf <= 0.0f ? 0 : (int)kf);
          while (k < kmax && __fmaf_rn(dt, 0.5f, __int_as_float(tb + k * (int)inc)) < target) ++k;
          while (k > 0 && !(__fmaf_rn(dt, 0.5f, __int_as_float(tb + (k - 1) * (int)inc)) < target)) --k;
          if (k > 0) {
            t = __int_as_float(tb + k * (int)inc);
            jumped = true;
          }
        }
      }
    }
    if (!jumped) t = __fadd_rn(t, dt);
  }
  return t;
}

// March one ray.  `sink(t_last, t_next, continuous)` is called once per emitted sample, in
// order; it returns nothing (counting is done here).  `limit` <= 0 means unlimited.
// Returns the number of samples; `n_intervals` gets the number of interval edges
// (#samples + #runs) and `t_term` the terminate plane (grid.cu:274-280).
template <class Sink>
__device__ __forceinline__ int march_ray(const GridView& g, const float o[3], const float d[3],
                                         float near_plane, float far_plane,
                                         const uint8_t* __restrict__ hits,        // [n_grids] of this ray
                                         const float* __restrict__ t_sorted,      // [2*n_grids]
                                         const int64_t* __restrict__ t_indices,   // [2*n_grids] or nullptr (identity)
                                         float step_size, float cone_angle, int limit, Sink& sink,
                                         int& n_intervals, float& t_term) {
  const float eps = 1e-6f;
  const float inv[3] = {__frcp_rn(d[0]), __frcp_rn(d[1]), __frcp_rn(d[2])};
  int n_samples = 0;
  n_intervals = 0;
  float t_last = near_plane;
  bool continuous = false;
  const int n_grids = g.n_grids;
  for (int i = 0; i < n_grids * 2 - 1; ++i) {  // grid.cu:129
    const int64_t ti = t_indices ? t_indices[i] : (int64_t)i;
    const bool is_entering = ti < n_grids;
    int level = (int)(ti % n_grids);
    if (!hits[level]) continue;
    if (!is_entering) {
      const int64_t tn = t_indices ? t_indices[i + 1] : (int64_t)(i + 1);
      if (tn < n_grids) continue;
      level = (int)(tn % n_grids);
      if (!hits[level]) continue;
    }
    const float this_tmin = fmaxf(t_sorted[i], near_plane);
    const float this_tmax = fminf(t_sorted[i + 1], far_plane);
    if (this_tmin >= this_tmax) continue;

    if (!continuous) {  // grid.cu:153-163
      if (step_size <= 0.0f) {
        t_last = this_tmin;
      } else {
        const float dt = calc_dt(t_last, cone_angle, step_size);
        t_last = skip_to(t_last, dt, this_tmin, g.skip_min_steps);
      }
    }

    // setup_traversal, include/utils_grid.cuh:58-114
    const float* __restrict__ ab = g.aabbs + level * 6;
    const int resv[3] = {g.rx, g.ry, g.rz};
    float tdist[3], delta[3];
    int cur[3], stp[3], ovf[3];
    const float ts = __fadd_rn(this_tmin, eps), te = __fsub_rn(this_tmax, eps);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float fres = (float)resv[a];
      const float ext = __fsub_rn(ab[3 + a], ab[a]);
      const float voxel = __fdiv_rn(ext, fres);
      const float ray_start = __fmaf_rn(d[a], ts, o[a]);
      const float ray_end = __fmaf_rn(d[a], te, o[a]);
      cur[a] = clampi(__float2int_rz(__fmul_rn(__fdiv_rn(__fsub_rn(ray_start, ab[a]), ext), fres)), 0, resv[a] - 1);
      const int fin = clampi(__float2int_rz(__fmul_rn(__fdiv_rn(__fsub_rn(ray_end, ab[a]), ext), fres)), 0, resv[a] - 1);
      const int start_index = cur[a] + (d[a] > 0.0f ? 1 : 0);
      const float inner = __fmaf_rn((float)start_index, voxel, -ray_start);
      const float tm = __fmaf_rn(__fadd_rn(ab[a], inner), inv[a], this_tmin);
      const float sf = (d[a] == 0.0f) ? 0.0f : (d[a] > 0.0f ? 1.0f : -1.0f);
      tdist[a] = (d[a] == 0.0f) ? this_tmax : tm;
      stp[a] = (int)sf;
      delta[a] = (d[a] == 0.0f) ? this_tmax : __fmul_rn(__fmul_rn(voxel, inv[a]), sf);
      ovf[a] = fin + stp[a];
    }
    // The cell id is kept incrementally (cur[] itself is only needed for the start cell) and the
    // "stepped onto the overflow cell" test cur[a] == ovf[a] (utils_grid.cuh:131-141) as a per-axis
    // difference that a step reduces by stp[a]: the same predicate, fewer instructions per cell.
    const uint8_t* __restrict__ lvl_bin = g.binaries + (int64_t)level * g.rx * g.ry * g.rz;
    int cid = cur[0] * g.ry * g.rz + cur[1] * g.rz + cur[2];
    const int idstep[3] = {stp[0] * g.ry * g.rz, stp[1] * g.rz, stp[2]};
    int diff[3] = {ovf[0] - cur[0], ovf[1] - cur[1], ovf[2] - cur[2]};

    while (limit <= 0 || n_samples < limit) {  // grid.cu:184
      const float t_traverse = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
      if (!lvl_bin[cid]) {
        if (step_size <= 0.0f) {
          t_last = t_traverse;
        } else {
          const float dt = calc_dt(t_last, cone_angle, step_size);
          t_last = skip_to(t_last, dt, t_traverse, g.skip_min_steps);
        }
        continuous = false;
      } else {
        while (limit <= 0 || n_samples < limit) {  // grid.cu:208
          float t_next;
          if (step_size <= 0.0f) {
            t_next = t_traverse;
          } else {
            const float dt = calc_dt(t_last, cone_angle, step_size);
            if (__fmaf_rn(dt, 0.5f, t_last) >= t_traverse) break;
            t_next = __fadd_rn(t_last, dt);
          }
          sink(t_last, t_next, continuous, n_samples, n_intervals);
          n_intervals += continuous ? 1 : 2;
          n_samples++;
          continuous = true;
          t_last = t_next;
          if (t_next >= t_traverse) break;
        }
      }
      // single_traversal, include/utils_grid.cuh:116-142: step along the axis with the nearest boundary.
      // Three tiny predicated bodies instead of a three-way branch (the lanes of a warp pick different axes).
      const bool m0 = (tdist[0] < tdist[1]) && (tdist[0] < tdist[2]);
      const bool m1 = !m0 && (tdist[1] < tdist[2]);
      const bool m2 = !m0 && !m1;
      if (m0) { tdist[0] = __fadd_rn(tdist[0], delta[0]); diff[0] -= stp[0]; cid += idstep[0]; }
      if (m1) { tdist[1] = __fadd_rn(tdist[1], delta[1]); diff[1] -= stp[1]; cid += idstep[1]; }
      if (m2) { tdist[2] = __fadd_rn(tdist[2], delta[2]); diff[2] -= stp[2]; cid += idstep[2]; }
      if ((m0 ? diff[0] : (m1 ? diff[1] : diff[2])) == 0) break;
    }
  }
  t_term = t_last;
  return n_samples;
}

}  // namespace apnerf
