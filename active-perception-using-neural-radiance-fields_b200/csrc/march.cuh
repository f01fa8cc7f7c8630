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
  const uint32_t* bits;  // optional: the same occupancy, one bit per cell (cell i = bit i & 31 of word i >> 5 of its level);
                         // 8x less cache footprint than the bool bytes, which is what the serial per-cell walk waits on
};

template <bool BITS>
__device__ __forceinline__ bool cell_occupied(const uint8_t* __restrict__ lvl_bin, const uint32_t* __restrict__ lvl_bits,
                                              int cid) {
  if (BITS) return (__ldg(lvl_bits + (cid >> 5)) >> (cid & 31)) & 1u;
  return lvl_bin[cid] != 0;
}

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


// The reference's empty-space skip loop itself (grid.cu:157-161, 199-203), every add rounded to nearest.
__device__ __forceinline__ float skip_loop(float t, float dt, float target) {
  while (__fmaf_rn(dt, 0.5f, t) < target) t = __fadd_rn(t, dt);
  return t;
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
// FAST (the renderer's instantiation): step_size > 0 is known and the closed-form skip is compiled out (an empty
// 0.1 m cell is at most ~100 steps; the closed form only pays for pathological skips, see skip_to), which removes
// eight instructions of uniform tests from every cell visit; it also reads the occupancy from the bit-packed copy
// (g.bits must be set).
template <bool FAST = false, class Sink>
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

    const bool free_step = !FAST && step_size <= 0.0f;
    if (!continuous) {  // grid.cu:153-163
      if (free_step) {
        t_last = this_tmin;
      } else {
        const float dt = calc_dt(t_last, cone_angle, step_size);
        t_last = FAST ? skip_loop(t_last, dt, this_tmin) : skip_to(t_last, dt, this_tmin, g.skip_min_steps);
      }
    }

    // setup_traversal, include/utils_grid.cuh:58-114
    const float* __restrict__ ab = g.aabbs + level * 6;
    const int resv[3] = {g.rx, g.ry, g.rz};
    const int cell_stride[3] = {g.ry * g.rz, g.rz, 1};
    // Per-axis DDA state, all kept in registers: distance to the next boundary, its increment, the cell-id increment and
    // a countdown to the "stepped onto the overflow cell" exit.  The reference tests cur[a] == overflow_index[a] after
    // cur[a] += step[a] (utils_grid.cuh:131-141); with diff = overflow - cur that is `diff -= step; diff == 0`, and
    // cnt = diff * step counts the same thing down by one per step (step = 0: the reference exits the first time the
    // axis is chosen, cnt = 1 does the same; diff * step <= 0 never reaches zero in either form).
    float tdist[3], delta[3];
    int idstep[3], cnt[3];
    int cid = 0;
    const float ts = __fadd_rn(this_tmin, eps), te = __fsub_rn(this_tmax, eps);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float fres = (float)resv[a];
      const float ext = __fsub_rn(ab[3 + a], ab[a]);
      const float voxel = __fdiv_rn(ext, fres);
      const float ray_start = __fmaf_rn(d[a], ts, o[a]);
      const float ray_end = __fmaf_rn(d[a], te, o[a]);
      const int cur = clampi(__float2int_rz(__fmul_rn(__fdiv_rn(__fsub_rn(ray_start, ab[a]), ext), fres)), 0, resv[a] - 1);
      const int fin = clampi(__float2int_rz(__fmul_rn(__fdiv_rn(__fsub_rn(ray_end, ab[a]), ext), fres)), 0, resv[a] - 1);
      const int start_index = cur + (d[a] > 0.0f ? 1 : 0);
      const float inner = __fmaf_rn((float)start_index, voxel, -ray_start);
      const float tm = __fmaf_rn(__fadd_rn(ab[a], inner), inv[a], this_tmin);
      const int stp = (d[a] == 0.0f) ? 0 : (d[a] > 0.0f ? 1 : -1);
      const float sf = (float)stp;
      tdist[a] = (d[a] == 0.0f) ? this_tmax : tm;
      delta[a] = (d[a] == 0.0f) ? this_tmax : __fmul_rn(__fmul_rn(voxel, inv[a]), sf);
      idstep[a] = stp * cell_stride[a];
      cnt[a] = stp == 0 ? 1 : (fin + stp - cur) * stp;
      cid += cur * cell_stride[a];
    }
    const int n_cells = g.rx * g.ry * g.rz;
    const uint8_t* __restrict__ lvl_bin = g.binaries + (int64_t)level * n_cells;
    const uint32_t* __restrict__ lvl_bits = FAST ? g.bits + (int64_t)level * ((n_cells + 31) >> 5) : nullptr;

    // The walk over the cells is a serial chain whose longest link is the occupancy load, so the NEXT cell's occupancy
    // is fetched before the current cell is processed: which cell comes next depends only on tdist, not on the samples.
    bool occ = cell_occupied<FAST>(lvl_bin, lvl_bits, cid);
    while (limit <= 0 || n_samples < limit) {  // grid.cu:184
      const float t_traverse = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
      // single_traversal, include/utils_grid.cuh:116-142: step along the axis with the nearest boundary
      const bool m0 = (tdist[0] < tdist[1]) && (tdist[0] < tdist[2]);
      const bool m1 = !m0 && (tdist[1] < tdist[2]);
      const int next_cid = cid + (m0 ? idstep[0] : (m1 ? idstep[1] : idstep[2]));
      const int left = (m0 ? cnt[0] : (m1 ? cnt[1] : cnt[2])) - 1;
      bool occ_next = false;
      if (left != 0 && (unsigned)next_cid < (unsigned)n_cells) occ_next = cell_occupied<FAST>(lvl_bin, lvl_bits, next_cid);
      if (!occ) {
        if (free_step) {
          t_last = t_traverse;
        } else {
          const float dt = calc_dt(t_last, cone_angle, step_size);
          t_last = FAST ? skip_loop(t_last, dt, t_traverse) : skip_to(t_last, dt, t_traverse, g.skip_min_steps);
        }
        continuous = false;
      } else {
        while (limit <= 0 || n_samples < limit) {  // grid.cu:208
          float t_next;
          if (free_step) {
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
      if (m0) { tdist[0] = __fadd_rn(tdist[0], delta[0]); cnt[0] = left; }
      else if (m1) { tdist[1] = __fadd_rn(tdist[1], delta[1]); cnt[1] = left; }
      else { tdist[2] = __fadd_rn(tdist[2], delta[2]); cnt[2] = left; }
      cid = next_cid;
      occ = occ_next;
      if (left == 0) break;
    }
  }
  t_term = t_last;
  return n_samples;
}

}  // namespace apnerf
