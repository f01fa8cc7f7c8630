// The device-driven test-mode renderer and the predictive-information scorer (kernels 1, 4, 5 in
// their fused form).  Replaces, on the reference side,
//   Dataset.generate_image_rays                     perception/data_proc/habitat_to_data.py:274-301
//   render_probablistic_image_with_occgrid_test     perception/models/utils.py:782-1032
//   render_image_with_occgrid_test                  perception/models/utils.py:555-779
//   ActiveNeRFMapper.probablistic_uncertainty       scripts/pipeline.py:727-781 (the arithmetic)
//
// A "call" is one view rendered through one ensemble member (R rays).  The reference marches a
// call in iterations of n = max(min(R // n_alive, 64), min_samples) samples per live ray, with
// three host synchronisations per iteration.  Here a whole batch of calls advances in lock step
// with the SAME per-call schedule, but every decision (n_alive, n, termination, compaction of
// the live-ray list, sample counts) is taken on the device: the host only enqueues a fixed
// kernel sequence per iteration and never reads anything back until the scores are done.
//
// Per-ray running state lives in HBM as structure-of-arrays [channel][ray]:
//   0-2 rgb | 3 opacity | 4 depth | 5-7 rgb_var | 8 depth_var | 9.. semantic logits (C)
#include "march.cuh"
#include "field.cuh"

namespace apnerf {


// counters[0] live rays this iteration, [1] live rays being collected for the next one,
// [2] samples emitted this iteration, [3] total iterations executed with work,
// [8] sample rows evaluated so far in this render (sum of [2] over the finished iterations; 16 ints in all)
__global__ void render_schedule_kernel(int n_calls, int rays_per_call, int max_samples, int min_samples,
                                       int* __restrict__ n_alive_acc, int* __restrict__ n_samp,
                                       int* __restrict__ iter_samples, int* __restrict__ counters,
                                       int* __restrict__ call_rows) {
  for (int c = threadIdx.x; c < n_calls; c += blockDim.x) {
    const int na = n_alive_acc[c];
    n_alive_acc[c] = 0;
    int n = 0;
    if (iter_samples[c] < max_samples && na > 0) {  // utils.py:896-903
      n = max(min(rays_per_call / na, MAX_ITER_SAMPLES), min_samples);
      iter_samples[c] += n;
      if (call_rows) call_rows[c] += n * na;
    }
    n_samp[c] = n;
  }
  if (threadIdx.x == 0) {
    counters[0] = counters[1];
    counters[1] = 0;
    counters[8] += counters[6] > 0 ? counters[6] : counters[2];  // rows of the iteration that just finished ([6]: real rows of the tile layout)
    counters[2] = 0;
    if (counters[0] > 0) counters[3] += 1;
    counters[4] = 0;  // ticket counter of render_compact_kernel
    counters[6] = 0;  // real (non-padding) samples emitted this iteration
    counters[7] = (counters[7] + 1) & 0x3fffffff;  // generation tag of render_compact_kernel (never 0 after this)
    if (counters[7] == 0) counters[7] = 1;
  }
}

// rays from camera poses (OpenGL convention), optionally subsampled by an index list.
__global__ void __launch_bounds__(256) generate_rays_kernel(int n_views, const float* __restrict__ c2w,  // [V,3,4]
                                                            int width, int height, float focal, int n_keep,
                                                            const int* __restrict__ keep_idx,
                                                            float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const long long total = (long long)n_views * n_keep;
  const float cx = (float)(width * 0.5), cy = (float)(height * 0.5);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)blockDim.x * gridDim.x) {
    const int v = (int)(t / n_keep), k = (int)(t % n_keep);
    const int pix = keep_idx ? keep_idx[k] : k;
    const float x = (float)(pix % width), y = (float)(pix / width);
    const float* m = c2w + 12 * v;
    // habitat_to_data.py:285-295: [(x - cx + 0.5) / fx, (y - cy + 0.5) / fy * -1, -1]
    const float c0 = __fdiv_rn(__fadd_rn(__fsub_rn(x, cx), 0.5f), focal);
    const float c1 = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(y, cy), 0.5f), focal), -1.0f);
    const float c2 = -1.0f;
    float d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      d[i] = __fadd_rn(__fadd_rn(__fmul_rn(c0, m[4 * i + 0]), __fmul_rn(c1, m[4 * i + 1])), __fmul_rn(c2, m[4 * i + 2]));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rays_o[3 * t + i] = m[4 * i + 3];
      rays_d[3 * t + i] = __fdiv_rn(d[i], nrm);
    }
  }
}

// occupancy bools -> one bit per cell (word w, bit b = cell 32 w + b): the marcher's copy of the grid
__global__ void __launch_bounds__(256) pack_occupancy_kernel(int n_cells, const uint8_t* __restrict__ binaries,
                                                             uint32_t* __restrict__ bits) {
  const int n_words = (n_cells + 31) >> 5;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += blockDim.x * gridDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int c = 32 * w + b;
      if (c < n_cells && binaries[c]) v |= 1u << b;
    }
    bits[w] = v;
  }
}

__global__ void __launch_bounds__(256) render_init_kernel(int n_rays, int rays_per_call, const float* __restrict__ rays_o,
                                                          const float* __restrict__ rays_d, GridView g,
                                                          float near_plane, int n_state, float* __restrict__ state,
                                                          float* __restrict__ t_min, float* __restrict__ t_max,
                                                          uint8_t* __restrict__ hit, float* __restrict__ near,
                                                          int* __restrict__ alive, int* __restrict__ n_alive_acc,
                                                          int* __restrict__ iter_samples, int* __restrict__ total_samples,
                                                          int n_calls, int* __restrict__ counters) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += blockDim.x * gridDim.x) {
    const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
    const float inv[3] = {__frcp_rn(rays_d[3 * r]), __frcp_rn(rays_d[3 * r + 1]), __frcp_rn(rays_d[3 * r + 2])};
    float t0, t1;
    // ray_aabb_intersect with the default +-inf planes, miss value +inf (utils.py:882, grid.py:14-21)
    const bool h = ray_aabb(o, inv, -INFINITY, INFINITY, g.aabbs, t0, t1);
    t_min[r] = h ? t0 : INFINITY;
    t_max[r] = h ? t1 : INFINITY;
    hit[r] = h ? 1 : 0;
    near[r] = near_plane;
    alive[r] = r;
    for (int c = 0; c < n_state; ++c) state[(size_t)c * n_rays + r] = 0.f;
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < n_calls; c += blockDim.x) {
      n_alive_acc[c] = rays_per_call;
      iter_samples[c] = 0;
      total_samples[c] = 0;
    }
    if (threadIdx.x == 0) counters[0] = 0, counters[1] = n_rays, counters[2] = 0, counters[3] = 0, counters[4] = 0, counters[6] = 0, counters[8] = 0;  // [5] (overflow) is sticky: the host clears it
  }
}

// The sample's aabb-normalised midpoint, op for op as perception/models/utils.py:833-836 (positions = o + d *
// (t_start + t_end) / 2) followed by ngp.py:175-176 (x = (x - aabb_min) / (aabb_max - aabb_min)): written once per
// sample here so that the field kernel's four encoder parts and its epilogue read it instead of re-deriving it.
struct SamplePointer {
  float o[3], d[3], lo[3], ext[3];
  __device__ __forceinline__ SamplePointer(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int ray,
                                           const FieldConst& fc) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      o[a] = rays_o[3 * (size_t)ray + a], d[a] = rays_d[3 * (size_t)ray + a];
      lo[a] = fc.aabb[a], ext[a] = __fsub_rn(fc.aabb[3 + a], fc.aabb[a]);
    }
  }
  __device__ __forceinline__ float4 at(float t0, float t1) const {
    const float tsum = __fadd_rn(t0, t1);
    float x[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = __fadd_rn(o[a], __fmul_rn(__fmul_rn(d[a], tsum), 0.5f));
      x[a] = __fdiv_rn(__fsub_rn(p, lo[a]), ext[a]);
    }
    return make_float4(x[0], x[1], x[2], 0.f);
  }
};

struct LocalSink {
  float* ts;
  float* te;
  __device__ __forceinline__ void operator()(float t_last, float t_next, bool, int i_sample, int) {
    ts[i_sample] = t_last;
    te[i_sample] = t_next;
  }
};

// One thread per live ray: march up to n samples from the ray's previous terminate plane and
// append them to the compact per-iteration sample list (warp-aggregated reservation keeps the
// samples of neighbouring rays adjacent, which is what gives the hash-grid gather its locality).
template <bool FAST>
__global__ void __launch_bounds__(256, 4) render_march_kernel(const int* counters_in, int rays_per_call,
                                                           const int* __restrict__ alive, const int* __restrict__ n_samp,
                                                           const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           GridView g, const float* __restrict__ t_min,
                                                           const float* __restrict__ t_max, const uint8_t* __restrict__ hit,
                                                           float* __restrict__ near, float far_plane, float step_size,
                                                           float cone_angle, int* __restrict__ entry_base,
                                                           int* __restrict__ entry_cnt, int* __restrict__ s_ray,
                                                           float* __restrict__ s_ts, float* __restrict__ s_te,
                                                           FieldConst fc, float4* __restrict__ s_x, int* counters) {
  const int n_live = counters_in[0];
  const int lane = threadIdx.x & 31;
  const int n_round = (n_live + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += blockDim.x * gridDim.x) {
    float ts[MAX_ITER_SAMPLES], te[MAX_ITER_SAMPLES];
    int k = 0, ray = -1;
    if (i < n_live) {
      ray = alive[i];
      const int n = n_samp[ray / rays_per_call];
      if (n > 0) {
        const float o[3] = {rays_o[3 * ray], rays_o[3 * ray + 1], rays_o[3 * ray + 2]};
        const float d[3] = {rays_d[3 * ray], rays_d[3 * ray + 1], rays_d[3 * ray + 2]};
        const float tsorted[2] = {t_min[ray], t_max[ray]};
        const uint8_t h = hit[ray];
        LocalSink sink{ts, te};
        int n_iv;
        float t_term;
        k = march_ray<FAST>(g, o, d, near[ray], far_plane, &h, tsorted, nullptr, step_size, cone_angle, n, sink, n_iv, t_term);
        near[ray] = t_term;  // utils.py:1002
      }
    }
    // warp-aggregated reservation of k slots
    int inc = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
    int warp_base = 0;
    if (lane == 31 && warp_total > 0) warp_base = atomicAdd(counters + 2, warp_total);
    warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
    const int base = warp_base + inc - k;
    if (i < n_live) {
      entry_base[i] = base;
      entry_cnt[i] = k;
      if (k > 0) {
        const SamplePointer sp(rays_o, rays_d, ray, fc);
        for (int j = 0; j < k; ++j) {
          s_ray[base + j] = ray;
          s_ts[base + j] = ts[j];
          s_te[base + j] = te[j];
          s_x[base + j] = sp.at(ts[j], te[j]);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// "v2" marching step of the fused pipeline: the compositor runs inside the field kernel's epilogue,
// which needs every ray's samples inside ONE 128-row tile.  Each warp therefore reserves a
// 128-aligned run of rows and places its 32 rays' samples one after the other, starting a new
// tile whenever a ray would straddle a tile boundary; unused rows are marked ray = -1.
//   s_ray [row]  ray id or -1        s_cnt [row]  k at a ray's first row, 0x80 | j at its j-th row, 0 = padding
// counters[2] = rows reserved this iteration (a multiple of 128).  keep_flag[ray] is cleared for
// every live ray here and set by the fused compositor for the rays that stay live.
template <bool FAST>
__global__ void __launch_bounds__(256) render_march_tiles_kernel(
    const int* counters_in, int rays_per_call, const int* __restrict__ alive, const int* __restrict__ n_samp,
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, GridView g, const float* __restrict__ t_min,
    const float* __restrict__ t_max, const uint8_t* __restrict__ hit, float* __restrict__ near, float far_plane,
    float step_size, float cone_angle, int* __restrict__ s_ray, uint8_t* __restrict__ s_cnt,
    float* __restrict__ s_ts, float* __restrict__ s_te, FieldConst fc, float4* __restrict__ s_x,
    uint8_t* __restrict__ keep_flag, int s_cap, int* counters) {
  const int n_live = counters_in[0];
  const int lane = threadIdx.x & 31;
  const int n_round = (n_live + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += blockDim.x * gridDim.x) {
    float ts[MAX_ITER_SAMPLES], te[MAX_ITER_SAMPLES];
    int k = 0, ray = -1;
    if (i < n_live) {
      ray = alive[i];
      keep_flag[ray] = 0;
      const int n = n_samp[ray / rays_per_call];
      if (n > 0) {
        const float o[3] = {rays_o[3 * ray], rays_o[3 * ray + 1], rays_o[3 * ray + 2]};
        const float d[3] = {rays_d[3 * ray], rays_d[3 * ray + 1], rays_d[3 * ray + 2]};
        const float tsorted[2] = {t_min[ray], t_max[ray]};
        const uint8_t h = hit[ray];
        LocalSink sink{ts, te};
        int n_iv;
        float t_term;
        k = march_ray<FAST>(g, o, d, near[ray], far_plane, &h, tsorted, nullptr, step_size, cone_angle, n, sink, n_iv, t_term);
        near[ray] = t_term;  // utils.py:1002
      }
    }
    // Rows are reserved with ONE atomicAdd per warp at an arbitrary offset, then the warp's rays are
    // placed one after the other from that offset, skipping to the next 128-row tile whenever a ray
    // would straddle a tile boundary.  The reservation covers the worst case of that skipping:
    // every tile touched can waste at most kmax - 1 rows and therefore holds at least 129 - kmax.
    const int real = __reduce_add_sync(0xffffffffu, k);
    const int kmax = __reduce_max_sync(0xffffffffu, k);
    const int rows = real > 0 ? real + (real / (129 - kmax) + 2) * (kmax - 1) : 0;  // tiles used <= real/(129-kmax)+2
    int warp_base = 0;
    if (lane == 0 && rows > 0) {
      warp_base = atomicAdd(counters + 2, rows);
      atomicAdd(counters + 6, real);
    }
    warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
    if (warp_base + rows > s_cap) {  // capacity exceeded: drop this warp's samples and raise the flag
      if (lane == 0 && rows > 0) counters[5] = 1;
      continue;
    }
    int pos = 0, run = warp_base & 127;  // `run` is tracked relative to the start of warp_base's tile
    const int run0 = run;
    for (int l = 0; l < 32; ++l) {
      const int kl = __shfl_sync(0xffffffffu, k, l);
      if ((run & 127) + kl > 128) run = (run + 127) & ~127;
      if (l == lane) pos = run - run0;
      run += kl;
    }
    // rows after this lane's samples up to the next lane's first row (or the end of the run) are padding
    int next_pos = __shfl_down_sync(0xffffffffu, pos, 1);
    if (lane == 31) next_pos = rows;  // the unused tail of the reservation is padding too
    if (rows > 0) {
      const int base = warp_base + pos;
      const float4 pad = make_float4(0.5f, 0.5f, 0.5f, 0.f);
      if (k > 0) {
        const SamplePointer sp(rays_o, rays_d, ray, fc);
        for (int j = 0; j < k; ++j) {
          s_ray[base + j] = ray;
          s_cnt[base + j] = (j == 0) ? (uint8_t)k : (uint8_t)(0x80 | j);  // ray head: #samples; else: offset in ray
          s_ts[base + j] = ts[j];
          s_te[base + j] = te[j];
          s_x[base + j] = sp.at(ts[j], te[j]);
        }
      }
      for (int r = pos + k; r < next_pos; ++r) {
        s_ray[warp_base + r] = -1;
        s_cnt[warp_base + r] = 0;
        s_x[warp_base + r] = pad;
      }
      if (lane == 0)  // rows skipped before the first ray (it did not fit into the partly used tile)
        for (int r = 0; r < pos; ++r) {
          s_ray[warp_base + r] = -1;
          s_cnt[warp_base + r] = 0;
          s_x[warp_base + r] = pad;
        }
    }
  }
}

// Ordered compaction of the live-ray list by keep_flag (single-pass scan with decoupled look-back
// over chunks taken in ticket order), so the list stays sorted by ray id: neighbouring rays stay
// neighbours (hash-grid gather locality) and the result does not depend on block scheduling.  Also
// accumulates the per-call live counts the schedule kernel needs.  `chain` holds one status word per
// chunk: (generation << 34) | (kind << 32) | value, kind 1 = chunk aggregate, 2 = inclusive prefix;
// the generation is counters[7], bumped by the schedule kernel, so the array never needs clearing.
constexpr int CMP_T = 256, CMP_ITEMS = 8, CMP_CHUNK = CMP_T * CMP_ITEMS;
__global__ void __launch_bounds__(CMP_T) render_compact_kernel(const int* counters_in, int rays_per_call,
                                                               const int* __restrict__ alive,
                                                               const uint8_t* __restrict__ keep_flag,
                                                               int* __restrict__ alive_next, int* __restrict__ n_alive_acc,
                                                               unsigned long long* chain, int* counters) {
  __shared__ int s_ticket, s_base, s_warp[CMP_T / 32];
  const int n_live = counters_in[0];
  const unsigned long long gen = (unsigned long long)(unsigned int)counters_in[7];
  const int n_chunks = (n_live + CMP_CHUNK - 1) / CMP_CHUNK;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  while (true) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(counters + 4, 1);
    __syncthreads();
    const int ticket = s_ticket;
    if (ticket >= n_chunks) break;
    const int first = ticket * CMP_CHUNK + threadIdx.x * CMP_ITEMS;
    int rays[CMP_ITEMS];
    unsigned keepmask = 0;
#pragma unroll
    for (int j = 0; j < CMP_ITEMS; ++j) {
      const int idx = first + j;
      rays[j] = idx < n_live ? alive[idx] : -1;
      if (rays[j] >= 0 && keep_flag[rays[j]]) keepmask |= 1u << j;
    }
    const int cnt = __popc(keepmask);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int warp_off = 0, total = 0;
    for (int w = 0; w < CMP_T / 32; ++w) {
      if (w < warp) warp_off += s_warp[w];
      total += s_warp[w];
    }
    if (threadIdx.x == 0) {
      volatile unsigned long long* vchain = chain;
      vchain[ticket] = (gen << 34) | (1ull << 32) | (unsigned long long)total;  // publish the aggregate first
      __threadfence();
      long long prev = 0;
      for (int t = ticket - 1; t >= 0; --t) {  // look back until an inclusive prefix is found
        unsigned long long w;
        do { w = vchain[t]; } while ((w >> 34) != gen);
        prev += (long long)(w & 0xffffffffull);
        if (((w >> 32) & 3ull) == 2ull) break;
      }
      vchain[ticket] = (gen << 34) | (2ull << 32) | (unsigned long long)(prev + total);
      s_base = (int)prev;
      if (ticket == n_chunks - 1) counters[1] = (int)prev + total;
    }
    __syncthreads();
    int out = s_base + warp_off + inc - cnt;
#pragma unroll
    for (int j = 0; j < CMP_ITEMS; ++j)
      if (keepmask & (1u << j)) alive_next[out++] = rays[j];
    // per-call live counts (rays of one thread are consecutive list entries: mostly the same call)
    int cur_call = -1, cur_n = 0;
#pragma unroll
    for (int j = 0; j < CMP_ITEMS; ++j)
      if (keepmask & (1u << j)) {
        const int c = rays[j] / rays_per_call;
        if (c != cur_call) {
          if (cur_n) atomicAdd(n_alive_acc + cur_call, cur_n);
          cur_call = c, cur_n = 0;
        }
        ++cur_n;
      }
    if (cur_n) atomicAdd(n_alive_acc + cur_call, cur_n);
    __syncthreads();
  }
}

// One thread per live ray: transmittance weights of this iteration's samples (prefix = 1 - the
// accumulated opacity, utils.py:937-944), alpha_thre filter, accumulation of rgb / opacity / depth /
// semantic logits, then the variance terms against the UPDATED running rgb / depth
// (utils.py:957-999), next ray mask (utils.py:1004-1009) and compaction of the live list.
// (A 4-lanes-per-ray variant was measured slower: the weights' exp() calls were then issued by every
// lane and the kernel became MUFU-bound.)  The weights of the first pass are kept in a small per-thread
// array and the 32 semantic accumulators are processed 16 at a time, which lets the kernel run at 64
// registers (8 CTAs of 128 threads per SM instead of 3 at 168): ncu shows it memory-latency bound (long
// scoreboard 12 of 16 stall cycles per issue), so resident warps are what it needs; 48 registers measured slower.
// Sample row (80 B): [sigma, r, g, b as fp32, already activated by the field kernel | 32 sem logits fp16].
struct SampleTerms {
  float w, col[3], tmid;
  bool vis;
};

__device__ __forceinline__ SampleTerms sample_terms(const uint4& r0, float t0, float t1, float esum, float prefix,
                                                    float alpha_thre, float& sdt_out) {
  SampleTerms o;
  const float sigma = __uint_as_float(r0.x);  // exp(logit - 1) * selector, computed by the field kernel
  const float sdt = __fmul_rn(sigma, __fsub_rn(t1, t0));
  const float alpha = __fsub_rn(1.0f, expf(-sdt));
  o.w = __fmul_rn(__fmul_rn(expf(-esum), prefix), alpha);
  o.vis = !(alpha_thre > 0.f && !(alpha >= alpha_thre));
  o.tmid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  o.col[0] = __uint_as_float(r0.y), o.col[1] = __uint_as_float(r0.z), o.col[2] = __uint_as_float(r0.w);
  sdt_out = sdt;
  return o;
}

template <bool PROB>
__global__ void __launch_bounds__(128, 8) render_composite_kernel(
    const int* counters_in, int n_rays, int rays_per_call, int n_sem,
    const int* __restrict__ alive, const int* __restrict__ entry_base, const int* __restrict__ entry_cnt,
    const float* __restrict__ s_ts, const float* __restrict__ s_te, const uint4* __restrict__ rows,
    float* __restrict__ state, float alpha_thre,
    float opc_thre, const int* __restrict__ n_samp, const int* __restrict__ iter_samples, int max_samples,
    int* __restrict__ alive_next, int* __restrict__ n_alive_acc, int* __restrict__ total_samples,
    int* counters, int* __restrict__ ray_counts) {
  const int n_live = counters_in[0];
  const int lane = threadIdx.x & 31;
  const int n_round = (n_live + 31) & ~31;
  const size_t NR = (size_t)n_rays;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += blockDim.x * gridDim.x) {
    bool keep = false;
    int ray = -1, call = -1, n_vis = 0, sem_base = 0, sem_k = 0;
    float wloc[MAX_ITER_SAMPLES];  // weight of sample j, or -1 when filtered by alpha_thre
    if (i < n_live) {
      ray = alive[i];
      call = ray / rays_per_call;
      const int base = entry_base[i], k = entry_cnt[i];
      float* st = state + ray;
      float opac = st[ST_OPA * NR];
      if (k > 0) {
        // pass 1: weights (kept in a small per-thread array: registers stay low, occupancy high),
        // rgb / opacity / depth
        const float prefix = __fsub_rn(1.0f, opac);
        float rgb[3] = {st[0], st[NR], st[2 * NR]};
        float depth = st[ST_DEPTH * NR];
        float esum = 0.f;
        for (int j = 0; j < k; ++j) {
          const int s = base + j;
          const uint4 r0 = __ldg(rows + (size_t)s * 5);
          float sdt;
          const SampleTerms tm = sample_terms(r0, s_ts[s], s_te[s], esum, prefix, alpha_thre, sdt);
          esum = __fadd_rn(esum, sdt);
          wloc[j] = tm.vis ? tm.w : -1.0f;
          if (!tm.vis) continue;
          ++n_vis;
#pragma unroll
          for (int c = 0; c < 3; ++c) rgb[c] = __fadd_rn(rgb[c], __fmul_rn(tm.w, tm.col[c]));
          opac = __fadd_rn(opac, tm.w);
          depth = __fadd_rn(depth, __fmul_rn(tm.w, tm.tmid));
        }
        // a ray whose samples were all filtered by alpha_thre leaves its state untouched: skip the
        // read-modify-write of the 38 state planes (half of this kernel's traffic in free space)
        if (n_vis > 0) {
        if (PROB) {  // variance terms against the UPDATED running rgb / depth
          float rv[3] = {st[ST_RGBVAR * NR], st[(ST_RGBVAR + 1) * NR], st[(ST_RGBVAR + 2) * NR]};
          float dv = st[ST_DVAR * NR];
          for (int j = 0; j < k; ++j) {
            const float w = wloc[j];
            if (w < 0.f) continue;
            const int s = base + j;
            const uint4 r0 = __ldg(rows + (size_t)s * 5);
            const float cols[3] = {__uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w)};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float df = __fsub_rn(cols[c], rgb[c]);
              rv[c] = __fadd_rn(rv[c], __fmul_rn(w, __fmul_rn(df, df)));
            }
            const float dd = __fsub_rn(__fmul_rn(__fadd_rn(s_ts[s], s_te[s]), 0.5f), depth);
            dv = __fadd_rn(dv, __fmul_rn(w, __fmul_rn(dd, dd)));
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) st[(ST_RGBVAR + c) * NR] = rv[c];
          st[ST_DVAR * NR] = dv;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) st[c * NR] = rgb[c];
        st[ST_OPA * NR] = opac;
        st[ST_DEPTH * NR] = depth;
        }  // n_vis > 0
      }
      sem_base = base, sem_k = k;
      const int n = n_samp[call];
      keep = (n > 0) && (opac <= opc_thre) && (k == n) && (iter_samples[call] < max_samples);
      if (ray_counts) ray_counts[ray] += k, ray_counts[NR + ray] += n_vis;  // parity instrumentation (tests)
    }
    // pass 2: semantic logits of the rays that composited something.  In the long tail of a render most live rays
    // march through transparent space and only a few lanes of a warp have visible samples: looping over those rays
    // with the WARP (8 lanes x 4 class channels per ray, four rays at a time) instead of letting each such lane run the
    // 32-channel loop alone keeps the lanes busy there.  Dense warps (the first iterations) keep the per-thread form,
    // 16 channels at a time, whose state accesses are coalesced across rays.  Same sums, same order, either way.
    const bool has_sem = n_vis > 0 && n_sem > 0;
    const unsigned sem_mask = __ballot_sync(0xffffffffu, has_sem);
    if (__popc(sem_mask) > 16) {
      if (has_sem) {
        float* st = state + ray;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          if (16 * g >= n_sem) break;
          float acc[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = (16 * g + c < n_sem) ? st[(ST_SEM + 16 * g + c) * NR] : 0.f;
          for (int j = 0; j < sem_k; ++j) {
            const float w = wloc[j];
            if (w < 0.f) continue;
            __align__(16) __half h[16];
            reinterpret_cast<uint4*>(h)[0] = __ldg(rows + (size_t)(sem_base + j) * 5 + 1 + 2 * g);
            reinterpret_cast<uint4*>(h)[1] = __ldg(rows + (size_t)(sem_base + j) * 5 + 2 + 2 * g);
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(w, __half2float(h[c])));
          }
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (16 * g + c < n_sem) st[(ST_SEM + 16 * g + c) * NR] = acc[c];
        }
      }
    } else if (sem_mask) {
      // four rays per warp step: 8 lanes x 4 channels each, so four rays' state / row loads are in flight together
      const __half* rows_h = reinterpret_cast<const __half*>(rows);
      const int grp = lane >> 3, sub = lane & 7, n_act = __popc(sem_mask);
      for (int t0 = 0; t0 < n_act; t0 += 4) {
        const bool on = t0 + grp < n_act;
        const int owner = on ? (int)__fns(sem_mask, 0, t0 + grp + 1) : 0;  // lane of this group's ray
        const int rb = __shfl_sync(0xffffffffu, ray, owner), bb = __shfl_sync(0xffffffffu, sem_base, owner);
        const int kb_owner = __shfl_sync(0xffffffffu, sem_k, owner);
        const int kb = on ? kb_owner : 0;
        const int kmax = __reduce_max_sync(0xffffffffu, kb);
        float* sp = state + (size_t)(ST_SEM + 4 * sub) * NR + (on ? rb : 0);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (on) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (4 * sub + c < n_sem) acc[c] = sp[(size_t)c * NR];
        }
        for (int j = 0; j < kmax; ++j) {
          const float w = __shfl_sync(0xffffffffu, wloc[j], owner);
          if (j < kb && w >= 0.f) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(rows_h + (size_t)(bb + j) * 40 + 8 + 4 * sub));
            const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
            const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
            acc[0] = __fadd_rn(acc[0], __fmul_rn(w, f01.x));
            acc[1] = __fadd_rn(acc[1], __fmul_rn(w, f01.y));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(w, f23.x));
            acc[3] = __fadd_rn(acc[3], __fmul_rn(w, f23.y));
          }
        }
        if (on) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (4 * sub + c < n_sem) sp[(size_t)c * NR] = acc[c];
        }
      }
    }
    // compaction of the live list (order-preserving inside a warp) + per-call live counts
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    int warp_base = 0;
    if (lane == 0 && ballot) warp_base = atomicAdd(counters + 1, __popc(ballot));
    warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
    if (keep) alive_next[warp_base + __popc(ballot & ((1u << lane) - 1u))] = ray;
    const unsigned peers = __match_any_sync(0xffffffffu, call);
    const int vis_sum = __reduce_add_sync(peers, n_vis);
    if (call >= 0 && lane == __ffs(peers) - 1) {
      const int kept = __popc(ballot & peers);
      if (kept) atomicAdd(n_alive_acc + call, kept);
      if (vis_sum) atomicAdd(total_samples + call, vis_sum);
    }
  }
}

// rgb += bkgd * (1 - opacity); depth /= max(opacity, eps)  (utils.py:1012-1013) and
// structure-of-arrays -> the reference's [n_rays, D] outputs.
__global__ void __launch_bounds__(256) render_finalize_kernel(int n_rays, int n_sem, const float* __restrict__ state,
                                                              float b0, float b1, float b2, float* __restrict__ rgb,
                                                              float* __restrict__ rgb_var, float* __restrict__ opacity,
                                                              float* __restrict__ depth, float* __restrict__ depth_var,
                                                              float* __restrict__ sem) {
  const size_t NR = (size_t)n_rays;
  const float bk[3] = {b0, b1, b2};
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += blockDim.x * gridDim.x) {
    const float op = state[ST_OPA * NR + r];
    if (rgb)
      for (int c = 0; c < 3; ++c) rgb[3 * r + c] = __fadd_rn(state[c * NR + r], __fmul_rn(bk[c], __fsub_rn(1.0f, op)));
    if (rgb_var)
      for (int c = 0; c < 3; ++c) rgb_var[3 * r + c] = state[(ST_RGBVAR + c) * NR + r];
    if (opacity) opacity[r] = op;
    if (depth) depth[r] = __fdiv_rn(state[ST_DEPTH * NR + r], fmaxf(op, 1.1920928955078125e-07f));
    if (depth_var) depth_var[r] = state[ST_DVAR * NR + r];
    if (sem)
      for (int c = 0; c < n_sem; ++c) sem[(size_t)r * n_sem + c] = state[(ST_SEM + c) * NR + r];
  }
}

// Predictive information of an ensemble's renders (scripts/pipeline.py:727-781), reduced to four
// sums per trajectory:
//   [0] sum over pixels x 3 channels of  H(ensemble rgb var) - mean_m H(rgb var_m)
//   [1] sum over pixels of the same for depth
//   [2] sum over pixels of  H(mean_m softmax) - mean_m H(softmax_m)
//   [3] sum over pixels of  Hb(mean_m acc) - mean_m Hb(acc_m)
// The host divides by the element counts and applies the x3 / x2 weights (pipeline.py:772-790).
// The reference does this in float64 numpy on float32 renders; here the per-pixel entropies use
// full-precision fp32 logf / expf (<= 2 ulp, i.e. <= ~2e-6 absolute per pixel term, far inside the
// 1e-3 entropy tolerance) and every accumulation across pixels is float64.
struct EnsembleStates {
  const float* s[4];
};

__device__ __forceinline__ float gauss_entropy(float var) { return 0.5f * logf(17.079468445347132f * var + 1e-4f); }
__device__ __forceinline__ float plogp(float p) { return (p + 1e-4f) * logf(p + 1e-4f); }
__device__ __forceinline__ float bern_entropy(float a) { return -plogp(a) - plogp(1.0f - a); }

__global__ void __launch_bounds__(128) score_views_kernel(int n_members, EnsembleStates es, int n_rays,
                                                          int rays_per_view, int n_sem,
                                                          const int* __restrict__ view_traj, int n_traj,
                                                          double* __restrict__ sums) {
  extern __shared__ double acc_s[];  // [n_traj][4]
  for (int i = threadIdx.x; i < n_traj * 4; i += blockDim.x) acc_s[i] = 0.0;
  __syncthreads();
  const size_t NR = (size_t)n_rays;
  const float inv_m = 1.0f / (float)n_members;
  int cur = -1;
  double a_rgb = 0.0, a_depth = 0.0, a_sem = 0.0, a_occ = 0.0;
  auto flush = [&]() {
    if (cur >= 0) {
      atomicAdd(&acc_s[cur * 4 + 0], a_rgb);
      atomicAdd(&acc_s[cur * 4 + 1], a_depth);
      atomicAdd(&acc_s[cur * 4 + 2], a_sem);
      atomicAdd(&acc_s[cur * 4 + 3], a_occ);
    }
    a_rgb = a_depth = a_sem = a_occ = 0.0;
  };
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rays; r += blockDim.x * gridDim.x) {
    const int traj = view_traj[r / rays_per_view];
    if (traj < 0) continue;
    if (traj != cur) {
      flush();
      cur = traj;
    }
    float t_rgb = 0.f, t_depth, t_sem = 0.f, t_occ;
    for (int c = 0; c < 3; ++c) {
      float vs = 0.f, hs = 0.f;
      for (int m = 0; m < n_members; ++m) {
        const float v = es.s[m][(ST_RGBVAR + c) * NR + r];
        vs += v;
        hs += gauss_entropy(v);
      }
      t_rgb += gauss_entropy(vs * 0.5f) - hs * inv_m;  // pipeline.py:733 hard-codes "/ 2"
    }
    {
      float vs = 0.f, hs = 0.f;
      for (int m = 0; m < n_members; ++m) {
        const float v = es.s[m][ST_DVAR * NR + r];
        vs += v;
        hs += gauss_entropy(v);
      }
      t_depth = gauss_entropy(vs * 0.5f) - hs * inv_m;
    }
    if (n_sem > 0) {
      float pm[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) pm[c] = 0.f;
      float hs = 0.f;
      for (int m = 0; m < n_members; ++m) {
        const float* sp = es.s[m] + ST_SEM * NR + r;
        float lg[32];
        float mx = -3.0e38f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          lg[c] = (c < n_sem) ? sp[c * NR] : -3.0e38f;
          mx = fmaxf(mx, lg[c]);
        }
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          lg[c] = (c < n_sem) ? expf(lg[c] - mx) : 0.f;
          den += lg[c];
        }
        const float inv_den = 1.0f / den;
        float h = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < n_sem) {
            const float p = lg[c] * inv_den;
            pm[c] += p;
            h -= plogp(p);
          }
        hs += h;
      }
      float he = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < n_sem) he -= plogp(pm[c] * inv_m);
      t_sem = he - hs * inv_m;
    }
    {
      float as = 0.f, hs = 0.f;
      for (int m = 0; m < n_members; ++m) {
        const float a = es.s[m][ST_OPA * NR + r];
        as += a;
        hs += bern_entropy(a);
      }
      t_occ = bern_entropy(as * inv_m) - hs * inv_m;
    }
    a_rgb += (double)t_rgb, a_depth += (double)t_depth, a_sem += (double)t_sem, a_occ += (double)t_occ;
  }
  flush();
  __syncthreads();
  for (int i = threadIdx.x; i < n_traj * 4; i += blockDim.x)
    if (acc_s[i] != 0.0) atomicAdd(sums + i, acc_s[i]);
}

}  // namespace apnerf

using namespace apnerf;

APNERF_API int apnerf_generate_rays(int n_views, const float* c2w, int width, int height, float focal, int n_keep,
                                    const int* keep_idx, float* rays_o, float* rays_d, void* stream) {
  const long long total = (long long)n_views * n_keep;
  if (total == 0) return 0;
  generate_rays_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(n_views, c2w, width, height, focal,
                                                                                  n_keep, keep_idx, rays_o, rays_d);
  APNERF_CHECK_LAUNCH("generate_rays_kernel");
  return 0;
}

APNERF_API int apnerf_render_init(int n_rays, int rays_per_call, const float* rays_o, const float* rays_d, int rx,
                                  int ry, int rz, const uint8_t* binaries, const float* aabbs, float near_plane,
                                  int n_state, float* state, float* t_min, float* t_max, uint8_t* hit, float* near,
                                  int* alive, int* n_alive_acc, int* iter_samples, int* total_samples, int n_calls,
                                  int* counters, uint32_t* occ_bits, void* stream) {
  if (n_rays == 0) return 0;
  GridView g{binaries, aabbs, 1, rx, ry, rz, apnerf_skip_min_steps()};
  if (occ_bits) {
    const int n_cells = rx * ry * rz;
    pack_occupancy_kernel<<<grid_for((n_cells + 31) / 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(n_cells, binaries, occ_bits);
    APNERF_CHECK_LAUNCH("pack_occupancy_kernel");
  }
  render_init_kernel<<<grid_for(n_rays, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, rays_per_call, rays_o, rays_d, g, near_plane, n_state, state, t_min, t_max, hit, near, alive,
      n_alive_acc, iter_samples, total_samples, n_calls, counters);
  APNERF_CHECK_LAUNCH("render_init_kernel");
  return 0;
}

APNERF_API int apnerf_render_schedule(int n_calls, int rays_per_call, int max_samples, int min_samples,
                                      int* n_alive_acc, int* n_samp, int* iter_samples, int* counters, int* call_rows,
                                      void* stream) {
  render_schedule_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n_calls, rays_per_call, max_samples, min_samples,
                                                             n_alive_acc, n_samp, iter_samples, counters, call_rows);
  APNERF_CHECK_LAUNCH("render_schedule_kernel");
  return 0;
}

APNERF_API int apnerf_render_march(int max_live, int rays_per_call, const int* alive, const int* n_samp,
                                   const float* rays_o, const float* rays_d, int rx, int ry, int rz,
                                   const uint8_t* binaries, const float* aabbs, const float* t_min, const float* t_max,
                                   const uint8_t* hit, float* near, float far_plane, float step_size, float cone_angle,
                                   int* entry_base, int* entry_cnt, int* s_ray, float* s_ts, float* s_te,
                                   const float* field_aabb_host, void* s_x, int* counters, const uint32_t* occ_bits,
                                   void* stream) {
  if (max_live == 0) return 0;
  GridView g{binaries, aabbs, 1, rx, ry, rz, apnerf_skip_min_steps(), occ_bits};
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = field_aabb_host[i];
  int mt, mc;
  apnerf_march_cfg(mt, mc);
  auto kernel = (step_size > 0.0f && occ_bits) ? render_march_kernel<true> : render_march_kernel<false>;
  kernel<<<grid_for(max_live, mt, mc), mt, 0, (cudaStream_t)stream>>>(
      counters, rays_per_call, alive, n_samp, rays_o, rays_d, g, t_min, t_max, hit, near, far_plane, step_size,
      cone_angle, entry_base, entry_cnt, s_ray, s_ts, s_te, fc, (float4*)s_x, counters);
  APNERF_CHECK_LAUNCH("render_march_kernel");
  return 0;
}


APNERF_API int apnerf_render_march_tiles(int max_live, int rays_per_call, const int* alive, const int* n_samp,
                                         const float* rays_o, const float* rays_d, int rx, int ry, int rz,
                                         const uint8_t* binaries, const float* aabbs, const float* t_min,
                                         const float* t_max, const uint8_t* hit, float* near, float far_plane,
                                         float step_size, float cone_angle, int* s_ray, uint8_t* s_cnt, float* s_ts,
                                         float* s_te, const float* field_aabb_host, void* s_x, uint8_t* keep_flag,
                                         int s_cap, int* counters, const uint32_t* occ_bits, void* stream) {
  if (max_live == 0) return 0;
  GridView g{binaries, aabbs, 1, rx, ry, rz, apnerf_skip_min_steps(), occ_bits};
  FieldConst fc;
  for (int i = 0; i < 6; ++i) fc.aabb[i] = field_aabb_host[i];
  int mt, mc;
  apnerf_march_cfg(mt, mc);
  auto kernel = (step_size > 0.0f && occ_bits) ? render_march_tiles_kernel<true> : render_march_tiles_kernel<false>;
  kernel<<<grid_for(max_live, mt, mc), mt, 0, (cudaStream_t)stream>>>(
      counters, rays_per_call, alive, n_samp, rays_o, rays_d, g, t_min, t_max, hit, near, far_plane, step_size,
      cone_angle, s_ray, s_cnt, s_ts, s_te, fc, (float4*)s_x, keep_flag, s_cap, counters);
  APNERF_CHECK_LAUNCH("render_march_tiles_kernel");
  return 0;
}

APNERF_API int apnerf_render_compact(int max_live, int rays_per_call, const int* alive, const uint8_t* keep_flag,
                                     int* alive_next, int* n_alive_acc, void* chain, int* counters, void* stream) {
  if (max_live == 0) return 0;
  const int chunks = (max_live + CMP_CHUNK - 1) / CMP_CHUNK;
  const int cap = apnerf_num_sms() * 4;
  render_compact_kernel<<<chunks < cap ? chunks : cap, CMP_T, 0, (cudaStream_t)stream>>>(
      counters, rays_per_call, alive, keep_flag, alive_next, n_alive_acc, (unsigned long long*)chain, counters);
  APNERF_CHECK_LAUNCH("render_compact_kernel");
  return 0;
}

APNERF_API int apnerf_render_composite(int max_live, int n_rays, int rays_per_call, int n_sem,
                                       const int* alive, const int* entry_base, const int* entry_cnt,
                                       const float* s_ts, const float* s_te, const void* rows, float* state,
                                       float alpha_thre, float opc_thre,
                                       const int* n_samp, const int* iter_samples, int max_samples, int* alive_next,
                                       int* n_alive_acc, int* total_samples, int* counters, int probabilistic,
                                       int* ray_counts, void* stream) {
  if (max_live == 0) return 0;
  APNERF_REQUIRE(n_sem >= 0 && n_sem <= 32, "render_composite: at most 32 semantic classes");
  const int grid = grid_for(max_live, 128, 16);
  if (probabilistic)
    render_composite_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(
        counters, n_rays, rays_per_call, n_sem, alive, entry_base, entry_cnt, s_ts, s_te, (const uint4*)rows,
        state, alpha_thre, opc_thre, n_samp, iter_samples, max_samples, alive_next, n_alive_acc, total_samples, counters,
        ray_counts);
  else
    render_composite_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(
        counters, n_rays, rays_per_call, n_sem, alive, entry_base, entry_cnt, s_ts, s_te, (const uint4*)rows,
        state, alpha_thre, opc_thre, n_samp, iter_samples, max_samples, alive_next, n_alive_acc, total_samples, counters,
        ray_counts);
  APNERF_CHECK_LAUNCH("render_composite_kernel");
  return 0;
}

APNERF_API int apnerf_render_finalize(int n_rays, int n_sem, const float* state, float bkgd_r, float bkgd_g,
                                      float bkgd_b, float* rgb, float* rgb_var, float* opacity, float* depth,
                                      float* depth_var, float* sem, void* stream) {
  if (n_rays == 0) return 0;
  render_finalize_kernel<<<grid_for(n_rays, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, n_sem, state, bkgd_r, bkgd_g, bkgd_b, rgb, rgb_var, opacity, depth, depth_var, sem);
  APNERF_CHECK_LAUNCH("render_finalize_kernel");
  return 0;
}

APNERF_API int apnerf_score_views(int n_members, const float* state0, const float* state1, const float* state2,
                                  const float* state3, int n_rays, int rays_per_view, int n_sem,
                                  const int* view_traj, int n_traj, double* sums, void* stream) {
  if (n_rays == 0) return 0;
  APNERF_REQUIRE(n_members >= 1 && n_members <= 4, "score_views: 1..4 ensemble members");
  APNERF_REQUIRE(n_traj >= 1 && n_traj <= 1024, "score_views: 1..1024 trajectories");
  EnsembleStates es{{state0, state1, state2, state3}};
  score_views_kernel<<<grid_for(n_rays, 128, 8), 128, n_traj * 4 * sizeof(double), (cudaStream_t)stream>>>(
      n_members, es, n_rays, rays_per_view, n_sem, view_traj, n_traj, sums);
  APNERF_CHECK_LAUNCH("score_views_kernel");
  return 0;
}
