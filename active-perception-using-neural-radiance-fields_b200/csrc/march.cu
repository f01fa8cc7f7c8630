// C-ABI entry points for ray/AABB intersection and occupancy-grid traversal (kernel 1).
// Drop-in for the reference's nerfacc_cuda.ray_aabb_intersect / traverse_grids
// (perception/nerfacc/nerfacc/cuda/csrc/grid.cu:284-313, 68-282, host code :320-519).
#include "march.cuh"

namespace apnerf {

__global__ void __launch_bounds__(256) ray_aabb_kernel(int n_rays, const float* __restrict__ rays_o,
                                                       const float* __restrict__ rays_d, float near, float far,
                                                       int n_aabbs, const float* __restrict__ aabbs, float miss,
                                                       float* __restrict__ t_mins, float* __restrict__ t_maxs,
                                                       uint8_t* __restrict__ hits) {
  const long long numel = (long long)n_rays * n_aabbs;
  for (long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x; tid < numel;
       tid += (long long)blockDim.x * gridDim.x) {
    const int r = (int)(tid / n_aabbs), b = (int)(tid % n_aabbs);
    const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
    const float inv[3] = {__frcp_rn(rays_d[3 * r]), __frcp_rn(rays_d[3 * r + 1]), __frcp_rn(rays_d[3 * r + 2])};
    float t0, t1;
    const bool hit = ray_aabb(o, inv, near, far, aabbs + 6 * b, t0, t1);
    t_mins[tid] = hit ? t0 : miss;
    t_maxs[tid] = hit ? t1 : miss;
    hits[tid] = hit ? 1 : 0;
  }
}

struct Segments {  // mirrors RaySegmentsSpec (csrc/include/data_spec.hpp:6-14); pointers may be null
  float* vals;
  int64_t* ray_indices;
  uint8_t* is_left;
  uint8_t* is_right;
  uint8_t* is_valid;
  const int64_t* chunk_starts;
  int64_t* chunk_cnts;
};

struct FillSink {
  Segments iv, sm;
  int64_t iv_base, sm_base;
  int64_t tid;
  bool write;
  __device__ __forceinline__ void operator()(float t_last, float t_next, bool continuous, int i_sample, int i_edge) {
    if (!write) return;
    if (iv.chunk_cnts) {  // grid.cu:219-246
      int64_t idx = iv_base + i_edge;
      if (!continuous) {
        iv.vals[idx] = t_last;
        iv.ray_indices[idx] = tid;
        iv.is_left[idx] = 1;
        ++idx;
        iv.vals[idx] = t_next;
        iv.ray_indices[idx] = tid;
        iv.is_right[idx] = 1;
      } else {
        iv.vals[idx] = t_next;
        iv.ray_indices[idx] = tid;
        iv.is_left[idx - 1] = 1;
        iv.is_right[idx] = 1;
      }
    }
    if (sm.chunk_cnts) {  // grid.cu:249-256
      const int64_t idx = sm_base + i_sample;
      sm.vals[idx] = __fmul_rn(__fadd_rn(t_next, t_last), 0.5f);
      sm.ray_indices[idx] = tid;
      sm.is_valid[idx] = 1;
    }
  }
};

__global__ void __launch_bounds__(256) traverse_kernel(int n_rays, const float* __restrict__ rays_o,
                                                       const float* __restrict__ rays_d,
                                                       const uint8_t* __restrict__ rays_mask, GridView g,
                                                       const uint8_t* __restrict__ hits,
                                                       const float* __restrict__ t_sorted,
                                                       const int64_t* __restrict__ t_indices,
                                                       const float* __restrict__ near_planes,
                                                       const float* __restrict__ far_planes, float step_size,
                                                       float cone_angle, int limit, bool first_pass, Segments iv,
                                                       Segments sm, float* __restrict__ terminate_planes) {
  for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n_rays; tid += blockDim.x * gridDim.x) {
    if (rays_mask && !rays_mask[tid]) continue;  // grid.cu:100
    if (iv.chunk_cnts && !first_pass && iv.chunk_cnts[tid] == 0) continue;
    if (sm.chunk_cnts && !first_pass && sm.chunk_cnts[tid] == 0) continue;
    FillSink sink;
    sink.iv = iv;
    sink.sm = sm;
    sink.tid = tid;
    sink.write = !first_pass;
    sink.iv_base = (!first_pass && iv.chunk_cnts) ? iv.chunk_starts[tid] : 0;
    sink.sm_base = (!first_pass && sm.chunk_cnts) ? sm.chunk_starts[tid] : 0;
    const float o[3] = {rays_o[3 * tid], rays_o[3 * tid + 1], rays_o[3 * tid + 2]};
    const float d[3] = {rays_d[3 * tid], rays_d[3 * tid + 1], rays_d[3 * tid + 2]};
    int n_intervals;
    float t_term;
    const int n_samples =
        march_ray(g, o, d, near_planes[tid], far_planes[tid], hits + (size_t)tid * g.n_grids,
                  t_sorted + (size_t)tid * g.n_grids * 2, t_indices ? t_indices + (size_t)tid * g.n_grids * 2 : nullptr,
                  step_size, cone_angle, limit, sink, n_intervals, t_term);
    if (terminate_planes) terminate_planes[tid] = t_term;
    if (iv.chunk_cnts) iv.chunk_cnts[tid] = n_intervals;
    if (sm.chunk_cnts) sm.chunk_cnts[tid] = n_samples;
  }
}

// OccGridEstimator.sampling's use of the traversal (occ_grid.py:117-131): only the samples' (ray, t_start, t_end)
// triples are needed there, i.e. intervals.vals[is_left] / [is_right] of the reference -- written directly, without the
// interval edges, masks and the two boolean compactions (each a host synchronisation) in between.
struct PairSink {
  int64_t* ray_indices;
  float* t_starts;
  float* t_ends;
  int64_t base, tid;
  bool write;
  __device__ __forceinline__ void operator()(float t_last, float t_next, bool, int i_sample, int) {
    if (!write) return;
    ray_indices[base + i_sample] = tid;
    t_starts[base + i_sample] = t_last;
    t_ends[base + i_sample] = t_next;
  }
};

__global__ void __launch_bounds__(256) sample_rays_kernel(int n_rays, const float* __restrict__ rays_o,
                                                          const float* __restrict__ rays_d, GridView g,
                                                          const uint8_t* __restrict__ hits,
                                                          const float* __restrict__ t_sorted,
                                                          const int64_t* __restrict__ t_indices,
                                                          const float* __restrict__ near_planes,
                                                          const float* __restrict__ far_planes, float step_size,
                                                          float cone_angle, int limit,
                                                          const int64_t* __restrict__ chunk_starts,
                                                          int64_t* __restrict__ chunk_cnts,
                                                          int64_t* __restrict__ ray_indices,
                                                          float* __restrict__ t_starts, float* __restrict__ t_ends) {
  const bool fill = chunk_starts != nullptr;
  for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n_rays; tid += blockDim.x * gridDim.x) {
    if (fill && chunk_cnts[tid] == 0) continue;
    PairSink sink{ray_indices, t_starts, t_ends, fill ? chunk_starts[tid] : 0, tid, fill};
    const float o[3] = {rays_o[3 * tid], rays_o[3 * tid + 1], rays_o[3 * tid + 2]};
    const float d[3] = {rays_d[3 * tid], rays_d[3 * tid + 1], rays_d[3 * tid + 2]};
    int n_intervals;
    float t_term;
    const int n_samples =
        march_ray(g, o, d, near_planes[tid], far_planes[tid], hits + (size_t)tid * g.n_grids,
                  t_sorted + (size_t)tid * g.n_grids * 2, t_indices ? t_indices + (size_t)tid * g.n_grids * 2 : nullptr,
                  step_size, cone_angle, limit, sink, n_intervals, t_term);
    if (!fill) chunk_cnts[tid] = n_samples;
  }
}

// ---- int64 exclusive scan (chunk_cnts -> chunk_starts + total), replaces the reference's
// torch::cumsum in RaySegmentsSpec::memalloc_data_from_chunk (data_spec.hpp:86-96).
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* smem, long long& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long w = (lane < SCAN_T / 32) ? smem[lane] : 0, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    if (lane < SCAN_T / 32) smem[lane] = winc - w;
    if (lane == SCAN_T / 32 - 1) smem[32] = winc;
  }
  __syncthreads();
  total = smem[32];
  const long long r = inc - v + smem[warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_sums(long long n, const int64_t* __restrict__ in,
                                                         int64_t* __restrict__ tile_sums) {
  __shared__ long long smem[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  long long s = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const long long idx = base + (long long)threadIdx.x * SCAN_ITEMS + i;
    if (idx < n) s += in[idx];
  }
  long long total;
  block_exclusive_scan(s, smem, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_offsets(int n_tiles, int64_t* __restrict__ tile_sums,
                                                            int64_t* __restrict__ total_out) {
  __shared__ long long smem[33];
  long long carry = 0;
  for (int base = 0; base < n_tiles; base += SCAN_T) {
    const int idx = base + threadIdx.x;
    const long long v = idx < n_tiles ? tile_sums[idx] : 0;
    long long total;
    const long long ex = block_exclusive_scan(v, smem, total);
    if (idx < n_tiles) tile_sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_T) scan_apply(long long n, const int64_t* __restrict__ in,
                                                     const int64_t* __restrict__ tile_offsets,
                                                     int64_t* __restrict__ out) {
  __shared__ long long smem[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  long long v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  long long total;
  long long ex = block_exclusive_scan(s, smem, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

}  // namespace apnerf

using namespace apnerf;

APNERF_API int apnerf_ray_aabb_intersect(int n_rays, const float* rays_o, const float* rays_d, int n_aabbs,
                                         const float* aabbs, float near_plane, float far_plane, float miss_value,
                                         float* t_mins, float* t_maxs, uint8_t* hits, void* stream) {
  const long long numel = (long long)n_rays * n_aabbs;
  if (numel == 0) return 0;
  ray_aabb_kernel<<<grid_for(numel, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, rays_o, rays_d, near_plane, far_plane, n_aabbs, aabbs, miss_value, t_mins, t_maxs, hits);
  APNERF_CHECK_LAUNCH("ray_aabb_kernel");
  return 0;
}

APNERF_API int apnerf_traverse_grids(int n_rays, const float* rays_o, const float* rays_d, const uint8_t* rays_mask,
                                     int n_grids, int rx, int ry, int rz, const uint8_t* binaries, const float* aabbs,
                                     const uint8_t* hits, const float* t_sorted, const int64_t* t_indices,
                                     const float* near_planes, const float* far_planes, float step_size,
                                     float cone_angle, int traverse_steps_limit, int first_pass,
                                     float* iv_vals, int64_t* iv_ray_indices, uint8_t* iv_is_left, uint8_t* iv_is_right,
                                     const int64_t* iv_chunk_starts, int64_t* iv_chunk_cnts,
                                     float* sm_vals, int64_t* sm_ray_indices, uint8_t* sm_is_valid,
                                     const int64_t* sm_chunk_starts, int64_t* sm_chunk_cnts,
                                     float* terminate_planes, void* stream) {
  if (n_rays == 0) return 0;
  APNERF_REQUIRE(n_grids >= 1 && n_grids <= 8, "traverse_grids: n_grids must be in [1, 8]");
  GridView g{binaries, aabbs, n_grids, rx, ry, rz, apnerf_skip_min_steps()};
  Segments iv{iv_vals, iv_ray_indices, iv_is_left, iv_is_right, nullptr, iv_chunk_starts, iv_chunk_cnts};
  Segments sm{sm_vals, sm_ray_indices, nullptr, nullptr, sm_is_valid, sm_chunk_starts, sm_chunk_cnts};
  traverse_kernel<<<grid_for(n_rays, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, rays_o, rays_d, rays_mask, g, hits, t_sorted, t_indices, near_planes, far_planes, step_size,
      cone_angle, traverse_steps_limit, first_pass != 0, iv, sm, terminate_planes);
  APNERF_CHECK_LAUNCH("traverse_kernel");
  return 0;
}

// Samples of every ray as packed (ray_indices, t_starts, t_ends): chunk_starts == NULL counts (chunk_cnts written),
// otherwise fills at chunk_starts[ray].
APNERF_API int apnerf_sample_rays(int n_rays, const float* rays_o, const float* rays_d, int n_grids, int rx, int ry,
                                  int rz, const uint8_t* binaries, const float* aabbs, const uint8_t* hits,
                                  const float* t_sorted, const int64_t* t_indices, const float* near_planes,
                                  const float* far_planes, float step_size, float cone_angle, int traverse_steps_limit,
                                  const int64_t* chunk_starts, int64_t* chunk_cnts, int64_t* ray_indices,
                                  float* t_starts, float* t_ends, void* stream) {
  if (n_rays == 0) return 0;
  APNERF_REQUIRE(n_grids >= 1 && n_grids <= 8, "sample_rays: n_grids must be in [1, 8]");
  APNERF_REQUIRE(chunk_cnts != nullptr, "sample_rays: chunk_cnts is required");
  APNERF_REQUIRE(chunk_starts == nullptr || (ray_indices && t_starts && t_ends), "sample_rays: null output");
  GridView g{binaries, aabbs, n_grids, rx, ry, rz, apnerf_skip_min_steps()};
  sample_rays_kernel<<<grid_for(n_rays, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, rays_o, rays_d, g, hits, t_sorted, t_indices, near_planes, far_planes, step_size, cone_angle,
      traverse_steps_limit, chunk_starts, chunk_cnts, ray_indices, t_starts, t_ends);
  APNERF_CHECK_LAUNCH("sample_rays_kernel");
  return 0;
}

// out[i] = sum_{j<i} in[i]; *total = sum of all.  `scratch` needs ceil(n / 2048) int64 slots.
APNERF_API int apnerf_exclusive_scan_i64(long long n, const int64_t* in, int64_t* out, int64_t* total,
                                         int64_t* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (total) APNERF_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st));
    return 0;
  }
  const int n_tiles = ceil_div_i(n, SCAN_TILE);
  scan_tile_sums<<<n_tiles, SCAN_T, 0, st>>>(n, in, scratch);
  scan_tile_offsets<<<1, SCAN_T, 0, st>>>(n_tiles, scratch, total);
  scan_apply<<<n_tiles, SCAN_T, 0, st>>>(n, in, scratch, out);
  APNERF_CHECK_LAUNCH("exclusive_scan_i64");
  return 0;
}

APNERF_API long long apnerf_scan_scratch_elems(long long n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }
