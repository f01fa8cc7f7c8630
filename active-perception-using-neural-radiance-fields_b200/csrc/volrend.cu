// Packed (per-ray segmented) volume-rendering ops (kernel 4, standalone form).
//
// Replaces, on the reference side:
//   exclusive_sum / inclusive_sum        perception/nerfacc/nerfacc/cuda/csrc/scan.cu:9-125,
//                                        include/utils_scan.cuh:21-263  (fwd + reverse "backward")
//   render_weight_from_density           perception/nerfacc/nerfacc/volrend.py:212-267,315-365
//   accumulate_along_rays(_)             perception/nerfacc/nerfacc/volrend.py:486-576 (index_add_)
//   pack_info                            perception/nerfacc/nerfacc/pack.py:10-49
//
// A warp owns one ray: its samples are contiguous in the packed arrays, so loads are
// coalesced 128-byte lines and the transmittance scan is a shuffle scan with a carried
// running total -- no shared memory, no block barriers.
#include "common.cuh"

namespace apnerf {

__device__ __forceinline__ float warp_inclusive_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = __fadd_rn(v, n);
  }
  return v;
}

template <bool INCLUSIVE, bool BACKWARD>
__global__ void __launch_bounds__(256) packed_sum_kernel(int n_rays, const int64_t* __restrict__ starts,
                                                         const int64_t* __restrict__ cnts,
                                                         const float* __restrict__ in, float* __restrict__ out,
                                                         bool normalize) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t s = starts[r], c = cnts[r];
    float carry = 0.f;
    for (int64_t base = 0; base < c; base += 32) {
      const int64_t k = base + lane;  // position in scan order
      const int64_t idx = BACKWARD ? (s + c - 1 - k) : (s + k);
      const float v = (k < c) ? in[idx] : 0.f;
      const float inc = warp_inclusive_sum(v, lane);
      // exclusive value = the left neighbour's inclusive value (not inc - v, which re-rounds)
      float ex = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) ex = 0.f;
      const float o = __fadd_rn(carry, INCLUSIVE ? inc : ex);
      if (k < c) out[idx] = o;
      carry = __fadd_rn(carry, __shfl_sync(0xffffffffu, inc, 31));
    }
    if (normalize && c > 0) {  // utils_scan.cuh normalize: divide by the ray total
      for (int64_t k = lane; k < c; k += 32) {
        const int64_t idx = BACKWARD ? (s + c - 1 - k) : (s + k);
        out[idx] = __fdiv_rn(out[idx], carry);
      }
    }
  }
}

// weights / transmittance / alphas from density, one warp per ray.
//   sdt = sigma * (t_end - t_start);  alpha = 1 - exp(-sdt);
//   trans = exp(-exclusive_sum(sdt)) [* prefix_trans];  weight = trans * alpha
__global__ void __launch_bounds__(256) weights_from_density_kernel(
    int n_rays, const int64_t* __restrict__ starts, const int64_t* __restrict__ cnts,
    const float* __restrict__ t_starts, const float* __restrict__ t_ends, const float* __restrict__ sigmas,
    const float* __restrict__ prefix_trans, float* __restrict__ weights, float* __restrict__ trans,
    float* __restrict__ alphas) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t s = starts[r], c = cnts[r];
    float carry = 0.f;
    for (int64_t base = 0; base < c; base += 32) {
      const int64_t k = base + lane, idx = s + k;
      const bool ok = k < c;
      const float sdt = ok ? __fmul_rn(sigmas[idx], __fsub_rn(t_ends[idx], t_starts[idx])) : 0.f;
      const float inc = warp_inclusive_sum(sdt, lane);
      float ex = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) ex = 0.f;
      const float esum = __fadd_rn(carry, ex);
      if (ok) {
        const float a = __fsub_rn(1.0f, expf(-sdt));
        float T = expf(-esum);
        if (prefix_trans) T = __fmul_rn(T, prefix_trans[idx]);
        if (alphas) alphas[idx] = a;
        if (trans) trans[idx] = T;
        if (weights) weights[idx] = __fmul_rn(T, a);
      }
      carry = __fadd_rn(carry, __shfl_sync(0xffffffffu, inc, 31));
    }
  }
}

// Backward of the above w.r.t. sigmas (and prefix_trans).  With q_k = (gw_k a_k + gT_k) T_k:
//   d/d sdt_i = (gw_i T_i + ga_i) exp(-sdt_i) - sum_{k>i} q_k ;  d/d sigma_i = that * dt_i
//   d/d prefix_i = (gw_i a_i + gT_i) exp(-E_i)
__global__ void __launch_bounds__(256) weights_from_density_bwd_kernel(
    int n_rays, const int64_t* __restrict__ starts, const int64_t* __restrict__ cnts,
    const float* __restrict__ t_starts, const float* __restrict__ t_ends, const float* __restrict__ sigmas,
    const float* __restrict__ prefix_trans, const float* __restrict__ g_weights, const float* __restrict__ g_trans,
    const float* __restrict__ g_alphas, float* __restrict__ g_sigmas, float* __restrict__ g_prefix) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps_per_block) {
    const int64_t s = starts[r], c = cnts[r];
    // pass 1 (forward order): total of sdt, to rebuild E_i when walking backwards
    float tot = 0.f;
    for (int64_t k = lane; k < c; k += 32)
      tot += sigmas[s + k] * (t_ends[s + k] - t_starts[s + k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    // pass 2 (reverse order): suffix sums of sdt (-> E_i = tot - suffix_incl_i) and of q
    float carry_s = 0.f, carry_q = 0.f;
    for (int64_t base = 0; base < c; base += 32) {
      const int64_t k = base + lane, idx = s + c - 1 - k;
      const bool ok = k < c;
      const float dt = ok ? (t_ends[idx] - t_starts[idx]) : 0.f;
      const float sdt = ok ? sigmas[idx] * dt : 0.f;
      const float inc_s = warp_inclusive_sum(sdt, lane);
      const float E = tot - (carry_s + inc_s);  // exclusive prefix sum in forward order
      const float a = 1.0f - expf(-sdt);
      const float eE = expf(-E);
      const float T = prefix_trans && ok ? eE * prefix_trans[idx] : eE;
      const float gw = (ok && g_weights) ? g_weights[idx] : 0.f;
      const float gT = (ok && g_trans) ? g_trans[idx] : 0.f;
      const float ga = (ok && g_alphas) ? g_alphas[idx] : 0.f;
      const float q = ok ? (gw * a + gT) * T : 0.f;
      const float inc_q = warp_inclusive_sum(q, lane);
      const float suffix_q = carry_q + (inc_q - q);  // sum over k' > i in forward order
      if (ok) {
        if (g_sigmas) g_sigmas[idx] = ((gw * T + ga) * expf(-sdt) - suffix_q) * dt;
        if (g_prefix) g_prefix[idx] = (gw * a + gT) * eE;
      }
      carry_s += __shfl_sync(0xffffffffu, inc_s, 31);
      carry_q += __shfl_sync(0xffffffffu, inc_q, 31);
    }
  }
}

// outputs[ray_indices[i], :] += weights[i] * values[i, :]   (values == nullptr -> D = 1, value 1)
// Channels on lanes: a warp walks a contiguous run of samples, keeps the running sum in a
// register while the ray id does not change and flushes with one atomicAdd per (ray, channel)
// per run -- for packed (sorted) ray_indices that is one atomic per ray per warp chunk.
__global__ void __launch_bounds__(256) accumulate_wide_kernel(long long n, int D, const float* __restrict__ weights,
                                                              const float* __restrict__ values,
                                                              const int64_t* __restrict__ ray_indices,
                                                              float* __restrict__ out, int chunk) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long b = warp * chunk; b < n; b += n_warps * chunk) {
    const long long e = (b + chunk < n) ? b + chunk : n;
    for (int d0 = 0; d0 < D; d0 += 32) {
      const int d = d0 + lane;
      int64_t cur = ray_indices[b];
      float acc = 0.f;
      for (long long i = b; i < e; ++i) {
        const int64_t r = ray_indices[i];
        if (r != cur) {
          if (d < D) atomicAdd(out + cur * D + d, acc);
          acc = 0.f;
          cur = r;
        }
        if (d < D) acc = __fadd_rn(acc, __fmul_rn(weights[i], values[i * D + d]));
      }
      if (d < D) atomicAdd(out + cur * D + d, acc);
    }
  }
}

// Samples on lanes (D <= 4): segmented shuffle reduction keyed by ray id, one atomic per
// (segment, channel).
__global__ void __launch_bounds__(256) accumulate_narrow_kernel(long long n, int D, const float* __restrict__ weights,
                                                                const float* __restrict__ values,
                                                                const int64_t* __restrict__ ray_indices,
                                                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long n_round = (n + 31) / 32 * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_round;
       i += (long long)blockDim.x * gridDim.x) {
    const bool ok = i < n;
    const int64_t r = ok ? ray_indices[i] : -1;
    const float w = ok ? weights[i] : 0.f;
    const int64_t r_prev = __shfl_up_sync(0xffffffffu, r, 1);
    const bool head = (lane == 0) || (r != r_prev);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    // segment end (exclusive) for this lane's segment
    const unsigned later = heads & ~((2u << lane) - 1u);  // heads strictly after this lane
    const int seg_end = later ? (__ffs(later) - 1) : 32;
    for (int d = 0; d < D; ++d) {
      float v = ok ? (values ? __fmul_rn(w, values[i * D + d]) : w) : 0.f;
      // segmented suffix-sum by doubling: lane accumulates lanes [lane, seg_end)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float nv = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < seg_end) v += nv;
      }
      if (ok && head) atomicAdd(out + r * D + d, v);
    }
  }
}

// grad_weights[i] = sum_d go[ray[i], d] * values[i, d];  grad_values[i, d] = weights[i] * go[ray[i], d]
__global__ void __launch_bounds__(256) accumulate_bwd_kernel(long long n, int D, const float* __restrict__ weights,
                                                             const float* __restrict__ values,
                                                             const int64_t* __restrict__ ray_indices,
                                                             const float* __restrict__ g_out,
                                                             float* __restrict__ g_weights,
                                                             float* __restrict__ g_values) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)blockDim.x * gridDim.x) {
    const int64_t r = ray_indices[i];
    const float w = weights[i];
    float gw = 0.f;
    for (int d = 0; d < D; ++d) {
      const float g = g_out[r * D + d];
      if (values) gw += g * values[i * D + d];
      else gw += g;
      if (g_values) g_values[i * D + d] = w * g;
    }
    if (g_weights) g_weights[i] = gw;
  }
}

__global__ void __launch_bounds__(256) histogram_kernel(long long n, const int64_t* __restrict__ ray_indices,
                                                        int n_rays, unsigned long long* __restrict__ cnts) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)blockDim.x * gridDim.x) {
    const int64_t r = ray_indices[i];
    if (r >= 0 && r < n_rays) atomicAdd(cnts + r, 1ull);
  }
}

__global__ void __launch_bounds__(256) interleave_kernel(int n_rays, const int64_t* __restrict__ a,
                                                         const int64_t* __restrict__ b, int64_t* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += blockDim.x * gridDim.x) {
    out[2 * i] = a[i];
    out[2 * i + 1] = b[i];
  }
}

}  // namespace apnerf

using namespace apnerf;

extern "C" int apnerf_exclusive_scan_i64(long long n, const int64_t* in, int64_t* out, int64_t* total,
                                         int64_t* scratch, void* stream);

// Packed inclusive / exclusive sum, forward or reverse ("backward") order.
APNERF_API int apnerf_packed_sum(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                                 long long n_edges, const float* inputs, float* outputs, int inclusive,
                                 int normalize, int backward, void* stream) {
  if (n_edges == 0 || n_rays == 0) return 0;  // scan.cu:32-34
  APNERF_REQUIRE(!(backward && normalize), "packed_sum: backward does not support normalize");
  const int grid = grid_for((long long)n_rays * 32, 256, 8);
  cudaStream_t st = (cudaStream_t)stream;
  const bool nz = normalize != 0;
  if (inclusive) {
    if (backward) packed_sum_kernel<true, true><<<grid, 256, 0, st>>>(n_rays, chunk_starts, chunk_cnts, inputs, outputs, nz);
    else packed_sum_kernel<true, false><<<grid, 256, 0, st>>>(n_rays, chunk_starts, chunk_cnts, inputs, outputs, nz);
  } else {
    if (backward) packed_sum_kernel<false, true><<<grid, 256, 0, st>>>(n_rays, chunk_starts, chunk_cnts, inputs, outputs, nz);
    else packed_sum_kernel<false, false><<<grid, 256, 0, st>>>(n_rays, chunk_starts, chunk_cnts, inputs, outputs, nz);
  }
  APNERF_CHECK_LAUNCH("packed_sum_kernel");
  return 0;
}

APNERF_API int apnerf_weights_from_density(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                                           long long n_samples, const float* t_starts, const float* t_ends,
                                           const float* sigmas, const float* prefix_trans, float* weights,
                                           float* trans, float* alphas, void* stream) {
  if (n_samples == 0 || n_rays == 0) return 0;
  weights_from_density_kernel<<<grid_for((long long)n_rays * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, chunk_starts, chunk_cnts, t_starts, t_ends, sigmas, prefix_trans, weights, trans, alphas);
  APNERF_CHECK_LAUNCH("weights_from_density_kernel");
  return 0;
}

APNERF_API int apnerf_weights_from_density_bwd(int n_rays, const int64_t* chunk_starts, const int64_t* chunk_cnts,
                                               long long n_samples, const float* t_starts, const float* t_ends,
                                               const float* sigmas, const float* prefix_trans,
                                               const float* g_weights, const float* g_trans, const float* g_alphas,
                                               float* g_sigmas, float* g_prefix, void* stream) {
  if (n_samples == 0 || n_rays == 0) return 0;
  weights_from_density_bwd_kernel<<<grid_for((long long)n_rays * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_rays, chunk_starts, chunk_cnts, t_starts, t_ends, sigmas, prefix_trans, g_weights, g_trans, g_alphas,
      g_sigmas, g_prefix);
  APNERF_CHECK_LAUNCH("weights_from_density_bwd_kernel");
  return 0;
}

// outputs [n_rays, D] is accumulated IN PLACE (callers zero it for the non-underscore form).
APNERF_API int apnerf_accumulate_along_rays(long long n_samples, int D, const float* weights, const float* values,
                                            const int64_t* ray_indices, float* outputs, void* stream) {
  if (n_samples == 0) return 0;
  APNERF_REQUIRE(D >= 1, "accumulate_along_rays: D must be >= 1");
  APNERF_REQUIRE(values != nullptr || D == 1, "accumulate_along_rays: values == NULL requires D == 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (D <= 4) {
    accumulate_narrow_kernel<<<grid_for(n_samples, 256, 8), 256, 0, st>>>(n_samples, D, weights, values,
                                                                          ray_indices, outputs);
  } else {
    const int chunk = 16;
    accumulate_wide_kernel<<<grid_for((n_samples + chunk - 1) / chunk * 32, 256, 8), 256, 0, st>>>(
        n_samples, D, weights, values, ray_indices, outputs, chunk);
  }
  APNERF_CHECK_LAUNCH("accumulate_kernel");
  return 0;
}

APNERF_API int apnerf_accumulate_along_rays_bwd(long long n_samples, int D, const float* weights,
                                                const float* values, const int64_t* ray_indices,
                                                const float* g_outputs, float* g_weights, float* g_values,
                                                void* stream) {
  if (n_samples == 0) return 0;
  accumulate_bwd_kernel<<<grid_for(n_samples, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      n_samples, D, weights, values, ray_indices, g_outputs, g_weights, g_values);
  APNERF_CHECK_LAUNCH("accumulate_bwd_kernel");
  return 0;
}

// packed_info [n_rays, 2] = (start, count) per ray.  scratch: 2 * n_rays + scan scratch int64.
APNERF_API int apnerf_pack_info(long long n_samples, const int64_t* ray_indices, int n_rays, int64_t* packed_info,
                                int64_t* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) return 0;
  int64_t* cnts = scratch;
  int64_t* starts = scratch + n_rays;
  int64_t* scan_scratch = scratch + 2 * (long long)n_rays;
  APNERF_CUDA(cudaMemsetAsync(cnts, 0, sizeof(int64_t) * n_rays, st));
  if (n_samples > 0)
    histogram_kernel<<<grid_for(n_samples, 256, 8), 256, 0, st>>>(n_samples, ray_indices, n_rays,
                                                                  (unsigned long long*)cnts);
  int rc = apnerf_exclusive_scan_i64(n_rays, cnts, starts, nullptr, scan_scratch, stream);
  if (rc) return rc;
  interleave_kernel<<<grid_for(n_rays, 256, 8), 256, 0, st>>>(n_rays, starts, cnts, packed_info);
  APNERF_CHECK_LAUNCH("pack_info");
  return 0;
}
