// TEST INFRASTRUCTURE ONLY (oracle).  Python binding for the UNMODIFIED reference
// nerfacc CUDA sources, compiled in place from /root/reference by oracle/build_ref.py.
// Only the hot-path entry points are bound (grid.cu + scan.cu); the reference's own
// binding file (perception/nerfacc/nerfacc/cuda/csrc/nerfacc.cpp:100-129) also pulls in
// pdf.cu / camera.cu which are out of scope.
// No reference source is copied into this repository: the .cu files are compiled where
// they lie and only the resulting .so lands in oracle/_ref/ (git-ignored).
#include "include/data_spec.hpp"
#include <torch/extension.h>

torch::Tensor inclusive_sum(torch::Tensor chunk_starts, torch::Tensor chunk_cnts,
                            torch::Tensor inputs, bool normalize, bool backward);
torch::Tensor exclusive_sum(torch::Tensor chunk_starts, torch::Tensor chunk_cnts,
                            torch::Tensor inputs, bool normalize, bool backward);
std::vector<torch::Tensor> ray_aabb_intersect(const torch::Tensor rays_o, const torch::Tensor rays_d,
                                              const torch::Tensor aabbs, const float near_plane,
                                              const float far_plane, const float miss_value);
std::tuple<RaySegmentsSpec, RaySegmentsSpec, torch::Tensor> traverse_grids(
    const torch::Tensor rays_o, const torch::Tensor rays_d, const torch::Tensor rays_mask,
    const torch::Tensor binaries, const torch::Tensor aabbs, const torch::Tensor t_sorted,
    const torch::Tensor t_indices, const torch::Tensor hits, const torch::Tensor near_planes,
    const torch::Tensor far_planes, const float step_size, const float cone_angle,
    const bool compute_intervals, const bool compute_samples, const bool compute_terminate_planes,
    const int32_t traverse_steps_limit, const bool over_allocate);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("inclusive_sum", &inclusive_sum);
  m.def("exclusive_sum", &exclusive_sum);
  m.def("ray_aabb_intersect", &ray_aabb_intersect);
  m.def("traverse_grids", &traverse_grids);
  py::class_<RaySegmentsSpec>(m, "RaySegmentsSpec")
      .def(py::init<>())
      .def_readwrite("vals", &RaySegmentsSpec::vals)
      .def_readwrite("is_left", &RaySegmentsSpec::is_left)
      .def_readwrite("is_right", &RaySegmentsSpec::is_right)
      .def_readwrite("is_valid", &RaySegmentsSpec::is_valid)
      .def_readwrite("chunk_starts", &RaySegmentsSpec::chunk_starts)
      .def_readwrite("chunk_cnts", &RaySegmentsSpec::chunk_cnts)
      .def_readwrite("ray_indices", &RaySegmentsSpec::ray_indices);
}
