// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// pybind11 table for the UNMODIFIED reference CUDA kernels of the hot path
// (sources compiled in place from /root/reference/gsplat/cuda/csrc by
// oracle/Makefile; nothing is copied into this repo).  The resulting module
// oracle/_ref/gsplat_ref_cuda.so is the GPU-side oracle / "reference CUDA on the
// same B200" baseline.  Signatures come from the reference header
// gsplat/cuda/include/bindings.h:94-309 which this file includes directly.
#include "bindings.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    py::enum_<gsplat::CameraModelType>(m, "CameraModelType")
        .value("PINHOLE", gsplat::CameraModelType::PINHOLE)
        .value("ORTHO", gsplat::CameraModelType::ORTHO)
        .value("FISHEYE", gsplat::CameraModelType::FISHEYE);
    m.def("compute_sh_fwd", &gsplat::compute_sh_fwd_tensor);
    m.def("compute_sh_bwd", &gsplat::compute_sh_bwd_tensor);
    m.def("fully_fused_projection_fwd", &gsplat::fully_fused_projection_fwd_tensor);
    m.def("fully_fused_projection_bwd", &gsplat::fully_fused_projection_bwd_tensor);
    m.def("fully_fused_projection_packed_fwd", &gsplat::fully_fused_projection_packed_fwd_tensor);
    m.def("fully_fused_projection_packed_bwd", &gsplat::fully_fused_projection_packed_bwd_tensor);
    m.def("isect_tiles", &gsplat::isect_tiles_tensor);
    m.def("isect_offset_encode", &gsplat::isect_offset_encode_tensor);
    m.def("rasterize_to_pixels_fwd", &gsplat::rasterize_to_pixels_fwd_tensor);
    m.def("rasterize_to_pixels_bwd", &gsplat::rasterize_to_pixels_bwd_tensor);
    // SURVEY.md section 8(f) rows
    m.def("rasterize_to_indices_in_range", &gsplat::rasterize_to_indices_in_range_tensor);
    m.def("quat_scale_to_covar_preci_fwd", &gsplat::quat_scale_to_covar_preci_fwd_tensor);
    m.def("quat_scale_to_covar_preci_bwd", &gsplat::quat_scale_to_covar_preci_bwd_tensor);
    m.def("proj_fwd", &gsplat::proj_fwd_tensor);
    m.def("proj_bwd", &gsplat::proj_bwd_tensor);
    m.def("world_to_cam_fwd", &gsplat::world_to_cam_fwd_tensor);
    m.def("world_to_cam_bwd", &gsplat::world_to_cam_bwd_tensor);
    m.def("selective_adam_update", &gsplat::selective_adam_update);
    m.def("compute_relocation", &gsplat::compute_relocation_tensor);
}
