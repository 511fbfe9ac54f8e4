// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Python entry table for the UNMODIFIED reference CUDA kernels (sources compiled in place from
// /root/reference/gsplat/cuda/csrc by oracle/Makefile; nothing is copied into this repo).  The resulting module
// oracle/_ref/gsplat_ref_cuda.so is the GPU-side oracle / "reference CUDA on the same B200" baseline.  The C++
// signatures come from the reference header gsplat/cuda/include/bindings.h, which this file includes directly; the
// Python names are the ones gsplat/cuda/_wrapper.py looks up on its `_C` module.
#include "bindings.h"

// ops whose C++ entry point is `gsplat::<python name>_tensor`
#define GSB_REF_TENSOR_OPS(OP)                                                                               \
    OP(compute_sh_fwd) OP(compute_sh_bwd)                           /* colour */                             \
    OP(fully_fused_projection_fwd) OP(fully_fused_projection_bwd)   /* stage 1 / 4b, dense */                \
    OP(fully_fused_projection_packed_fwd) OP(fully_fused_projection_packed_bwd)                              \
    OP(isect_tiles) OP(isect_offset_encode)                         /* stage 2 */                            \
    OP(rasterize_to_pixels_fwd) OP(rasterize_to_pixels_bwd)         /* stage 3 / 4a */                       \
    OP(rasterize_to_indices_in_range)                               /* SURVEY.md section 8(f) rows below */  \
    OP(quat_scale_to_covar_preci_fwd) OP(quat_scale_to_covar_preci_bwd)                                      \
    OP(proj_fwd) OP(proj_bwd) OP(world_to_cam_fwd) OP(world_to_cam_bwd)                                      \
    OP(compute_relocation)                                                                                   \
    OP(fully_fused_projection_fwd_2dgs) OP(fully_fused_projection_bwd_2dgs) /* 2D Gaussian (surfel) family */ \
    OP(fully_fused_projection_packed_fwd_2dgs) OP(fully_fused_projection_packed_bwd_2dgs)                    \
    OP(rasterize_to_pixels_fwd_2dgs) OP(rasterize_to_pixels_bwd_2dgs) OP(rasterize_to_indices_in_range_2dgs)

PYBIND11_MODULE(TORCH_EXTENSION_NAME, mod) {
    namespace ref = gsplat;
    py::enum_<ref::CameraModelType>(mod, "CameraModelType")
        .value("PINHOLE", ref::CameraModelType::PINHOLE)
        .value("ORTHO", ref::CameraModelType::ORTHO)
        .value("FISHEYE", ref::CameraModelType::FISHEYE);
#define GSB_BIND(name) mod.def(#name, &ref::name##_tensor);
    GSB_REF_TENSOR_OPS(GSB_BIND)
#undef GSB_BIND
    mod.def("selective_adam_update", &ref::selective_adam_update);  // the one op without the _tensor suffix
}
