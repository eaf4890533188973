// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// C-ABI shim around the *reference's own* CUDA rasterizer and simple-knn.  The
// reference sources are compiled in place from /root/reference by oracle/Makefile
// (outputs only under oracle/_ref/); nothing of them is copied into this repo.
// This file is ours: it only adapts
//   CudaRasterizer::Rasterizer::forward / backward
//       (submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:24-95)
//   SimpleKNN::knn (submodules/simple-knn/simple_knn.h:15-19)
// to plain-pointer entry points so tests can drive them through ctypes with the
// same torch device buffers they hand to libdgs_b200.so.
#include <cuda_runtime.h>
#include <functional>
#include <cstddef>
#include <cstdint>
#include "cuda_rasterizer/rasterizer.h"
#include "simple_knn.h"

typedef char* (*ref_alloc_fn)(size_t bytes);

extern "C" {

int ref_forward(ref_alloc_fn geom, ref_alloc_fn binning, ref_alloc_fn img,
                int P, int D, int M,
                const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos,
                float tan_fovx, float tan_fovy, float z_near, float z_far, int prefiltered,
                float* out_color, float* out_depth, int* radii, int use_sigmoid)
{
    std::function<char*(size_t)> g = [geom](size_t n) { return geom(n); };
    std::function<char*(size_t)> b = [binning](size_t n) { return binning(n); };
    std::function<char*(size_t)> i = [img](size_t n) { return img(n); };
    int r = CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, background, W, H,
        means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
        cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far,
        prefiltered != 0, out_color, out_depth, radii, use_sigmoid != 0, false);
    return r;
}

void ref_backward(int P, int D, int M, int R,
                  const float* background, int W, int H,
                  const float* means3D, const float* shs, const float* colors_precomp,
                  const float* scales, float scale_modifier, const float* rotations,
                  const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const float* campos, float tan_fovx, float tan_fovy, float z_near, float z_far,
                  const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, const float* dL_dpixdepth,
                  float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                  float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                  float* dL_drot, float* dL_ddepths, float* dL_dviewmatrix, float* dL_dprojmatrix,
                  int use_sigmoid)
{
    CudaRasterizer::Rasterizer::backward(P, D, M, R, background, W, H, means3D, shs,
        colors_precomp, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
        projmatrix, campos, tan_fovx, tan_fovy, z_near, z_far, radii, geom_buffer,
        binning_buffer, image_buffer, dL_dpix, dL_dpixdepth, dL_dmean2D, dL_dconic,
        dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
        dL_ddepths, dL_dviewmatrix, dL_dprojmatrix, use_sigmoid != 0, false);
}

void ref_knn(int P, float* points, float* mean_dists)
{
    SimpleKNN::knn(P, (float3*)points, mean_dists);
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
