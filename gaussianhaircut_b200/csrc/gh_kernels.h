// Internal launcher declarations (one per stage); the public C-ABI lives in include/gh_rasterizer.h.
#pragma once
#include "gh_common.cuh"

void gh_launch_preprocess(int P, const float* means3D, const float* scales, float scale_modifier,
                          const float* rotations, const float* opacities, const float* cov3D_precomp,
                          const float* conic_precomp, const float* viewmatrix, const float* projmatrix,
                          int W, int H, float tan_fovx, float tan_fovy, int* radii,
                          GhGeomWS geom, GhImgWS img, int prefiltered, cudaStream_t stream);

void gh_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, bool* present,
                            cudaStream_t stream);

// exclusive scan of the tile histogram -> ranges, cursors, R, longest list
void gh_launch_tile_scan(int T, GhImgWS img, cudaStream_t stream);

// scatter (depth|idx) records into their tile buckets
void gh_launch_emit(int P, const int* radii, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                    int gx, int gy, cudaStream_t stream);

// sort every tile bucket by (depth bits, gaussian idx)
// returns the number of kernels launched
int gh_launch_tile_sort(int T, unsigned int max_tile_len, long long R, GhImgWS img, GhBinWS bin, cudaStream_t stream);

void gh_launch_blend_forward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                             const float* features, const float* bg, float* out_color,
                             cudaStream_t stream);

// accumulates into geom.acc16 (must be zero on entry)
void gh_launch_blend_backward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                              const float* features, const float* bg, const float* dL_dpix,
                              cudaStream_t stream);

// geom.acc16 -> dL_dmean2D (P,3), dL_dconic (P,4), dL_dopacity (P), dL_dcolor (P,10); writes every row
// (also zero-fills the geometry gradients that are non-null: used when the conic was supplied)
void gh_launch_unpack_grads(int P, GhGeomWS geom, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                            float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale, float* dL_drot,
                            cudaStream_t stream);

void gh_launch_preprocess_backward(int P, const float* means3D, const int* radii,
                                   const float* scales, float scale_modifier, const float* rotations,
                                   const float* cov3D_precomp, const float* conic_precomp,
                                   const float* viewmatrix, const float* projmatrix,
                                   int W, int H, float tan_fovx, float tan_fovy,
                                   const float* acc16, float* dL_dmean2D, float* dL_dconic,
                                   float* dL_dopacity, float* dL_dcolor,
                                   float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale, float* dL_drot,
                                   cudaStream_t stream);

// per-thread error message behind gh_last_error(): every extern "C" entry point clears it on entry and
// sets it before returning a GH_E_* code (defined in gh_api.cu)
void gh_clear_error();
int gh_set_error(int code, const char* msg);

// kernels launched outside gh_api.cu (optimizer, image losses) report themselves to gh_kernel_launch_count()
void gh_count_launches(int n);
