// "Next" row 1: the caller-side projection preamble as ONE forward and ONE backward kernel.
//
// Every training iteration the reference's `render()` / `render_hair()` restate stage 1 of the rasterizer in
// PyTorch before calling it (~100 elementwise / bmm kernels with (P,3,3) temporaries and six boolean-mask
// gathers): src/gaussian_renderer/__init__.py:29-83,122-186 calling
//   GaussianModel.get_conic / get_covariance_2d / get_covariance   src/scene/gaussian_model.py:230-315
//   get_mean_2d :317-337, get_depths :339-342, get_direction_2d :344-393, filter_points :143-228
//   eval_sh                                                         src/utils/sh_utils.py:57-112
//   build_rotation (normalises the quaternion)                      src/utils/general_utils.py:79-109
// (the strand models restate the same functions: src/scene/gaussian_model_latent_strands.py:150-440).
// Here one thread owns one Gaussian from the RAW model parameters to everything the rasterizer consumes
// (conic, NDC mean, opacity, the 10-channel feature row) and back -- including the gradients w.r.t. the
// camera (world-view matrix, full projection matrix, camera centre, tan(fov/2)), which the reference
// obtains through autograd because cameras are trainable (src/scene/cameras.py:124-150).
//
// Row-vector convention of the Python code: t = x . V[:3,:3] + V[3,:3],  h = x . Pm[:3,:] + Pm[3,:].
//   s = exp(log s) * mod (or the activated scale),  R = build_rotation(q / |q|),  Sigma = sum_i s_i^2 R[i]^T R[i]
//   u0 = T[:,0], u1 = T[:,1] with T = V[:3,:3] @ J  (J from the clamped view-space position)
//   a = sum_i s_i^2 (R[i].u0)^2 + 0.3,  b = sum_i s_i^2 (R[i].u0)(R[i].u1),  c = sum_i s_i^2 (R[i].u1)^2 + 0.3
//   conic = (c, -b, a) / (ac - b^2 + eps)        eps = 1e-12 (GaussianModel :312) or 1e-7 (strand models :355)
//   dir2D = d3 . T,  d3 = s_max R[argmax s] (GaussianModel :385-391) or normalize(dir) (strand models :437-438)
//   feature row = [max(SH(v)+0.5, 0), label, 1, dir2D (3), orientation confidence, view depth]
// Gaussians failing the caller's prefilter (filter_points: near plane, det != 0, empty tile rectangle) are
// not compacted away: they get conic = 0, which the rasterizer drops (zero determinant, forward.cu:243-245),
// so indices stay the model's own (radii / visibility need no scatter) and every gradient row is written.
//
// f_rest rows are 180 bytes (45 floats): a CTA stages its 128 rows through shared memory with coalesced
// 128-bit loads / stores; the odd row stride (45 words) makes the per-thread accesses conflict free.
#include "gh_common.cuh"
#include "gh_kernels.h"
#include "../../include/gh_rasterizer.h"
#include "gh_project_math.h"

#include <cstdio>

namespace {

#define GH_PJ_THREADS 128

// stage the CTA's f_rest rows (or their gradients) through shared memory with coalesced 128-bit accesses
__device__ __forceinline__ void gh_rest_load(float* s_rest, const float* __restrict__ f_rest, int P, int row0) {
    const size_t base = (size_t)row0 * GH_PJ_REST;
    const size_t total = (size_t)P * GH_PJ_REST;
    const int n = (int)min((size_t)GH_PJ_THREADS * GH_PJ_REST, total - base);
    if ((reinterpret_cast<size_t>(f_rest + base) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(f_rest + base);
        for (int v = threadIdx.x; v < n / 4; v += GH_PJ_THREADS) reinterpret_cast<float4*>(s_rest)[v] = src[v];
        for (int e = (n & ~3) + threadIdx.x; e < n; e += GH_PJ_THREADS) s_rest[e] = f_rest[base + e];
    } else {
        for (int e = threadIdx.x; e < n; e += GH_PJ_THREADS) s_rest[e] = f_rest[base + e];
    }
}
__device__ __forceinline__ void gh_rest_store(const float* s_rest, float* __restrict__ out, int P, int row0) {
    const size_t base = (size_t)row0 * GH_PJ_REST;
    const size_t total = (size_t)P * GH_PJ_REST;
    const int n = (int)min((size_t)GH_PJ_THREADS * GH_PJ_REST, total - base);
    if ((reinterpret_cast<size_t>(out + base) & 15) == 0) {
        float4* dst = reinterpret_cast<float4*>(out + base);
        for (int v = threadIdx.x; v < n / 4; v += GH_PJ_THREADS) dst[v] = reinterpret_cast<const float4*>(s_rest)[v];
        for (int e = (n & ~3) + threadIdx.x; e < n; e += GH_PJ_THREADS) out[base + e] = s_rest[e];
    } else {
        for (int e = threadIdx.x; e < n; e += GH_PJ_THREADS) out[base + e] = s_rest[e];
    }
}

// ------------------------------------------------------------------------------------------------ forward
// BIN: also run stage 1 of the rasterizer for the Gaussian just projected (what gh_preprocess_kernel does when it is
// handed the conic): splat radius, tile rectangle, the 32-byte state record of the blend kernels, depth, and the
// per-tile instance histogram -- the conic, mean and opacity are still in registers, so the rasterizer's own
// preprocess launch (and its re-read of 28 bytes per Gaussian) disappears from a fused render.  Same device functions
// as gh_preprocess_kernel (gh_common.cuh), hence bit-identical radii / records / keys.
template <bool BIN>
__global__ void __launch_bounds__(GH_PJ_THREADS)
gh_project_forward_kernel(GhProjArgs A, float* __restrict__ means2D, float* __restrict__ colors,
                          float* __restrict__ opac_out, float* __restrict__ conic_out, float* __restrict__ cov3D_out,
                          unsigned char* __restrict__ mask_out,
                          int* __restrict__ radii, GhGeo* __restrict__ geo, float* __restrict__ depth,
                          uint32_t* __restrict__ tile_count, int gx, int gy)
{
    __shared__ __align__(16) float s_rest[GH_PJ_THREADS * GH_PJ_REST];
    const int row0 = blockIdx.x * GH_PJ_THREADS;
    const int i = row0 + threadIdx.x;
    if (A.sh_degree > 0) gh_rest_load(s_rest, A.f_rest, A.P, row0);
    __syncthreads();
    int rect_minx = 0, rect_miny = 0, rect_maxx = 0, rect_maxy = 0;
    if (i < A.P) {
        GhProjOut o;
        gh_project_forward_one(A, i, s_rest + threadIdx.x * GH_PJ_REST, cov3D_out != nullptr, o);
        means2D[3 * (size_t)i] = o.m2[0]; means2D[3 * (size_t)i + 1] = o.m2[1]; means2D[3 * (size_t)i + 2] = o.m2[2];
        conic_out[3 * (size_t)i] = o.conic[0]; conic_out[3 * (size_t)i + 1] = o.conic[1]; conic_out[3 * (size_t)i + 2] = o.conic[2];
        mask_out[i] = o.visible ? 1 : 0;
        opac_out[i] = o.opacity;
        if (cov3D_out) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov3D_out[6 * (size_t)i + k] = o.cov3D[k];
        }
        float2* crow = reinterpret_cast<float2*>(colors + (size_t)i * GH_NUM_CHANNELS);
#pragma unroll
        for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) crow[k] = make_float2(o.color[2 * k], o.color[2 * k + 1]);
        if (BIN) {
            int out_radius = 0;
            GhGeo g = gh_geo_not_rendered();
            float zview = 0.f, projx, projy, covx, covz, det;
            const float px = A.xyz[3 * (size_t)i], py = A.xyz[3 * (size_t)i + 1], pz = A.xyz[3 * (size_t)i + 2];
            // (culled Gaussians carry conic = 0 -> singular -> radius 0, exactly as in the two-kernel path)
            if (gh_pre_project(px, py, pz, A.V, A.Pm, zview, projx, projy) &&
                gh_pre_from_conic(o.conic[0], o.conic[1], o.conic[2], covx, covz, det))
                out_radius = gh_pre_finish(covx, covz, det, o.conic[0], o.conic[1], o.conic[2], projx, projy, o.opacity,
                                           A.W, A.H, gx, gy, rect_minx, rect_miny, rect_maxx, rect_maxy, g);
            radii[i] = out_radius;
            depth[i] = zview;
            float4* gp = reinterpret_cast<float4*>(geo + i);
            gp[0] = make_float4(g.x, g.y, g.ca, g.cb);
            gp[1] = make_float4(g.cc, g.op, g.thr, g.pd);
        }
    }
    if (BIN) gh_warp_tile_histogram(rect_minx, rect_miny, rect_maxx, rect_maxy, gx, tile_count);
}

// ------------------------------------------------------------------------------------------------ backward
// Incoming gradients: either the four API-shaped tensors of gh_backward (dL_dmean2D (P,3) NDC units,
// dL_dconic (P,4) = (d/da, HALF d/db, unused, d/dc), dL_dcolor (P,10), dL_dopacity (P)), or -- when acc16 is
// given -- the blend backward's 64-byte accumulation records themselves (colors 0..9, mean2D x y, conic x y w,
// opacity), which skips the unpack kernel and its round trip through HBM.
__global__ void __launch_bounds__(GH_PJ_THREADS)
gh_project_backward_kernel(GhProjArgs A, const unsigned char* __restrict__ mask,
                           const float* __restrict__ acc16,
                           const float* __restrict__ g_mean2D, const float* __restrict__ g_conic4,
                           const float* __restrict__ g_color, const float* __restrict__ g_opacity,
                           float* __restrict__ d_xyz, float* __restrict__ d_scaling, float* __restrict__ d_rotation,
                           float* __restrict__ d_dirs, float* __restrict__ d_fdc, float* __restrict__ d_frest,
                           float* __restrict__ d_opacity, float* __restrict__ d_label, float* __restrict__ d_conf,
                           float* __restrict__ d_mean2D_out,
                           float* __restrict__ cam_partial, unsigned int* __restrict__ cam_ticket,
                           float* __restrict__ d_cam,     // 16 V + 16 Pm + 3 campos + 2 tan, or NULL
                           unsigned int* __restrict__ nan_flag)   // set to 1 if any parameter gradient is NaN, or NULL
{
    __shared__ __align__(16) float s_rest[GH_PJ_THREADS * GH_PJ_REST];
    __shared__ float s_cam[GH_PJ_THREADS / 32][GH_PJ_NCAM];
    const int row0 = blockIdx.x * GH_PJ_THREADS;
    const int i = row0 + threadIdx.x;
    const bool want_cam = (d_cam != nullptr);
    if (A.sh_degree > 0) gh_rest_load(s_rest, A.f_rest, A.P, row0);
    __syncthreads();

    float cam[GH_PJ_NCAM];
#pragma unroll
    for (int k = 0; k < GH_PJ_NCAM; k++) cam[k] = 0.f;
    GhProjGradOut go;
#pragma unroll
    for (int k = 0; k < 3; k++) { go.xyz[k] = 0.f; go.scaling[k] = 0.f; go.dirs[k] = 0.f; go.f_dc[k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; k++) go.rotation[k] = 0.f;
    go.opacity = 0.f; go.label = 0.f; go.conf = 0.f;
#pragma unroll
    for (int k = 0; k < GH_PJ_REST; k++) go.rest[k] = 0.f;

    if (i < A.P) {
        const bool live = (mask[i] != 0);
        GhProjGradIn gi;
        gi.m2x = 0.f; gi.m2y = 0.f; gi.con[0] = 0.f; gi.con[1] = 0.f; gi.con[2] = 0.f; gi.opacity = 0.f;
#pragma unroll
        for (int k = 0; k < GH_NUM_CHANNELS; k++) gi.color[k] = 0.f;
        if (live) {
            if (acc16) {
                const float4* rec = reinterpret_cast<const float4*>(acc16 + (size_t)i * 16);
                const float4 a0 = rec[0], a1 = rec[1], a2 = rec[2], a3 = rec[3];
                gi.color[0] = a0.x; gi.color[1] = a0.y; gi.color[2] = a0.z; gi.color[3] = a0.w; gi.color[4] = a1.x; gi.color[5] = a1.y;
                gi.color[6] = a1.z; gi.color[7] = a1.w; gi.color[8] = a2.x; gi.color[9] = a2.y;
                gi.m2x = a2.z; gi.m2y = a2.w; gi.con[0] = a3.x; gi.con[1] = 2.f * a3.y; gi.con[2] = a3.z; gi.opacity = a3.w;
            } else {
                gi.m2x = g_mean2D[3 * (size_t)i]; gi.m2y = g_mean2D[3 * (size_t)i + 1];
                const float4 c4 = reinterpret_cast<const float4*>(g_conic4)[i];
                gi.con[0] = c4.x; gi.con[1] = 2.f * c4.y; gi.con[2] = c4.w;     // python restack [g00, 2 g01, g11] (__init__.py:149-153)
                gi.opacity = g_opacity[i];
                const float2* cr = reinterpret_cast<const float2*>(g_color + (size_t)i * GH_NUM_CHANNELS);
#pragma unroll
                for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) { const float2 v = cr[k]; gi.color[2 * k] = v.x; gi.color[2 * k + 1] = v.y; }
            }
            gh_project_backward_one(A, i, s_rest + threadIdx.x * GH_PJ_REST, gi, go, cam);
        }
        // every row of every per-Gaussian gradient is written (zeros for culled Gaussians)
        if (d_mean2D_out) { d_mean2D_out[3 * (size_t)i] = gi.m2x; d_mean2D_out[3 * (size_t)i + 1] = gi.m2y; d_mean2D_out[3 * (size_t)i + 2] = 0.f; }
        d_xyz[3 * (size_t)i] = go.xyz[0]; d_xyz[3 * (size_t)i + 1] = go.xyz[1]; d_xyz[3 * (size_t)i + 2] = go.xyz[2];
        d_scaling[3 * (size_t)i] = go.scaling[0]; d_scaling[3 * (size_t)i + 1] = go.scaling[1]; d_scaling[3 * (size_t)i + 2] = go.scaling[2];
        reinterpret_cast<float4*>(d_rotation)[i] = make_float4(go.rotation[0], go.rotation[1], go.rotation[2], go.rotation[3]);
        if (d_dirs) { d_dirs[3 * (size_t)i] = go.dirs[0]; d_dirs[3 * (size_t)i + 1] = go.dirs[1]; d_dirs[3 * (size_t)i + 2] = go.dirs[2]; }
        d_fdc[3 * (size_t)i] = go.f_dc[0]; d_fdc[3 * (size_t)i + 1] = go.f_dc[1]; d_fdc[3 * (size_t)i + 2] = go.f_dc[2];
        if (d_opacity) d_opacity[i] = go.opacity;
        if (d_label) d_label[i] = go.label;
        if (d_conf) d_conf[i] = go.conf;
    }
    if (nan_flag != nullptr) {
        // the optimizer's NaN guard (train_gaussians.py:174-181) rides on the kernel that produces the gradients
        float acc = go.opacity + go.label + go.conf;
#pragma unroll
        for (int k = 0; k < 3; k++) acc += go.xyz[k] + go.scaling[k] + go.dirs[k] + go.f_dc[k];
#pragma unroll
        for (int k = 0; k < 4; k++) acc += go.rotation[k];
#pragma unroll
        for (int k = 0; k < GH_PJ_REST; k++) acc += go.rest[k];
        // a sum is NaN iff a term is NaN or +inf and -inf meet: both must stop the step
        if (__any_sync(0xffffffffu, acc != acc) && (threadIdx.x & 31) == 0) atomicOr(nan_flag, 1u);
    }
    // f_rest gradients leave through shared memory (coalesced 128-bit stores)
    __syncthreads();                                   // every thread is done reading its coefficients
    if (i < A.P) {
        float* row = s_rest + threadIdx.x * GH_PJ_REST;
#pragma unroll
        for (int e = 0; e < GH_PJ_REST; e++) row[e] = go.rest[e];
    }
    __syncthreads();
    gh_rest_store(s_rest, d_frest, A.P, row0);

    if (!want_cam) return;
    // ---- camera gradients: warp shuffle -> CTA partial -> the last CTA sums all partials in double, in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < GH_PJ_NCAM; k++) {
        float v = cam[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_cam[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < GH_PJ_NCAM) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < GH_PJ_THREADS / 32; w++) v += s_cam[w][threadIdx.x];
        cam_partial[(size_t)blockIdx.x * GH_PJ_NCAM + threadIdx.x] = v;
    }
    __threadfence();
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(cam_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_sum[GH_PJ_THREADS / 32][GH_PJ_NCAM];
    // warp w sums blocks w, w+4, ...; lane k < 29 owns accumulator k: a fixed order, hence deterministic
    if (lane < GH_PJ_NCAM) {
        double acc = 0.0;
        for (unsigned int b = warp; b < gridDim.x; b += GH_PJ_THREADS / 32) acc += (double)__ldcg(cam_partial + (size_t)b * GH_PJ_NCAM + lane);
        s_sum[warp][lane] = acc;
    }
    __syncthreads();
    if (threadIdx.x < GH_PJ_NCAM) {
        double acc = 0.0;
        for (int w = 0; w < GH_PJ_THREADS / 32; w++) acc += s_sum[w][threadIdx.x];
        const int k = threadIdx.x;
        const float v = (float)acc;
        if (k < 12) d_cam[4 * (k / 3) + (k % 3)] = v;                                  // V[i][j], j < 3
        else if (k < 24) { const int q = k - 12; const int col = q % 3; d_cam[16 + 4 * (q / 3) + (col == 2 ? 3 : col)] = v; }   // Pm[i][0,1,3]
        else if (k < 27) d_cam[32 + (k - 24)] = v;
        else d_cam[35 + (k - 27)] = v;
    }
    if (threadIdx.x == 0) {
        // entries that never receive a gradient
        d_cam[3] = 0.f; d_cam[7] = 0.f; d_cam[11] = 0.f; d_cam[15] = 0.f;
        d_cam[16 + 2] = 0.f; d_cam[16 + 6] = 0.f; d_cam[16 + 10] = 0.f; d_cam[16 + 14] = 0.f;
        *cam_ticket = 0u;
    }
}

int gh_proj_check(GhProjArgs& A, const char* who)
{
    char msg[160];
    if (A.P <= 0 || A.W <= 0 || A.H <= 0) { snprintf(msg, sizeof msg, "%s: P, width, height must be positive", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if (!A.xyz || !A.scaling || !A.rotation || !A.f_dc || !A.V || !A.Pm || !A.campos) { snprintf(msg, sizeof msg, "%s: missing mandatory pointer", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if (A.sh_degree < 0 || A.sh_degree > 3) { snprintf(msg, sizeof msg, "%s: sh_degree must be 0..3", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if (A.sh_degree > 0 && !A.f_rest) { snprintf(msg, sizeof msg, "%s: features_rest required for sh_degree > 0", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if ((size_t)A.rotation & 15) { snprintf(msg, sizeof msg, "%s: rotation must be 16-byte aligned", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if (A.dir_mode == 1 && !A.dirs) { snprintf(msg, sizeof msg, "%s: dirs required for dir_mode 1", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    if ((A.opacity_act < 2 && !A.opacity) || (A.label_act < 2 && !A.label) || (A.conf_act < 2 && !A.conf)) {
        snprintf(msg, sizeof msg, "%s: opacity / label / orient_conf pointer missing for the chosen activation", who);
        return gh_set_error(GH_E_INVALID_ARG, msg);
    }
    if (!(A.tanx > 0.f) || !(A.tany > 0.f)) { snprintf(msg, sizeof msg, "%s: tan_fov must be positive", who); return gh_set_error(GH_E_INVALID_ARG, msg); }
    return GH_OK;
}

GhProjArgs gh_proj_args(int P, int width, int height, const float* xyz, const float* scaling, const float* rotation,
                        const float* dirs, const float* f_dc, const float* f_rest, const float* opacity,
                        const float* label, const float* conf, const float* V, const float* Pm, const float* campos,
                        float tanx, float tany, float mod, int sh_degree, unsigned int flags, float det_eps)
{
    GhProjArgs A;
    A.P = P; A.W = width; A.H = height; A.mod = mod; A.det_eps = det_eps; A.tanx = tanx; A.tany = tany; A.sh_degree = sh_degree;
    A.scale_act = (int)(flags & 3u); A.opacity_act = (int)((flags >> 2) & 3u); A.label_act = (int)((flags >> 4) & 3u);
    A.conf_act = (int)((flags >> 6) & 3u); A.dir_mode = (int)((flags >> 8) & 3u);
    A.xyz = xyz; A.scaling = scaling; A.rotation = rotation; A.dirs = dirs; A.f_dc = f_dc; A.f_rest = f_rest;
    A.opacity = opacity; A.label = label; A.conf = conf; A.V = V; A.Pm = Pm; A.campos = campos;
    return A;
}

}  // namespace

extern "C" int gh_project_workspace_size(int P, size_t* bytes)
{
    gh_clear_error();
    if (P < 0 || !bytes) return gh_set_error(GH_E_INVALID_ARG, "gh_project_workspace_size: bad arguments");
    const size_t blocks = ((size_t)P + GH_PJ_THREADS - 1) / GH_PJ_THREADS;
    *bytes = 256 + blocks * GH_PJ_NCAM * sizeof(float);
    return GH_OK;
}

extern "C" int gh_project_forward(
    int P, int width, int height,
    const float* xyz, const float* scaling, const float* rotation, const float* dirs,
    const float* features_dc, const float* features_rest,
    const float* opacity, const float* label, const float* orient_conf,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float scale_modifier, int sh_degree, unsigned int flags, float det_eps,
    float* means2D, float* colors, float* opacities, float* conic, float* cov3D, unsigned char* visible,
    gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    gh_clear_error();
    GhProjArgs A = gh_proj_args(P, width, height, xyz, scaling, rotation, dirs, features_dc, features_rest, opacity, label,
                                orient_conf, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, scale_modifier, sh_degree, flags, det_eps);
    const int rc = gh_proj_check(A, "gh_project_forward");
    if (rc != GH_OK) return rc;
    if (!means2D || !colors || !opacities || !conic || !visible) return gh_set_error(GH_E_INVALID_ARG, "gh_project_forward: missing output pointer");
    if ((size_t)colors & 7) return gh_set_error(GH_E_INVALID_ARG, "gh_project_forward: colors must be 8-byte aligned");
    gh_project_forward_kernel<false><<<(P + GH_PJ_THREADS - 1) / GH_PJ_THREADS, GH_PJ_THREADS, 0, stream>>>(
        A, means2D, colors, opacities, conic, cov3D, visible, nullptr, nullptr, nullptr, nullptr, 0, 0);
    gh_count_launches(1);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GH_OK : gh_set_error(GH_E_CUDA, cudaGetErrorString(e));
}

// gh_project_forward + the rasterizer's first phase (gh_forward_preprocess) in one pass over the Gaussians: fills
// radii and the geometry / image workspaces, runs the tile scan and reads back R like gh_forward_preprocess does;
// continue with gh_forward_render.
extern "C" int gh_project_forward_binned(
    int P, int width, int height,
    const float* xyz, const float* scaling, const float* rotation, const float* dirs,
    const float* features_dc, const float* features_rest,
    const float* opacity, const float* label, const float* orient_conf,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float scale_modifier, int sh_degree, unsigned int flags, float det_eps,
    float* means2D, float* colors, float* opacities, float* conic, float* cov3D, unsigned char* visible,
    int* radii, char* geom_buffer, char* img_buffer, int* num_rendered, int* max_tile_len,
    gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    gh_clear_error();
    GhProjArgs A = gh_proj_args(P, width, height, xyz, scaling, rotation, dirs, features_dc, features_rest, opacity, label,
                                orient_conf, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, scale_modifier, sh_degree, flags, det_eps);
    const int rc = gh_proj_check(A, "gh_project_forward_binned");
    if (rc != GH_OK) return rc;
    if (!means2D || !colors || !opacities || !conic || !visible || !radii || !geom_buffer || !img_buffer || !num_rendered)
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_forward_binned: missing output pointer");
    if ((size_t)colors & 7) return gh_set_error(GH_E_INVALID_ARG, "gh_project_forward_binned: colors must be 8-byte aligned");
    const int gx = (width + GH_BLOCK_X - 1) / GH_BLOCK_X, gy = (height + GH_BLOCK_Y - 1) / GH_BLOCK_Y;
    const int T = gx * gy;
    if ((unsigned long long)gx * gx * gy >= (1ull << 32))     // exactness bound of the tile enumeration (gh_warp_rects)
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_forward_binned: image too large (tile grid gx * gx * gy must stay below 2^32)");
    GhGeomWS geom = GhGeomWS::carve(geom_buffer, (size_t)P);
    GhImgWS img = GhImgWS::carve(img_buffer, (size_t)width * height, (size_t)T);
    // ctrl + tile histogram are contiguous: one memset
    cudaError_t e = cudaMemsetAsync(img.ctrl, 0, 256 + gh_align_up((size_t)T * 4, 256), stream);
    if (e != cudaSuccess) return gh_set_error(GH_E_CUDA, "gh_project_forward_binned: memset(tile histogram) failed");
    gh_project_forward_kernel<true><<<(P + GH_PJ_THREADS - 1) / GH_PJ_THREADS, GH_PJ_THREADS, 0, stream>>>(
        A, means2D, colors, opacities, conic, cov3D, visible, radii, geom.geo, geom.depth, img.tile_count, gx, gy);
    gh_launch_tile_scan(T, img, stream);
    gh_count_launches(2);
    GhCtrl h;
    e = cudaMemcpyAsync(&h, img.ctrl, sizeof(GhCtrl), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return gh_set_error(GH_E_CUDA, cudaGetErrorString(e));
    *num_rendered = (int)h.num_rendered;
    if (max_tile_len) *max_tile_len = (int)h.max_tile_len;
    return GH_OK;
}

extern "C" int gh_project_backward(
    int P, int width, int height,
    const float* xyz, const float* scaling, const float* rotation, const float* dirs,
    const float* features_dc, const float* features_rest,
    const float* opacity, const float* label, const float* orient_conf,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, float scale_modifier, int sh_degree, unsigned int flags, float det_eps,
    const unsigned char* visible,
    const char* geom_buffer,
    const float* dL_dmeans2D, const float* dL_dconic, const float* dL_dcolors, const float* dL_dopacity,
    float* d_xyz, float* d_scaling, float* d_rotation, float* d_dirs, float* d_features_dc, float* d_features_rest,
    float* d_opacity, float* d_label, float* d_orient_conf, float* d_means2D, float* d_camera,
    unsigned int* nan_flag, void* workspace, gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    gh_clear_error();
    GhProjArgs A = gh_proj_args(P, width, height, xyz, scaling, rotation, dirs, features_dc, features_rest, opacity, label,
                                orient_conf, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, scale_modifier, sh_degree, flags, det_eps);
    const int rc = gh_proj_check(A, "gh_project_backward");
    if (rc != GH_OK) return rc;
    if (!visible || !d_xyz || !d_scaling || !d_rotation || !d_features_dc || !d_features_rest)
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_backward: missing mandatory pointer");
    if (geom_buffer == nullptr && (!dL_dmeans2D || !dL_dconic || !dL_dcolors || !dL_dopacity))
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_backward: pass the geometry workspace of gh_backward or the four incoming gradients");
    if (((size_t)d_rotation & 15) || (dL_dconic && ((size_t)dL_dconic & 15)) || (dL_dcolors && ((size_t)dL_dcolors & 7)))
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_backward: d_rotation / dL_dconic must be 16-byte, dL_dcolors 8-byte aligned");
    if (d_camera != nullptr && (workspace == nullptr || ((size_t)workspace & 15)))
        return gh_set_error(GH_E_INVALID_ARG, "gh_project_backward: camera gradients need the 16-byte aligned workspace");
    const float* acc16 = nullptr;
    if (geom_buffer != nullptr) acc16 = GhGeomWS::carve(const_cast<char*>(geom_buffer), (size_t)P).acc16;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    float* partial = workspace ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256) : nullptr;
    if (d_camera != nullptr) {
        // the ticket word is reset by the kernel's last CTA; zero it here for the very first use of a workspace
        if (cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream) != cudaSuccess)
            return gh_set_error(GH_E_CUDA, "gh_project_backward: memset(ticket) failed");
    }
    gh_project_backward_kernel<<<(P + GH_PJ_THREADS - 1) / GH_PJ_THREADS, GH_PJ_THREADS, 0, stream>>>(
        A, visible, acc16, dL_dmeans2D, dL_dconic, dL_dcolors, dL_dopacity,
        d_xyz, d_scaling, d_rotation, d_dirs, d_features_dc, d_features_rest, d_opacity, d_label, d_orient_conf,
        d_means2D, partial, ticket, d_camera, nan_flag);
    gh_count_launches(1);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GH_OK : gh_set_error(GH_E_CUDA, cudaGetErrorString(e));
}
