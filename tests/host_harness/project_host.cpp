// TEST INFRASTRUCTURE: compiles the product's per-Gaussian projection arithmetic
// (gaussianhaircut_b200/csrc/gh_project_math.h, the functions the CUDA kernels call) for the host, so that the
// hand-derived backward can be checked against PyTorch autograd where there is no GPU (tests/test_project_cpu.py).
// Not part of libgh_raster.so; the product has no CPU path.
#include "../../gaussianhaircut_b200/csrc/gh_project_math.h"

#include <cstring>

static GhProjArgs make_args(int P, int W, int H, const float* xyz, const float* scaling, const float* rotation,
                            const float* dirs, const float* f_dc, const float* f_rest, const float* opacity,
                            const float* label, const float* conf, const float* V, const float* Pm, const float* campos,
                            float tanx, float tany, float mod, int sh_degree, unsigned int flags, float det_eps)
{
    GhProjArgs A;
    A.P = P; A.W = W; A.H = H; A.mod = mod; A.det_eps = det_eps; A.tanx = tanx; A.tany = tany; A.sh_degree = sh_degree;
    A.scale_act = (int)(flags & 3u); A.opacity_act = (int)((flags >> 2) & 3u); A.label_act = (int)((flags >> 4) & 3u);
    A.conf_act = (int)((flags >> 6) & 3u); A.dir_mode = (int)((flags >> 8) & 3u);
    A.xyz = xyz; A.scaling = scaling; A.rotation = rotation; A.dirs = dirs; A.f_dc = f_dc; A.f_rest = f_rest;
    A.opacity = opacity; A.label = label; A.conf = conf; A.V = V; A.Pm = Pm; A.campos = campos;
    return A;
}

extern "C" void gh_host_project_forward(
    int P, int W, int H, const float* xyz, const float* scaling, const float* rotation, const float* dirs,
    const float* f_dc, const float* f_rest, const float* opacity, const float* label, const float* conf,
    const float* V, const float* Pm, const float* campos, float tanx, float tany, float mod, int sh_degree,
    unsigned int flags, float det_eps,
    float* means2D, float* colors, float* opac, float* conic, float* cov3D, unsigned char* visible)
{
    GhProjArgs A = make_args(P, W, H, xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, V, Pm, campos,
                             tanx, tany, mod, sh_degree, flags, det_eps);
    static const float zero_rest[GH_PJ_REST] = {0};
    for (int i = 0; i < P; i++) {
        GhProjOut o;
        gh_project_forward_one(A, i, f_rest ? f_rest + (size_t)i * GH_PJ_REST : zero_rest, cov3D != nullptr, o);
        for (int k = 0; k < 3; k++) { means2D[3 * i + k] = o.m2[k]; conic[3 * i + k] = o.conic[k]; }
        opac[i] = o.opacity;
        for (int k = 0; k < GH_PJ_CHANNELS; k++) colors[GH_PJ_CHANNELS * i + k] = o.color[k];
        if (cov3D) for (int k = 0; k < 6; k++) cov3D[6 * i + k] = o.cov3D[k];
        visible[i] = o.visible ? 1 : 0;
    }
}

// g_conic3 is the PUBLIC 3-vector conic gradient [g00, 2 g01, g11]
extern "C" void gh_host_project_backward(
    int P, int W, int H, const float* xyz, const float* scaling, const float* rotation, const float* dirs,
    const float* f_dc, const float* f_rest, const float* opacity, const float* label, const float* conf,
    const float* V, const float* Pm, const float* campos, float tanx, float tany, float mod, int sh_degree,
    unsigned int flags, float det_eps, const unsigned char* visible,
    const float* g_means2D, const float* g_conic3, const float* g_colors, const float* g_opacity,
    float* d_xyz, float* d_scaling, float* d_rotation, float* d_dirs, float* d_fdc, float* d_frest,
    float* d_opacity, float* d_label, float* d_conf, double* d_cam29)
{
    GhProjArgs A = make_args(P, W, H, xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, V, Pm, campos,
                             tanx, tany, mod, sh_degree, flags, det_eps);
    static const float zero_rest[GH_PJ_REST] = {0};
    for (int k = 0; k < GH_PJ_NCAM; k++) d_cam29[k] = 0.0;
    for (int i = 0; i < P; i++) {
        GhProjGradOut go;
        std::memset(&go, 0, sizeof go);
        float cam[GH_PJ_NCAM] = {0};
        if (visible[i]) {
            GhProjGradIn gi;
            gi.m2x = g_means2D[3 * i]; gi.m2y = g_means2D[3 * i + 1];
            for (int k = 0; k < 3; k++) gi.con[k] = g_conic3[3 * i + k];
            for (int k = 0; k < GH_PJ_CHANNELS; k++) gi.color[k] = g_colors[GH_PJ_CHANNELS * i + k];
            gi.opacity = g_opacity[i];
            gh_project_backward_one(A, i, f_rest ? f_rest + (size_t)i * GH_PJ_REST : zero_rest, gi, go, cam);
        }
        for (int k = 0; k < 3; k++) { d_xyz[3 * i + k] = go.xyz[k]; d_scaling[3 * i + k] = go.scaling[k]; d_fdc[3 * i + k] = go.f_dc[k]; }
        if (d_dirs) for (int k = 0; k < 3; k++) d_dirs[3 * i + k] = go.dirs[k];
        for (int k = 0; k < 4; k++) d_rotation[4 * i + k] = go.rotation[k];
        for (int k = 0; k < GH_PJ_REST; k++) d_frest[(size_t)GH_PJ_REST * i + k] = go.rest[k];
        d_opacity[i] = go.opacity; d_label[i] = go.label; d_conf[i] = go.conf;
        for (int k = 0; k < GH_PJ_NCAM; k++) d_cam29[k] += (double)cam[k];
    }
}
