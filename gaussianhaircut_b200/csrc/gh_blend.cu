// Stage 3 / 4: per-tile front-to-back alpha compositing of the 10-channel feature vector
// (RGB, label, ones, dir2D(3), orientation confidence, depth) and its back-to-front backward.
//
// Per-pixel arithmetic and every skip/stop decision follow the reference kernels renderCUDA
// (cuda_rasterizer/forward.cu:287-400, backward.cu:403-561) bit for bit; what is new is the work
// decomposition:
//   * a tile is blended by 8 warps, each owning an 8x4 pixel block;
//   * each warp first tests the staged Gaussians against its own 8x4 block with an exact-conservative
//     "can this splat reach alpha >= 1/255 anywhere in the block" test (one Gaussian per lane) and
//     then walks only the survivors -- strand Gaussians are long thin ellipses whose 3-sigma square
//     over-covers most of the tile, so most (pixel, Gaussian) pairs of the reference are skipped here;
//   * all 10 feature channels are staged in shared memory with the geometry (the reference forward
//     re-reads them from global memory per contributing pair, forward.cu:381);
//   * the backward reduces the 16 gradient components of a Gaussian over the 32 pixels of a warp
//     with a 16-shuffle transposing butterfly and issues ONE red.global.add per component per
//     (warp, Gaussian) instead of one atomicAdd per component per (pixel, Gaussian)
//     (backward.cu:527,549-558).
#include "gh_common.cuh"
#include "gh_kernels.h"

#define GH_CHUNK 256

namespace {

struct GhPixEval {
    float dx, dy, power, G, alpha;
    bool ok;   // passes the reference's `power > 0` and `alpha < 1/255` skips
};

__device__ __forceinline__ GhPixEval gh_eval(const float4 g0, const float4 g1, float pxf, float pyf) {
    GhPixEval e;
    e.dx = GH_SUB(g0.x, pxf);
    e.dy = GH_SUB(g0.y, pyf);
    e.power = gh_power(e.dx, e.dy, g0.z, g0.w, g1.x);
    e.G = 0.f; e.alpha = 0.f;
    e.ok = false;
    if (!(e.power > 0.0f)) {
        e.G = expf(e.power);
        e.alpha = fminf(0.99f, GH_MUL(g1.y, e.G));
        e.ok = !(e.alpha < 1.0f / 255.0f);
    }
    return e;
}

// stage `cnt` instances (list positions first .. first+cnt-1, ascending or descending) into smem
__device__ __forceinline__ void gh_stage_chunk(const uint64_t* __restrict__ inst, long long pos, bool valid,
                                               const GhGeo* __restrict__ geo, const float* __restrict__ features,
                                               float4* s_g0, float4* s_g1, float2* s_feat, uint32_t* s_id, int slot)
{
    if (valid) {
        const uint32_t id = (uint32_t)inst[pos];
        const float4* gp = reinterpret_cast<const float4*>(geo + id);
        const float4 a = __ldg(gp), b = __ldg(gp + 1);
        const float2* fp = reinterpret_cast<const float2*>(features + (size_t)id * GH_NUM_CHANNELS);
        float2 f[GH_NUM_CHANNELS / 2];
#pragma unroll
        for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) f[k] = __ldg(fp + k);
        s_g0[slot] = a; s_g1[slot] = b;
#pragma unroll
        for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) s_feat[slot * (GH_NUM_CHANNELS / 2) + k] = f[k];
        if (s_id) s_id[slot] = id;
    }
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256)
gh_blend_forward_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ inst,
                        const GhGeo* __restrict__ geo, const float* __restrict__ features,
                        int W, int H, int gx, const float* __restrict__ bg,
                        float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                        float* __restrict__ out)
{
    __shared__ float4 s_g0[GH_CHUNK];
    __shared__ float4 s_g1[GH_CHUNK];
    __shared__ float2 s_feat[GH_CHUNK * GH_NUM_CHANNELS / 2];

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx0 = tx * GH_BLOCK_X + (warp & 1) * 8, by0 = ty * GH_BLOCK_Y + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;
    const float rx0 = (float)bx0, rx1 = (float)(bx0 + 7), ry0 = (float)by0, ry1 = (float)(by0 + 3);

    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);

    float T = 1.0f;
    float C[GH_NUM_CHANNELS];
#pragma unroll
    for (int ch = 0; ch < GH_NUM_CHANNELS; ch++) C[ch] = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    bool warp_done = (__ballot_sync(0xffffffffu, done) == 0xffffffffu);

    for (int base = 0; base < n; base += GH_CHUNK) {
        // block-wide early exit (reference: __syncthreads_count(done) == BLOCK_SIZE); also the
        // barrier that protects the staging buffers of the previous chunk
        if (__syncthreads_and(warp_done)) break;
        const int cnt = min(GH_CHUNK, n - base);
        gh_stage_chunk(inst, (long long)rg.x + base + tid, tid < cnt, geo, features,
                       s_g0, s_g1, s_feat, nullptr, tid);
        __syncthreads();
        if (warp_done) continue;

        for (int sub = 0; sub < cnt; sub += 32) {
            const int j = sub + lane;
            bool hit = false;
            if (j < cnt) {
                const float4 a = s_g0[j], b = s_g1[j];
                GhGeo g; g.x = a.x; g.y = a.y; g.ca = a.z; g.cb = a.w; g.cc = b.x; g.op = b.y; g.thr = b.z; g.pd = b.w;
                hit = gh_cull_hit(g, rx0, rx1, ry0, ry1);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int bpos = __ffs(mask) - 1;
                mask &= mask - 1;
                const int jj = sub + bpos;
                if (!done) {
                    const float4 g0 = s_g0[jj], g1 = s_g1[jj];
                    const GhPixEval e = gh_eval(g0, g1, pxf, pyf);
                    if (e.ok) {
                        const float test_T = GH_MUL(T, GH_SUB(1.0f, e.alpha));
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
#pragma unroll
                            for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) {
                                const float2 f = s_feat[jj * (GH_NUM_CHANNELS / 2) + k];
                                C[2 * k + 0] = GH_FMA(T, GH_MUL(e.alpha, f.x), C[2 * k + 0]);
                                C[2 * k + 1] = GH_FMA(T, GH_MUL(e.alpha, f.y), C[2 * k + 1]);
                            }
                            T = test_T;
                            last = (uint32_t)(base + jj + 1);
                        }
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, done) == 0xffffffffu) { warp_done = true; break; }
        }
    }

    if (inside) {
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int ch = 0; ch < GH_NUM_CHANNELS; ch++)
            out[ch * plane + pix] = GH_FMA(T, __ldg(bg + ch), C[ch]);
    }
}

// ------------------------------------------------------------------------------------------ backward
// 16 values per lane -> lane l ends with the warp-wide sum of value  v(l) = l>>1  (both lanes of a pair hold it)
__device__ __forceinline__ float gh_warp_reduce16(float (&v)[16], int lane) {
    float w8[8], w4[4], w2[2], w1;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float send = b4 ? v[i] : v[i + 8];
        const float keep = b4 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float send = b3 ? w8[i] : w8[i + 4];
        const float keep = b3 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float send = b2 ? w4[i] : w4[i + 2];
        const float keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float send = b1 ? w2[0] : w2[1];
        const float keep = b1 ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    return w1;   // component index = (b4?8:0) + (b3?4:0) + (b2?2:0) + (b1?1:0)
}

__global__ void __launch_bounds__(256)
gh_blend_backward_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ inst,
                         const GhGeo* __restrict__ geo, const float* __restrict__ features,
                         int W, int H, int gx, const float* __restrict__ bg,
                         const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                         const float* __restrict__ dL_dpix,
                         float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic,
                         float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor)
{
    __shared__ float4 s_g0[GH_CHUNK];
    __shared__ float4 s_g1[GH_CHUNK];
    __shared__ float2 s_feat[GH_CHUNK * GH_NUM_CHANNELS / 2];
    __shared__ uint32_t s_id[GH_CHUNK];
    __shared__ uint32_t s_warp_last[8];

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx0 = tx * GH_BLOCK_X + (warp & 1) * 8, by0 = ty * GH_BLOCK_Y + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;
    const float rx0 = (float)bx0, rx1 = (float)(bx0 + 7), ry0 = (float)by0, ry1 = (float)(by0 + 3);
    const size_t pix = (size_t)py * W + px;
    const size_t plane = (size_t)H * W;

    const uint2 rg = ranges[tile];

    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const uint32_t last = inside ? n_contrib[pix] : 0u;   // pixel blends list positions 1..last

    float dL_dpixel[GH_NUM_CHANNELS];
    float bg_dot_dpixel = 0.f;
#pragma unroll
    for (int ch = 0; ch < GH_NUM_CHANNELS; ch++) {
        dL_dpixel[ch] = inside ? dL_dpix[ch * plane + pix] : 0.f;
        bg_dot_dpixel += __ldg(bg + ch) * dL_dpixel[ch];
    }
    float accum_rec[GH_NUM_CHANNELS], last_color[GH_NUM_CHANNELS];
#pragma unroll
    for (int ch = 0; ch < GH_NUM_CHANNELS; ch++) { accum_rec[ch] = 0.f; last_color[ch] = 0.f; }
    float last_alpha = 0.f;

    // pixel-coordinate -> NDC chain rule factors (backward.cu:464-465)
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    uint32_t warp_last = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    if (lane == 0) s_warp_last[warp] = warp_last;
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) tile_last = max(tile_last, s_warp_last[w]);
    // nothing beyond tile_last is blended by any pixel of this tile
    const int n = (int)tile_last;
    const int nchunks = (n + GH_CHUNK - 1) / GH_CHUNK;

    for (int c = nchunks - 1; c >= 0; c--) {
        const int base = c * GH_CHUNK;
        const int cnt = min(GH_CHUNK, n - base);
        __syncthreads();
        gh_stage_chunk(inst, (long long)rg.x + base + tid, tid < cnt, geo, features,
                       s_g0, s_g1, s_feat, s_id, tid);
        __syncthreads();
        if ((uint32_t)base >= warp_last) continue;

        for (int sub = ((cnt - 1) >> 5) << 5; sub >= 0; sub -= 32) {
            const int j = sub + lane;
            bool hit = false;
            if (j < cnt && (uint32_t)(base + j) < warp_last) {
                const float4 a = s_g0[j], b = s_g1[j];
                GhGeo g; g.x = a.x; g.y = a.y; g.ca = a.z; g.cb = a.w; g.cc = b.x; g.op = b.y; g.thr = b.z; g.pd = b.w;
                hit = gh_cull_hit(g, rx0, rx1, ry0, ry1);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int bpos = 31 - __clz(mask);   // back to front
                mask &= ~(1u << bpos);
                const int jj = sub + bpos;
                const float4 g0 = s_g0[jj], g1 = s_g1[jj];
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = 0.f;
                bool contrib = false;
                // reference: contributor--; if (contributor >= last_contributor) continue;
                if ((uint32_t)(base + jj) < last) {
                    const GhPixEval e = gh_eval(g0, g1, pxf, pyf);
                    if (e.ok) {
                        contrib = true;
                        const float alpha = e.alpha, G = e.G;
                        T = GH_DIV(T, GH_SUB(1.0f, alpha));
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.f;
#pragma unroll
                        for (int k = 0; k < GH_NUM_CHANNELS / 2; k++) {
                            const float2 f = s_feat[jj * (GH_NUM_CHANNELS / 2) + k];
                            {
                                const int ch = 2 * k;
                                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                                last_color[ch] = f.x;
                                dL_dalpha += (f.x - accum_rec[ch]) * dL_dpixel[ch];
                                v[ch] = dchannel_dcolor * dL_dpixel[ch];
                            }
                            {
                                const int ch = 2 * k + 1;
                                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                                last_color[ch] = f.y;
                                dL_dalpha += (f.y - accum_rec[ch]) * dL_dpixel[ch];
                                v[ch] = dchannel_dcolor * dL_dpixel[ch];
                            }
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        // alpha also scales how much background shows through (backward.cu:535-538)
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                        const float dL_dG = g1.y * dL_dalpha;
                        const float gdx = G * e.dx, gdy = G * e.dy;
                        const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                        const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                        v[10] = dL_dG * dG_ddelx * ddelx_dx;
                        v[11] = dL_dG * dG_ddely * ddely_dy;
                        v[12] = -0.5f * gdx * e.dx * dL_dG;
                        v[13] = -0.5f * gdx * e.dy * dL_dG;
                        v[14] = -0.5f * gdy * e.dy * dL_dG;
                        v[15] = G * dL_dalpha;
                    }
                }
                if (__ballot_sync(0xffffffffu, contrib) == 0u) continue;
                const float sum = gh_warp_reduce16(v, lane);
                if ((lane & 1) == 0) {
                    const int comp = lane >> 1;
                    const uint32_t id = s_id[jj];
                    float* dst;
                    if (comp < 10)       dst = dL_dcolor + (size_t)id * GH_NUM_CHANNELS + comp;
                    else if (comp < 12)  dst = dL_dmean2D + (size_t)id * 3 + (comp - 10);
                    else if (comp < 14)  dst = dL_dconic + (size_t)id * 4 + (comp - 12);
                    else if (comp == 14) dst = dL_dconic + (size_t)id * 4 + 3;
                    else                 dst = dL_dopacity + id;
                    atomicAdd(dst, sum);
                }
            }
        }
    }
}

}  // namespace

void gh_launch_blend_forward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                             const float* features, const float* bg, float* out_color,
                             cudaStream_t stream)
{
    gh_blend_forward_kernel<<<gx * gy, 256, 0, stream>>>(img.ranges, bin.inst, geom.geo, features,
                                                         W, H, gx, bg, img.final_T, img.n_contrib, out_color);
}

void gh_launch_blend_backward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                              const float* features, const float* bg, const float* dL_dpix,
                              float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                              cudaStream_t stream)
{
    gh_blend_backward_kernel<<<gx * gy, 256, 0, stream>>>(img.ranges, bin.inst, geom.geo, features,
                                                          W, H, gx, bg, img.final_T, img.n_contrib, dL_dpix,
                                                          dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
}
