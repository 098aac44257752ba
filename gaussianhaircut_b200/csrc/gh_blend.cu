// Stage 3 / 4: per-tile front-to-back alpha compositing of the 10-channel feature vector
// (RGB, label, ones, dir2D(3), orientation confidence, depth) and its back-to-front backward.
//
// Per-pixel arithmetic and every skip/stop decision follow the reference kernels renderCUDA
// (cuda_rasterizer/forward.cu:287-400, backward.cu:403-561) bit for bit.  The work decomposition is
// new and built around the shape of strand-aligned Gaussians (long thin ellipses: of the 256 pixels
// of a tile a splat typically reaches alpha >= 1/255 on ~15):
//
//   forward  (gh_blend_forward_kernel, CTA = tile, 8 warps, one pixel per lane)
//   * lists of <= 2048 records are depth-sorted by the CTA itself (gh_bucket_sort_tile) before blending;
//   * the sorted list is staged in chunks of 256 into shared memory with cp.async (LDGSTS), double
//     buffered, indices prefetched two chunks ahead;
//   * each staged Gaussian is scan-converted ONCE per tile: the exact-conservative x-span of
//     {alpha >= 1/255} on every pixel row (two rows per packed-FP32 instruction) gives 16-bit row masks,
//     a 5-step shuffle transpose turns the 32 (Gaussian) x 32 (pixel) bit matrix into one hit list per
//     pixel, and every lane walks only ITS pixel's list -- a skipped Gaussian costs nothing;
//   * the 10 colour channels are accumulated as 5 x (FMUL2 + FFMA2), bit-identical to FMUL + FFMA.
//   backward (gh_blend_backward_kernel, CTA = tile, 4 warps, a vertical pixel pair per lane)
//   * one 512-record window staged at a time; Gaussians are scan-converted against the tile's 32 blocks
//     of 4x2 pixels and a shuffle transpose leaves one bit list per block;
//   * a block is owned by 4 lanes; the 8 blocks of a warp walk their own lists in lock-step (one warp
//     instruction works on EIGHT different Gaussians); blocks are assigned to warps by list length;
//   * the pixel pair is evaluated in packed FP32 (.x upper pixel, .y lower pixel); the 16 gradient
//     components are summed over the block's 4 lanes through shared memory and leave as ONE
//     REDG.E.ADD.F32x4 per lane into a 64-byte per-Gaussian record -- instead of 16 scalar atomicAdd
//     per (pixel, Gaussian) pair in the reference (backward.cu:527,549-558);
//   * dL/dalpha uses the scalar form of the reference's per-channel suffix recursion
//     (backward.cu:519-523):  sum_ch (c - accum_rec)[ch] dL[ch]  =  c.dL  -  A,   A' = a_last (c_last.dL) + (1-a_last) A.
#include "gh_common.cuh"
#include "gh_kernels.h"

#define GH_HALF_C (GH_NUM_CHANNELS / 2)

// ---- build-time tunables (defaults = the measured best; tools/bench_variants.py rebuilds with -D overrides)
#ifndef GH_TILE_ORDER
#define GH_TILE_ORDER 1          // 1: blend CTAs take tiles by decreasing list length (gh_tile_scan_kernel's permutation)
#endif
#ifndef GH_FWD_NBUF
#define GH_FWD_NBUF 2            // forward staging: 2 = double-buffered chunks, 1 = one single-buffered window (other CTAs cover the gather)
#endif
#ifndef GH_CHUNK
#define GH_CHUNK 256             // forward staging chunk / window (records); 512 with GH_FWD_NBUF 1 keeps the same shared memory
#endif
#ifndef GH_BWD_CHUNK
#define GH_BWD_CHUNK 384         // backward staging window (records)
#endif
#ifndef GH_BWD_MIN_CTAS
#define GH_BWD_MIN_CTAS 5        // __launch_bounds__ second argument of the backward kernel (register cap)
#endif
#ifndef GH_BWD_LPB
#define GH_BWD_LPB 2             // lanes per backward block: 4 = 4x2-pixel blocks (8 per warp), 2 = 2x2-pixel blocks (16 per warp)
#endif
#define GH_BWD_NBLK (128 / GH_BWD_LPB)          // blocks per tile
#define GH_BWD_BPW (32 / GH_BWD_LPB)            // blocks per warp
#define GH_BWD_BW GH_BWD_LPB                    // block width in pixels (a lane owns a vertical pixel pair)
#define GH_BWD_NBX (16 / GH_BWD_BW)             // blocks across the tile
#ifndef GH_BWD_FAST_EXP
#define GH_BWD_FAST_EXP 0        // 1: ex2.approx based exp in the backward (gradients only; decisions may differ in the last ulp)
#endif

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

__device__ __forceinline__ void gh_cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void gh_cp_async8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void gh_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void gh_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE-rounded results per issue slot).  The blend
// kernels are bound by instruction issue, not by the FP32 pipe, so pairing halves their FP32 cost.
__device__ __forceinline__ float2 gh_f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 gh_neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 gh_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 gh_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 gh_sub2(float2 a, float2 b) { return __fadd2_rn(a, gh_neg2(b)); }
__device__ __forceinline__ float2 gh_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float gh_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct GhStage {
    float4 g0[GH_FWD_NBUF][GH_CHUNK];
    float4 g1[GH_FWD_NBUF][GH_CHUNK];
    float2 feat[GH_FWD_NBUF][GH_CHUNK * GH_HALF_C];
    uint32_t id[GH_FWD_NBUF][GH_CHUNK];
};
static_assert(sizeof(float4) * 2 * GH_FWD_NBUF * GH_CHUNK >= 8u * GH_INKERNEL_SORT_MAX, "sort buffer A (g0+g1) too small");
static_assert(sizeof(float2) * GH_FWD_NBUF * GH_CHUNK * GH_HALF_C >= 8u * GH_INKERNEL_SORT_MAX, "sort buffer B (feat) too small");

// issue the gather of one instance (geometry record + feature row) into slot `slot` of buffer `buf`
__device__ __forceinline__ void gh_stage_issue(GhStage& st, int buf, int slot, uint32_t id,
                                               const GhGeo* __restrict__ geo, const float* __restrict__ features) {
    const float4* gp = reinterpret_cast<const float4*>(geo + id);
    gh_cp_async16(&st.g0[buf][slot], gp);
    gh_cp_async16(&st.g1[buf][slot], gp + 1);
    const float2* fp = reinterpret_cast<const float2*>(features + (size_t)id * GH_NUM_CHANNELS);
#pragma unroll
    for (int k = 0; k < GH_HALF_C; k++) gh_cp_async8(&st.feat[buf][slot * GH_HALF_C + k], fp + k);
    st.id[buf][slot] = id;
}

// Per-PIXEL instance lists of one staged chunk (forward pass).  Warp w scan-converts Gaussians
// [32w, 32w+32) of the chunk (one per lane): on every pixel row the exact-conservative x-span of
// {alpha >= 1/255} (same quadratic as gh_block_mask) becomes a 16-bit column mask; two rows form the
// 32-bit coverage word of the 32 pixels owned by one warp (pixel p = 32*(row/2) + 16*(row%2) + col).
// The 32 (Gaussian) x 32 (pixel) bit matrix is then transposed inside the warp with 5 shuffle steps,
// so lane P ends with "which of my warp's 32 Gaussians cover pixel P" and stores it: no atomics, no
// clearing pass.  pixbits is [word = builder warp][pixel] -> a warp reads consecutive banks.
__device__ __forceinline__ uint32_t gh_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = (s == 16) ? 0x0000ffffu : (s == 8) ? 0x00ff00ffu : (s == 4) ? 0x0f0f0f0fu
                         : (s == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? (((y >> s) & m) | (x & ~m)) : ((x & m) | ((y & m) << s));
    }
    return x;
}

__device__ __forceinline__ void gh_build_pixel_lists(const GhStage& st, uint32_t* pixbits, int buf, int cnt,
                                                     int warp, int lane, float tx0, float ty0) {
    // `warp` = index of the batch of 32 staged Gaussians this warp scan-converts (= list word index)
    const int j = warp * 32 + lane;
    uint32_t rowmask[16];
#pragma unroll
    for (int r = 0; r < 16; r++) rowmask[r] = 0u;
    float gx = 0.f, gy = 0.f, b = 0.f, ia = 0.f, aQ = -1.f, bbac = 0.f;
    int r0 = 16, r1 = -1;            // rows of the tile this Gaussian can reach
    bool all_pixels = false;
    if (j < cnt) {
        const float4 g0 = st.g0[buf][j], g1 = st.g1[buf][j];
        gx = g0.x; gy = g0.y; b = g0.w;
        const float a = g0.z, c = g1.x, thr = g1.z, pd = g1.w;
        const float mxd = fmaxf(fabsf(gx - tx0), fabsf(gx - (tx0 + 15.f)));
        const float myd = fmaxf(fabsf(gy - ty0), fabsf(gy - (ty0 + 15.f)));
        const float slack = 1e-3f + 2e-5f * (a * mxd * mxd + c * myd * myd + 2.f * fabsf(b) * mxd * myd);
        const float Q = 2.f * (thr + slack);
        if (pd == 0.f || Q != Q) {
            all_pixels = true; r0 = 0; r1 = 15;                     // no bound available: every pixel
        } else if (Q >= 0.f) {
            ia = __frcp_rn(a);
            aQ = a * Q;
            bbac = b * b - a * c;                                   // < 0 for a PD conic
            // rows with a real span: dy^2 < aQ / (ac - b^2)
            const float ry = sqrtf(aQ / (-bbac)) + 0.01f;
            r0 = 0; r1 = 15;
            if (ry == ry) { r0 = max(0, __float2int_ru(gy - ry - ty0)); r1 = min(15, __float2int_rd(gy + ry - ty0)); }
        }
    }
    // warp-uniform row window: the 32 Gaussians of a warp are neighbours in depth, not in space, but
    // strand Gaussians are small, so most rows are empty for the whole warp
    int wr0 = r0, wr1 = r1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wr0 = min(wr0, __shfl_xor_sync(0xffffffffu, wr0, o));
        wr1 = max(wr1, __shfl_xor_sync(0xffffffffu, wr1, o));
    }
    // two pixel rows per instruction (FFMA2 / FMUL2 / FADD2); a row without a real span has
    // disc <= 0 and is masked out below
    const float2 bbac2 = gh_f2(bbac), aQ2 = gh_f2(aQ), b2 = gh_f2(b), ia2 = gh_f2(ia), gx2 = gh_f2(gx);
    const float2 off2 = gh_f2(-0.01f), ntx2 = gh_f2(-tx0);
#pragma unroll
    for (int t = 0; t < 8; t++) {
        if (2 * t + 1 < wr0 || 2 * t > wr1) continue;               // uniform
        const float2 dy = make_float2(gy - (ty0 + (float)(2 * t)), gy - (ty0 + (float)(2 * t + 1)));
        // a dx^2 + 2 b dy dx + (c dy^2 - Q) <= 0,  dx = gx - px
        const float2 disc = gh_fma2(gh_mul2(dy, dy), bbac2, aQ2);
        const float2 sq = gh_mul2(disc, make_float2(rsqrtf(disc.x), rsqrtf(disc.y)));
        const float2 hb = gh_mul2(b2, dy);
        const float2 lo = gh_add2(gh_add2(gh_fma2(gh_sub2(hb, sq), ia2, gx2), off2), ntx2);
        const float2 hi = gh_add2(gh_add2(gh_fma2(gh_add2(hb, sq), ia2, gx2), gh_neg2(off2)), ntx2);
        const int p0a = max(0, __float2int_ru(lo.x)), p1a = min(15, __float2int_rd(hi.x));
        const int p0b = max(0, __float2int_ru(lo.y)), p1b = min(15, __float2int_rd(hi.y));
        const uint32_t ma = (disc.x > 0.f && p0a <= p1a) ? (((2u << (p1a - p0a)) - 1u) << p0a) : 0u;
        const uint32_t mb = (disc.y > 0.f && p0b <= p1b) ? (((2u << (p1b - p0b)) - 1u) << p0b) : 0u;
        rowmask[2 * t] = all_pixels ? 0xffffu : ma;      // rows outside [r0, r1] have disc <= 0
        rowmask[2 * t + 1] = all_pixels ? 0xffffu : mb;
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t word = rowmask[2 * t] | (rowmask[2 * t + 1] << 16);   // my coverage of warp t's pixels
        uint32_t col = 0u;
        if (__any_sync(0xffffffffu, word != 0u)) col = gh_transpose32(word, lane);
        pixbits[warp * 256 + t * 32 + lane] = col;                           // pixel 32t+lane: bits = Gaussians
    }
}

// ------------------------------------------------------------------------------------------ forward
#ifndef GH_FWD_MIN_CTAS
#define GH_FWD_MIN_CTAS 4        // __launch_bounds__ second argument of the forward kernel (register cap)
#endif
__global__ void __launch_bounds__(256, GH_FWD_MIN_CTAS)
gh_blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_perm, uint64_t* inst,
                        const GhGeo* __restrict__ geo, const float* __restrict__ features,
                        int W, int H, int gx, const float* __restrict__ bg,
                        float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                        float* __restrict__ out)
{
    // Forward needs no cross-pixel communication, so every lane walks the list of ITS OWN pixel:
    // a lane only ever touches Gaussians whose alpha >= 1/255 footprint (conservatively) contains
    // its pixel, whatever the other lanes of the warp are doing.
    extern __shared__ __align__(16) unsigned char gh_fwd_smem[];
    GhStage& st = *reinterpret_cast<GhStage*>(gh_fwd_smem);
    uint32_t* pixbits = reinterpret_cast<uint32_t*>(gh_fwd_smem + sizeof(GhStage));   // [word][pixel] (GH_CHUNK / 32 words x 256); also the sort's scratch

    const int tile = GH_TILE_ORDER ? (int)tile_perm[blockIdx.x] : (int)blockIdx.x;      // heaviest tiles first
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a warp owns two pixel rows: lanes 0..15 row 2*warp, lanes 16..31 row 2*warp+1
    const int px = tx * GH_BLOCK_X + (lane & 15);
    const int py = ty * GH_BLOCK_Y + 2 * warp + (lane >> 4);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;
    const float tx0 = (float)(tx * GH_BLOCK_X), ty0 = (float)(ty * GH_BLOCK_Y);

    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    const int nchunks = (n + GH_CHUNK - 1) / GH_CHUNK;

    float T = 1.0f;
    float2 C2[GH_HALF_C];                      // channel pairs (2k, 2k+1)
#pragma unroll
    for (int k = 0; k < GH_HALF_C; k++) C2[k] = make_float2(0.f, 0.f);
    uint32_t last = 0;
    bool done = !inside;
    bool warp_done = (__ballot_sync(0xffffffffu, done) == 0xffffffffu);

    // Sort this tile's bucket by (depth bits, Gaussian index) right here when it fits the staging buffers
    // (<= 2048 records = 16 KB): the sort is latency/barrier bound, the blend is issue bound, and as
    // one kernel the two overlap across the CTAs of an SM.  The sorted bucket is written back for the
    // backward pass.  Longer lists were sorted by gh_tile_split_long_kernel + gh_segment_sort_kernel
    // (gh_binning.cu) before this launch.
#ifndef GH_ABLATE_SORT     // (timing experiments only: tools/bench_variants.py)
#define GH_ABLATE_SORT 0
#endif
#ifndef GH_ABLATE_TRAVERSE
#define GH_ABLATE_TRAVERSE 0
#endif
#ifndef GH_ABLATE_BUILD
#define GH_ABLATE_BUILD 0
#endif
#ifndef GH_ABLATE_GATHER
#define GH_ABLATE_GATHER 0
#endif
    if (!GH_ABLATE_SORT && n >= 2 && n <= (int)GH_INKERNEL_SORT_MAX) {
        uint64_t* sbase = reinterpret_cast<uint64_t*>(&st.g0[0][0]);     // g0 + g1 = 16 KB contiguous
        uint64_t* spong = reinterpret_cast<uint64_t*>(&st.feat[0][0]);   // next 16 KB (feat is 20 KB)
        uint64_t* gl = inst + rg.x;
        // The bucket is one contiguous run of 8-byte records: the TMA engine copies it into shared memory
        // (cp.async.bulk, one request, completion on an mbarrier) instead of a load/store loop.  Bulk copies
        // move 16-byte granules, so the run is widened to even record indices; `skeys` skips the extra
        // leading record.  The rare run that would not fit after widening takes the loop.
        const uint32_t odd = rg.x & 1u;
        const uint32_t nrec = ((uint32_t)n + odd + 1u) & ~1u;
        uint64_t* skeys = sbase;                  // the loop path below uses the whole 2048-record buffer
        if (nrec <= GH_INKERNEL_SORT_MAX) {
            skeys = sbase + odd;
            __shared__ __align__(8) uint64_t s_mbar;
            const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) {
                const uint32_t bytes = nrec * 8u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"((uint32_t)__cvta_generic_to_shared(sbase)), "l"(gl - odd), "r"(bytes), "r"(mbar) : "memory");
            }
            uint32_t landed = 0;
            for (uint32_t it = 0; it < (1u << 24) && !landed; it++)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(landed) : "r"(mbar) : "memory");
            if (!landed) __trap();            // the copy engine never reported completion
        } else {
            for (int i = tid; i < n; i += 256) skeys[i] = gl[i];
            __syncthreads();
        }
        if (n <= 64) gh_bitonic_sort(skeys, (uint32_t)n, tid, 256);
        else gh_bucket_sort_tile<256>(skeys, spong, pixbits, n, tid);
        for (int i = tid; i < n; i += 256) gl[i] = skeys[i];
        __syncthreads();
    }

#if GH_FWD_NBUF == 2
    // prologue: chunk 0 in flight, indices of chunk 1 in a register
    uint32_t next_id = 0;
    if (nchunks > 0) {
        if (!GH_ABLATE_GATHER && tid < min(GH_CHUNK, n)) gh_stage_issue(st, 0, tid, (uint32_t)inst[(size_t)rg.x + tid], geo, features);
        gh_cp_async_commit();
        if (GH_CHUNK + tid < n) next_id = (uint32_t)inst[(size_t)rg.x + GH_CHUNK + tid];
    }
#else
    constexpr int PER_THREAD = GH_CHUNK / 256;
    uint32_t next_id[PER_THREAD];
#pragma unroll
    for (int h = 0; h < PER_THREAD; h++) next_id[h] = (tid + h * 256 < n) ? (uint32_t)inst[(size_t)rg.x + tid + h * 256] : 0u;
#endif

    for (int c = 0; c < nchunks; c++) {
        const int base = c * GH_CHUNK;
        const int cnt = min(GH_CHUNK, n - base);
#if GH_FWD_NBUF == 2
        const int buf = c & 1;
        gh_cp_async_wait_all();
        // chunk c landed for every thread; block-wide early exit like the reference's
        // __syncthreads_count(done) == BLOCK_SIZE
        if (__syncthreads_and(warp_done)) break;
        if (c + 1 < nchunks) {
            if (!GH_ABLATE_GATHER && base + GH_CHUNK + tid < n) gh_stage_issue(st, buf ^ 1, tid, next_id, geo, features);
            gh_cp_async_commit();
            if (base + 2 * GH_CHUNK + tid < n) next_id = (uint32_t)inst[(size_t)rg.x + base + 2 * GH_CHUNK + tid];
        }
        if (!GH_ABLATE_BUILD) gh_build_pixel_lists(st, pixbits, buf, cnt, warp, lane, tx0, ty0);
        __syncthreads();
#else
        constexpr int buf = 0;
        // everyone is done with the previous window; block-wide early exit like the reference's
        // __syncthreads_count(done) == BLOCK_SIZE
        if (__syncthreads_and(warp_done)) break;
#pragma unroll
        for (int h = 0; h < PER_THREAD; h++)
            if (tid + h * 256 < cnt) gh_stage_issue(st, 0, tid + h * 256, next_id[h], geo, features);
        gh_cp_async_commit();
#pragma unroll
        for (int h = 0; h < PER_THREAD; h++) {
            const int nx = base + GH_CHUNK + tid + h * 256;
            next_id[h] = (nx < n) ? (uint32_t)inst[(size_t)rg.x + nx] : 0u;
        }
        gh_cp_async_wait_all();
        __syncthreads();
        for (int word = warp; word * 32 < cnt; word += 8) gh_build_pixel_lists(st, pixbits, 0, cnt, word, lane, tx0, ty0);
        __syncthreads();
#endif

        if (!done && !GH_ABLATE_TRAVERSE) {
            // which of my pixel's 8 list words are non-empty
            uint32_t nz = 0u;
#pragma unroll
            for (int k = 0; k < GH_CHUNK / 32; k++)
                if (GH_FWD_NBUF == 2 || k * 32 < cnt) nz |= (pixbits[k * 256 + tid] != 0u) ? (1u << k) : 0u;
            int wi = 0;
            uint32_t cur = 0u;
            while (true) {
                if (cur == 0u) {
                    if (nz == 0u) break;
                    wi = __ffs(nz) - 1;
                    nz &= nz - 1;
                    cur = pixbits[wi * 256 + tid];
                }
                const int bpos = __ffs(cur) - 1;
                cur &= cur - 1;
                const int jj = wi * 32 + bpos;
                const float4 g0 = st.g0[buf][jj], g1 = st.g1[buf][jj];
                const GhPixEval e = gh_eval(g0, g1, pxf, pyf);
                if (e.ok) {
                    const float test_T = GH_MUL(T, GH_SUB(1.0f, e.alpha));
                    if (test_T < 0.0001f) { done = true; break; }
                    // C[ch] = fma(T, alpha * f[ch], C[ch]) (forward.cu:379-380 as ptxas fuses it), two
                    // channels per instruction with Blackwell's packed FP32 pipe: FMUL2 + FFMA2 round each
                    // half exactly like FMUL + FFMA
                    const float2 a2 = make_float2(e.alpha, e.alpha), T2 = make_float2(T, T);
#pragma unroll
                    for (int k = 0; k < GH_HALF_C; k++)
                        C2[k] = __ffma2_rn(T2, __fmul2_rn(a2, st.feat[buf][jj * GH_HALF_C + k]), C2[k]);
                    T = test_T;
                    last = (uint32_t)(base + jj + 1);
                }
            }
        }
        warp_done = (__ballot_sync(0xffffffffu, done) == 0xffffffffu);
    }
    gh_cp_async_wait_all();

    if (inside) {
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int ch = 0; ch < GH_NUM_CHANNELS; ch++)
            out[ch * plane + pix] = GH_FMA(T, __ldg(bg + ch), (ch & 1) ? C2[ch >> 1].y : C2[ch >> 1].x);
    }
}

// ------------------------------------------------------------------------------------------ backward
// Backward staging: ONE window of GH_BWD_CHUNK instances (single-buffered; the other resident CTAs of
// the SM cover the gather latency).  The 8 blocks of a warp walk their lists in lock-step, so a warp
// spends max(list length over its blocks) steps per window: a 512-wide window wastes fewer lanes on
// that maximum than two 256-wide ones (tools/imbalance.py: 0.74 vs 0.69 lane efficiency at 500k).
struct GhStageB {
    float4 g0[GH_BWD_CHUNK];
    float4 g1[GH_BWD_CHUNK];
    float2 feat[GH_BWD_CHUNK * GH_HALF_C];
    uint32_t id[GH_BWD_CHUNK];
    uint32_t bits[GH_BWD_NBLK][GH_BWD_CHUNK / 32 + 1];   // [block][word]; +1 pad against bank conflicts
};

__device__ __forceinline__ void gh_stage_issue_b(GhStageB& st, int slot, uint32_t id,
                                                 const GhGeo* __restrict__ geo, const float* __restrict__ features) {
    const float4* gp = reinterpret_cast<const float4*>(geo + id);
    gh_cp_async16(&st.g0[slot], gp);
    gh_cp_async16(&st.g1[slot], gp + 1);
    const float2* fp = reinterpret_cast<const float2*>(features + (size_t)id * GH_NUM_CHANNELS);
#pragma unroll
    for (int k = 0; k < GH_HALF_C; k++) gh_cp_async8(&st.feat[slot * GH_HALF_C + k], fp + k);
    st.id[slot] = id;
}

// Which of the tile's 32 blocks (4 wide x 2 high; bit = by*4 + bx) can this Gaussian reach with
// alpha >= 1/255?  Exact-conservative: on each pixel row the set {q(d) <= 2(thr+slack)} is an
// x-interval obtained from the quadratic; the interval is widened by 0.01 px and rounded outwards to
// whole pixels.  Returns all ones when the conic is not positive definite (no bound available).
// Skipping a Gaussian in a block whose bit is clear cannot change any pixel of that block: each of
// them would take the reference's `alpha < 1/255 -> continue` (forward.cu:370, backward.cu:503).
// Two pixel rows per instruction; a row without a real span has disc <= 0 -> s = NaN -> fminf / fmaxf
// ignore it.
#if GH_BWD_LPB == 4      // (the 4x2-block variant of the measured A/B; the product builds 2x2)
__device__ __forceinline__ uint32_t gh_block_mask2(const float4 g0, const float4 g1, float tx0, float ty0) {
    const float gx = g0.x, gy = g0.y, a = g0.z, b = g0.w, c = g1.x, thr = g1.z, pd = g1.w;
    if (pd == 0.f) return 0xffffffffu;
    const float mxd = fmaxf(fabsf(gx - tx0), fabsf(gx - (tx0 + 15.f)));
    const float myd = fmaxf(fabsf(gy - ty0), fabsf(gy - (ty0 + 15.f)));
    const float slack = 1e-3f + 2e-5f * (a * mxd * mxd + c * myd * myd + 2.f * fabsf(b) * mxd * myd);
    const float Q = 2.f * (thr + slack);
    if (!(Q >= 0.f)) return (Q != Q) ? 0xffffffffu : 0u;
    const float2 ia = gh_f2(__frcp_rn(a)), aQ = gh_f2(a * Q), bbac = gh_f2(b * b - a * c), gx2 = gh_f2(gx), b2 = gh_f2(b);
    float2 dy = make_float2(gy - ty0, gy - (ty0 + 1.f));     // rows 2 by, 2 by + 1
    uint32_t mask = 0;
#pragma unroll
    for (int by = 0; by < 8; by++) {
        // a dx^2 + 2 b dy dx + (c dy^2 - Q) <= 0,  dx = gx - px
        const float2 disc = gh_fma2(gh_mul2(dy, dy), bbac, aQ);
        const float2 s = gh_mul2(disc, make_float2(rsqrtf(disc.x), rsqrtf(disc.y)));   // NaN unless disc > 0
        const float2 hb = gh_mul2(b2, dy);
        const float2 lo2 = gh_fma2(gh_sub2(hb, s), ia, gx2), hi2 = gh_fma2(gh_add2(hb, s), ia, gx2);
        const float xlo = fminf(fminf(1e30f, lo2.x), lo2.y), xhi = fmaxf(fmaxf(-1e30f, hi2.x), hi2.y);
        // tile-local integer pixel columns inside [xlo - .01, xhi + .01]
        const float lo = (xlo - 0.01f) - tx0, hi = (xhi + 0.01f) - tx0;
        const int p0 = max(0, __float2int_ru(lo)), p1 = min(15, __float2int_rd(hi));
        if (p0 <= p1) {
            const int b0 = p0 >> 2, b1 = p1 >> 2;
            mask |= (((2u << (b1 - b0)) - 1u) << b0) << (4 * by);
        }
        dy = gh_add2(dy, gh_f2(-2.f));
    }
    return mask;
}
#endif

// The same test at 2x2-block granularity: 64 blocks, bit = by * 8 + bx; .x = block rows 0..3, .y = rows 4..7.
#if GH_BWD_LPB != 4
__device__ __forceinline__ uint2 gh_block_mask_2x2(const float4 g0, const float4 g1, float tx0, float ty0) {
    const float gx = g0.x, gy = g0.y, a = g0.z, b = g0.w, c = g1.x, thr = g1.z, pd = g1.w;
    if (pd == 0.f) return make_uint2(0xffffffffu, 0xffffffffu);
    const float mxd = fmaxf(fabsf(gx - tx0), fabsf(gx - (tx0 + 15.f)));
    const float myd = fmaxf(fabsf(gy - ty0), fabsf(gy - (ty0 + 15.f)));
    const float slack = 1e-3f + 2e-5f * (a * mxd * mxd + c * myd * myd + 2.f * fabsf(b) * mxd * myd);
    const float Q = 2.f * (thr + slack);
    if (!(Q >= 0.f)) return (Q != Q) ? make_uint2(0xffffffffu, 0xffffffffu) : make_uint2(0u, 0u);
    const float2 ia = gh_f2(__frcp_rn(a)), aQ = gh_f2(a * Q), bbac = gh_f2(b * b - a * c), gx2 = gh_f2(gx), b2 = gh_f2(b);
    float2 dy = make_float2(gy - ty0, gy - (ty0 + 1.f));     // rows 2 by, 2 by + 1
    uint32_t m[2] = {0u, 0u};
#pragma unroll
    for (int by = 0; by < 8; by++) {
        const float2 disc = gh_fma2(gh_mul2(dy, dy), bbac, aQ);
        const float2 s = gh_mul2(disc, make_float2(rsqrtf(disc.x), rsqrtf(disc.y)));   // NaN unless disc > 0
        const float2 hb = gh_mul2(b2, dy);
        const float2 lo2 = gh_fma2(gh_sub2(hb, s), ia, gx2), hi2 = gh_fma2(gh_add2(hb, s), ia, gx2);
        const float xlo = fminf(fminf(1e30f, lo2.x), lo2.y), xhi = fmaxf(fmaxf(-1e30f, hi2.x), hi2.y);
        const float lo = (xlo - 0.01f) - tx0, hi = (xhi + 0.01f) - tx0;
        const int p0 = max(0, __float2int_ru(lo)), p1 = min(15, __float2int_rd(hi));
        if (p0 <= p1) {
            const int b0 = p0 >> 1, b1 = p1 >> 1;
            m[by >> 2] |= (((2u << (b1 - b0)) - 1u) << b0) << (8 * (by & 3));
        }
        dy = gh_add2(dy, gh_f2(-2.f));
    }
    return make_uint2(m[0], m[1]);
}
#endif

// batch `word` = instances [32 word, 32 word + 32) of the window, one per lane; a 5-step shuffle
// transpose turns the 32 block masks into one list word per block (lane = block).  The word is cut at
// the block's deepest blended list position `glast` here, once, rather than in the traversal loop.
__device__ __forceinline__ uint32_t gh_valid_bits(uint32_t glast, int first_pos) {
    const int lim = (int)glast - first_pos;
    return lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
}

// glast_lo / glast_hi: deepest blended list position of block `lane` (and, with 64 blocks, of block 32 + lane)
__device__ __forceinline__ void gh_build_lists_b(GhStageB& st, int cnt, int word, int lane, float tx0, float ty0,
                                                 int base, uint32_t glast_lo, uint32_t glast_hi) {
    const int j = word * 32 + lane;
#if GH_BWD_LPB == 4
    uint32_t m = 0;
    if (j < cnt) m = gh_block_mask2(st.g0[j], st.g1[j], tx0, ty0);
    uint32_t mine = 0;
    if (__any_sync(0xffffffffu, m != 0u)) mine = gh_transpose32(m, lane);
    st.bits[lane][word] = mine & gh_valid_bits(glast_lo, base + word * 32);
    (void)glast_hi;
#else
    uint2 m = make_uint2(0u, 0u);
    if (j < cnt) m = gh_block_mask_2x2(st.g0[j], st.g1[j], tx0, ty0);
    uint32_t lo = 0, hi = 0;
    if (__any_sync(0xffffffffu, m.x != 0u)) lo = gh_transpose32(m.x, lane);
    if (__any_sync(0xffffffffu, m.y != 0u)) hi = gh_transpose32(m.y, lane);
    st.bits[lane][word] = lo & gh_valid_bits(glast_lo, base + word * 32);
    st.bits[32 + lane][word] = hi & gh_valid_bits(glast_hi, base + word * 32);
#endif
}

// CTA = tile = 4 warps.  A lane owns a vertical pixel pair (x, y0), (x, y0+1); 4 lanes own a 4x2 block;
// the 8 blocks of a warp walk their own lists in lock-step, i.e. every warp instruction works on EIGHT
// different Gaussians.  The two pixels of a lane are summed in registers; the 16 gradient components
// are then summed over the 4 lanes through shared memory (gh_group4_reduce16), leaving four adjacent
// components per lane = one REDG.E.ADD.F32x4 per lane.
#define GH_BWD_THREADS 128

// Sum the 16 gradient components over the 4 lanes of a block and leave components 4*(lane&3)+{0..3}
// in each lane.  Through shared memory: 4 STS.128 + 4 LDS.128 + 6 FADD2 per lane, against 12 SHFL +
// 24 SEL + 6 FADD2 for a transposing butterfly (the kernel is bound by issue slots).  A lane's record
// is 80 B (16 floats + pad): the 8 lanes of a quarter-warp store to, and load from, 8 disjoint groups
// of 4 banks.
#define GH_RED_STRIDE4 5       // float4 per lane record
#if GH_BWD_LPB == 4
__device__ __forceinline__ float4 gh_group4_reduce16(const float (&v)[16], float4* warp_buf, int lane) {
    float4* mine = warp_buf + lane * GH_RED_STRIDE4;
    mine[0] = make_float4(v[0], v[1], v[2], v[3]);
    mine[1] = make_float4(v[4], v[5], v[6], v[7]);
    mine[2] = make_float4(v[8], v[9], v[10], v[11]);
    mine[3] = make_float4(v[12], v[13], v[14], v[15]);
    __syncwarp();
    const float4* grp = warp_buf + (lane & 28) * GH_RED_STRIDE4 + (lane & 3);
    const float4 a = grp[0], b = grp[GH_RED_STRIDE4], c = grp[2 * GH_RED_STRIDE4], d = grp[3 * GH_RED_STRIDE4];
    __syncwarp();              // the next step overwrites the records
    const float2 lo = gh_add2(gh_add2(make_float2(a.x, a.y), make_float2(b.x, b.y)),
                              gh_add2(make_float2(c.x, c.y), make_float2(d.x, d.y)));
    const float2 hi = gh_add2(gh_add2(make_float2(a.z, a.w), make_float2(b.z, b.w)),
                              gh_add2(make_float2(c.z, c.w), make_float2(d.z, d.w)));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
#endif

// 2-lane blocks: the even lane keeps components 0..7, the odd lane 8..15 (two REDG.E.ADD.F32x4 each).  Each lane
// stores the half its partner keeps (2 STS.128), loads the partner's contribution to its own half (2 LDS.128) and
// adds (4 FADD2).  Record stride 48 B: the 8 lanes of a quarter-warp hit 8 disjoint groups of 4 banks.
#define GH_RED2_STRIDE4 3
#if GH_BWD_LPB != 4
__device__ __forceinline__ void gh_group2_reduce16(const float (&v)[16], float4* warp_buf, int lane, float4& out0, float4& out1) {
    const bool odd = (lane & 1) != 0;
    float4* mine = warp_buf + lane * GH_RED2_STRIDE4;
    // what the partner keeps: the even lane's partner (odd) keeps 8..15, the odd lane's partner keeps 0..7
    mine[0] = odd ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(v[8], v[9], v[10], v[11]);
    mine[1] = odd ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(v[12], v[13], v[14], v[15]);
    __syncwarp();
    const float4* theirs = warp_buf + (lane ^ 1) * GH_RED2_STRIDE4;
    const float4 a = theirs[0], b = theirs[1];
    __syncwarp();              // the next step overwrites the records
    const float4 k0 = odd ? make_float4(v[8], v[9], v[10], v[11]) : make_float4(v[0], v[1], v[2], v[3]);
    const float4 k1 = odd ? make_float4(v[12], v[13], v[14], v[15]) : make_float4(v[4], v[5], v[6], v[7]);
    const float2 s0 = gh_add2(make_float2(k0.x, k0.y), make_float2(a.x, a.y)), s1 = gh_add2(make_float2(k0.z, k0.w), make_float2(a.z, a.w));
    const float2 s2 = gh_add2(make_float2(k1.x, k1.y), make_float2(b.x, b.y)), s3 = gh_add2(make_float2(k1.z, k1.w), make_float2(b.z, b.w));
    out0 = make_float4(s0.x, s0.y, s1.x, s1.y);
    out1 = make_float4(s2.x, s2.y, s3.x, s3.y);
}
#endif

// State of the lane's pixel pair; .x = pixel (x, y0), .y = pixel (x, y0 + 1).
struct GhBwdPair {
    float2 T, A, last_alpha, last_cdot;
    float2 ntf_bg;                 // -T_final * (bg . dL_dpix)
    float2 npy;                    // -(pixel y)
    uint32_t last0, last1;
    float2 dL0[GH_HALF_C], dL1[GH_HALF_C];   // dL/dpixel of pixel 0 / pixel 1, channel pairs (2k, 2k+1)
};

// The pixel pair's share of one Gaussian, branch-free: every lane runs the same instruction stream and
// a pixel that does not blend this Gaussian (behind its last contributor, power > 0, alpha < 1/255 --
// the reference's three `continue`s, backward.cu:490-505) has alpha = G = 0, contributes exact zeros
// and keeps its state.  power / alpha are bit-identical to the forward pass (same operation order,
// FMUL2/FFMA2 round like FMUL/FFMA), so both passes agree on which pixels blend.  v = the 16 gradient
// components summed over the two pixels.
__device__ __forceinline__ bool gh_bwd_pair(GhBwdPair& p, const float4 g0, const float4 g1, const float2* feat,
                                            float pxf, uint32_t pos, float ddelx_dx, float ddely_dy, float (&v)[16]) {
    const float dx = GH_SUB(g0.x, pxf);
    const float2 dx2 = gh_f2(dx);
    const float2 dy = gh_add2(gh_f2(g0.y), p.npy);
    // gh_power: fma(fma(dx, dx*ca, dy*(dy*cc)), -0.5, -(dy*(dx*cb)))
    const float t1 = GH_MUL(dx, g0.z), dxcb = GH_MUL(dx, g0.w);
    const float2 t3 = gh_mul2(dy, gh_mul2(dy, gh_f2(g1.x)));
    const float2 sq = gh_fma2(dx2, gh_f2(t1), t3);
    const float2 power = gh_fma2(sq, gh_f2(-0.5f), gh_mul2(dy, gh_f2(-dxcb)));
#if GH_BWD_FAST_EXP
    const float Gx0 = __expf((power.x > 0.0f) ? 0.0f : power.x);
    const float Gx1 = __expf((power.y > 0.0f) ? 0.0f : power.y);
#else
    const float Gx0 = expf((power.x > 0.0f) ? 0.0f : power.x);   // expf(power) whenever the reference evaluates it
    const float Gx1 = expf((power.y > 0.0f) ? 0.0f : power.y);
#endif
    const float2 ao = gh_mul2(gh_f2(g1.y), make_float2(Gx0, Gx1));
    const float alpha0 = fminf(0.99f, ao.x), alpha1 = fminf(0.99f, ao.y);
    const bool ok0 = (pos < p.last0) & !(power.x > 0.0f) & !(alpha0 < 1.0f / 255.0f);
    const bool ok1 = (pos < p.last1) & !(power.y > 0.0f) & !(alpha1 < 1.0f / 255.0f);
    const float2 am = make_float2(ok0 ? alpha0 : 0.f, ok1 ? alpha1 : 0.f);
    const float2 G = make_float2(ok0 ? Gx0 : 0.f, ok1 ? Gx1 : 0.f);
    // T_{before this Gaussian} = T / (1 - alpha); gradients only need ~1 ulp here: MUFU.RCP, 1 - alpha
    // is in [0.01, 1] and rcp(1) == 1 exactly
    const float2 om = gh_sub2(gh_f2(1.0f), am);
    const float2 r = make_float2(gh_rcp_approx(om.x), gh_rcp_approx(om.y));
    p.T = gh_mul2(p.T, r);
    const float2 w = gh_mul2(am, p.T);                          // d(out)/d(color)
    const float2 w0 = gh_f2(w.x), w1 = gh_f2(w.y);
    float2 c0, c1;
#pragma unroll
    for (int k = 0; k < GH_HALF_C; k++) {
        const float2 f = feat[k];
        c0 = (k == 0) ? gh_mul2(f, p.dL0[0]) : gh_fma2(f, p.dL0[k], c0);
        c1 = (k == 0) ? gh_mul2(f, p.dL1[0]) : gh_fma2(f, p.dL1[k], c1);
        const float2 vc = gh_fma2(w1, p.dL1[k], gh_mul2(w0, p.dL0[k]));
        v[2 * k] = vc.x; v[2 * k + 1] = vc.y;
    }
    const float2 cdot = make_float2(c0.x + c0.y, c1.x + c1.y);
    // suffix recursion on the dot product (backward.cu:519-523), state advances only when blended
    const float2 A_new = gh_fma2(p.last_alpha, p.last_cdot, gh_mul2(gh_sub2(gh_f2(1.0f), p.last_alpha), p.A));
    // alpha also scales how much background shows through (backward.cu:535-538).  Not masked: every use
    // below is multiplied by G, which is 0 for a pixel that does not blend.
    const float2 dL_dalpha = gh_fma2(gh_sub2(cdot, A_new), p.T, gh_mul2(p.ntf_bg, r));
    // State update WITHOUT selects: a pixel that does not blend stores last_alpha = 0, so the next step computes
    // A_new = fma(0, last_cdot, 1 * A) = A exactly (cdot is finite) -- the recursion skips it just as the
    // reference's `continue` does, and A_new of this step is what the next blended step would have computed.
    p.A = A_new;
    p.last_cdot = cdot;
    p.last_alpha = am;
    const float2 dL_dG = gh_mul2(gh_f2(g1.y), dL_dalpha);
    const float2 gdx = gh_mul2(G, dx2), gdy = gh_mul2(G, dy);
    const float2 ncb = gh_f2(-g0.w);
    const float2 dG_ddelx = gh_fma2(gdx, gh_f2(-g0.z), gh_mul2(gdy, ncb));
    const float2 dG_ddely = gh_fma2(gdy, gh_f2(-g1.x), gh_mul2(gdx, ncb));
    const float2 t10 = gh_mul2(dL_dG, dG_ddelx), t11 = gh_mul2(dL_dG, dG_ddely);
    const float2 hg = gh_mul2(gh_f2(-0.5f), dL_dG);
    const float2 hgdx = gh_mul2(hg, gdx);
    const float2 t12 = gh_mul2(hgdx, dx2), t13 = gh_mul2(hgdx, dy), t14 = gh_mul2(gh_mul2(hg, gdy), dy);
    const float2 t15 = gh_mul2(G, dL_dalpha);
    v[10] = (t10.x + t10.y) * ddelx_dx; v[11] = (t11.x + t11.y) * ddely_dy;
    v[12] = t12.x + t12.y; v[13] = t13.x + t13.y; v[14] = t14.x + t14.y; v[15] = t15.x + t15.y;
    return ok0 | ok1;
}

__global__ void __launch_bounds__(GH_BWD_THREADS, GH_BWD_MIN_CTAS)
gh_blend_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_perm, const uint64_t* __restrict__ inst,
                         const GhGeo* __restrict__ geo, const float* __restrict__ features,
                         int W, int H, int gx, const float* __restrict__ bg,
                         const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                         const float* __restrict__ dL_dpix,
                         float* __restrict__ acc16)   // [P][16]: colors 0..9, mean2D x,y, conic x,y,w, opacity
{
    extern __shared__ __align__(16) unsigned char gh_bwd_smem[];
    GhStageB& st = *reinterpret_cast<GhStageB*>(gh_bwd_smem);
    float4* red_buf = reinterpret_cast<float4*>(gh_bwd_smem + sizeof(GhStageB)) + (threadIdx.x >> 5) * 32 * GH_RED_STRIDE4;
    __shared__ uint32_t s_warp_last[GH_BWD_THREADS / 32];
    __shared__ uint32_t s_glast[GH_BWD_NBLK];              // per block: deepest list position any of its pixels blended
    __shared__ uint8_t s_perm[GH_BWD_THREADS / 32][GH_BWD_NBLK];   // per warp (redundant copies): rank -> block

    const int tile = GH_TILE_ORDER ? (int)tile_perm[blockIdx.x] : (int)blockIdx.x;      // heaviest tiles first
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float tx0 = (float)(tx * GH_BLOCK_X), ty0 = (float)(ty * GH_BLOCK_Y);
    const unsigned gshift = lane & (32 - GH_BWD_LPB);      // first lane of my block
    const unsigned gmask = (1u << GH_BWD_LPB) - 1u;
    const size_t plane = (size_t)H * W;
    const uint2 rg = ranges[tile];

    // ---- how far do the blocks / the tile reach into the list?  (natural block order here)
    {
        const int nb = GH_BWD_BPW * warp + lane / GH_BWD_LPB;
        const int nx = tx * GH_BLOCK_X + GH_BWD_BW * (nb % GH_BWD_NBX) + (lane % GH_BWD_LPB);
        const int ny = ty * GH_BLOCK_Y + 2 * (nb / GH_BWD_NBX);
        uint32_t gl = 0;
        if (nx < W && ny < H) gl = n_contrib[(size_t)ny * W + nx];
        if (nx < W && ny + 1 < H) gl = max(gl, n_contrib[(size_t)(ny + 1) * W + nx]);
#pragma unroll
        for (int o = GH_BWD_LPB / 2; o > 0; o >>= 1) gl = max(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        if ((lane % GH_BWD_LPB) == 0) s_glast[nb] = gl;
        uint32_t wl = gl;
#pragma unroll
        for (int o = GH_BWD_LPB; o < 32; o <<= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
        if (lane == 0) s_warp_last[warp] = wl;
    }
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int w = 0; w < GH_BWD_THREADS / 32; w++) tile_last = max(tile_last, s_warp_last[w]);
    const int n = (int)tile_last;             // nothing beyond is blended by any pixel of this tile
    const int nchunks = (n + GH_BWD_CHUNK - 1) / GH_BWD_CHUNK;
    if (nchunks == 0) return;
    constexpr int PER_THREAD = GH_BWD_CHUNK / GH_BWD_THREADS;
    // the builder's lane == block after the transpose (with 64 blocks: blocks `lane` and `32 + lane`)
    const uint32_t glast_lo = s_glast[lane];
    const uint32_t glast_hi = (GH_BWD_NBLK > 32) ? s_glast[(32 + lane) % GH_BWD_NBLK] : 0u;

    // ---- stage the last window (windows are visited last to first) and build its lists
    uint32_t next_id[PER_THREAD];
    {
        const int base = (nchunks - 1) * GH_BWD_CHUNK;
        const int cnt = n - base;
#pragma unroll
        for (int h = 0; h < PER_THREAD; h++) {
            const int slot = tid + h * GH_BWD_THREADS;
            if (slot < cnt) gh_stage_issue_b(st, slot, (uint32_t)inst[(size_t)rg.x + base + slot], geo, features);
        }
        gh_cp_async_commit();
        if (nchunks > 1) {
#pragma unroll
            for (int h = 0; h < PER_THREAD; h++)
                next_id[h] = (uint32_t)inst[(size_t)rg.x + base - GH_BWD_CHUNK + tid + h * GH_BWD_THREADS];
        }
        gh_cp_async_wait_all();
        __syncthreads();
        for (int word = warp; word * 32 < cnt; word += GH_BWD_THREADS / 32)
            gh_build_lists_b(st, cnt, word, lane, tx0, ty0, base, glast_lo, glast_hi);
        __syncthreads();
    }

    // ---- block -> (warp, lane group) assignment.  The blocks of a warp advance in lock-step, so a warp pays for
    // its LONGEST list: blocks are ranked by the length of their list in the window just built (lane b counts block
    // b -- and block 32 + b; every warp computes the same ranking for itself) and each warp takes GH_BWD_BPW
    // neighbours of that ranking.  Any assignment gives the same gradients up to the order of the float atomics.
    int blk;
    {
        const int nw = (n - (nchunks - 1) * GH_BWD_CHUNK + 31) >> 5;
        uint32_t len0 = 0, len1 = 0;
        for (int w = 0; w < nw; w++) {
            len0 += __popc(st.bits[lane][w]);
            if (GH_BWD_NBLK > 32) len1 += __popc(st.bits[(32 + lane) % GH_BWD_NBLK][w]);
        }
        const uint32_t key0 = (len0 << 6) | (uint32_t)lane, key1 = (len1 << 6) | (uint32_t)(32 + lane);
        int rank0 = 0, rank1 = 0;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const uint32_t a = __shfl_sync(0xffffffffu, key0, k);
            rank0 += (a < key0) ? 1 : 0;
            if (GH_BWD_NBLK > 32) {
                const uint32_t b = __shfl_sync(0xffffffffu, key1, k);
                rank0 += (b < key0) ? 1 : 0;
                rank1 += ((a < key1) ? 1 : 0) + ((b < key1) ? 1 : 0);
            }
        }
        s_perm[warp][rank0] = (uint8_t)lane;
        if (GH_BWD_NBLK > 32) s_perm[warp][rank1] = (uint8_t)(32 + lane);
        __syncwarp();
        blk = s_perm[warp][GH_BWD_BPW * warp + lane / GH_BWD_LPB];
    }
    const int px = tx * GH_BLOCK_X + GH_BWD_BW * (blk % GH_BWD_NBX) + (lane % GH_BWD_LPB);
    const int py0 = ty * GH_BLOCK_Y + 2 * (blk / GH_BWD_NBX);
    const float pxf = (float)px;

    GhBwdPair pr;
    {
        float tf[2], bgd[2];
        uint32_t lastv[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int py = py0 + r;
            const bool inside = (px < W) && (py < H);
            const size_t pi = (size_t)py * W + px;
            tf[r] = inside ? final_T[pi] : 0.f;
            lastv[r] = inside ? n_contrib[pi] : 0u;     // pixel blends list positions 1..last
            float d[GH_NUM_CHANNELS];
            bgd[r] = 0.f;
#pragma unroll
            for (int ch = 0; ch < GH_NUM_CHANNELS; ch++) {
                d[ch] = inside ? dL_dpix[ch * plane + pi] : 0.f;
                bgd[r] += __ldg(bg + ch) * d[ch];
            }
#pragma unroll
            for (int k = 0; k < GH_HALF_C; k++) {
                if (r == 0) pr.dL0[k] = make_float2(d[2 * k], d[2 * k + 1]);
                else pr.dL1[k] = make_float2(d[2 * k], d[2 * k + 1]);
            }
        }
        pr.T = make_float2(tf[0], tf[1]);
        pr.ntf_bg = make_float2(-tf[0] * bgd[0], -tf[1] * bgd[1]);
        pr.npy = make_float2(-(float)py0, -(float)(py0 + 1));
        pr.A = gh_f2(0.f); pr.last_alpha = gh_f2(0.f); pr.last_cdot = gh_f2(0.f);
        pr.last0 = lastv[0]; pr.last1 = lastv[1];
    }
    // pixel-coordinate -> NDC chain rule factors (backward.cu:464-465)
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    uint32_t wlast = s_glast[blk];
#pragma unroll
    for (int o = GH_BWD_LPB; o < 32; o <<= 1) wlast = max(wlast, __shfl_xor_sync(0xffffffffu, wlast, o));

    for (int c = nchunks - 1; c >= 0; c--) {
        const int base = c * GH_BWD_CHUNK;
        const int cnt = min(GH_BWD_CHUNK, n - base);
        if (c != nchunks - 1) {
            __syncthreads();                       // everyone is done with the previous window
#pragma unroll
            for (int h = 0; h < PER_THREAD; h++) gh_stage_issue_b(st, tid + h * GH_BWD_THREADS, next_id[h], geo, features);
            gh_cp_async_commit();
            if (c > 0) {
#pragma unroll
                for (int h = 0; h < PER_THREAD; h++)
                    next_id[h] = (uint32_t)inst[(size_t)rg.x + base - GH_BWD_CHUNK + tid + h * GH_BWD_THREADS];
            }
            gh_cp_async_wait_all();
            __syncthreads();
            // per-block lists: each of the 4 warps scan-converts every fourth batch of 32 Gaussians
            for (int word = warp; word * 32 < cnt; word += GH_BWD_THREADS / 32)
                gh_build_lists_b(st, cnt, word, lane, tx0, ty0, base, glast_lo, glast_hi);
            __syncthreads();
        }
        if ((uint32_t)base >= wlast) continue;

        int wi = (cnt + 31) >> 5;
        uint32_t cur = 0;
        while (true) {
            while (cur == 0 && wi > 0) cur = st.bits[blk][--wi];
            const bool act = (cur != 0);
            if (!__any_sync(0xffffffffu, act)) break;

            // blocks whose list is exhausted run the same code on slot 0 with pos = UINT_MAX (-> all zeros)
            const int bpos = act ? 31 - __clz(cur) : 0;   // back to front
            cur &= ~(1u << bpos);
            const int jj = act ? wi * 32 + bpos : 0;
            const uint32_t id = st.id[jj];
            const float4 g0 = st.g0[jj], g1 = st.g1[jj];
            const float2* feat = &st.feat[jj * GH_HALF_C];
            const uint32_t pos = act ? (uint32_t)(base + jj) : 0xffffffffu;
            float v[16];
            const bool contrib = gh_bwd_pair(pr, g0, g1, feat, pxf, pos, ddelx_dx, ddely_dy, v);
            const uint32_t cm = __ballot_sync(0xffffffffu, contrib);
            if (cm == 0u) continue;
#if GH_BWD_LPB == 4
            const float4 s = gh_group4_reduce16(v, red_buf, lane);
            if ((cm >> gshift) & gmask) {
                float4* dst = reinterpret_cast<float4*>(acc16 + (size_t)id * 16) + (lane & 3);
                atomicAdd(dst, s);    // REDG.E.ADD.F32x4
            }
#else
            float4 s0, s1;
            gh_group2_reduce16(v, red_buf, lane, s0, s1);
            if ((cm >> gshift) & gmask) {
                float4* dst = reinterpret_cast<float4*>(acc16 + (size_t)id * 16) + 2 * (lane & 1);
                atomicAdd(dst, s0);       // REDG.E.ADD.F32x4: components 0..3 / 8..11
                atomicAdd(dst + 1, s1);   //                              4..7 / 12..15
            }
#endif
        }
    }
    gh_cp_async_wait_all();
}

// ---------------------------------------------------------------------------- gradient unpack
// acc16 -> the reference binding's layouts (rasterize_points.cu:160-168): dL_dcolor (P,10),
// dL_dmean2D (P,3) [z never written by the reference: 0], dL_dconic (P,2,2) [.z unused: 0], dL_dopacity (P,1)
__global__ void __launch_bounds__(256)
gh_unpack_grads_kernel(int P, const float* __restrict__ acc16, float* __restrict__ dL_dmean2D,
                       float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                       float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D,
                       float* __restrict__ dL_dscale, float* __restrict__ dL_drot)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4* a = reinterpret_cast<const float4*>(acc16 + (size_t)idx * 16);
    const float4 a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
    float2* dc = reinterpret_cast<float2*>(dL_dcolor + (size_t)idx * GH_NUM_CHANNELS);
    dc[0] = make_float2(a0.x, a0.y); dc[1] = make_float2(a0.z, a0.w);
    dc[2] = make_float2(a1.x, a1.y); dc[3] = make_float2(a1.z, a1.w);
    dc[4] = make_float2(a2.x, a2.y);
    dL_dmean2D[3 * idx + 0] = a2.z; dL_dmean2D[3 * idx + 1] = a2.w; dL_dmean2D[3 * idx + 2] = 0.f;
    reinterpret_cast<float4*>(dL_dconic)[idx] = make_float4(a3.x, a3.y, 0.f, a3.z);
    dL_dopacity[idx] = a3.w;
    // conic supplied by the caller: nothing flows through the geometry (reference backward.cu:371,398,588)
    if (dL_dmean3D) { dL_dmean3D[3 * idx + 0] = 0.f; dL_dmean3D[3 * idx + 1] = 0.f; dL_dmean3D[3 * idx + 2] = 0.f; }
    if (dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * idx + k] = 0.f;
    }
    if (dL_dscale) { dL_dscale[3 * idx + 0] = 0.f; dL_dscale[3 * idx + 1] = 0.f; dL_dscale[3 * idx + 2] = 0.f; }
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

void gh_launch_blend_forward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                             const float* features, const float* bg, float* out_color,
                             cudaStream_t stream)
{
    // ask for the largest shared-memory carve-out so that 4 CTAs (4 x 48 KB) are resident per SM
    const int smem = (int)(sizeof(GhStage) + (GH_CHUNK / 32) * 256 * sizeof(uint32_t));
    cudaFuncSetAttribute(gh_blend_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gh_blend_forward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    gh_blend_forward_kernel<<<gx * gy, 256, smem, stream>>>(img.ranges, img.tile_perm, bin.inst, geom.geo, features,
                                                         W, H, gx, bg, img.final_T, img.n_contrib, out_color);
}

void gh_launch_blend_backward(int W, int H, int gx, int gy, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                              const float* features, const float* bg, const float* dL_dpix,
                              cudaStream_t stream)
{
    const int smem = (int)(sizeof(GhStageB) + GH_BWD_THREADS * GH_RED_STRIDE4 * sizeof(float4));
    cudaFuncSetAttribute(gh_blend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gh_blend_backward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    gh_blend_backward_kernel<<<gx * gy, GH_BWD_THREADS, smem, stream>>>(img.ranges, img.tile_perm, bin.inst, geom.geo, features,
                                                          W, H, gx, bg, img.final_T, img.n_contrib, dL_dpix,
                                                          geom.acc16);
}

void gh_launch_unpack_grads(int P, GhGeomWS geom, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                            float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dscale, float* dL_drot,
                            cudaStream_t stream)
{
    gh_unpack_grads_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, geom.acc16, dL_dmean2D, dL_dconic,
                                                                dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D,
                                                                dL_dscale, dL_drot);
}
