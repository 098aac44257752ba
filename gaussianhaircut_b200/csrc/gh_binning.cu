// Stage 2: binning.  Produces, per 16x16 tile, the list of Gaussian instances overlapping it,
// ordered by the reference's 64-bit key  (tile_id << 32) | float_bits(depth)  with ties broken by
// ascending Gaussian index -- exactly the order the reference obtains from duplicateWithKeys +
// a stable cub::DeviceRadixSort::SortPairs over bits [0, 32+bit) + identifyTileRanges
// (rasterizer_impl.cu:70-138, 293-321).
//
// New design (no multi-pass device-wide radix sort): the tile id is an exact bucket, so
//   1. stage 1 already histogrammed instances per tile;
//   2. one block scans the histogram -> tile ranges (= identifyTileRanges' output) and R;
//   3. each Gaussian scatters (depth_bits << 32 | idx) into its tiles' buckets (atomic cursor);
//   4. each bucket is sorted in shared memory on the composite 64-bit (depth, idx) key, which
//      reproduces the stable order because a Gaussian appears at most once per tile.
// HBM traffic is 8 B write + 8 B read + 8 B write per instance instead of ~156 B for 6 radix passes.
#include "gh_common.cuh"
#include "gh_kernels.h"

namespace {

// ---------------------------------------------------------------- tile scan (1 block)
__global__ void __launch_bounds__(1024)
gh_tile_scan_kernel(int T, const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_cursor,
                    uint2* __restrict__ ranges, GhCtrl* __restrict__ ctrl)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    __shared__ uint32_t max_s[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    uint32_t local_max = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int i = base + tid;
        const uint32_t c = (i < T) ? tile_count[i] : 0u;
        local_max = max(local_max, c);
        uint32_t v = c;   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_sums[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            warp_sums[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = carry + v + (wid > 0 ? warp_sums[wid - 1] : 0u);
        const uint32_t excl = incl - c;
        if (i < T) {
            tile_cursor[i] = excl;
            // empty tiles keep (0,0) like the reference's memset + identifyTileRanges
            ranges[i] = (c > 0) ? make_uint2(excl, incl) : make_uint2(0u, 0u);
        }
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) max_s[wid] = local_max;
    __syncthreads();
    if (tid == 0) {
        uint32_t m = 0;
        for (int w = 0; w < 32; w++) m = max(m, max_s[w]);
        ctrl->num_rendered = carry_s;
        ctrl->max_tile_len = m;
    }
}

// ---------------------------------------------------------------- emit
// Consecutive Gaussians of a strand fall into the same tiles, so the lanes of a warp mostly ask for
// slots of the same few buckets: lanes that want the same tile in the same loop step are grouped
// with match.any and their leader reserves all their slots with ONE atomic.
__global__ void __launch_bounds__(256)
gh_emit_kernel(int P, const int* __restrict__ radii, const GhGeo* __restrict__ geo,
               const float* __restrict__ depth, uint32_t* __restrict__ tile_cursor,
               uint64_t* __restrict__ inst, int gx, int gy)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int minx = 0, miny = 0, maxx = 0, maxy = 0;
    uint64_t rec = 0;
    if (idx < P) {
        const int r = radii[idx];
        if (r > 0) {
            const float4 g0 = reinterpret_cast<const float4*>(geo + idx)[0];
            gh_get_rect(g0.x, g0.y, r, gx, gy, minx, miny, maxx, maxy);
            rec = ((uint64_t)__float_as_uint(depth[idx]) << 32) | (uint32_t)idx;
        }
    }
    const int w = maxx - minx;
    const int count = w * (maxy - miny);
    int maxcount = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxcount = max(maxcount, __shfl_xor_sync(0xffffffffu, maxcount, o));
    int x = minx, y = miny;
    for (int t = 0; t < maxcount; t++) {
        const bool have = t < count;
        const int tile = have ? (y * gx + x) : -1;
        const uint32_t peers = __match_any_sync(0xffffffffu, tile);
        if (have) {
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&tile_cursor[tile], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            inst[base + __popc(peers & ((1u << lane) - 1u))] = rec;
            if (++x == maxx) { x = minx; y++; }
        }
    }
}

// ---------------------------------------------------------------- per-tile sort
// Normalised bitonic network (every comparator puts the smaller key at the lower index), so
// elements beyond n behave as +inf without being materialised: a comparator whose upper index
// is >= n is simply skipped.
__device__ __forceinline__ void gh_ce(uint64_t& a, uint64_t& b) {
    const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo; b = hi;
}

template <typename KeyPtr>
__device__ __forceinline__ void gh_bitonic_sort(KeyPtr keys, const uint32_t n, const int tid, const int nt)
{
    const uint64_t INF = ~0ull;             // virtual padding: never stored, never moves
    uint32_t ln2 = 0;                       // n2 = 1 << ln2 >= n
    while ((1u << ln2) < n) ln2++;
    const uint32_t half = (1u << ln2) >> 1, quarter = half >> 1;
    for (uint32_t lk = 1; lk <= ln2; lk++) {      // merge blocks of size k = 1 << lk
        {   // first step of the merge: partner = mirror inside the block of size k
            const uint32_t k = 1u << lk, hk = k >> 1;
            for (uint32_t t = tid; t < half; t += nt) {
                const uint32_t blk = t >> (lk - 1), off = t & (hk - 1);
                const uint32_t i = (blk << lk) + off, j = (blk << lk) + (k - 1 - off);
                if (j < n) {
                    const uint64_t a = keys[i], b = keys[j];
                    if (a > b) { keys[i] = b; keys[j] = a; }
                }
            }
            __syncthreads();
        }
        int ls = (int)lk - 2;                      // remaining strides s = k/4 ... 1
        // two strides (s, s/2) per pass: a thread owns i, i+s/2, i+s, i+3s/2 and does 4 exchanges in registers
        for (; ls >= 1; ls -= 2) {
            const uint32_t sft = (uint32_t)ls, s = 1u << sft, h = s >> 1;
            for (uint32_t t = tid; t < quarter; t += nt) {
                const uint32_t i = ((t >> (sft - 1)) << (sft + 1)) | (t & (h - 1));
                if (i < n) {
                    const uint32_t ib = i + h, ic = i + s, id = ic + h;
                    uint64_t ka = keys[i];
                    uint64_t kb = ib < n ? keys[ib] : INF;
                    uint64_t kc = ic < n ? keys[ic] : INF;
                    uint64_t kd = id < n ? keys[id] : INF;
                    gh_ce(ka, kc); gh_ce(kb, kd);
                    gh_ce(ka, kb); gh_ce(kc, kd);
                    keys[i] = ka;
                    if (ib < n) keys[ib] = kb;
                    if (ic < n) keys[ic] = kc;
                    if (id < n) keys[id] = kd;
                }
            }
            __syncthreads();
        }
        if (ls == 0) {                             // odd number of strides left: the stride-1 stage
            for (uint32_t t = tid; t < half; t += nt) {
                const uint32_t i = t << 1, j = i + 1;
                if (j < n) {
                    const uint64_t a = keys[i], b = keys[j];
                    if (a > b) { keys[i] = b; keys[j] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// SMEM_KEYS: capacity of the shared staging buffer; handles tiles with lo < n <= hi.
// Tiles longer than the buffer are sorted in place in global memory by the same network.
template <int NT>
__global__ void __launch_bounds__(NT)
gh_tile_sort_kernel(const uint2* __restrict__ ranges, uint64_t* inst,
                    uint32_t lo, uint32_t hi, uint32_t smem_keys)
{
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint2 rg = ranges[blockIdx.x];
    const uint32_t n = rg.y - rg.x;
    if (n <= lo || n > hi || n < 2) return;
    uint64_t* g = inst + rg.x;
    const int tid = threadIdx.x;
    if (n <= smem_keys) {
        for (uint32_t i = tid; i < n; i += NT) skeys[i] = g[i];
        __syncthreads();
        gh_bitonic_sort(skeys, n, tid, NT);
        for (uint32_t i = tid; i < n; i += NT) g[i] = skeys[i];
    } else {
        __syncthreads();
        gh_bitonic_sort(g, n, tid, NT);
    }
}

}  // namespace

void gh_launch_tile_scan(int T, GhImgWS img, cudaStream_t stream)
{
    gh_tile_scan_kernel<<<1, 1024, 0, stream>>>(T, img.tile_count, img.tile_cursor, img.ranges, img.ctrl);
}

void gh_launch_emit(int P, const int* radii, GhGeomWS geom, GhImgWS img, GhBinWS bin,
                    int gx, int gy, cudaStream_t stream)
{
    gh_emit_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, radii, geom.geo, geom.depth,
                                                        img.tile_cursor, bin.inst, gx, gy);
}

int gh_launch_tile_sort(int T, unsigned int max_tile_len, GhImgWS img, GhBinWS bin, cudaStream_t stream)
{
    if (max_tile_len < 2) return 0;
    constexpr uint32_t SMALL = 2048;     // 16 KB of keys, 256 threads
    constexpr uint32_t LARGE = 24576;    // 192 KB of keys, 1024 threads
    gh_tile_sort_kernel<256><<<T, 256, SMALL * 8, stream>>>(img.ranges, bin.inst, 0u, SMALL, SMALL);
    if (max_tile_len > SMALL) {
        cudaFuncSetAttribute(gh_tile_sort_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(LARGE * 8));
        gh_tile_sort_kernel<1024><<<T, 1024, LARGE * 8, stream>>>(img.ranges, bin.inst, SMALL, 0xffffffffu, LARGE);
        return 2;
    }
    return 1;
}
