// Gradient all-reduce over NVLink peer memory (views are sharded one per GPU and the
// per-Gaussian gradients are summed once per step).
//
// One kernel per rank on the step's stream, no NCCL on the data path.  Every rank's gradient arena lives
// in symmetric memory (same allocation on every GPU, all of them mapped into every process), so a rank
// can load from and store to its peers' arenas directly through NVSwitch:
//     entry barrier   "my gradients are complete" flags, release/acquire at system scope
//     two-shot sum    rank r owns slice r of the arena: it reads slice r of every peer (world-1 remote
//                     loads in flight per element), adds, and writes the sum into slice r of EVERY arena
//                     (or, when the allocation has an NVLS multicast mapping, one multimem.ld_reduce and
//                     one multimem.st per 16 bytes: the switch does the reduction and the broadcast)
//     exit barrier    "my stores have landed everywhere" flags; the kernel ends when all peers said so,
//                     so whatever follows on the stream (the optimizer) sees the reduced gradients
// Per GPU and direction the links carry 2 (world-1)/world of the arena (peers reading my copy + my stores
// to them), the same as a ring all-reduce, but in one launch with every pair of GPUs talking at once; the
// multimem path sends each reduced slice only once.  Measured on 48 MB (500k Gaussians), B200 NVLink 5,
// us per all-reduce  NCCL / peer ld-st / multimem:  N=2 114 / 86 / 146,  N=4 145 / 125 / 138,  N=8 226 / 161 / 148
// (tools/allreduce_case.py); at N=2 each direction carries the full 48 MB, i.e. ~650 GB/s sustained.  Spins are bounded: a peer that never arrives
// raises the error word instead of hanging the GPU.
#include "gh_common.cuh"
#include "gh_kernels.h"
#include "../../include/gh_rasterizer.h"

namespace {

#define GH_AR_THREADS 512
#define GH_AR_MAX_WORLD 16
#define GH_AR_SPIN_LIMIT (1u << 22)      // ~ seconds: a missing peer raises the error word, it never hangs the GPU

struct GhArPeers {
    float* buf[GH_AR_MAX_WORLD];
    unsigned int* flag[GH_AR_MAX_WORLD];     // per rank: uint[2 * world] = entry flags, exit flags
};

__device__ __forceinline__ void gh_st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int gh_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int gh_ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wait until *p >= epoch (epochs only grow; compare with wrap-around)
__device__ __forceinline__ bool gh_spin_sys(const unsigned int* p, unsigned int epoch) {
    for (unsigned int it = 0; it < GH_AR_SPIN_LIMIT; it++) {
        if ((int)(gh_ld_acquire_sys(p) - epoch) >= 0) return true;
        __nanosleep(20);
    }
    return false;
}

// local[0]: "go" word (entry barrier passed), local[1]: exit arrival counter, local[2]: error word
// WORLD_T > 0: world size known at compile time (2, 4, 8), U = 8 / WORLD_T elements per thread per round so
// that 8 remote 16-byte loads are in flight per thread (NVLink latency is a few microseconds).
template <int WORLD_T>
__global__ void __launch_bounds__(GH_AR_THREADS)
gh_allreduce_p2p_kernel(GhArPeers peers, float* __restrict__ mc, int rank, int world_rt, size_t off4, size_t n4,
                        unsigned int epoch, unsigned int* __restrict__ local)
{
    const int world = WORLD_T > 0 ? WORLD_T : world_rt;
    const int tid = threadIdx.x;
    // ---------------------------------------------------------------- entry barrier
    if (blockIdx.x == 0) {
        if (tid < world) {
            gh_st_release_sys(peers.flag[tid] + rank, epoch);                  // tell peer `tid`: my gradients are final
            if (!gh_spin_sys(peers.flag[rank] + tid, epoch)) atomicExch(local + 2, 1u);
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicExch(local + 0, epoch); }
    } else if (tid == 0) {
        for (unsigned int it = 0; it < GH_AR_SPIN_LIMIT; it++) {
            if ((int)(gh_ld_acquire_gpu(local + 0) - epoch) >= 0) break;
            __nanosleep(20);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- my slice: sum over ranks, write everywhere
    const size_t lo = off4 + n4 * (size_t)rank / (size_t)world, hi = off4 + n4 * (size_t)(rank + 1) / (size_t)world;
    const size_t stride = (size_t)gridDim.x * GH_AR_THREADS;
    if (mc != nullptr) {
        // NVLS: the switch reduces the `world` copies on the load and replicates the store; 4 requests in flight
        float4* m4 = reinterpret_cast<float4*>(mc);
        for (size_t i0 = lo + (size_t)blockIdx.x * GH_AR_THREADS + tid; i0 < hi; i0 += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t i = i0 + u * stride;
                if (i < hi)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(m4 + i) : "memory");
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t i = i0 + u * stride;
                if (i < hi)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                                 ::"l"(m4 + i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
            }
        }
    } else if (WORLD_T > 0) {
        constexpr int U = WORLD_T > 0 ? (8 / WORLD_T > 0 ? 8 / WORLD_T : 1) : 1;
        for (size_t i0 = lo + (size_t)blockIdx.x * GH_AR_THREADS + tid; i0 < hi; i0 += U * stride) {
            float4 part[U][WORLD_T > 0 ? WORLD_T : 1];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t i = i0 + u * stride;
#pragma unroll
                for (int p = 0; p < WORLD_T; p++)
                    part[u][p] = (i < hi) ? __ldcg(reinterpret_cast<const float4*>(peers.buf[p]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t i = i0 + u * stride;
                float4 acc = part[u][0];
#pragma unroll
                for (int p = 1; p < WORLD_T; p++) { acc.x += part[u][p].x; acc.y += part[u][p].y; acc.z += part[u][p].z; acc.w += part[u][p].w; }   // rank order: same bits everywhere
                if (i < hi) {
#pragma unroll
                    for (int p = 0; p < WORLD_T; p++) __stcg(reinterpret_cast<float4*>(peers.buf[p]) + i, acc);
                }
            }
        }
    } else {
        for (size_t i = lo + (size_t)blockIdx.x * GH_AR_THREADS + tid; i < hi; i += stride) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < world; p++) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(peers.buf[p]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            for (int p = 0; p < world; p++) __stcg(reinterpret_cast<float4*>(peers.buf[p]) + i, acc);
        }
    }

    // ---------------------------------------------------------------- exit barrier
    __threadfence_system();                       // my stores are visible system-wide before I report
    __syncthreads();
    __shared__ bool s_last;
    if (tid == 0) s_last = (atomicAdd(local + 1, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) { local[1] = 0u; __threadfence_system(); }
    __syncthreads();
    if (tid < world) {
        gh_st_release_sys(peers.flag[tid] + world + rank, epoch);
        if (!gh_spin_sys(peers.flag[rank] + world + tid, epoch)) atomicExch(local + 2, 1u);
    }
}

}  // namespace

extern "C" int gh_allreduce_p2p(const unsigned long long* peer_bufs, const unsigned long long* peer_flags,
                                unsigned long long multicast_buf, int rank, int world,
                                size_t offset_floats, size_t n_floats, unsigned int epoch,
                                unsigned int* local_sync, gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!peer_bufs || !peer_flags || !local_sync || world < 1 || world > GH_AR_MAX_WORLD || rank < 0 || rank >= world ||
        (offset_floats & 3) || (n_floats & 3) || epoch == 0)
        return GH_E_INVALID_ARG;
    if (n_floats == 0) return GH_OK;
    GhArPeers peers;
    for (int p = 0; p < GH_AR_MAX_WORLD; p++) {
        peers.buf[p] = p < world ? reinterpret_cast<float*>(peer_bufs[p]) : nullptr;
        peers.flag[p] = p < world ? reinterpret_cast<unsigned int*>(peer_flags[p]) : nullptr;
        if (p < world && (!peers.buf[p] || !peers.flag[p] || (peer_bufs[p] & 15))) return GH_E_INVALID_ARG;
    }
    // every CTA spins on the entry barrier, so all of them must be resident: one CTA per SM
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return GH_E_CUDA;
    const size_t n4 = n_floats / 4, per_rank = (n4 + world - 1) / world;
    size_t want = (per_rank + GH_AR_THREADS - 1) / GH_AR_THREADS;
    const int grid = (int)(want < 1 ? 1 : (want > (size_t)sms ? (size_t)sms : want));
    float* mc = reinterpret_cast<float*>(multicast_buf);
#define GH_AR_LAUNCH(WT) gh_allreduce_p2p_kernel<WT><<<grid, GH_AR_THREADS, 0, stream>>>(peers, mc, rank, world, offset_floats / 4, n4, epoch, local_sync)
    if (world == 2) GH_AR_LAUNCH(2);
    else if (world == 4) GH_AR_LAUNCH(4);
    else if (world == 8) GH_AR_LAUNCH(8);
    else GH_AR_LAUNCH(0);
#undef GH_AR_LAUNCH
    gh_count_launches(1);
    return cudaGetLastError() == cudaSuccess ? GH_OK : GH_E_CUDA;
}
