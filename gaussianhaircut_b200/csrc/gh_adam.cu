// "Next" row 2: the optimiser step that follows the rasterizer backward.
// One launch updates every parameter group with torch.optim.Adam's arithmetic
// (reference: torch.optim.Adam(l, lr=0.0, eps=1e-15), SRC/scene/gaussian_model.py:431-448) and the
// reference's NaN guard (SRC/train_gaussians.py:174-181: skip the step if any gradient has a NaN,
// done there with seven blocking `.isnan().any()` host syncs) is a device-side flag: no host sync.
#include "gh_common.cuh"
#include "gh_kernels.h"
#include "../../include/gh_rasterizer.h"

namespace {

struct GhAdamGroups {
    float* param[GH_ADAM_MAX_GROUPS];
    const float* grad[GH_ADAM_MAX_GROUPS];
    float* exp_avg[GH_ADAM_MAX_GROUPS];
    float* exp_avg_sq[GH_ADAM_MAX_GROUPS];
    unsigned long long end[GH_ADAM_MAX_GROUPS];   // exclusive prefix of element counts
    float lr[GH_ADAM_MAX_GROUPS];
    int n;
};

// MUFU-based square root / reciprocal (about 1 ulp each): the update is bandwidth bound only if the
// per-element arithmetic stays short.  The result differs from torch's IEEE sqrt + divide by a few
// ulp of the UPDATE, i.e. ~1e-7 * lr relative to the parameter.
__device__ __forceinline__ float gh_sqrt_approx(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float gh_rcp_approx(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct GhAdamConst { float beta1, beta2, omb1, omb2, eps, inv_bc2s, step_size; };

__device__ __forceinline__ void gh_adam_elem(float& p, float gr, float& m, float& v, const GhAdamConst& c) {
    m = fmaf(c.omb1, gr - m, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(v, c.beta2, (c.omb2 * gr) * gr);              // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = fmaf(gh_sqrt_approx(v), c.inv_bc2s, c.eps);   // sqrt(v) / sqrt(bias_correction2) + eps
    p = fmaf(-(c.step_size * m), gh_rcp_approx(denom), p);            // addcdiv_(exp_avg, denom, -lr / bias_correction1)
}

// grid = (blocks, groups): blockIdx.y is the parameter group, float4 per thread when the group's four
// arrays are 16-byte aligned (torch allocations are), scalar tail / fallback otherwise.
__global__ void __launch_bounds__(256)
gh_adam_nan_kernel(GhAdamGroups g, unsigned int* __restrict__ flag)
{
    const int k = blockIdx.y;
    const unsigned long long n = g.end[k] - (k ? g.end[k - 1] : 0ull);
    const float* __restrict__ gr = g.grad[k];
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    bool bad = false;
    unsigned long long done = 0;
    if ((reinterpret_cast<size_t>(gr) & 15) == 0) {
        const unsigned long long n4 = n >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(gr);
        for (unsigned long long i = tid; i < n4; i += nthreads) {
            const float4 x = g4[i];
            bad |= (x.x != x.x) | (x.y != x.y) | (x.z != x.z) | (x.w != x.w);
        }
        done = n4 << 2;
    }
    for (unsigned long long i = done + tid; i < n; i += nthreads) { const float x = gr[i]; bad |= (x != x); }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

__global__ void __launch_bounds__(256)
gh_adam_update_kernel(GhAdamGroups g, float beta1, float beta2, float eps,
                      float bc1, float bc2_sqrt, const unsigned int* __restrict__ flag,
                      const unsigned int* __restrict__ skip_flag, int* step_state)
{
    if (flag != nullptr && *flag != 0u) return;     // a gradient held a NaN: skip this step entirely
    if (skip_flag != nullptr && *skip_flag != 0u) return;   // the producer of the gradients reported a failure
    if (step_state != nullptr) {
        // device-resident step count (only advanced by steps that were not skipped, like torch's state['step'])
        const int step = step_state[0] + 1;
        bc1 = 1.0f - powf(beta1, (float)step);
        bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    }
    const int k = blockIdx.y;
    const unsigned long long n = g.end[k] - (k ? g.end[k - 1] : 0ull);
    float* __restrict__ P = g.param[k];
    const float* __restrict__ G = g.grad[k];
    float* __restrict__ M = g.exp_avg[k];
    float* __restrict__ V = g.exp_avg_sq[k];
    GhAdamConst c;
    c.beta1 = beta1; c.beta2 = beta2; c.omb1 = 1.0f - beta1; c.omb2 = 1.0f - beta2; c.eps = eps;
    c.inv_bc2s = 1.0f / bc2_sqrt; c.step_size = g.lr[k] / bc1;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long done = 0;
    if (((reinterpret_cast<size_t>(P) | reinterpret_cast<size_t>(G) | reinterpret_cast<size_t>(M) | reinterpret_cast<size_t>(V)) & 15) == 0) {
        const unsigned long long n4 = n >> 2;
        for (unsigned long long i = tid; i < n4; i += nthreads) {
            float4 p = reinterpret_cast<float4*>(P)[i];
            const float4 gr = reinterpret_cast<const float4*>(G)[i];
            float4 m = reinterpret_cast<float4*>(M)[i], v = reinterpret_cast<float4*>(V)[i];
            gh_adam_elem(p.x, gr.x, m.x, v.x, c); gh_adam_elem(p.y, gr.y, m.y, v.y, c);
            gh_adam_elem(p.z, gr.z, m.z, v.z, c); gh_adam_elem(p.w, gr.w, m.w, v.w, c);
            reinterpret_cast<float4*>(P)[i] = p;
            reinterpret_cast<float4*>(M)[i] = m;
            reinterpret_cast<float4*>(V)[i] = v;
        }
        done = n4 << 2;
    }
    for (unsigned long long i = done + tid; i < n; i += nthreads) {
        float p = P[i], m = M[i], v = V[i];
        gh_adam_elem(p, G[i], m, v, c);
        P[i] = p; M[i] = m; V[i] = v;
    }
    if (step_state != nullptr) {
        // the last CTA to finish advances the counter (every CTA has read it by then)
        __threadfence();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(reinterpret_cast<unsigned int*>(step_state + 1), 1u);
            if (t == gridDim.x * gridDim.y - 1) { step_state[0] += 1; step_state[1] = 0; }
        }
    }
}

}  // namespace

extern "C" int gh_adam_step(int n_groups, float* const* params, const float* const* grads,
                            float* const* exp_avg, float* const* exp_avg_sq,
                            const unsigned long long* sizes, const float* lrs,
                            float beta1, float beta2, float eps, int step, int* step_state,
                            unsigned int* nan_flag, const unsigned int* skip_flag, gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    gh_clear_error();
    if (n_groups <= 0 || n_groups > GH_ADAM_MAX_GROUPS || (step < 1 && step_state == nullptr) || !params || !grads || !exp_avg || !exp_avg_sq || !sizes || !lrs)
        return gh_set_error(GH_E_INVALID_ARG, "gh_adam_step: bad group count / step or missing array");
    GhAdamGroups g;
    unsigned long long total = 0;
    for (int k = 0; k < GH_ADAM_MAX_GROUPS; k++) {
        const bool on = k < n_groups;
        g.param[k] = on ? params[k] : nullptr; g.grad[k] = on ? grads[k] : nullptr;
        g.exp_avg[k] = on ? exp_avg[k] : nullptr; g.exp_avg_sq[k] = on ? exp_avg_sq[k] : nullptr;
        g.lr[k] = on ? lrs[k] : 0.f;
        if (on) {
            if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k])
                return gh_set_error(GH_E_INVALID_ARG, "gh_adam_step: NULL parameter / gradient / moment pointer");
            total += sizes[k];
        }
        g.end[k] = total;
    }
    g.n = n_groups;
    if (total == 0) return GH_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)(step < 1 ? 1 : step));
    const double bc2 = 1.0 - pow((double)beta2, (double)(step < 1 ? 1 : step));
    // one grid row per group; enough CTAs per row for the largest group to fill the machine
    unsigned long long largest = 0;
    for (int k = 0; k < n_groups; k++) largest = sizes[k] > largest ? sizes[k] : largest;
    const unsigned long long want = (largest / 4 + 255) / 256;
    const dim3 grid((unsigned int)(want < 1 ? 1 : (want > 148ull * 8 ? 148ull * 8 : want)), (unsigned int)n_groups);
    if (nan_flag) {
        if (cudaMemsetAsync(nan_flag, 0, sizeof(unsigned int), stream) != cudaSuccess)
            return gh_set_error(GH_E_CUDA, "gh_adam_step: memset of the NaN flag failed");
        gh_adam_nan_kernel<<<grid, 256, 0, stream>>>(g, nan_flag);
        gh_count_launches(1);
    }
    gh_adam_update_kernel<<<grid, 256, 0, stream>>>(g, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), nan_flag, skip_flag, step_state);
    gh_count_launches(1);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GH_OK : gh_set_error(GH_E_CUDA, cudaGetErrorString(e));
}
