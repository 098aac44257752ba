// "Next" row 2: the optimiser step that follows the rasterizer backward.
// One launch updates every parameter group with torch.optim.Adam's arithmetic
// (reference: torch.optim.Adam(l, lr=0.0, eps=1e-15), SRC/scene/gaussian_model.py:431-448) and the
// reference's NaN guard (SRC/train_gaussians.py:174-181: skip the step if any gradient has a NaN,
// done there with seven blocking `.isnan().any()` host syncs) is a device-side flag: no host sync.
#include "gh_common.cuh"
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

__device__ __forceinline__ int gh_adam_find(const GhAdamGroups& g, unsigned long long i) {
    int k = 0;
#pragma unroll
    for (int j = 0; j < GH_ADAM_MAX_GROUPS; j++) k += (j < g.n - 1 && i >= g.end[j]) ? 1 : 0;
    return k;
}

__global__ void __launch_bounds__(256)
gh_adam_nan_kernel(GhAdamGroups g, unsigned long long total, unsigned int* __restrict__ flag)
{
    bool bad = false;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int k = gh_adam_find(g, i);
        const unsigned long long o = i - (k ? g.end[k - 1] : 0ull);
        const float v = g.grad[k][o];
        bad |= (v != v);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

__global__ void __launch_bounds__(256)
gh_adam_update_kernel(GhAdamGroups g, unsigned long long total, float beta1, float beta2, float eps,
                      float bc1, float bc2_sqrt, const unsigned int* __restrict__ flag, int* step_state)
{
    if (flag != nullptr && *flag != 0u) return;     // a gradient held a NaN: skip this step entirely
    if (step_state != nullptr) {
        // device-resident step count (only advanced by steps that were not skipped, like torch's state['step'])
        const int step = step_state[0] + 1;
        bc1 = 1.0f - powf(beta1, (float)step);
        bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    }
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int k = gh_adam_find(g, i);
        const unsigned long long o = i - (k ? g.end[k - 1] : 0ull);
        const float gr = g.grad[k][o];
        float m = g.exp_avg[k][o], v = g.exp_avg_sq[k][o];
        m = m + (1.0f - beta1) * (gr - m);                       // exp_avg.lerp_(grad, 1 - beta1)
        v = v * beta2 + (1.0f - beta2) * gr * gr;                // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        const float step_size = g.lr[k] / bc1;
        g.param[k][o] = g.param[k][o] - step_size * (m / denom); // addcdiv_(exp_avg, denom, -step_size)
        g.exp_avg[k][o] = m;
        g.exp_avg_sq[k][o] = v;
    }
    if (step_state != nullptr) {
        // the last CTA to finish advances the counter (every CTA has read it by then)
        __threadfence();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(reinterpret_cast<unsigned int*>(step_state + 1), 1u);
            if (t == gridDim.x - 1) { step_state[0] += 1; step_state[1] = 0; }
        }
    }
}

}  // namespace

extern "C" int gh_adam_step(int n_groups, float* const* params, const float* const* grads,
                            float* const* exp_avg, float* const* exp_avg_sq,
                            const unsigned long long* sizes, const float* lrs,
                            float beta1, float beta2, float eps, int step, int* step_state,
                            unsigned int* nan_flag, gh_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_groups <= 0 || n_groups > GH_ADAM_MAX_GROUPS || (step < 1 && step_state == nullptr) || !params || !grads || !exp_avg || !exp_avg_sq || !sizes || !lrs)
        return GH_E_INVALID_ARG;
    GhAdamGroups g;
    unsigned long long total = 0;
    for (int k = 0; k < GH_ADAM_MAX_GROUPS; k++) {
        const bool on = k < n_groups;
        g.param[k] = on ? params[k] : nullptr; g.grad[k] = on ? grads[k] : nullptr;
        g.exp_avg[k] = on ? exp_avg[k] : nullptr; g.exp_avg_sq[k] = on ? exp_avg_sq[k] : nullptr;
        g.lr[k] = on ? lrs[k] : 0.f;
        if (on) { if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k]) return GH_E_INVALID_ARG; total += sizes[k]; }
        g.end[k] = total;
    }
    g.n = n_groups;
    if (total == 0) return GH_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)(step < 1 ? 1 : step));
    const double bc2 = 1.0 - pow((double)beta2, (double)(step < 1 ? 1 : step));
    const int blocks = (int)((total + 255) / 256 < 148ull * 16 ? (total + 255) / 256 : 148ull * 16);
    if (nan_flag) {
        if (cudaMemsetAsync(nan_flag, 0, sizeof(unsigned int), stream) != cudaSuccess) return GH_E_CUDA;
        gh_adam_nan_kernel<<<blocks, 256, 0, stream>>>(g, total, nan_flag);
    }
    gh_adam_update_kernel<<<blocks, 256, 0, stream>>>(g, total, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), nan_flag, step_state);
    return cudaGetLastError() == cudaSuccess ? GH_OK : GH_E_CUDA;
}
