// M5 optimiser: Adam over one flat fp32 parameter buffer (replaces torch.optim.Adam's per-tensor loop,
// /root/reference/run_graphcount.py:478,505; run_zinc.py:263; run_ogb_mol.py:436). Same update rule as
// torch.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  Also counts library launches (bench evidence).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/escgnn_b200.h"
#include "launch.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale) {
    escgnn::pdl_enter();
    const int64_t n4 = n >> 2;
    const float step = lr / bc1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float* pp = &pv.x; float* mm = &mv.x; float* vp = &vv.x; const float* gg = &gv.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * grad_scale;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vp[k] = b2 * vp[k] + (1.f - b2) * gk * gk;
            pp[k] -= step * mm[k] / (sqrtf(vp[k]) / bc2_sqrt + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        const float gk = g[i] * grad_scale;
        m[i] = b1 * m[i] + (1.f - b1) * gk;
        v[i] = b2 * v[i] + (1.f - b2) * gk * gk;
        p[i] -= step * m[i] / (sqrtf(v[i]) / bc2_sqrt + eps);
    }
}

// device-side step counter + bias corrections, so a captured CUDA graph stays valid from step to step
// hyper: [0] lr, [1] beta1, [2] beta2, [3] eps, [4] grad_scale, [5] bc1 (out), [6] sqrt(bc2) (out); state: [0] step
__global__ void adam_tick_kernel(float* hyper, long long* state) {
    escgnn::pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const long long t = ++state[0];
        hyper[5] = (float)(1.0 - pow((double)hyper[1], (double)t));
        hyper[6] = (float)sqrt(1.0 - pow((double)hyper[2], (double)t));
    }
}

__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                int64_t n, const float* __restrict__ hyper) {
    escgnn::pdl_enter();
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], gs = hyper[4], bc1 = hyper[5], bc2s = hyper[6];
    const float step = lr / bc1;
    const int64_t n4 = n >> 2;                      // n is padded to a multiple of 4 by FlatAdam
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float* pp = &pv.x; float* mm = &mv.x; float* vp = &vv.x; const float* gg = &gv.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * gs;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vp[k] = b2 * vp[k] + (1.f - b2) * gk * gk;
            pp[k] -= step * mm[k] / (sqrtf(vp[k]) / bc2s + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
}

}  // namespace

extern "C" int escgnn_adam_step_device(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                       int64_t n, float* d_hyper, long long* d_state, void* stream) {
    if (n <= 0 || (n & 3)) return ESCGNN_ERR_BAD_ARG;
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    escgnn::launch_pdl(adam_tick_kernel, 1, 32, 0, (cudaStream_t)stream, d_hyper, d_state);
    escgnn::launch_pdl(adam_dev_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, d_hyper);
    return (int)cudaGetLastError();
}

extern "C" int escgnn_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                                float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                                void* stream) {
    if (n <= 0 || step < 1) return ESCGNN_ERR_BAD_ARG;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    int64_t blocks = ((n >> 2) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    escgnn::launch_pdl(adam_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, lr, beta1,
                                                                    beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
    return (int)cudaGetLastError();
}
