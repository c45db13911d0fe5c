// K2 `edge_distance`: Euclidean length of every edge from node positions (/root/reference/distance.py:29-47).
//   dist[e] = || pos[col_e] - pos[row_e] ||_2  (or its square), optionally divided by max_e dist[e] or a given value;
//   rel[e]  = pos[col_e] - pos[row_e]          (the `relative_pos` option, :43-45)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/escgnn_b200.h"

namespace {

__global__ void edge_distance_kernel(const float* __restrict__ pos, int dim, const int64_t* __restrict__ row,
                                     const int64_t* __restrict__ col, int64_t n_edges, int squared,
                                     float* __restrict__ dist, float* __restrict__ rel, unsigned* __restrict__ max_bits) {
    float local_max = 0.f;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const float* a = pos + row[e] * dim;
        const float* b = pos + col[e] * dim;
        float s = 0.f;
        for (int k = 0; k < dim; ++k) {
            const float d = b[k] - a[k];
            if (rel) rel[e * dim + k] = d;
            s += d * d;
        }
        const float v = squared ? s : sqrtf(s);
        dist[e] = v;
        local_max = fmaxf(local_max, v);
    }
    #pragma unroll
    for (int d = 16; d; d >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, d));
    // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0 && max_bits) atomicMax(max_bits, __float_as_uint(local_max));
}

__global__ void scale_by_max_kernel(float* __restrict__ dist, int64_t n, const unsigned* __restrict__ max_bits,
                                    float fixed) {
    const float m = max_bits ? __uint_as_float(*max_bits) : fixed;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        dist[e] = dist[e] / m;
}

}  // namespace

extern "C" int escgnn_edge_distance(const float* d_pos, int dim, const int64_t* d_row, const int64_t* d_col,
                                    int64_t n_edges, int squared, int norm, float max_value, float* d_dist,
                                    float* d_rel, unsigned* d_max_scratch, void* stream) {
    if (n_edges <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned blocks = (unsigned)((n_edges + 255) / 256);
    if (blocks > 2368) blocks = 2368;
    const bool data_max = norm && !(max_value > 0.f);
    if (data_max) cudaMemsetAsync(d_max_scratch, 0, sizeof(unsigned), st);
    edge_distance_kernel<<<blocks, 256, 0, st>>>(d_pos, dim, d_row, d_col, n_edges, squared, d_dist, d_rel,
                                                 data_max ? d_max_scratch : nullptr);
    if (norm) scale_by_max_kernel<<<blocks, 256, 0, st>>>(d_dist, n_edges, data_max ? d_max_scratch : nullptr, max_value);
    return (int)cudaGetLastError();
}
