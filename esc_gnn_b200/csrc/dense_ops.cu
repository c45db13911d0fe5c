// Row-wise dense kernels of the static-shape training engine: training-mode
// BatchNorm1d + activation (forward and backward), column sums for bias gradients, embedding lookup, L1 / BCE loss.
//
// Everything works on row-major fp32 [rows_cap, C] buffers with a leading dimension, and reads the ACTUAL row count
// from device memory (`d_rows`), so one captured CUDA graph serves batches of any size up to the capacity; rows at
// or beyond the actual count are written as zeros, which keeps them inert in the GEMMs that follow.
// Reductions are two-stage and ordered (per-tile partials, then a fixed-order sum): results are run-to-run
// deterministic.  Reference semantics: torch.nn.BatchNorm1d (momentum 0.1, biased variance for normalisation,
// unbiased for the running estimate) as used at run_graphcount.py:54-61,78-87; zinc_models.py:513-522;
// ogb_mol_gnn.py:331-336,672.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/escgnn_b200.h"

namespace {

constexpr int kTileRows = 128;     // rows per CTA (512 was measured slower: too few CTAs in flight)
constexpr int kCols = 32;          // columns per CTA (one per lane)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : expm1f(v);
    return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
    if (act == 1) return v > 0.f ? 1.f : 0.f;
    if (act == 2) return v > 0.f ? 1.f : expf(v);
    return 1.f;
}

// cross-warp reduction of two per-lane values; result valid in warp 0
__device__ __forceinline__ void cta_reduce2(float& a, float& b, float (*s)[2][kCols]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    s[warp][0][lane] = a; s[warp][1][lane] = b;
    __syncthreads();
    if (warp == 0) {
        float x = 0.f, y = 0.f;
        for (int w = 0; w < nw; ++w) { x += s[w][0][lane]; y += s[w][1][lane]; }
        a = x; b = y;
    }
}

// partial[tile][0][c] = sum_r x[r][c], partial[tile][1][c] = sum_r x[r][c]^2 over the tile's valid rows
__global__ void __launch_bounds__(256)
colstats_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ d_rows, int C, float* __restrict__ partial) {
    __shared__ float s[8][2][kCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = *d_rows;
    float a = 0.f, b = 0.f;
    if (c < C)
        #pragma unroll 8
        for (int r = r0 + warp; r < min(r0 + kTileRows, rows); r += 8) { const float v = x[(size_t)r * ldx + c]; a += v; b += v * v; }
    cta_reduce2(a, b, s);
    if (warp == 0 && c < C) { partial[((size_t)blockIdx.y * 2 + 0) * C + c] = a; partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b; }
}

// y = act((x - mean) * rstd * gamma + beta); the first row tile also finalises mean / rstd / running stats.
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ partial, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* running_mean, float* running_var, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, int act, float eps, float momentum, int training,
                  const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ y, int ldy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    const int tiles = (rows + kTileRows - 1) / kTileRows;
    float mean = 0.f, rstd = 1.f, g = 1.f, bt = 0.f;
    if (c < C) {
        if (training) {
            float s1 = 0.f, s2 = 0.f;
            for (int t = 0; t < tiles; ++t) { s1 += partial[((size_t)t * 2 + 0) * C + c]; s2 += partial[((size_t)t * 2 + 1) * C + c]; }
            const float m = (float)max(rows, 1);
            mean = s1 / m;
            const float var = fmaxf(s2 / m - mean * mean, 0.f);
            rstd = rsqrtf(var + eps);
            if (blockIdx.y == 0 && warp == 0) {
                mean_out[c] = mean; rstd_out[c] = rstd;
                if (rows > 0) {
                    const float unbiased = rows > 1 ? var * m / (m - 1.f) : var;
                    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
                }
            }
        } else {
            mean = running_mean[c]; rstd = rsqrtf(running_var[c] + eps);
            if (blockIdx.y == 0 && warp == 0) { mean_out[c] = mean; rstd_out[c] = rstd; }
        }
        g = gamma ? gamma[c] : 1.f; bt = beta ? beta[c] : 0.f;
    }
    if (c >= C) return;
    const float sc = rstd * g, sh = bt - mean * sc;
    #pragma unroll 8
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8) {
        float o = 0.f;
        if (r < rows) o = act_fwd(x[(size_t)r * ldx + c] * sc + sh, act);
        y[(size_t)r * ldy + c] = o;               // rows >= actual count are zeroed (inert in the next GEMM)
    }
}

// pass 1 of the backward: partial sums of dz and dz * xhat, dz = (dy [+ dy2]) * act'(bn(x))
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                         const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         int act, const int* __restrict__ d_rows, int C, float* __restrict__ partial) {
    __shared__ float s[8][2][kCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = *d_rows;
    float a = 0.f, b = 0.f;
    if (c < C) {
        const float mu = mean[c], rs = rstd[c], g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
        #pragma unroll 4
        for (int r = r0 + warp; r < min(r0 + kTileRows, rows); r += 8) {
            const float xhat = (x[(size_t)r * ldx + c] - mu) * rs;
            float d = dy[(size_t)r * lddy + c];
            if (dy2) d += dy2[(size_t)r * lddy2 + c];
            const float dz = d * act_grad(xhat * g + bt, act);
            a += dz; b += dz * xhat;
        }
    }
    cta_reduce2(a, b, s);
    if (warp == 0 && c < C) { partial[((size_t)blockIdx.y * 2 + 0) * C + c] = a; partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b; }
}

// pass 2: dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat));  dgamma / dbeta from the first row tile.
// training == 0 (eval-mode BN): dx = gamma * rstd * dz.
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                        const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                        int act, int training, const float* __restrict__ partial, const int* __restrict__ d_rows,
                        int rows_cap, int C, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx,
                        int lddx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    if (c >= C) return;
    const int tiles = (rows + kTileRows - 1) / kTileRows;
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < tiles; ++t) { s1 += partial[((size_t)t * 2 + 0) * C + c]; s2 += partial[((size_t)t * 2 + 1) * C + c]; }
    if (blockIdx.y == 0 && warp == 0) { if (dgamma) dgamma[c] = s2; if (dbeta) dbeta[c] = s1; }
    const float mu = mean[c], rs = rstd[c], g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    const float inv_m = training ? 1.f / (float)max(rows, 1) : 0.f;
    const float m1 = s1 * inv_m, m2 = s2 * inv_m, k = g * rs;
    #pragma unroll 4
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8) {
        float o = 0.f;
        if (r < rows) {
            const float xhat = (x[(size_t)r * ldx + c] - mu) * rs;
            float d = dy[(size_t)r * lddy + c];
            if (dy2) d += dy2[(size_t)r * lddy2 + c];
            const float dz = d * act_grad(xhat * g + bt, act);
            o = k * (dz - m1 - xhat * m2);
        }
        dx[(size_t)r * lddx + c] = o;
    }
}

// activation only (no BatchNorm): y = act(x), rows beyond the count zeroed; backward: dx = dy * act'(x)
__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ x, int ldx, int act, const int* __restrict__ d_rows, int rows_cap, int C,
               float* __restrict__ y, int ldy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    if (c >= C) return;
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8) y[(size_t)r * ldy + c] = r < rows ? act_fwd(x[(size_t)r * ldx + c], act) : 0.f;
}
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy, int act,
               const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ dx, int lddx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    if (c >= C) return;
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8)
        dx[(size_t)r * lddx + c] = r < rows ? dy[(size_t)r * lddy + c] * act_grad(x[(size_t)r * ldx + c], act) : 0.f;
}

// column sums (bias gradients): partial per tile, then an ordered final sum
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ d_rows, int C, float* __restrict__ partial) {
    __shared__ float s[8][2][kCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = *d_rows;
    float a = 0.f, b = 0.f;
    if (c < C) for (int r = r0 + warp; r < min(r0 + kTileRows, rows); r += 8) a += x[(size_t)r * ldx + c];
    cta_reduce2(a, b, s);
    if (warp == 0 && c < C) partial[(size_t)blockIdx.y * C + c] = a;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, const int* __restrict__ d_rows, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int tiles = (*d_rows + kTileRows - 1) / kTileRows;
    float s = 0.f;
    for (int t = 0; t < tiles; ++t) s += partial[(size_t)t * C + c];
    out[c] = s;
}

// embedding rows: y[r, :] = table[idx[r]] (rows beyond the count zeroed); backward scatters with atomics (tiny tables)
__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int n_cols_idx,
                                     const int64_t* __restrict__ col_offsets, const int* __restrict__ d_rows, int rows_cap,
                                     int C, float* __restrict__ y, int ldy) {
    const int rows = *d_rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows_cap * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        float v = 0.f;
        if (r < rows)
            for (int k = 0; k < n_cols_idx; ++k)     // sum of per-column embeddings (AtomEncoder / BondEncoder)
                v += table[(size_t)(idx[(size_t)r * n_cols_idx + k] + (col_offsets ? col_offsets[k] : 0)) * C + c];
        y[(size_t)r * ldy + c] = v;
    }
}
__global__ void embedding_bwd_kernel(const float* __restrict__ dy, int lddy, const int64_t* __restrict__ idx, int n_cols_idx,
                                     const int64_t* __restrict__ col_offsets, const int* __restrict__ d_rows, int C,
                                     float* __restrict__ dtable) {
    const int rows = *d_rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        const float g = dy[(size_t)r * lddy + c];
        for (int k = 0; k < n_cols_idx; ++k)
            atomicAdd(&dtable[(size_t)(idx[(size_t)r * n_cols_idx + k] + (col_offsets ? col_offsets[k] : 0)) * C + c], g);
    }
}

// losses over `rows` predictions [rows, T]: kind 0 = L1 mean (run_graphcount.py:498, run_zinc.py:283),
// kind 1 = BCE-with-logits mean over labelled entries y == y (run_ogb_mol.py:58-74). Single CTA, ordered sum.
__global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ target, int kind, const int* __restrict__ d_rows,
            int T, float* __restrict__ loss, float* __restrict__ dpred, int lddp, int rows_cap) {
    __shared__ float s_sum[32], s_cnt[32];
    const int rows = *d_rows;
    float acc = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < rows * T; i += blockDim.x) {
        const int r = i / T, c = i % T;
        const float p = pred[(size_t)r * ldp + c], t = target[i];
        if (kind == 0) { acc += fabsf(p - t); cnt += 1.f; }
        else if (t == t) { acc += fmaxf(p, 0.f) - p * t + log1pf(expf(-fabsf(p))); cnt += 1.f; }
    }
    for (int d = 16; d; d >>= 1) { acc += __shfl_xor_sync(kFull, acc, d); cnt += __shfl_xor_sync(kFull, cnt, d); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    float tot = 0.f, n = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tot += s_sum[w]; n += s_cnt[w]; }
    n = fmaxf(n, 1.f);
    if (threadIdx.x == 0) loss[0] = tot / n;
    for (int i = threadIdx.x; i < rows_cap * T; i += blockDim.x) {
        const int r = i / T, c = i % T;
        float g = 0.f;
        if (r < rows) {
            const float p = pred[(size_t)r * ldp + c], t = target[i];
            if (kind == 0) g = (p > t ? 1.f : (p < t ? -1.f : 0.f)) / n;
            else if (t == t) g = (1.f / (1.f + expf(-p)) - t) / n;
        }
        dpred[(size_t)r * lddp + c] = g;
    }
}

inline dim3 tile_grid(int rows_cap, int C) { return dim3((unsigned)((C + kCols - 1) / kCols), (unsigned)((rows_cap + kTileRows - 1) / kTileRows)); }

}  // namespace

extern "C" {

int escgnn_dense_tile_rows(void) { return kTileRows; }

int escgnn_bn_act_fwd(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float* d_running_mean,
                      float* d_running_var, float* d_mean, float* d_rstd, float* d_partial, int act, float eps,
                      float momentum, int training, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 g = tile_grid(rows_cap, channels);
    if (training) colstats_kernel<<<g, 256, 0, st>>>(d_x, ldx, d_rows, channels, d_partial);
    bn_act_fwd_kernel<<<g, 256, 0, st>>>(d_x, ldx, d_partial, d_gamma, d_beta, d_running_mean, d_running_var, d_mean, d_rstd,
                                         act, eps, momentum, training, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_bn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, const float* d_dy2, int lddy2,
                      const float* d_mean, const float* d_rstd, const float* d_gamma, const float* d_beta, int act,
                      int training, float* d_partial, const int* d_rows, int rows_cap, int channels, float* d_dgamma,
                      float* d_dbeta, float* d_dx, int lddx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 g = tile_grid(rows_cap, channels);
    bn_act_bwd_reduce_kernel<<<g, 256, 0, st>>>(d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                                d_rows, channels, d_partial);
    bn_act_bwd_apply_kernel<<<g, 256, 0, st>>>(d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                               training, d_partial, d_rows, rows_cap, channels, d_dgamma, d_dbeta, d_dx, lddx);
    return (int)cudaGetLastError();
}

int escgnn_act_fwd(const float* d_x, int ldx, int act, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                   void* stream) {
    act_fwd_kernel<<<tile_grid(rows_cap, channels), 256, 0, (cudaStream_t)stream>>>(d_x, ldx, act, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, int act, const int* d_rows, int rows_cap,
                   int channels, float* d_dx, int lddx, void* stream) {
    act_bwd_kernel<<<tile_grid(rows_cap, channels), 256, 0, (cudaStream_t)stream>>>(d_x, ldx, d_dy, lddy, act, d_rows, rows_cap, channels, d_dx, lddx);
    return (int)cudaGetLastError();
}

int escgnn_colsum(const float* d_x, int ldx, const int* d_rows, int rows_cap, int channels, float* d_partial, float* d_out,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    colsum_partial_kernel<<<tile_grid(rows_cap, channels), 256, 0, st>>>(d_x, ldx, d_rows, channels, d_partial);
    colsum_final_kernel<<<(unsigned)((channels + 127) / 128), 128, 0, st>>>(d_partial, d_rows, channels, d_out);
    return (int)cudaGetLastError();
}

int escgnn_embedding_fwd(const float* d_table, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_y, int ldy, void* stream) {
    int64_t total = (int64_t)rows_cap * channels;
    unsigned b = (unsigned)((total + 255) / 256); if (b > 1184) b = 1184; if (b < 1) b = 1;
    embedding_fwd_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(d_table, d_idx, idx_cols, d_col_offsets, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_embedding_bwd(const float* d_dy, int lddy, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_dtable, void* stream) {
    int64_t total = (int64_t)rows_cap * channels;
    unsigned b = (unsigned)((total + 255) / 256); if (b > 1184) b = 1184; if (b < 1) b = 1;
    embedding_bwd_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(d_dy, lddy, d_idx, idx_cols, d_col_offsets, d_rows, channels, d_dtable);
    return (int)cudaGetLastError();
}

int escgnn_loss_fwd_bwd(const float* d_pred, int ldp, const float* d_target, int kind, const int* d_rows, int rows_cap,
                        int n_targets, float* d_loss, float* d_dpred, int lddp, void* stream) {
    loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_pred, ldp, d_target, kind, d_rows, n_targets, d_loss, d_dpred, lddp, rows_cap);
    return (int)cudaGetLastError();
}

}  // extern "C"
