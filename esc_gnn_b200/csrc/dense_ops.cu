// Row-wise dense kernels of the static-shape training engine: training-mode
// BatchNorm1d + activation (forward and backward), column sums for bias gradients, embedding lookup, L1 / BCE loss.
//
// Everything works on row-major fp32 [rows_cap, C] buffers with a leading dimension, and reads the ACTUAL row count
// from device memory (`d_rows`), so one captured CUDA graph serves batches of any size up to the capacity; rows at
// or beyond the actual count are written as zeros, which keeps them inert in the GEMMs that follow.
// Reductions are two-stage and ordered (per-tile partials, then a fixed-order sum): results are run-to-run
// deterministic.  Batch statistics are SHIFTED sums -- sum (x - x[0]) and sum (x - x[0])^2 per column, x[0] = the column's first row --
// so the variance S2/m - (S1/m)^2 does not cancel when |mean| >> std (and is exactly 0 for a constant column), and the
// normalisation is applied centred, (x - mean) * (rstd * gamma) + beta: both within ~1 ulp of torch's Welford statistics.
// Reference semantics: torch.nn.BatchNorm1d (momentum 0.1, biased variance for normalisation,
// unbiased for the running estimate) as used at run_graphcount.py:54-61,78-87; zinc_models.py:513-522;
// ogb_mol_gnn.py:331-336,672.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cooperative_groups.h>

#include <initializer_list>

#include "../../include/escgnn_b200.h"
#include "launch.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kTileRows = 128;     // rows per CTA (512 was measured slower: too few CTAs in flight)
constexpr int kCols = 32;          // columns per CTA (one per lane)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float act_fwd(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : expm1f(v);   // (a Taylor / exp(v) - 1 hybrid was measured SLOWER than the library expm1f: tools/trace_bn.py)
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
    escgnn::pdl_enter();
    __shared__ float s[8][2][kCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = *d_rows;
    float a = 0.f, b = 0.f;
    if (c < C && rows > 0) {
        const float shift = x[c];
        #pragma unroll 8
        for (int r = r0 + warp; r < min(r0 + kTileRows, rows); r += 8) { const float v = x[(size_t)r * ldx + c] - shift; a += v; b += v * v; }
    }
    cta_reduce2(a, b, s);
    if (warp == 0 && c < C) { partial[((size_t)blockIdx.y * 2 + 0) * C + c] = a; partial[((size_t)blockIdx.y * 2 + 1) * C + c] = b; }
}

// y = act((x - mean) * rstd * gamma + beta); the first row tile also finalises mean / rstd / running stats.
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ partial, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* running_mean, float* running_var, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, int act, float eps, float momentum, int training,
                  const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ y, int ldy) {
    escgnn::pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    const int tiles = (rows + kTileRows - 1) / kTileRows;
    float mean = 0.f, rstd = 1.f, g = 1.f, bt = 0.f;
    if (c < C) {
        if (training) {
            float s1 = 0.f, s2 = 0.f;
            for (int t = 0; t < tiles; ++t) { s1 += partial[((size_t)t * 2 + 0) * C + c]; s2 += partial[((size_t)t * 2 + 1) * C + c]; }
            const float m = (float)max(rows, 1), m1 = s1 / m;
            mean = (rows > 0 ? x[c] : 0.f) + m1;
            const float var = fmaxf(s2 / m - m1 * m1, 0.f);
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
    const float sc = rstd * g;
    #pragma unroll 8
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8) {
        float o = 0.f;
        if (r < rows) o = act_fwd((x[(size_t)r * ldx + c] - mean) * sc + bt, act);
        y[(size_t)r * ldy + c] = o;               // rows >= actual count are zeroed (inert in the next GEMM)
    }
}

// pass 1 of the backward: partial sums of dz and dz * xhat, dz = (dy [+ dy2]) * act'(bn(x))
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                         const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         int act, const int* __restrict__ d_rows, int C, float* __restrict__ partial) {
    escgnn::pdl_enter();
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
    escgnn::pdl_enter();
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

// ---------------------------------------------------------------------------------------------------------------
// float4 variants (every operand 16-byte aligned, C and all leading dimensions multiples of 4 - the engine's case).
// One CTA = kVRows rows x kVCols columns, a lane owns 4 adjacent columns and a warp a row, so a warp reads 512
// contiguous bytes per row and keeps 8 independent 16-byte loads in flight per thread.  The reduction kernels write
// per-tile partials, and the LAST tile to arrive (a ticket per column block at the head of the workspace) sums them
// in a fixed order and finalises the statistics, so the apply kernels read two numbers per column instead of
// re-summing every tile in every CTA.  Workspace (floats): [0,64) tickets (left zero), [64, 64+2C) final sums,
// then [tile][2][C] partials.
constexpr int kVRows = 64, kVCols = 128, kHdr = 64;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
// act((v - mean) * sc + bt), per component
__device__ __forceinline__ float4 bn_apply4(const float4& v, const float4& mean, const float4& sc, const float4& bt) {
    return make_float4((v.x - mean.x) * sc.x + bt.x, (v.y - mean.y) * sc.y + bt.y, (v.z - mean.z) * sc.z + bt.z, (v.w - mean.w) * sc.w + bt.w);
}
// (S1, S2) shifted sums over m rows -> mean (shift added back) and biased variance, per component
__device__ __forceinline__ void bn_finish4(const float4& s1, const float4& s2, const float4& shift, float m, float4& mean, float4& var) {
    const float4 m1 = make_float4(s1.x / m, s1.y / m, s1.z / m, s1.w / m);
    mean = make_float4(shift.x + m1.x, shift.y + m1.y, shift.z + m1.z, shift.w + m1.w);
    var = make_float4(fmaxf(s2.x / m - m1.x * m1.x, 0.f), fmaxf(s2.y / m - m1.y * m1.y, 0.f), fmaxf(s2.z / m - m1.z * m1.z, 0.f),
                      fmaxf(s2.w / m - m1.w * m1.w, 0.f));
}

// Per-CTA sums of (a, b) over its 8 warps -> tile partial; returns true (uniformly) in the last CTA of the column block.
__device__ __forceinline__ bool tile_commit(const float4& a, const float4& b, float* ws, int C, float (*s)[2][kVCols], int* s_last) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    st4(&s[warp][0][lane * 4], a);
    st4(&s[warp][1][lane * 4], b);
    __syncthreads();
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127, c = blockIdx.x * kVCols + col;
    float v = 0.f;
    #pragma unroll
    for (int w = 0; w < 8; ++w) v += s[w][which][col];
    if (c < C) ws[kHdr + 2 * C + ((size_t)blockIdx.y * 2 + which) * C + c] = v;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_last = atomicAdd(reinterpret_cast<unsigned*>(ws) + blockIdx.x, 1u) == gridDim.y - 1;
    __syncthreads();
    return *s_last != 0;
}

// Last CTA: ordered sum of the first `tiles` tile partials; valid in threads [0, 128) (one column each).
__device__ __forceinline__ void final_sums(float* ws, int C, int tiles, float (*s)[2][kVCols], float& s1, float& s2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c4 = blockIdx.x * kVCols + lane * 4;
    __threadfence();
    float4 a = zero4(), b = zero4();
    if (c4 < C) {
        const float* base = ws + kHdr + 2 * C + c4;
        #pragma unroll 4
        for (int t = warp; t < tiles; t += 8) {
            const float4 u = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(t * 2) * C));
            const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)(t * 2 + 1) * C));
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
    }
    st4(&s[warp][0][lane * 4], a);
    st4(&s[warp][1][lane * 4], b);
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
    if (threadIdx.x < kVCols) {
        #pragma unroll
        for (int w = 0; w < 8; ++w) { s1 += s[w][0][threadIdx.x]; s2 += s[w][1][threadIdx.x]; }
    }
    if (threadIdx.x == 0) reinterpret_cast<unsigned*>(ws)[blockIdx.x] = 0u;      // ticket ready for the next launch
}

// mode 0: BatchNorm statistics (sum x, sum x^2 -> mean / rstd / running stats); mode 1: column sums only (-> out_sum)
template <int MODE>
__global__ void __launch_bounds__(256)
colstats_v4_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ d_rows, int rows_cap, int C, float* ws,
                   float* running_mean, float* running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                   float eps, float momentum, float* __restrict__ out_sum) {
    escgnn::pdl_enter();
    __shared__ __align__(16) float s[8][2][kVCols];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = blockIdx.x * kVCols + lane * 4, rows = min(*d_rows, rows_cap);
    float4 a = zero4(), b = zero4();
    const float4 sh = (MODE == 0 && c4 < C && rows > 0) ? ld4(x + c4) : zero4();       // statistics: sums shifted by the first row
    if (c4 < C) {
        for (int r0 = blockIdx.y * kVRows + warp; r0 < rows; r0 += gridDim.y * kVRows) {      // row chunks of this tile
            float4 v[kVRows / 8];
            #pragma unroll
            for (int i = 0; i < kVRows / 8; ++i) v[i] = r0 + 8 * i < rows ? sub4(ld4(x + (size_t)(r0 + 8 * i) * ldx + c4), sh) : zero4();
            #pragma unroll
            for (int i = 0; i < kVRows / 8; ++i) {
                a.x += v[i].x; a.y += v[i].y; a.z += v[i].z; a.w += v[i].w;
                if (MODE == 0) { b.x += v[i].x * v[i].x; b.y += v[i].y * v[i].y; b.z += v[i].z * v[i].z; b.w += v[i].w * v[i].w; }
            }
        }
    }
    if (!tile_commit(a, b, ws, C, s, &s_last)) return;
    float s1, s2;
    final_sums(ws, C, min((int)gridDim.y, (rows + kVRows - 1) / kVRows), s, s1, s2);
    const int c = blockIdx.x * kVCols + threadIdx.x;
    if (threadIdx.x >= kVCols || c >= C) return;
    if (MODE == 1) { out_sum[c] = s1; return; }
    const float m = (float)max(rows, 1), m1 = s1 / m, mean = (rows > 0 ? x[c] : 0.f) + m1, var = fmaxf(s2 / m - m1 * m1, 0.f);
    mean_out[c] = mean; rstd_out[c] = rsqrtf(var + eps);
    if (rows > 0) {
        const float unbiased = rows > 1 ? var * m / (m - 1.f) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

__device__ __forceinline__ float4 act_fwd4(const float4& v, int act) {
    return make_float4(act_fwd(v.x, act), act_fwd(v.y, act), act_fwd(v.z, act), act_fwd(v.w, act));
}

// y = act((x - mean) * rstd * gamma + beta) with the statistics already final (training) or the running ones (eval)
__global__ void __launch_bounds__(256)
bn_act_fwd_v4_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ running_mean, const float* __restrict__ running_var, float* mean_io, float* rstd_io,
                     int act, float eps, int training, const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ y,
                     int ldy) {
    escgnn::pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = blockIdx.x * kVCols + lane * 4, r0 = blockIdx.y * kVRows + warp, rows = min(*d_rows, rows_cap);
    if (c4 >= C) return;
    float4 mean, rstd;
    if (training) { mean = ld4(mean_io + c4); rstd = ld4(rstd_io + c4); }
    else {
        mean = ld4(running_mean + c4);
        const float4 rv = ld4(running_var + c4);
        rstd = make_float4(rsqrtf(rv.x + eps), rsqrtf(rv.y + eps), rsqrtf(rv.z + eps), rsqrtf(rv.w + eps));
        if (blockIdx.y == 0 && warp == 0) { st4(mean_io + c4, mean); st4(rstd_io + c4, rstd); }
    }
    const float4 g = gamma ? ld4(gamma + c4) : make_float4(1.f, 1.f, 1.f, 1.f), bt = beta ? ld4(beta + c4) : zero4();
    const float4 sc = make_float4(rstd.x * g.x, rstd.y * g.y, rstd.z * g.z, rstd.w * g.w);
    float4 v[kVRows / 8];
    #pragma unroll
    for (int i = 0; i < kVRows / 8; ++i) v[i] = r0 + 8 * i < rows ? ld4(x + (size_t)(r0 + 8 * i) * ldx + c4) : zero4();
    #pragma unroll
    for (int i = 0; i < kVRows / 8; ++i) {
        const int r = r0 + 8 * i;
        if (r >= rows_cap) break;
        float4 o = zero4();                           // rows >= actual count are zeroed (inert in the next GEMM)
        if (r < rows) o = act_fwd4(bn_apply4(v[i], mean, sc, bt), act);
        st4(y + (size_t)r * ldy + c4, o);
    }
}

struct BnCols { float4 mu, rs, g, bt; };
__device__ __forceinline__ BnCols bn_cols(const float* mean, const float* rstd, const float* gamma, const float* beta, int c4) {
    BnCols k;
    k.mu = ld4(mean + c4); k.rs = ld4(rstd + c4);
    k.g = gamma ? ld4(gamma + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
    k.bt = beta ? ld4(beta + c4) : zero4();
    return k;
}
// dz = d * act'(xhat * g + bt) and xhat, per component
__device__ __forceinline__ void bn_dz(const float4& x, const float4& d, const BnCols& k, int act, float4& dz, float4& xh) {
    xh = make_float4((x.x - k.mu.x) * k.rs.x, (x.y - k.mu.y) * k.rs.y, (x.z - k.mu.z) * k.rs.z, (x.w - k.mu.w) * k.rs.w);
    dz = make_float4(d.x * act_grad(xh.x * k.g.x + k.bt.x, act), d.y * act_grad(xh.y * k.g.y + k.bt.y, act),
                     d.z * act_grad(xh.z * k.g.z + k.bt.z, act), d.w * act_grad(xh.w * k.g.w + k.bt.w, act));
}

__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_v4_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                            const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                            int act, const int* __restrict__ d_rows, int rows_cap, int C, float* ws, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
    escgnn::pdl_enter();
    __shared__ __align__(16) float s[8][2][kVCols];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = blockIdx.x * kVCols + lane * 4, rows = min(*d_rows, rows_cap);
    float4 a = zero4(), b = zero4();
    if (c4 < C) {
        const BnCols k = bn_cols(mean, rstd, gamma, beta, c4);
        for (int r0 = blockIdx.y * kVRows + warp; r0 < rows; r0 += gridDim.y * kVRows)       // row chunks of this tile
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            float4 xv[kVRows / 16], dv[kVRows / 16];
            #pragma unroll
            for (int i = 0; i < kVRows / 16; ++i) {
                const int r = r0 + 8 * (h * (kVRows / 16) + i);
                xv[i] = zero4(); dv[i] = zero4();
                if (r < rows) {
                    xv[i] = ld4(x + (size_t)r * ldx + c4);
                    dv[i] = ld4(dy + (size_t)r * lddy + c4);
                    if (dy2) { const float4 e = ld4(dy2 + (size_t)r * lddy2 + c4); dv[i].x += e.x; dv[i].y += e.y; dv[i].z += e.z; dv[i].w += e.w; }
                }
            }
            #pragma unroll
            for (int i = 0; i < kVRows / 16; ++i) {
                const int r = r0 + 8 * (h * (kVRows / 16) + i);
                if (r < rows) {
                    float4 dz, xh;
                    bn_dz(xv[i], dv[i], k, act, dz, xh);
                    a.x += dz.x; a.y += dz.y; a.z += dz.z; a.w += dz.w;
                    b.x += dz.x * xh.x; b.y += dz.y * xh.y; b.z += dz.z * xh.z; b.w += dz.w * xh.w;
                }
            }
        }
    }
    if (!tile_commit(a, b, ws, C, s, &s_last)) return;
    float s1, s2;
    final_sums(ws, C, min((int)gridDim.y, (rows + kVRows - 1) / kVRows), s, s1, s2);
    const int c = blockIdx.x * kVCols + threadIdx.x;
    if (threadIdx.x >= kVCols || c >= C) return;
    ws[kHdr + c] = s1; ws[kHdr + C + c] = s2;
    if (dgamma) dgamma[c] = s2;
    if (dbeta) dbeta[c] = s1;
}

__global__ void __launch_bounds__(256)
bn_act_bwd_apply_v4_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                           const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                           int act, int training, const float* __restrict__ ws, const int* __restrict__ d_rows, int rows_cap, int C,
                           float* __restrict__ dx, int lddx) {
    escgnn::pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = blockIdx.x * kVCols + lane * 4, r0 = blockIdx.y * kVRows + warp, rows = min(*d_rows, rows_cap);
    if (c4 >= C) return;
    const BnCols k = bn_cols(mean, rstd, gamma, beta, c4);
    const float inv_m = training ? 1.f / (float)max(rows, 1) : 0.f;
    const float4 s1 = ld4(ws + kHdr + c4), s2 = ld4(ws + kHdr + C + c4);
    const float4 m1 = make_float4(s1.x * inv_m, s1.y * inv_m, s1.z * inv_m, s1.w * inv_m);
    const float4 m2 = make_float4(s2.x * inv_m, s2.y * inv_m, s2.z * inv_m, s2.w * inv_m);
    const float4 kk = make_float4(k.g.x * k.rs.x, k.g.y * k.rs.y, k.g.z * k.rs.z, k.g.w * k.rs.w);
    #pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 xv[kVRows / 16], dv[kVRows / 16];
        #pragma unroll
        for (int i = 0; i < kVRows / 16; ++i) {
            const int r = r0 + 8 * (h * (kVRows / 16) + i);
            xv[i] = zero4(); dv[i] = zero4();
            if (r < rows) {
                xv[i] = ld4(x + (size_t)r * ldx + c4);
                dv[i] = ld4(dy + (size_t)r * lddy + c4);
                if (dy2) { const float4 e = ld4(dy2 + (size_t)r * lddy2 + c4); dv[i].x += e.x; dv[i].y += e.y; dv[i].z += e.z; dv[i].w += e.w; }
            }
        }
        #pragma unroll
        for (int i = 0; i < kVRows / 16; ++i) {
            const int r = r0 + 8 * (h * (kVRows / 16) + i);
            if (r >= rows_cap) break;
            float4 o = zero4();
            if (r < rows) {
                float4 dz, xh;
                bn_dz(xv[i], dv[i], k, act, dz, xh);
                o = make_float4(kk.x * (dz.x - m1.x - xh.x * m2.x), kk.y * (dz.y - m1.y - xh.y * m2.y),
                                kk.z * (dz.z - m1.z - xh.z * m2.z), kk.w * (dz.w - m1.w - xh.w * m2.w));
            }
            st4(dx + (size_t)r * lddx + c4, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One-launch BatchNorm (+activation) for the row counts of a reference-sized batch: a thread-block CLUSTER of 8 CTAs
// owns 4*LANES columns and splits the rows; per-CTA partial sums meet through distributed shared memory, every CTA
// then holds the final statistics and normalises its rows (second read of x hits L2).  Replaces the statistics kernel
// + apply kernel pair, i.e. one launch and one dependent-launch gap per BatchNorm, forward and backward.
constexpr int kClRanks = 8;
constexpr int kClMaxRows = 1 << 16;          // above this the two-kernel path (all SMs busy) wins
constexpr int kClFwdUnroll = 8, kClBwdUnroll = 4;    // independent row loads in flight per thread

// debugging (escgnn_bn_set_trace, tools/trace_bn.py): 6 %globaltimer stamps per CTA of the register-resident cluster kernels
__constant__ unsigned long long* g_bn_trace = nullptr;     // (constant bank: the disabled check costs one cached load)
__device__ __forceinline__ void bn_stamp(int slot) {
    if (g_bn_trace != nullptr && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_bn_trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 6 + slot] = t;
    }
}

template <int LANES>
struct ClusterTile {
    static constexpr int kCols = 4 * LANES, kSlots = 256 / LANES, kSweep = kClRanks * kSlots;
    float (*s_w)[2][kCols];
    float (*s_part)[kCols];
    float (*s_tot)[kCols];
    // (a, b): this thread's partial sums for its 4 columns -> s_tot[0..1][cols] = cluster-wide sums (all threads may read)
    __device__ __forceinline__ void reduce(float4 a, float4 b, cg::cluster_group& cluster) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cl = threadIdx.x % LANES;
        #pragma unroll
        for (int d = LANES; d < 32; d <<= 1) {
            a.x += __shfl_xor_sync(kFull, a.x, d); a.y += __shfl_xor_sync(kFull, a.y, d);
            a.z += __shfl_xor_sync(kFull, a.z, d); a.w += __shfl_xor_sync(kFull, a.w, d);
            b.x += __shfl_xor_sync(kFull, b.x, d); b.y += __shfl_xor_sync(kFull, b.y, d);
            b.z += __shfl_xor_sync(kFull, b.z, d); b.w += __shfl_xor_sync(kFull, b.w, d);
        }
        if (lane < LANES) { st4(&s_w[warp][0][4 * cl], a); st4(&s_w[warp][1][4 * cl], b); }
        __syncthreads();
        const int which = threadIdx.x / kCols, col = threadIdx.x % kCols;
        if (threadIdx.x < 2 * kCols) {
            float v = 0.f;
            #pragma unroll
            for (int w = 0; w < 8; ++w) v += s_w[w][which][col];
            s_part[which][col] = v;
        }
        cluster.sync();
        if (threadIdx.x < 2 * kCols) {
            float v = 0.f;
            #pragma unroll
            for (int r = 0; r < kClRanks; ++r) v += *cluster.map_shared_rank(&s_part[which][col], r);     // fixed order: deterministic
            s_tot[which][col] = v;
        }
        __syncthreads();
    }
};

template <int LANES>
__global__ void __launch_bounds__(256)
bn_act_fwd_cluster_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* running_mean, float* running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                          int act, float eps, float momentum, const int* __restrict__ d_rows, int rows_cap, int C,
                          float* __restrict__ y, int ldy) {
    escgnn::pdl_enter();
    using T = ClusterTile<LANES>;
    __shared__ __align__(16) float s_w[8][2][T::kCols];
    __shared__ __align__(16) float s_part[2][T::kCols];
    __shared__ __align__(16) float s_tot[2][T::kCols];
    cg::cluster_group cluster = cg::this_cluster();
    T t{s_w, s_part, s_tot};
    const int cl = threadIdx.x % LANES, rs = threadIdx.x / LANES, rank = (int)cluster.block_rank();
    const int c4 = blockIdx.x * T::kCols + 4 * cl, rows = min(*d_rows, rows_cap);
    const bool col_ok = c4 < C;
    float4 a = zero4(), b = zero4();
    const float4 shift = (col_ok && rows > 0) ? ld4(x + c4) : zero4();
    if (col_ok)
        for (int r0 = rank * T::kSlots + rs; r0 < rows; r0 += kClFwdUnroll * T::kSweep) {
            float4 v[kClFwdUnroll];
            #pragma unroll
            for (int k = 0; k < kClFwdUnroll; ++k) v[k] = r0 + k * T::kSweep < rows ? sub4(ld4(x + (size_t)(r0 + k * T::kSweep) * ldx + c4), shift) : zero4();
            #pragma unroll
            for (int k = 0; k < kClFwdUnroll; ++k) {
                a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
                b.x += v[k].x * v[k].x; b.y += v[k].y * v[k].y; b.z += v[k].z * v[k].z; b.w += v[k].w * v[k].w;
            }
        }
    t.reduce(a, b, cluster);
    if (col_ok) {
        const float m = (float)max(rows, 1);
        const float4 s1 = ld4(&s_tot[0][4 * cl]), s2 = ld4(&s_tot[1][4 * cl]);
        float4 mean, var;
        bn_finish4(s1, s2, shift, m, mean, var);
        const float4 rstd = make_float4(rsqrtf(var.x + eps), rsqrtf(var.y + eps), rsqrtf(var.z + eps), rsqrtf(var.w + eps));
        if (rank == 0 && rs == 0) {
            st4(mean_out + c4, mean); st4(rstd_out + c4, rstd);
            if (rows > 0) {
                const float ub = rows > 1 ? m / (m - 1.f) : 1.f, k0 = 1.f - momentum;
                const float4 rm = ld4(running_mean + c4), rv = ld4(running_var + c4);
                st4(running_mean + c4, make_float4(k0 * rm.x + momentum * mean.x, k0 * rm.y + momentum * mean.y,
                                                   k0 * rm.z + momentum * mean.z, k0 * rm.w + momentum * mean.w));
                st4(running_var + c4, make_float4(k0 * rv.x + momentum * var.x * ub, k0 * rv.y + momentum * var.y * ub,
                                                  k0 * rv.z + momentum * var.z * ub, k0 * rv.w + momentum * var.w * ub));
            }
        }
        const float4 g = gamma ? ld4(gamma + c4) : make_float4(1.f, 1.f, 1.f, 1.f), bt = beta ? ld4(beta + c4) : zero4();
        const float4 sc = make_float4(rstd.x * g.x, rstd.y * g.y, rstd.z * g.z, rstd.w * g.w);
        for (int r0 = rank * T::kSlots + rs; r0 < rows_cap; r0 += kClFwdUnroll * T::kSweep) {
            float4 v[kClFwdUnroll];
            #pragma unroll
            for (int k = 0; k < kClFwdUnroll; ++k) v[k] = r0 + k * T::kSweep < rows ? ld4(x + (size_t)(r0 + k * T::kSweep) * ldx + c4) : zero4();
            #pragma unroll
            for (int k = 0; k < kClFwdUnroll; ++k) {
                const int r = r0 + k * T::kSweep;
                if (r >= rows_cap) break;
                float4 o = zero4();                   // rows >= actual count are zeroed (inert in the next GEMM)
                if (r < rows) o = act_fwd4(bn_apply4(v[k], mean, sc, bt), act);
                st4(y + (size_t)r * ldy + c4, o);
            }
        }
    }
    cluster.sync();                                   // nobody leaves while a peer may still read its partial sums
}

template <int LANES>
__global__ void __launch_bounds__(256)
bn_act_bwd_cluster_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                          const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                          int act, const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dx, int lddx) {
    escgnn::pdl_enter();
    using T = ClusterTile<LANES>;
    __shared__ __align__(16) float s_w[8][2][T::kCols];
    __shared__ __align__(16) float s_part[2][T::kCols];
    __shared__ __align__(16) float s_tot[2][T::kCols];
    cg::cluster_group cluster = cg::this_cluster();
    T t{s_w, s_part, s_tot};
    const int cl = threadIdx.x % LANES, rs = threadIdx.x / LANES, rank = (int)cluster.block_rank();
    const int c4 = blockIdx.x * T::kCols + 4 * cl, rows = min(*d_rows, rows_cap);
    const bool col_ok = c4 < C;
    BnCols k;
    k.mu = k.rs = k.g = k.bt = zero4();
    float4 a = zero4(), b = zero4();
    if (col_ok) {
        k = bn_cols(mean, rstd, gamma, beta, c4);
        for (int r0 = rank * T::kSlots + rs; r0 < rows; r0 += kClBwdUnroll * T::kSweep) {
            float4 xv[kClBwdUnroll], dv[kClBwdUnroll];
            #pragma unroll
            for (int i = 0; i < kClBwdUnroll; ++i) {
                const int r = r0 + i * T::kSweep;
                xv[i] = zero4(); dv[i] = zero4();
                if (r < rows) {
                    xv[i] = ld4(x + (size_t)r * ldx + c4);
                    dv[i] = ld4(dy + (size_t)r * lddy + c4);
                    if (dy2) { const float4 e = ld4(dy2 + (size_t)r * lddy2 + c4); dv[i].x += e.x; dv[i].y += e.y; dv[i].z += e.z; dv[i].w += e.w; }
                }
            }
            #pragma unroll
            for (int i = 0; i < kClBwdUnroll; ++i)
                if (r0 + i * T::kSweep < rows) {
                    float4 dz, xh;
                    bn_dz(xv[i], dv[i], k, act, dz, xh);
                    a.x += dz.x; a.y += dz.y; a.z += dz.z; a.w += dz.w;
                    b.x += dz.x * xh.x; b.y += dz.y * xh.y; b.z += dz.z * xh.z; b.w += dz.w * xh.w;
                }
        }
    }
    t.reduce(a, b, cluster);
    if (col_ok) {
        const float4 s1 = ld4(&s_tot[0][4 * cl]), s2 = ld4(&s_tot[1][4 * cl]);
        if (rank == 0 && rs == 0) { if (dgamma) st4(dgamma + c4, s2); if (dbeta) st4(dbeta + c4, s1); }
        const float inv_m = 1.f / (float)max(rows, 1);
        const float4 m1 = make_float4(s1.x * inv_m, s1.y * inv_m, s1.z * inv_m, s1.w * inv_m);
        const float4 m2 = make_float4(s2.x * inv_m, s2.y * inv_m, s2.z * inv_m, s2.w * inv_m);
        const float4 kk = make_float4(k.g.x * k.rs.x, k.g.y * k.rs.y, k.g.z * k.rs.z, k.g.w * k.rs.w);
        for (int r0 = rank * T::kSlots + rs; r0 < rows_cap; r0 += kClBwdUnroll * T::kSweep) {
            float4 xv[kClBwdUnroll], dv[kClBwdUnroll];
            #pragma unroll
            for (int i = 0; i < kClBwdUnroll; ++i) {
                const int r = r0 + i * T::kSweep;
                xv[i] = zero4(); dv[i] = zero4();
                if (r < rows) {
                    xv[i] = ld4(x + (size_t)r * ldx + c4);
                    dv[i] = ld4(dy + (size_t)r * lddy + c4);
                    if (dy2) { const float4 e = ld4(dy2 + (size_t)r * lddy2 + c4); dv[i].x += e.x; dv[i].y += e.y; dv[i].z += e.z; dv[i].w += e.w; }
                }
            }
            #pragma unroll
            for (int i = 0; i < kClBwdUnroll; ++i) {
                const int r = r0 + i * T::kSweep;
                if (r >= rows_cap) break;
                float4 o = zero4();
                if (r < rows) {
                    float4 dz, xh;
                    bn_dz(xv[i], dv[i], k, act, dz, xh);
                    o = make_float4(kk.x * (dz.x - m1.x - xh.x * m2.x), kk.y * (dz.y - m1.y - xh.y * m2.y),
                                    kk.z * (dz.z - m1.z - xh.z * m2.z), kk.w * (dz.w - m1.w - xh.w * m2.w));
                }
                st4(dx + (size_t)r * lddx + c4, o);
            }
        }
    }
    cluster.sync();
}

// Register-resident variants: when the rows a thread owns fit in R float4 registers (rows_cap <= R * kSweep) the tile is
// read ONCE -- statistics and normalisation both come from registers -- and the kernel is one L2 round trip, one
// cluster exchange and the stores.
template <int LANES, int R>
__global__ void __launch_bounds__(256, 2)
bn_act_fwd_cluster_reg_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* running_mean, float* running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                              int act, float eps, float momentum, const int* __restrict__ d_rows, int rows_cap, int C,
                              float* __restrict__ y, int ldy) {
    bn_stamp(0);
    escgnn::pdl_enter();
    bn_stamp(1);
    using T = ClusterTile<LANES>;
    __shared__ __align__(16) float s_w[8][2][T::kCols];
    __shared__ __align__(16) float s_part[2][T::kCols];
    __shared__ __align__(16) float s_tot[2][T::kCols];
    cg::cluster_group cluster = cg::this_cluster();
    T t{s_w, s_part, s_tot};
    const int cl = threadIdx.x % LANES, rs = threadIdx.x / LANES, rank = (int)cluster.block_rank();
    const int c4 = blockIdx.x * T::kCols + 4 * cl;
    const bool col_ok = c4 < C;
    const int r_first = rank * T::kSlots + rs;
    float4 v[R];
    float4 a = zero4(), b = zero4();
    // every row below the CAPACITY is valid memory: issue the tile loads together with the load of the row count instead of
    // behind it (one L2 round trip less on the critical path), then mask
    #pragma unroll
    for (int k = 0; k < R; ++k) v[k] = (col_ok && r_first + k * T::kSweep < rows_cap) ? ld4(x + (size_t)(r_first + k * T::kSweep) * ldx + c4) : zero4();
    const float4 shift = col_ok ? ld4(x + c4) : zero4();          // row 0 is below the capacity: valid memory, used only when rows > 0
    const int rows = min(*d_rows, rows_cap);
    #pragma unroll
    for (int k = 0; k < R; ++k) v[k] = r_first + k * T::kSweep >= rows ? zero4() : sub4(v[k], shift);     // registers hold x - shift
    #pragma unroll
    for (int k = 0; k < R; ++k) {
        a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
        b.x += v[k].x * v[k].x; b.y += v[k].y * v[k].y; b.z += v[k].z * v[k].z; b.w += v[k].w * v[k].w;
    }
    bn_stamp(2);
    t.reduce(a, b, cluster);
    bn_stamp(3);
    cluster.barrier_arrive();                         // this CTA is done reading its peers' shared memory ...
    if (col_ok) {
        const float m = (float)max(rows, 1);
        const float4 s1 = ld4(&s_tot[0][4 * cl]), s2 = ld4(&s_tot[1][4 * cl]);
        float4 mean, var;
        bn_finish4(s1, s2, shift, m, mean, var);
        const float4 m1c = make_float4(s1.x / m, s1.y / m, s1.z / m, s1.w / m);      // mean - shift (the registers hold x - shift)
        const float4 rstd = make_float4(rsqrtf(var.x + eps), rsqrtf(var.y + eps), rsqrtf(var.z + eps), rsqrtf(var.w + eps));
        if (rank == 0 && rs == 0) {
            st4(mean_out + c4, mean); st4(rstd_out + c4, rstd);
            if (rows > 0) {
                const float ub = rows > 1 ? m / (m - 1.f) : 1.f, k0 = 1.f - momentum;
                const float4 rm = ld4(running_mean + c4), rv = ld4(running_var + c4);
                st4(running_mean + c4, make_float4(k0 * rm.x + momentum * mean.x, k0 * rm.y + momentum * mean.y,
                                                   k0 * rm.z + momentum * mean.z, k0 * rm.w + momentum * mean.w));
                st4(running_var + c4, make_float4(k0 * rv.x + momentum * var.x * ub, k0 * rv.y + momentum * var.y * ub,
                                                  k0 * rv.z + momentum * var.z * ub, k0 * rv.w + momentum * var.w * ub));
            }
        }
        const float4 g = gamma ? ld4(gamma + c4) : make_float4(1.f, 1.f, 1.f, 1.f), bt = beta ? ld4(beta + c4) : zero4();
        const float4 sc = make_float4(rstd.x * g.x, rstd.y * g.y, rstd.z * g.z, rstd.w * g.w);
        #pragma unroll
        for (int k = 0; k < R; ++k) {
            const int r = r_first + k * T::kSweep;
            if (r < rows_cap) {
                float4 o = zero4();                   // rows >= actual count are zeroed (inert in the next GEMM)
                if (r < rows) o = act_fwd4(bn_apply4(v[k], m1c, sc, bt), act);      // v = x - shift, m1c = mean - shift
                st4(y + (size_t)r * ldy + c4, o);
            }
        }
    }
    bn_stamp(4);
    cluster.barrier_wait();                           // ... and leaves only when every peer is done reading its own
    bn_stamp(5);
}

template <int LANES, int R>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_cluster_reg_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                              const float* __restrict__ dy2, int lddy2, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                              int act, const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, float* __restrict__ dx, int lddx) {
    bn_stamp(0);
    escgnn::pdl_enter();
    bn_stamp(1);
    using T = ClusterTile<LANES>;
    __shared__ __align__(16) float s_w[8][2][T::kCols];
    __shared__ __align__(16) float s_part[2][T::kCols];
    __shared__ __align__(16) float s_tot[2][T::kCols];
    cg::cluster_group cluster = cg::this_cluster();
    T t{s_w, s_part, s_tot};
    const int cl = threadIdx.x % LANES, rs = threadIdx.x / LANES, rank = (int)cluster.block_rank();
    const int c4 = blockIdx.x * T::kCols + 4 * cl;
    const bool col_ok = c4 < C;
    const int r_first = rank * T::kSlots + rs;
    BnCols k;
    k.mu = k.rs = k.g = k.bt = zero4();
    if (col_ok) k = bn_cols(mean, rstd, gamma, beta, c4);
    float4 dz[R], xh[R];                              // after the loads: dz and xhat of this thread's rows
    #pragma unroll
    for (int i = 0; i < R; ++i) {                     // loads bounded by the capacity, issued alongside the row count (see forward)
        const int r = r_first + i * T::kSweep;
        dz[i] = zero4(); xh[i] = zero4();
        if (col_ok && r < rows_cap) {
            xh[i] = ld4(x + (size_t)r * ldx + c4);
            dz[i] = ld4(dy + (size_t)r * lddy + c4);
            if (dy2) { const float4 e = ld4(dy2 + (size_t)r * lddy2 + c4); dz[i].x += e.x; dz[i].y += e.y; dz[i].z += e.z; dz[i].w += e.w; }
        }
    }
    const int rows = min(*d_rows, rows_cap);
    float4 a = zero4(), b = zero4();
    #pragma unroll
    for (int i = 0; i < R; ++i)
        if (col_ok && r_first + i * T::kSweep < rows) {
            float4 z, h;
            bn_dz(xh[i], dz[i], k, act, z, h);
            dz[i] = z; xh[i] = h;
            a.x += z.x; a.y += z.y; a.z += z.z; a.w += z.w;
            b.x += z.x * h.x; b.y += z.y * h.y; b.z += z.z * h.z; b.w += z.w * h.w;
        }
    bn_stamp(2);
    t.reduce(a, b, cluster);
    bn_stamp(3);
    cluster.barrier_arrive();
    if (col_ok) {
        const float4 s1 = ld4(&s_tot[0][4 * cl]), s2 = ld4(&s_tot[1][4 * cl]);
        if (rank == 0 && rs == 0) { if (dgamma) st4(dgamma + c4, s2); if (dbeta) st4(dbeta + c4, s1); }
        const float inv_m = 1.f / (float)max(rows, 1);
        const float4 m1 = make_float4(s1.x * inv_m, s1.y * inv_m, s1.z * inv_m, s1.w * inv_m);
        const float4 m2 = make_float4(s2.x * inv_m, s2.y * inv_m, s2.z * inv_m, s2.w * inv_m);
        const float4 kk = make_float4(k.g.x * k.rs.x, k.g.y * k.rs.y, k.g.z * k.rs.z, k.g.w * k.rs.w);
        #pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = r_first + i * T::kSweep;
            if (r < rows_cap) {
                float4 o = zero4();
                if (r < rows)
                    o = make_float4(kk.x * (dz[i].x - m1.x - xh[i].x * m2.x), kk.y * (dz[i].y - m1.y - xh[i].y * m2.y),
                                    kk.z * (dz[i].z - m1.z - xh[i].z * m2.z), kk.w * (dz[i].w - m1.w - xh[i].w * m2.w));
                st4(dx + (size_t)r * lddx + c4, o);
            }
        }
    }
    bn_stamp(4);
    cluster.barrier_wait();
    bn_stamp(5);
}

// (lanes, registers) plan of the one-launch BatchNorm: as many clusters as the column count allows, rows in registers when they fit
template <class K2_8, class K2_16, class KS2, class KS4, class KS8, class... Args>
inline void launch_cluster_bn(bool backward, int rows_cap, int channels, cudaStream_t st, K2_8 k2_8, K2_16 k2_16, KS2 ks2,
                              KS4 ks4, KS8 ks8, Args... args) {
    auto go = [&](auto kern, int lanes) {
        escgnn::launch_pdl_cluster(kern, dim3((unsigned)((channels + 4 * lanes - 1) / (4 * lanes)), kClRanks), 256, 0, st, kClRanks, args...);
    };
    if (channels <= 512) {
        if (rows_cap <= 8 * 1024) return go(k2_8, 2);                       // lanes 2: 1024 rows per sweep
        if (!backward && rows_cap <= 16 * 1024) return go(k2_16, 2);
        return go(ks2, 2);
    }
    if (channels <= 1024) return go(ks4, 4);
    return go(ks8, 8);
}

inline int& cluster_bn_enabled() {
    static int on = 1;
    return on;
}
inline bool vec_ok(int C, std::initializer_list<const void*> ptrs, std::initializer_list<int> lds) {
    if (C % 4) return false;
    for (const void* p : ptrs) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return false;
    for (int l : lds) if (l % 4) return false;
    return true;
}
constexpr int kMaxTiles = 1024;   // reduction tiles per column block (more rows -> several 64-row chunks per CTA): bounds the last tile's sum
inline dim3 reduce_grid(int rows_cap, int C) {
    const int t = (rows_cap + kVRows - 1) / kVRows;
    return dim3((unsigned)((C + kVCols - 1) / kVCols), (unsigned)(t < kMaxTiles ? (t < 1 ? 1 : t) : kMaxTiles));
}
inline dim3 vec_grid(int rows_cap, int C) { return dim3((unsigned)((C + kVCols - 1) / kVCols), (unsigned)((rows_cap + kVRows - 1) / kVRows)); }

// activation only (no BatchNorm): y = act(x), rows beyond the count zeroed; backward: dx = dy * act'(x)
__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ x, int ldx, int act, const int* __restrict__ d_rows, int rows_cap, int C,
               float* __restrict__ y, int ldy) {
    escgnn::pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    if (c >= C) return;
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8) y[(size_t)r * ldy + c] = r < rows ? act_fwd(x[(size_t)r * ldx + c], act) : 0.f;
}
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy, int act,
               const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ dx, int lddx) {
    escgnn::pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = min(*d_rows, rows_cap);
    if (c >= C) return;
    for (int r = r0 + warp; r < min(r0 + kTileRows, rows_cap); r += 8)
        dx[(size_t)r * lddx + c] = r < rows ? dy[(size_t)r * lddy + c] * act_grad(x[(size_t)r * ldx + c], act) : 0.f;
}

// inverted dropout with a counter-based mask: keep(r, c) = hash(salt, step, r * C + c) >= p, y = keep ? x / (1 - p) : 0.  The mask
// is a pure function of (salt, step counter, position), so the backward pass applies the SAME kernel to the gradient and nothing
// is stored; `step` lives in device memory (the optimiser's counter), which keeps the captured graph valid from step to step.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, int ldx, float p, uint32_t salt, const long long* __restrict__ d_step,
               const int* __restrict__ d_rows, int rows_cap, int C, float* __restrict__ y, int ldy) {
    escgnn::pdl_enter();
    const int rows = min(*d_rows, rows_cap);
    const uint32_t step = d_step ? (uint32_t)*d_step : 0u;
    const uint32_t key = mix32(salt * 0x9e3779b9u + step);
    const uint32_t thresh = p >= 1.f ? 0xffffffffu : (uint32_t)((double)p * 4294967296.0);
    const float scale = p >= 1.f ? 0.f : 1.f / (1.f - p);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows_cap * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        float o = 0.f;
        if (r < rows && mix32((uint32_t)i ^ key) >= thresh) o = x[(size_t)r * ldx + c] * scale;
        y[(size_t)r * ldy + c] = o;
    }
}

// column sums (bias gradients): partial per tile, then an ordered final sum
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ d_rows, int C, float* __restrict__ partial) {
    escgnn::pdl_enter();
    __shared__ float s[8][2][kCols];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kCols + lane, r0 = blockIdx.y * kTileRows, rows = *d_rows;
    float a = 0.f, b = 0.f;
    if (c < C) for (int r = r0 + warp; r < min(r0 + kTileRows, rows); r += 8) a += x[(size_t)r * ldx + c];
    cta_reduce2(a, b, s);
    if (warp == 0 && c < C) partial[(size_t)blockIdx.y * C + c] = a;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, const int* __restrict__ d_rows, int C, float* __restrict__ out) {
    escgnn::pdl_enter();
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
    escgnn::pdl_enter();
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
    escgnn::pdl_enter();
    const int rows = *d_rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        const float g = dy[(size_t)r * lddy + c];
        for (int k = 0; k < n_cols_idx; ++k)
            atomicAdd(&dtable[(size_t)(idx[(size_t)r * n_cols_idx + k] + (col_offsets ? col_offsets[k] : 0)) * C + c], g);
    }
}

// small tables (BondEncoder: 13 rows, the type embeddings: 100): thousands of rows hammer a handful of addresses, so every CTA
// accumulates into a private copy of the table in shared memory and adds it to the global gradient once
__global__ void __launch_bounds__(256)
embedding_bwd_small_kernel(const float* __restrict__ dy, int lddy, const int64_t* __restrict__ idx, int n_cols_idx,
                           const int64_t* __restrict__ col_offsets, const int* __restrict__ d_rows, int C, int table_rows,
                           float* __restrict__ dtable) {
    escgnn::pdl_enter();
    extern __shared__ float s_tab[];
    const int total = table_rows * C;
    for (int i = threadIdx.x; i < total; i += blockDim.x) s_tab[i] = 0.f;
    __syncthreads();
    const int rows = *d_rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)rows * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        const float g = dy[(size_t)r * lddy + c];
        for (int k = 0; k < n_cols_idx; ++k) {
            const int64_t t = idx[(size_t)r * n_cols_idx + k] + (col_offsets ? col_offsets[k] : 0);
            if (t >= 0 && t < table_rows) atomicAdd(&s_tab[(int)t * C + c], g);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const float v = s_tab[i];
        if (v != 0.f) atomicAdd(&dtable[i], v);
    }
}

// losses over `rows` predictions [rows, T]: kind 0 = L1 mean (run_graphcount.py:498, run_zinc.py:283),
// kind 1 = BCE-with-logits mean over labelled entries y == y (run_ogb_mol.py:58-74). Single CTA, ordered sum.
__global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ target, int kind, const int* __restrict__ d_rows,
            int T, float* __restrict__ loss, float* __restrict__ dpred, int lddp, int rows_cap) {
    escgnn::pdl_enter();
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

// ---------------------------------------------------------------------------------------------------------------
// Graph-level readout tail in ONE launch (zinc_models.py:604-609 and the loss / its backward, run_zinc.py:276-281):
//   p2 = act(BN(x));  pred = p2 w2 + b2;  loss = mean |pred - y|;  and straight away the backward of all of it:
//   d pred, d w2, d b2, d p2, BatchNorm backward -> dx, d gamma, d beta.
// A few hundred rows x <= 256 columns: five dependent launches (BN, GEMV, loss, GEMV^T, BN backward) of ~6 us each for
// microseconds of work.  One 8-CTA cluster: a thread owns one column of RPC rows (registers), the two column reductions
// (statistics; dz sums + d w2) cross the cluster through distributed shared memory, the row dot products stay inside a CTA.
template <int RPC>
__global__ void __launch_bounds__(256)
head_bn_linear_l1_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* running_mean, float* running_var, int act, float eps, float momentum,
                         const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ target,
                         const int* __restrict__ d_rows, int rows_cap, int H, float* __restrict__ pred, float* __restrict__ loss,
                         float* __restrict__ dx, int lddx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dw2, float* __restrict__ db2) {
    escgnn::pdl_enter();
    __shared__ float s_stat[2][256], s_back[3][256], s_row[8][RPC], s_dpred[RPC], s_scal[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5, rank = (int)cluster.block_rank(), r_base = rank * RPC;
    const bool col_ok = c < H;
    float xv[RPC];
    #pragma unroll
    for (int i = 0; i < RPC; ++i) xv[i] = (col_ok && r_base + i < rows_cap) ? x[(size_t)(r_base + i) * ldx + c] : 0.f;
    const int rows = min(*d_rows, rows_cap);
    const float m = (float)max(rows, 1);
    float s1 = 0.f, s2 = 0.f;
    const float shift = (col_ok && rows > 0) ? x[c] : 0.f;         // shifted sums (see the header): xv holds x - x[0]
    #pragma unroll
    for (int i = 0; i < RPC; ++i) {
        xv[i] = r_base + i >= rows ? 0.f : xv[i] - shift;
        s1 += xv[i]; s2 += xv[i] * xv[i];
    }
    s_stat[0][c] = s1; s_stat[1][c] = s2;
    cluster.sync();
    float t1 = 0.f, t2 = 0.f;
    #pragma unroll
    for (int r = 0; r < kClRanks; ++r) { t1 += *cluster.map_shared_rank(&s_stat[0][c], r); t2 += *cluster.map_shared_rank(&s_stat[1][c], r); }
    const float mean = t1 / m, var = fmaxf(t2 / m - mean * mean, 0.f), rstd = rsqrtf(var + eps);      // mean of x - shift
    if (rank == 0 && col_ok && rows > 0) {
        const float ub = rows > 1 ? m / (m - 1.f) : 1.f;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + shift);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * ub;
    }
    const float g = col_ok ? (gamma ? gamma[c] : 1.f) : 0.f, bt = col_ok ? (beta ? beta[c] : 0.f) : 0.f, w = col_ok ? w2[c] : 0.f;
    // forward: xv becomes xhat, pv the activation (ONE transcendental per element: act' follows from the activation value);
    // row dot products pred[r] = sum_c p2[r][c] w2[c]
    float pv[RPC];
    #pragma unroll
    for (int i = 0; i < RPC; ++i) {
        xv[i] = (xv[i] - mean) * rstd;
        pv[i] = (col_ok && r_base + i < rows) ? act_fwd(xv[i] * g + bt, act) : 0.f;
    }
    #pragma unroll
    for (int i = 0; i < RPC; ++i) {
        float v = pv[i] * w;
        #pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
        if (lane == 0) s_row[warp][i] = v;
    }
    __syncthreads();
    if (c < RPC) {
        const int r = r_base + c;
        float p = b2[0];
        #pragma unroll
        for (int wv = 0; wv < 8; ++wv) p += s_row[wv][c];
        float dp = 0.f, l = 0.f;
        if (r < rows) {
            const float diff = p - target[r];
            l = fabsf(diff);
            dp = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) / m;
        }
        if (r < rows_cap) pred[r] = r < rows ? p : 0.f;
        s_dpred[c] = dp;
        s_row[0][c] = l;                                        // this thread's own slot: reused for the loss terms
    }
    __syncthreads();
    if (c == 0) {
        float ls = 0.f, ds = 0.f;
        for (int i = 0; i < RPC; ++i) { ls += s_row[0][i]; ds += s_dpred[i]; }
        s_scal[0] = ls; s_scal[1] = ds;
    }
    // backward: dz = d p2 * act'(z), d p2 = d pred (x) w2
    // act'(z) from the activation value: ReLU / ELU are positive exactly where z is, and elu'(z) = elu(z) + 1 for z <= 0
    auto slope = [act](float pval) { return act == 0 ? 1.f : (pval > 0.f ? 1.f : (act == 2 ? pval + 1.f : 0.f)); };
    float a = 0.f, b = 0.f, dw = 0.f;
    #pragma unroll
    for (int i = 0; i < RPC; ++i) {
        if (col_ok && r_base + i < rows) {
            const float dp = s_dpred[i];
            dw += dp * pv[i];
            const float dz = dp * w * slope(pv[i]);
            a += dz; b += dz * xv[i];
        }
    }
    s_back[0][c] = a; s_back[1][c] = b; s_back[2][c] = dw;
    cluster.sync();
    float A = 0.f, B = 0.f, DW = 0.f;
    #pragma unroll
    for (int r = 0; r < kClRanks; ++r) {
        A += *cluster.map_shared_rank(&s_back[0][c], r); B += *cluster.map_shared_rank(&s_back[1][c], r);
        DW += *cluster.map_shared_rank(&s_back[2][c], r);
    }
    if (rank == 0) {
        if (col_ok) { if (dgamma) dgamma[c] = B; if (dbeta) dbeta[c] = A; dw2[c] = DW; }
        if (c == 0) {
            float ls = 0.f, ds = 0.f;
            for (int r = 0; r < kClRanks; ++r) { ls += *cluster.map_shared_rank(&s_scal[0], r); ds += *cluster.map_shared_rank(&s_scal[1], r); }
            loss[0] = ls / m; db2[0] = ds;
        }
    }
    cluster.barrier_arrive();
    if (col_ok) {
        const float k = g * rstd, m1 = A / m, m2 = B / m;
        #pragma unroll
        for (int i = 0; i < RPC; ++i) {
            const int r = r_base + i;
            if (r < rows_cap) {
                float o = 0.f;
                if (r < rows) {
                    const float dz = s_dpred[i] * w * slope(pv[i]);
                    o = k * (dz - m1 - xv[i] * m2);
                }
                dx[(size_t)r * lddx + c] = o;
            }
        }
    }
    cluster.barrier_wait();
}

inline dim3 tile_grid(int rows_cap, int C) { return dim3((unsigned)((C + kCols - 1) / kCols), (unsigned)((rows_cap + kTileRows - 1) / kTileRows)); }

}  // namespace

extern "C" {

int escgnn_dense_tile_rows(void) { return kTileRows; }

int escgnn_head_bn_linear_l1(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float* d_running_mean,
                             float* d_running_var, int act, float eps, float momentum, const float* d_w2, const float* d_b2,
                             const float* d_target, const int* d_rows, int rows_cap, int channels, float* d_pred, float* d_loss,
                             float* d_dx, int lddx, float* d_dgamma, float* d_dbeta, float* d_dw2, float* d_db2, void* stream) {
    if (channels > 256 || rows_cap > kClRanks * 64 || rows_cap < 1) return ESCGNN_ERR_TOO_LARGE;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows_cap <= kClRanks * 32)
        escgnn::launch_pdl_cluster(head_bn_linear_l1_kernel<32>, dim3(1, kClRanks), 256, 0, st, kClRanks, d_x, ldx, d_gamma, d_beta,
                                   d_running_mean, d_running_var, act, eps, momentum, d_w2, d_b2, d_target, d_rows, rows_cap, channels,
                                   d_pred, d_loss, d_dx, lddx, d_dgamma, d_dbeta, d_dw2, d_db2);
    else
        escgnn::launch_pdl_cluster(head_bn_linear_l1_kernel<64>, dim3(1, kClRanks), 256, 0, st, kClRanks, d_x, ldx, d_gamma, d_beta,
                                   d_running_mean, d_running_var, act, eps, momentum, d_w2, d_b2, d_target, d_rows, rows_cap, channels,
                                   d_pred, d_loss, d_dx, lddx, d_dgamma, d_dbeta, d_dw2, d_db2);
    return (int)cudaGetLastError();
}

int escgnn_bn_set_trace(unsigned long long* d_stamps) {
    return (int)cudaMemcpyToSymbol(g_bn_trace, &d_stamps, sizeof(d_stamps));
}

int escgnn_set_cluster_bn(int on) {
    const int was = cluster_bn_enabled();
    cluster_bn_enabled() = on ? 1 : 0;
    return was;
}

int escgnn_set_pdl(int on) {
    const int was = escgnn::pdl_enabled();
    escgnn::pdl_enabled() = on ? 1 : 0;
    return was;
}

int64_t escgnn_dense_partial_floats(int rows_cap, int channels) {
    const int64_t c = (channels + 3) / 4 * 4;
    return kHdr + 2 * c + ((int64_t)(rows_cap + kVRows - 1) / kVRows + 1) * 2 * c;
}

int escgnn_bn_act_fwd(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float* d_running_mean,
                      float* d_running_var, float* d_mean, float* d_rstd, float* d_partial, int act, float eps,
                      float momentum, int training, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (vec_ok(channels, {d_x, d_y, d_gamma, d_beta, d_running_mean, d_running_var, d_mean, d_rstd, d_partial}, {ldx, ldy})) {
        if (training && rows_cap <= kClMaxRows && cluster_bn_enabled()) {       // one launch: cluster-wide statistics
            launch_cluster_bn(false, rows_cap, channels, st, bn_act_fwd_cluster_reg_kernel<2, 8>, bn_act_fwd_cluster_reg_kernel<2, 16>,
                              bn_act_fwd_cluster_kernel<2>, bn_act_fwd_cluster_kernel<4>,
                              bn_act_fwd_cluster_kernel<8>, d_x, ldx, d_gamma, d_beta, d_running_mean, d_running_var, d_mean, d_rstd, act,
                              eps, momentum, d_rows, rows_cap, channels, d_y, ldy);
            return (int)cudaGetLastError();
        }
        const dim3 g = vec_grid(rows_cap, channels);
        if (training)
            escgnn::launch_pdl(colstats_v4_kernel<0>, reduce_grid(rows_cap, channels), 256, 0, st, d_x, ldx, d_rows, rows_cap, channels, d_partial, d_running_mean, d_running_var,
                                                     d_mean, d_rstd, eps, momentum, nullptr);
        escgnn::launch_pdl(bn_act_fwd_v4_kernel, g, 256, 0, st, d_x, ldx, d_gamma, d_beta, d_running_mean, d_running_var, d_mean, d_rstd, act, eps,
                                                training, d_rows, rows_cap, channels, d_y, ldy);
        return (int)cudaGetLastError();
    }
    d_partial += kHdr;                                       // scalar fallback: plain [tile][2][C] partials behind the tickets
    const dim3 g = tile_grid(rows_cap, channels);
    if (training) escgnn::launch_pdl(colstats_kernel, g, 256, 0, st, d_x, ldx, d_rows, channels, d_partial);
    escgnn::launch_pdl(bn_act_fwd_kernel, g, 256, 0, st, d_x, ldx, d_partial, d_gamma, d_beta, d_running_mean, d_running_var, d_mean, d_rstd,
                                         act, eps, momentum, training, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_bn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, const float* d_dy2, int lddy2,
                      const float* d_mean, const float* d_rstd, const float* d_gamma, const float* d_beta, int act,
                      int training, float* d_partial, const int* d_rows, int rows_cap, int channels, float* d_dgamma,
                      float* d_dbeta, float* d_dx, int lddx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (vec_ok(channels, {d_x, d_dy, d_dy2, d_mean, d_rstd, d_gamma, d_beta, d_partial, d_dx}, {ldx, lddy, d_dy2 ? lddy2 : 0, lddx})) {
        if (training && rows_cap <= 8 * 1024 && cluster_bn_enabled()) {      // larger tiles do not fit the registers: the two-kernel path below is faster (19.8 vs 23.3 us at 12 k rows)
            launch_cluster_bn(true, rows_cap, channels, st, bn_act_bwd_cluster_reg_kernel<2, 8>, bn_act_bwd_cluster_reg_kernel<2, 8>,
                              bn_act_bwd_cluster_kernel<2>, bn_act_bwd_cluster_kernel<4>,
                              bn_act_bwd_cluster_kernel<8>, d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act, d_rows,
                              rows_cap, channels, d_dgamma, d_dbeta, d_dx, lddx);
            return (int)cudaGetLastError();
        }
        const dim3 g = vec_grid(rows_cap, channels);
        escgnn::launch_pdl(bn_act_bwd_reduce_v4_kernel, reduce_grid(rows_cap, channels), 256, 0, st, d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                                       d_rows, rows_cap, channels, d_partial, d_dgamma, d_dbeta);
        escgnn::launch_pdl(bn_act_bwd_apply_v4_kernel, g, 256, 0, st, d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                                      training, d_partial, d_rows, rows_cap, channels, d_dx, lddx);
        return (int)cudaGetLastError();
    }
    d_partial += kHdr;
    const dim3 g = tile_grid(rows_cap, channels);
    escgnn::launch_pdl(bn_act_bwd_reduce_kernel, g, 256, 0, st, d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                                d_rows, channels, d_partial);
    escgnn::launch_pdl(bn_act_bwd_apply_kernel, g, 256, 0, st, d_x, ldx, d_dy, lddy, d_dy2, lddy2, d_mean, d_rstd, d_gamma, d_beta, act,
                                               training, d_partial, d_rows, rows_cap, channels, d_dgamma, d_dbeta, d_dx, lddx);
    return (int)cudaGetLastError();
}

int escgnn_dropout(const float* d_x, int ldx, float p, uint32_t salt, const long long* d_step, const int* d_rows, int rows_cap,
                   int channels, float* d_y, int ldy, void* stream) {
    int64_t total = (int64_t)rows_cap * channels;
    if (total <= 0) return 0;
    unsigned b = (unsigned)((total + 255) / 256); if (b > 148 * 16) b = 148 * 16;
    escgnn::launch_pdl(dropout_kernel, b, 256, 0, (cudaStream_t)stream, d_x, ldx, p, salt, d_step, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_act_fwd(const float* d_x, int ldx, int act, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                   void* stream) {
    escgnn::launch_pdl(act_fwd_kernel, tile_grid(rows_cap, channels), 256, 0, (cudaStream_t)stream, d_x, ldx, act, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, int act, const int* d_rows, int rows_cap,
                   int channels, float* d_dx, int lddx, void* stream) {
    escgnn::launch_pdl(act_bwd_kernel, tile_grid(rows_cap, channels), 256, 0, (cudaStream_t)stream, d_x, ldx, d_dy, lddy, act, d_rows, rows_cap, channels, d_dx, lddx);
    return (int)cudaGetLastError();
}

int escgnn_colsum(const float* d_x, int ldx, const int* d_rows, int rows_cap, int channels, float* d_partial, float* d_out,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (vec_ok(channels, {d_x, d_partial}, {ldx})) {         // one launch: the last tile to arrive writes the sums
        escgnn::launch_pdl(colstats_v4_kernel<1>, reduce_grid(rows_cap, channels), 256, 0, st, d_x, ldx, d_rows, rows_cap, channels, d_partial, nullptr,
                                                                           nullptr, nullptr, nullptr, 0.f, 0.f, d_out);
        return (int)cudaGetLastError();
    }
    d_partial += kHdr;
    escgnn::launch_pdl(colsum_partial_kernel, tile_grid(rows_cap, channels), 256, 0, st, d_x, ldx, d_rows, channels, d_partial);
    escgnn::launch_pdl(colsum_final_kernel, (unsigned)((channels + 127) / 128), 128, 0, st, d_partial, d_rows, channels, d_out);
    return (int)cudaGetLastError();
}

int escgnn_embedding_fwd(const float* d_table, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_y, int ldy, void* stream) {
    int64_t total = (int64_t)rows_cap * channels;
    unsigned b = (unsigned)((total + 255) / 256); if (b > 1184) b = 1184; if (b < 1) b = 1;
    escgnn::launch_pdl(embedding_fwd_kernel, b, 256, 0, (cudaStream_t)stream, d_table, d_idx, idx_cols, d_col_offsets, d_rows, rows_cap, channels, d_y, ldy);
    return (int)cudaGetLastError();
}

int escgnn_embedding_bwd(const float* d_dy, int lddy, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_dtable, void* stream) {
    int64_t total = (int64_t)rows_cap * channels;
    unsigned b = (unsigned)((total + 255) / 256); if (b > 1184) b = 1184; if (b < 1) b = 1;
    escgnn::launch_pdl(embedding_bwd_kernel, b, 256, 0, (cudaStream_t)stream, d_dy, lddy, d_idx, idx_cols, d_col_offsets, d_rows, channels, d_dtable);
    return (int)cudaGetLastError();
}

int escgnn_embedding_bwd_small(const float* d_dy, int lddy, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                               const int* d_rows, int rows_cap, int channels, int table_rows, float* d_dtable, void* stream) {
    const size_t smem = (size_t)table_rows * channels * sizeof(float);
    if (smem > 48 * 1024)
        return escgnn_embedding_bwd(d_dy, lddy, d_idx, idx_cols, d_col_offsets, d_rows, rows_cap, channels, d_dtable, stream);
    int64_t total = (int64_t)rows_cap * channels;
    unsigned b = (unsigned)((total + 256 * 16 - 1) / (256 * 16)); if (b > 296) b = 296; if (b < 1) b = 1;
    escgnn::launch_pdl(embedding_bwd_small_kernel, b, 256, smem, (cudaStream_t)stream, d_dy, lddy, d_idx, idx_cols, d_col_offsets, d_rows,
                       channels, table_rows, d_dtable);
    return (int)cudaGetLastError();
}

int escgnn_loss_fwd_bwd(const float* d_pred, int ldp, const float* d_target, int kind, const int* d_rows, int rows_cap,
                        int n_targets, float* d_loss, float* d_dpred, int lddp, void* stream) {
    escgnn::launch_pdl(loss_kernel, 1, 1024, 0, (cudaStream_t)stream, d_pred, ldp, d_target, kind, d_rows, n_targets, d_loss, d_dpred, lddp, rows_cap);
    return (int)cudaGetLastError();
}

}  // extern "C"
