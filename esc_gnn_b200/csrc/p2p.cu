// Data-parallel exchange of the train step as ONE kernel over NVLink peer memory: gradient
// reduce-scatter + Adam on the owned slice + all-gather of the updated parameters, with the cross-GPU barriers inside the
// kernel.  Replaces "NCCL all-reduce between two captured graphs, then an Adam launch" -- the exchange becomes a node of the
// step's single CUDA graph (the reference itself is single-GPU: run_zinc.py:266-289 steps torch.optim.Adam on one device).
//
//   every rank r owns slice r of the flat parameter vector (n / world elements):
//   (escgnn_allreduce_adam_range splits the vector into buckets that are exchanged independently: the engine sends everything but
//   the parameters whose gradients arrive last on a side branch, under the tail of the backward pass.)
//     phase 0  announce "my gradients are complete" to every peer (release store of this step's epoch into the peer's flag
//              word), wait until every peer has announced the same epoch (acquire spin on local memory);
//     phase 1  for the owned slice: g = sum over ranks (fixed order -> bit-identical replicas) of the peers' gradient slices,
//              read straight from their HBM over NVLink; Adam update of m, v (kept for the owned slice only, ZeRO-1 style)
//              and p; the new p is stored into EVERY rank's parameter buffer;
//     phase 2  fence, then the last CTA announces "done" to every peer and waits for theirs: on return every peer has finished
//              reading this rank's gradients (they may be zeroed for the next step) and all slices of the new parameters
//              have landed here.
// Traffic per rank: (world-1)/world * n reads + the same in writes (6.2 MB each way for the 1.77 M-parameter ZINC model
// on 8 GPUs), against 2 * (world-1)/world * n for a ring all-reduce PLUS a separate read-modify-write Adam pass.
// Buffers that peers touch (gradients, parameters, flags) are cudaMalloc allocations shared through CUDA IPC handles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>

#include "../../include/escgnn_b200.h"
#include "launch.cuh"

namespace {

constexpr int kMaxWorld = 16;
constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;      // a peer that never arrives must not hang the GPU

struct Peers {
    const float* grad[kMaxWorld];
    float* param[kMaxWorld];
    unsigned long long* flags[kMaxWorld];     // per rank: [0, world) ready epochs, [world, 2 world) done epochs, [2 world] epoch,
};                                            // [2 world + 1] CTA ticket, [2 world + 2] error word

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// wait until *p >= epoch; on timeout record an error and give up (the step's result is then wrong, but the GPU stays alive)
__device__ __forceinline__ void spin_until(const unsigned long long* p, unsigned long long epoch, unsigned long long* err) {
    const unsigned long long t0 = now_ns();
    while (ld_acquire_sys(p) < epoch) {
        __nanosleep(64);
        if (now_ns() - t0 > kSpinTimeoutNs) { atomicExch(err, 1ull); break; }
    }
}

// One bucket [lo4, hi4) (float4 units) of the flat vector.  `tick` != 0: this launch opens a new optimiser step -- every CTA derives
// t = state + 1 and the bias corrections itself, and the last CTA to finish publishes them (state, hyper[5..6]) for the launches of
// the same step that follow (tick == 0: they read what the first one stored).  Each bucket has its own flag block (epoch counter
// included), so a bucket whose gradients are complete early can be exchanged on a side branch while the backward pass continues.
__global__ void __launch_bounds__(256)
allreduce_adam_kernel(const Peers peers, int rank, int world, int64_t lo4, int64_t hi4, int64_t flag_base, int tick, float* __restrict__ m,
                      float* __restrict__ v, float* hyper, long long* state) {
    escgnn::pdl_enter();
    unsigned long long* mine = peers.flags[rank] + flag_base;
    unsigned long long* err = mine + 2 * world + 2;
    const unsigned long long epoch = mine[2 * world] + 1ull;          // the last CTA stores it back in phase 2
    __shared__ int s_last;
    // ---- phase 0: gradients of every rank complete
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + flag_base + rank, epoch);
    if (threadIdx.x < world) spin_until(mine + threadIdx.x, epoch, err);
    __syncthreads();
    // ---- phase 1: reduce the owned slice of the bucket, Adam, broadcast the new parameters
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], gs = hyper[4];
    float bc1 = hyper[5], bc2s = hyper[6];
    const long long t_now = state[0] + (tick ? 1 : 0);
    if (tick) {
        bc1 = (float)(1.0 - pow((double)b1, (double)t_now));
        bc2s = (float)sqrt(1.0 - pow((double)b2, (double)t_now));
    }
    const float step = lr / bc1;
    const int64_t n4 = hi4 - lo4, slice4 = (n4 + world - 1) / world;
    const int64_t lo = lo4 + (int64_t)rank * slice4, hi = min(lo + slice4, hi4);
    for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        #pragma unroll 4
        for (int p = 0; p < world; ++p) {                             // fixed order: every slice is summed by exactly one rank
            const float4 t = __ldcg(reinterpret_cast<const float4*>(peers.grad[p]) + i);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        float4 pv = reinterpret_cast<float4*>(peers.param[rank])[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pp = &pv.x; float* mm = &mv.x; float* vp = &vv.x; const float* gg = &g.x;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * gs;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vp[k] = b2 * vp[k] + (1.f - b2) * gk * gk;
            pp[k] -= step * mm[k] / (sqrtf(vp[k]) / bc2s + eps);
        }
        reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
        for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(peers.param[p])[i] = pv;
    }
    // ---- phase 2: all of this rank's reads and remote stores are done -> tell the peers, wait for theirs
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(mine + 2 * world + 1, 1ull) == (unsigned long long)gridDim.x - 1ull;
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + flag_base + world + rank, epoch);
        if (threadIdx.x < world) spin_until(mine + world + threadIdx.x, epoch, err);
        if (threadIdx.x == 0) {
            mine[2 * world + 1] = 0ull;                              // ticket re-armed for the next launch
            mine[2 * world] = epoch;
            if (tick) { state[0] = t_now; hyper[5] = bc1; hyper[6] = bc2s; }
        }
    }
}

}  // namespace

extern "C" {

int escgnn_p2p_alloc(int64_t bytes, void** d_ptr, unsigned char* handle64) {
    if (bytes <= 0 || !d_ptr || !handle64) return ESCGNN_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(d_ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaMemset(*d_ptr, 0, (size_t)bytes)) != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, *d_ptr)) != cudaSuccess) return (int)e;
    memcpy(handle64, &h, 64);
    return (int)cudaDeviceSynchronize();
}

int escgnn_p2p_open(const unsigned char* handle64, void** d_ptr) {
    if (!d_ptr || !handle64) return ESCGNN_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int escgnn_p2p_close(void* d_ptr) { return d_ptr ? (int)cudaIpcCloseMemHandle(d_ptr) : 0; }
int escgnn_p2p_free(void* d_ptr) { return d_ptr ? (int)cudaFree(d_ptr) : 0; }

int64_t escgnn_p2p_flag_words(int world) { return 4 * (2 * (int64_t)world + 8); }      // four independent buckets

int escgnn_allreduce_adam(const float* const* h_peer_grads, float* const* h_peer_params, unsigned long long* const* h_peer_flags, int rank,
                          int world, int64_t n, float* d_exp_avg, float* d_exp_avg_sq, float* d_hyper, long long* d_state, void* stream) {
    return escgnn_allreduce_adam_range(h_peer_grads, h_peer_params, h_peer_flags, rank, world, 0, n, 0, 1, d_exp_avg, d_exp_avg_sq, d_hyper,
                                       d_state, stream);
}

int escgnn_allreduce_adam_range(const float* const* h_peer_grads, float* const* h_peer_params, unsigned long long* const* h_peer_flags,
                                int rank, int world, int64_t begin, int64_t end, int bucket, int tick, float* d_exp_avg,
                                float* d_exp_avg_sq, float* d_hyper, long long* d_state, void* stream) {
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || begin < 0 || end <= begin || ((begin | end) & 3) || bucket < 0 ||
        bucket > 3)
        return ESCGNN_ERR_BAD_ARG;
    Peers p;
    for (int i = 0; i < kMaxWorld; ++i) { p.grad[i] = nullptr; p.param[i] = nullptr; p.flags[i] = nullptr; }
    for (int i = 0; i < world; ++i) { p.grad[i] = h_peer_grads[i]; p.param[i] = h_peer_params[i]; p.flags[i] = h_peer_flags[i]; }
    const int64_t lo4 = begin >> 2, hi4 = end >> 2, slice4 = (hi4 - lo4 + world - 1) / world;
    int64_t blocks = (slice4 + 255) / 256;
    if (blocks > 148) blocks = 148;                      // every CTA spins on the ready flags: all must be resident at once
    if (blocks < 1) blocks = 1;
    escgnn::launch_pdl(allreduce_adam_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, p, rank, world, lo4, hi4,
                       (int64_t)bucket * (2 * world + 8), tick, d_exp_avg, d_exp_avg_sq, d_hyper, d_state);
    return (int)cudaGetLastError();
}

}  // extern "C"
