// All-pairs shortest-path lengths of every graph of a batch (the `attn_bias` field of the GraphGPS twin of the
// transform, /root/reference/GraphGPS/graphgps/loader/utils_escgnn.py:29-38: networkx
// all_pairs_shortest_path_length on the UNDIRECTED graph, unreachable pairs = 100, flattened [n*n] int64).
//
// One CTA per graph (grid-stride over graphs): undirected CSR in shared memory, then one warp per root runs a
// level-synchronous BFS over its own 16-bit distance row (also in shared memory) and streams the finished row to
// global memory as int64 (coalesced: lane = column).  Integer work, bound by the n*n*8 output bytes per graph.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/escgnn_b200.h"

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kUnseen = 0xffffu;

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
__host__ __device__ inline size_t spd_smem_bytes(int64_t max_nodes, int64_t max_edges) {
    return align16((size_t)(max_nodes + 1) * 4) * 2 + align16((size_t)max_edges * 2 * 2) + (size_t)kWarps * align16((size_t)max_nodes * 2);
}

__global__ void __launch_bounds__(kThreads)
all_pairs_spd_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, const int64_t* __restrict__ edge_ptr,
                     const int64_t* __restrict__ node_ptr, int64_t n_graphs, const int64_t* __restrict__ out_ptr,
                     int64_t* __restrict__ out, int max_nodes, int max_edges, long long unreachable,
                     unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* ptr = reinterpret_cast<uint32_t*>(smem);
    uint32_t* cur = reinterpret_cast<uint32_t*>(smem + align16((size_t)(max_nodes + 1) * 4));
    uint16_t* adj = reinterpret_cast<uint16_t*>(smem + 2 * align16((size_t)(max_nodes + 1) * 4));
    uint16_t* rows = reinterpret_cast<uint16_t*>(smem + 2 * align16((size_t)(max_nodes + 1) * 4) + align16((size_t)max_edges * 4));
    __shared__ int s_bad;
    for (int64_t g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const int n = (int)(node_ptr[g + 1] - node_ptr[g]);
        const int64_t e0 = edge_ptr[g];
        const int e = (int)(edge_ptr[g + 1] - e0);
        if (n > max_nodes || e > max_edges) {
            if (tid == 0) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)ESCGNN_DATA_NODE);
            continue;
        }
        for (int i = tid; i <= n; i += kThreads) { ptr[i] = 0; cur[i] = 0; }
        if (tid == 0) s_bad = 0;
        __syncthreads();
        for (int i = tid; i < e; i += kThreads) {
            const long long s = src[e0 + i], t = dst[e0 + i];
            if (s < 0 || s >= n || t < 0 || t >= n) { s_bad = 1; continue; }
            if (s == t) continue;                                   // loops do not change any path length
            atomicAdd(&ptr[s + 1], 1u);
            atomicAdd(&ptr[t + 1], 1u);
        }
        __syncthreads();
        if (s_bad) {
            if (tid == 0) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)ESCGNN_DATA_NODE);
            __syncthreads();
            continue;
        }
        if (warp == 0) {                                            // inclusive scan of the degree counts
            uint32_t carry = 0;
            for (int i0 = 0; i0 <= n; i0 += 32) {
                const int i = i0 + lane;
                uint32_t v = i <= n ? ptr[i] : 0u;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, v, d);
                    if (lane >= d) v += t;
                }
                if (i <= n) ptr[i] = v + carry;
                carry += __shfl_sync(kFull, v, 31);
            }
        }
        __syncthreads();
        for (int i = tid; i < e; i += kThreads) {
            const int s = (int)src[e0 + i], t = (int)dst[e0 + i];
            if (s == t) continue;
            adj[ptr[s] + atomicAdd(&cur[s], 1u)] = (uint16_t)t;
            adj[ptr[t] + atomicAdd(&cur[t], 1u)] = (uint16_t)s;
        }
        __syncthreads();
        volatile uint16_t* row = rows + (size_t)warp * (align16((size_t)max_nodes * 2) / 2);
        int64_t* o = out + out_ptr[g];
        for (int r = warp; r < n; r += kWarps) {
            for (int w = lane; w < n; w += 32) row[w] = kUnseen;
            __syncwarp();
            if (lane == 0) row[r] = 0;
            __syncwarp();
            for (int level = 0; level < n; ++level) {
                bool grew = false;
                for (int w = lane; w < n; w += 32) {
                    if (row[w] != (uint16_t)level) continue;
                    for (uint32_t k = ptr[w]; k < ptr[w + 1]; ++k) {
                        const int s = adj[k];
                        if (row[s] == kUnseen) { row[s] = (uint16_t)(level + 1); grew = true; }   // racers write the same value
                    }
                }
                __syncwarp();
                if (!__any_sync(kFull, grew)) break;
            }
            for (int w = lane; w < n; w += 32) {
                const uint16_t d = row[w];
                o[(size_t)r * n + w] = d == kUnseen ? unreachable : (long long)d;
            }
            __syncwarp();
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int64_t escgnn_all_pairs_spd_smem_bytes(int64_t max_nodes, int64_t max_edges) {
    return (int64_t)spd_smem_bytes(max_nodes, max_edges);
}

extern "C" int escgnn_all_pairs_spd(const int64_t* d_src, const int64_t* d_dst, const int64_t* d_edge_ptr,
                                    const int64_t* d_node_ptr, int64_t n_graphs, const int64_t* d_out_ptr, int64_t* d_out,
                                    int64_t max_nodes, int64_t max_edges, int64_t unreachable,
                                    unsigned long long* d_counters, void* stream) {
    if (n_graphs <= 0) return 0;
    if (max_nodes < 1) max_nodes = 1;
    if (max_nodes > 65534) return ESCGNN_ERR_TOO_LARGE;
    const size_t smem = spd_smem_bytes(max_nodes, max_edges);
    if (smem > 227 * 1024) return ESCGNN_ERR_TOO_LARGE;
    cudaError_t e = cudaFuncSetAttribute(all_pairs_spd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    unsigned blocks = (unsigned)(n_graphs < 148 * 8 ? n_graphs : 148 * 8);
    all_pairs_spd_kernel<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(d_src, d_dst, d_edge_ptr, d_node_ptr, n_graphs,
                                                                           d_out_ptr, d_out, (int)max_nodes, (int)max_edges,
                                                                           (long long)unreachable, d_counters);
    return (int)cudaGetLastError();
}
