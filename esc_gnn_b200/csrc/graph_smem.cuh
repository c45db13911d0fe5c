// Per-graph on-chip working set shared by the encoder kernels (encode.cu, rd.cu).
//
// One CTA owns one graph at a time: both CSRs (by target for the BFS, by source for the induced edge walk)
// and the N x N hop-distance matrix, 4 bits per entry (15 = farther than h), live in shared memory, or in a
// per-CTA global slab when the graph is too large.  Reference semantics: utils_edge_efficient.py:201-294
// (k_hop_subgraph: BFS walks target->source, `col, row = edge_index`, :207-210).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/escgnn_b200.h"

namespace escgnn {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kFar = 15u;   // nibble value of "not within h hops"

__host__ __device__ inline int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }
__host__ __device__ inline int row_words(int64_t n) { return (int)((n + 7) >> 3); }

// Byte layout of one graph's working set (same formula on host and device).
struct GraphLayout {
    int64_t off_in_ptr, off_out_ptr, off_in_adj, off_out_adj, off_dist, total;
    __host__ __device__ GraphLayout(int64_t n, int64_t e) {
        int64_t o = 0;
        off_in_ptr = o;  o += align16((n + 1) * 4);
        off_out_ptr = o; o += align16((n + 1) * 4);
        off_in_adj = o;  o += align16(e * 2);
        off_out_adj = o; o += align16(e * 2);
        off_dist = o;
        int64_t dist = n * (int64_t)row_words(n) * 4;
        int64_t cur = 2 * n * 4;                       // fill cursors alias the distance matrix
        o += align16(dist > cur ? dist : cur);
        total = o;
    }
};

struct GraphView {
    int n, e, rw;                 // nodes, directed edges (after E1), words per distance row
    uint32_t* in_ptr;             // [n+1] CSR by target
    uint32_t* out_ptr;            // [n+1] CSR by source
    uint16_t* in_adj;             // [e] sources of edges into t
    uint16_t* out_adj;            // [e] targets of edges out of s
    uint32_t* dist;               // [n][rw] nibbles
    int max_out_deg;
};

__device__ __forceinline__ uint32_t nib(const uint32_t* row, int w) {
    return (row[w >> 3] >> ((w & 7) << 2)) & 15u;
}

// SWAR over the 8 nibbles of a word: bit 4k set where nibble k == v
__device__ __forceinline__ uint32_t nibbles_equal(uint32_t word, uint32_t v) {
    const uint32_t y = word ^ (v * 0x11111111u);            // matching nibbles become 0
    return ~(y | (y >> 1) | (y >> 2) | (y >> 3)) & 0x11111111u;
}
// bit 4k set where nibble k != 15 (node within h hops)
__device__ __forceinline__ uint32_t nibbles_near(uint32_t word) { return ~nibbles_equal(word, kFar) & 0x11111111u; }

// upper bound on the persistent encoder grids (0 = fill the machine): a pipelined training step confines the encoder of
// the NEXT batch to a few SMs so that it does not evict the latency-critical kernels of the current one
inline int& encoder_grid_cap() {
    static int cap = 0;
    return cap;
}

constexpr int kWideNodes = 64;      // graphs above this size scan 8 nodes per lane (one distance word) instead of 1

// Build both CSRs and run the N bounded BFS.  All threads of the CTA call this; `base` points at the graph's
// working set (shared or global), `s_misc` at >= 2 ints of shared scratch.  Returns false (uniformly) when the
// edge list holds a node id outside [0, n) (error bit already raised).
template <int H>
__device__ bool load_graph(GraphView& g, unsigned char* base, const int64_t* __restrict__ src,
                           const int64_t* __restrict__ dst, int n, int e, int* s_misc,
                           unsigned long long* counters) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    GraphLayout L(n, e);
    g.n = n; g.e = e; g.rw = row_words(n);
    g.in_ptr = reinterpret_cast<uint32_t*>(base + L.off_in_ptr);
    g.out_ptr = reinterpret_cast<uint32_t*>(base + L.off_out_ptr);
    g.in_adj = reinterpret_cast<uint16_t*>(base + L.off_in_adj);
    g.out_adj = reinterpret_cast<uint16_t*>(base + L.off_out_adj);
    g.dist = reinterpret_cast<uint32_t*>(base + L.off_dist);
    uint32_t* cur_in = g.dist;            // alias: cursors are dead before the matrix is initialised
    uint32_t* cur_out = g.dist + n;

    for (int i = tid; i <= n; i += nt) { g.in_ptr[i] = 0; g.out_ptr[i] = 0; }
    for (int i = tid; i < 2 * n; i += nt) cur_in[i] = 0;
    if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0; }
    __syncthreads();
    // degree counts (shifted by one so an inclusive scan yields the CSR offsets)
    bool bad = false;
    for (int i = tid; i < e; i += nt) {
        long long s = src[i], t = dst[i];
        if (s < 0 || s >= n || t < 0 || t >= n) { bad = true; continue; }
        atomicAdd(&g.in_ptr[t + 1], 1u);
        atomicAdd(&g.out_ptr[s + 1], 1u);
    }
    if (bad) s_misc[0] = 1;
    __syncthreads();
    if (s_misc[0]) {
        if (tid == 0) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)ESCGNN_DATA_NODE);
        return false;
    }
    // inclusive scans: warp 0 -> in_ptr, warp 1 (or warp 0 again) -> out_ptr
    for (int which = warp; which < 2; which += nw) {
        uint32_t* p = which == 0 ? g.in_ptr : g.out_ptr;
        uint32_t carry = 0;
        int mx = 0;
        for (int i0 = 0; i0 <= n; i0 += 32) {
            int i = i0 + lane;
            uint32_t v = i <= n ? p[i] : 0u;
            mx = max(mx, (int)v);
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(kFull, v, d);
                if (lane >= d) v += t;
            }
            if (i <= n) p[i] = v + carry;
            carry += __shfl_sync(kFull, v, 31);
        }
        if (which == 1) {
            #pragma unroll
            for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, d));
            if (lane == 0) s_misc[1] = mx;
        }
    }
    __syncthreads();
    g.max_out_deg = s_misc[1];
    // fill adjacency (order inside a list is irrelevant: every consumer is order-invariant)
    for (int i = tid; i < e; i += nt) {
        int s = (int)src[i], t = (int)dst[i];
        g.in_adj[g.in_ptr[t] + atomicAdd(&cur_in[t], 1u)] = (uint16_t)s;
        g.out_adj[g.out_ptr[s] + atomicAdd(&cur_out[s], 1u)] = (uint16_t)t;
    }
    __syncthreads();
    const int rw = g.rw;
    for (int i = tid; i < n * rw; i += nt) g.dist[i] = 0xffffffffu;
    __syncthreads();
    // E2: one warp per root; frontier = nodes whose nibble equals the current level
    for (int r = warp; r < n; r += nw) {
        volatile uint32_t* row = g.dist + (size_t)r * rw;
        if (lane == 0) row[r >> 3] = row[r >> 3] & ~(15u << ((r & 7) << 2));
        __syncwarp();
        for (int level = 0; level < H; ++level) {
            bool grew = false;
            auto expand = [&](int w) {
                const uint32_t a = g.in_ptr[w], b = g.in_ptr[w + 1];
                for (uint32_t k = a; k < b; ++k) {
                    const int s = g.in_adj[k];
                    const int sh = (s & 7) << 2;
                    if (((row[s >> 3] >> sh) & 15u) == kFar) {
                        // 15 -> level+1: clearing the zero bits of (level+1) is idempotent among same-level racers
                        atomicAnd(const_cast<uint32_t*>(&row[s >> 3]), ~((15u ^ (uint32_t)(level + 1)) << sh));
                        grew = true;
                    }
                }
            };
            if (n <= kWideNodes) {
                for (int w = lane; w < n; w += 32)
                    if (((row[w >> 3] >> ((w & 7) << 2)) & 15u) == (uint32_t)level) expand(w);
            } else {                                   // frontier = nibbles equal to `level`, 8 nodes per lane per step
                for (int wi = lane; wi < rw; wi += 32) {
                    uint32_t m = nibbles_equal(row[wi], (uint32_t)level);
                    while (m) {
                        const int w = wi * 8 + ((__ffs(m) - 1) >> 2);
                        m &= m - 1;
                        if (w < n) expand(w);
                    }
                }
            }
            __syncwarp();
            if (!__any_sync(kFull, grew)) break;
        }
    }
    __syncthreads();
    return true;
}

}  // namespace escgnn
