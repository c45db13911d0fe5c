// K1 `ego_encode`: per-edge h-hop ego-network structural encoding, integer blocks (E2-E4 of SURVEY.md section 8a).
//
// Replaces /root/reference/utils_edge_efficient.py:41-90 (two k_hop_subgraph calls per edge, union, relabel, degree)
// and :122-144 (one-hot sums -> nonzero).  B200-first shape: N bounded BFS per graph instead of 2E, distance
// matrix resident in shared memory (4 bit / entry), one warp per directed edge histogramming into a private
// shared-memory table with first-touch counting, one global atomic per edge to reserve its records, ascending
// emission by scanning the table.  Also here: E1 rewrite, exclusive scan, E6 expansion to the int64 triple.
#include <cuda_runtime.h>
#include <stdint.h>

#include "graph_smem.cuh"

namespace escgnn {

// ---- per-warp histogram table -------------------------------------------------------------------------------
// bins: [0,200) degree | [200,206) d0 | [208,214) d1 | [216, 216+(H+2)^4) compact distance-pair codes.
// Two 16-bit counters per 32-bit word (counts < 65536 is enforced by the host: E_graph, N <= 65535).
constexpr int kBinZ0 = 200, kBinZ1 = 208, kBinCode = 216;
template <int H> struct Bins {
    static constexpr int B = H + 2;
    static constexpr int kCodes = B * B * B * B;
    static constexpr int kBins = kBinCode + kCodes;
    static constexpr int kWords = (kBins + 1) / 2;
};

// add 1 to bin b; returns 1 when the bin was empty before (first touch -> one more output record)
__device__ __forceinline__ int hist_add(uint32_t* hist, int b) {
    const uint32_t old = atomicAdd(&hist[b >> 1], 1u << ((b & 1) << 4));
    return ((old >> ((b & 1) << 4)) & 0xffffu) == 0u;
}

template <int H, int THREADS>
__global__ void __launch_bounds__(THREADS)
ego_encode_kernel(const int64_t* __restrict__ eo_src, const int64_t* __restrict__ eo_dst,
                  const int64_t* __restrict__ eo_ptr, const int64_t* __restrict__ node_ptr, int n_graphs,
                  const uint16_t* __restrict__ rdh, uint32_t* __restrict__ rec, long long rec_cap,
                  int64_t* __restrict__ rec_off, int32_t* __restrict__ rec_nnz, int32_t* __restrict__ edge_graph,
                  unsigned long long* counters, long long graph_smem_bytes, unsigned char* scratch,
                  long long slab_bytes, const int32_t* __restrict__ graph_ids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket;
    __shared__ int s_misc[2];
    constexpr int B = Bins<H>::B;
    constexpr int kWords = Bins<H>::kWords;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem) + warp * kWords;
    unsigned char* graph_smem = smem + align16((int64_t)nw * kWords * 4);
    const int use_rd = rdh != nullptr;
    const int code_off = use_rd ? 500 : 400;

    for (int i = lane; i < kWords; i += 32) hist[i] = 0;   // table is kept all-zero between edges

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ticket = (int)atomicAdd(&counters[ESCGNN_CTR_TICKET], 1ull);
        __syncthreads();
        if (s_ticket >= n_graphs) break;
        const int gi = graph_ids ? graph_ids[s_ticket] : s_ticket;      // size-class launches walk an id list
        const long long e0 = eo_ptr[gi];
        const int e = (int)(eo_ptr[gi + 1] - e0);
        const int n = (int)(node_ptr[gi + 1] - node_ptr[gi]);
        for (int i = tid; i < e; i += blockDim.x) edge_graph[e0 + i] = gi;
        if (e == 0) continue;
        GraphLayout L(n, e);
        unsigned char* base = (L.total <= graph_smem_bytes) ? graph_smem : scratch + (size_t)blockIdx.x * slab_bytes;
        GraphView g;
        if (!load_graph<H>(g, base, eo_src + e0, eo_dst + e0, n, e, s_misc, counters)) {
            for (int i = tid; i < e; i += blockDim.x) { rec_off[e0 + i] = 0; rec_nnz[e0 + i] = 0; }
            continue;
        }
        const int rw = g.rw;
        const int deg_words = min(100, (g.max_out_deg + 2) >> 1);       // words holding degree bins 0..max_out_deg

        for (int ed = warp; ed < e; ed += nw) {
            const int u = (int)eo_src[e0 + ed], v = (int)eo_dst[e0 + ed];
            const uint32_t* rowU = g.dist + (size_t)u * rw;
            const uint32_t* rowV = g.dist + (size_t)v * rw;
            int fresh = 0;
            bool deg_err = false;
            if (u == v && lane == 0) {          // phantom duplicate root: degree 0, z = (0,0)
                fresh += hist_add(hist, 0) + hist_add(hist, kBinZ0) + hist_add(hist, kBinZ1);
            }
            auto visit = [&](int w, uint32_t du, uint32_t dv) {       // one member w of S = B_u U B_v
                const int z0 = min((int)du, H + 1), z1 = min((int)dv, H + 1);
                fresh += hist_add(hist, kBinZ0 + z0) + hist_add(hist, kBinZ1 + z1);
                const int cbase = kBinCode + (z0 * B + z1) * B * B;
                int deg = 0;
                const uint32_t ka = g.out_ptr[w], kb = g.out_ptr[w + 1];
                for (uint32_t k = ka; k < kb; ++k) {
                    const int b = g.out_adj[k];
                    const uint32_t bu = nib(rowU, b), bv = nib(rowV, b);
                    // F = induced(B_u) OR induced(B_v)   (utils_edge_efficient.py:55)
                    if (!((du != kFar && bu != kFar) || (dv != kFar && bv != kFar))) continue;
                    ++deg;                                  // loops count once in the degree (:86) ...
                    if (b == w) continue;                   // ... but are removed before the pair code (:138)
                    fresh += hist_add(hist, cbase + min((int)bu, H + 1) * B + min((int)bv, H + 1));
                }
                if (deg >= 200) deg_err = true; else fresh += hist_add(hist, deg);
            };
            if (n <= kWideNodes) {                       // small graphs: one node per lane
                for (int w = lane; w < n; w += 32) {
                    const uint32_t du = nib(rowU, w), dv = nib(rowV, w);
                    if (du == kFar && dv == kFar) continue;
                    visit(w, du, dv);
                }
            } else {                                     // large graphs: one distance word (8 nodes) per lane, members only
                for (int wi = lane; wi < rw; wi += 32) {
                    const uint32_t xu = rowU[wi], xv = rowV[wi];
                    uint32_t m = nibbles_near(xu) | nibbles_near(xv);
                    while (m) {
                        const int sh = __ffs(m) - 1;
                        m &= m - 1;
                        const int w = wi * 8 + (sh >> 2);
                        if (w < n) visit(w, (xu >> sh) & 15u, (xv >> sh) & 15u);
                    }
                }
            }
            // rd block comes pre-binned from K1b
            uint32_t rdc = 0;
            if (use_rd && lane < ESCGNN_RD_SLOTS) {
                rdc = rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS + lane];
                fresh += rdc != 0;
            }
            #pragma unroll
            for (int d = 16; d; d >>= 1) fresh += __shfl_xor_sync(kFull, fresh, d);
            if (__any_sync(kFull, deg_err) && lane == 0)
                atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)ESCGNN_DATA_DEG);
            long long off = 0;
            if (lane == 0) off = (long long)atomicAdd(&counters[ESCGNN_CTR_NNZ], (unsigned long long)fresh);
            off = __shfl_sync(kFull, off, 0);
            const bool room = off + fresh <= rec_cap;
            // an edge without room owns NO records (offset 0, count 0): consumers of the compact form never index past the
            // capacity; the overflow itself is visible in counters[NNZ] > rec_cap (and in the sticky slots, escgnn_make_dims)
            if (lane == 0) { rec_off[e0 + ed] = room ? off : 0; rec_nnz[e0 + ed] = room ? fresh : 0; }
            __syncwarp();
            // ---- ascending emission; each lane owns one word = two consecutive bins; table is zeroed on the way
            int pos = 0;
            const unsigned lt = (1u << lane) - 1u;
            auto emit_words = [&](int w_lo, int w_hi, auto bin_to_index) {
                for (int w0 = w_lo; w0 < w_hi; w0 += 32) {
                    const int wi = w0 + lane;
                    uint32_t word = 0;
                    if (wi < w_hi) word = hist[wi];
                    const unsigned any = __ballot_sync(kFull, word != 0u);
                    if (any == 0u) continue;                       // most 32-word spans of the code block are empty
                    if (word) hist[wi] = 0;
                    const uint32_t c0 = word & 0xffffu, c1 = word >> 16;
                    const unsigned m0 = __ballot_sync(kFull, c0 != 0u), m1 = __ballot_sync(kFull, c1 != 0u);
                    // records of lower lanes come first; inside a lane the even bin precedes the odd one
                    int p = pos + __popc(m0 & lt) + __popc(m1 & lt);
                    if (room) {
                        if (c0) rec[off + p++] = (uint32_t)bin_to_index(2 * wi) | (c0 << ESCGNN_REC_IDX_BITS);
                        if (c1) rec[off + p] = (uint32_t)bin_to_index(2 * wi + 1) | (c1 << ESCGNN_REC_IDX_BITS);
                    }
                    pos += __popc(m0) + __popc(m1);
                }
            };
            emit_words(0, deg_words, [](int b) { return b; });                           // degree   [0,200)
            emit_words(kBinZ0 / 2, kBinZ0 / 2 + 8, [](int b) {                            // d0 -> 200+, d1 -> 300+
                return b < kBinZ1 ? 200 + (b - kBinZ0) : 300 + (b - kBinZ1); });
            if (use_rd) {                                                                  // rd       [400,500)
                const unsigned m = __ballot_sync(kFull, rdc != 0);
                if (rdc && room) rec[off + pos + __popc(m & ((1u << lane) - 1))] =
                    (uint32_t)(400 + lane) | (rdc << ESCGNN_REC_IDX_BITS);
                pos += __popc(m);
            }
            emit_words(kBinCode / 2, kWords, [code_off](int b) {                           // pair codes
                const int c = b - kBinCode;
                const int b1 = c % B, b0 = (c / B) % B, a1 = (c / (B * B)) % B, a0 = c / (B * B * B);
                return code_off + 216 * a0 + 36 * a1 + 6 * b0 + b1; });
            __syncwarp();
        }
    }
}

// ---- E1 ------------------------------------------------------------------------------------------------------
__global__ void count_rewritten_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                       const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ node_ptr,
                                       int64_t n_graphs, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t gi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gi >= n_graphs) return;
    const int64_t a = edge_ptr[gi], b = edge_ptr[gi + 1];
    int c = 0;
    for (int64_t i = a + lane; i < b; i += 32) c += src[i] != dst[i];
    #pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
    if (lane == 0) counts[gi] = c + (int)(node_ptr[gi + 1] - node_ptr[gi]);
}

__global__ void rewrite_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                               const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ node_ptr,
                               int64_t n_graphs, const int64_t* __restrict__ eo_ptr, int64_t* __restrict__ eo_src,
                               int64_t* __restrict__ eo_dst) {
    const int lane = threadIdx.x & 31;
    const int64_t gi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gi >= n_graphs) return;
    const int64_t a = edge_ptr[gi], b = edge_ptr[gi + 1];
    int64_t o = eo_ptr[gi];
    for (int64_t i0 = a; i0 < b; i0 += 32) {         // order-preserving compaction of the non-loop edges
        const int64_t i = i0 + lane;
        int64_t s = 0, t = 0;
        bool keep = false;
        if (i < b) { s = src[i]; t = dst[i]; keep = s != t; }
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) { const int p = __popc(m & ((1u << lane) - 1)); eo_src[o + p] = s; eo_dst[o + p] = t; }
        o += __popc(m);
    }
    const int64_t n = node_ptr[gi + 1] - node_ptr[gi];
    for (int64_t i = lane; i < n; i += 32) { eo_src[o + i] = i; eo_dst[o + i] = i; }   // appended (i,i), node order
}

// ---- exclusive scan int32 -> int64, three small kernels -------------------------------------------------------
constexpr int kScanBlock = 256, kScanItems = 4, kScanTile = kScanBlock * kScanItems;   // 1024 per block

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* s_warp, long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long t = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0, wi = w;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long t = __shfl_up_sync(kFull, wi, d); if (lane >= d) wi += t; }
        if (lane < (blockDim.x >> 5)) s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    const long long r = s_warp[warp] + incl - v;
    __syncthreads();
    return r;
}

__global__ void scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
    __shared__ long long s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    long long v = 0;
    #pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) v += in[base + k];
    long long total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_tile_offsets_kernel(int64_t* tile_sums, int64_t n_tiles) {   // single block, in place
    __shared__ long long s_warp[33];
    long long carry = 0;
    for (int64_t i0 = 0; i0 < n_tiles; i0 += blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const long long v = i < n_tiles ? tile_sums[i] : 0;
        long long total;
        const long long ex = block_exclusive_scan(v, s_warp, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_off,
                                  int64_t n_tiles, int64_t* __restrict__ out) {
    __shared__ long long s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    long long x[kScanItems], v = 0;
    #pragma unroll
    for (int k = 0; k < kScanItems; ++k) { x[k] = base + k < n ? in[base + k] : 0; v += x[k]; }
    long long total;
    long long ex = block_exclusive_scan(v, s_warp, total) + tile_off[blockIdx.x];
    #pragma unroll
    for (int k = 0; k < kScanItems; ++k) { if (base + k < n) out[base + k] = ex; ex += x[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_off[n_tiles];
}

// ---- E6: records -> (pos_enc, pos_index, pos_batch) int64 ------------------------------------------------------
__global__ void expand_records_kernel(const uint32_t* __restrict__ rec, const int64_t* __restrict__ rec_off,
                                      const int32_t* __restrict__ rec_nnz, const int32_t* __restrict__ edge_graph,
                                      const int64_t* __restrict__ eo_ptr, const int64_t* __restrict__ out_off,
                                      int64_t n_edges, int local_ordinals, int64_t* __restrict__ pos_enc,
                                      int64_t* __restrict__ pos_index, int64_t* __restrict__ pos_batch) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t ed = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ed < n_edges; ed += wpg) {
        const int k = rec_nnz[ed];
        const int64_t ri = rec_off[ed], oi = out_off[ed];
        const int64_t ordinal = local_ordinals ? ed - eo_ptr[edge_graph[ed]] : ed;
        for (int i = lane; i < k; i += 32) {
            const uint32_t r = rec[ri + i];
            pos_index[oi + i] = r & ((1u << ESCGNN_REC_IDX_BITS) - 1);
            pos_enc[oi + i] = r >> ESCGNN_REC_IDX_BITS;
            pos_batch[oi + i] = ordinal;
        }
    }
}

template <int H, int kThreads>
static int launch_encode_t(const int64_t* eo_src, const int64_t* eo_dst, const int64_t* eo_ptr, const int64_t* node_ptr,
                         int64_t n_graphs, const uint16_t* rdh, uint32_t* rec, int64_t rec_cap, int64_t* rec_off,
                         int32_t* rec_nnz, int32_t* edge_graph, unsigned long long* counters, int64_t max_nodes,
                         int64_t max_edges, void* scratch, int64_t scratch_bytes, const int32_t* graph_ids, cudaStream_t st) {
    constexpr int kNw = kThreads / 32;
    const int64_t hist_bytes = align16((int64_t)kNw * Bins<H>::kWords * 4);
    const int64_t need = GraphLayout(max_nodes, max_edges).total;
    int dev = 0, sms = 148, smem_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int64_t cap = (int64_t)smem_optin - hist_bytes - 1024;
    int64_t graph_bytes = need <= cap ? need : 0;          // too large: every such graph works from the global slab
    if (need > cap) {
        // keep a modest on-chip area so the small graphs of a mixed batch still run from shared memory
        graph_bytes = cap < 64 * 1024 ? cap : 64 * 1024;
    }
    const int64_t smem = hist_bytes + graph_bytes;
    auto kern = ego_encode_kernel<H, kThreads>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, (size_t)smem);
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > n_graphs) grid = n_graphs;
    if (encoder_grid_cap() > 0 && grid > encoder_grid_cap()) grid = encoder_grid_cap();
    if (grid < 1) grid = 1;
    int64_t slab = 0;
    if (need > graph_bytes) {
        slab = need;
        if (scratch == nullptr || scratch_bytes < slab) return ESCGNN_ERR_BAD_ARG;
        const int64_t fit = scratch_bytes / slab;
        if (grid > fit) grid = fit;
    }
    cudaMemsetAsync(counters + ESCGNN_CTR_TICKET, 0, sizeof(unsigned long long), st);     // every launch starts its own queue
    kern<<<(unsigned)grid, kThreads, (size_t)smem, st>>>(eo_src, eo_dst, eo_ptr, node_ptr, (int)n_graphs, rdh, rec,
                                                          (long long)rec_cap, rec_off, rec_nnz, edge_graph, counters,
                                                          (long long)graph_bytes, (unsigned char*)scratch,
                                                          (long long)slab, graph_ids);
    return (int)cudaGetLastError();
}

// large graphs get 16 warps per CTA: their distance matrix allows one CTA per SM, so parallelism has to come from warps
template <int H>
static int launch_encode(const int64_t* eo_src, const int64_t* eo_dst, const int64_t* eo_ptr, const int64_t* node_ptr,
                         int64_t n_graphs, const uint16_t* rdh, uint32_t* rec, int64_t rec_cap, int64_t* rec_off,
                         int32_t* rec_nnz, int32_t* edge_graph, unsigned long long* counters, int64_t max_nodes,
                         int64_t max_edges, void* scratch, int64_t scratch_bytes, const int32_t* graph_ids, cudaStream_t st) {
    if (max_nodes > 160)
        return launch_encode_t<H, 512>(eo_src, eo_dst, eo_ptr, node_ptr, n_graphs, rdh, rec, rec_cap, rec_off, rec_nnz, edge_graph,
                                        counters, max_nodes, max_edges, scratch, scratch_bytes, graph_ids, st);
    return launch_encode_t<H, 256>(eo_src, eo_dst, eo_ptr, node_ptr, n_graphs, rdh, rec, rec_cap, rec_off, rec_nnz, edge_graph,
                                   counters, max_nodes, max_edges, scratch, scratch_bytes, graph_ids, st);
}

}  // namespace escgnn

using namespace escgnn;

extern "C" {

int escgnn_encode_subset(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                         const int64_t* d_node_ptr, int64_t n_graphs, const int32_t* d_graph_ids, int h, const uint16_t* d_rdh,
                         uint32_t* d_rec, int64_t rec_cap, int64_t* d_rec_off, int32_t* d_rec_nnz, int32_t* d_edge_graph,
                         unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                         int64_t scratch_bytes, void* stream);

int escgnn_version(void) { return 100; }

int escgnn_set_encoder_grid_cap(int ctas) {
    const int was = encoder_grid_cap();
    encoder_grid_cap() = ctas > 0 ? ctas : 0;
    return was;
}

int64_t escgnn_encode_scratch_bytes(int64_t max_nodes, int64_t max_edges, int h) {
    (void)h;
    const int64_t need = GraphLayout(max_nodes, max_edges).total;
    if (need <= 150 * 1024) return 0;                       // always fits beside the histogram tables
    return need * 148 * 2;                                   // one slab per resident CTA (upper bound)
}

int escgnn_rewrite_self_loops(const int64_t* d_src, const int64_t* d_dst, const int64_t* d_edge_ptr,
                              const int64_t* d_node_ptr, int64_t n_graphs, int64_t* d_eo_ptr, int64_t* d_eo_src,
                              int64_t* d_eo_dst, void* d_tmp, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_graphs <= 0) return (int)cudaMemsetAsync(d_eo_ptr, 0, sizeof(int64_t), st);
    int64_t* tile = reinterpret_cast<int64_t*>(d_tmp);        // scan scratch: n_graphs/1024 + 2 entries
    int32_t* counts = reinterpret_cast<int32_t*>(tile + n_graphs / 1024 + 2);
    const int wpb = 8;
    const unsigned blocks = (unsigned)((n_graphs + wpb - 1) / wpb);
    count_rewritten_kernel<<<blocks, wpb * 32, 0, st>>>(d_src, d_dst, d_edge_ptr, d_node_ptr, n_graphs, counts);
    int rc = escgnn_exclusive_scan_i32(counts, n_graphs, d_eo_ptr, tile, stream);
    if (rc) return rc;
    rewrite_kernel<<<blocks, wpb * 32, 0, st>>>(d_src, d_dst, d_edge_ptr, d_node_ptr, n_graphs, d_eo_ptr, d_eo_src,
                                                 d_eo_dst);
    return (int)cudaGetLastError();
}

int escgnn_exclusive_scan_i32(const int32_t* d_in, int64_t n, int64_t* d_out, int64_t* d_tmp, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0) return ESCGNN_ERR_BAD_ARG;
    if (n == 0) { return (int)cudaMemsetAsync(d_out, 0, sizeof(int64_t), st); }
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(d_in, n, d_tmp);
    scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(d_tmp, tiles);
    scan_apply_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(d_in, n, d_tmp, tiles, d_out);
    return (int)cudaGetLastError();
}

int escgnn_encode(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                  const int64_t* d_node_ptr, int64_t n_graphs, int h, const uint16_t* d_rdh, uint32_t* d_rec,
                  int64_t rec_cap, int64_t* d_rec_off, int32_t* d_rec_nnz, int32_t* d_edge_graph,
                  unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                  int64_t scratch_bytes, void* stream) {
    return escgnn_encode_subset(d_eo_src, d_eo_dst, d_eo_ptr, d_node_ptr, n_graphs, nullptr, h, d_rdh, d_rec, rec_cap, d_rec_off,
                                d_rec_nnz, d_edge_graph, d_counters, max_nodes, max_edges, d_scratch, scratch_bytes, stream);
}

int escgnn_encode_subset(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                         const int64_t* d_node_ptr, int64_t n_graphs, const int32_t* d_graph_ids, int h, const uint16_t* d_rdh,
                         uint32_t* d_rec, int64_t rec_cap, int64_t* d_rec_off, int32_t* d_rec_nnz, int32_t* d_edge_graph,
                         unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                         int64_t scratch_bytes, void* stream) {
    if (n_graphs <= 0) return 0;
    if (h < 1 || h > 4) return ESCGNN_ERR_BAD_ARG;            // reference: one_hot(code, 1300) raises for h >= 5
    if (max_nodes > 65535 || max_edges > 65535) return ESCGNN_ERR_TOO_LARGE;
    cudaStream_t st = (cudaStream_t)stream;
#define ESC_GO(H) launch_encode<H>(d_eo_src, d_eo_dst, d_eo_ptr, d_node_ptr, n_graphs, d_rdh, d_rec, rec_cap, \
                                   d_rec_off, d_rec_nnz, d_edge_graph, d_counters, max_nodes, max_edges, d_scratch, \
                                   scratch_bytes, d_graph_ids, st)
    switch (h) {
        case 1: return ESC_GO(1);
        case 2: return ESC_GO(2);
        case 3: return ESC_GO(3);
        default: return ESC_GO(4);
    }
#undef ESC_GO
}

int escgnn_expand_records(const uint32_t* d_rec, const int64_t* d_rec_off, const int32_t* d_rec_nnz,
                          const int32_t* d_edge_graph, const int64_t* d_eo_ptr, const int64_t* d_out_off,
                          int64_t n_edges, int use_rd, int local_ordinals, int64_t* d_pos_enc, int64_t* d_pos_index,
                          int64_t* d_pos_batch, void* stream) {
    (void)use_rd;
    if (n_edges <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (n_edges + 7) / 8;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    expand_records_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_rec, d_rec_off, d_rec_nnz, d_edge_graph, d_eo_ptr, d_out_off, n_edges, local_ordinals, d_pos_enc,
        d_pos_index, d_pos_batch);
    return (int)cudaGetLastError();
}

}  // extern "C"
