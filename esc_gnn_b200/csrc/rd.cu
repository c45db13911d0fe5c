// K1b `ego_rd`: resistance-distance histogram of every directed edge's union subgraph (E5).
//
// Replaces /root/reference/utils_edge_efficient.py:92-107 (float32 scipy laplacian -> pinv per edge) and :130-131
// (one_hot(rd.long(), 100)) under parity policy E5: float64 arithmetic, bin = trunc((float)rd).
//
// For a symmetric, connected (S, F):  u != v: rd(w) = R(u,w) = [(L with row/col u removed)^-1]_ww, rd(u) = 0;
// u == v (phantom root): rd(w) = [pinv(L_ball)]_ww = [(L_ball + J/m)^-1]_ww - 1/m, rd(phantom) = 0.
// Both are "diagonal of the inverse of an SPD matrix": LDL^T in place on a packed lower triangle held in shared
// memory, then the Takahashi recurrence Z_ij = delta_ij/D_j - sum_{k>j} Z_ik L_kj run from the last column back.
// Each thread owns matrix rows (row i -> thread i mod group), so assembly and both sweeps need no atomics.
// Small graphs: one warp per edge; larger ones: the whole CTA works on one edge; too large for shared memory: the
// matrix lives in a per-CTA global slab.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "graph_smem.cuh"

namespace escgnn {

__host__ __device__ inline int64_t tri(int64_t m) { return m * (m + 1) / 2; }
__device__ __forceinline__ int tidx(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

template <bool kCta> __device__ __forceinline__ void group_sync() {
    if (kCta) __syncthreads(); else __syncwarp();
}

// sum over the group; s_red: one double per warp of shared scratch (CTA mode only)
template <bool kCta> __device__ __forceinline__ double group_sum(double v, double* s_red) {
    #pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    if (!kCta) return v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_red[i];
    return t;
}

// One edge (u, v): fills hist[ESCGNN_RD_SLOTS] (shared, ints, zeroed here). Returns error bits (uniform in group).
template <int H, bool kCta>
__device__ unsigned rd_edge(const GraphView& g, int u, int v, double* M, uint16_t* sub, int* hist, double* s_red,
                            int* s_cnt) {
    const int gt = kCta ? threadIdx.x : (threadIdx.x & 31);        // thread id inside the group
    const int gn = kCta ? blockDim.x : 32;                         // group size
    const int lane = threadIdx.x & 31;
    const int n = g.n, rw = g.rw;
    const uint32_t* rowU = g.dist + (size_t)u * rw;
    const uint32_t* rowV = g.dist + (size_t)v * rw;
    const bool phantom = u == v;
    // ---- matrix index of every member of S (node order); u is grounded (no row) unless phantom
    if (gt < ESCGNN_RD_SLOTS) hist[gt] = 0;
    int m = 0;
    if (!kCta || threadIdx.x < 32) {
        for (int w0 = 0; w0 < n; w0 += 32) {
            const int w = w0 + lane;
            bool in = false;
            if (w < n) in = (nib(rowU, w) != kFar || nib(rowV, w) != kFar) && (phantom || w != u);
            const unsigned b = __ballot_sync(kFull, in);
            if (w < n) sub[w] = in ? (uint16_t)(m + __popc(b & ((1u << lane) - 1))) : (uint16_t)0xffff;
            m += __popc(b);
        }
        if (kCta && lane == 0) *s_cnt = m;
    }
    group_sync<kCta>();
    if (kCta) m = *s_cnt;
    if (m == 0) {                                   // S = {u} only (cannot happen for an edge, kept for safety)
        if (gt == 0) hist[0] = 1;
        group_sync<kCta>();
        return 0u;
    }
    const double fill = phantom ? 1.0 / (double)m : 0.0;
    for (int t = gt; t < (int)tri(m); t += gn) M[t] = fill;
    group_sync<kCta>();
    // ---- assemble: thread owning node w writes row sub[w] (diagonal = degree in F without loops)
    for (int w = gt; w < n; w += gn) {
        const int i = sub[w];
        if (i == 0xffff) continue;
        const uint32_t du = nib(rowU, w), dv = nib(rowV, w);
        double deg = 0.0;
        const uint32_t ka = g.out_ptr[w], kb = g.out_ptr[w + 1];
        for (uint32_t k = ka; k < kb; ++k) {
            const int b = g.out_adj[k];
            if (b == w) continue;                                      // scipy laplacian ignores loops
            const uint32_t bu = nib(rowU, b), bv = nib(rowV, b);
            if (!((du != kFar && bu != kFar) || (dv != kFar && bv != kFar))) continue;
            deg += 1.0;
            const int j = sub[b];
            if (j != 0xffff && j < i) M[tidx(i, j)] -= 1.0;
        }
        M[tidx(i, i)] += deg;
    }
    group_sync<kCta>();
    // ---- LDL^T, column k keeps W_ik = L_ik * D_k (unscaled), diagonal keeps D_k
    bool bad = false;
    for (int k = 0; k < m; ++k) {
        const double d = M[tidx(k, k)];
        if (!(d > 1e-12)) { bad = true; break; }                      // uniform: every thread reads the same value
        const double invd = 1.0 / d;
        for (int i = k + 1 + gt; i < m; i += gn) {
            const double f = M[tidx(i, k)] * invd;
            if (f != 0.0) {
                double* row = M + tidx(i, 0);
                for (int j = k + 1; j <= i; ++j) row[j] -= f * M[tidx(j, k)];
            }
        }
        group_sync<kCta>();
    }
    if (bad) return ESCGNN_DATA_RD;
    // ---- Takahashi: columns from the last to the first; Z overwrites the factor column by column
    for (int j = m - 1; j >= 0; --j) {
        const double invd = 1.0 / M[tidx(j, j)];
        double part = 0.0;                      // sum_k W_kj * Z_kj over the rows this thread owns
        double zmine[4];                        // up to 4 rows per thread in flight (m <= 4 * group size)
        int cnt = 0;
        for (int i = j + 1 + gt; i < m; i += gn) {
            double acc = 0.0;
            for (int k = j + 1; k < m; ++k) {
                const double z = k <= i ? M[tidx(i, k)] : M[tidx(k, i)];
                acc += z * M[tidx(k, j)];
            }
            const double zij = -acc * invd;
            part += M[tidx(i, j)] * zij;
            if (cnt < 4) zmine[cnt] = zij;
            ++cnt;
        }
        const double s = group_sum<kCta>(part, s_red);     // (also orders the reads of column j before its overwrite)
        group_sync<kCta>();
        cnt = 0;
        for (int i = j + 1 + gt; i < m; i += gn) { M[tidx(i, j)] = zmine[cnt < 4 ? cnt : 3]; ++cnt; }
        if (gt == 0) M[tidx(j, j)] = invd - invd * s;
        group_sync<kCta>();
    }
    // ---- bin: trunc((float)rd)   (torch.FloatTensor(...) then .long(), utils_edge_efficient.py:105,131)
    unsigned err = 0;
    for (int w = gt; w < n; w += gn) {
        const int i = sub[w];
        if (i == 0xffff) continue;
        const double rd = M[tidx(i, i)] - fill;
        const float rf = (float)rd;
        const int b = (int)truncf(rf);
        if (b < 0 || b >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD; else atomicAdd(&hist[b], 1);
    }
    if (gt == 0) atomicAdd(&hist[0], 1);           // the grounded root u (rd = 0) or the phantom root (rd = 0)
    group_sync<kCta>();
    return err;
}

template <int H>
__global__ void __launch_bounds__(256)
ego_rd_kernel(const int64_t* __restrict__ eo_src, const int64_t* __restrict__ eo_dst,
              const int64_t* __restrict__ eo_ptr, const int64_t* __restrict__ node_ptr, int n_graphs,
              uint16_t* __restrict__ rdh, unsigned long long* counters, long long graph_smem_bytes,
              long long mat_region_doubles, int sub_stride, unsigned char* scratch, long long slab_bytes,
              long long slab_graph_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket;
    __shared__ int s_misc[2];
    __shared__ int s_hist[8][ESCGNN_RD_SLOTS];
    __shared__ double s_red[8];
    __shared__ int s_cnt;
    __shared__ unsigned s_err;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    double* mat_region = reinterpret_cast<double*>(smem);
    uint16_t* sub_region = reinterpret_cast<uint16_t*>(smem + mat_region_doubles * 8);
    unsigned char* graph_smem = smem + align16(mat_region_doubles * 8 + (long long)nw * sub_stride * 2);
    const long long warp_cap = mat_region_doubles / nw;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_ticket = (int)atomicAdd(&counters[ESCGNN_CTR_TICKET_RD], 1ull); s_err = 0; }
        __syncthreads();
        const int gi = s_ticket;
        if (gi >= n_graphs) break;
        const long long e0 = eo_ptr[gi];
        const int e = (int)(eo_ptr[gi + 1] - e0);
        const int n = (int)(node_ptr[gi + 1] - node_ptr[gi]);
        if (e == 0) continue;
        GraphLayout L(n, e);
        unsigned char* slab = scratch + (size_t)blockIdx.x * slab_bytes;
        unsigned char* base = (L.total <= graph_smem_bytes) ? graph_smem : slab;
        GraphView g;
        if (!load_graph<H>(g, base, eo_src + e0, eo_dst + e0, n, e, s_misc, counters)) {
            for (int i = tid; i < e * ESCGNN_RD_SLOTS; i += blockDim.x) rdh[(size_t)e0 * ESCGNN_RD_SLOTS + i] = 0;
            continue;
        }
        // symmetry of the edge multiset (every induced F inherits it): mult(w->b) == mult(b->w)
        bool asym = false;
        for (int w = tid; w < n; w += blockDim.x) {
            const uint32_t ka = g.out_ptr[w], kb = g.out_ptr[w + 1];
            for (uint32_t k = ka; k < kb; ++k) {
                const int b = g.out_adj[k];
                int c1 = 0, c2 = 0;
                for (uint32_t q = ka; q < kb; ++q) c1 += g.out_adj[q] == b;
                for (uint32_t q = g.out_ptr[b]; q < g.out_ptr[b + 1]; ++q) c2 += g.out_adj[q] == w;
                if (c1 != c2) asym = true;
            }
        }
        if (asym) atomicOr(&s_err, ESCGNN_DATA_ASYM);
        __syncthreads();
        if (s_err) {
            if (tid == 0) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)s_err);
            for (int i = tid; i < e * ESCGNN_RD_SLOTS; i += blockDim.x) rdh[(size_t)e0 * ESCGNN_RD_SLOTS + i] = 0;
            continue;
        }
        const long long need = tri(n + 1);
        unsigned err = 0;
        if (need <= warp_cap && n <= sub_stride) {
            // warp per edge
            double* M = mat_region + (size_t)warp * warp_cap;
            uint16_t* sub = sub_region + (size_t)warp * sub_stride;
            for (int ed = warp; ed < e; ed += nw) {
                const int u = (int)eo_src[e0 + ed], v = (int)eo_dst[e0 + ed];
                err |= rd_edge<H, false>(g, u, v, M, sub, s_hist[warp], nullptr, nullptr);
                if (lane < ESCGNN_RD_SLOTS)
                    rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS + lane] = (uint16_t)s_hist[warp][lane];
                __syncwarp();
            }
        } else {
            // whole CTA per edge; matrix in shared memory when it fits, else in the global slab
            double* M = need <= mat_region_doubles ? mat_region
                                                   : reinterpret_cast<double*>(slab + slab_graph_bytes);
            uint16_t* sub = (n <= nw * sub_stride) ? sub_region
                                                   : reinterpret_cast<uint16_t*>(slab + slab_graph_bytes + align16(need * 8));
            for (int ed = 0; ed < e; ++ed) {
                const int u = (int)eo_src[e0 + ed], v = (int)eo_dst[e0 + ed];
                err |= rd_edge<H, true>(g, u, v, M, sub, s_hist[0], s_red, &s_cnt);
                if (tid < ESCGNN_RD_SLOTS)
                    rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS + tid] = (uint16_t)s_hist[0][tid];
                __syncthreads();
            }
        }
        if (err) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)err);
    }
}

struct RdPlan {
    int64_t mat_doubles, smem, graph_bytes, slab, slab_graph;
    int sub_stride;
};

static RdPlan plan_rd(int64_t max_nodes, int64_t max_edges, int smem_optin) {
    constexpr int kNw = 8;
    RdPlan p;
    const int64_t need_graph = GraphLayout(max_nodes, max_edges).total;
    const int64_t t = tri(max_nodes + 1);
    const int64_t budget = (int64_t)smem_optin - 2048;
    p.sub_stride = (int)((max_nodes + 7) & ~int64_t(7));
    if (p.sub_stride > 1024) p.sub_stride = 1024;
    int64_t sub_bytes = (int64_t)kNw * p.sub_stride * 2;
    p.graph_bytes = need_graph <= 48 * 1024 ? need_graph : 0;
    int64_t avail = budget - sub_bytes - p.graph_bytes - 16;
    if (t * 8 * kNw <= 96 * 1024 && t * 8 * kNw <= avail) p.mat_doubles = t * kNw;      // warp mode for every graph
    else if (t * 8 <= avail) p.mat_doubles = ((t + kNw - 1) / kNw) * kNw;               // CTA mode, on chip
    else p.mat_doubles = (64 * 1024 / 8 / kNw) * kNw;                                    // small graphs stay on chip
    p.smem = align16(p.mat_doubles * 8 + sub_bytes) + p.graph_bytes;
    p.slab_graph = align16(need_graph);
    p.slab = 0;
    if (p.graph_bytes < need_graph || t > p.mat_doubles || max_nodes > (int64_t)kNw * p.sub_stride)
        p.slab = p.slab_graph + align16(t * 8) + align16(max_nodes * 2);
    return p;
}

template <int H>
static int launch_rd(const int64_t* eo_src, const int64_t* eo_dst, const int64_t* eo_ptr, const int64_t* node_ptr,
                     int64_t n_graphs, uint16_t* rdh, unsigned long long* counters, int64_t max_nodes,
                     int64_t max_edges, void* scratch, int64_t scratch_bytes, cudaStream_t st) {
    constexpr int kThreads = 256;
    int dev = 0, sms = 148, smem_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const RdPlan p = plan_rd(max_nodes, max_edges, smem_optin);
    auto kern = ego_rd_kernel<H>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (err != cudaSuccess) return (int)err;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, (size_t)p.smem);
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > n_graphs) grid = n_graphs;
    if (p.slab) {
        if (scratch == nullptr || scratch_bytes < p.slab) return ESCGNN_ERR_BAD_ARG;
        const int64_t fit = scratch_bytes / p.slab;
        if (grid > fit) grid = fit;
    }
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kThreads, (size_t)p.smem, st>>>(eo_src, eo_dst, eo_ptr, node_ptr, (int)n_graphs, rdh,
                                                            counters, (long long)p.graph_bytes,
                                                            (long long)p.mat_doubles, p.sub_stride,
                                                            (unsigned char*)scratch, (long long)p.slab,
                                                            (long long)p.slab_graph);
    return (int)cudaGetLastError();
}

}  // namespace escgnn

using namespace escgnn;

extern "C" {

int64_t escgnn_encode_rd_scratch_bytes(int64_t max_nodes, int64_t max_edges, int h) {
    (void)h;
    const RdPlan p = plan_rd(max_nodes, max_edges, 227 * 1024);
    return p.slab * 148 * 2;
}

int escgnn_encode_rd(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                     const int64_t* d_node_ptr, int64_t n_graphs, int h, uint16_t* d_rdh,
                     unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                     int64_t scratch_bytes, void* stream) {
    if (n_graphs <= 0) return 0;
    if (h < 1 || h > 4) return ESCGNN_ERR_BAD_ARG;
    if (max_nodes > 1023 || max_edges > 65535) return ESCGNN_ERR_TOO_LARGE;   // rd is a molecule-scale feature (m <= 4 rows/thread)
    cudaStream_t st = (cudaStream_t)stream;
#define ESC_GO(H) launch_rd<H>(d_eo_src, d_eo_dst, d_eo_ptr, d_node_ptr, n_graphs, d_rdh, d_counters, max_nodes, \
                               max_edges, d_scratch, scratch_bytes, st)
    switch (h) {
        case 1: return ESC_GO(1);
        case 2: return ESC_GO(2);
        case 3: return ESC_GO(3);
        default: return ESC_GO(4);
    }
#undef ESC_GO
}

}  // extern "C"
