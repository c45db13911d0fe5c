// K1b `ego_rd`: resistance-distance histogram of every directed edge's union subgraph (E5).
//
// Both directions of an undirected edge share S and F, so each unordered pair is solved once (full selected inverse).
// Replaces /root/reference/utils_edge_efficient.py:92-107 (float32 scipy laplacian -> pinv per edge) and :130-131
// (one_hot(rd.long(), 100)) under parity policy E5: float64 arithmetic, bin = trunc((float)rd).
//
// For a symmetric, connected (S, F):  u != v: rd(w) = R(u,w) = [(L with row/col u removed)^-1]_ww, rd(u) = 0;
// u == v (phantom root): rd(w) = [pinv(L_ball)]_ww, rd(phantom) = 0; pinv from the same grounded inverse (see rd_pair).
// Both are "diagonal of the inverse of an SPD matrix": LDL^T in place on a packed lower triangle held in shared
// memory, then the Takahashi recurrence Z_ij = delta_ij/D_j - sum_{k>j} Z_ik L_kj run from the last column back.
// Each thread owns matrix rows (row i -> thread i mod group), so assembly and both sweeps need no atomics.
// Small graphs: one warp per edge; larger ones: the whole CTA works on one edge; too large for shared memory: the
// matrix lives in a per-CTA global slab.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "graph_smem.cuh"
#include "rd_fast.cuh"

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

// 64-bit "which rows k of column c are non-zero" mask for the warp-per-pair solver (m <= 64): rows lane and lane + 32
__device__ __forceinline__ unsigned long long column_mask(const double* M, int m, int c, int first_row, int lane) {
    const int r0 = lane, r1 = lane + 32;
    const bool a = r0 >= first_row && r0 < m && M[tidx(r0, c)] != 0.0;
    const bool b = r1 >= first_row && r1 < m && M[tidx(r1, c)] != 0.0;
    return (unsigned long long)__ballot_sync(kFull, a) | ((unsigned long long)__ballot_sync(kFull, b) << 32);
}

// One unordered pair {u, v} (or the self-loop edge u == v): fills hist_u[ESCGNN_RD_SLOTS] = rd histogram of the directed edge
// (u, v) and hist_v = that of (v, u) -- both directions share S and F, only the root differs.
//   u != v : ground u.  M = L without row/col u (SPD, and as sparse as the molecule), Z = M^-1:
//            R(u,w) = Z_ww,  R(v,w) = Z_vv + Z_ww - 2 Z_vw,  R(v,u) = Z_vv.
//   u == v : phantom root: ground u as well; rd(w) = pinv(L)_ww = Z_ww - 2 r_w / mm + S / mm^2, rd(phantom) = 0.
// LDL^T in place on the packed lower triangle, then the Takahashi recurrence Z_ij = delta_ij/D_j - (1/D_j) sum_{k>j} Z_ik W_kj
// from the last column back.  The warp-per-pair variant walks only the NON-ZERO entries of each factor column (ballot
// masks): molecular graphs are nearly trees, so a column holds 1-3 entries instead of m/2.
// Returns error bits (uniform in the group).
// Pendant trees are peeled before the linear algebra (warp mode, u != v): a node whose only remaining neighbour p is joined to it by
// a single edge adds exactly 1 to every resistance measured from the rest of the graph, R(r, w) = R(r, p) + 1, so leaves are
// removed round by round (u and v are kept) and only the remaining core -- the rings and the paths between them, often just {u, v} for
// a molecule -- goes through LDL^T / Takahashi.  Eliminating a unit-weight leaf leaves the Laplacian of the induced core
// (Schur complement), and the added path lengths are exact integers in fp64, so the histogram bins are unchanged.
// aux (per warp, may be nullptr = no peeling): rnd[stride] u8 (0 = core, else the round it was peeled in), dep[stride] u8 (path
// length to the core), par[stride] u16, anc[stride] u16 (the core node the pendant tree hangs from).
template <int H, bool kCta>
__device__ unsigned rd_pair(const GraphView& g, int u, int v, double* M, uint16_t* sub, int* hist_u, int* hist_v, double* s_red,
                            int* s_cnt, unsigned char* aux = nullptr, int aux_stride = 0) {
    const int gt = kCta ? threadIdx.x : (threadIdx.x & 31);        // thread id inside the group
    const int gn = kCta ? blockDim.x : 32;                         // group size
    const int lane = threadIdx.x & 31;
    const int n = g.n, rw = g.rw;
    const uint32_t* rowU = g.dist + (size_t)u * rw;
    const uint32_t* rowV = g.dist + (size_t)v * rw;
    const bool phantom = u == v;
    if (gt < ESCGNN_RD_SLOTS) { hist_u[gt] = 0; hist_v[gt] = 0; }
    // ---- peel pendant trees
    // (worth it on sparse, molecule-like graphs only: with ~4 edges per node almost nothing is pendant and the scan is pure overhead)
    const bool peel = !kCta && aux != nullptr && !phantom && n <= 64 && 10 * g.e <= 33 * n;
    uint8_t* rnd = aux;
    uint8_t* dep = aux + aux_stride;
    uint16_t* par = reinterpret_cast<uint16_t*>(aux + 2 * aux_stride);
    uint16_t* anc = par + aux_stride;
    if (peel) {
        for (int w = lane; w < n; w += 32) { rnd[w] = 0; dep[w] = 0; }
        __syncwarp();
        int rounds = 0;
        for (int r = 1; r < 250; ++r) {
            int cand[2] = {-1, -1}, cpar[2] = {0, 0};
            int slot = 0;
            for (int w = lane; w < n; w += 32, ++slot) {
                const uint32_t du = nib(rowU, w), dv = nib(rowV, w);
                if ((du == kFar && dv == kFar) || rnd[w] != 0 || w == u || w == v) continue;
                int b0 = -1, mult = 0;
                bool other = false;
                const uint32_t ka = g.out_ptr[w], kb = g.out_ptr[w + 1];
                for (uint32_t k = ka; k < kb; ++k) {
                    const int b = g.out_adj[k];
                    if (b == w) continue;
                    const uint32_t bu = nib(rowU, b), bv = nib(rowV, b);
                    if (!((du != kFar && bu != kFar) || (dv != kFar && bv != kFar))) continue;      // edge not in F
                    if (rnd[b] != 0) continue;                                                      // neighbour already peeled
                    if (b0 < 0) { b0 = b; mult = 1; } else if (b == b0) ++mult; else other = true;
                }
                if (b0 >= 0 && !other && mult == 1) { cand[slot] = w; cpar[slot] = b0; }            // a unit-weight leaf
            }
            if (!__any_sync(kFull, cand[0] >= 0 || cand[1] >= 0)) break;
            __syncwarp();                                        // every lane has finished reading this round's state
            #pragma unroll
            for (int q = 0; q < 2; ++q)
                if (cand[q] >= 0) { rnd[cand[q]] = (uint8_t)r; par[cand[q]] = (uint16_t)cpar[q]; }
            __syncwarp();
            rounds = r;
        }
        for (int r = rounds; r >= 1; --r) {                      // parents were peeled later (or never): latest rounds first
            for (int w = lane; w < n; w += 32)
                if (rnd[w] == r) {
                    const int pp = par[w];
                    if (rnd[pp] == 0) { anc[w] = (uint16_t)pp; dep[w] = 1; }
                    else { anc[w] = anc[pp]; dep[w] = (uint8_t)(dep[pp] + 1); }
                }
            __syncwarp();
        }
    }
    // ---- matrix index of every member of S (node order); u is grounded (no row) unless phantom
    int m = 0;
    if (!kCta || threadIdx.x < 32) {
        for (int w0 = 0; w0 < n; w0 += 32) {
            const int w = w0 + lane;
            bool in = false;
            if (w < n) in = (nib(rowU, w) != kFar || nib(rowV, w) != kFar) && w != u && (!peel || rnd[w] == 0);
            const unsigned b = __ballot_sync(kFull, in);
            if (w < n) sub[w] = in ? (uint16_t)(m + __popc(b & ((1u << lane) - 1))) : (uint16_t)0xffff;
            m += __popc(b);
        }
        if (kCta && lane == 0) *s_cnt = m;
    }
    group_sync<kCta>();
    if (kCta) m = *s_cnt;
    for (int t = gt; t < (int)tri(m); t += gn) M[t] = 0.0;
    group_sync<kCta>();
    // ---- assemble: the thread owning node w writes row sub[w] (diagonal = degree in F without loops)
    for (int w = gt; w < n; w += gn) {
        const int i = sub[w];
        if (i == 0xffff) continue;
        const uint32_t du = nib(rowU, w), dv = nib(rowV, w);
        double deg = 0.0;
        double* rowi = M + tidx(i, 0);
        const uint32_t ka = g.out_ptr[w], kb = g.out_ptr[w + 1];
        for (uint32_t k = ka; k < kb; ++k) {
            const int b = g.out_adj[k];
            if (b == w) continue;                                      // scipy laplacian ignores loops
            const uint32_t bu = nib(rowU, b), bv = nib(rowV, b);
            if (!((du != kFar && bu != kFar) || (dv != kFar && bv != kFar))) continue;
            if (peel && rnd[b] != 0) continue;                         // peeled leaves are gone from the core's Laplacian
            deg += 1.0;
            const int j = sub[b];
            if (j != 0xffff && j < i) rowi[j] -= 1.0;
        }
        rowi[i] += deg;
    }
    group_sync<kCta>();
    // ---- LDL^T, column k keeps W_ik = L_ik * D_k (unscaled), diagonal keeps D_k
    bool bad = false;
    for (int k = 0; k < m; ++k) {
        const double d = M[tidx(k, k)];
        if (!(d > 1e-12)) { bad = true; break; }                      // uniform: every thread reads the same value
        const double invd = 1.0 / d;
        if (!kCta && m <= 64) {
            const unsigned long long nz = column_mask(M, m, k, k + 1, lane);       // rows j > k with W_jk != 0
            const bool dense_col = 2 * __popcll(nz) > m - k;                        // uniform: walk the column contiguously
            for (int i = k + 1 + ((lane - k - 1) & 31); i < m; i += 32) {             // rows owned by this lane (i % 32 == lane)
                if (!((nz >> i) & 1ull)) continue;
                double* rowi = M + tidx(i, 0);
                const double f = rowi[k] * invd;
                if (dense_col) {
                    int p = tidx(k + 1, k);
                    for (int j = k + 1; j <= i; ++j) { rowi[j] -= f * M[p]; p += j + 1; }
                } else {
                    unsigned long long todo = nz & ((i >= 63 ? ~0ull : ((1ull << (i + 1)) - 1ull)));   // j <= i
                    while (todo) {
                        const int jj = __ffsll((long long)todo) - 1;
                        todo &= todo - 1;
                        rowi[jj] -= f * M[tidx(jj, k)];
                    }
                }
            }
        } else {
            for (int i = k + 1 + gt; i < m; i += gn) {
                double* rowi = M + tidx(i, 0);
                const double f = rowi[k] * invd;
                if (f != 0.0) {
                    int p = tidx(k + 1, k);                                // M[j][k], advancing one row at a time
                    for (int j = k + 1; j <= i; ++j) { rowi[j] -= f * M[p]; p += j + 1; }
                }
            }
        }
        group_sync<kCta>();
    }
    if (bad) return ESCGNN_DATA_RD;
    // ---- Takahashi, columns from the last to the first; Z overwrites the factor column by column
    for (int j = m - 1; j >= 0; --j) {
        const double invd = 1.0 / M[tidx(j, j)];
        double part = 0.0;                      // sum_k W_kj * Z_kj over the rows this thread owns
        double zmine[4];                        // up to 4 rows per thread in flight (m <= 4 * group size)
        int cnt = 0;
        if (!kCta && m <= 64) {
            const unsigned long long nz = column_mask(M, m, j, j + 1, lane);       // rows k > j with W_kj != 0
            const bool dense_col = 2 * __popcll(nz) > m - j;
            for (int i = j + 1 + ((lane - j - 1) & 31); i < m; i += 32) {
                const double* rowi = M + tidx(i, 0);
                double acc = 0.0;
                if (dense_col) {
                    int pj = tidx(j + 1, j);
                    for (int k = j + 1; k <= i; ++k) { acc += rowi[k] * M[pj]; pj += k + 1; }
                    int pi = tidx(i + 1, i);
                    for (int k = i + 1; k < m; ++k) { acc += M[pi] * M[pj]; pi += k + 1; pj += k + 1; }
                } else {
                    unsigned long long todo = nz;
                    while (todo) {
                        const int k = __ffsll((long long)todo) - 1;
                        todo &= todo - 1;
                        const double z = k <= i ? rowi[k] : M[tidx(k, i)];
                        acc += z * M[tidx(k, j)];
                    }
                }
                const double zij = -acc * invd;
                part += rowi[j] * zij;
                if (cnt < 4) zmine[cnt] = zij;
                ++cnt;
            }
        } else {
            for (int i = j + 1 + gt; i < m; i += gn) {
                const double* rowi = M + tidx(i, 0);
                double acc = 0.0;
                int pj = tidx(j + 1, j);                                   // W[k][j]
                for (int k = j + 1; k <= i; ++k) { acc += rowi[k] * M[pj]; pj += k + 1; }          // Z_ik with k <= i: row i
                int pi = tidx(i + 1, i);                                   // Z_ki with k > i: column i
                for (int k = i + 1; k < m; ++k) { acc += M[pi] * M[pj]; pi += k + 1; pj += k + 1; }
                const double zij = -acc * invd;
                part += rowi[j] * zij;
                if (cnt < 4) zmine[cnt] = zij;
                ++cnt;
            }
        }
        const double s = group_sum<kCta>(part, s_red);     // (also orders the reads of column j before its overwrite)
        group_sync<kCta>();
        cnt = 0;
        if (!kCta && m <= 64) { for (int i = j + 1 + ((lane - j - 1) & 31); i < m; i += 32) { M[tidx(i, j)] = zmine[cnt < 4 ? cnt : 3]; ++cnt; } }
        else { for (int i = j + 1 + gt; i < m; i += gn) { M[tidx(i, j)] = zmine[cnt < 4 ? cnt : 3]; ++cnt; } }
        if (gt == 0) M[tidx(j, j)] = invd - invd * s;
        group_sync<kCta>();
    }
    // ---- bin: trunc((float)rd)   (torch.FloatTensor(...) then .long(), utils_edge_efficient.py:105,131)
    unsigned err = 0;
    const int iv = phantom ? 0 : sub[v];
    const double zvv = phantom ? 0.0 : M[tidx(iv, iv)];
    // phantom root: pinv(L_ball) from the SAME grounded inverse (any symmetric generalised inverse G of L gives L^+ = H G H with the
    // centring H = I - J/mm): L^+_ww = Z_ww - (2/mm) r_w + S/mm^2 with r_w = sum_b Z_wb, S = sum_w r_w, and Z_u. = 0 for the ground
    // u -- the sparse grounded system replaces the dense (L + J/mm) one the first version factorised for every node of the graph.
    double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0, rs3 = 0.0, S = 0.0;       // row sums of the nodes this thread owns (<= 4: registers)
    const double mm = (double)(m + 1);
    if (phantom) {
        double part = 0.0;
        int cnt = 0;
        for (int w = gt; w < n; w += gn, ++cnt) {
            const int i = sub[w];
            if (i == 0xffff) continue;
            double r = 0.0;
            int pr = tidx(i, 0);
            for (int b = 0; b <= i; ++b) r += M[pr + b];
            int pc = tidx(i + 1, i);
            for (int b = i + 1; b < m; ++b) { r += M[pc]; pc += b + 1; }
            if (cnt == 0) rs0 = r; else if (cnt == 1) rs1 = r; else if (cnt == 2) rs2 = r; else rs3 = r;
            part += r;
        }
        S = group_sum<kCta>(part, s_red);
    }
    int cnt_w = 0;
    for (int w = gt; w < n; w += gn, ++cnt_w) {
        const int i = sub[w];
        if (i == 0xffff) {
            if (peel && rnd[w] != 0) {                               // pendant node: its anchor's resistances plus the path length
                const int a = anc[w];
                const double d = (double)dep[w];
                const int ia = a == u ? 0 : sub[a];
                const double zaa = a == u ? 0.0 : M[tidx(ia, ia)];
                double rv;
                if (a == v) rv = d;
                else if (a == u) rv = zvv + d;
                else rv = zvv + zaa - 2.0 * (ia >= iv ? M[tidx(ia, iv)] : M[tidx(iv, ia)]) + d;
                const int bu = (int)truncf((float)(zaa + d)), bv = (int)truncf((float)rv);
                if (bu < 0 || bu >= ESCGNN_RD_SLOTS || bv < 0 || bv >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD;
                else { atomicAdd(&hist_u[bu], 1); atomicAdd(&hist_v[bv], 1); }
            }
            continue;
        }
        const double zww = M[tidx(i, i)];
        if (phantom) {
            const double rw = cnt_w == 0 ? rs0 : cnt_w == 1 ? rs1 : cnt_w == 2 ? rs2 : rs3;
            const int b = (int)truncf((float)(zww - 2.0 * rw / mm + S / (mm * mm)));
            if (b < 0 || b >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD; else atomicAdd(&hist_u[b], 1);
        } else {
            const double zvw = i >= iv ? M[tidx(i, iv)] : M[tidx(iv, i)];
            const int bu = (int)truncf((float)zww);                                        // R(u, w), u grounded
            const int bv = w == v ? 0 : (int)truncf((float)(zvv + zww - 2.0 * zvw));        // R(v, w)
            if (bu < 0 || bu >= ESCGNN_RD_SLOTS || bv < 0 || bv >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD;
            else { atomicAdd(&hist_u[bu], 1); atomicAdd(&hist_v[bv], 1); }
        }
    }
    if (gt == 0) {
        if (phantom) {
            atomicAdd(&hist_u[0], 1);                      // the phantom root itself (rd = 0)
            const int b = (int)truncf((float)(S / (mm * mm)));       // the real root u: the grounded node, Z_u. = 0
            if (b < 0 || b >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD; else atomicAdd(&hist_u[b], 1);
        } else {
            atomicAdd(&hist_u[0], 1);                      // rd_u(u) = 0
            const int b = (int)truncf((float)zvv);          // rd_v(u) = R(v, u) = Z_vv
            if (b < 0 || b >= ESCGNN_RD_SLOTS) err = ESCGNN_DATA_RD; else atomicAdd(&hist_v[b], 1);
        }
    }
    group_sync<kCta>();
    return err;
}

template <int H>
__global__ void __launch_bounds__(256, 3)          // three CTAs per SM is what the shared-memory plan of molecule-sized batches allows
ego_rd_kernel(const int64_t* __restrict__ eo_src, const int64_t* __restrict__ eo_dst,
              const int64_t* __restrict__ eo_ptr, const int64_t* __restrict__ node_ptr, int n_graphs,
              uint16_t* __restrict__ rdh, unsigned long long* counters, long long graph_smem_bytes,
              long long mat_region_doubles, int sub_stride, unsigned char* scratch, long long slab_bytes,
              long long slab_graph_bytes, int parts, int aux_stride, int n_lo, int n_hi, int fast_done) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ticket;
    __shared__ int s_misc[2];
    __shared__ int s_hist[8][2][ESCGNN_RD_SLOTS];
    __shared__ double s_red[8];
    __shared__ int s_cnt;
    __shared__ unsigned s_err;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (fast_done && counters[ESCGNN_CTR_RD_DECLINED] == 0ull) return;       // the cycle-space kernel solved every edge of the batch
    double* mat_region = reinterpret_cast<double*>(smem);
    uint16_t* sub_region = reinterpret_cast<uint16_t*>(smem + mat_region_doubles * 8);
    unsigned char* aux_region = smem + align16(mat_region_doubles * 8 + (long long)nw * sub_stride * 2);      // aux_stride == 0: none
    unsigned char* graph_smem = aux_region + align16((long long)nw * aux_stride * 6);
    const long long warp_cap = mat_region_doubles / nw;

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_ticket = (int)atomicAdd(&counters[ESCGNN_CTR_TICKET_RD], 1ull); s_err = 0; }
        __syncthreads();
        // small batches: `parts` CTAs share a graph (each rebuilds the cheap per-graph state and solves every parts-th pair
        // system), because a graph's pair systems are the only parallelism there is when the batch does not fill the machine
        const int gi = s_ticket / parts, part = s_ticket % parts;
        if (gi >= n_graphs) break;
        const long long e0 = eo_ptr[gi];
        const int e = (int)(eo_ptr[gi + 1] - e0);
        const int n = (int)(node_ptr[gi + 1] - node_ptr[gi]);
        if (e == 0 || n < n_lo || n > n_hi) continue;             // (size classes: each launch takes the graphs of its node range)
        if (fast_done) {
            // the cycle-space kernel (ego_rd_fast_kernel below) ran first: only the edges it marked are still to be solved
            int any = 0;
            for (int i = tid; i < e; i += blockDim.x) any |= rdh[(size_t)(e0 + i) * ESCGNN_RD_SLOTS] == rdfast::kSentinel;
            if (!__syncthreads_or(any)) continue;
        }
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
            for (int ed = warp + nw * part; ed < e; ed += nw * parts) {
                const int u = (int)eo_src[e0 + ed], v = (int)eo_dst[e0 + ed];
                if (u > v) continue;                                   // handled together with its reverse edge (v, u)
                // the first occurrence of (u, v) does the work for every copy of (u, v) and (v, u) in the edge list
                bool first = true;
                for (int q0 = 0; q0 < ed && first; q0 += 32) {
                    const int q = q0 + lane;
                    const bool dup = q < ed && (int)eo_src[e0 + q] == u && (int)eo_dst[e0 + q] == v;
                    if (__any_sync(kFull, dup)) first = false;
                }
                if (!first) continue;
                if (fast_done && rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS] != rdfast::kSentinel) continue;     // solved by the fast path
                err |= rd_pair<H, false>(g, u, v, M, sub, s_hist[warp][0], s_hist[warp][1], nullptr, nullptr,
                                         aux_stride ? aux_region + (size_t)warp * aux_stride * 6 : nullptr, aux_stride);
                for (int q0 = 0; q0 < e; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < e) {
                        const int a = (int)eo_src[e0 + q], b = (int)eo_dst[e0 + q];
                        const int which = (a == u && b == v) ? 0 : ((a == v && b == u) ? 1 : -1);
                        if (which >= 0)
                            for (int t = 0; t < ESCGNN_RD_SLOTS; ++t)
                                rdh[(size_t)(e0 + q) * ESCGNN_RD_SLOTS + t] = (uint16_t)s_hist[warp][which][t];
                    }
                }
                __syncwarp();
            }
        } else {
            // whole CTA per pair; matrix in shared memory when it fits, else in the global slab
            double* M = need <= mat_region_doubles ? mat_region
                                                   : reinterpret_cast<double*>(slab + slab_graph_bytes);
            uint16_t* sub = (n <= nw * sub_stride) ? sub_region
                                                   : reinterpret_cast<uint16_t*>(slab + slab_graph_bytes + align16(need * 8));
            for (int ed = part; ed < e; ed += parts) {
                const int u = (int)eo_src[e0 + ed], v = (int)eo_dst[e0 + ed];
                if (u > v) continue;
                bool first = true;
                for (int q = tid; q < ed; q += blockDim.x)
                    if ((int)eo_src[e0 + q] == u && (int)eo_dst[e0 + q] == v) first = false;
                if (__syncthreads_or(!first)) continue;
                if (fast_done && rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS] != rdfast::kSentinel) continue;
                err |= rd_pair<H, true>(g, u, v, M, sub, s_hist[0][0], s_hist[0][1], s_red, &s_cnt);
                for (int q = tid; q < e; q += blockDim.x) {
                    const int a = (int)eo_src[e0 + q], b = (int)eo_dst[e0 + q];
                    const int which = (a == u && b == v) ? 0 : ((a == v && b == u) ? 1 : -1);
                    if (which >= 0)
                        for (int t = 0; t < ESCGNN_RD_SLOTS; ++t)
                            rdh[(size_t)(e0 + q) * ESCGNN_RD_SLOTS + t] = (uint16_t)s_hist[0][which][t];
                }
                __syncthreads();
            }
        }
        if (err) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)err);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path (rd_fast.cuh): a warp per graph, a lane per unordered pair / self-loop edge, cycle-space resistance distances.
// Takes symmetric simple graphs of at most `nf` nodes / `ef` directed edges; per pair at most kFastCmax independent cycles inside
// the ego-net.  Whatever it declines gets the sentinel in slot 0 of its histogram and is solved by ego_rd_kernel afterwards.
constexpr int kFastCmax = 4;
constexpr int kFastSmall = 41;        // size classes, as for the general solver: molecule-sized graphs keep a small workspace

__device__ __forceinline__ unsigned long long spread4(uint32_t x) {         // four 8-bit counts -> four uint16
    return (unsigned long long)(x & 0xffu) | ((unsigned long long)((x >> 8) & 0xffu) << 16) |
           ((unsigned long long)((x >> 16) & 0xffu) << 32) | ((unsigned long long)(x >> 24) << 48);
}
__device__ __forceinline__ void store_hist(uint16_t* rdh, long long edge, const rdfast::Hist& h) {
    unsigned long long* out = reinterpret_cast<unsigned long long*>(rdh + (size_t)edge * ESCGNN_RD_SLOTS);     // 24-byte rows
    out[0] = spread4((uint32_t)h.lo);
    out[1] = spread4((uint32_t)(h.lo >> 32));
    out[2] = spread4(h.hi);
}

// kCta = false: a warp per graph (molecule-sized graphs: the pairs of one graph fill one or two chunks of 32 lanes);
// kCta = true: a CTA per graph -- the graph's CSR / distance matrix are built once by all warps (a thread per BFS root) and the
// chunks of 32 pair systems are dealt to the warps, so a 120-node graph does not serialise 8 chunks on one warp.
template <int H, bool kCta>
__global__ void __launch_bounds__(256, kCta ? 1 : 3)
ego_rd_fast_kernel(const int64_t* __restrict__ eo_src, const int64_t* __restrict__ eo_dst, const int64_t* __restrict__ eo_ptr,
                   const int64_t* __restrict__ node_ptr, int n_graphs, uint16_t* __restrict__ rdh, unsigned long long* counters,
                   int nf, int ef, int graph_part, int lane_part, int n_lo, int n_hi) {
    static_assert(ESCGNN_RD_SLOTS == rdfast::kSlots, "histogram width");
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_share[4];                                    // CTA mode: ticket, loops, pairs, n_pair
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gt = kCta ? (int)threadIdx.x : lane, gn = kCta ? (int)blockDim.x : 32;      // thread id / size of the group owning a graph
    auto gsync = [&]() { if (kCta) __syncthreads(); else __syncwarp(); };
    auto gany = [&](bool x) -> bool { return kCta ? __syncthreads_or(x) != 0 : __any_sync(kFull, x) != 0; };
    rdfast::Ws ws;
    if (kCta) {
        rdfast::carve_graph(ws, smem, nf, ef);
        rdfast::carve_lanes(ws, smem + graph_part + (size_t)warp * lane_part, nf);
    } else {
        unsigned char* base = smem + (size_t)warp * (graph_part + lane_part);
        rdfast::carve_graph(ws, base, nf, ef);
        rdfast::carve_lanes(ws, base + graph_part, nf);
    }
    for (;;) {
        int gi = 0;
        if (kCta) {
            __syncthreads();                                      // everyone is done with the previous graph
            if (threadIdx.x == 0) { s_share[0] = (int)atomicAdd(&counters[ESCGNN_CTR_TICKET_RD], 1ull); s_share[1] = 0; }
            __syncthreads();
            gi = s_share[0];
        } else {
            if (lane == 0) gi = (int)atomicAdd(&counters[ESCGNN_CTR_TICKET_RD], 1ull);
            gi = __shfl_sync(kFull, gi, 0);
        }
        if (gi >= n_graphs) break;
        const long long e0 = eo_ptr[gi];
        const int e = (int)(eo_ptr[gi + 1] - e0);
        const int n = (int)(node_ptr[gi + 1] - node_ptr[gi]);
        if (e == 0 || n < n_lo || n > n_hi) continue;              // another launch's size class
        const int64_t* src = eo_src + e0;
        const int64_t* dst = eo_dst + e0;
        gsync();
        bool bad = n > nf || e > ef || n < 1;                      // (uniform)
        if (!bad) {
            // ---- CSR by source (counting sort; the cursors alias the distance matrix, which is initialised afterwards)
            ws.n = n; ws.e = e; ws.rws = rdfast::row_stride(n);
            uint32_t* cnt = ws.dist;
            for (int i = gt; i <= n; i += gn) cnt[i] = 0;
            gsync();
            int loops = 0;
            for (int i = gt; i < e; i += gn) {
                const long long s = src[i], t = dst[i];
                if (s < 0 || s >= n || t < 0 || t >= n) bad = true;
                else { atomicAdd(&cnt[s + 1], 1u); loops += s == t; }
            }
            if (kCta) {
                if (loops) atomicAdd(&s_share[1], loops);
                __syncthreads();
                loops = s_share[1];
            } else {
                #pragma unroll
                for (int d = 16; d; d >>= 1) loops += __shfl_xor_sync(kFull, loops, d);
            }
            // cyclomatic number of the whole graph (if connected): an upper bound for every ego-net.  Dense graphs (count_cycle-like,
            // ~1.66 edges per node) would be declined pair by pair anyway: skip their setup
            bad |= (e - loops) / 2 - n + 1 > 2 * kFastCmax;
            bad = gany(bad);
        }
        if (!bad) {
            if (warp == 0 || !kCta) {
                uint32_t carry = 0;
                for (int i0 = 0; i0 <= n; i0 += 32) {
                    const int i = i0 + lane;
                    uint32_t v = i <= n ? ws.dist[i] : 0u;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(kFull, v, d);
                        if (lane >= d) v += t;
                    }
                    if (i <= n) ws.optr[i] = (uint16_t)(v + carry);
                    carry += __shfl_sync(kFull, v, 31);
                }
            }
            gsync();
            uint32_t* cur = ws.dist + (n + 1);
            for (int i = gt; i < n; i += gn) cur[i] = ws.optr[i];
            gsync();
            for (int i = gt; i < e; i += gn) {
                const int s = (int)src[i], t = (int)dst[i];
                const uint32_t slot = atomicAdd(&cur[s], 1u);
                ws.oadj[slot] = (uint8_t)t;
                ws.oeid[slot] = (uint16_t)i;
            }
            gsync();
            // ---- ascending adjacency (deterministic trees), simple, symmetric
            for (int w = gt; w < n; w += gn) {
                const int ka = ws.optr[w], kb = ws.optr[w + 1];
                for (int k = ka + 1; k < kb; ++k) {
                    const uint8_t t = ws.oadj[k];
                    const uint16_t id = ws.oeid[k];
                    int j = k - 1;
                    while (j >= ka && ws.oadj[j] > t) { ws.oadj[j + 1] = ws.oadj[j]; ws.oeid[j + 1] = ws.oeid[j]; --j; }
                    ws.oadj[j + 1] = t; ws.oeid[j + 1] = id;
                }
                for (int k = ka + 1; k < kb; ++k) bad |= ws.oadj[k] == ws.oadj[k - 1];        // multi-edge
            }
            gsync();
            for (int w = gt; w < n; w += gn) {
                const int ka = ws.optr[w], kb = ws.optr[w + 1];
                for (int k = ka; k < kb; ++k) {
                    const int b = ws.oadj[k];
                    bool back = false;
                    for (int q = ws.optr[b]; q < ws.optr[b + 1]; ++q) back |= ws.oadj[q] == w;
                    bad |= !back;                                                            // asymmetric
                }
            }
            bad = gany(bad);
        }
        if (bad) {                                                 // left to the general solver (which also raises the data errors)
            for (int i = gt; i < e; i += gn) rdh[(size_t)(e0 + i) * ESCGNN_RD_SLOTS] = rdfast::kSentinel;
            if (gt == 0) atomicAdd(&counters[ESCGNN_CTR_RD_DECLINED], (unsigned long long)e);
            continue;
        }
        gsync();
        // ---- E2: a thread per root
        for (int i = gt; i < n * ws.rws; i += gn) ws.dist[i] = 0xffffffffu;
        gsync();
        for (int r = gt; r < n; r += gn) rdfast::bfs_root<H>(ws, r);
        // ---- one system per unordered pair / self-loop edge (pairs first, self-loops after them, in separate chunks of 32: the two
        // kinds of system run different code)
        int pairs = 0, n_pair = 0;
        if (warp == 0 || !kCta) {
            for (int kind = 0; kind < 2; ++kind) {
                for (int i0 = 0; i0 < e; i0 += 32) {
                    const int i = i0 + lane;
                    const bool is = i < e && (kind == 0 ? src[i] < dst[i] : src[i] == dst[i]);
                    const unsigned b = __ballot_sync(kFull, is);
                    if (is) ws.list[pairs + __popc(b & ((1u << lane) - 1u))] = (uint16_t)i;
                    pairs += __popc(b);
                }
                if (kind == 0) n_pair = pairs;
            }
            if (kCta && lane == 0) { s_share[2] = pairs; s_share[3] = n_pair; }
        }
        gsync();
        if (kCta) { pairs = s_share[2]; n_pair = s_share[3]; }
        const int c_pair = (n_pair + 31) >> 5, c_all = c_pair + ((pairs - n_pair + 31) >> 5);
        unsigned err = 0;
        for (int ci = kCta ? warp : 0; ci < c_all; ci += kCta ? nw : 1) {
            const int p = (ci < c_pair ? ci * 32 : n_pair + (ci - c_pair) * 32) + lane;
            if (p < (ci < c_pair ? n_pair : pairs)) {
                const int ed = ws.list[p];
                const int u = (int)src[ed], v = (int)dst[ed];
                rdfast::Hist hu, hv;
                const int rc = rdfast::solve_pair<H, kFastCmax>(ws, lane, u, v, hu, hv);
                int rev = -1;
                if (u != v)
                    for (int q = ws.optr[v]; q < ws.optr[v + 1]; ++q) if (ws.oadj[q] == u) rev = ws.oeid[q];
                if (rc == -1) {
                    rdh[(size_t)(e0 + ed) * ESCGNN_RD_SLOTS] = rdfast::kSentinel;
                    if (rev >= 0) rdh[(size_t)(e0 + rev) * ESCGNN_RD_SLOTS] = rdfast::kSentinel;
                    atomicAdd(&counters[ESCGNN_CTR_RD_DECLINED], 1ull);
                } else {
                    if (rc < 0) err = ESCGNN_DATA_RD;
                    store_hist(rdh, e0 + ed, hu);
                    if (rev >= 0) store_hist(rdh, e0 + rev, hv);
                }
            }
            __syncwarp();
        }
        if (err) atomicOr(&counters[ESCGNN_CTR_ERROR], (unsigned long long)err);
    }
}

inline int& rd_fast_enabled() {
    static int on = 1;
    return on;
}

template <int H, bool kCta>
static int launch_rd_fast_class(const int64_t* eo_src, const int64_t* eo_dst, const int64_t* eo_ptr, const int64_t* node_ptr,
                                int64_t n_graphs, uint16_t* rdh, unsigned long long* counters, int nf, int ef, int n_lo, int n_hi,
                                int sms, cudaStream_t st) {
    const int warps = 8;
    const int64_t gp = (rdfast::graph_bytes(nf, ef) + 15) & ~int64_t(15), lp = (rdfast::lane_bytes(nf, kFastCmax) + 15) & ~int64_t(15);
    const size_t smem = (size_t)(kCta ? gp + warps * lp : warps * (gp + lp));
    auto kern = ego_rd_fast_kernel<H, kCta>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem);
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)sms * occ;
    if (encoder_grid_cap() > 0 && grid > encoder_grid_cap()) grid = encoder_grid_cap();
    const int64_t need = kCta ? n_graphs : (n_graphs + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, warps * 32, smem, st>>>(eo_src, eo_dst, eo_ptr, node_ptr, (int)n_graphs, rdh, counters, nf, ef, (int)gp, (int)lp,
                                                   n_lo, n_hi);
    return (int)cudaGetLastError();
}

template <int H>
static int launch_rd_fast(const int64_t* eo_src, const int64_t* eo_dst, const int64_t* eo_ptr, const int64_t* node_ptr,
                          int64_t n_graphs, uint16_t* rdh, unsigned long long* counters, int64_t max_nodes, int64_t max_edges,
                          cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // class 0: molecule-sized graphs (<= kFastSmall nodes), a warp per graph; class 1 (only when the batch holds larger graphs): up to
    // 128 nodes, a CTA per graph.  The last class owns every graph above its range too (it marks them for the general solver).
    int rc = (int)cudaMemsetAsync(counters + ESCGNN_CTR_RD_DECLINED, 0, sizeof(unsigned long long), st);
    if (rc != 0) return rc;
    const bool two = max_nodes > kFastSmall;
    const int nf0 = (int)(max_nodes < kFastSmall ? (max_nodes < 8 ? 8 : max_nodes) : kFastSmall);
    const int64_t cap0 = 8 * kFastSmall;
    const int ef0 = (int)(max_edges < cap0 ? (max_edges < 16 ? 16 : max_edges) : cap0);
    rc = launch_rd_fast_class<H, false>(eo_src, eo_dst, eo_ptr, node_ptr, n_graphs, rdh, counters, nf0, ef0, 0,
                                            two ? kFastSmall : 0x7fffffff, sms, st);
    if (rc != 0) return rc;
    if (two) {
        rc = (int)cudaMemsetAsync(counters + ESCGNN_CTR_TICKET_RD, 0, sizeof(unsigned long long), st);
        if (rc != 0) return rc;
        const int nf1 = (int)(max_nodes < rdfast::kMaxNodes ? max_nodes : rdfast::kMaxNodes);
        const int ef1 = (int)(max_edges < 1536 ? max_edges : 1536);
        rc = launch_rd_fast_class<H, true>(eo_src, eo_dst, eo_ptr, node_ptr, n_graphs, rdh, counters, nf1, ef1, kFastSmall + 1, 0x7fffffff,
                                           sms, st);
        if (rc != 0) return rc;
    }
    return (int)cudaMemsetAsync(counters + ESCGNN_CTR_TICKET_RD, 0, sizeof(unsigned long long), st);     // for the general solver
}

inline int& rd_peel_enabled() {
    static int on = 1;
    return on;
}

struct RdPlan {
    int64_t mat_doubles, smem, graph_bytes, slab, slab_graph;
    int sub_stride, aux_stride;
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
    p.aux_stride = max_nodes <= 64 ? (p.sub_stride + 15) & ~15 : 64;                    // pendant-tree peeling state (graphs of <= 64 nodes)
    const int64_t aux_bytes = align16((int64_t)kNw * p.aux_stride * 6);
    int64_t avail = budget - sub_bytes - aux_bytes - p.graph_bytes - 16;
    if (t * 8 * kNw <= 96 * 1024 && t * 8 * kNw <= avail) p.mat_doubles = t * kNw;      // warp mode for every graph
    else if (t * 8 <= avail) {                                                          // CTA mode on chip for the largest graphs ...
        p.mat_doubles = ((t + kNw - 1) / kNw) * kNw;
        const int64_t t64 = tri(65) * kNw;                                              // ... and a warp per pair up to 64 nodes, if it fits
        if (max_nodes > 64 && t64 > p.mat_doubles && t64 * 8 <= avail) p.mat_doubles = t64;
    }
    else p.mat_doubles = (64 * 1024 / 8 / kNw) * kNw;                                    // small graphs stay on chip
    p.smem = align16(p.mat_doubles * 8 + sub_bytes) + aux_bytes + p.graph_bytes;
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
    // the cycle-space kernel first (rd_fast.cuh): it solves the sparse systems and marks the rest for the general solver below
    const int fast_done = rd_fast_enabled() && max_edges <= 65535 ? 1 : 0;
    if (fast_done) {
        const int rc = launch_rd_fast<H>(eo_src, eo_dst, eo_ptr, node_ptr, n_graphs, rdh, counters, max_nodes, max_edges, st);
        if (rc != 0) return rc;
    }
    int dev = 0, sms = 148, smem_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // size classes: a batch whose largest graph needs the whole-CTA path (more than kSmall nodes) is encoded by two launches, so the
    // small graphs keep the small plan (three CTAs per SM, a warp per pair, peeling) instead of inheriting the large one
    constexpr int kSmall = 41;
    const int n_class = max_nodes > kSmall ? 2 : 1;
    for (int cls = 0; cls < n_class; ++cls) {
        const int n_lo = cls == 0 ? 0 : kSmall + 1, n_hi = (n_class == 2 && cls == 0) ? kSmall : (int)max_nodes;
        const RdPlan p = plan_rd(n_hi, max_edges, smem_optin);
        auto kern = ego_rd_kernel<H>;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (err != cudaSuccess) return (int)err;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, (size_t)p.smem);
        if (occ < 1) occ = 1;
        int64_t grid = (int64_t)sms * occ;
        if (encoder_grid_cap() > 0 && grid > encoder_grid_cap()) grid = encoder_grid_cap();
        int parts = (int)(grid / n_graphs);                      // CTAs left over by a small batch share its graphs
        parts = parts < 1 ? 1 : parts > 8 ? 8 : parts;
        if (grid > n_graphs * parts) grid = n_graphs * parts;
        if (p.slab) {
            if (scratch == nullptr || scratch_bytes < p.slab) return ESCGNN_ERR_BAD_ARG;
            const int64_t fit = scratch_bytes / p.slab;
            if (grid > fit) grid = fit;
        }
        if (grid < 1) grid = 1;
        if (cls > 0) {                                            // the work ticket starts again for the second class
            err = cudaMemsetAsync(counters + ESCGNN_CTR_TICKET_RD, 0, sizeof(unsigned long long), st);
            if (err != cudaSuccess) return (int)err;
        }
        kern<<<(unsigned)grid, kThreads, (size_t)p.smem, st>>>(eo_src, eo_dst, eo_ptr, node_ptr, (int)n_graphs, rdh, counters,
                                                                (long long)p.graph_bytes, (long long)p.mat_doubles, p.sub_stride,
                                                                (unsigned char*)scratch, (long long)p.slab, (long long)p.slab_graph,
                                                                parts, rd_peel_enabled() ? p.aux_stride : 0, n_lo, n_hi, fast_done);
        if ((err = cudaGetLastError()) != cudaSuccess) return (int)err;
    }
    return 0;
}

}  // namespace escgnn

using namespace escgnn;

extern "C" {

int escgnn_set_rd_fast(int on) {
    const int was = rd_fast_enabled();
    rd_fast_enabled() = on ? 1 : 0;
    return was;
}

int escgnn_set_rd_peel(int on) {
    const int was = rd_peel_enabled();
    rd_peel_enabled() = on ? 1 : 0;
    return was;
}

int64_t escgnn_encode_rd_scratch_bytes(int64_t max_nodes, int64_t max_edges, int h) {
    (void)h;
    const RdPlan p = plan_rd(max_nodes, max_edges, 227 * 1024);
    const RdPlan q = plan_rd(max_nodes > 41 ? 41 : max_nodes, max_edges, 227 * 1024);
    return (p.slab > q.slab ? p.slab : q.slab) * 148 * 3;
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
