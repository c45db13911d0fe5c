// K1b fast path: resistance distances of sparse (molecule-like) pair systems from the CYCLE SPACE, one thread per system.
//
// Same contract as rd.cu (E5, /root/reference/utils_edge_efficient.py:92-107,130-131 under parity policy E5: float64, bin =
// trunc((float)rd)), different mathematics.  For a connected graph (S, F) with a spanning tree T rooted at u and c = |F| - |S| + 1
// chords, route a unit current from w to u along the tree (p_w = indicator of w's root path) and project out the cycle space:
//     Z_wx = p_w^T (I - C G^-1 C^T) p_x = lambda(w, x) - q(w)^T G^-1 q(x),        R(u, w) = Z_ww,
//     R(v, w) = d_T(v, w) - (q(w) - q(v))^T G^-1 (q(w) - q(v)),
// where lambda(x, y) = depth of the lowest common ancestor, C the tree-edge x cycle incidence of the fundamental cycles,
// q_i(w) = lambda(w, b_i) - lambda(w, a_i) for chord i = (a_i, b_i), and G = C^T C + I the c x c Gram matrix
// (G_ij = lambda(b_i,b_j) - lambda(b_i,a_j) - lambda(a_i,b_j) + lambda(a_i,a_j) + [i == j]).  Everything but the Cholesky factor of
// G is a small integer: a tree costs no linear algebra at all, a molecule with c <= 4 rings inside the ego-net a 4 x 4 factor --
// against an LDL^T + Takahashi sweep over a |S| x |S| matrix per pair in rd.cu (7 100 warp instructions per system, ncu r02).
// Phantom root (u == v): pinv(L)_ww = Z_ww - 2 r_w / m + s / m^2 with r_w = sum_x Z_wx = A(w) - y(w).Y,
// A(w) = sum over w's non-root ancestors-or-self of their subtree sizes, Y = sum_x y(x), y = L^-1 q.
//
// Work mapping: a WARP owns a graph, a LANE owns an unordered pair {u, v} (or a self-loop edge).  The loops over nodes and adjacency
// entries have the same trip counts in every lane (they depend on the graph only), so the 32 systems advance in lockstep and
// differ only in predicates; per-lane state (parent / depth / subtree size per node) sits in shared memory as [node][lane] bytes.
// The spanning tree needs no search: the hop-distance rows of u and v (E2) already order S -- the parent of w is its smallest
// neighbour one hop closer to u (inside B_u) or to v (outside), so every root path is at most h + 1 long and fits one 64-bit word.
//
// The per-lane routines are plain C++ (host + device): tests/test_rd_fast_cpu.py compiles them for the CPU and checks the histograms.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define RDF_HD __host__ __device__ __forceinline__
#else
#define RDF_HD inline
#endif

namespace escgnn {
namespace rdfast {

constexpr int kLanes = 32;
constexpr int kMaxNodes = 128;              // node ids are 7 bits inside the packed root paths
constexpr int kSlots = 12;                  // = ESCGNN_RD_SLOTS
constexpr uint32_t kFarD = 15u;             // = kFar of graph_smem.cuh
constexpr uint8_t kNone = 0xff;             // par[]: node not in S
constexpr uint16_t kSentinel = 0xffff;      // rdh slot 0 of an edge this path did not solve (slot 0 of a solved edge is >= 1)

RDF_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// lambda(x, y) from two packed root paths: the ancestors agree from the root down to the lowest common ancestor, so the number of
// equal bytes is the index of the first differing byte (the filler marks of any two paths differ, so there always is one)
RDF_HD int equal_bytes(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return (__ffsll((long long)(a ^ b)) - 1) >> 3;
#else
    return __builtin_ctzll(a ^ b) >> 3;
#endif
}

// one warp's workspace
struct Ws {
    int n, e, rws;        // nodes, directed edges, words per distance row (odd: the 32 rows a warp touches fall into distinct banks)
    uint16_t* optr;       // [n + 1]   CSR by source
    uint8_t* oadj;        // [e]       targets, ascending inside a node
    uint16_t* oeid;       // [e]       position of the entry in the graph's edge list
    uint16_t* list;       // [e]       edges with src <= dst: one per unordered pair / self-loop
    uint32_t* dist;       // [n][rws]  hop distances, 4 bits each, 15 = farther than h
    uint8_t* par;         // [n][32]   per lane: parent in the spanning tree (kNone: not in S; the root is its own parent)
    uint8_t* dep;         // [n][32]   per lane: depth
    uint8_t* siz;         // [n][32]   per lane: subtree size (phantom systems)
    uint64_t* path;       // [2 CMAX][32] per lane: packed root paths of the chord endpoints (byte k-1 = the ancestor at depth k)
};

RDF_HD int row_stride(int n) { return ((n + 7) >> 3) | 1; }

RDF_HD uint32_t nibr(const uint32_t* row, int w) { return (row[w >> 3] >> ((w & 7) << 2)) & 15u; }

RDF_HD int64_t al8(int64_t x) { return (x + 7) & ~int64_t(7); }
RDF_HD int64_t dist_words(int64_t nf) {            // the counting-sort cursors (2 nf + 2 words) alias the distance matrix
    const int64_t d = nf * row_stride((int)nf);
    return d < 2 * nf + 2 ? 2 * nf + 2 : d;
}
// bytes of the per-graph part (CSR + distance matrix) and of one warp's per-lane part, for graphs of <= nf nodes / ef directed edges
RDF_HD int64_t graph_bytes(int64_t nf, int64_t ef) { return al8(4 * dist_words(nf)) + al8(2 * (nf + 1)) + 2 * al8(2 * ef) + al8(ef); }
RDF_HD int64_t lane_bytes(int64_t nf, int cmax) { return 3 * al8(nf * kLanes) + 8 * 2 * (int64_t)cmax * kLanes; }

RDF_HD void carve_graph(Ws& ws, unsigned char* base, int64_t nf, int64_t ef) {        // base 8-byte aligned
    ws.dist = reinterpret_cast<uint32_t*>(base); base += al8(4 * dist_words(nf));
    ws.optr = reinterpret_cast<uint16_t*>(base); base += al8(2 * (nf + 1));
    ws.oeid = reinterpret_cast<uint16_t*>(base); base += al8(2 * ef);
    ws.list = reinterpret_cast<uint16_t*>(base); base += al8(2 * ef);
    ws.oadj = base;
}
RDF_HD void carve_lanes(Ws& ws, unsigned char* base, int64_t nf) {
    ws.par = base; base += al8(nf * kLanes);
    ws.dep = base; base += al8(nf * kLanes);
    ws.siz = base; base += al8(nf * kLanes);
    ws.path = reinterpret_cast<uint64_t*>(base);
}

// E2 for one root (lane-private row of the distance matrix; uniform loops)
template <int H>
RDF_HD void bfs_root(const Ws& ws, int r) {
    uint32_t* row = ws.dist + (size_t)r * ws.rws;
    row[r >> 3] &= ~(15u << ((r & 7) << 2));
    for (int level = 0; level < H; ++level) {
        bool grew = false;
        for (int w = 0; w < ws.n; ++w) {
            if (nibr(row, w) != (uint32_t)level) continue;
            const int ka = ws.optr[w], kb = ws.optr[w + 1];
            for (int k = ka; k < kb; ++k) {
                const int b = ws.oadj[k];
                const int sh = (b & 7) << 2;
                if (((row[b >> 3] >> sh) & 15u) == kFarD) {
                    row[b >> 3] &= ~((15u ^ (uint32_t)(level + 1)) << sh);
                    grew = true;
                }
            }
        }
        if (!grew) break;
    }
}

// 12 histogram slots of 8 bits (counts <= |S| + 1 <= 129)
struct Hist {
    uint64_t lo;      // slots 0..7
    uint32_t hi;      // slots 8..11
    RDF_HD void clear() { lo = 0; hi = 0; }
    RDF_HD bool add(double rd) {                       // bin = trunc((float)rd)  (torch.FloatTensor(...).long(), :105,131)
        const int b = (int)truncf((float)rd);
        if (b < 0 || b >= kSlots) return false;
        if (b < 8) lo += 1ull << (8 * b); else hi += 1u << (8 * (b - 8));
        return true;
    }
    RDF_HD uint16_t get(int s) const { return (uint16_t)(s < 8 ? (lo >> (8 * s)) & 0xff : (hi >> (8 * (s - 8))) & 0xff); }
};

// Solve the system of lane `lane`: pair {u, v} (u < v, adjacent) or the self-loop edge (u == v).
// Returns the number of chords (>= 0), -1 when the system has more than CMAX independent cycles (not solved), -2 on an rd bin
// outside [0, kSlots) (ESCGNN_DATA_RD).  hu = histogram of the directed edge (u, v), hv = that of (v, u) (unused when u == v).
template <int H, int CMAX>
RDF_HD int solve_pair(const Ws& ws, int lane, int u, int v, Hist& hu, Hist& hv) {
    const int n = ws.n;
    const uint32_t* rowU = ws.dist + (size_t)u * ws.rws;
    const uint32_t* rowV = ws.dist + (size_t)v * ws.rws;
    uint8_t* par = ws.par + lane;
    uint8_t* dep = ws.dep + lane;
    uint8_t* siz = ws.siz + lane;
    uint64_t* path = ws.path + lane;
    const bool phantom = u == v;
    constexpr int L = kLanes;
    hu.clear(); hv.clear();

    // ---- spanning tree of (S, F): parent = smallest neighbour one hop closer to u (inside B_u) / to v (outside B_u)
    int members = 0;
    for (int w = 0; w < n; ++w) {
        const uint32_t du = nibr(rowU, w), dv = nibr(rowV, w);
        if (du == kFarD && dv == kFarD) { par[w * L] = kNone; continue; }
        ++members;
        if (w == u) { par[w * L] = (uint8_t)u; dep[w * L] = 0; continue; }
        const bool in_u = du != kFarD;
        const uint32_t want = (in_u ? du : dv) - 1u;
        const uint32_t* row = in_u ? rowU : rowV;
        int p = kNone;
        const int ka = ws.optr[w], kb = ws.optr[w + 1];
        for (int k = ka; k < kb; ++k) {
            const int b = ws.oadj[k];
            if (p == kNone && b != w && nibr(row, b) == want) p = b;
        }
        par[w * L] = (uint8_t)p;
        dep[w * L] = in_u ? (uint8_t)du : kNone;
    }
    if (!phantom) {
        for (int w = 0; w < n; ++w) {                      // depth outside B_u: walk towards v until B_u is entered
            if (par[w * L] == kNone || dep[w * L] != kNone) continue;
            int x = par[w * L], s = 1;
            while (dep[x * L] == kNone) { x = par[x * L]; ++s; }
            dep[w * L] = (uint8_t)(s + dep[x * L]);
        }
    }
    // packed root path of x: byte k-1 = ancestor-or-self at depth k, other bytes = mark; also A = sum of subtree sizes on the path
    auto pack = [&](int x, uint8_t mark, int* sum_siz) -> uint64_t {
        uint64_t P = 0x0101010101010101ull * mark;
        int a = 0;
        for (int k = dep[x * L]; k >= 1; --k) {               // x itself enters at byte 0 and is shifted up to byte depth - 1
            P = (P << 8) | (uint64_t)x;
            if (sum_siz) a += siz[x * L];
            x = par[x * L];
        }
        if (sum_siz) *sum_siz = a;
        return P;
    };
    // ---- chords: the edges of F outside the tree, each undirected edge seen from its smaller endpoint
    int c = 0;
    bool over = false;
    for (int w = 0; w < n; ++w) {
        if (par[w * L] == kNone) continue;
        const uint32_t du = nibr(rowU, w), dv = nibr(rowV, w);
        const int ka = ws.optr[w], kb = ws.optr[w + 1];
        for (int k = ka; k < kb; ++k) {
            const int b = ws.oadj[k];
            if (b <= w || par[b * L] == kNone) continue;                 // (also drops loops: scipy's laplacian ignores them)
            const uint32_t bu = nibr(rowU, b), bv = nibr(rowV, b);
            if (!((du != kFarD && bu != kFarD) || (dv != kFarD && bv != kFarD))) continue;      // not in F (:55, :283-285)
            if (par[w * L] == b || par[b * L] == w) continue;            // tree edge
            if (c == CMAX) { over = true; continue; }
            path[(2 * c) * L] = pack(w, (uint8_t)(0x80 | (2 * c)), nullptr);
            path[(2 * c + 1) * L] = pack(b, (uint8_t)(0x80 | (2 * c + 1)), nullptr);
            ++c;
        }
    }
    // the unrolled per-chord code below runs for the largest chord count among the lanes of the warp only (uniform branches): a chunk
    // of tree-like ego-nets skips the Gram matrix and the triangular solves altogether
#ifdef __CUDA_ARCH__
    const int cw = __reduce_max_sync(__activemask(), c);
#else
    const int cw = c;
#endif
    if (over) return -1;
    // ---- Gram matrix of the fundamental cycles and its Cholesky factor (identity beyond c: the unrolled code is uniform)
    uint64_t pa[CMAX], pb[CMAX];
    #pragma unroll
    for (int i = 0; i < CMAX; ++i) {
        pa[i] = i < c ? path[(2 * i) * L] : 0x0101010101010101ull * (uint8_t)(0xc0 | (2 * i));         // marks never equal anything
        pb[i] = i < c ? path[(2 * i + 1) * L] : 0x0101010101010101ull * (uint8_t)(0xc0 | (2 * i + 1));
    }
    double g[CMAX][CMAX];
    #pragma unroll
    for (int i = 0; i < CMAX; ++i) {
        #pragma unroll
        for (int j = 0; j <= i; ++j) {
            int val;
            if (i >= cw) {
                val = i == j ? 1 : 0;
            } else if (i == j) {
                const int da = 8 - popc64(pa[i] & 0x8080808080808080ull), db = 8 - popc64(pb[i] & 0x8080808080808080ull);
                val = i < c ? da + db - 2 * equal_bytes(pa[i], pb[i]) + 1 : 1;                          // the cycle's length
            } else {
                val = equal_bytes(pb[i], pb[j]) - equal_bytes(pb[i], pa[j]) - equal_bytes(pa[i], pb[j]) + equal_bytes(pa[i], pa[j]);
            }
            g[i][j] = (double)val;
        }
    }
    double inv_d[CMAX];
    bool bad = false;
    #pragma unroll
    for (int j = 0; j < CMAX; ++j) {
        inv_d[j] = 1.0;
        if (j < cw) {
            double d = g[j][j];
            #pragma unroll
            for (int k = 0; k < j; ++k) d -= g[j][k] * g[j][k];
            if (!(d > 1e-12)) { bad = true; d = 1.0; }
            const double l = sqrt(d);
            inv_d[j] = 1.0 / l;
            g[j][j] = l;
            #pragma unroll
            for (int i = j + 1; i < CMAX; ++i) {
                double s = g[i][j];
                #pragma unroll
                for (int k = 0; k < j; ++k) s -= g[i][k] * g[j][k];
                g[i][j] = s * inv_d[j];
            }
        }
    }
    if (bad) return -2;
    // y(x) = L^-1 q(x), q_i(x) = lambda(x, b_i) - lambda(x, a_i)
    auto solve_y = [&](uint64_t P, double (&y)[CMAX]) {
        #pragma unroll
        for (int i = 0; i < CMAX; ++i) {
            y[i] = 0.0;
            if (i < cw) {
                double s = (double)(equal_bytes(P, pb[i]) - equal_bytes(P, pa[i]));
                #pragma unroll
                for (int k = 0; k < i; ++k) s -= g[i][k] * y[k];
                y[i] = s * inv_d[i];
            }
        }
    };
    bool ok = true;
    if (!phantom) {
        double yv[CMAX];
        solve_y(pack(v, 0xfe, nullptr), yv);
        for (int w = 0; w < n; ++w) {
            if (par[w * L] == kNone) continue;
            const uint64_t P = pack(w, 0xfe, nullptr);
            const int d = dep[w * L];
            double y[CMAX];
            solve_y(P, y);
            double n2 = 0.0, m2 = 0.0;
            #pragma unroll
            for (int i = 0; i < CMAX; ++i) { n2 += y[i] * y[i]; const double t = y[i] - yv[i]; m2 += t * t; }
            const int on_v = (d >= 1 && (int)(P & 0xff) == v) ? 1 : 0;           // v (depth 1) is an ancestor-or-self of w
            ok &= hu.add((double)d - n2);                                        // R(u, w)
            ok &= hv.add((double)(d + 1 - 2 * on_v) - m2);                       // R(v, w)
        }
    } else {
        for (int w = 0; w < n; ++w) if (par[w * L] != kNone) siz[w * L] = 1;
        for (int lev = H; lev >= 1; --lev)
            for (int w = 0; w < n; ++w)
                if (par[w * L] != kNone && dep[w * L] == lev) siz[par[w * L] * L] = (uint8_t)(siz[par[w * L] * L] + siz[w * L]);
        double Y[CMAX];
        #pragma unroll
        for (int i = 0; i < CMAX; ++i) Y[i] = 0.0;
        long long a_sum = 0;
        for (int w = 0; w < n; ++w) {
            if (par[w * L] == kNone) continue;
            int a;
            const uint64_t P = pack(w, 0xfe, &a);
            double y[CMAX];
            solve_y(P, y);
            #pragma unroll
            for (int i = 0; i < CMAX; ++i) Y[i] += y[i];
            a_sum += a;
        }
        double yy = 0.0;
        #pragma unroll
        for (int i = 0; i < CMAX; ++i) yy += Y[i] * Y[i];
        const double mm = (double)members, S = (double)a_sum - yy;
        for (int w = 0; w < n; ++w) {
            if (par[w * L] == kNone) continue;
            int a;
            const uint64_t P = pack(w, 0xfe, &a);
            double y[CMAX];
            solve_y(P, y);
            double n2 = 0.0, yd = 0.0;
            #pragma unroll
            for (int i = 0; i < CMAX; ++i) { n2 += y[i] * y[i]; yd += y[i] * Y[i]; }
            const double zww = (double)dep[w * L] - n2, rw = (double)a - yd;
            ok &= hu.add(zww - 2.0 * rw / mm + S / (mm * mm));                   // pinv(L_ball)_ww
        }
        ok &= hu.add(0.0);                                                       // the phantom root itself
    }
    return ok ? c : -2;
}

}  // namespace rdfast
}  // namespace escgnn
