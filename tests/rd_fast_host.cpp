// Host mirror of the ego_rd fast path (esc_gnn_b200/csrc/rd_fast.cuh): the SAME per-lane routines the sm_100a kernel runs
// (bfs_root, solve_pair), driven by a serial loader, so that the cycle-space mathematics can be checked against the oracle on the
// CPU (tests/test_rd_fast_cpu.py).  Test infrastructure only: the product never loads this.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../esc_gnn_b200/csrc/rd_fast.cuh"

using namespace escgnn::rdfast;

template <int H, int CMAX>
static int run(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, uint16_t* rdh, int* max_chords) {
    if (n < 1 || n > kMaxNodes || e < 1 || e > 60000) return 1;
    std::vector<unsigned char> buf(graph_bytes(n, e) + lane_bytes(n, CMAX) + 16);
    Ws ws;
    ws.n = (int)n; ws.e = (int)e; ws.rws = row_stride((int)n);
    unsigned char* base = buf.data();
    base += (8 - (reinterpret_cast<uintptr_t>(base) & 7)) & 7;
    carve_graph(ws, base, n, e);
    carve_lanes(ws, base + graph_bytes(n, e), n);
    std::vector<std::vector<std::pair<int, int>>> adj(n);
    for (int64_t i = 0; i < e; ++i) {
        if (src[i] < 0 || src[i] >= n || dst[i] < 0 || dst[i] >= n) return 2;
        adj[src[i]].push_back({(int)dst[i], (int)i});
    }
    int k = 0;
    for (int w = 0; w < n; ++w) {
        std::sort(adj[w].begin(), adj[w].end());
        ws.optr[w] = (uint16_t)k;
        for (size_t j = 0; j < adj[w].size(); ++j) {
            if (j && adj[w][j].first == adj[w][j - 1].first) return 3;       // multi-edge: not eligible
            ws.oadj[k] = (uint8_t)adj[w][j].first; ws.oeid[k] = (uint16_t)adj[w][j].second; ++k;
        }
    }
    ws.optr[n] = (uint16_t)k;
    for (int w = 0; w < n; ++w)
        for (auto& pr : adj[w]) {
            bool back = false;
            for (auto& q : adj[pr.first]) back |= q.first == w;
            if (!back) return 4;                                             // asymmetric: not eligible
        }
    for (int64_t i = 0; i < n * ws.rws; ++i) ws.dist[i] = 0xffffffffu;
    for (int r = 0; r < n; ++r) bfs_root<H>(ws, r);
    int lane = 0;
    *max_chords = 0;
    for (int64_t i = 0; i < e; ++i) rdh[i * kSlots] = kSentinel;
    for (int64_t i = 0; i < e; ++i) {
        const int u = (int)src[i], v = (int)dst[i];
        if (u > v) continue;
        Hist hu, hv;
        const int rc = solve_pair<H, CMAX>(ws, lane, u, v, hu, hv);
        lane = (lane + 1) & 31;
        if (rc == -2) return 5;
        if (rc < 0) continue;
        *max_chords = std::max(*max_chords, rc);
        for (int s = 0; s < kSlots; ++s) rdh[i * kSlots + s] = hu.get(s);
        if (u != v) {
            int rev = -1;
            for (int q = ws.optr[v]; q < ws.optr[v + 1]; ++q) if (ws.oadj[q] == u) rev = ws.oeid[q];
            if (rev < 0) return 6;
            for (int s = 0; s < kSlots; ++s) rdh[(int64_t)rev * kSlots + s] = hv.get(s);
        }
    }
    return 0;
}

extern "C" int rdfast_host_graph(const int64_t* src, const int64_t* dst, int64_t e, int64_t n, int h, int cmax, uint16_t* rdh,
                                 int* max_chords) {
#define GO(H, C) if (h == H && cmax == C) return run<H, C>(src, dst, e, n, rdh, max_chords);
    GO(1, 4) GO(2, 4) GO(3, 4) GO(4, 4) GO(1, 6) GO(2, 6) GO(3, 6) GO(4, 6) GO(3, 2) GO(4, 2)
#undef GO
    return 9;
}
