// Host-buffer front end of the encoder: the C-ABI calls a Python `create_subgraphs` / dataset `process()` replacement
// makes when it holds CPU tensors (reference call sites: run_zinc.py:141-146, run_graphcount.py:404-408,
// run_ogb_mol.py:329-332; the per-graph loop of GraphCountDataset.py:111-117).
//
// Two ways in:
//  * escgnn_encode_host_run / _fetch: one synchronous call pair that returns the reference's int64 triple (expanded on
//    the device, 24 bytes per record over PCIe) -- the literal contract, for single graphs and small batches.
//  * escgnn_encode_host_submit / _wait (+ escgnn_expand_records_host): the throughput path.  A context owns TWO slots,
//    each with its own stream, device workspace and PINNED staging / result arenas that only ever grow, so after warm-up
//    no call allocates.  submit() stages the inputs and queues H2D + kernels and returns; wait() fetches the COMPACT
//    result (4 bytes per record: index | count << 11, plus 12 bytes per edge) into the slot's pinned arena.  With the
//    next chunk submitted on the other slot before waiting, the D2H of chunk k runs under the kernels of chunk k+1.
//    The int64 triple is produced on the host, by all cores, only when somebody asks for it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/escgnn_b200.h"

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return (int)e;
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {            // grow-only page-locked host memory
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) return (int)e;
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct Slot {
    cudaStream_t stream = nullptr;
    Buf src, dst, edge_ptr, node_ptr, eo_src, eo_dst, eo_ptr, tmp, rdh, rec, rec_off, rec_nnz, edge_graph, out_off,
        scan_tmp, counters, scratch, pos_enc, pos_index, pos_batch;
    PinnedBuf h_in, h_rec, h_edges;             // staging of the inputs; records; per-edge arrays + rewritten edge list
    unsigned long long* h_counters = nullptr;   // pinned
    int64_t* h_ptr_tail = nullptr;              // pinned: eo_ptr[G]
    // the run in flight / last run
    int64_t n_graphs = 0, e_in = 0, e_cap = 0, e_out = 0, nnz = 0, max_n = 0, max_e = 0;
    int h = 0, use_rd = 0, self_loop = 0;
    bool alias_input = false, busy = false;
    void release() {
        Buf* all[] = {&src, &dst, &edge_ptr, &node_ptr, &eo_src, &eo_dst, &eo_ptr, &tmp, &rdh, &rec, &rec_off, &rec_nnz,
                      &edge_graph, &out_off, &scan_tmp, &counters, &scratch, &pos_enc, &pos_index, &pos_batch};
        for (Buf* b : all) b->release();
        h_in.release(); h_rec.release(); h_edges.release();
        if (h_counters) cudaFreeHost(h_counters);
        if (h_ptr_tail) cudaFreeHost(h_ptr_tail);
        if (stream) cudaStreamDestroy(stream);
    }
};

constexpr int kSlots = 2;

// the context's device is made current for the duration of a call and the caller's device restored afterwards (the host thread
// usually belongs to a framework that tracks its own current device)
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct escgnn_ctx {
    int device = 0;
    Slot slot[kSlots];
};

#define ESC_TRY(x) do { int _rc = (int)(x); if (_rc != 0) return _rc; } while (0)

namespace {

// sizes, capacity checks, H2D of the inputs (from `h_*`, which must stay valid until the stream has consumed them:
// caller memory for the synchronous path, the slot's pinned staging for the pipelined one) and the kernels E1 / E5 / E2-E4
int launch_encode(escgnn_ctx* c, Slot& s, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                  const int64_t* h_node_ptr, int64_t G, int h, int use_rd, int self_loop) {
    cudaStream_t st = s.stream;
    const int64_t e_in = s.e_in, e_cap = s.e_cap;
    ESC_TRY(s.src.ensure((size_t)(e_in + 1) * 8));
    ESC_TRY(s.dst.ensure((size_t)(e_in + 1) * 8));
    ESC_TRY(s.edge_ptr.ensure((size_t)(G + 1) * 8));
    ESC_TRY(s.node_ptr.ensure((size_t)(G + 1) * 8));
    ESC_TRY(s.counters.ensure(ESCGNN_NUM_COUNTERS * 8));
    ESC_TRY(s.rec_off.ensure((size_t)(e_cap + 1) * 8));
    ESC_TRY(s.rec_nnz.ensure((size_t)(e_cap + 1) * 4));
    ESC_TRY(s.edge_graph.ensure((size_t)(e_cap + 1) * 4));
    ESC_TRY(s.out_off.ensure((size_t)(e_cap + 2) * 8));
    ESC_TRY(s.scan_tmp.ensure((size_t)(e_cap / 1024 + 4) * 8));
    ESC_TRY(cudaMemcpyAsync(s.src.p, h_src, (size_t)e_in * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(s.dst.p, h_dst, (size_t)e_in * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(s.edge_ptr.p, h_edge_ptr, (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(s.node_ptr.p, h_node_ptr, (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemsetAsync(s.counters.p, 0, ESCGNN_NUM_COUNTERS * 8, st));
    s.alias_input = !self_loop;
    if (self_loop) {
        ESC_TRY(s.eo_src.ensure((size_t)(e_cap + 1) * 8));
        ESC_TRY(s.eo_dst.ensure((size_t)(e_cap + 1) * 8));
        ESC_TRY(s.eo_ptr.ensure((size_t)(G + 1) * 8));
        ESC_TRY(s.tmp.ensure((size_t)(4 * G + 8 * (G / 1024 + 2) + 64)));
        ESC_TRY(escgnn_rewrite_self_loops(s.src.as<int64_t>(), s.dst.as<int64_t>(), s.edge_ptr.as<int64_t>(),
                                          s.node_ptr.as<int64_t>(), G, s.eo_ptr.as<int64_t>(),
                                          s.eo_src.as<int64_t>(), s.eo_dst.as<int64_t>(), s.tmp.p, st));
    }
    int64_t sb = escgnn_encode_scratch_bytes(s.max_n, s.max_e, h);
    if (use_rd) {
        const int64_t sb_rd = escgnn_encode_rd_scratch_bytes(s.max_n, s.max_e, h);
        if (sb_rd > sb) sb = sb_rd;
    }
    if (sb > 0) ESC_TRY(s.scratch.ensure((size_t)sb));
    const int64_t want = e_cap * 48 + 1024;
    if ((int64_t)(s.rec.cap / 4) < want) ESC_TRY(s.rec.ensure((size_t)want * 4));
    if (use_rd) ESC_TRY(s.rdh.ensure((size_t)(e_cap + 1) * ESCGNN_RD_SLOTS * 2));
    return 0;
}

const int64_t* eo_src_of(Slot& s) { return s.alias_input ? s.src.as<int64_t>() : s.eo_src.as<int64_t>(); }
const int64_t* eo_dst_of(Slot& s) { return s.alias_input ? s.dst.as<int64_t>() : s.eo_dst.as<int64_t>(); }
const int64_t* eo_ptr_of(Slot& s) { return s.alias_input ? s.edge_ptr.as<int64_t>() : s.eo_ptr.as<int64_t>(); }

int run_kernels(Slot& s) {
    cudaStream_t st = s.stream;
    const uint16_t* rdh = nullptr;
    if (s.use_rd) {
        ESC_TRY(escgnn_encode_rd(eo_src_of(s), eo_dst_of(s), eo_ptr_of(s), s.node_ptr.as<int64_t>(), s.n_graphs, s.h,
                                 s.rdh.as<uint16_t>(), s.counters.as<unsigned long long>(), s.max_n, s.max_e, s.scratch.p,
                                 (int64_t)s.scratch.cap, st));
        rdh = s.rdh.as<uint16_t>();
    }
    ESC_TRY(escgnn_encode(eo_src_of(s), eo_dst_of(s), eo_ptr_of(s), s.node_ptr.as<int64_t>(), s.n_graphs, s.h, rdh,
                          s.rec.as<uint32_t>(), (int64_t)(s.rec.cap / 4), s.rec_off.as<int64_t>(), s.rec_nnz.as<int32_t>(),
                          s.edge_graph.as<int32_t>(), s.counters.as<unsigned long long>(), s.max_n, s.max_e, s.scratch.p,
                          (int64_t)s.scratch.cap, st));
    ESC_TRY(cudaMemcpyAsync(s.h_counters, s.counters.p, ESCGNN_NUM_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
    ESC_TRY(cudaMemcpyAsync(s.h_ptr_tail, eo_ptr_of(s) + s.n_graphs, 8, cudaMemcpyDeviceToHost, st));
    return 0;
}

// wait for the kernels; if the record buffer was too small (data dependent), grow it and encode once more
int finish_kernels(Slot& s) {
    cudaStream_t st = s.stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        ESC_TRY(cudaStreamSynchronize(st));
        const int64_t nnz = (int64_t)s.h_counters[ESCGNN_CTR_NNZ];
        if (nnz <= (int64_t)(s.rec.cap / 4)) break;
        if (attempt == 1) return ESCGNN_ERR_CAPACITY;
        ESC_TRY(s.rec.ensure((size_t)nnz * 4));
        ESC_TRY(cudaMemsetAsync(s.counters.as<unsigned long long>() + ESCGNN_CTR_NNZ, 0, 8, st));
        ESC_TRY(cudaMemsetAsync(s.counters.as<unsigned long long>() + ESCGNN_CTR_TICKET, 0, 8, st));
        ESC_TRY(escgnn_encode(eo_src_of(s), eo_dst_of(s), eo_ptr_of(s), s.node_ptr.as<int64_t>(), s.n_graphs, s.h,
                              s.use_rd ? s.rdh.as<uint16_t>() : nullptr, s.rec.as<uint32_t>(), (int64_t)(s.rec.cap / 4),
                              s.rec_off.as<int64_t>(), s.rec_nnz.as<int32_t>(), s.edge_graph.as<int32_t>(),
                              s.counters.as<unsigned long long>(), s.max_n, s.max_e, s.scratch.p, (int64_t)s.scratch.cap, st));
        ESC_TRY(cudaMemcpyAsync(s.h_counters, s.counters.p, ESCGNN_NUM_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
    }
    s.e_out = *s.h_ptr_tail;
    s.nnz = (int64_t)s.h_counters[ESCGNN_CTR_NNZ];
    return 0;
}

int scan_sizes(Slot& s, const int64_t* h_edge_ptr, const int64_t* h_node_ptr, int64_t G, int h, int use_rd, int self_loop) {
    if (h_edge_ptr[0] != 0 || h_node_ptr[0] != 0) return ESCGNN_ERR_BAD_ARG;
    int64_t max_n = 0, max_e = 0;
    for (int64_t g = 0; g < G; ++g) {
        const int64_t n = h_node_ptr[g + 1] - h_node_ptr[g], e = h_edge_ptr[g + 1] - h_edge_ptr[g];
        if (n < 0 || e < 0) return ESCGNN_ERR_BAD_ARG;
        if (n > max_n) max_n = n;
        const int64_t eo = self_loop ? e + n : e;
        if (eo > max_e) max_e = eo;
    }
    s.n_graphs = G; s.h = h; s.use_rd = use_rd; s.self_loop = self_loop;
    s.e_in = h_edge_ptr[G]; s.max_n = max_n; s.max_e = max_e;
    s.e_cap = self_loop ? s.e_in + h_node_ptr[G] : s.e_in;
    s.e_out = 0; s.nnz = 0;
    return 0;
}

}  // namespace

extern "C" {

escgnn_ctx* escgnn_ctx_create(int device) {
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return nullptr;
    escgnn_ctx* c = new escgnn_ctx();
    c->device = device;
    for (int i = 0; i < kSlots; ++i) {
        Slot& s = c->slot[i];
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMallocHost((void**)&s.h_counters, ESCGNN_NUM_COUNTERS * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMallocHost((void**)&s.h_ptr_tail, sizeof(int64_t)) != cudaSuccess) {
            escgnn_ctx_destroy(c);
            return nullptr;
        }
    }
    return c;
}

void escgnn_ctx_destroy(escgnn_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (int i = 0; i < kSlots; ++i) {
        if (c->slot[i].stream) cudaStreamSynchronize(c->slot[i].stream);
        c->slot[i].release();
    }
    delete c;
}

int escgnn_encode_host_run(escgnn_ctx* c, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                           const int64_t* h_node_ptr, int64_t G, int h, int use_rd, int self_loop,
                           int local_ordinals, int64_t* out_num_edges, int64_t* out_nnz, uint32_t* out_error_bits) {
    if (!c || G < 0 || h < 1 || h > 4) return ESCGNN_ERR_BAD_ARG;
    DeviceGuard guard(c->device);
    ESC_TRY(guard.err);
    Slot& s = c->slot[0];
    if (s.busy) return ESCGNN_ERR_BAD_ARG;      // a pipelined submit is in flight on this slot
    *out_num_edges = 0; *out_nnz = 0; *out_error_bits = 0;
    s.n_graphs = G; s.e_out = 0; s.nnz = 0;
    if (G == 0) return 0;
    ESC_TRY(scan_sizes(s, h_edge_ptr, h_node_ptr, G, h, use_rd, self_loop));
    ESC_TRY(launch_encode(c, s, h_src, h_dst, h_edge_ptr, h_node_ptr, G, h, use_rd, self_loop));
    ESC_TRY(run_kernels(s));
    ESC_TRY(finish_kernels(s));
    *out_error_bits = (uint32_t)s.h_counters[ESCGNN_CTR_ERROR];
    *out_num_edges = s.e_out;
    *out_nnz = s.nnz;
    if (*out_error_bits) return ESCGNN_ERR_DATA;
    cudaStream_t st = s.stream;
    ESC_TRY(escgnn_exclusive_scan_i32(s.rec_nnz.as<int32_t>(), s.e_out, s.out_off.as<int64_t>(), s.scan_tmp.as<int64_t>(), st));
    ESC_TRY(s.pos_enc.ensure((size_t)(s.nnz + 1) * 8));
    ESC_TRY(s.pos_index.ensure((size_t)(s.nnz + 1) * 8));
    ESC_TRY(s.pos_batch.ensure((size_t)(s.nnz + 1) * 8));
    ESC_TRY(escgnn_expand_records(s.rec.as<uint32_t>(), s.rec_off.as<int64_t>(), s.rec_nnz.as<int32_t>(),
                                  s.edge_graph.as<int32_t>(), eo_ptr_of(s), s.out_off.as<int64_t>(), s.e_out, use_rd,
                                  local_ordinals, s.pos_enc.as<int64_t>(), s.pos_index.as<int64_t>(),
                                  s.pos_batch.as<int64_t>(), st));
    return 0;
}

int escgnn_encode_host_fetch(escgnn_ctx* c, int64_t* h_eo_src, int64_t* h_eo_dst, int64_t* h_eo_ptr,
                             int64_t* h_pos_enc, int64_t* h_pos_index, int64_t* h_pos_batch) {
    if (!c) return ESCGNN_ERR_BAD_ARG;
    DeviceGuard guard(c->device);
    ESC_TRY(guard.err);
    Slot& s = c->slot[0];
    cudaStream_t st = s.stream;
    if (s.n_graphs > 0) {
        if (h_eo_src) ESC_TRY(cudaMemcpyAsync(h_eo_src, eo_src_of(s), (size_t)s.e_out * 8, cudaMemcpyDeviceToHost, st));
        if (h_eo_dst) ESC_TRY(cudaMemcpyAsync(h_eo_dst, eo_dst_of(s), (size_t)s.e_out * 8, cudaMemcpyDeviceToHost, st));
        if (h_eo_ptr) ESC_TRY(cudaMemcpyAsync(h_eo_ptr, eo_ptr_of(s), (size_t)(s.n_graphs + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_enc) ESC_TRY(cudaMemcpyAsync(h_pos_enc, s.pos_enc.p, (size_t)s.nnz * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_index) ESC_TRY(cudaMemcpyAsync(h_pos_index, s.pos_index.p, (size_t)s.nnz * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_batch) ESC_TRY(cudaMemcpyAsync(h_pos_batch, s.pos_batch.p, (size_t)s.nnz * 8, cudaMemcpyDeviceToHost, st));
    }
    return (int)cudaStreamSynchronize(st);
}

int escgnn_encode_host_device_results(escgnn_ctx* c, const uint32_t** d_rec, const int64_t** d_rec_off,
                                      const int32_t** d_rec_nnz, const int64_t** d_eo_src, const int64_t** d_eo_dst) {
    if (!c) return ESCGNN_ERR_BAD_ARG;
    Slot& s = c->slot[0];
    if (d_rec) *d_rec = s.rec.as<uint32_t>();
    if (d_rec_off) *d_rec_off = s.rec_off.as<int64_t>();
    if (d_rec_nnz) *d_rec_nnz = s.rec_nnz.as<int32_t>();
    if (d_eo_src) *d_eo_src = eo_src_of(s);
    if (d_eo_dst) *d_eo_dst = eo_dst_of(s);
    return 0;
}

int escgnn_encode_host_submit(escgnn_ctx* c, int slot, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                              const int64_t* h_node_ptr, int64_t G, int h, int use_rd, int self_loop) {
    if (!c || slot < 0 || slot >= kSlots || G < 1 || h < 1 || h > 4) return ESCGNN_ERR_BAD_ARG;
    DeviceGuard guard(c->device);
    ESC_TRY(guard.err);
    Slot& s = c->slot[slot];
    if (s.busy) return ESCGNN_ERR_BAD_ARG;      // wait() first
    ESC_TRY(scan_sizes(s, h_edge_ptr, h_node_ptr, G, h, use_rd, self_loop));
    // stage the inputs in pinned memory: the caller's (pageable) buffers are free as soon as this call returns, and the H2D
    // copies are truly asynchronous
    const size_t b_e = (size_t)s.e_in * 8, b_p = (size_t)(G + 1) * 8;
    ESC_TRY(s.h_in.ensure(2 * b_e + 2 * b_p + 64));
    int64_t* st_src = s.h_in.as<int64_t>();
    int64_t* st_dst = st_src + s.e_in;
    int64_t* st_ep = st_dst + s.e_in;
    int64_t* st_np = st_ep + (G + 1);
    memcpy(st_src, h_src, b_e); memcpy(st_dst, h_dst, b_e); memcpy(st_ep, h_edge_ptr, b_p); memcpy(st_np, h_node_ptr, b_p);
    ESC_TRY(launch_encode(c, s, st_src, st_dst, st_ep, st_np, G, h, use_rd, self_loop));
    ESC_TRY(run_kernels(s));
    s.busy = true;
    return 0;
}

int escgnn_encode_host_wait(escgnn_ctx* c, int slot, int64_t* out_num_edges, int64_t* out_nnz, uint32_t* out_error_bits,
                            const uint32_t** h_rec, const int64_t** h_rec_off, const int32_t** h_rec_nnz,
                            const int64_t** h_eo_src, const int64_t** h_eo_dst, const int64_t** h_eo_ptr) {
    if (!c || slot < 0 || slot >= kSlots) return ESCGNN_ERR_BAD_ARG;
    DeviceGuard guard(c->device);
    ESC_TRY(guard.err);
    Slot& s = c->slot[slot];
    if (!s.busy) return ESCGNN_ERR_BAD_ARG;
    s.busy = false;
    ESC_TRY(finish_kernels(s));
    *out_error_bits = (uint32_t)s.h_counters[ESCGNN_CTR_ERROR];
    *out_num_edges = s.e_out;
    *out_nnz = s.nnz;
    if (*out_error_bits) return ESCGNN_ERR_DATA;
    cudaStream_t st = s.stream;
    const int64_t E = s.e_out, G = s.n_graphs;
    ESC_TRY(s.h_rec.ensure((size_t)(s.nnz + 1) * 4));
    // per-edge arrays: rec_off int64 [E], then (rewritten edge list only) eo_src, eo_dst int64 [E], eo_ptr int64 [G+1], rec_nnz int32 [E]
    const size_t n64 = (size_t)E + (s.alias_input ? 0 : 2 * (size_t)E + (size_t)(G + 1));
    ESC_TRY(s.h_edges.ensure(n64 * 8 + (size_t)(E + 1) * 4 + 64));
    int64_t* p_off = s.h_edges.as<int64_t>();
    int64_t* p_src = p_off + E;
    int64_t* p_dst = p_src + (s.alias_input ? 0 : E);
    int64_t* p_ptr = p_dst + (s.alias_input ? 0 : E);
    int32_t* p_nnz = reinterpret_cast<int32_t*>(s.h_edges.as<int64_t>() + n64);
    ESC_TRY(cudaMemcpyAsync(s.h_rec.p, s.rec.p, (size_t)s.nnz * 4, cudaMemcpyDeviceToHost, st));
    ESC_TRY(cudaMemcpyAsync(p_off, s.rec_off.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
    ESC_TRY(cudaMemcpyAsync(p_nnz, s.rec_nnz.p, (size_t)E * 4, cudaMemcpyDeviceToHost, st));
    if (!s.alias_input) {
        ESC_TRY(cudaMemcpyAsync(p_src, s.eo_src.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
        ESC_TRY(cudaMemcpyAsync(p_dst, s.eo_dst.p, (size_t)E * 8, cudaMemcpyDeviceToHost, st));
        ESC_TRY(cudaMemcpyAsync(p_ptr, s.eo_ptr.p, (size_t)(G + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    ESC_TRY(cudaStreamSynchronize(st));
    *h_rec = s.h_rec.as<uint32_t>();
    *h_rec_off = p_off;
    *h_rec_nnz = p_nnz;
    if (s.alias_input) {                          // no self-loop rewrite: the output edge list IS the (staged) input
        int64_t* st_src = s.h_in.as<int64_t>();
        *h_eo_src = st_src; *h_eo_dst = st_src + s.e_in; *h_eo_ptr = st_src + 2 * s.e_in;
    } else {
        *h_eo_src = p_src; *h_eo_dst = p_dst; *h_eo_ptr = p_ptr;
    }
    return 0;
}

// E6 on the host (utils_edge_efficient.py:139-151): compact records -> the int64 triple, edges in order, indices ascending.
// Two passes over graphs: record counts per graph (parallel), exclusive scan over graphs (serial, G entries), fill (parallel).
int escgnn_expand_records_host(const uint32_t* rec, const int64_t* rec_off, const int32_t* rec_nnz, const int64_t* eo_ptr,
                               int64_t n_graphs, int local_ordinals, int64_t* pos_enc, int64_t* pos_index, int64_t* pos_batch,
                               int threads) {
    if (n_graphs < 0 || !rec_off || !rec_nnz || !eo_ptr) return ESCGNN_ERR_BAD_ARG;
    if (n_graphs == 0) return 0;
    int64_t* base = (int64_t*)malloc((size_t)(n_graphs + 1) * sizeof(int64_t));
    if (!base) return ESCGNN_ERR_CAPACITY;
    const int nt = threads > 0 ? threads : 1;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t g = 0; g < n_graphs; ++g) {
        int64_t t = 0;
        for (int64_t e = eo_ptr[g]; e < eo_ptr[g + 1]; ++e) t += rec_nnz[e];
        base[g + 1] = t;
    }
    base[0] = 0;
    for (int64_t g = 0; g < n_graphs; ++g) base[g + 1] += base[g];
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int64_t g = 0; g < n_graphs; ++g) {
        int64_t o = base[g];
        for (int64_t e = eo_ptr[g]; e < eo_ptr[g + 1]; ++e) {
            const uint32_t* r = rec + rec_off[e];
            const int64_t ord = local_ordinals ? e - eo_ptr[g] : e;
            const int n = rec_nnz[e];
            for (int k = 0; k < n; ++k, ++o) {
                const uint32_t v = r[k];
                pos_index[o] = (int64_t)(v & ((1u << ESCGNN_REC_IDX_BITS) - 1u));
                pos_enc[o] = (int64_t)(v >> ESCGNN_REC_IDX_BITS);
                pos_batch[o] = ord;
            }
        }
    }
    free(base);
    return 0;
}

}  // extern "C"
