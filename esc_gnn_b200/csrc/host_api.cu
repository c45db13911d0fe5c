// Host-buffer front end of the encoder: the C-ABI call a Python `create_subgraphs` replacement makes when it holds
// CPU tensors (reference call sites: run_zinc.py:141-146, run_graphcount.py:404-408, run_ogb_mol.py:329-332).
// Owns a stream and a grow-only device workspace; one context per host thread.
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

}  // namespace

struct escgnn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Buf src, dst, edge_ptr, node_ptr, eo_src, eo_dst, eo_ptr, tmp, rdh, rec, rec_off, rec_nnz, edge_graph, out_off,
        scan_tmp, counters, scratch, pos_enc, pos_index, pos_batch;
    unsigned long long* h_counters = nullptr;   // pinned
    int64_t* h_ptr_tail = nullptr;              // pinned: eo_ptr[G]
    // last run
    int64_t n_graphs = 0, e_out = 0, nnz = 0;
    bool alias_input = false;
};

#define ESC_TRY(x) do { int _rc = (int)(x); if (_rc != 0) return _rc; } while (0)

extern "C" {

escgnn_ctx* escgnn_ctx_create(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    escgnn_ctx* c = new escgnn_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
    cudaMallocHost((void**)&c->h_counters, ESCGNN_NUM_COUNTERS * sizeof(unsigned long long));
    cudaMallocHost((void**)&c->h_ptr_tail, sizeof(int64_t));
    return c;
}

void escgnn_ctx_destroy(escgnn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    Buf* all[] = {&c->src, &c->dst, &c->edge_ptr, &c->node_ptr, &c->eo_src, &c->eo_dst, &c->eo_ptr, &c->tmp, &c->rdh,
                  &c->rec, &c->rec_off, &c->rec_nnz, &c->edge_graph, &c->out_off, &c->scan_tmp, &c->counters,
                  &c->scratch, &c->pos_enc, &c->pos_index, &c->pos_batch};
    for (Buf* b : all) b->release();
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_ptr_tail) cudaFreeHost(c->h_ptr_tail);
    cudaStreamDestroy(c->stream);
    delete c;
}

int escgnn_encode_host_run(escgnn_ctx* c, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                           const int64_t* h_node_ptr, int64_t G, int h, int use_rd, int self_loop,
                           int local_ordinals, int64_t* out_num_edges, int64_t* out_nnz, uint32_t* out_error_bits) {
    if (!c || G < 0 || h < 1 || h > 4) return ESCGNN_ERR_BAD_ARG;
    ESC_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->n_graphs = G; c->e_out = 0; c->nnz = 0;
    *out_num_edges = 0; *out_nnz = 0; *out_error_bits = 0;
    if (G == 0) return 0;
    const int64_t e_in = h_edge_ptr[G] - h_edge_ptr[0], n_tot = h_node_ptr[G] - h_node_ptr[0];
    if (h_edge_ptr[0] != 0 || h_node_ptr[0] != 0) return ESCGNN_ERR_BAD_ARG;
    int64_t max_n = 0, max_e = 0;
    for (int64_t g = 0; g < G; ++g) {
        const int64_t n = h_node_ptr[g + 1] - h_node_ptr[g], e = h_edge_ptr[g + 1] - h_edge_ptr[g];
        if (n < 0 || e < 0) return ESCGNN_ERR_BAD_ARG;
        if (n > max_n) max_n = n;
        const int64_t eo = self_loop ? e + n : e;
        if (eo > max_e) max_e = eo;
    }
    const int64_t e_cap = self_loop ? e_in + n_tot : e_in;
    ESC_TRY(c->src.ensure((size_t)(e_in + 1) * 8));
    ESC_TRY(c->dst.ensure((size_t)(e_in + 1) * 8));
    ESC_TRY(c->edge_ptr.ensure((size_t)(G + 1) * 8));
    ESC_TRY(c->node_ptr.ensure((size_t)(G + 1) * 8));
    ESC_TRY(c->counters.ensure(ESCGNN_NUM_COUNTERS * 8));
    ESC_TRY(c->rec_off.ensure((size_t)(e_cap + 1) * 8));
    ESC_TRY(c->rec_nnz.ensure((size_t)(e_cap + 1) * 4));
    ESC_TRY(c->edge_graph.ensure((size_t)(e_cap + 1) * 4));
    ESC_TRY(c->out_off.ensure((size_t)(e_cap + 2) * 8));
    ESC_TRY(c->scan_tmp.ensure((size_t)(e_cap / 1024 + 4) * 8));
    ESC_TRY(cudaMemcpyAsync(c->src.p, h_src, (size_t)e_in * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(c->dst.p, h_dst, (size_t)e_in * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(c->edge_ptr.p, h_edge_ptr, (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemcpyAsync(c->node_ptr.p, h_node_ptr, (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, st));
    ESC_TRY(cudaMemsetAsync(c->counters.p, 0, ESCGNN_NUM_COUNTERS * 8, st));
    const int64_t *eo_src, *eo_dst, *eo_ptr;
    c->alias_input = !self_loop;
    if (self_loop) {
        ESC_TRY(c->eo_src.ensure((size_t)(e_cap + 1) * 8));
        ESC_TRY(c->eo_dst.ensure((size_t)(e_cap + 1) * 8));
        ESC_TRY(c->eo_ptr.ensure((size_t)(G + 1) * 8));
        ESC_TRY(c->tmp.ensure((size_t)(4 * G + 8 * (G / 1024 + 2) + 64)));
        ESC_TRY(escgnn_rewrite_self_loops(c->src.as<int64_t>(), c->dst.as<int64_t>(), c->edge_ptr.as<int64_t>(),
                                          c->node_ptr.as<int64_t>(), G, c->eo_ptr.as<int64_t>(),
                                          c->eo_src.as<int64_t>(), c->eo_dst.as<int64_t>(), c->tmp.p, st));
        eo_src = c->eo_src.as<int64_t>(); eo_dst = c->eo_dst.as<int64_t>(); eo_ptr = c->eo_ptr.as<int64_t>();
    } else {
        eo_src = c->src.as<int64_t>(); eo_dst = c->dst.as<int64_t>(); eo_ptr = c->edge_ptr.as<int64_t>();
    }
    const uint16_t* rdh = nullptr;
    int64_t sb = escgnn_encode_scratch_bytes(max_n, max_e, h);
    if (use_rd) {
        const int64_t sb_rd = escgnn_encode_rd_scratch_bytes(max_n, max_e, h);
        if (sb_rd > sb) sb = sb_rd;
    }
    if (sb > 0) ESC_TRY(c->scratch.ensure((size_t)sb));
    if (use_rd) {
        ESC_TRY(c->rdh.ensure((size_t)(e_cap + 1) * ESCGNN_RD_SLOTS * 2));
        ESC_TRY(escgnn_encode_rd(eo_src, eo_dst, eo_ptr, c->node_ptr.as<int64_t>(), G, h, c->rdh.as<uint16_t>(),
                                 c->counters.as<unsigned long long>(), max_n, max_e, c->scratch.p,
                                 (int64_t)c->scratch.cap, st));
        rdh = c->rdh.as<uint16_t>();
    }
    int64_t want = e_cap * 48 + 1024;
    if ((int64_t)(c->rec.cap / 4) < want) ESC_TRY(c->rec.ensure((size_t)want * 4));
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int64_t rec_cap = (int64_t)(c->rec.cap / 4);
        ESC_TRY(escgnn_encode(eo_src, eo_dst, eo_ptr, c->node_ptr.as<int64_t>(), G, h, rdh, c->rec.as<uint32_t>(),
                              rec_cap, c->rec_off.as<int64_t>(), c->rec_nnz.as<int32_t>(),
                              c->edge_graph.as<int32_t>(), c->counters.as<unsigned long long>(), max_n, max_e,
                              c->scratch.p, (int64_t)c->scratch.cap, st));
        ESC_TRY(cudaMemcpyAsync(c->h_counters, c->counters.p, ESCGNN_NUM_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
        ESC_TRY(cudaMemcpyAsync(c->h_ptr_tail, eo_ptr + G, 8, cudaMemcpyDeviceToHost, st));
        ESC_TRY(cudaStreamSynchronize(st));
        const int64_t nnz = (int64_t)c->h_counters[ESCGNN_CTR_NNZ];
        if (nnz <= rec_cap) break;
        if (attempt == 1) return ESCGNN_ERR_CAPACITY;
        ESC_TRY(c->rec.ensure((size_t)nnz * 4));
        ESC_TRY(cudaMemsetAsync(c->counters.as<unsigned long long>() + ESCGNN_CTR_NNZ, 0, 8, st));
        ESC_TRY(cudaMemsetAsync(c->counters.as<unsigned long long>() + ESCGNN_CTR_TICKET, 0, 8, st));
    }
    c->e_out = *c->h_ptr_tail;
    c->nnz = (int64_t)c->h_counters[ESCGNN_CTR_NNZ];
    *out_error_bits = (uint32_t)c->h_counters[ESCGNN_CTR_ERROR];
    *out_num_edges = c->e_out;
    *out_nnz = c->nnz;
    if (*out_error_bits) return ESCGNN_ERR_DATA;
    ESC_TRY(escgnn_exclusive_scan_i32(c->rec_nnz.as<int32_t>(), c->e_out, c->out_off.as<int64_t>(),
                                      c->scan_tmp.as<int64_t>(), st));
    ESC_TRY(c->pos_enc.ensure((size_t)(c->nnz + 1) * 8));
    ESC_TRY(c->pos_index.ensure((size_t)(c->nnz + 1) * 8));
    ESC_TRY(c->pos_batch.ensure((size_t)(c->nnz + 1) * 8));
    ESC_TRY(escgnn_expand_records(c->rec.as<uint32_t>(), c->rec_off.as<int64_t>(), c->rec_nnz.as<int32_t>(),
                                  c->edge_graph.as<int32_t>(), eo_ptr, c->out_off.as<int64_t>(), c->e_out, use_rd,
                                  local_ordinals, c->pos_enc.as<int64_t>(), c->pos_index.as<int64_t>(),
                                  c->pos_batch.as<int64_t>(), st));
    return 0;
}

int escgnn_encode_host_fetch(escgnn_ctx* c, int64_t* h_eo_src, int64_t* h_eo_dst, int64_t* h_eo_ptr,
                             int64_t* h_pos_enc, int64_t* h_pos_index, int64_t* h_pos_batch) {
    if (!c) return ESCGNN_ERR_BAD_ARG;
    ESC_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const void* es = c->alias_input ? c->src.p : c->eo_src.p;
    const void* ed = c->alias_input ? c->dst.p : c->eo_dst.p;
    const void* ep = c->alias_input ? c->edge_ptr.p : c->eo_ptr.p;
    if (c->n_graphs > 0) {
        if (h_eo_src) ESC_TRY(cudaMemcpyAsync(h_eo_src, es, (size_t)c->e_out * 8, cudaMemcpyDeviceToHost, st));
        if (h_eo_dst) ESC_TRY(cudaMemcpyAsync(h_eo_dst, ed, (size_t)c->e_out * 8, cudaMemcpyDeviceToHost, st));
        if (h_eo_ptr) ESC_TRY(cudaMemcpyAsync(h_eo_ptr, ep, (size_t)(c->n_graphs + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_enc) ESC_TRY(cudaMemcpyAsync(h_pos_enc, c->pos_enc.p, (size_t)c->nnz * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_index) ESC_TRY(cudaMemcpyAsync(h_pos_index, c->pos_index.p, (size_t)c->nnz * 8, cudaMemcpyDeviceToHost, st));
        if (h_pos_batch) ESC_TRY(cudaMemcpyAsync(h_pos_batch, c->pos_batch.p, (size_t)c->nnz * 8, cudaMemcpyDeviceToHost, st));
    }
    return (int)cudaStreamSynchronize(st);
}

int escgnn_encode_host_device_results(escgnn_ctx* c, const uint32_t** d_rec, const int64_t** d_rec_off,
                                      const int32_t** d_rec_nnz, const int64_t** d_eo_src, const int64_t** d_eo_dst) {
    if (!c) return ESCGNN_ERR_BAD_ARG;
    if (d_rec) *d_rec = c->rec.as<uint32_t>();
    if (d_rec_off) *d_rec_off = c->rec_off.as<int64_t>();
    if (d_rec_nnz) *d_rec_nnz = c->rec_nnz.as<int32_t>();
    if (d_eo_src) *d_eo_src = c->alias_input ? c->src.as<int64_t>() : c->eo_src.as<int64_t>();
    if (d_eo_dst) *d_eo_dst = c->alias_input ? c->dst.as<int64_t>() : c->eo_dst.as<int64_t>();
    return 0;
}

}  // extern "C"
