/* escgnn_b200 -- C-ABI of the B200-native ESC-GNN hot paths (libescgnn_b200.so).
 *
 * The reference (pkuyzy/ESC-GNN) is pure Python on these paths and has no native FFI boundary of its own
 * (SURVEY.md section 8b); the entry points below are what a maintainer binds from Python (ctypes stub in
 * INTEGRATION.md) to replace the bodies of the reference functions cited on each declaration.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every `d_` pointer is DEVICE memory, every `h_` pointer HOST memory
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); device entry points never synchronise
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or an ESCGNN_ERR_* code (<0); never throws
 *   - data-dependent errors (degree >= 200, bad node id ...) are reported through `d_counters[ESCGNN_CTR_ERROR]`
 */
#ifndef ESCGNN_B200_H
#define ESCGNN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESCGNN_ERR_BAD_ARG (-1)
#define ESCGNN_ERR_TOO_LARGE (-2)     /* a graph exceeds the supported node/edge count of the kernels */
#define ESCGNN_ERR_CAPACITY (-3)      /* caller-provided record capacity too small (required size reported) */
#define ESCGNN_ERR_DATA (-4)          /* data-dependent error, see ESCGNN_CTR_ERROR bits */

/* bits of d_counters[ESCGNN_CTR_ERROR] -- mirror the exceptions the reference raises */
#define ESCGNN_DATA_DEG 1u            /* sub_degree >= 200: F.one_hot(..., 200) raises (utils_edge_efficient.py:128) */
#define ESCGNN_DATA_NODE 2u           /* node id outside [0, N) */
#define ESCGNN_DATA_RD 4u             /* rd bin outside the representable range (:131) / singular system */
#define ESCGNN_DATA_ASYM 8u           /* use_rd on a non-symmetric edge multiset (unsupported, documented) */

/* slots of the uint64 counter block every encoder launch works on (caller zeroes it, or uses *_host) */
#define ESCGNN_CTR_NNZ 0              /* total records emitted (bump cursor) */
#define ESCGNN_CTR_ERROR 1            /* OR of ESCGNN_DATA_* */
#define ESCGNN_CTR_TICKET 2           /* persistent-CTA work ticket */
#define ESCGNN_CTR_TICKET_RD 3
#define ESCGNN_CTR_STICKY_ERROR 4     /* OR of every ERROR word seen by escgnn_make_dims since the caller last cleared it */
#define ESCGNN_CTR_MAX_NNZ 5          /* max of every NNZ seen by escgnn_make_dims (record-capacity overflow detection) */
#define ESCGNN_CTR_RD_DECLINED 6      /* edges the cycle-space rd kernel left to the general solver (zeroed by escgnn_encode_rd itself) */
#define ESCGNN_CTR_PER_CALL 4         /* slots [0, 4) are per-call state: a caller that keeps the sticky slots zeroes only these */
#define ESCGNN_NUM_COUNTERS 8

#define ESCGNN_RD_SLOTS 12            /* rd histogram slots kept per edge (bins 0..11; R(u,w) <= 2h+1 <= 9) */

/* A record packs one non-zero of the per-edge encoding vector: index (11 bits, < 1800) | count << 11. */
#define ESCGNN_REC_IDX_BITS 11

int escgnn_version(void);

/* ---- E1: self-loop rewrite.  Replaces utils_edge_efficient.py:33-38 (remove_self_loops + add_self_loops).
 * Input: G graphs, graph-local int64 node ids, d_edge_ptr[G+1], d_node_ptr[G+1].
 * Output: d_eo_ptr[G+1] and the rewritten edge list (capacity E_in + total nodes). With self_loop == 0 the
 * caller aliases the input instead of calling this. d_tmp: scratch of 4*G + 8*(G/1024 + 2) + 16 bytes. */
int escgnn_rewrite_self_loops(const int64_t* d_src, const int64_t* d_dst, const int64_t* d_edge_ptr,
                              const int64_t* d_node_ptr, int64_t n_graphs, int64_t* d_eo_ptr, int64_t* d_eo_src,
                              int64_t* d_eo_dst, void* d_tmp, void* stream);

/* ---- E2-E4 (+E6 compact form): per-edge ego-net encoding.  Replaces utils_edge_efficient.py:41-90,122-144
 * and k_hop_subgraph (:201-294).  One CTA per graph (persistent, ticketed); N bounded BFS per graph; warp per
 * edge builds the histogram in shared memory and emits ascending (index,count) records.
 *   d_eo_*      edge list after E1, d_eo_ptr[G+1] / d_node_ptr[G+1]
 *   d_rdh       optional [E_out][ESCGNN_RD_SLOTS] uint16 rd histogram from escgnn_encode_rd (NULL = use_rd off)
 *   d_rec       uint32 records, capacity rec_cap; d_rec_off[E_out] int64, d_rec_nnz[E_out] int32,
 *               d_edge_graph[E_out] int32
 *   d_counters  ESCGNN_NUM_COUNTERS uint64, zeroed by the caller; NNZ may exceed rec_cap (then nothing past the
 *               capacity was written and the caller retries with a larger buffer)
 *   d_scratch   global scratch for graphs that do not fit in shared memory (scratch_bytes, may be 0/NULL when
 *               max_nodes/max_edges fit); max_nodes/max_edges: maxima over the batch (host knows them from ptrs) */
int escgnn_encode(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                  const int64_t* d_node_ptr, int64_t n_graphs, int h, const uint16_t* d_rdh, uint32_t* d_rec,
                  int64_t rec_cap, int64_t* d_rec_off, int32_t* d_rec_nnz, int32_t* d_edge_graph,
                  unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                  int64_t scratch_bytes, void* stream);
/* The same over a SUBSET of the batch's graphs (d_graph_ids[n_graphs], int32): callers bucket graphs by size and launch
 * once per bucket, so small graphs are not forced to the shared-memory footprint (= occupancy) of the largest one.
 * max_nodes / max_edges are the maxima over the subset. All subsets share the record buffer and the counters. */
int escgnn_encode_subset(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                         const int64_t* d_node_ptr, int64_t n_graphs, const int32_t* d_graph_ids, int h, const uint16_t* d_rdh,
                         uint32_t* d_rec, int64_t rec_cap, int64_t* d_rec_off, int32_t* d_rec_nnz, int32_t* d_edge_graph,
                         unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                         int64_t scratch_bytes, void* stream);
/* scratch bytes escgnn_encode / escgnn_encode_rd need for a batch with these maxima (0 if everything fits on chip) */
int64_t escgnn_encode_scratch_bytes(int64_t max_nodes, int64_t max_edges, int h);
int64_t escgnn_encode_rd_scratch_bytes(int64_t max_nodes, int64_t max_edges, int h);

/* ---- E5: resistance-distance histogram per edge.  Replaces utils_edge_efficient.py:92-107,130-131 under parity
 * policy E5 (float64 solve, bin = trunc((float)rd)).  Writes d_rdh[E_out][ESCGNN_RD_SLOTS]. */
int escgnn_encode_rd(const int64_t* d_eo_src, const int64_t* d_eo_dst, const int64_t* d_eo_ptr,
                     const int64_t* d_node_ptr, int64_t n_graphs, int h, uint16_t* d_rdh,
                     unsigned long long* d_counters, int64_t max_nodes, int64_t max_edges, void* d_scratch,
                     int64_t scratch_bytes, void* stream);

/* ---- E6: expand compact records into the reference's int64 triple (pos_enc, pos_index, pos_batch), edges in
 * order, indices ascending (utils_edge_efficient.py:139-151).  d_out_off[E_out+1] = exclusive scan of d_rec_nnz
 * (escgnn_exclusive_scan_i32).  pos_batch = edge ordinal minus d_eo_ptr[graph] when local_ordinals != 0
 * (per-graph Data), else the batch-wide ordinal (what batch.py:70-71 produces after collation). */
int escgnn_exclusive_scan_i32(const int32_t* d_in, int64_t n, int64_t* d_out /* [n+1] */, int64_t* d_tmp /* [n/1024+2] */,
                              void* stream);
int escgnn_expand_records(const uint32_t* d_rec, const int64_t* d_rec_off, const int32_t* d_rec_nnz,
                          const int32_t* d_edge_graph, const int64_t* d_eo_ptr, const int64_t* d_out_off,
                          int64_t n_edges, int use_rd, int local_ordinals, int64_t* d_pos_enc, int64_t* d_pos_index,
                          int64_t* d_pos_batch, void* stream);

/* ---- Host-buffer front end (what `create_subgraphs` binds when the caller holds CPU tensors): H2D, E1, E5, E2-E4,
 * scan, E6, D2H in one call on an internal stream with a grow-only device workspace.
 * Two-step because the output size is data dependent: `_run` leaves results on the device and returns sizes,
 * `_fetch` copies them into caller buffers of exactly those sizes. */
typedef struct escgnn_ctx escgnn_ctx;
escgnn_ctx* escgnn_ctx_create(int device);
void escgnn_ctx_destroy(escgnn_ctx* ctx);
int escgnn_encode_host_run(escgnn_ctx* ctx, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                           const int64_t* h_node_ptr, int64_t n_graphs, int h, int use_rd, int self_loop,
                           int local_ordinals, int64_t* out_num_edges, int64_t* out_nnz, uint32_t* out_error_bits);
int escgnn_encode_host_fetch(escgnn_ctx* ctx, int64_t* h_eo_src, int64_t* h_eo_dst, int64_t* h_eo_ptr,
                             int64_t* h_pos_enc, int64_t* h_pos_index, int64_t* h_pos_batch);
/* device pointers of the last run's compact results (valid until the next run on this ctx) */
int escgnn_encode_host_device_results(escgnn_ctx* ctx, const uint32_t** d_rec, const int64_t** d_rec_off,
                                      const int32_t** d_rec_nnz, const int64_t** d_eo_src, const int64_t** d_eo_dst);

/* Throughput form of the host front end: two slots per context, each with its own stream, device workspace and grow-only
 * PINNED staging / result arenas (no allocation after warm-up).
 *   _submit(slot): copies the inputs into the slot's pinned staging (the caller's buffers are free on return), queues H2D and the
 *                  E1 / E5 / E2-E4 kernels on the slot's stream and returns without waiting.
 *   _wait(slot):   waits for the kernels, then fetches the COMPACT result into the slot's pinned arena -- records uint32 [nnz]
 *                  (index | count << ESCGNN_REC_IDX_BITS, ascending inside an edge; the records of edge e are
 *                  h_rec[h_rec_off[e] .. + h_rec_nnz[e])), and the output edge list (graph-local ids, after E1) with its
 *                  per-graph pointers.  4 bytes per record cross PCIe instead of the 24 of the int64 triple.  The returned
 *                  pointers stay valid until the slot is submitted again.
 * Submitting chunk k+1 on the other slot before waiting for chunk k overlaps the D2H of k with the kernels of k+1.
 * escgnn_expand_records_host: E6 on the host (OpenMP over graphs) for callers that need the reference's int64 triple
 * (utils_edge_efficient.py:139-151); pos_batch = per-graph edge ordinal when local_ordinals != 0, else the batch-wide one. */
int escgnn_encode_host_submit(escgnn_ctx* ctx, int slot, const int64_t* h_src, const int64_t* h_dst, const int64_t* h_edge_ptr,
                              const int64_t* h_node_ptr, int64_t n_graphs, int h, int use_rd, int self_loop);
int escgnn_encode_host_wait(escgnn_ctx* ctx, int slot, int64_t* out_num_edges, int64_t* out_nnz, uint32_t* out_error_bits,
                            const uint32_t** h_rec, const int64_t** h_rec_off, const int32_t** h_rec_nnz,
                            const int64_t** h_eo_src, const int64_t** h_eo_dst, const int64_t** h_eo_ptr);
int escgnn_expand_records_host(const uint32_t* h_rec, const int64_t* h_rec_off, const int32_t* h_rec_nnz, const int64_t* h_eo_ptr,
                               int64_t n_graphs, int local_ordinals, int64_t* h_pos_enc, int64_t* h_pos_index,
                               int64_t* h_pos_batch, int threads);

/* ================= model step; fp32, row-major, device pointers ========
 * Static-shape convention: where an entry point takes `const int* d_count` (may be NULL), the size argument before
 * it is a CAPACITY and the actual count is read from device memory, so one captured CUDA graph serves every batch;
 * rows between the count and the capacity are written as zeros. */

/* Deterministic CSR of an int64 key vector (edge_index[1] for the forward aggregation, edge_index[0] for the
 * backward): d_ptr[n_nodes+1], d_perm[n_edges] = edge ids grouped by key, ascending inside a group.
 * d_tmp: n_nodes+1 int32 scratch. d_err (optional): bit 0 set when a key is outside [0, n_nodes). */
int escgnn_csr_build(const int64_t* d_keys, int64_t n_edges, int64_t n_nodes, int32_t* d_ptr, int32_t* d_perm,
                     int32_t* d_tmp, unsigned long long* d_err, const int* d_count, void* stream);
/* Segment pointers of a sorted id vector (pos_batch -> per-edge record ranges, batch -> per-graph node ranges;
 * replaces the `int(batch.max())+1` + scatter bookkeeping of PyG global_add_pool). d_ptr[n_segments+1]. */
int escgnn_sorted_ids_to_ptr(const int64_t* d_ids, int64_t n, int64_t n_segments, int32_t* d_ptr, const int* d_count,
                             void* stream);

/* M1 sparse bag-embed.  Replaces global_add_pool(z_initial.weight[pos_index] * pos_enc[:,None], pos_batch)
 * (run_graphcount.py:155, zinc_models.py:590, ogb_mol_gnn.py:716) without the [nnz,H] intermediate.
 * Sparse operand either as the reference's int64 triple (d_pos_index, d_pos_enc, d_ptr from
 * escgnn_sorted_ids_to_ptr(pos_batch)) or, when d_rec != NULL, as the encoder's packed records. hidden % 4 == 0.
 * bwd accumulates into d_grad_weight [1800,hidden] (caller zeroes it). */
int escgnn_bag_embed_fwd(const float* d_weight, int hidden, const int64_t* d_pos_index, const int64_t* d_pos_enc,
                         const int32_t* d_ptr, const uint32_t* d_rec, const int64_t* d_rec_off,
                         const int32_t* d_rec_nnz, int64_t n_edges, float* d_out, const int* d_count, void* stream);
int escgnn_bag_embed_bwd(const float* d_grad, int hidden, const int64_t* d_pos_index, const int64_t* d_pos_enc,
                         const int32_t* d_ptr, const uint32_t* d_rec, const int64_t* d_rec_off,
                         const int32_t* d_rec_nnz, int64_t n_edges, float* d_grad_weight, const int* d_count,
                         void* stream);

/* Same gradient, index-major: the records are transposed on the device (count / scan / fill) and reduced in balanced
 * chunks, so the indices present in every edge do not serialise on atomics. Packed records only. rec_cap bounds the
 * record count (grid sizing); d_work: 3*1800+1 int32; d_sorted_edge / d_sorted_cnt: rec_cap entries each. */
int escgnn_bag_embed_bwd_sorted(const float* d_grad, int hidden, const uint32_t* d_rec, const int64_t* d_rec_off,
                                const int32_t* d_rec_nnz, int64_t n_edges, int64_t rec_cap, float* d_grad_weight,
                                int32_t* d_work, int32_t* d_sorted_edge, float* d_sorted_cnt, const int* d_count,
                                void* stream);

/* The two halves of escgnn_bag_embed_bwd_sorted: the index-major transposition depends only on the encoding (it can be
 * built once per batch, off the critical path), the reduction on the incoming gradient. */
int escgnn_bag_index_build(const uint32_t* d_rec, const int64_t* d_rec_off, const int32_t* d_rec_nnz, int64_t n_edges,
                           int32_t* d_work, int32_t* d_sorted_edge, float* d_sorted_cnt, const int* d_count, void* stream);
int escgnn_bag_embed_bwd_indexed(const float* d_grad, int hidden, int64_t rec_cap, float* d_grad_weight,
                                 const int32_t* d_work, const int32_t* d_sorted_edge, const float* d_sorted_cnt, void* stream);

/* M3 GINE aggregation.  Replaces PyG GINEConv.propagate + the (1+eps)*x residual (in-tree twin:
 * GraphGPS/graphgps/layer/gine_conv_layer.py:56-84; ogb_mol_gnn.py:346-358):
 *   out[i] = (1+eps) x[i] + sum_{e: dst_e = i} relu(x[src_e] + edge_feat[e]).
 * d_src/d_dst: the int64 rows of edge_index; (d_dst_ptr, d_dst_perm) / (d_src_ptr, d_src_perm) from escgnn_csr_build.
 * bwd writes d_grad_x [N,C], d_grad_edge_feat [E,C], per-node <g_out, x> into d_node_dots [N] and their sum into
 * d_grad_eps[0] (optional). */
int escgnn_gine_aggregate_fwd(const float* d_x, const float* d_edge_feat, const int64_t* d_src, const int32_t* d_dst_ptr,
                              const int32_t* d_dst_perm, const float* d_eps, int64_t n_nodes, int channels,
                              float* d_out, const int* d_count, void* stream);
int escgnn_gine_aggregate_bwd(const float* d_grad_out, const float* d_x, const float* d_edge_feat, const int64_t* d_dst,
                              const int32_t* d_src_ptr, const int32_t* d_src_perm, const float* d_eps, int64_t n_nodes,
                              int channels, float* d_grad_x, float* d_grad_edge_feat, float* d_node_dots,
                              float* d_grad_eps, const int* d_count, void* stream);

/* the same with explicit leading dimensions: x / out / gradients may be column slices of wider buffers (the JK concat
 * buffer, one projection of all layers' edge features), so no slice is ever copied */
int escgnn_gine_aggregate_fwd_ld(const float* d_x, int ldx, const float* d_edge_feat, int lde, const int64_t* d_src,
                                 const int32_t* d_dst_ptr, const int32_t* d_dst_perm, const float* d_eps, int64_t n_nodes,
                                 int channels, float* d_out, int ldo, const int* d_count, void* stream);
int escgnn_gine_aggregate_bwd_ld(const float* d_grad_out, int ldg, const float* d_x, int ldx, const float* d_edge_feat, int lde,
                                 const int64_t* d_dst, const int32_t* d_src_ptr, const int32_t* d_src_perm, const float* d_eps,
                                 int64_t n_nodes, int channels, float* d_grad_x, int ldgx, float* d_grad_edge_feat,
                                 float* d_node_dots, float* d_grad_eps, const int* d_count, void* stream);

/* M4 pooling over the sorted `batch` vector.  Replaces global_add_pool / global_mean_pool
 * (run_graphcount.py:179, zinc_models.py:602, ogb_mol_gnn.py:124,768). mean divides by max(count,1). */
int escgnn_segment_pool_fwd(const float* d_x, const int32_t* d_ptr, int64_t n_segments, int channels, int mean,
                            float* d_out, void* stream);
int escgnn_segment_pool_bwd(const float* d_grad, const int32_t* d_ptr, int64_t n_segments, int channels, int mean,
                            float* d_grad_x, void* stream);

/* D1 `Distance` edge transform.  Replaces distance.py:29-47: dist = ||pos[col]-pos[row]|| (squared != 0: its
 * square), divided by the data maximum (norm != 0, max_value <= 0) or by max_value; d_rel (optional) receives
 * pos[col]-pos[row] [E,dim]. d_max_scratch: one uint32. */
int escgnn_edge_distance(const float* d_pos, int dim, const int64_t* d_row, const int64_t* d_col, int64_t n_edges,
                         int squared, int norm, float max_value, float* d_dist, float* d_rel, unsigned* d_max_scratch,
                         void* stream);

/* All-pairs shortest-path lengths per graph on the UNDIRECTED edge set, unreachable pairs = `unreachable` (100 in the
 * reference). Replaces the networkx loop that fills `attn_bias` in the GraphGPS twin of the transform
 * (GraphGPS/graphgps/loader/utils_escgnn.py:29-38). d_out_ptr[g] = sum_{g'<g} n_g'^2; graph g's [n_g, n_g] block is
 * written row-major at d_out + d_out_ptr[g]. max_nodes / max_edges: per-graph maxima (shared-memory sizing). */
int64_t escgnn_all_pairs_spd_smem_bytes(int64_t max_nodes, int64_t max_edges);
int escgnn_all_pairs_spd(const int64_t* d_src, const int64_t* d_dst, const int64_t* d_edge_ptr, const int64_t* d_node_ptr,
                         int64_t n_graphs, const int64_t* d_out_ptr, int64_t* d_out, int64_t max_nodes, int64_t max_edges,
                         int64_t unreachable, unsigned long long* d_counters, void* stream);

/* M5 Adam over one flat fp32 buffer (torch.optim.Adam rule, no amsgrad / weight decay; reference call sites
 * run_graphcount.py:478,505, run_zinc.py:263, run_ogb_mol.py:436). `step` counts from 1; grad_scale multiplies
 * the gradient first (1/world_size after a sum all-reduce). */
int escgnn_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream);
/* same update with the step counter and hyper-parameters in device memory (graph-capturable):
 * d_hyper[7] = {lr, beta1, beta2, eps, grad_scale, bc1 (out), sqrt(bc2) (out)}, d_state[1] = step (incremented). n % 4 == 0. */
int escgnn_adam_step_device(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                            float* d_hyper, long long* d_state, void* stream);

/* B1 collation on the device (reference batch.py:52-123 + Data.__inc__): shift graph-local edge ids by the
 * graph's node offset (d_edge_graph from escgnn_encode), and expand node_ptr into the `batch` vector. */
int escgnn_collate_edges(const int64_t* d_src, const int64_t* d_dst, const int32_t* d_edge_graph,
                         const int64_t* d_node_ptr, int64_t n_edges, int64_t* d_out_src, int64_t* d_out_dst,
                         const int* d_count, void* stream);
int escgnn_ptr_to_ids(const int64_t* d_ptr, int64_t n_segments, int64_t n, int64_t* d_ids, const int* d_count,
                      void* stream);
/* out[i] = x[i] + v[segment of i] (virtual-node broadcast, ogb_mol_gnn.py:737; x == out allowed); d_ptr int32 [n_segments+1] */
int escgnn_add_segment_rows(const float* d_x, int ldx, const float* d_v, int ldv, const int32_t* d_ptr, int64_t n_segments,
                            int channels, float* d_out, int ldo, void* stream);
/* E1 on int64 edge attributes [E_in, attr_cols]: the permutation of escgnn_rewrite_self_loops (loop rows dropped, order kept,
 * `fill` rows appended per node; utils_edge_efficient.py:35-36 -> add_self_loops fills 1) */
int escgnn_rewrite_edge_attr(const int64_t* d_src, const int64_t* d_dst, const int64_t* d_edge_ptr, const int64_t* d_node_ptr,
                             int64_t n_graphs, const int64_t* d_eo_ptr, const int64_t* d_attr, int attr_cols, int64_t fill,
                             int64_t* d_out, void* stream);
/* d_out[0] (+)= sum of d_v[0..n) in a fixed order (the eps gradient of a GINE layer from its per-node dot products) */
int escgnn_reduce_sum(const float* d_v, int64_t n, float* d_out, int accumulate, void* stream);
/* rows [*d_rows, rows_cap) of a [rows_cap, ld] fp32 buffer := 0 (first `cols` columns) */
int escgnn_zero_tail_rows(float* d_x, int ld, int cols, const int* d_rows, int64_t rows_cap, void* stream);
/* d_dims[4] = {total nodes, total edges after E1, graphs, records}: the device-side sizes the static-shape
 * entry points read through their d_count / d_rows arguments (no host sync between encoder and model).
 * Also folds this call's ERROR / NNZ words into the sticky slots ESCGNN_CTR_STICKY_ERROR / ESCGNN_CTR_MAX_NNZ of
 * d_counters, so a caller that zeroes only the per-call slots [0, ESCGNN_CTR_PER_CALL) every step can check ONCE per
 * epoch whether ANY batch had a data error or exceeded its record capacity. */
int escgnn_make_dims(const int64_t* d_eo_ptr, const int64_t* d_node_ptr, int64_t n_graphs,
                     unsigned long long* d_counters, int* d_dims, void* stream);

/* ---- dense row-wise kernels of the static-shape engine (M2, M4, M5); `d_rows` = actual row count on the device,
 * rows_cap = capacity; ld* = leading dimensions (outputs can be column slices of a wider buffer = free concat) ---- */
/* Programmatic dependent launch for the model-side kernels (csrc/launch.cuh): on by default; 0 launches them with plain
 * stream ordering (same results, used for A/B timing). Returns the previous setting. */
int escgnn_set_pdl(int on);
/* Cap the persistent encoder grids (ego_encode / ego_rd) at `ctas` CTAs (0 = fill the machine, the default): used by the
 * pipelined training step, where the encoder of the next batch has a whole step of slack and should leave most SMs to the
 * latency-critical kernels of the current batch. Returns the previous cap. */
int escgnn_set_encoder_grid_cap(int ctas);
/* Pendant-tree peeling in the resistance-distance kernel (graphs of at most 64 nodes; see csrc/rd.cu): on by default, 0 solves the
 * full system of every pair (same histograms; A/B timing and tests). Returns the previous setting. */
int escgnn_set_rd_peel(int on);
/* Cycle-space fast path of the resistance-distance block (csrc/rd_fast.cuh: a warp per graph, a thread per pair system; symmetric
 * simple graphs of at most 128 nodes whose ego-nets hold at most 4 independent cycles -- everything else falls through to the
 * LDL^T / Takahashi solver): on by default, 0 = general solver for every pair (same histograms; A/B timing and tests).
 * Replaces utils_edge_efficient.py:92-107,130-131 like escgnn_encode_rd itself. Returns the previous setting. */
int escgnn_set_rd_fast(int on);
/* One-launch cluster BatchNorm kernels (rows_cap <= 65536, training mode): on by default; 0 = statistics + apply kernel pair
 * (same results up to summation order). Returns the previous setting. */
int escgnn_set_cluster_bn(int on);
/* Debugging aid (tools/trace_bn.py): every CTA of the following register-resident cluster BatchNorm launches writes 6 %globaltimer
 * values (ns) to d_stamps[cta * 6 + i]: 0 CTA start, 1 dependency wait returned, 2 tile loaded and summed, 3 cluster reduction done,
 * 4 outputs stored, 5 cluster released.  NULL (default) switches it off. */
int escgnn_bn_set_trace(unsigned long long* d_stamps);
int escgnn_dense_tile_rows(void);     /* rows per reduction tile of the scalar fallback kernels */
/* floats the reduction workspace `d_partial` of bn_act_fwd / bn_act_bwd / colsum needs. It must be zero before its
 * first use (its first 64 words are arrival tickets, which every launch leaves at zero again), and two launches that
 * may run concurrently (different streams) need different workspaces. */
int64_t escgnn_dense_partial_floats(int rows_cap, int channels);
/* training-mode BatchNorm1d + activation (act: 0 none, 1 ReLU, 2 ELU). Replaces the BN,act pairs of
 * nn.Sequential(Linear,Dropout,BN,act,...) (run_graphcount.py:54-61,78-87; zinc_models.py:513-522). Updates the
 * running statistics like torch (momentum, unbiased variance); saves mean / rstd for the backward. */
int escgnn_bn_act_fwd(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float* d_running_mean,
                      float* d_running_var, float* d_mean, float* d_rstd, float* d_partial, int act, float eps,
                      float momentum, int training, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                      void* stream);
/* backward through activation + BatchNorm; the incoming gradient is d_dy (+ d_dy2 when not NULL: the second
 * consumer of a concatenated layer output). Writes d_dx, d_dgamma, d_dbeta. */
int escgnn_bn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, const float* d_dy2, int lddy2,
                      const float* d_mean, const float* d_rstd, const float* d_gamma, const float* d_beta, int act,
                      int training, float* d_partial, const int* d_rows, int rows_cap, int channels, float* d_dgamma,
                      float* d_dbeta, float* d_dx, int lddx, void* stream);
/* Graph-level readout tail of the ZINC model in one launch (zinc_models.py:604-609 + L1 loss, run_zinc.py:276-281), forward
 * AND backward: p2 = act(BN(x)), pred = p2 w2 + b2, loss = mean |pred - target|, then d_dx (gradient wrt x), d_dgamma,
 * d_dbeta, d_dw2 [channels], d_db2 [1]. channels <= 256, rows_cap <= 512 (ESCGNN_ERR_TOO_LARGE otherwise: use the separate
 * entry points). Training-mode statistics; running statistics updated like torch. */
int escgnn_head_bn_linear_l1(const float* d_x, int ldx, const float* d_gamma, const float* d_beta, float* d_running_mean,
                             float* d_running_var, int act, float eps, float momentum, const float* d_w2, const float* d_b2,
                             const float* d_target, const int* d_rows, int rows_cap, int channels, float* d_pred, float* d_loss,
                             float* d_dx, int lddx, float* d_dgamma, float* d_dbeta, float* d_dw2, float* d_db2, void* stream);
/* inverted dropout (F.dropout, ogb_mol_gnn.py:756-758,774; run_graphcount.py:54-61): y = keep ? x / (1 - p) : 0 with a
 * counter-based mask keep = hash(salt, *d_step, position) >= p. The same call on the gradient is the backward pass (nothing
 * is stored); d_step (may be NULL) is a device-side step counter, so a captured graph draws a fresh mask every replay. */
int escgnn_dropout(const float* d_x, int ldx, float p, uint32_t salt, const long long* d_step, const int* d_rows, int rows_cap,
                   int channels, float* d_y, int ldy, void* stream);
int escgnn_act_fwd(const float* d_x, int ldx, int act, const int* d_rows, int rows_cap, int channels, float* d_y, int ldy,
                   void* stream);
int escgnn_act_bwd(const float* d_x, int ldx, const float* d_dy, int lddy, int act, const int* d_rows, int rows_cap,
                   int channels, float* d_dx, int lddx, void* stream);
/* ordered column sum over the valid rows (bias gradients of nn.Linear) */
int escgnn_colsum(const float* d_x, int ldx, const int* d_rows, int rows_cap, int channels, float* d_partial, float* d_out,
                  void* stream);
/* sum of per-column embeddings: y[r] = sum_k table[idx[r,k] + col_offsets[k]] (nn.Embedding when idx_cols == 1;
 * AtomEncoder / BondEncoder of ogb_mol_gnn.py:264-282 otherwise); bwd accumulates into d_dtable (caller zeroes) */
int escgnn_embedding_fwd(const float* d_table, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_y, int ldy, void* stream);
int escgnn_embedding_bwd(const float* d_dy, int lddy, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                         const int* d_rows, int rows_cap, int channels, float* d_dtable, void* stream);
/* same gradient for a SMALL table (table_rows * channels * 4 <= 48 KB, e.g. BondEncoder's 13 rows): per-CTA accumulation in
 * shared memory, one global add per CTA and entry; larger tables fall through to escgnn_embedding_bwd */
int escgnn_embedding_bwd_small(const float* d_dy, int lddy, const int64_t* d_idx, int idx_cols, const int64_t* d_col_offsets,
                               const int* d_rows, int rows_cap, int channels, int table_rows, float* d_dtable, void* stream);
/* loss + its gradient: kind 0 = L1Loss mean (run_graphcount.py:498, run_zinc.py:283); kind 1 = BCEWithLogitsLoss
 * over labelled targets y == y (run_ogb_mol.py:58-74) */
int escgnn_loss_fwd_bwd(const float* d_pred, int ldp, const float* d_target, int kind, const int* d_rows, int rows_cap,
                        int n_targets, float* d_loss, float* d_dpred, int lddp, void* stream);

/* K6 dense contraction on tcgen05 tensor cores, 3xTF32 (fp32-grade accuracy): C[M,N] (+)= A * B^T (+ bias[n]).
 * Replaces nn.Linear forward / dgrad / wgrad (run_graphcount.py:54-118, zinc_models.py:513-566, GINEConv.lin).
 *   a_mn_major == 0: A is [M,K] row-major (lda);  != 0: A is stored [K,M] row-major (i.e. the caller holds A^T)
 *   b_mn_major == 0: B is [N,K] row-major (ldb);  != 0: B is stored [K,N] row-major
 * Operands are plain fp32; the low tf32 planes are produced on chip. lda / ldb multiples of 4 and 16-byte aligned bases
 * (TMA). Rows beyond the tensor extents read as zero. Split-K (small output, long K: wgrad) uses d_workspace
 * (escgnn_gemm_workspace_floats(M,N,K) floats; NULL disables it) with an ordered reduction.
 * accumulate: 0 C = ..., 1 C += ... (ordered), 2 C += ... with the K-slices added by fp32 vector reductions
 * (red.global.add.v4.f32): no workspace, no reduction launch, summation order not fixed. */
int escgnn_gemm_tf32x3(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c, int ldc,
                       const float* d_bias, int M, int N, int K, int accumulate, float* d_workspace, int64_t workspace_floats,
                       void* stream);
/* Same product for the static-shape engine, where M (forward / dgrad) or K (wgrad) is a row CAPACITY and the actual row
 * count lives in device memory: rows_dim = 1 -> output tiles that start at or past *d_rows leave without touching C (their
 * rows of C keep their previous contents); rows_dim = 2 -> k-blocks past *d_rows are skipped. Keeps the launch shape static
 * (CUDA-graph capturable) while the work follows the batch; d_rows = NULL: identical to escgnn_gemm_tf32x3.
 * rows_dim | 4: *d_rows was final before the kernel PRECEDING this launch on the stream was launched (a static-shape engine writes
 * its counts once at the start of the step): the kernel then reads it ahead of its programmatic-dependency wait, which takes an
 * L2-missing load off the path between the wait and the first TMA issue. */
int escgnn_gemm_tf32x3_bounded(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c,
                               int ldc, const float* d_bias, int M, int N, int K, int accumulate, float* d_workspace,
                               int64_t workspace_floats, const int* d_rows, int rows_dim, void* stream);
int64_t escgnn_gemm_workspace_floats(int M, int N, int K);
/* Fused Linear -> BatchNorm1d(training) -> activation, ONE launch each way (replaces the Linear, BN, act triples of
 * nn.Sequential(Linear, Dropout, BN, act, ...): run_graphcount.py:54-61,78-87; zinc_models.py:513-522,538-566). The GEMM
 * (tcgen05 3xTF32, as escgnn_gemm_tf32x3_bounded) keeps its tile in tensor memory; column statistics are reduced per tile,
 * exchanged through d_ws behind a grid barrier and summed in a fixed order (deterministic).
 *   fwd: y = x W^T + b -> d_y (pre-BN, kept for the backward; may be NULL), d_out = act(BN(y)); saves d_mean / d_rstd, updates the
 *        running statistics like torch. Rows [*d_rows, rows_cap) of d_out are written as zeros.
 *   bwd: given d_dy = gradient wrt the output of the NEXT Linear (weights d_w [n_out, n_in]) whose input was act(BN(x)),
 *        computes d(act(BN(x))) = d_dy W on the tensor cores and, in the epilogue, the gradient wrt x (the saved pre-BN d_x):
 *        d_dx [rows_cap, n_in]; d_dgamma / d_dbeta [bn_cols]. Columns [bn_cols, n_in) of d_dx receive the plain product
 *        (concatenated inputs whose tail did not come through the BatchNorm).
 * The grid barrier needs the whole grid resident: escgnn_linear_bn_fusable() tells whether a shape qualifies on this device
 * (ESCGNN_ERR_TOO_LARGE from the launchers otherwise: use the separate entry points), and two such launches must not run
 * concurrently on different streams. d_ws: escgnn_linear_bn_workspace_floats() floats, zero before first use. */
int escgnn_linear_bn_fusable(int rows_cap, int n_out, int k_in);
/* CTAs of the fused kernel for this shape that the current device keeps resident at once (<= 0: CUDA error, negated) */
int escgnn_linear_bn_resident_ctas(int n_cols, int rows_cap, int backward);
int64_t escgnn_linear_bn_workspace_floats(int rows_cap, int n_cols);
int escgnn_linear_bn_act_fwd(const float* d_x, int ldx, const float* d_w, int ldw, const float* d_bias, int rows_cap, int n_out, int k_in,
                             const int* d_rows, const float* d_gamma, const float* d_beta, float* d_running_mean, float* d_running_var,
                             float* d_mean, float* d_rstd, int act, float eps, float momentum, float* d_y, int ldy, float* d_out,
                             int ldo, float* d_ws, int64_t ws_floats, void* stream);
int escgnn_linear_bn_act_bwd(const float* d_dy, int lddy, const float* d_w, int ldw, int rows_cap, int n_in, int n_out, const int* d_rows,
                             const float* d_x, int ldx, const float* d_mean, const float* d_rstd, const float* d_gamma,
                             const float* d_beta, int act, int bn_cols, float* d_dgamma, float* d_dbeta, float* d_dx, int lddx,
                             float* d_ws, int64_t ws_floats, void* stream);
/* shared-memory plan of the GEMM: -1 auto, 0 = 2 stages (2 CTAs/SM), 1 = 4 stages (1 CTA/SM); +2 = the variant that keeps
 * both planes of A in shared memory instead of tensor memory (experiments / tests) */
int escgnn_gemm_set_plan(int plan);
/* 128 x 256 output tiles (one CTA per SM, 3-stage ring, the whole tensor memory) for products with a 256-column output, K-major A and
 * 64..148 row tiles -- the edge-level Linear layers of a reference batch (zinc_models.py:513-522, GINEConv.lin dgrad): on by default,
 * 0 = always 128-wide tiles (A/B timing and tests).  Returns the previous setting. */
int escgnn_gemm_set_wide(int on);
/* Debugging aid (tools/trace_gemm.py): when d_stamps is not NULL every CTA of the following escgnn_gemm_tf32x3* launches writes
 * %globaltimer values (ns): 0 CTA start, 1 prologue done, 2 predecessor complete
 * (griddepcontrol.wait returned), 3 first operand stage landed, 4 first k-block split, 5 last MMA issued, 6 accumulator complete,
 * 7 epilogue stored -- escgnn_gemm_trace_slots() words per CTA, d_stamps[(tile_n * tiles_m + tile_m) * slots + i].  A library built
 * with -DESCGNN_TRACE_KB (72 slots) adds, for k-block i < 16: 8+4i stage free (TMA issued), 9+4i landed, 10+4i split, 11+4i MMAs
 * issued.  NULL switches the trace off (the default; one predictable branch per stamp). */
int escgnn_gemm_set_trace(unsigned long long* d_stamps);
int escgnn_gemm_trace_slots(void);
/* Warps that split the operand tiles into their tf32 planes during the main loop and drain the accumulator afterwards: 8 (default:
 * two per tensor-memory lane quarter; with the staged epilogue 1.072 against 1.116 ms per training step at batch 256) or 4 (A/B
 * timing and tests; the fused-BatchNorm epilogues always run with 4).  Returns the previous value. */
int escgnn_gemm_set_split_warps(int warps);
/* With 8 warps and the one-CTA-per-SM plan: 2 (default) = the two groups of four warps split ALTERNATE k-blocks (four lo buffers), so
 * two of the ~0.6 us split latency chains are in flight per 0.39 us of MMA time; 1 = both groups work on halves of the same k-block.
 * Returns the previous value. */
int escgnn_gemm_set_kb_groups(int groups);
/* Epilogue of the GEMM: 1 (default) = full output tiles are parked in the dead operand stages and written out row-contiguously
 * (512 bytes per warp instruction); 0 = every thread stores its accumulator row straight from registers (A/B timing and tests; tiles
 * that cross the N extent or an unaligned C always take this path).  Returns the previous setting. */
int escgnn_gemm_set_staged_store(int on);
/* fp32 accumulation OUTSIDE the tensor core (the "drain" kernel: per k-block the hi*hi products start a fresh TMEM accumulator that
 * four extra warps add into registers with round-to-nearest; the A planes are rounded, not truncated).  tcgen05.mma accumulates with
 * truncation: -5.9e-6 mean signed relative error at K = 256 for the plain kernels, -8e-8 (rms 1.0e-7; cuBLAS fp32: 2.3e-7) with the
 * drain kernel (tools/bench_linear_bn.py); on the model's own tensors 4.3e-7 against 2.4e-6 (tools/debug_linear_error.py).
 * 0 = off (default: the step is 4-5 % faster and losses / predictions are inside the 1e-4 parity tolerance either way), 1 = grids of
 * at most one wave, 2 = every product that is not split along K.  Also ESCGNN_GEMM_DRAIN in the environment.  Returns the previous mode. */
int escgnn_gemm_set_drain(int mode);
/* number of CTAs a split-K product (weight gradients) spreads over; default 296 = two per SM. Fewer, longer slices leave
 * room for kernels running concurrently on other streams. Returns the previous value. */
int escgnn_gemm_set_split_target(int ctas);
/* x - tf32_trunc(x): the low plane of the 3xTF32 split (diagnostics; the GEMM computes it on chip) */
int escgnn_tf32_split_lo(const float* d_x, int ldx, float* d_lo, int ldlo, int64_t rows, int cols, void* stream);
/* CUDA-core GEMM with the same contract (any strides): odd shapes (K or N = 10, 1) and the test reference; accumulate = 2
 * spreads a long K over up to 64 slices that add into C with atomics (weight gradients of the odd-shaped layers) */
int escgnn_gemm_simple(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c, int ldc,
                       const float* d_bias, int M, int N, int K, int accumulate, void* stream);

/* ---- data-parallel exchange over NVLink peer memory (the reference is single-GPU, run_zinc.py:266-289) ----
 * escgnn_p2p_alloc: cudaMalloc'd, zero-filled buffer + its 64-byte CUDA IPC handle; the peers open it with escgnn_p2p_open
 * (one process per GPU; peer access is enabled lazily).  escgnn_allreduce_adam: ONE exchange + optimiser step, graph-capturable:
 * reduce-scatter of the gradients read from the peers' HBM, Adam on the owned slice (same update rule as escgnn_adam_step_device;
 * moments are only maintained for the owned slice), all-gather of the new parameters by stores into every rank's buffer, with
 * the two cross-GPU barriers inside the kernel (epoch flags; a peer that never arrives sets the error word after 4 s instead of
 * hanging).  h_peer_* are HOST arrays of `world` device pointers (index = rank; entry `rank` is this rank's own buffer):
 * gradients [n], parameters [n], flags [escgnn_p2p_flag_words(world)] (zero-initialised; word 2*world+2 of a bucket's block
 * of 2*world+8 words != 0 after a timeout).
 * n must be a multiple of 4; every rank must call it the same number of times. */
int escgnn_p2p_alloc(int64_t bytes, void** d_ptr, unsigned char* handle64);
int escgnn_p2p_open(const unsigned char* handle64, void** d_ptr);
int escgnn_p2p_close(void* d_ptr);
int escgnn_p2p_free(void* d_ptr);
int64_t escgnn_p2p_flag_words(int world);
/* The same over the element range [begin, end) only (multiples of 4), with flag block `bucket` (0..3: buckets of one step may be in
 * flight concurrently on different streams).  tick != 0 on the FIRST launch of an optimiser step (it advances the step counter and
 * the bias corrections), 0 on the others, which must be ordered after it on the device.  The union of a step's ranges must cover
 * [0, n) exactly once. */
int escgnn_allreduce_adam_range(const float* const* h_peer_grads, float* const* h_peer_params, unsigned long long* const* h_peer_flags,
                                int rank, int world, int64_t begin, int64_t end, int bucket, int tick, float* d_exp_avg,
                                float* d_exp_avg_sq, float* d_hyper, long long* d_state, void* stream);
int escgnn_allreduce_adam(const float* const* h_peer_grads, float* const* h_peer_params, unsigned long long* const* h_peer_flags, int rank,
                          int world, int64_t n, float* d_exp_avg, float* d_exp_avg_sq, float* d_hyper, long long* d_state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESCGNN_B200_H */
