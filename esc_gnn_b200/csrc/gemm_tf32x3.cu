// K6 dense contraction on the 5th-generation tensor cores: C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias), fp32 in / fp32 out,
// computed as 3xTF32 (a = a_hi + a_lo, b = b_hi + b_lo; a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulated in fp32 TMEM), which
// keeps the result within ~2^-19 of an fp32 GEMM -- the reference's Linear layers run in fp32 on its CPU path and the
// parity target is 1e-4 relative (BASELINE.json), which single-pass TF32 (~1e-3) would miss.
//
// Replaces every nn.Linear forward / dgrad / wgrad of the NestedGIN_eff step (run_graphcount.py:54-109,113-118;
// zinc_models.py:513-566; GINEConv.lin, gine_conv_layer.py:31): forward  Y = X W^T + b        (A = X,  B = W,  both K-major)
//                                                                 dgrad    dX = dY W            (A = dY K-major, B = W MN-major)
//                                                                 wgrad    dW = dY^T X          (A = dY MN-major, B = X MN-major, split-K)
// so no operand is ever transposed in memory.
//
// Shape: one CTA per 128 x BLOCK_N output tile (x split-K slice); 6 warps: TMA producer, MMA issuer (+TMEM allocator),
// 4 epilogue warps (one per TMEM lane quarter).  Operand tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into
// a shared-memory ring guarded by mbarriers; `tcgen05.mma.kind::tf32` (M=128, N=BLOCK_N<=128, K=8) is issued by one
// thread; the accumulator lives in TMEM and is read back with tcgen05.ld for the epilogue.  The "hi" operand is the raw
// fp32 tile (the tensor core ignores the 13 low mantissa bits); the "lo" plane x - tf32_trunc(x) (exact in fp32) is
// produced ON CHIP by the four epilogue warps while the tiles sit in shared memory, so only fp32 A and B are ever read.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/escgnn_b200.h"
#include "launch.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                 // fp32 elements = 128 bytes = one swizzle row
constexpr int kSlabBytes = kBlockK * 128;   // one 32(MN) x 32(K) fp32 slab of an MN-major operand
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | SW128 (2) <<61
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only layout tf32 MN-major operands may use,
// cutlass sm100_common.inl:92; it pairs with TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct Params {
    float* C; int ldc; const float* bias;
    int M, N, K;                  // problem (M = row capacity; TMA zero-fills beyond the tensor extents)
    int kb_per_split, kb_total;   // k-blocks of 32
    float* partial;               // split-K partial tiles [splits][M][N] (nullptr when splits == 1 or atomic)
    int accumulate;
    int atomic;                   // split-K slices add their tile into C with fp32 vector reductions (C holds the initial value)
    const int* rows_ptr;          // optional device-side row count (static-shape engine): rows_dim 1 = bounds M (tiles past it
    int rows_dim;                 // leave without touching C), 2 = bounds K (k-blocks past it are skipped)
    int rows_early;               // the count was final before the PRECEDING kernel was launched: read it ahead of griddepcontrol.wait
    unsigned long long* trace;    // debugging (escgnn_gemm_set_trace): 8 globaltimer stamps per CTA of the TS kernel, nullptr = off
    int stage_out;                // full tiles are written out row-contiguously through shared memory (escgnn_gemm_set_staged_store)
};

// 8 kernel-level stamps; a library built with -DESCGNN_TRACE_KB (ESCGNN_NVCC_FLAGS, see tools/trace_gemm.py) adds 4 per k-block for
// the first 16 k-blocks of the TS kernel: 8+4i stage free (TMA issued), 9+4i landed, 10+4i split, 11+4i MMAs issued
#ifdef ESCGNN_TRACE_KB
constexpr int kTraceSlots = 72;
#define ESC_TRACE_KB(cond, slot) do { if (cond) trace_stamp(p, slot); } while (0)
#else
constexpr int kTraceSlots = 8;
#define ESC_TRACE_KB(cond, slot) do { } while (0)
#endif
__device__ __forceinline__ void trace_stamp(const Params& p, int slot) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kTraceSlots + slot] = t;
    }
}

// 3xTF32 split: hi = the raw fp32 value (the tensor core reads its top 19 bits, i.e. truncates), lo = x - trunc_tf32(x), exact in
// fp32.  In the plain kernels the error floor is not the split but the tensor core's truncating accumulation (-5.9e-6 mean signed
// relative error at K = 256, tools/bench_linear_bn.py); rounding B in place cost 15 % of the main loop and was removed.  The
// drain kernel below fixes the accumulation and rounds what it can round for free.
__device__ __forceinline__ float rna_tf32(float x) {          // round to nearest (ties away) to the 10-bit tf32 mantissa
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& v, float4& l) {
    l = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                    v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
}

__device__ __forceinline__ void red_add4(float4* dst, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Fused epilogues of the "TS" kernel (EPI template parameter): the Linear -> BatchNorm(training) -> activation chains of the
// model (run_graphcount.py:54-61,78-87; zinc_models.py:513-522,538-566) as ONE launch each way.
//   EPI 1 (forward):  y = A B^T + bias is kept in tensor memory; every CTA writes y (saved for the backward), reduces its
//                     tile's column sums (sum y, sum y^2) with a shuffle-transposed warp reduction, publishes them, waits at a
//                     grid barrier of its column block, sums the tile partials in a fixed order (deterministic), and normalises
//                     + activates straight from tensor memory:  C = act(BN(y)).
//   EPI 2 (backward): the tile is dgrad's output d(act(BN(x))); with the saved pre-BN x, mean, rstd the epilogue forms
//                     dz = d * act'(BN(x)), reduces sum dz and sum dz*xhat the same way, and writes the gradient with respect
//                     to x:  C = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)); dgamma / dbeta from the sums.
//                     Columns >= bn_cols are stored as plain dgrad output (the edge-type columns of the concatenated z).
// The grid barrier needs every CTA of a column block resident at once: the launcher only takes this path when the whole grid
// fits the machine (cudaOccupancyMaxActiveBlocksPerMultiprocessor x SM count), and callers keep such launches on ONE stream.
struct BnParams {
    float* Y; int ldy;                         // EPI 1: pre-BN output (saved)
    const float* X; int ldx;                   // EPI 2: pre-BN input saved by the forward
    const float* gamma; const float* beta;
    float* running_mean; float* running_var;   // EPI 1
    float* mean; float* rstd;                  // EPI 1: written; EPI 2: read
    float* dgamma; float* dbeta;               // EPI 2
    float eps, momentum;
    int act, bn_cols;
    float* ws;                                 // [0,32) arrive tickets, [32,64) depart tickets (uint32, left zero), then [tile][2][ldp]
    int ldp;
};

__device__ __forceinline__ float epi_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : expm1f(v);   // (a Taylor / exp(v) - 1 hybrid was measured SLOWER than the library expm1f: tools/trace_bn.py)
    return v;
}
__device__ __forceinline__ float epi_act_grad(float v, int act) {
    if (act == 1) return v > 0.f ? 1.f : 0.f;
    if (act == 2) return v > 0.f ? 1.f : expf(v);
    return 1.f;
}

// v[j] of lane l = element (row l, column j) of a 32 x 32 block; returns in lane l the sum of column l over the 32 rows
// (31 shuffles: every step halves the columns a lane is responsible for; fixed order -> deterministic)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
    #pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
        #pragma unroll
        for (int j = 0; j < half; ++j) {
            const float send = up ? v[j] : v[j + half];
            const float keep = up ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one row chunk of 32 columns [nb, nb + 32) of a row-major buffer: 16-byte accesses when possible
__device__ __forceinline__ void row_store32(float* rowp, int nb, int N, const float (&f)[32]) {
    if (nb + 32 <= N && ((reinterpret_cast<uintptr_t>(rowp + nb) & 15) == 0)) {
        #pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(rowp + nb + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
        #pragma unroll
        for (int j = 0; j < 32; ++j) if (nb + j < N) rowp[nb + j] = f[j];
    }
}
__device__ __forceinline__ void row_load32(const float* rowp, int nb, int N, float (&f)[32]) {
    if (nb + 32 <= N && ((reinterpret_cast<uintptr_t>(rowp + nb) & 15) == 0)) {
        #pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(rowp + nb + j));
            f[j] = v.x; f[j + 1] = v.y; f[j + 2] = v.z; f[j + 3] = v.w;
        }
    } else {
        #pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = nb + j < N ? __ldcg(rowp + nb + j) : 0.f;
    }
}

// grid barrier over the `expected` active CTAs of one column block (called by the 128 epilogue threads; t = 0..127).
// arrive: release (fence + atomic), spin: acquire; the last CTA to depart re-arms both tickets for the next launch.
__device__ __forceinline__ void column_block_barrier(float* ws, int block_y, unsigned expected, int t) {
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t == 0) {
        unsigned* arrive = reinterpret_cast<unsigned*>(ws) + block_y;
        unsigned* depart = arrive + 32;
        atomicAdd(arrive, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
            if (seen < expected) __nanosleep(40);
        } while (seen < expected);
        if (atomicAdd(depart, 1u) == expected - 1u) { *depart = 0u; __threadfence(); atomicExch(arrive, 0u); }
        __threadfence();
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM owner), 2..5 = tf32 splitters during the main loop, then epilogue.
// Shared memory: STAGES raw stages {A 16 KB, B BLOCK_N*128 B} filled by TMA, ONE lo buffer of the same shape written by the
// splitter warps (lo = x - trunc_tf32(x), element-wise, so the swizzled placement is simply preserved).  Two CTAs are
// resident per SM (<= 113 KB each): while one waits on TMA / split / epilogue the other keeps the tensor pipe busy.
template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS>
__global__ void __launch_bounds__(kThreads, LO_BUFS == 1 ? 2 : 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int kABytes = kBlockM * 128;           // 16 KB
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* lo_buf = smem + (size_t)STAGES * kStageBytes;
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], split_bar[LO_BUFS], lo_free_bar[LO_BUFS], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[BLOCK_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM, n0 = blockIdx.y * BLOCK_N;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : 128;

    // ---- prologue without global-memory traffic: overlaps the tail of the preceding kernel (launch.cuh)
    if (threadIdx.x == 0) {
        // (the descriptors live in the kernel's parameter space: fetching them is legal ahead of the dependency wait)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < LO_BUFS; ++b) { mbar_init(&split_bar[b], 128); mbar_init(&lo_free_bar[b], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    escgnn::pdl_wait();                              // operands (and the buffers written below) belong to earlier kernels until here
    escgnn::pdl_trigger();
    const int rows_now = p.rows_ptr ? *p.rows_ptr : 0;
    const bool skip = p.rows_dim == 1 && m0 >= rows_now;                      // tile past the actual row count: nothing to do
    const int kb_total = p.rows_dim == 2 ? min(p.kb_total, (rows_now + kBlockK - 1) / kBlockK) : p.kb_total;
    const int num_kb = skip ? 0 : max(min(kb_begin + p.kb_per_split, kb_total) - kb_begin, 0);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * kStageBytes;
                mbar_expect_tx(&full_bar[s], kStageBytes);
                const int k0 = (kb_begin + i) * kBlockK;
                if (!A_MN) tma_load_2d(st, &tmA, &full_bar[s], k0, m0);              // A stored [M, K]: box {32 k, 128 rows}
                else {                                                                 // A stored [K, M]: four slabs {32 m, 32 k}
                    #pragma unroll
                    for (int j = 0; j < kBlockM / 32; ++j) tma_load_2d(st + j * kSlabBytes, &tmA, &full_bar[s], m0 + 32 * j, k0);
                }
                uint8_t* sb = st + kABytes;
                if (!B_MN) tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
                else {
                    #pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * kSlabBytes, &tmB, &full_bar[s], n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, majors, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES, lb = i % LO_BUFS;
                const uint32_t a_lo = smem_u32(lo_buf + (size_t)lb * kStageBytes), b_lo = a_lo + kABytes;
                mbar_wait(&full_bar[s], (i / STAGES) & 1);       // raw tiles landed (hi operands)
                mbar_wait(&split_bar[lb], (i / LO_BUFS) & 1);    // lo planes of this k-block written
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes), b_hi = a_hi + kABytes;
                #pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t ab = pass == 2 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
                    #pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        // K-major (SW128): 8 rows x 128 B atoms, next 8-row group 1024 B further (SBO); a K step of 8 tf32 = +32 B.
                        // MN-major (SW128 / 32 B atoms): 32(MN) x 4(K) atoms of 512 B; next K group 512 B further (SBO), next
                        // MN atom one slab (4096 B) further (LBO); a K step of 8 = +1024 B.
                        const uint64_t ad = A_MN ? make_desc(ab + k * 1024, kSlabBytes, 512, 1) : make_desc(ab + k * 32, 0, 1024, 2);
                        const uint64_t bd = B_MN ? make_desc(bb + k * 1024, kSlabBytes, 512, 1) : make_desc(bb + k * 32, 0, 1024, 2);
                        mma_tf32(tmem_d, ad, bd, idesc, (i | pass | k) ? 1u : 0u);
                    }
                }
                tcgen05_commit(&empty_bar[s]);          // raw stage reusable once these MMAs have read it
                tcgen05_commit(&lo_free_bar[lb]);       // ... and so is the lo buffer
            }
            tcgen05_commit(&tmem_full_bar);             // accumulator complete
        }
    } else {
        // ===== splitters (main loop), then epilogue: 4 warps, TMEM lane quarter = warp % 4 =====
        const int t = threadIdx.x - 64;                 // 0..127
        for (int i = t; i < BLOCK_N; i += 128)          // bias tile for the epilogue (these four warps are its only readers)
            s_bias[i] = (p.bias && !p.partial && (!p.atomic || blockIdx.z == 0) && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % STAGES, lb = i % LO_BUFS;
            mbar_wait(&full_bar[s], (i / STAGES) & 1);
            mbar_wait(&lo_free_bar[lb], ((i / LO_BUFS) & 1) ^ 1);     // the MMAs that last read this lo buffer are done
            const float4* src = reinterpret_cast<const float4*>(smem + (size_t)s * kStageBytes);
            float4* dst = reinterpret_cast<float4*>(lo_buf + (size_t)lb * kStageBytes);
            #pragma unroll 4
            for (int q = t; q < kStageBytes / 16; q += 128) {
                float4 l;
                split4(src[q], l);
                dst[q] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to tcgen05.mma
            mbar_arrive(&split_bar[lb]);
        }
        const int q = warp & 3;
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_bias complete
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M && !skip;
        float* out = p.partial ? p.partial + ((size_t)blockIdx.z * p.M + row) * p.N : p.C + (size_t)row * p.ldc;
        #pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            if (num_kb > 0) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
                #pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (row_ok) {
                const int nb = n0 + c;
                const bool acc = p.accumulate && !p.partial;
                if (nb + 32 <= p.N && ((reinterpret_cast<uintptr_t>(out + nb) & 15) == 0)) {
                    #pragma unroll
                    for (int j = 0; j < 32; j += 4) {            // 128 contiguous bytes per thread, 16-byte stores
                        float4 r = make_float4(__uint_as_float(v[j]) + s_bias[c + j], __uint_as_float(v[j + 1]) + s_bias[c + j + 1],
                                               __uint_as_float(v[j + 2]) + s_bias[c + j + 2], __uint_as_float(v[j + 3]) + s_bias[c + j + 3]);
                        float4* dst = reinterpret_cast<float4*>(out + nb + j);
                        if (p.atomic) { red_add4(dst, r); continue; }
                        if (acc) { const float4 o = *dst; r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                        *dst = r;
                    }
                } else {
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < p.N) {
                            float r = __uint_as_float(v[j]) + s_bias[c + j];
                            if (p.atomic) { atomicAdd(out + n, r); continue; }
                            if (acc) r += out[n];
                            out[n] = r;
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// "TS" variant (default): A never returns to shared memory.  The four splitter warps own one
// row of the 128 x 32 A tile each (TMEM lane = row), read it from the swizzled stage, and store BOTH planes -- hi (raw
// fp32) and lo = x - trunc_tf32(x) -- into tensor memory with tcgen05.st; the MMA then takes A from TMEM
// (tcgen05.mma [d], [a], b_desc) for all three passes.  Only B's lo plane is still written to shared memory.  Per k-block
// that removes the A_lo write and three A reads from the shared-memory pipe (192 KB -> 128 KB of traffic for
// BLOCK_N = 128), which is what bounds the "SS" kernel above, and the freed space holds a second lo buffer.
// TMEM columns: [0, 128) accumulator, [128 + 64 b, 128 + 64 b + 32) A_hi and the next 32 A_lo of buffer b.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// SW = splitter / epilogue warps: 4 (one per tensor-memory lane quarter; the fused-BatchNorm epilogues are written for this shape) or
// 8 (two per quarter, each taking half of the A tile's 32 columns, half of the B_lo split and every other 32-column chunk of the
// epilogue).  In-kernel %globaltimer stamps (tools/trace_gemm.py) showed the 4-warp splitter, not the tensor pipe, pacing the main
// loop of a node-level product (0.68 us per k-block against 0.39 us of MMA time) and the epilogue taking 2.5 of its 9.5 us.
// KBG = 2 (needs SW = 8 and LO_BUFS = 4): the two groups of four warps split ALTERNATE k-blocks (each group a whole 128 x 32 A tile and
// B_lo plane, two lo buffers per group) instead of halves of the same one.  One split is a latency chain of ~0.6 us (stage wait,
// shared-memory reads, tcgen05.st + wait, proxy fence, barrier) against 0.39 us of MMA time per k-block: two of them in flight make
// the tensor pipe the pacer of a node-level product's main loop.
template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS, int EPI = 0, int SW = 4, int KBG = 1>
__global__ void __launch_bounds__(64 + 32 * SW, STAGES == 2 ? 2 : 1)
gemm_tf32x3_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p, const BnParams bn) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int kABytes = kBlockM * 128;           // 16 KB
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* lo_buf = smem + (size_t)STAGES * kStageBytes;          // LO_BUFS x B_lo
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], split_bar[LO_BUFS], lo_free_bar[LO_BUFS], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[BLOCK_N];
    // fused epilogues: their scratch lives in the first raw stage, which is dead once the accumulator is complete (every TMA fill
    // has been consumed by an MMA that has finished) -- the fused instantiations keep the shared-memory footprint, and with it
    // the CTAs per SM, of the plain kernel
    float (*s_part)[2][BLOCK_N] = reinterpret_cast<float (*)[2][BLOCK_N]>(smem);                        // [4] per-warp column sums
    float (*s_col)[BLOCK_N] = reinterpret_cast<float (*)[BLOCK_N]>(smem + 8 * BLOCK_N * sizeof(float));   // [4] per-column constants
    float (*s_fin)[BLOCK_N] = reinterpret_cast<float (*)[BLOCK_N]>(smem + 12 * BLOCK_N * sizeof(float));  // [2] EPI 2: mean(dz), mean(dz xhat)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM, n0 = blockIdx.y * BLOCK_N;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    if (threadIdx.x == 0) trace_stamp(p, 0);                                   // CTA start
    // accumulator [0, BLOCK_N) then the A planes; a 256-wide tile takes the whole tensor memory (one CTA per SM)
    constexpr uint32_t kTmemA = BLOCK_N <= 128 ? 128 : 256, kTmemCols = kTmemA + 64 * LO_BUFS <= 256 ? 256 : 512;
    static_assert(kTmemA + 64 * LO_BUFS <= 512, "tensor memory");
    static_assert(KBG == 1 || (KBG == 2 && SW == 8 && LO_BUFS % 2 == 0 && EPI == 0), "k-block groups");

    if (threadIdx.x == 0) {
        // (the descriptors live in the kernel's parameter space: fetching them is legal ahead of the dependency wait)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < LO_BUFS; ++b) { mbar_init(&split_bar[b], 32 * SW / KBG); mbar_init(&lo_free_bar[b], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    // the row count of a static-shape engine is written at the start of the step, many kernels ahead of this one: its (L2-missing)
    // read need not sit between the dependency wait and the first TMA issue
    int rows_now = (p.rows_ptr && p.rows_early) ? __ldg(p.rows_ptr) : 0;
    if (threadIdx.x == 0) trace_stamp(p, 1);                                   // prologue done (barriers, tensor memory)
    escgnn::pdl_wait();
    escgnn::pdl_trigger();
    if (threadIdx.x == 0) trace_stamp(p, 2);                                   // predecessor complete
    if (p.rows_ptr && !p.rows_early) rows_now = *p.rows_ptr;
    const bool skip = p.rows_dim == 1 && m0 >= rows_now;
    const int kb_total = p.rows_dim == 2 ? min(p.kb_total, (rows_now + kBlockK - 1) / kBlockK) : p.kb_total;
    const int num_kb = skip ? 0 : max(min(kb_begin + p.kb_per_split, kb_total) - kb_begin, 0);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                ESC_TRACE_KB(i < 16, 8 + 4 * i);                                // stage free: TMA for k-block i issued
                uint8_t* st = smem + (size_t)s * kStageBytes;
                mbar_expect_tx(&full_bar[s], kStageBytes);
                const int k0 = (kb_begin + i) * kBlockK;
                if (!A_MN) tma_load_2d(st, &tmA, &full_bar[s], k0, m0);
                else {
                    #pragma unroll
                    for (int j = 0; j < kBlockM / 32; ++j) tma_load_2d(st + j * kSlabBytes, &tmA, &full_bar[s], m0 + 32 * j, k0);
                }
                uint8_t* sb = st + kABytes;
                if (!B_MN) tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
                else {
                    #pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * kSlabBytes, &tmB, &full_bar[s], n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // A comes from tensor memory, where it is K-major by construction (the splitter transposes an MN-major tile on the fly)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES, lb = i % LO_BUFS;
                mbar_wait(&split_bar[lb], (i / LO_BUFS) & 1);    // A planes in TMEM, B_lo in shared memory (implies the stage landed)
                if (i == 0) trace_stamp(p, 4);                                // first k-block split
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = smem_u32(smem + (size_t)s * kStageBytes) + kABytes;
                const uint32_t b_lo = smem_u32(lo_buf + (size_t)lb * kBBytes);
                const uint32_t a_hi = tmem_d + kTmemA + (uint32_t)lb * 64u, a_lo = a_hi + 32u;
                #pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t at = pass == 2 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
                    #pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        const uint64_t bd = B_MN ? make_desc(bb + k * 1024, kSlabBytes, 512, 1) : make_desc(bb + k * 32, 0, 1024, 2);
                        mma_tf32_ts(tmem_d, at + (uint32_t)(k * 8), bd, idesc, (i | pass | k) ? 1u : 0u);
                    }
                }
                tcgen05_commit(&empty_bar[s]);
                tcgen05_commit(&lo_free_bar[lb]);
                ESC_TRACE_KB(i < 16, 11 + 4 * i);                               // MMAs of k-block i issued
            }
            tcgen05_commit(&tmem_full_bar);
            trace_stamp(p, 5);                                                // last MMA issued
        }
    } else {
        static_assert(SW == 4 || (SW == 8 && EPI == 0), "the fused epilogues are written for four warps");
        const int t = threadIdx.x - 64, q = warp & 3, r = q * 32 + lane;       // r: this thread's row of the A tile = its TMEM lane
        constexpr int kSplitThreads = 32 * SW / KBG;                           // threads working on one k-block
        constexpr int kACols = 4096 / kSplitThreads;                           // columns of the A tile per thread: 32 or 16
        const int half = SW == 8 ? (warp - 2) >> 2 : 0;                        // epilogue: which 32-column chunks this warp takes
        const int grp = KBG == 2 ? half : 0;                                   // main loop: which k-blocks (i % KBG == grp) ...
        const int ahalf = KBG == 2 ? 0 : half;                                 // ... or which half of every A tile's columns
        const int tg = KBG == 2 ? (t & 127) : t;                               // thread index within the group splitting a k-block
        for (int i = t; i < BLOCK_N; i += 32 * SW)
            s_bias[i] = (p.bias && !p.partial && (!p.atomic || blockIdx.z == 0) && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        for (int i = grp; i < num_kb; i += KBG) {
            const int s = i % STAGES, lb = i % LO_BUFS;
            mbar_wait(&full_bar[s], (i / STAGES) & 1);
            if (i == 0 && t == 0) trace_stamp(p, 3);                            // first stage landed
            ESC_TRACE_KB(i < 16 && tg == 0, 9 + 4 * i);                         // k-block i landed
            mbar_wait(&lo_free_bar[lb], ((i / LO_BUFS) & 1) ^ 1);     // MMAs that read TMEM A buffer / B_lo buffer `lb` are done
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint8_t* stage = smem + (size_t)s * kStageBytes;
            uint32_t hi[kACols], lo[kACols];
            if (!A_MN) {
                #pragma unroll
                for (int cc = 0; cc < kACols / 4; ++cc) {  // 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
                    const int c = ahalf * (kACols / 4) + cc;
                    const float4 v = *reinterpret_cast<const float4*>(stage + r * 128 + ((c ^ (r & 7)) << 4));
                    hi[4 * cc + 0] = __float_as_uint(v.x); hi[4 * cc + 1] = __float_as_uint(v.y);
                    hi[4 * cc + 2] = __float_as_uint(v.z); hi[4 * cc + 3] = __float_as_uint(v.w);
                }
            } else {
                // MN-major tile: slab q holds rows 32q..32q+31 as [k][32 m] lines of 128 B whose 32-byte units are XOR-ed with
                // (k & 3) (SWIZZLE_128B_ATOM_32B); a warp reads one permuted line per k -- conflict-free, and transposed for free
                const uint8_t* slab = stage + q * kSlabBytes + (lane & 7) * 4;
                #pragma unroll
                for (int kk = 0; kk < kACols; ++kk) {
                    const int k = ahalf * kACols + kk;
                    hi[kk] = *reinterpret_cast<const uint32_t*>(slab + k * 128 + ((((lane >> 3) ^ (k & 3))) << 5));
                }
            }
            #pragma unroll
            for (int k = 0; k < kACols; ++k) {
                const float x = __uint_as_float(hi[k]);
                lo[k] = __float_as_uint(x - __uint_as_float(hi[k] & 0xffffe000u));
            }
            const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) + kTmemA + (uint32_t)lb * 64u + (uint32_t)(ahalf * kACols);
            if constexpr (kACols == 32) { tmem_st32(ta, hi); tmem_st32(ta + 32u, lo); }
            else { tmem_st16(ta, hi); tmem_st16(ta + 32u, lo); }
            const float4* src = reinterpret_cast<const float4*>(stage + kABytes);
            float4* dst = reinterpret_cast<float4*>(lo_buf + (size_t)lb * kBBytes);
            #pragma unroll 4
            for (int e = tg; e < kBBytes / 16; e += kSplitThreads) {
                float4 l;
                split4(src[e], l);
                dst[e] = l;
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&split_bar[lb]);
            ESC_TRACE_KB(i < 16 && tg == 0, 10 + 4 * i);                        // k-block i split
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * SW) : "memory");
        const int row = m0 + r;
        if constexpr (EPI != 0) {
            const int rows_eff = p.rows_ptr ? min(rows_now, p.M) : p.M;
            float* crow = p.C + (size_t)row * p.ldc;
            if (m0 >= rows_eff) {
                // tile past the actual row count: its rows of C are kept zero (inert in the GEMMs that follow), no barrier
                if (row < p.M) {
                    float z[32];
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) z[j] = 0.f;
                    for (int c = 0; c < BLOCK_N; c += 32) row_store32(crow, n0 + c, p.N, z);
                }
            } else {
                const bool row_in = row < rows_eff;
                const unsigned expected = (unsigned)((rows_eff + kBlockM - 1) / kBlockM);
                const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
                float* part = bn.ws + 64;
                mbar_wait(&tmem_full_bar, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (EPI == 2) {
                    for (int i = t; i < BLOCK_N; i += 128) {
                        const int n = n0 + i;
                        const bool on = n < bn.bn_cols;
                        s_col[0][i] = on ? bn.mean[n] : 0.f; s_col[1][i] = on ? bn.rstd[n] : 0.f;
                        s_col[2][i] = on ? (bn.gamma ? bn.gamma[n] : 1.f) : 0.f; s_col[3][i] = on ? (bn.beta ? bn.beta[n] : 0.f) : 0.f;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                // ---- pass 1: per-tile column statistics.  EPI 1: tile mean, then the CENTRED sum of squares about it (a second read
                // of tensor memory; no cancellation), combined across tiles with Chan's update; also stores the pre-BN output y.
                // EPI 2: sum dz and sum dz * xhat.
                #pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(trow + (uint32_t)c, v);
                    float a[32];
                    if (EPI == 1) {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) a[j] = row_in ? __uint_as_float(v[j]) + s_bias[c + j] : 0.f;
                        if (row_in && bn.Y) row_store32(bn.Y + (size_t)row * bn.ldy, n0 + c, p.N, a);
                        s_part[q][0][c + lane] = warp_transpose_sum(a, lane);
                    } else {
                        float b[32];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) b[j] = 0.f;
                        if (row_in) row_load32(bn.X + (size_t)row * bn.ldx, n0 + c, min(p.N, bn.bn_cols), b);
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float xh = (b[j] - s_col[0][c + j]) * s_col[1][c + j];
                            const float dz = __uint_as_float(v[j]) * epi_act_grad(xh * s_col[2][c + j] + s_col[3][c + j], bn.act);
                            a[j] = (row_in && s_col[1][c + j] != 0.f) ? dz : 0.f;      // columns >= bn_cols (rstd slot 0) take no part
                            b[j] = a[j] * xh;
                        }
                        const float sa = warp_transpose_sum(a, lane), sb = warp_transpose_sum(b, lane);
                        s_part[q][0][c + lane] = sa; s_part[q][1][c + lane] = sb;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (EPI == 1) {
                    const float n_t = (float)min(kBlockM, rows_eff - m0);
                    for (int col = t; col < BLOCK_N; col += 128)
                        s_col[2][col] = ((s_part[0][0][col] + s_part[1][0][col]) + (s_part[2][0][col] + s_part[3][0][col])) / n_t;     // tile mean
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    #pragma unroll 1
                    for (int c = 0; c < BLOCK_N; c += 32) {
                        uint32_t v[32];
                        tmem_ld32(trow + (uint32_t)c, v);
                        float b[32];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = row_in ? (__uint_as_float(v[j]) + s_bias[c + j]) - s_col[2][c + j] : 0.f;
                            b[j] = d * d;
                        }
                        s_part[q][1][c + lane] = warp_transpose_sum(b, lane);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                for (int i = t; i < 2 * BLOCK_N; i += 128) {
                    const int which = i / BLOCK_N, col = i % BLOCK_N;
                    part[((size_t)blockIdx.x * 2 + which) * bn.ldp + n0 + col] = (EPI == 1 && which == 0) ? s_col[2][col] :
                        (s_part[0][which][col] + s_part[1][which][col]) + (s_part[2][which][col] + s_part[3][which][col]);
                }
                column_block_barrier(bn.ws, blockIdx.y, expected, t);
                // ---- every CTA combines the tile partials of its columns in a fixed order (deterministic): warp q takes tiles
                // q, q+4, ... with a lane per 4 columns (independent 16-byte loads), then the four warps' results are merged
                {
                    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), w = u;         // EPI 1: running mean / M2 (Chan); EPI 2: the two sums
                    float cnt = 0.f;
                    if (lane * 4 < BLOCK_N) {
                        const float* base = part + n0 + lane * 4;
                        #pragma unroll 4
                        for (unsigned tl = q; tl < expected; tl += 4) {
                            const float4 pa = __ldcg(reinterpret_cast<const float4*>(base + ((size_t)tl * 2 + 0) * bn.ldp));
                            const float4 pb = __ldcg(reinterpret_cast<const float4*>(base + ((size_t)tl * 2 + 1) * bn.ldp));
                            if (EPI == 1) {
                                const float nb = (float)min(kBlockM, rows_eff - (int)tl * kBlockM), nab = cnt + nb, f = nb / nab, g2 = cnt * f;
                                float d;
                                d = pa.x - u.x; u.x += d * f; w.x += pb.x + d * d * g2;
                                d = pa.y - u.y; u.y += d * f; w.y += pb.y + d * d * g2;
                                d = pa.z - u.z; u.z += d * f; w.z += pb.z + d * d * g2;
                                d = pa.w - u.w; u.w += d * f; w.w += pb.w + d * d * g2;
                                cnt = nab;
                            } else {
                                u.x += pa.x; u.y += pa.y; u.z += pa.z; u.w += pa.w;
                                w.x += pb.x; w.y += pb.y; w.z += pb.z; w.w += pb.w;
                            }
                        }
                        *reinterpret_cast<float4*>(&s_part[q][0][lane * 4]) = u;
                        *reinterpret_cast<float4*>(&s_part[q][1][lane * 4]) = w;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int col = t; col < BLOCK_N; col += 128) {
                    const int n = n0 + col;
                    const float m = (float)max(rows_eff, 1);
                    if (EPI == 1) {
                        float mean = 0.f, m2 = 0.f, cnt = 0.f;
                        #pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            float nb = 0.f;                                     // rows behind warp w4's tiles
                            for (unsigned tl = w4; tl < expected; tl += 4) nb += (float)min(kBlockM, rows_eff - (int)tl * kBlockM);
                            if (nb > 0.f) {
                                const float nab = cnt + nb, f = nb / nab, d = s_part[w4][0][col] - mean;
                                mean += d * f; m2 += s_part[w4][1][col] + d * d * cnt * f; cnt = nab;
                            }
                        }
                        const float var = fmaxf(m2 / m, 0.f), rstd = rsqrtf(var + bn.eps);
                        const float g = (bn.gamma && n < p.N) ? bn.gamma[n] : 1.f, bt = (bn.beta && n < p.N) ? bn.beta[n] : 0.f;
                        s_col[0][col] = rstd * g; s_col[1][col] = bt; s_col[3][col] = mean;
                        if (blockIdx.x == 0 && n < p.N) {
                            bn.mean[n] = mean; bn.rstd[n] = rstd;
                            const float ub = rows_eff > 1 ? m / (m - 1.f) : 1.f;
                            bn.running_mean[n] = (1.f - bn.momentum) * bn.running_mean[n] + bn.momentum * mean;
                            bn.running_var[n] = (1.f - bn.momentum) * bn.running_var[n] + bn.momentum * var * ub;
                        }
                    } else {
                        const float s1 = (s_part[0][0][col] + s_part[1][0][col]) + (s_part[2][0][col] + s_part[3][0][col]);
                        const float s2 = (s_part[0][1][col] + s_part[1][1][col]) + (s_part[2][1][col] + s_part[3][1][col]);
                        s_fin[0][col] = s1 / m; s_fin[1][col] = s2 / m;
                        if (blockIdx.x == 0 && n < min(p.N, bn.bn_cols)) { if (bn.dgamma) bn.dgamma[n] = s2; if (bn.dbeta) bn.dbeta[n] = s1; }
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // ---- pass 2: normalise (+ activation) / BatchNorm backward straight from tensor memory
                #pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(trow + (uint32_t)c, v);
                    float o[32];
                    if (EPI == 1) {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j)
                            o[j] = row_in ? epi_act(((__uint_as_float(v[j]) + s_bias[c + j]) - s_col[3][c + j]) * s_col[0][c + j] + s_col[1][c + j], bn.act) : 0.f;
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = 0.f;
                        if (row_in) row_load32(bn.X + (size_t)row * bn.ldx, n0 + c, min(p.N, bn.bn_cols), o);
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float rs = s_col[1][c + j], g = s_col[2][c + j];
                            const float xh = (o[j] - s_col[0][c + j]) * rs;
                            const float d = __uint_as_float(v[j]);
                            const float dz = d * epi_act_grad(xh * g + s_col[3][c + j], bn.act);
                            const float bnv = g * rs * (dz - s_fin[0][c + j] - xh * s_fin[1][c + j]);
                            o[j] = row_in ? (rs != 0.f ? bnv : d) : 0.f;
                        }
                    }
                    if (row < p.M) row_store32(crow, n0 + c, p.N, o);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        } else {
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t == 0) trace_stamp(p, 6);                                          // accumulator complete
        // Full tiles leave through shared memory: a thread owns a ROW of the accumulator (its tensor-memory lane), so storing straight
        // from the registers makes every 16-byte store instruction touch 32 different rows -- 4096 half-sector writes per tile,
        // 2.6 of the 9.5 us of a node-level product (tools/trace_gemm.py).  The operand stages are dead by now (every fill was consumed
        // by an MMA that has completed): the tile is parked there with a 4-float row pad (conflict-free 16-byte stores), then written
        // out row-contiguously, 512 bytes per warp instruction.
        const int ldo = p.partial ? p.N : p.ldc;
        float* obase = p.partial ? p.partial + (size_t)blockIdx.z * p.M * p.N : p.C;
        const bool staged = p.stage_out && n0 + BLOCK_N <= p.N && (ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(obase + n0) & 15) == 0);
        if (staged) {
            constexpr int kLd = BLOCK_N + 4;
            float* stg = reinterpret_cast<float*>(smem);
            if (!skip) {
                #pragma unroll 1
                for (int c = 32 * half; c < BLOCK_N; c += 32 * (SW / 4)) {
                    uint32_t v[32];
                    if (num_kb > 0) {
                        tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0u;
                    }
                    float* dst = stg + r * kLd + c;
                    #pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) =
                            make_float4(__uint_as_float(v[j]) + s_bias[c + j], __uint_as_float(v[j + 1]) + s_bias[c + j + 1],
                                        __uint_as_float(v[j + 2]) + s_bias[c + j + 2], __uint_as_float(v[j + 3]) + s_bias[c + j + 3]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * SW) : "memory");
            if (!skip) {
                constexpr int kLanesPerRow = BLOCK_N / 4, kRowsPerPass = (32 * SW) / kLanesPerRow;
                const bool acc = p.accumulate && !p.partial;
                if (t < kRowsPerPass * kLanesPerRow) {
                    const int col = (t % kLanesPerRow) * 4;
                    #pragma unroll 4
                    for (int rr = t / kLanesPerRow; rr < kBlockM; rr += kRowsPerPass) {
                        const int grow = m0 + rr;
                        if (grow >= p.M) break;
                        float4 w = *reinterpret_cast<const float4*>(stg + rr * kLd + col);
                        float4* dst = reinterpret_cast<float4*>(obase + (size_t)grow * ldo + n0 + col);
                        if (p.atomic) { red_add4(dst, w); continue; }
                        if (acc) { const float4 o = *dst; w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                        *dst = w;
                    }
                }
            }
        } else {
            const bool row_ok = row < p.M && !skip;
            float* out = p.partial ? p.partial + ((size_t)blockIdx.z * p.M + row) * p.N : p.C + (size_t)row * p.ldc;
            #pragma unroll 1
            for (int c = 32 * half; c < BLOCK_N; c += 32 * (SW / 4)) {      // SW == 8: the two warps of a lane quarter alternate chunks

                uint32_t v[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
                if (num_kb > 0) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (row_ok) {
                    const int nb = n0 + c;
                    const bool acc = p.accumulate && !p.partial;
                    if (nb + 32 <= p.N && ((reinterpret_cast<uintptr_t>(out + nb) & 15) == 0)) {
                        #pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 w = make_float4(__uint_as_float(v[j]) + s_bias[c + j], __uint_as_float(v[j + 1]) + s_bias[c + j + 1],
                                                   __uint_as_float(v[j + 2]) + s_bias[c + j + 2], __uint_as_float(v[j + 3]) + s_bias[c + j + 3]);
                            float4* dst = reinterpret_cast<float4*>(out + nb + j);
                            if (p.atomic) { red_add4(dst, w); continue; }
                            if (acc) { const float4 o = *dst; w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                            *dst = w;
                        }
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = nb + j;
                            if (n < p.N) {
                                float w = __uint_as_float(v[j]) + s_bias[c + j];
                                if (p.atomic) { atomicAdd(out + n, w); continue; }
                                if (acc) w += out[n];
                                out[n] = w;
                            }
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_stamp(p, 7);                                   // epilogue stored
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// "Drain" variant of the TS kernel: fp32-GEMM accuracy from the tensor cores.
// tcgen05.mma adds into its fp32 accumulator with TRUNCATION (measured: -2.6e-6 mean signed relative error for K = 256 --
// 96 accumulating MMAs -- against +3e-10 for an FFMA GEMM, tools/bench_linear_bn.py), and the BatchNorm stacks of the model
// amplify that one-sided error into gradient errors ~10-40x those of the reference's fp32 CPU run.  The cure (Ootomo &
// Yokota, "Recovering single precision accuracy from Tensor Cores", 2022): never let the big products accumulate inside the
// tensor core for long.  Per k-block the four hi*hi MMAs start a FRESH accumulator (two alternate in tensor memory), which
// four extra "drainer" warps read back (tcgen05.ld) and add to register accumulators with round-to-nearest FADDs while the
// next k-block's MMAs run; the small hi*lo / lo*hi products (2^-11 of the result, so their truncation is harmless) accumulate
// in a third TMEM region over the whole K and are added once at the end.  The truncating chain on the large terms is 4 MMAs
// long instead of 3 K / 8.
// Tensor memory (512 columns, one CTA per SM): [0,128) D1a, [128,256) D1b, [256,384) D2, [384,512) A planes (2 x {hi 32, lo 32}).
// Warps: 0 TMA producer, 1 MMA issuer, 2-5 splitters, 6-9 drainers + epilogue (TMEM lane quarter = warp % 4).
constexpr int kDrainThreads = 320;

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kDrainThreads, 1)
gemm_tf32x3_ts_drain_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
    constexpr int STAGES = 4, LO_BUFS = 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int kABytes = kBlockM * 128;
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* lo_buf = smem + (size_t)STAGES * kStageBytes;
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], split_bar[LO_BUFS], lo_free_bar[LO_BUFS], d1_full_bar[2], d1_free_bar[2],
        tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[BLOCK_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBlockM, n0 = blockIdx.y * BLOCK_N;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    constexpr uint32_t kTmemCols = 512, kD1 = 0, kD2 = 256, kTmemA = 384;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < LO_BUFS; ++b) { mbar_init(&split_bar[b], 128); mbar_init(&lo_free_bar[b], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&d1_full_bar[b], 1); mbar_init(&d1_free_bar[b], 128); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    // the row count of a static-shape engine is written at the start of the step, many kernels ahead of this one: its (L2-missing)
    // read need not sit between the dependency wait and the first TMA issue
    int rows_now = (p.rows_ptr && p.rows_early) ? __ldg(p.rows_ptr) : 0;
    escgnn::pdl_wait();
    escgnn::pdl_trigger();
    if (p.rows_ptr && !p.rows_early) rows_now = *p.rows_ptr;
    const bool skip = p.rows_dim == 1 && m0 >= rows_now;
    const int kb_total = p.rows_dim == 2 ? min(p.kb_total, (rows_now + kBlockK - 1) / kBlockK) : p.kb_total;
    const int num_kb = skip ? 0 : max(min(kb_begin + p.kb_per_split, kb_total) - kb_begin, 0);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&empty_bar[s], ((i / STAGES) & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * kStageBytes;
                mbar_expect_tx(&full_bar[s], kStageBytes);
                const int k0 = (kb_begin + i) * kBlockK;
                if (!A_MN) tma_load_2d(st, &tmA, &full_bar[s], k0, m0);
                else {
                    #pragma unroll
                    for (int j = 0; j < kBlockM / 32; ++j) tma_load_2d(st + j * kSlabBytes, &tmA, &full_bar[s], m0 + 32 * j, k0);
                }
                uint8_t* sb = st + kABytes;
                if (!B_MN) tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
                else {
                    #pragma unroll
                    for (int j = 0; j < BLOCK_N / 32; ++j) tma_load_2d(sb + j * kSlabBytes, &tmB, &full_bar[s], n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % STAGES, lb = i % LO_BUFS, b = i & 1;
                mbar_wait(&split_bar[lb], (i / LO_BUFS) & 1);
                mbar_wait(&d1_free_bar[b], ((i >> 1) & 1) ^ 1);       // the drainers have read what this accumulator held two k-blocks ago
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_hi = smem_u32(smem + (size_t)s * kStageBytes) + kABytes;
                const uint32_t b_lo = smem_u32(lo_buf + (size_t)lb * kBBytes);
                const uint32_t a_hi = tmem_d + kTmemA + (uint32_t)lb * 64u, a_lo = a_hi + 32u;
                const uint32_t d1 = tmem_d + kD1 + (uint32_t)b * 128u, d2 = tmem_d + kD2;
                #pragma unroll
                for (int k = 0; k < kBlockK / 8; ++k) {               // hi * hi: a fresh accumulator per k-block
                    const uint64_t bd = B_MN ? make_desc(b_hi + k * 1024, kSlabBytes, 512, 1) : make_desc(b_hi + k * 32, 0, 1024, 2);
                    mma_tf32_ts(d1, a_hi + (uint32_t)(k * 8), bd, idesc, k ? 1u : 0u);
                }
                tcgen05_commit(&d1_full_bar[b]);
                #pragma unroll
                for (int pass = 1; pass < 3; ++pass) {                // hi * lo, lo * hi: one accumulator for the whole K
                    const uint32_t at = pass == 2 ? a_lo : a_hi, bb = pass == 1 ? b_lo : b_hi;
                    #pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k) {
                        const uint64_t bd = B_MN ? make_desc(bb + k * 1024, kSlabBytes, 512, 1) : make_desc(bb + k * 32, 0, 1024, 2);
                        mma_tf32_ts(d2, at + (uint32_t)(k * 8), bd, idesc, (i | (pass - 1) | k) ? 1u : 0u);
                    }
                }
                tcgen05_commit(&empty_bar[s]);
                tcgen05_commit(&lo_free_bar[lb]);
            }
            tcgen05_commit(&tmem_full_bar);
        }
    } else if (warp < 6) {
        // ===== splitters: A planes -> tensor memory, B_lo -> shared memory (as in gemm_tf32x3_ts_kernel) =====
        const int t = threadIdx.x - 64, q = warp & 3, r = q * 32 + lane;
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % STAGES, lb = i % LO_BUFS;
            mbar_wait(&full_bar[s], (i / STAGES) & 1);
            mbar_wait(&lo_free_bar[lb], ((i / LO_BUFS) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint8_t* stage = smem + (size_t)s * kStageBytes;
            uint32_t hi[32], lo[32];
            if (!A_MN) {
                #pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(stage + r * 128 + ((c ^ (r & 7)) << 4));
                    hi[4 * c + 0] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
                    hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
                }
            } else {
                const uint8_t* slab = stage + q * kSlabBytes + (lane & 7) * 4;
                #pragma unroll
                for (int k = 0; k < 32; ++k)
                    hi[k] = *reinterpret_cast<const uint32_t*>(slab + k * 128 + ((((lane >> 3) ^ (k & 3))) << 5));
            }
            // A planes are built in registers, so both can be ROUNDED for free: hi = rn_tf32(x), lo = rn_tf32(x - hi) (|lo| <= 2^-12 |x|,
            // either sign); B keeps hi = the raw value (truncated by the tensor core) with lo = rn_tf32(x - trunc(x)).  Every dropped
            // term (representation error of the lo planes, lo_a * lo_b) then has zero mean instead of pulling the result towards zero.
            #pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float x = __uint_as_float(hi[k]), h = rna_tf32(x);
                hi[k] = __float_as_uint(h);
                lo[k] = __float_as_uint(rna_tf32(x - h));
            }
            const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) + kTmemA + (uint32_t)lb * 64u;
            tmem_st32(ta, hi);
            tmem_st32(ta + 32u, lo);
            const float4* src = reinterpret_cast<const float4*>(stage + kABytes);
            float4* dst = reinterpret_cast<float4*>(lo_buf + (size_t)lb * kBBytes);
            #pragma unroll 4
            for (int e = t; e < kBBytes / 16; e += 128) {
                float4 l;
                split4(src[e], l);
                dst[e] = make_float4(rna_tf32(l.x), rna_tf32(l.y), rna_tf32(l.z), rna_tf32(l.w));
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&split_bar[lb]);
        }
    } else {
        // ===== drainers: fp32 accumulation outside the tensor core, then the epilogue =====
        const int t = threadIdx.x - 192, q = warp & 3, r = q * 32 + lane;
        for (int i = t; i < BLOCK_N; i += 128)
            s_bias[i] = (p.bias && !p.partial && (!p.atomic || blockIdx.z == 0) && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        float acc[BLOCK_N];
        #pragma unroll
        for (int j = 0; j < BLOCK_N; ++j) acc[j] = 0.f;
        const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);
        for (int i = 0; i < num_kb; ++i) {
            const int b = i & 1;
            mbar_wait(&d1_full_bar[b], (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            #pragma unroll
            for (int c = 0; c < BLOCK_N; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + kD1 + (uint32_t)b * 128u + (uint32_t)c, v);
                #pragma unroll
                for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&d1_free_bar[b]);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");          // s_bias complete
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + r;
        const bool row_ok = row < p.M && !skip;
        float* out = p.partial ? p.partial + ((size_t)blockIdx.z * p.M + row) * p.N : p.C + (size_t)row * p.ldc;
        #pragma unroll
        for (int c = 0; c < BLOCK_N; c += 32) {
            if (num_kb > 0) {
                uint32_t v[32];
                tmem_ld32(trow + kD2 + (uint32_t)c, v);
                #pragma unroll
                for (int j = 0; j < 32; ++j) acc[c + j] += __uint_as_float(v[j]);
            }
            if (row_ok) {
                const int nb = n0 + c;
                const bool accum = p.accumulate && !p.partial;
                if (nb + 32 <= p.N && ((reinterpret_cast<uintptr_t>(out + nb) & 15) == 0)) {
                    #pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 w = make_float4(acc[c + j] + s_bias[c + j], acc[c + j + 1] + s_bias[c + j + 1],
                                               acc[c + j + 2] + s_bias[c + j + 2], acc[c + j + 3] + s_bias[c + j + 3]);
                        float4* dst = reinterpret_cast<float4*>(out + nb + j);
                        if (p.atomic) { red_add4(dst, w); continue; }
                        if (accum) { const float4 o = *dst; w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w; }
                        *dst = w;
                    }
                } else {
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = nb + j;
                        if (n < p.N) {
                            float w = acc[c + j] + s_bias[c + j];
                            if (p.atomic) { atomicAdd(out + n, w); continue; }
                            if (accum) w += out[n];
                            out[n] = w;
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// split-K reduction: C = (accumulate ? C : 0) + bias + sum_s partial[s]   (fixed order -> deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C, int ldc,
                                     const float* __restrict__ bias, int accumulate) {
    escgnn::pdl_enter();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)M * N; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / N), c = (int)(i % N);
        float s = accumulate ? C[(size_t)r * ldc + c] : 0.f;
        if (bias) s += bias[c];
        for (int k = 0; k < splits; ++k) s += partial[(size_t)k * M * N + i];
        C[(size_t)r * ldc + c] = s;
    }
}

// lo plane of the 3xTF32 split: x - trunc_tf32(x), exact in fp32
__global__ void tf32_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t rows, int cols, int ldx, int ldlo) {
    escgnn::pdl_enter();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols; const int c = (int)(i % cols);
        const float v = x[r * ldx + c];
        lo[r * ldlo + c] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    }
}

// simple CUDA-core fallback / test reference: C[M,N] (+)= A[M,K] B[N,K]^T (+bias), arbitrary strides and majors
__global__ void __launch_bounds__(256)
gemm_simple_kernel(const float* __restrict__ A, int lda, int a_mn, const float* __restrict__ B, int ldb, int b_mn,
                   float* __restrict__ C, int ldc, const float* __restrict__ bias, int M, int N, int K, int accumulate) {
    escgnn::pdl_enter();
    __shared__ float sA[16][65], sB[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    // gridDim.z > 1: split-K, every slice adds its tile with atomics (accumulate == 2: C holds the initial value)
    const int k_per = ((K + (int)gridDim.z - 1) / (int)gridDim.z + 15) / 16 * 16;
    const int k_lo = blockIdx.z * k_per, k_hi = min(K, k_lo + k_per);
    for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int kk = i & 15, mm = i >> 4;
            const int m = m0 + mm, n = n0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < k_hi) ? (a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]) : 0.f;
            sB[kk][mm] = (n < N && k < k_hi) ? (b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]) : 0.f;
        }
        __syncthreads();
        #pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) {
                float r = acc[i][j] + ((bias && blockIdx.z == 0) ? bias[n] : 0.f);
                if (gridDim.z > 1) { atomicAdd(&C[(size_t)m * ldc + n], r); continue; }
                if (accumulate) r += C[(size_t)m * ldc + n];
                C[(size_t)m * ldc + n] = r;
            }
        }
}

// cuTensorMapEncodeTiled is a driver entry point: fetch it through the runtime so the library has no link-time
// dependency on libcuda (it must load on machines without a driver, e.g. for the symbol check of the CPU test suite)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor map, 128-byte swizzle; inner extent / stride in elements
int make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
             bool mn_major = false) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return 699;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

// two shared-memory plans: "dual" = 2 raw stages + 1 lo buffer (<= 96 KB, two CTAs per SM hide each other's latencies),
// "deep" = 4 raw stages + 2 lo buffers (<= 192 KB, one CTA per SM, split overlaps the MMAs of the previous k-block)
template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS>
int launch_cfg(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
    constexpr int stage = kBlockM * 128 + BLOCK_N * 128;
    const int smem = (STAGES + LO_BUFS) * stage + 1024;
    auto kern = gemm_tf32x3_kernel<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS>;
    static bool configured = false;          // once per instantiation (keeps stream capture free of attribute calls)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    escgnn::launch_pdl(kern, grid, kThreads, smem, st, a, b, p);
    return (int)cudaGetLastError();
}

int g_gemm_plan = -1;      // -1 auto, 0 dual, 1 deep (escgnn_gemm_set_plan; experiments / tests); +2: force the "SS" kernel for K-major A

// One-time setup of a TS instantiation: opt in to its dynamic shared memory; returns how many of its CTAs the device holds at once
// (<= 0: a CUDA error, negated) -- the bound the grid-barrier epilogues need.
int g_gemm_kb_groups = 2;     // escgnn_gemm_set_kb_groups: 2 = with 8 warps and one CTA per SM the warp groups take alternate k-blocks
int g_gemm_split_warps = 8;   // escgnn_gemm_set_split_warps: splitter / epilogue warps of the plain TS kernel (4 or 8)

template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS, int EPI, int SW = 4, int KBG = 1>
int ts_resident_ctas() {
    static int resident = 0;
    if (resident != 0) return resident;
    const int smem = STAGES * (kBlockM * 128 + BLOCK_N * 128) + LO_BUFS * BLOCK_N * 128 + 1024;
    auto kern = gemm_tf32x3_ts_kernel<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, EPI, SW, KBG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return -(int)e;
    int per_sm = 0, dev = 0, sms = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 64 + 32 * SW, smem)) != cudaSuccess) return -(int)e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return -(int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return -(int)e;
    if (getenv("ESCGNN_DEBUG_OCC")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        int smem_sm = 0, smem_blk = 0, resv = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&resv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
        fprintf(stderr, "[escgnn] ts<%d,%d,%d,%d,%d,%d> per_sm %d dyn %d static %d regs %d maxdyn %d | sm %d blk %d reserved %d\n", BLOCK_N, (int)A_MN,
                (int)B_MN, STAGES, LO_BUFS, EPI, per_sm, smem, (int)fa.sharedSizeBytes, fa.numRegs, fa.maxDynamicSharedSizeBytes, smem_sm, smem_blk, resv);
    }
    resident = (per_sm > 2 ? 2 : per_sm) * sms;          // 256 of the 512 tensor-memory columns per CTA
    if (resident <= 0) resident = -(int)cudaErrorLaunchOutOfResources;
    return resident;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS, int EPI = 0, int SW = 4, int KBG = 1>
int launch_cfg_ts_sw(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st, const BnParams& bn) {
    const int smem = STAGES * (kBlockM * 128 + BLOCK_N * 128) + LO_BUFS * BLOCK_N * 128 + 1024;
    auto kern = gemm_tf32x3_ts_kernel<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, EPI, SW, KBG>;
    const int resident = ts_resident_ctas<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, EPI, SW, KBG>();     // cached after the first call
    if (resident <= 0) return -resident;
    if (EPI != 0 && (int)(grid.x * grid.y * grid.z) > resident) return ESCGNN_ERR_TOO_LARGE;   // the grid barrier would deadlock
    escgnn::launch_pdl(kern, grid, 64 + 32 * SW, smem, st, a, b, p, bn);
    return (int)cudaGetLastError();
}

template <int BLOCK_N, bool A_MN, bool B_MN, int STAGES, int LO_BUFS, int EPI = 0>
int launch_cfg_ts(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st, const BnParams& bn = BnParams()) {
    if constexpr (EPI == 0) {
        // one CTA per SM (the deep ring): room for four lo buffers, the two warp groups split alternate k-blocks
        if constexpr (STAGES == 4 && LO_BUFS == 2 && BLOCK_N <= 128) {
            if (g_gemm_split_warps == 8 && g_gemm_kb_groups == 2)
                return launch_cfg_ts_sw<BLOCK_N, A_MN, B_MN, 4, 4, 0, 8, 2>(a, b, p, grid, st, bn);
        }
        if (g_gemm_split_warps == 8) return launch_cfg_ts_sw<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, 0, 8>(a, b, p, grid, st, bn);
    }
    return launch_cfg_ts_sw<BLOCK_N, A_MN, B_MN, STAGES, LO_BUFS, EPI, 4>(a, b, p, grid, st, bn);
}

template <int BLOCK_N, bool A_MN, bool B_MN>
int launch_cfg_drain(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
    const int smem = 4 * (kBlockM * 128 + BLOCK_N * 128) + 2 * BLOCK_N * 128 + 1024;
    auto kern = gemm_tf32x3_ts_drain_kernel<BLOCK_N, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    escgnn::launch_pdl(kern, grid, kDrainThreads, smem, st, a, b, p);
    return (int)cudaGetLastError();
}

template <int EPI>
int bn_resident(int block_n, bool deep) {
    constexpr bool B_MN = EPI == 2;
    switch (block_n) {
        case 32: return deep ? ts_resident_ctas<32, false, B_MN, 4, 2, EPI>() : ts_resident_ctas<32, false, B_MN, 2, 2, EPI>();
        case 64: return deep ? ts_resident_ctas<64, false, B_MN, 4, 2, EPI>() : ts_resident_ctas<64, false, B_MN, 2, 2, EPI>();
        case 96: return deep ? ts_resident_ctas<96, false, B_MN, 4, 2, EPI>() : ts_resident_ctas<96, false, B_MN, 2, 2, EPI>();
        default: return deep ? ts_resident_ctas<128, false, B_MN, 4, 2, EPI>() : ts_resident_ctas<128, false, B_MN, 2, 2, EPI>();
    }
}

// fused Linear + BatchNorm epilogues (EPI 1: forward, K-major A and B; EPI 2: dgrad, MN-major B)
template <int EPI>
int launch_bn(int block_n, const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st, const BnParams& bn) {
    const bool deep = (int)(grid.x * grid.y) <= 148;
    constexpr bool B_MN = EPI == 2;
    switch (block_n) {
        case 32: return deep ? launch_cfg_ts<32, false, B_MN, 4, 2, EPI>(a, b, p, grid, st, bn) : launch_cfg_ts<32, false, B_MN, 2, 2, EPI>(a, b, p, grid, st, bn);
        case 64: return deep ? launch_cfg_ts<64, false, B_MN, 4, 2, EPI>(a, b, p, grid, st, bn) : launch_cfg_ts<64, false, B_MN, 2, 2, EPI>(a, b, p, grid, st, bn);
        case 96: return deep ? launch_cfg_ts<96, false, B_MN, 4, 2, EPI>(a, b, p, grid, st, bn) : launch_cfg_ts<96, false, B_MN, 2, 2, EPI>(a, b, p, grid, st, bn);
        default: return deep ? launch_cfg_ts<128, false, B_MN, 4, 2, EPI>(a, b, p, grid, st, bn) : launch_cfg_ts<128, false, B_MN, 2, 2, EPI>(a, b, p, grid, st, bn);
    }
}

int g_gemm_drain = 0;      // fp32 accumulation outside the tensor core (escgnn_gemm_set_drain): 0 off, 1 one-wave grids, 2 every non-split product

template <int BLOCK_N, bool A_MN, bool B_MN>
int launch(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
    const int ctas = (int)(grid.x * grid.y * grid.z);
    const int plan = g_gemm_plan >= 0 ? (g_gemm_plan & 1) : -1;
    const bool deep = plan >= 0 ? plan == 1 : (ctas <= 148 && grid.z == 1);   // one wave at 1 CTA/SM: deeper ring
    // long accumulation chains (one CTA walks the whole K) go through the drain kernel; split-K slices are short chains already
    if (g_gemm_plan < 0 && grid.z == 1 && p.kb_total > 2 && (g_gemm_drain == 2 || (g_gemm_drain == 1 && ctas <= 148)))
        return launch_cfg_drain<BLOCK_N, A_MN, B_MN>(a, b, p, grid, st);
    if (!(g_gemm_plan >= 2))           // planes of A in tensor memory (2 stages + 2 lo buffers still fit two CTAs per SM)
        return deep ? launch_cfg_ts<BLOCK_N, A_MN, B_MN, 4, 2>(a, b, p, grid, st) : launch_cfg_ts<BLOCK_N, A_MN, B_MN, 2, 2>(a, b, p, grid, st);
    return deep ? launch_cfg<BLOCK_N, A_MN, B_MN, 4, 2>(a, b, p, grid, st) : launch_cfg<BLOCK_N, A_MN, B_MN, 2, 1>(a, b, p, grid, st);
}

// "wide" plan: one CTA per 128 x 256 tile (3 raw stages + 2 lo buffers = 208 KB, the whole tensor memory).  For edge-level products
// with a 256-column output (z_embedding's Linear, the projection dgrad with K = 1056) it replaces two 128-wide CTAs per SM that
// share one tensor pipe through a 2-stage ring: the A tile is fetched once, the ring is deep enough to cover the TMA latency, and
// one UTCHMMA covers N = 256.
int g_gemm_wide = 1;       // escgnn_gemm_set_wide
unsigned long long* g_gemm_trace = nullptr;      // escgnn_gemm_set_trace
int g_gemm_staged = 1;     // escgnn_gemm_set_staged_store

template <bool A_MN, bool B_MN>
int dispatch_n(int block_n, const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
    if (block_n == 256) {
        if constexpr (!A_MN) return launch_cfg_ts<256, false, B_MN, 3, 2>(a, b, p, grid, st);
        else return ESCGNN_ERR_BAD_ARG;
    }
    switch (block_n) {
        case 32: return launch<32, A_MN, B_MN>(a, b, p, grid, st);
        case 64: return launch<64, A_MN, B_MN>(a, b, p, grid, st);
        case 96: return launch<96, A_MN, B_MN>(a, b, p, grid, st);
        default: return launch<128, A_MN, B_MN>(a, b, p, grid, st);
    }
}

// tile width along N: a multiple of 32 (MN-major slabs are 32 wide), at most 128, as few equal tiles as possible
int pick_block_n(int N) {
    const int tiles = (N + 127) / 128;
    int bn = ((N + tiles - 1) / tiles + 31) / 32 * 32;
    return bn > 128 ? 128 : bn;
}

int g_split_target = 296;   // CTAs a split-K product aims for (escgnn_gemm_set_split_target)

int pick_splits(int tiles, int kb_total, int M, int N, int64_t workspace_floats) {
    if (tiles >= 96 || kb_total < 16) return 1;
    int splits = g_split_target / tiles;
    if (splits > kb_total / 4) splits = kb_total / 4;
    if ((int64_t)splits * M * N > workspace_floats) splits = (int)(workspace_floats / ((int64_t)M * N));
    return splits < 1 ? 1 : splits;
}

}  // namespace

extern "C" {

int escgnn_tf32_split_lo(const float* d_x, int ldx, float* d_lo, int ldlo, int64_t rows, int cols, void* stream) {
    const int64_t total = rows * cols;
    if (total <= 0) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    escgnn::launch_pdl(tf32_lo_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, d_x, d_lo, rows, cols, ldx, ldlo);
    return (int)cudaGetLastError();
}

int escgnn_gemm_simple(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c, int ldc,
                       const float* d_bias, int M, int N, int K, int accumulate, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    if (accumulate == 2) {               // long-K, small-output products (weight gradients): spread K over up to 64 slices
        int splits = K / 128;
        const int tiles = (int)(grid.x * grid.y);
        if (splits * tiles > 296) splits = 296 / tiles;
        grid.z = (unsigned)(splits < 1 ? 1 : splits > 64 ? 64 : splits);
    }
    escgnn::launch_pdl(gemm_simple_kernel, grid, 256, 0, (cudaStream_t)stream, d_a, lda, a_mn_major, d_b, ldb, b_mn_major, d_c, ldc, d_bias, M, N, K, accumulate);
    return (int)cudaGetLastError();
}

int escgnn_gemm_set_split_target(int ctas) {
    const int was = g_split_target;
    if (ctas > 0) g_split_target = ctas;
    return was;
}

int escgnn_gemm_set_plan(int plan) { g_gemm_plan = plan; return 0; }

int escgnn_gemm_set_trace(unsigned long long* d_stamps) { g_gemm_trace = d_stamps; return 0; }
int escgnn_gemm_trace_slots(void) { return kTraceSlots; }

int escgnn_gemm_set_kb_groups(int groups) {
    const int was = g_gemm_kb_groups;
    if (groups == 1 || groups == 2) g_gemm_kb_groups = groups;
    return was;
}

int escgnn_gemm_set_staged_store(int on) {
    const int was = g_gemm_staged;
    g_gemm_staged = on ? 1 : 0;
    return was;
}

int escgnn_gemm_set_split_warps(int warps) {
    const int was = g_gemm_split_warps;
    if (warps == 4 || warps == 8) g_gemm_split_warps = warps;
    return was;
}

int escgnn_gemm_set_wide(int on) {
    const int was = g_gemm_wide;
    g_gemm_wide = on ? 1 : 0;
    return was;
}

int escgnn_gemm_set_drain(int mode) {
    const int was = g_gemm_drain;
    if (mode >= 0 && mode <= 2) g_gemm_drain = mode;
    return was;
}

/* how many floats of split-K workspace a call with these sizes can use (0 = never splits) */
int64_t escgnn_gemm_workspace_floats(int M, int N, int K) {
    const int block_n = pick_block_n(N);
    const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + block_n - 1) / block_n);
    const int kb = (K + kBlockK - 1) / kBlockK;
    const int splits = pick_splits(tiles, kb, M, N, (int64_t)1 << 40);
    return splits > 1 ? (int64_t)splits * M * N : 0;
}

int escgnn_gemm_tf32x3(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c, int ldc,
                       const float* d_bias, int M, int N, int K, int accumulate, float* d_workspace, int64_t workspace_floats,
                       void* stream) {
    return escgnn_gemm_tf32x3_bounded(d_a, lda, a_mn_major, d_b, ldb, b_mn_major, d_c, ldc, d_bias, M, N, K, accumulate, d_workspace,
                                      workspace_floats, nullptr, 0, stream);
}

int escgnn_gemm_tf32x3_bounded(const float* d_a, int lda, int a_mn_major, const float* d_b, int ldb, int b_mn_major, float* d_c, int ldc,
                               const float* d_bias, int M, int N, int K, int accumulate, float* d_workspace, int64_t workspace_floats,
                               const int* d_rows, int rows_dim, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if ((lda & 3) || (ldb & 3) || ((uintptr_t)d_a & 15) || ((uintptr_t)d_b & 15))
        return ESCGNN_ERR_BAD_ARG;            // TMA needs 16-byte aligned bases and row pitches
    cudaStream_t st = (cudaStream_t)stream;
    int block_n = pick_block_n(N);
    const int tiles_m = (M + kBlockM - 1) / kBlockM;
    int n_tiles = (N + block_n - 1) / block_n;
    const int kb_total = (K + kBlockK - 1) / kBlockK;
    const bool atomic = accumulate == 2;       // C += A B^T with the K-slices added by fp32 reductions (order not fixed)
    // one wave of 128 x 256 tiles when the output is 256 columns wide and the 128-wide tiles would need two CTAs per SM (more than 74
    // row tiles; below that the 128-wide grid fits one CTA per SM with the deep ring and is faster: 12.9 vs 19.0 us at 65 row tiles,
    // against 22.5 vs 20.2 us at 95, tools/bench_gemm_wide.py)
    const bool wide = g_gemm_wide && g_gemm_plan < 0 && g_gemm_drain == 0 && !a_mn_major && !atomic && N % 256 == 0 &&
                      tiles_m * (N / 128) > 148 && tiles_m * (N / 256) <= 148;
    if (wide) { block_n = 256; n_tiles = N / 256; }
    int splits = wide ? 1 : atomic ? pick_splits(tiles_m * n_tiles, kb_total, M, N, (int64_t)1 << 40)
                                   : d_workspace ? pick_splits(tiles_m * n_tiles, kb_total, M, N, workspace_floats) : 1;
    Params p;
    p.C = d_c; p.ldc = ldc; p.bias = d_bias; p.M = M; p.N = N; p.K = K;
    p.kb_total = kb_total; p.kb_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.partial = (splits > 1 && !atomic) ? d_workspace : nullptr;
    p.accumulate = accumulate;
    p.atomic = atomic ? 1 : 0;
    p.rows_ptr = d_rows; p.rows_dim = d_rows ? (rows_dim & 3) : 0; p.rows_early = (d_rows && (rows_dim & 4)) ? 1 : 0;
    p.trace = g_gemm_trace; p.stage_out = g_gemm_staged;
    CUtensorMap a, b;
    int rc = 0;
    if (!a_mn_major) rc |= make_map(&a, d_a, K, M, lda, kBlockK, kBlockM);          // [M, K] row-major: inner = K
    else rc |= make_map(&a, d_a, M, K, lda, 32, kBlockK, true);                     // [K, M] row-major: inner = M
    if (!b_mn_major) rc |= make_map(&b, d_b, K, N, ldb, kBlockK, block_n);
    else rc |= make_map(&b, d_b, N, K, ldb, 32, kBlockK, true);
    if (rc) return rc;
    dim3 grid((unsigned)tiles_m, (unsigned)n_tiles, (unsigned)splits);
    if (!a_mn_major && !b_mn_major) rc = dispatch_n<false, false>(block_n, a, b, p, grid, st);
    else if (!a_mn_major && b_mn_major) rc = dispatch_n<false, true>(block_n, a, b, p, grid, st);
    else if (a_mn_major && !b_mn_major) rc = dispatch_n<true, false>(block_n, a, b, p, grid, st);
    else rc = dispatch_n<true, true>(block_n, a, b, p, grid, st);
    if (rc) return rc;
    if (splits > 1 && !atomic) {
        int64_t blocks = ((int64_t)M * N + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        escgnn::launch_pdl(splitk_reduce_kernel, (unsigned)blocks, 256, 0, st, d_workspace, splits, M, N, d_c, ldc, d_bias, accumulate);
        rc = (int)cudaGetLastError();
    }
    return rc;
}

/* ---- fused Linear -> BatchNorm(training) -> activation, one launch each way (header: escgnn_linear_bn_act_fwd / _bwd) ---- */
int escgnn_linear_bn_resident_ctas(int n_cols, int rows_cap, int backward) {
    const int block_n = pick_block_n(n_cols);
    const int ctas = ((rows_cap + kBlockM - 1) / kBlockM) * ((n_cols + block_n - 1) / block_n);
    return backward ? bn_resident<2>(block_n, ctas <= 148) : bn_resident<1>(block_n, ctas <= 148);
}

int escgnn_linear_bn_fusable(int rows_cap, int n_out, int k_in) {
    if (rows_cap <= 0 || n_out <= 0 || k_in <= 0 || (k_in & 3)) return 0;
    const int block_n = pick_block_n(n_out);
    const int ctas = ((rows_cap + kBlockM - 1) / kBlockM) * ((n_out + block_n - 1) / block_n);
    if ((n_out + block_n - 1) / block_n > 32) return 0;          // one arrive / depart ticket per column block
    // the whole grid must be resident at once (the same bound the launchers enforce), for the forward and the dgrad instantiation
    return (ctas <= bn_resident<1>(block_n, ctas <= 148) && ctas <= bn_resident<2>(block_n, ctas <= 148)) ? 1 : 0;
}

static int bn_ldp(int n_cols) {            // tile partials are written for every column of every column block
    const int block_n = pick_block_n(n_cols);
    return (n_cols + block_n - 1) / block_n * block_n;
}

int64_t escgnn_linear_bn_workspace_floats(int rows_cap, int n_cols) {
    return 64 + (int64_t)((rows_cap + kBlockM - 1) / kBlockM) * 2 * bn_ldp(n_cols);
}

int escgnn_linear_bn_act_fwd(const float* d_x, int ldx, const float* d_w, int ldw, const float* d_bias, int rows_cap, int n_out, int k_in,
                             const int* d_rows, const float* d_gamma, const float* d_beta, float* d_running_mean, float* d_running_var,
                             float* d_mean, float* d_rstd, int act, float eps, float momentum, float* d_y, int ldy, float* d_out,
                             int ldo, float* d_ws, int64_t ws_floats, void* stream) {
    if (rows_cap <= 0 || n_out <= 0 || k_in <= 0) return 0;
    if ((ldx & 3) || (ldw & 3) || ((uintptr_t)d_x & 15) || ((uintptr_t)d_w & 15) || !d_ws || !d_mean || !d_rstd || !d_running_mean ||
        !d_running_var || !d_out)
        return ESCGNN_ERR_BAD_ARG;
    if (ws_floats < escgnn_linear_bn_workspace_floats(rows_cap, n_out)) return ESCGNN_ERR_CAPACITY;
    const int block_n = pick_block_n(n_out);
    const int tiles_m = (rows_cap + kBlockM - 1) / kBlockM, n_tiles = (n_out + block_n - 1) / block_n;
    if (n_tiles > 32) return ESCGNN_ERR_TOO_LARGE;
    Params p;
    p.C = d_out; p.ldc = ldo; p.bias = d_bias; p.M = rows_cap; p.N = n_out; p.K = k_in;
    p.kb_total = (k_in + kBlockK - 1) / kBlockK; p.kb_per_split = p.kb_total;
    p.partial = nullptr; p.accumulate = 0; p.atomic = 0; p.rows_ptr = d_rows; p.rows_dim = d_rows ? 1 : 0; p.rows_early = 0; p.trace = nullptr; p.stage_out = 0;
    BnParams bn = BnParams();
    bn.Y = d_y; bn.ldy = ldy; bn.gamma = d_gamma; bn.beta = d_beta; bn.running_mean = d_running_mean; bn.running_var = d_running_var;
    bn.mean = d_mean; bn.rstd = d_rstd; bn.eps = eps; bn.momentum = momentum; bn.act = act; bn.bn_cols = n_out; bn.ws = d_ws;
    bn.ldp = bn_ldp(n_out);
    CUtensorMap a, b;
    int rc = make_map(&a, d_x, k_in, rows_cap, ldx, kBlockK, kBlockM) | make_map(&b, d_w, k_in, n_out, ldw, kBlockK, block_n);
    if (rc) return rc;
    return launch_bn<1>(block_n, a, b, p, dim3((unsigned)tiles_m, (unsigned)n_tiles, 1), (cudaStream_t)stream, bn);
}

int escgnn_linear_bn_act_bwd(const float* d_dy, int lddy, const float* d_w, int ldw, int rows_cap, int n_in, int n_out, const int* d_rows,
                             const float* d_x, int ldx, const float* d_mean, const float* d_rstd, const float* d_gamma,
                             const float* d_beta, int act, int bn_cols, float* d_dgamma, float* d_dbeta, float* d_dx, int lddx,
                             float* d_ws, int64_t ws_floats, void* stream) {
    if (rows_cap <= 0 || n_out <= 0 || n_in <= 0) return 0;
    if ((lddy & 3) || (ldw & 3) || ((uintptr_t)d_dy & 15) || ((uintptr_t)d_w & 15) || !d_ws || !d_mean || !d_rstd || !d_x || !d_dx ||
        bn_cols < 0 || bn_cols > n_in)
        return ESCGNN_ERR_BAD_ARG;
    if (ws_floats < escgnn_linear_bn_workspace_floats(rows_cap, n_in)) return ESCGNN_ERR_CAPACITY;
    // dX[rows, n_in] = dY[rows, n_out] W[n_out, n_in]: output width n_in, contraction over n_out; W is the MN-major B operand
    const int block_n = pick_block_n(n_in);
    const int tiles_m = (rows_cap + kBlockM - 1) / kBlockM, n_tiles = (n_in + block_n - 1) / block_n;
    if (n_tiles > 32) return ESCGNN_ERR_TOO_LARGE;
    Params p;
    p.C = d_dx; p.ldc = lddx; p.bias = nullptr; p.M = rows_cap; p.N = n_in; p.K = n_out;
    p.kb_total = (n_out + kBlockK - 1) / kBlockK; p.kb_per_split = p.kb_total;
    p.partial = nullptr; p.accumulate = 0; p.atomic = 0; p.rows_ptr = d_rows; p.rows_dim = d_rows ? 1 : 0; p.rows_early = 0; p.trace = nullptr; p.stage_out = 0;
    BnParams bn = BnParams();
    bn.X = d_x; bn.ldx = ldx; bn.gamma = d_gamma; bn.beta = d_beta; bn.mean = const_cast<float*>(d_mean); bn.rstd = const_cast<float*>(d_rstd);
    bn.dgamma = d_dgamma; bn.dbeta = d_dbeta; bn.act = act; bn.bn_cols = bn_cols; bn.ws = d_ws; bn.ldp = bn_ldp(n_in);
    CUtensorMap a, b;
    int rc = make_map(&a, d_dy, n_out, rows_cap, lddy, kBlockK, kBlockM) | make_map(&b, d_w, n_in, n_out, ldw, 32, kBlockK, true);
    if (rc) return rc;
    return launch_bn<2>(block_n, a, b, p, dim3((unsigned)tiles_m, (unsigned)n_tiles, 1), (cudaStream_t)stream, bn);
}

}  // extern "C"
