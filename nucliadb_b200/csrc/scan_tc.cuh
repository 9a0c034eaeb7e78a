// nidx_b200 — K1 (batched): exact-scan scores on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// A batch of queries against a block of stored vectors IS a dense GEMM (scores = Q · Vᵀ), the one place on the
// nidx_vector path where tensor cores apply (segment.rs:581-597 evaluated for many queries at once).  To stay
// inside the 1e-5 similarity tolerance with f32 inputs the product is computed as a 3xTF32 split:
//     x = x_hi + x_lo,  x_hi = x with the low 13 mantissa bits cleared (exact in TF32),  x_lo = x - x_hi (exact in f32)
//     q·v ≈ q_hi·v_hi + q_hi·v_lo + q_lo·v_hi            (the dropped q_lo·v_lo term is < 2^-22 relative)
// accumulated in f32 in tensor memory.  The tensor core TRUNCATES its f32 accumulator after every instruction
// (measured: a single accumulator drifts by ~288 ulp = 1.6e-5 at d = 768), so the sum is spread over four TMEM
// accumulators -- three take the hi·hi products round-robin, one takes the two small cross terms -- that are added
// in f32 by the epilogue: |Δ| vs the lane-blocked CUDA-core path stays a few 1e-6 for unit vectors.  This path is
// only used for large batches (ground truth, bulk re-scoring); small batches keep the bit-exact kernel.
//
// One CTA = one 128-query × TC_N-vector tile of the score matrix.
//   * operands go global -> registers -> split hi/lo -> shared memory in the canonical K-major no-swizzle UMMA
//     layout (8-row × 16-byte core matrices; cute/atom/mma_traits_sm100.hpp "make_umma_desc<Major::K>"):
//         byte offset(row r, 16-byte k-chunk c) = c * LBO + (r / 8) * 128 + (r % 8) * 16,   SBO = 128
//     LBO is padded by 16 bytes so that the 8 k-chunks a quarter-warp stores fall into different banks;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = TC_N, K = 8 per instruction),
//     three per k-step (hi·hi, hi·lo, lo·hi), accumulating into TMEM (128 lanes × TC_N columns of f32);
//     tcgen05.commit arrives on an mbarrier when the tensor core has consumed the stage;
//   * two shared-memory stages plus a register stage: the global loads of k-block i+1 are issued (8 x 16 bytes per
//     thread in flight) before the barrier and the MMAs of k-block i, its split + stores overlap those MMAs;
//   * epilogue: 4 warps tcgen05.ld their 32 TMEM lanes (lane = query row), apply the cosine normalisation and
//     write scores[q][v].
#pragma once
#include "common.cuh"

namespace nidx {

constexpr int TC_M = 128;        // queries per tile (UMMA M)
constexpr int TC_N = 128;        // vectors per tile (UMMA N)
constexpr int TC_KB = 32;        // floats per k-block (8 chunks of 16 bytes = 4 MMA k-steps of K = 8)
constexpr int TC_THREADS = 256;
constexpr int TC_CHUNKS = TC_KB / 4;
constexpr uint32_t TC_LBO_A = (TC_M / 8) * 128 + 16;   // bytes between k-chunks of the A (query) tile, padded
constexpr uint32_t TC_LBO_B = (TC_N / 8) * 128 + 16;
constexpr uint32_t TC_A_BYTES = TC_CHUNKS * TC_LBO_A;  // one of {hi, lo} of one stage
constexpr uint32_t TC_B_BYTES = TC_CHUNKS * TC_LBO_B;
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;
constexpr size_t TC_SMEM_BYTES = 2 * TC_STAGE_BYTES + 1024;

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version = 1 [46,48), layout NONE [61,64)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t tc_instr_desc() {
    // cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6), a/b_format TF32 = 2 [7,10) [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(addr), "r"(parity) : "memory");
}

// split x into the TF32-exact head and the f32 remainder
__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = __fsub_rn(x, hi);
}

// A [rows][TC_KB] block (row stride ld floats, rows beyond n_rows read as zero) moves in two steps so that the
// global loads of the next k-block are in flight while the tensor core works: tc_fetch -> registers (4 consecutive
// rows x 8 chunks per warp: 128-byte coalesced reads), tc_store -> split hi/lo -> shared (conflict-free thanks to
// the padded LBO).
template <int ROWS>
__device__ __forceinline__ void tc_fetch(const float* __restrict__ src, uint64_t row0, uint64_t n_rows, int ld, int k0, float4 (&regs)[ROWS * TC_CHUNKS / TC_THREADS]) {
#pragma unroll
    for (int u = 0; u < ROWS * TC_CHUNKS / TC_THREADS; ++u) {
        int i = u * TC_THREADS + threadIdx.x;
        int r = i / TC_CHUNKS, c = i % TC_CHUNKS;
        regs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n_rows) regs[u] = ldg_stream(reinterpret_cast<const float4*>(src + (row0 + r) * (size_t)ld + k0) + c);
    }
}
template <int ROWS, uint32_t LBO>
__device__ __forceinline__ void tc_store(const float4 (&regs)[ROWS * TC_CHUNKS / TC_THREADS], unsigned char* hi_tile, unsigned char* lo_tile) {
#pragma unroll
    for (int u = 0; u < ROWS * TC_CHUNKS / TC_THREADS; ++u) {
        int i = u * TC_THREADS + threadIdx.x;
        int r = i / TC_CHUNKS, c = i % TC_CHUNKS;
        float4 v = regs[u], h, l;
        tc_split(v.x, h.x, l.x); tc_split(v.y, h.y, l.y); tc_split(v.z, h.z, l.z); tc_split(v.w, h.w, l.w);
        uint32_t off = (uint32_t)c * LBO + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
        *reinterpret_cast<float4*>(hi_tile + off) = h;
        *reinterpret_cast<float4*>(lo_tile + off) = l;
    }
}

// grid: (vector tiles, query tiles).  queries: [nq][ld] zero padded; ld % TC_KB == 0.
__global__ void __launch_bounds__(TC_THREADS, 1) scan_scores_tc_kernel(VecDev V, const float* __restrict__ queries, const float* __restrict__ qnorms, int nq,
                                                                       float* __restrict__ scores) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char* smem = tc_smem;
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t v0 = (uint64_t)blockIdx.x * TC_N;
    int q0 = blockIdx.y * TC_M;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(4 * TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_d = tmem_base_s;

    const uint32_t idesc = tc_instr_desc();
    int n_kb = V.ld / TC_KB;
    uint32_t smem_u32 = (uint32_t)__cvta_generic_to_shared(smem);
    // register stages: k-blocks kb+1 and kb+2 are in flight while k-block kb is split, stored and multiplied
    float4 ra[2][TC_M * TC_CHUNKS / TC_THREADS], rb[2][TC_N * TC_CHUNKS / TC_THREADS];
    tc_fetch<TC_M>(queries, (uint64_t)q0, (uint64_t)nq, V.ld, 0, ra[0]);
    tc_fetch<TC_N>(V.vecs, v0, (uint64_t)V.n, V.ld, 0, rb[0]);
    if (n_kb > 1) {
        tc_fetch<TC_M>(queries, (uint64_t)q0, (uint64_t)nq, V.ld, TC_KB, ra[1]);
        tc_fetch<TC_N>(V.vecs, v0, (uint64_t)V.n, V.ld, TC_KB, rb[1]);
    }
#pragma unroll 2
    for (int kb = 0; kb < n_kb; ++kb) {
        int st = kb & 1;
        // the tensor core must be done with this stage's previous contents (k-block kb - 2)
        if (kb >= 2) mbar_wait(&mbar[st], ((kb >> 1) - 1) & 1);
        unsigned char* stage = smem + (size_t)st * TC_STAGE_BYTES;
        unsigned char *a_hi = stage, *a_lo = stage + TC_A_BYTES, *b_hi = stage + 2 * TC_A_BYTES, *b_lo = stage + 2 * TC_A_BYTES + TC_B_BYTES;
        if (st == 0) { tc_store<TC_M, TC_LBO_A>(ra[0], a_hi, a_lo); tc_store<TC_N, TC_LBO_B>(rb[0], b_hi, b_lo); }
        else         { tc_store<TC_M, TC_LBO_A>(ra[1], a_hi, a_lo); tc_store<TC_N, TC_LBO_B>(rb[1], b_hi, b_lo); }
        if (kb + 2 < n_kb) {   // refill the register stage just consumed: two iterations to land
            if (st == 0) { tc_fetch<TC_M>(queries, (uint64_t)q0, (uint64_t)nq, V.ld, (kb + 2) * TC_KB, ra[0]); tc_fetch<TC_N>(V.vecs, v0, (uint64_t)V.n, V.ld, (kb + 2) * TC_KB, rb[0]); }
            else         { tc_fetch<TC_M>(queries, (uint64_t)q0, (uint64_t)nq, V.ld, (kb + 2) * TC_KB, ra[1]); tc_fetch<TC_N>(V.vecs, v0, (uint64_t)V.n, V.ld, (kb + 2) * TC_KB, rb[1]); }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t sbase = smem_u32 + (uint32_t)st * TC_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < TC_KB / 8; ++ks) {   // K = 8 per instruction = 2 chunks
                uint32_t ka = (uint32_t)(2 * ks) * TC_LBO_A, kbo = (uint32_t)(2 * ks) * TC_LBO_B;
                uint64_t d_ahi = tc_smem_desc(sbase + ka, TC_LBO_A, 128), d_alo = tc_smem_desc(sbase + TC_A_BYTES + ka, TC_LBO_A, 128);
                uint64_t d_bhi = tc_smem_desc(sbase + 2 * TC_A_BYTES + kbo, TC_LBO_B, 128), d_blo = tc_smem_desc(sbase + 2 * TC_A_BYTES + TC_B_BYTES + kbo, TC_LBO_B, 128);
                int step = kb * (TC_KB / 8) + ks;
                uint32_t main_acc = tmem_d + (uint32_t)(step % 3) * TC_N, small_acc = tmem_d + 3u * TC_N;
                tc_mma(small_acc, d_alo, d_bhi, idesc, step != 0);
                tc_mma(small_acc, d_ahi, d_blo, idesc, 1);
                tc_mma(main_acc, d_ahi, d_bhi, idesc, step >= 3);
            }
            tc_commit(&mbar[st]);   // arrives when every MMA issued so far has finished reading shared memory / writing TMEM
        }
    }
    // all MMAs done: the last commit covers everything issued before it
    int last = n_kb - 1;
    mbar_wait(&mbar[last & 1], (last >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w reads TMEM lanes 32*(w%4).. (lane = query row of the tile) and half of the columns (w/4).
    // Cosine: ab * (1/|q|) * (1/|v|) with simsimd's edge cases (zero norms, ab == 0, clamp) -- two multiplies per
    // score instead of a division; 1/|v| is computed once per tile into shared memory (the stages are free now).
    float* inv_vn = reinterpret_cast<float*>(smem);
    if (V.sim == SIM_COSINE && threadIdx.x < TC_N) {
        uint64_t v = v0 + threadIdx.x;
        float vn = v < V.n ? __ldg(V.norms + v) : 0.0f;
        inv_vn[threadIdx.x] = vn > 0.0f ? __frcp_rn(vn) : 0.0f;
    }
    __syncthreads();
    {
        int qrow = q0 + (warp & 3) * 32 + lane;
        float qn = (V.sim == SIM_COSINE && qrow < nq) ? qnorms[qrow] : 0.0f;
        float inv_qn = qn > 0.0f ? __frcp_rn(qn) : 0.0f;
        const int n_main = n_kb * (TC_KB / 8) >= 3 ? 3 : n_kb * (TC_KB / 8);   // accumulators that received at least one MMA
        const int cbase = (warp >> 2) * (TC_N / 2);
        for (int c0 = cbase; c0 < cbase + TC_N / 2; c0 += 16) {
            uint32_t r[4][16];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * TC_N + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[g][0]), "=r"(r[g][1]), "=r"(r[g][2]), "=r"(r[g][3]), "=r"(r[g][4]), "=r"(r[g][5]), "=r"(r[g][6]), "=r"(r[g][7]),
                               "=r"(r[g][8]), "=r"(r[g][9]), "=r"(r[g][10]), "=r"(r[g][11]), "=r"(r[g][12]), "=r"(r[g][13]), "=r"(r[g][14]), "=r"(r[g][15])
                             : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (qrow < nq) {
                float* out = scores + (size_t)qrow * V.n + v0 + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint64_t v = v0 + c0 + j;
                    if (v < V.n) {
                        float m0 = __uint_as_float(r[0][j]), m1 = n_main > 1 ? __uint_as_float(r[1][j]) : 0.f, m2 = n_main > 2 ? __uint_as_float(r[2][j]) : 0.f;
                        float ab = __fadd_rn(__fadd_rn(__fadd_rn(m0, m1), m2), __uint_as_float(r[3][j]));
                        float sc = ab;
                        if (V.sim == SIM_COSINE) {
                            float ivn = inv_vn[c0 + j];
                            if (inv_qn == 0.0f && ivn == 0.0f) sc = 1.0f;          // both norms zero: distance 0
                            else if (ab == 0.0f) sc = 0.0f;                       // distance 1
                            else {
                                float dist = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(ab, inv_qn), ivn));
                                sc = __fsub_rn(1.0f, dist > 0.0f ? dist : 0.0f);
                            }
                        }
                        out[j] = sc;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(4 * TC_N) : "memory");
}

}  // namespace nidx
