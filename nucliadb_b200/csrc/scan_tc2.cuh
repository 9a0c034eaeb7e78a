// nidx_b200 — K1 (batched): the exact scan for a large query batch as FILTER (tensor cores) + REFINE (bit-exact).
//
// brute_force_search (nidx/nidx_vector/src/segment.rs:569-623) evaluated for a batch of queries is a dense GEMM,
// scores = Q · Vᵀ -- the one place on the nidx_vector path where the tensor cores apply.  The product is needed only to
// find each query's top-k, so it is computed ONCE in TF32 (f32 operands straight from the stored rows, f32 accumulation in
// tensor memory) as a filter with a rigorous error bound, never written to HBM, and only the few survivors are re-scored
// with the lane-blocked f32 arithmetic of common.cuh.  Ids AND scores are therefore bit-identical to the small-batch kernel
// and to the oracle (round 1's 3xTF32 kernel was 1e-6 off and materialised the [Q x N] matrix).
//
//   |tf32(q) . tf32(v) - q . v| <= 2^-9 |q| |v|  (each operand keeps 10 mantissa bits: relative error < 2^-10 per factor)
//   eps = 2.2e-3 (cosine) or 2.2e-3 |q| max|v| (dot).  Let tau = the k-th largest APPROXIMATE score of a query over the
//   eligible vectors: every member of the true top-k has approx >= tau - 2 eps.  Each CTA keeps, per query row, the best
//   TC2_L approximate scores of ITS share of the vectors (a list in registers); the k best approximations overall are in
//   those lists (at most k - 1 entries of a share rank above any of them, and L >= k), so tau is known exactly from the
//   lists; survivors = list entries with approx >= tau - 2 eps.  A list that is full and whose smallest entry is still
//   >= tau - 2 eps may have dropped a survivor: OVERFLOW, the query is then scanned exactly (rare).
//
// scan_tc_filter_kernel: persistent CTAs; a CTA serves ONE 128-query block (slot s of grid / n_qblocks slots) and walks the
// vector chunks s, s + slots, ...; the CTAs of one slot run together and share every chunk (L2 reuse).  Warp-specialised:
//   warp 0   TMA producer: cp.async.bulk.tensor (128-byte swizzle) of a 128 x 32-float query tile and a 256 x 32-float
//            vector tile per stage into a 4-stage mbarrier ring (48 KB per stage);
//   warp 1   MMA issuer: one thread, tcgen05.mma.cta_group::1.kind::tf32, M = 128 (queries) x N = 256 (vectors) x K = 8,
//            four per stage; tcgen05.commit releases the stage; accumulators double-buffered in TMEM (2 x 256 columns);
//   warp 2   TMEM allocation;
//   warps 4-11 epilogue: warp % 4 = TMEM lane quadrant, thread = query row; warps 4-7 read columns 0-127 of every tile, warps 8-11
//            columns 128-255 (the read-out, not the tensor pipe, paces the kernel); tcgen05.ld 32 columns at a time, cosine scaling, eligibility bit,
//            running top-L of the row in REGISTERS (unsorted + its minimum), fed through a per-row staging area in shared memory
//            so that the (warp-wide) list update runs once per ~8-16 candidates of the busiest row, not once per column.
// scan_tc_refine_kernel: one CTA per query: overflow test, tau, survivors, exact re-scoring (one warp per survivor),
//   min_score, top-k -- or the exact scan of the whole segment for an overflowed query.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "scan_tc.cuh"
#include "topk.cuh"

namespace nidx {

constexpr int TC2_M = 128;            // queries per block (UMMA M, TMEM lanes)
constexpr int TC2_N = 256;            // vectors per tile (UMMA N)
constexpr int TC2_KB = 32;            // floats per k-block = one 128-byte swizzle row
constexpr int TC2_STAGES = 4;
constexpr int TC2_TILES = 8;          // vector tiles per chunk
constexpr int TC2_CHUNK = TC2_N * TC2_TILES;   // 2048 vectors
constexpr int TC2_L = 24;             // candidates kept per (query, chunk)
constexpr int TC2_KMAX = 16;          // the filter path serves k <= TC2_KMAX
constexpr int TC2_EPI_WARPS = 8;      // epilogue warps: TMEM lane quadrant = warp % 4, column half of the tile = warp / 4
constexpr int TC2_THREADS = (4 + TC2_EPI_WARPS) * 32;
constexpr uint32_t TC2_A_BYTES = TC2_M * 128, TC2_B_BYTES = TC2_N * 128, TC2_STAGE_BYTES = TC2_A_BYTES + TC2_B_BYTES;
constexpr int TC2_STAGE_ROWS = 12;    // staged candidates per (query row, column half) between two merges into the register list
constexpr int TC2_LISTS = TC2_EPI_WARPS / 4;   // candidate lists per query row and CTA (one per column half)
constexpr size_t TC2_SMEM_BYTES = 1024 /* alignment slack */ + (size_t)TC2_STAGES * TC2_STAGE_BYTES + 2 * TC2_N * 4 /* 1/|v| */ +
                                  2 * (TC2_N / 32) * 4 /* eligibility */ + (size_t)TC2_STAGE_ROWS * TC2_M * TC2_LISTS * 8 /* staging */ + 256;
constexpr float TC2_EPS = 2.2e-3f;
constexpr int TC2_SURV_CAP = 512;     // survivors per query the refine kernel re-scores; more => exact scan

__device__ __forceinline__ uint64_t tc2_desc(uint32_t smem_addr) {
    // K-major, SWIZZLE_128B (cute::UMMA::make_umma_desc<Major::K>): start >> 4, LBO = 1, SBO = 8 rows x 128 B = 1024 B, version 1, layout 2
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t tc2_idesc() {
    // c_format F32 = 1 [4,6), a/b_format TF32 = 2 [7,10) [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC2_N >> 3) << 17) | ((uint32_t)(TC2_M >> 4) << 24);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}

struct Tc2Args {
    int nq, n_qblocks, n_chunks, slots;   // slots = CTAs per query block (1 when there are more query blocks than CTAs)
    const float* qnorms;          // [nq] (cosine)
    const uint64_t* bits;         // eligibility (alive & filter) per vector, or nullptr
    float* cand_score;            // [nq][slots * TC2_LISTS][TC2_L] approx scores, unsorted, -inf padded
    uint32_t* cand_id;            // [nq][slots * TC2_LISTS][TC2_L]
    unsigned int* work_counter;
};

// The (query block, chunk) sequence of a CTA -- the same in the three roles.  n_qblocks <= gridDim: CTA c serves block c % n_qblocks
// as slot c / n_qblocks (CTAs beyond n_qblocks * slots idle); else one slot and CTA c serves blocks c, c + gridDim, ...
struct Tc2Sched {
    int g0, gstride, slot, slots;
    __device__ Tc2Sched(const Tc2Args& a) {
        slots = a.slots;
        if (a.n_qblocks <= (int)gridDim.x) { g0 = (int)blockIdx.x % a.n_qblocks; slot = (int)blockIdx.x / a.n_qblocks; gstride = a.n_qblocks; if (slot >= slots) g0 = a.n_qblocks; }
        else { g0 = (int)blockIdx.x; slot = 0; gstride = (int)gridDim.x; }
    }
};

__global__ void __launch_bounds__(TC2_THREADS, 1) scan_tc_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_v,
                                                                        VecDev V, Tc2Args a) {
    extern __shared__ unsigned char tc2_raw[];
    __shared__ uint64_t full[TC2_STAGES], empty[TC2_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tc2_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char* stages = smem;
    float* inv_vn = reinterpret_cast<float*>(smem + (size_t)TC2_STAGES * TC2_STAGE_BYTES);     // [2][256]
    uint32_t* elig = reinterpret_cast<uint32_t*>(inv_vn + 2 * TC2_N);                          // [2][8]
    float* stg_sc = reinterpret_cast<float*>(elig + 2 * (TC2_N / 32));                         // [TC2_STAGE_ROWS][128] staged scores ...
    uint32_t* stg_id = reinterpret_cast<uint32_t*>(stg_sc + TC2_STAGE_ROWS * TC2_M * TC2_LISTS);   // ... and ids of the epilogue threads
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = V.ld / TC2_KB;
    const Tc2Sched sch(a);

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
        mbar_init(&tmem_empty[0], TC2_EPI_WARPS); mbar_init(&tmem_empty[1], TC2_EPI_WARPS);   // one arrive per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(2 * TC2_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int qb = sch.g0; qb < a.n_qblocks; qb += sch.gstride)
                for (int ch = sch.slot; ch < a.n_chunks; ch += sch.slots)
                    for (int t = 0; t < TC2_TILES; ++t) {
                        int v0 = ch * TC2_CHUNK + t * TC2_N;
                        if ((uint32_t)v0 >= V.n) break;
                        for (int kb = 0; kb < n_kb; ++kb) {
                            mbar_wait(&empty[st], ph ^ 1);
                            mbar_expect_tx(&full[st], TC2_STAGE_BYTES);
                            unsigned char* sa = stages + (size_t)st * TC2_STAGE_BYTES;
                            tma_load_2d(sa, &map_q, kb * TC2_KB, qb * TC2_M, &full[st]);
                            tma_load_2d(sa + TC2_A_BYTES, &map_v, kb * TC2_KB, v0, &full[st]);
                            if (++st == TC2_STAGES) { st = 0; ph ^= 1; }
                        }
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = tc2_idesc();
            const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stages);
            uint32_t st = 0, ph = 0, tcount = 0;
            for (int qb = sch.g0; qb < a.n_qblocks; qb += sch.gstride)
                for (int ch = sch.slot; ch < a.n_chunks; ch += sch.slots)
                    for (int t = 0; t < TC2_TILES; ++t) {
                        int v0 = ch * TC2_CHUNK + t * TC2_N;
                        if ((uint32_t)v0 >= V.n) break;
                        uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                        mbar_wait(&tmem_empty[acc], aph ^ 1);               // the epilogue has drained this accumulator
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        for (int kb = 0; kb < n_kb; ++kb) {
                            mbar_wait(&full[st], ph);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            uint32_t sa = sbase + st * TC2_STAGE_BYTES, sb = sa + TC2_A_BYTES;
#pragma unroll
                            for (int ks = 0; ks < TC2_KB / 8; ++ks)    // K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
                                tc_mma(tmem_d + acc * TC2_N, tc2_desc(sa + ks * 32), tc2_desc(sb + ks * 32), idesc, (kb | ks) != 0);
                            tc_commit(&empty[st]);                          // frees the stage when these MMAs have read it
                            if (++st == TC2_STAGES) { st = 0; ph ^= 1; }
                        }
                        tc_commit(&tmem_full[acc]);                         // accumulator complete
                        ++tcount;
                    }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = query row =====
        const int row = (warp & 3) * 32 + lane;                 // TMEM lane = query row of the block
        const int half = (warp - 4) >> 2;                       // which TC2_N / TC2_LISTS columns of every tile this warp reads
        const int et = threadIdx.x - 128;                       // 0 .. 32 * TC2_EPI_WARPS - 1
        const int srow = half * TC2_M + row;                    // this thread's column of the staging area
        uint32_t tcount = 0;
        for (int qb = sch.g0; qb < a.n_qblocks; qb += sch.gstride) {
            int q = qb * TC2_M + row;
            float inv_qn = 1.0f;
            if (V.sim == SIM_COSINE) { float qn = q < a.nq ? a.qnorms[q] : 0.0f; inv_qn = qn > 0.0f ? __frcp_rn(qn) : 0.0f; }
            // the row's best TC2_L approximate scores over this CTA's share: unsorted, with the minimum and where it sits
            float ls[TC2_L];
            uint32_t li[TC2_L];
#pragma unroll
            for (int i = 0; i < TC2_L; ++i) { ls[i] = -INFINITY; li[i] = NIL; }
            int cnt = 0, minpos = 0, n_st = 0;
            float thr = -INFINITY;       // scores <= thr cannot enter the list (-inf until it is full)
            // Candidates that beat the row's threshold are first appended to a per-row staging area in shared memory ([slot][row]: the
            // lanes of a warp hit 32 different banks) and merged into the register list only when some row's area is nearly full: the
            // list update -- ~70 predicated instructions, executed by the whole warp whenever ANY lane needs it -- then runs a few
            // dozen times per row share instead of once per column.  Staged entries keep their column order and are re-tested against
            // the up-to-date threshold at merge time, so the list ends up exactly as if every column had been merged at once.
            auto merge_staged = [&]() {
                for (int t = 0; t < n_st; ++t) {
                    const float sc = stg_sc[t * (TC2_M * TC2_LISTS) + srow];
                    if (sc > thr) {
                        const uint32_t id = stg_id[t * (TC2_M * TC2_LISTS) + srow];
                        const int pos = cnt < TC2_L ? cnt : minpos;
#pragma unroll
                        for (int i = 0; i < TC2_L; ++i) if (i == pos) { ls[i] = sc; li[i] = id; }   // static indices: predicated moves
                        if (cnt < TC2_L) ++cnt;
                        if (cnt == TC2_L) {
                            float m = ls[0];
                            int mp = 0;
#pragma unroll
                            for (int i = 1; i < TC2_L; ++i) if (ls[i] < m) { m = ls[i]; mp = i; }
                            thr = m; minpos = mp;
                        }
                    }
                }
                n_st = 0;
            };
            for (int ch = sch.slot; ch < a.n_chunks; ch += sch.slots)
                for (int t = 0; t < TC2_TILES; ++t) {
                    int v0 = ch * TC2_CHUNK + t * TC2_N;
                    if ((uint32_t)v0 >= V.n) break;
                    uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                    // per-tile column data: 1 / |v| (cosine) and the eligibility bits, by the 128 epilogue threads
                    for (int j = et; j < TC2_N; j += 32 * TC2_EPI_WARPS) {
                        uint32_t v = (uint32_t)v0 + j;
                        float ivn = 1.0f;
                        if (V.sim == SIM_COSINE) { float vn = v < V.n ? __ldg(V.norms + v) : 0.0f; ivn = vn > 0.0f ? __frcp_rn(vn) : 0.0f; }
                        inv_vn[acc * TC2_N + j] = ivn;
                    }
                    if (et < TC2_N / 32) {
                        uint32_t w = 0xFFFFFFFFu;
                        uint32_t vb = (uint32_t)v0 + et * 32;
                        if (a.bits) w = vb < V.n ? reinterpret_cast<const uint32_t*>(a.bits)[vb >> 5] : 0u;
                        if (vb + 32 > V.n) w &= vb < V.n ? (0xFFFFFFFFu >> (32 - (V.n - vb))) : 0u;   // columns beyond the segment
                        elig[acc * (TC2_N / 32) + et] = w;
                    }
                    asm volatile("bar.sync 1, %0;" :: "n"(32 * TC2_EPI_WARPS) : "memory");
                    mbar_wait(&tmem_full[acc], aph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int c0 = half * (TC2_N / TC2_LISTS); c0 < (half + 1) * (TC2_N / TC2_LISTS); c0 += 32) {
                        uint32_t r[32];
                        uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + acc * TC2_N + c0;
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                     : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        const uint32_t ew = q < a.nq ? elig[acc * (TC2_N / 32) + (c0 >> 5)] : 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float sc = __uint_as_float(r[j]);
                            if (V.sim == SIM_COSINE) sc = sc * inv_qn * inv_vn[acc * TC2_N + c0 + j];
                            const bool pass = ((ew >> j) & 1u) && sc > thr;
                            if (pass) {                                   // predicated stores: the row's staging area, in column order
                                stg_sc[n_st * (TC2_M * TC2_LISTS) + srow] = sc;
                                stg_id[n_st * (TC2_M * TC2_LISTS) + srow] = (uint32_t)v0 + c0 + j;
                                ++n_st;
                            }
                            if ((j & 3) == 3 && __any_sync(0xFFFFFFFFu, n_st > TC2_STAGE_ROWS - 4)) merge_staged();
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    ++tcount;
                }
            merge_staged();
            if (q < a.nq) {
                float* os = a.cand_score + (((size_t)q * a.slots + sch.slot) * TC2_LISTS + half) * TC2_L;
                uint32_t* oi = a.cand_id + (((size_t)q * a.slots + sch.slot) * TC2_LISTS + half) * TC2_L;
#pragma unroll
                for (int i = 0; i < TC2_L; ++i) { os[i] = ls[i]; oi[i] = li[i]; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(2 * TC2_N) : "memory");
}

// One CTA per query.  dynamic smem: ld floats (query) + cap keys (top-k buffer) + survivors.
__host__ __device__ __forceinline__ size_t tc2_refine_smem(int ld, int cap) { return (size_t)ld * 4 + (size_t)cap * 8 + (size_t)TC2_SURV_CAP * 4 + 64; }

__global__ void __launch_bounds__(256) scan_tc_refine_kernel(VecDev V, const float* __restrict__ queries, const float* __restrict__ qnorms, int n_lists,
                                                             const float* __restrict__ cand_score, const uint32_t* __restrict__ cand_id,
                                                             const uint64_t* __restrict__ bits, float max_vnorm, float min_score, int k, int cap,
                                                             uint32_t* __restrict__ out_ids, float* __restrict__ out_scores, int* __restrict__ out_counts,
                                                             unsigned long long* __restrict__ stats /* [0] survivors [1] overflowed queries */) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    __shared__ int s_overflow, s_nsurv;
    float* qv = reinterpret_cast<float*>(rf_smem);
    uint64_t* tk_buf = reinterpret_cast<uint64_t*>(rf_smem + (size_t)V.ld * 4);
    uint32_t* surv = reinterpret_cast<uint32_t*>(tk_buf + cap);
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ng = V.ld >> 2;
    for (int i = threadIdx.x; i < ng; i += blockDim.x) reinterpret_cast<float4*>(qv)[i] = reinterpret_cast<const float4*>(queries + (size_t)q * V.ld)[i];
    const float qn = V.sim == SIM_COSINE ? qnorms[q] : 0.0f;
    // eps of this query: cosine scores are scale free; a dot product scales with both norms
    float eps = TC2_EPS;
    if (V.sim != SIM_COSINE) {
        float s = 0.0f;
        for (int i = lane; i < V.d; i += 32) { float x = queries[(size_t)q * V.ld + i]; s += x * x; }
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
        eps = TC2_EPS * sqrtf(s) * max_vnorm;
    }
    const float margin = 2.0f * eps;
    if (threadIdx.x == 0) { s_overflow = (V.sim == SIM_COSINE && !(qn > 0.0f)) ? 1 : 0; s_nsurv = 0; }
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, k, cap);
    const float* cs = cand_score + (size_t)q * n_lists * TC2_L;
    const uint32_t* ci = cand_id + (size_t)q * n_lists * TC2_L;
    // 1. tau = k-th largest approximate score among all candidates
    const int total = n_lists * TC2_L;
    for (int base = 0; base < total; base += blockDim.x) {
        int i = base + threadIdx.x;
        uint64_t key = 0;
        if (i < total && ci[i] != NIL) key = make_key(cs[i], (uint32_t)i, 0);
        tk.offer(key);
    }
    int c = tk.finish();
    const float tau = c >= k ? key_score(tk_buf[k - 1]) : -INFINITY;
    __syncthreads();
    // 2. overflow test per list: full (no NIL entry), and its smallest entry still within the margin of tau
    for (int l = threadIdx.x; l < n_lists; l += blockDim.x) {
        float mn = INFINITY;
        bool full = true;
        for (int i = 0; i < TC2_L; ++i) { full &= ci[(size_t)l * TC2_L + i] != NIL; mn = fminf(mn, cs[(size_t)l * TC2_L + i]); }
        if (full && mn >= tau - margin) s_overflow = 1;
    }
    // 3. survivors
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        if (ci[i] != NIL && cs[i] >= tau - margin) {
            int pos = atomicAdd(&s_nsurv, 1);
            if (pos < TC2_SURV_CAP) surv[pos] = ci[i];
        }
    }
    __syncthreads();
    const bool exact_all = s_overflow || s_nsurv > TC2_SURV_CAP;
    const int nsurv = exact_all ? (int)V.n : s_nsurv;
    if (threadIdx.x == 0 && stats) { atomicAdd(&stats[0], (unsigned long long)(exact_all ? 0 : nsurv)); if (exact_all) atomicAdd(&stats[1], 1ull); }
    // 4. exact scores (lane-blocked arithmetic: bit-identical to scan_scores_kernel), min_score, top-k
    tk.init(tk_buf, &tk_count, &tk_thr, k, cap);
    for (int base = 0; base < nsurv; base += blockDim.x) {
        // one warp per candidate, 8 candidates per round, keys offered by lane 0 of each warp in a full-block round
        uint64_t mykey = 0;
        for (int w8 = 0; w8 < 32; ++w8) {            // 32 sub-rounds x 8 warps = 256 candidates per offer round
            int i = base + w8 * 8 + warp;
            uint64_t key = 0;
            if (i < nsurv) {
                uint32_t v = exact_all ? (uint32_t)i : surv[i];
                bool ok = !exact_all || !bits || ((bits[v >> 6] >> (v & 63)) & 1);
                if (ok) {
                    float ab = warp_dot(reinterpret_cast<const float4*>(V.vecs + (size_t)v * V.ld), reinterpret_cast<const float4*>(qv), ng, lane);
                    float s = sim_from_parts(V.sim, ab, V.norms[v], qn);
                    if (s >= min_score) key = make_key(s, v, 0);
                }
            }
            if (lane == w8) mykey = key;             // lane w8 of warp `warp` carries candidate base + w8 * 8 + warp
        }
        tk.offer(mykey);
    }
    c = tk.finish();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_ids[(size_t)q * k + i] = i < c ? key_id(tk_buf[i]) : NIL;
        out_scores[(size_t)q * k + i] = i < c ? key_score(tk_buf[i]) : 0.0f;
    }
    if (threadIdx.x == 0 && out_counts) out_counts[q] = c;
}

}  // namespace nidx
