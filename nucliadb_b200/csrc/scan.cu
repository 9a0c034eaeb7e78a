// nidx_b200 — K1: exact scan (brute force) for nidx_vector.
//
// Replaces OpenSegment::brute_force_search (nidx/nidx_vector/src/segment.rs:569-623): for every
// alive paragraph take its best vector's similarity (dense_f32.rs:29-39), keep it iff
// score >= min_score, sort descending, take k.  The reference does this per query with one
// simsimd call per vector and a full sort of N pairs; here a batch of queries shares every pass
// over the [N][D] f32 block in HBM.
//
//   scan_scores_kernel   HBM-bound for small batches (N*D*4 bytes per pass), FMA/smem-bound for
//                        large ones; scores[q][v] f32 written coalesced (1/D of the read traffic).
//   scan_select_kernel   per (query, chunk): paragraph max + filter + min_score + block top-k.
//   topk_merge_kernel    per query: merge the chunk partials (also used across segments / GPUs).
#include "common.cuh"
#include "topk.cuh"

namespace nidx {

// [n][src_ld] (row stride in BYTES, rows may be 4-byte aligned only: vectors.bin records are
// 4*d+4 bytes, data_store/v2/vector_store.rs:33-68) -> [n][ld] f32 zero padded.
__global__ void pad_rows_kernel(const unsigned char* __restrict__ src, size_t src_stride_bytes, int d, float* __restrict__ dst, int ld,
                                uint64_t n) {
    uint64_t total = n * (uint64_t)ld;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / ld;
        int c = (int)(i % ld);
        dst[i] = c < d ? *reinterpret_cast<const float*>(src + r * src_stride_bytes + (size_t)c * 4) : 0.0f;
    }
}

// norms[r] = sqrt(dot_ordered(row, row)); one warp per row.
__global__ void row_norms_kernel(const float* __restrict__ rows, int ld, uint64_t n, float* __restrict__ norms) {
    int lane = threadIdx.x & 31;
    uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float4* a = reinterpret_cast<const float4*>(rows + r * ld);
        float s = warp_dot(a, a, ld >> 2, lane);
        if (lane == 0) norms[r] = __fsqrt_rn(s);
    }
}

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_QT = 8;      // queries per tile (shared memory)
constexpr int SCAN_VPW = 32;    // vectors per warp (one score per lane => coalesced stores)

// grid: 1-D, query tile fastest so CTAs that share a vector chunk run together (L2 reuse).
// dynamic smem: SCAN_QT * ld floats.
__global__ void __launch_bounds__(SCAN_WARPS * 32) scan_scores_kernel(VecDev V, const float* __restrict__ queries,
                                                                       const float* __restrict__ qnorms, int nq, int n_qtiles,
                                                                       float* __restrict__ scores) {
    extern __shared__ __align__(16) float qs[];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int qtile = blockIdx.x % n_qtiles;
    uint64_t chunk = blockIdx.x / n_qtiles;
    int q0 = qtile * SCAN_QT;
    int nqt = min(SCAN_QT, nq - q0);
    int ng = V.ld >> 2;
    for (int i = threadIdx.x; i < nqt * ng; i += blockDim.x)
        reinterpret_cast<float4*>(qs)[i] = reinterpret_cast<const float4*>(queries + (size_t)q0 * V.ld)[i];
    __syncthreads();

    uint64_t v0 = (chunk * SCAN_WARPS + warp) * SCAN_VPW;
    if (v0 >= V.n) return;
    int nv = (int)min((uint64_t)SCAN_VPW, (uint64_t)V.n - v0);
    float mine[SCAN_QT];
#pragma unroll
    for (int qi = 0; qi < SCAN_QT; ++qi) mine[qi] = 0.0f;

    for (int j = 0; j < nv; ++j) {
        uint32_t v = (uint32_t)(v0 + j);
        const float4* a = reinterpret_cast<const float4*>(V.vecs + (size_t)v * V.ld);
        float vnorm = V.sim != SIM_DOT ? V.norms[v] : 0.0f;
        if (ng <= 256) {
            float4 va[8];  // the whole row, lane-sliced, read from HBM once for all queries of the tile
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                int g = t * 32 + lane;
                va[t] = g < ng ? ldg_stream(a + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int qi = 0; qi < SCAN_QT; ++qi) {
                if (qi < nqt) {
                    const float4* b = reinterpret_cast<const float4*>(qs) + (size_t)qi * ng;
                    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        int g = t * 32 + lane;
                        if (g < ng) {
                            float4 vb = b[g];
                            ax = __fmaf_rn(va[t].x, vb.x, ax);
                            ay = __fmaf_rn(va[t].y, vb.y, ay);
                            az = __fmaf_rn(va[t].z, vb.z, az);
                            aw = __fmaf_rn(va[t].w, vb.w, aw);
                        }
                    }
                    float ab = butterfly_sum(__fadd_rn(__fadd_rn(ax, ay), __fadd_rn(az, aw)));
                    float s = sim_from_parts(V.sim, ab, vnorm, qnorms[q0 + qi]);
                    if (lane == j) mine[qi] = s;
                }
            }
        } else {
            for (int qi = 0; qi < nqt; ++qi) {
                float ab = warp_dot(a, reinterpret_cast<const float4*>(qs) + (size_t)qi * ng, ng, lane);
                float s = sim_from_parts(V.sim, ab, vnorm, qnorms[q0 + qi]);
                if (lane == j) mine[qi] = s;
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < SCAN_QT; ++qi)
        if (qi < nqt && lane < nv) scores[(size_t)(q0 + qi) * V.n + v0 + lane] = mine[qi];
}

// Compile-time row length (ld == NG * 128 floats), VU vectors per warp in flight: each query float4 read from
// shared memory feeds VU vectors (VU x fewer shared-memory bytes per FMA) and VU rows' loads overlap.
// Same lane-blocked arithmetic as warp_dot => bit-identical scores.
template <int NG, int VU>
__global__ void __launch_bounds__(SCAN_WARPS * 32) scan_scores_kernel_t(VecDev V, const float* __restrict__ queries, const float* __restrict__ qnorms, int nq,
                                                                         int n_qtiles, float* __restrict__ scores) {
    extern __shared__ __align__(16) float qs[];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int qtile = blockIdx.x % n_qtiles;
    uint64_t chunk = blockIdx.x / n_qtiles;
    int q0 = qtile * SCAN_QT;
    int nqt = min(SCAN_QT, nq - q0);
    constexpr int ng = NG * 32;
    for (int i = threadIdx.x; i < nqt * ng; i += blockDim.x)
        reinterpret_cast<float4*>(qs)[i] = reinterpret_cast<const float4*>(queries + (size_t)q0 * V.ld)[i];
    __syncthreads();
    uint64_t v0 = (chunk * SCAN_WARPS + warp) * SCAN_VPW;
    if (v0 >= V.n) return;
    int nv = (int)min((uint64_t)SCAN_VPW, (uint64_t)V.n - v0);
    float mine[SCAN_QT];
#pragma unroll
    for (int qi = 0; qi < SCAN_QT; ++qi) mine[qi] = 0.0f;
    for (int j0 = 0; j0 < nv; j0 += VU) {
        float4 va[VU][NG];
        float vnorm[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) {
            uint32_t v = (uint32_t)(v0 + min(j0 + u, nv - 1));   // tail: recompute the last row, result ignored
            const float4* a = reinterpret_cast<const float4*>(V.vecs + (size_t)v * V.ld);
#pragma unroll
            for (int t = 0; t < NG; ++t) va[u][t] = ldg_stream(a + t * 32 + lane);
            vnorm[u] = V.sim != SIM_DOT ? __ldg(V.norms + v) : 0.0f;
        }
#pragma unroll
        for (int qi = 0; qi < SCAN_QT; ++qi) {
            if (qi < nqt) {
                const float4* b = reinterpret_cast<const float4*>(qs) + (size_t)qi * ng;
                float acc[VU][4];
#pragma unroll
                for (int u = 0; u < VU; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
#pragma unroll
                for (int t = 0; t < NG; ++t) {
                    float4 vb = b[t * 32 + lane];
#pragma unroll
                    for (int u = 0; u < VU; ++u) {
                        acc[u][0] = __fmaf_rn(va[u][t].x, vb.x, acc[u][0]);
                        acc[u][1] = __fmaf_rn(va[u][t].y, vb.y, acc[u][1]);
                        acc[u][2] = __fmaf_rn(va[u][t].z, vb.z, acc[u][2]);
                        acc[u][3] = __fmaf_rn(va[u][t].w, vb.w, acc[u][3]);
                    }
                }
                float qn = V.sim != SIM_DOT ? qnorms[q0 + qi] : 0.0f;
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    float ab = butterfly_sum(__fadd_rn(__fadd_rn(acc[u][0], acc[u][1]), __fadd_rn(acc[u][2], acc[u][3])));
                    float s = sim_from_parts(V.sim, ab, vnorm[u], qn);
                    if (lane == j0 + u) mine[qi] = s;
                }
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < SCAN_QT; ++qi)
        if (qi < nqt && lane < nv) scores[(size_t)(q0 + qi) * V.n + v0 + lane] = mine[qi];
}

// Per (chunk, query): best vector per alive+filtered paragraph (segment.rs:581-597), min_score (>=),
// block top-k.  par_first == nullptr: one vector per paragraph.  partial: [nq][n_chunks][k] keys.
// dynamic smem: cap * 8 bytes.
__global__ void scan_select_kernel(const float* __restrict__ scores, uint32_t n_vec, uint32_t n_par, const uint32_t* __restrict__ par_first,
                                   const uint64_t* __restrict__ alive, const uint64_t* __restrict__ filter, float min_score, int k, int cap,
                                   int n_chunks, uint64_t* __restrict__ partial) {
    extern __shared__ __align__(16) uint64_t tk_buf[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    int q = blockIdx.y, chunk = blockIdx.x;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, k, cap);
    const float* sc = scores + (size_t)q * n_vec;
    uint32_t per = (n_par + n_chunks - 1) / n_chunks;
    uint32_t p0 = chunk * per, p1 = min(n_par, p0 + per);
    for (uint32_t base = p0; base < p1; base += blockDim.x) {
        uint32_t p = base + threadIdx.x;
        uint64_t key = 0;
        if (p < p1) {
            bool ok = true;
            if (alive) ok = (alive[p >> 6] >> (p & 63)) & 1;
            if (ok && filter) ok = (filter[p >> 6] >> (p & 63)) & 1;
            if (ok) {
                uint32_t va = par_first ? par_first[p] : p, vb = par_first ? par_first[p + 1] : p + 1;
                float best = 0.f;
                uint32_t bestv = NIL;
                for (uint32_t v = va; v < vb; ++v) {  // max_by(total_cmp): the later element wins ties
                    float s = sc[v];
                    if (bestv == NIL || ordered_bits(s) >= ordered_bits(best)) { best = s; bestv = v; }
                }
                if (bestv != NIL && best >= min_score) key = make_key(best, bestv, 0);
            }
        }
        tk.offer(key);
    }
    int c = tk.finish();
    uint64_t* out = partial + ((size_t)q * n_chunks + chunk) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = i < c ? tk_buf[i] : 0;
}

// Per query: top-k of n_in keys (0 = empty) -> ids / scores / count.  dynamic smem: cap * 8 bytes.
__global__ void topk_merge_kernel(const uint64_t* __restrict__ keys_in, int n_in, int k, int cap, uint32_t* __restrict__ out_ids,
                                  float* __restrict__ out_scores, int* __restrict__ out_counts) {
    extern __shared__ __align__(16) uint64_t tk_buf[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    int q = blockIdx.x;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, k, cap);
    const uint64_t* in = keys_in + (size_t)q * n_in;
    for (int base = 0; base < n_in; base += blockDim.x) {
        int i = base + threadIdx.x;
        tk.offer(i < n_in ? in[i] : 0);
    }
    int c = tk.finish();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_ids[(size_t)q * k + i] = i < c ? key_id(tk_buf[i]) : NIL;
        out_scores[(size_t)q * k + i] = i < c ? key_score(tk_buf[i]) : 0.0f;
    }
    if (threadIdx.x == 0 && out_counts) out_counts[q] = c;
}

// Cross-part merge (searcher.rs:241-290 / shard_merge.rs:332-348): parts [n_parts][nq][k] of (id, score)
// sorted desc -> [nq][k] plus the originating part.  Keys here rank (score desc, part asc, position asc),
// which is what kmerge_by(|a, b| a.score >= b.score) yields for inputs that are each sorted.
__global__ void parts_merge_kernel(const uint32_t* __restrict__ ids, const float* __restrict__ scores, int n_parts, size_t part_stride, int nq, int k, int cap,
                                   uint32_t* __restrict__ out_ids, float* __restrict__ out_scores, int* __restrict__ out_part) {
    extern __shared__ __align__(16) uint64_t tk_buf[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    int q = blockIdx.x;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, k, cap);
    int total = n_parts * k;
    for (int base = 0; base < total; base += blockDim.x) {
        int i = base + threadIdx.x;
        uint64_t key = 0;
        if (i < total) {
            int part = i / k, pos = i % k;
            size_t src = (size_t)part * part_stride + (size_t)q * k + pos;
            if (ids[src] != NIL) key = make_key(scores[src], (uint32_t)i, 0);  // id field = part*k+pos
        }
        tk.offer(key);
    }
    int c = tk.finish();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        size_t dst = (size_t)q * k + i;
        if (i < c) {
            uint32_t slot = key_id(tk_buf[i]);
            int part = slot / k, pos = slot % k;
            size_t src = (size_t)part * part_stride + (size_t)q * k + pos;
            out_ids[dst] = ids[src];
            out_scores[dst] = scores[src];
            if (out_part) out_part[dst] = part;
        } else {
            out_ids[dst] = NIL;
            out_scores[dst] = 0.0f;
            if (out_part) out_part[dst] = -1;
        }
    }
}

}  // namespace nidx
