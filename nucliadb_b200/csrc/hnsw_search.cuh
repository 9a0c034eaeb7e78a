// nidx_b200 — K2/K3: HNSW search for nidx_vector (sm_100a).
//
// One CTA walks one query (or, in build mode, one node to insert) through the graph:
//   HnswSearcher::layer_search      nidx/nidx_vector/src/hnsw/search.rs:242-304
//   HnswSearcher::search            search.rs:306-383  (descent with k=1, layer 0 with max(k, ef))
//   HnswSearcher::closest_up_nodes  search.rs:188-240  (filter / dedup aware expansion to k results)
//   NodeFilter::passes              search.rs:135-171
//   HnswBuilder::insert (search half) hnsw/build.rs:123-150
//
// Data structures per CTA, all in shared memory:
//   * ONE sorted list of 64-bit rank keys (score desc, id asc) with an "unexpanded" flag in bit 0.
//     It is the reference's two BinaryHeaps folded together: a candidate that is not among the best
//     `ef` results can never be expanded (popping it ends the search, search.rs:268-273), so the
//     candidates that matter are exactly the unexpanded members of the result list.
//   * an exact visited set (open-addressing hash of node ids) = the reference's FxHashSet;
//   * the query vector; up to 64 neighbour ids / keys of the node being expanded.
// Per expansion: warp 0 reads the adjacency row (one 128-byte line for M0=32) and filters it through
// the visited set; all warps compute one similarity each (3 KB coalesced, streamed past L1; the
// lane-blocked summation order of common.cuh makes scores bit-identical to the oracle's); the CTA
// merges the admitted keys into the list by rank (no sort).
#pragma once
#include <limits.h>

#include "common.cuh"

namespace nidx {

constexpr int HS_WARPS = 8;
constexpr int HS_THREADS = HS_WARPS * 32;
constexpr int HS_MAX_ROW = 64;      // stride of an adjacency row (M0 <= 64)
constexpr int HS_MAX_LAYERS = 8;    // found[] layers kept per inserted node (level > 7 has p < 1e-9)

struct SearchArgs {
    int mode;  // 0 = query (search.rs:306-383), 1 = build (build.rs:123-150)
    int nq;
    // query mode
    const float* queries;  // [nq][ld] zero padded
    const float* qnorms;   // [nq] (cosine)
    int k, ef0;            // ef0 = max(k, ef)
    float min_score;
    int with_duplicates, multi_vector;
    const uint64_t* filter;  // bitset over paragraphs (filter ∧ alive) or nullptr
    uint32_t* out_ids; float* out_scores; int* out_counts;
    // build mode
    const uint32_t* nodes;   // [nq]
    int efC;
    uint64_t* found;         // [nq][HS_MAX_LAYERS][efC] rank keys (flag bit 0)
    int* found_count;        // [nq][HS_MAX_LAYERS]
    // shared
    int hash_bits;           // visited table = 1 << hash_bits slots
    int list_cap;            // >= max(ef0, efC) and >= cu_cap
    int cu_cap;              // closest_up_nodes pending-candidate capacity
    unsigned int* work_counter;       // dynamic query scheduler (zeroed by the host)
    unsigned long long* counters;     // [0] similarities [1] expansions [2] visited overflows [3] cu overflows [4] RaBitQ estimates [5] exact similarities the sequential rerank_top needs
    // quantised walk (hnsw_rabitq.cuh; hnsw/search.rs:332-366 with SearchVector::RabitQ)
    const unsigned char* codes;       // [n][code_stride] vectors.quant records
    int code_stride;
    const uint32_t* planes;           // [nq][4][d/32] query bit planes
    const void* qparams;              // [nq] RabitqQueryParams
    uint32_t* gvisited;               // [grid][1 << gv_bits] layer-0 visited table in global memory (L2)
    int gv_bits;
    int last_k;                       // min(k * RERANKING_FACTOR, RERANKING_LIMIT)
    int rq_prefetch;                  // pull the predicted next node's neighbour codes / visited slots into L2
};

struct SearchCtx {
    float* qvec;
    uint64_t *A, *B;
    uint32_t* hash;
    uint32_t* todo_id;
    uint64_t* todo_key;
    uint32_t* pref_row;   // [2][HS_MAX_ROW] speculatively prefetched adjacency rows (double buffered by hop parity)
    uint32_t* pref_node;  // [2] node each buffer belongs to (NIL = none)
    int *s_len, *s_best, *s_best_next, *s_ntodo, *s_hash_count, *s_flag, *s_nadmit;
    int* s_bn;                       // [2] parity slots for s_best_next in the barrier-lean layer search
    unsigned long long* s_maxtodo;   // the largest admitted key of the expansion: list entries above it keep their position
    unsigned hop;
    uint32_t hash_mask;
    int hash_bits, hash_limit;
    float qnorm;
    unsigned long long n_dist, n_expand, n_overflow;
};

__host__ __device__ __forceinline__ size_t hs_smem_bytes(int ld, int list_cap, int hash_bits) {
    return (size_t)ld * 4 + (size_t)list_cap * 16 + ((size_t)4 << hash_bits) + HS_MAX_ROW * 12 + HS_MAX_ROW * 8 + 16 + 64;
}

// returns true iff y was not in the set (and is now).  A full table reports "already visited".
__device__ __forceinline__ bool hash_insert(SearchCtx& c, uint32_t y, bool& overflow) {
    if (*c.s_hash_count >= c.hash_limit) { overflow = true; return false; }
    uint32_t h = (y * 2654435761u) >> (32 - c.hash_bits);
    while (true) {
        uint32_t old = atomicCAS(&c.hash[h], NIL, y);
        if (old == NIL) return true;
        if (old == y) return false;
        h = (h + 1) & c.hash_mask;
    }
}

// Clear the visited set and seed it with the ids of the current list; mark every entry unexpanded.
__device__ inline void hs_reseed(SearchCtx& c) {
    __syncthreads();
    for (int i = threadIdx.x; i <= (int)c.hash_mask; i += blockDim.x) c.hash[i] = NIL;
    if (threadIdx.x == 0) { *c.s_hash_count = 0; *c.s_best = 0; c.pref_node[0] = NIL; c.pref_node[1] = NIL; }
    __syncthreads();
    int len = *c.s_len;
    bool ov = false;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        uint64_t key = c.A[i] | 1ull;
        c.A[i] = key;
        hash_insert(c, key_id(key), ov);
    }
    __syncthreads();
    if (threadIdx.x == 0) *c.s_hash_count = len;
    __syncthreads();
}

// Expand `node` (already chosen): gather unvisited neighbours (warp 0), score them (all warps).
// Admission: layer_search (search.rs:286) -- when the list holds `ef` entries only keys better than the
// worst survive; closest_up_nodes (search.rs:231) -- score >= min_score.
// Speculation: while warp 0 works, the last warp starts an asynchronous copy (cp.async) of the adjacency
// row of the node that will be expanded next if no new neighbour outranks it -- `pred_idx` in the list --
// into the other half of a double buffer; the row's HBM latency then overlaps this expansion's vector loads.
// LEAN (the layer search's hot loop): the list length arrives in a register and s_best_next alternates between two slots, so that
// hs_merge needs no barrier after thread 0 has published the new length -- three barriers per expansion instead of four.
template <bool CU, int NG, int W = HS_WARPS, bool PAIR = false, bool LEAN = false>
__device__ inline void hs_expand(const VecDev& V, const GraphDev& G, SearchCtx& c, uint32_t node, int layer, int ef, float min_score, int best, int len_in = -1) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int stride = G.stride(layer);
    unsigned cur = c.hop & 1u;
    if (LEAN) c.s_best_next = c.s_bn + cur;
    if (warp == 0) {
        const uint32_t* prow = c.pref_row + cur * HS_MAX_ROW;
        bool hit = c.pref_node[cur] == node;
        const uint32_t* row = G.row(node, layer);
        int ntodo = 0;
        bool ov = false;
        for (int e0 = 0; e0 < stride; e0 += 32) {
            uint32_t y = NIL;
            if (e0 + lane < stride) y = hit ? prow[e0 + lane] : __ldg(row + e0 + lane);
            bool fresh = (y != NIL) && hash_insert(c, y, ov);
            unsigned mask = __ballot_sync(0xFFFFFFFFu, fresh);
            if (fresh) c.todo_id[ntodo + __popc(mask & ((1u << lane) - 1))] = y;
            ntodo += __popc(mask);
        }
        if (__any_sync(0xFFFFFFFFu, ov) && lane == 0) c.n_overflow++;
        if (lane == 0) {
            *c.s_ntodo = ntodo;
            *c.s_hash_count += ntodo;
            *c.s_best_next = INT_MAX;
            *c.s_nadmit = 0;
            *c.s_maxtodo = 0;
            c.n_expand++;
        }
    } else if (warp == W - 1) {
        int len = LEAN ? len_in : *c.s_len;
        int pred = -1;
        if (CU) {
            pred = len > 1 ? 1 : -1;
        } else {
            for (int i0 = best + 1; i0 < len && pred < 0; i0 += 32) {
                int i = i0 + lane;
                unsigned m = __ballot_sync(0xFFFFFFFFu, i < len && (c.A[i] & 1ull));
                if (m) pred = i0 + __ffs(m) - 1;
            }
        }
        uint32_t pnode = NIL;
        if (pred >= 0) {
            pnode = key_id(c.A[pred]);
            const uint32_t* row2 = G.row(pnode, layer);
            uint32_t* dst = c.pref_row + (cur ^ 1u) * HS_MAX_ROW;
            for (int e = lane; e < stride; e += 32) cp_async4(dst + e, row2 + e);
        }
        if (lane == 0) c.pref_node[cur ^ 1u] = pnode;
    }
    __syncthreads();
    int ntodo = *c.s_ntodo, len = LEAN ? len_in : *c.s_len;
    uint64_t wkey = (!CU && len >= ef) ? c.A[len - 1] : 0;
    int ng = V.ld >> 2;
    auto finish = [&](int j, uint32_t y, float ab, float vnorm) {
        float s = sim_from_parts(V.sim, ab, vnorm, c.qnorm);
        uint64_t key = make_key(s, y, 1);
        bool admit = CU ? (s >= min_score) : (key > wkey);
        c.todo_key[j] = admit ? key : 0;
        if (admit) { atomicAdd(c.s_nadmit, 1); atomicMax(c.s_maxtodo, (unsigned long long)key); }
        c.n_dist++;
    };
    // A warp's later rows (j + W, j + 2W, ...) are pulled into L2 while it works on its first one: their loads then cost an L2
    // hit instead of a second and third HBM round trip on the expansion's critical path (one prefetch per 128-byte line, a lane
    // each; no registers held, unlike a second row in flight).
    if (!PAIR) {
        const int lines = (V.ld * 4 + 127) >> 7;
        for (int j = warp + W; j < ntodo; j += W) {
            const char* rowp = reinterpret_cast<const char*>(V.vecs + (size_t)c.todo_id[j] * V.ld);
            for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(rowp + (size_t)l * 128));
        }
    }
    if constexpr (PAIR) {   // two rows in flight per warp (fewer warps per query, more queries per SM)
        for (int j = warp; j < ntodo; j += 2 * W) {
            uint32_t y0 = c.todo_id[j];
            bool two = j + W < ntodo;
            uint32_t y1 = two ? c.todo_id[j + W] : y0;
            float n0 = V.sim != SIM_DOT ? __ldg(V.norms + y0) : 0.0f, n1 = (two && V.sim != SIM_DOT) ? __ldg(V.norms + y1) : 0.0f;
            float ab0, ab1;
            if (two) warp_dot2_t<NG>(reinterpret_cast<const float4*>(V.vecs + (size_t)y0 * V.ld), reinterpret_cast<const float4*>(V.vecs + (size_t)y1 * V.ld),
                                     reinterpret_cast<const float4*>(c.qvec), ng, lane, ab0, ab1);
            else { ab0 = warp_dot_t<NG>(reinterpret_cast<const float4*>(V.vecs + (size_t)y0 * V.ld), reinterpret_cast<const float4*>(c.qvec), ng, lane); ab1 = 0.0f; }
            if (lane == 0) { finish(j, y0, ab0, n0); if (two) finish(j + W, y1, ab1, n1); }
        }
    } else {
        for (int j = warp; j < ntodo; j += W) {
            uint32_t y = c.todo_id[j];
            float vnorm = V.sim != SIM_DOT ? __ldg(V.norms + y) : 0.0f;
            float ab = warp_dot_t<NG>(reinterpret_cast<const float4*>(V.vecs + (size_t)y * V.ld), reinterpret_cast<const float4*>(c.qvec), ng, lane);
            if (lane == 0) finish(j, y, ab, vnorm);
        }
    }
    if (warp == W - 1) cp_async_commit_wait_all();
    c.hop++;
    __syncthreads();
}

// Merge the admitted todo keys into the sorted list A -> B (rank merge, no sort), keep at most `cap`.
// CU: entry 0 (the popped candidate) is dropped.  Afterwards A/B are swapped and s_len/s_best updated.
// LEAN: `len_io` / `best_io` carry the list length and the next candidate in registers (every thread computes them); no barrier
// after thread 0's update of the shared copies, which only code outside the hot loop reads (after a barrier of its own).
template <bool CU, bool LEAN = false>
__device__ inline void hs_merge(SearchCtx& c, int cap, int best, int* len_io = nullptr, int* best_io = nullptr) {
    int len = LEAN ? *len_io : *c.s_len, ntodo = *c.s_ntodo;
    int first = CU ? 1 : 0;
    int my_best = INT_MAX;
    const uint64_t maxtodo = *c.s_maxtodo;
    const int nadmit = *c.s_nadmit;         // final since hs_expand's last barrier; re-zeroed by the next expansion
    int* const bn = c.s_best_next;
    for (int t = threadIdx.x; t < len - first + ntodo; t += blockDim.x) {
        uint64_t key;
        int p;
        if (t < len - first) {
            int i = t + first;
            key = c.A[i];
            if (!CU && i == best) key &= ~1ull;  // the entry just expanded
            int shift = 0;
            if (key < maxtodo)                      // entries above every admitted key keep their position
                for (int j = 0; j < ntodo; ++j) shift += (c.todo_key[j] > key);
            p = t + shift;
        } else {
            int j = t - (len - first);
            key = c.todo_key[j];
            if (key == 0) continue;
            int r = 0;
            for (int jj = 0; jj < ntodo; ++jj) r += (c.todo_key[jj] > key);
            int lo = first, hi = len;  // count of old keys greater than key (A sorted descending)
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (c.A[mid] > key) lo = mid + 1; else hi = mid;
            }
            p = r + (lo - first);
        }
        if (p < cap) {
            c.B[p] = key;
            if (key & 1ull) my_best = min(my_best, p);
        }
    }
    // first unexpanded entry of the merged list: one shared-memory atomic per warp, not per entry
    my_best = __reduce_min_sync(0xFFFFFFFFu, my_best);
    if ((threadIdx.x & 31) == 0 && my_best != INT_MAX) atomicMin(bn, my_best);
    __syncthreads();
    int nl = len - first + nadmit;
    const bool over = nl > cap;
    if (over) nl = cap;
    const int nb = *bn;
    if (threadIdx.x == 0) {
        if (over && CU) c.n_overflow += (1ull << 32);
        *c.s_len = nl;
        *c.s_best = nb;
    }
    uint64_t* t = c.A; c.A = c.B; c.B = t;
    if (LEAN) { *len_io = nl; *best_io = nb; }
    else __syncthreads();
}

// hnsw/search.rs:242-304 on the list held in shared memory.
template <int NG, int W = HS_WARPS, bool PAIR = false>
__device__ inline void hs_layer_search(const VecDev& V, const GraphDev& G, SearchCtx& c, int layer, int ef) {
    int best = *c.s_best, len = *c.s_len;     // published by hs_reseed (behind its barrier); from here on in registers
    while (best < len) {
        uint64_t ckey = c.A[best];
        hs_expand<false, NG, W, PAIR, true>(V, G, c, key_id(ckey), layer, ef, 0.0f, best, len);
        hs_merge<false, true>(c, ef, best, &len, &best);
    }
    c.s_best_next = c.s_bn;                    // (callers that use the shared copies come after a barrier)
}

// NodeFilter::passes (search.rs:147-170) for the popped candidate; warp 0 only, result broadcast by the caller.
__device__ inline bool hs_passes(const VecDev& V, const SearchArgs& a, uint32_t node, float score, const uint32_t* acc_ids,
                                 const float* acc_scores, int nacc, int lane) {
    uint32_t p = V.paragraph_of ? V.paragraph_of[node] : node;
    if (a.filter && !((a.filter[p >> 6] >> (p & 63)) & 1)) return false;
    if (!a.with_duplicates) {  // RepCounter: exact byte equality with an accepted vector (search.rs:388-412)
        const float4* x = reinterpret_cast<const float4*>(V.vecs + (size_t)node * V.ld);
        for (int i = 0; i < nacc; ++i) {
            if (__float_as_uint(acc_scores[i]) != __float_as_uint(score)) continue;  // equal bytes => equal score
            const float4* y = reinterpret_cast<const float4*>(V.vecs + (size_t)acc_ids[i] * V.ld);
            bool same = true;
            for (int g = lane; g < (V.ld >> 2); g += 32) {
                float4 u = x[g], w = y[g];
                same = same && __float_as_uint(u.x) == __float_as_uint(w.x) && __float_as_uint(u.y) == __float_as_uint(w.y) &&
                       __float_as_uint(u.z) == __float_as_uint(w.z) && __float_as_uint(u.w) == __float_as_uint(w.w);
            }
            if (__all_sync(0xFFFFFFFFu, same)) return false;
        }
    }
    if (a.multi_vector) {
        for (int i = 0; i < nacc; ++i) {
            uint32_t ap = V.paragraph_of ? V.paragraph_of[acc_ids[i]] : acc_ids[i];
            if (ap == p) return false;
        }
    }
    return true;
}

// hnsw/search.rs:188-240.  Results go straight to out_ids/out_scores (already descending).
template <int NG, int W = HS_WARPS, bool PAIR = false>
__device__ inline int hs_closest_up(const VecDev& V, const GraphDev& G, SearchCtx& c, const SearchArgs& a, uint32_t* out_ids, float* out_scores) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    hs_reseed(c);
    int nacc = 0;
    while (true) {
        int len = *c.s_len;
        if (len == 0) break;
        uint64_t ckey = c.A[0];
        float score = key_score(ckey);
        uint32_t node = key_id(ckey);
        if (score < a.min_score) break;  // 206
        __syncthreads();
        if (warp == 0) {
            bool pass = !(score != score) && hs_passes(V, a, node, score, out_ids, out_scores, nacc, lane);
            if (lane == 0) {
                *c.s_flag = pass;
                if (pass) { out_ids[nacc] = node; out_scores[nacc] = score; }
            }
        }
        __syncthreads();
        nacc += *c.s_flag;
        if (nacc == a.k) break;  // 214
        hs_expand<true, NG, W, PAIR>(V, G, c, node, 0, 0, a.min_score, 0);
        hs_merge<true>(c, a.cu_cap, 0);
    }
    return nacc;
}

// The tail of HnswSearcher::search for query q: closest_up_nodes on the list in c.A (search.rs:369-375), the final stable sort
// (search.rs:381) and the NIL padding of the outputs.
template <int NG, int W = HS_WARPS, bool PAIR = false>
__device__ inline void hs_emit_results(const VecDev& V, const GraphDev& G, SearchCtx& c, const SearchArgs& a, unsigned int q) {
    uint32_t* oi = a.out_ids + (size_t)q * a.k;
    float* os = a.out_scores + (size_t)q * a.k;
    int nacc = hs_closest_up<NG, W, PAIR>(V, G, c, a, oi, os);
    __syncthreads();
    // search.rs:381 `filtered_result.sort_by(|a, b| b.1.total_cmp(&a.1))`: stable, descending.
    // (closest_up_nodes can accept a late-found neighbour that outranks earlier results.)
    {
        uint32_t* tid = reinterpret_cast<uint32_t*>(c.B);
        float* tsc = reinterpret_cast<float*>(c.B) + a.k;
        for (int i = threadIdx.x; i < nacc; i += blockDim.x) { tid[i] = oi[i]; tsc[i] = os[i]; }
        __syncthreads();
        for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
            uint32_t oi_bits = ordered_bits(tsc[i]);
            int r = 0;
            for (int j = 0; j < nacc; ++j) {
                uint32_t oj = ordered_bits(tsc[j]);
                r += (oj > oi_bits) || (oj == oi_bits && j < i);
            }
            oi[r] = tid[i];
            os[r] = tsc[i];
        }
        __syncthreads();
    }
    for (int i = nacc + threadIdx.x; i < a.k; i += blockDim.x) { oi[i] = NIL; os[i] = 0.0f; }
    if (threadIdx.x == 0) a.out_counts[q] = nacc;
}

// W warps per CTA.  W = 8: one row per warp in flight, 4 CTAs per SM (592 queries resident).  W = 4 (PAIR): two rows per warp in
// flight, 7 CTAs per SM -- 1036 queries resident, so a batch of 1024 runs as ONE wave instead of 592 + 432 (the second wave of
// the 8-warp shape leaves 27 % of the CTA slots empty while it runs).
template <int NG, int W = HS_WARPS, bool PAIR = false>
__global__ void __launch_bounds__(W * 32, W == HS_WARPS ? 4 : 7) hnsw_search_kernel(VecDev V, GraphDev G, SearchArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_ints[8];
    __shared__ unsigned int s_work;
    __shared__ unsigned long long s_maxtodo;
    __shared__ int s_bn[2];
    SearchCtx c;
    c.s_bn = s_bn;
    unsigned char* p = smem;
    c.qvec = reinterpret_cast<float*>(p); p += (size_t)V.ld * 4;
    c.A = reinterpret_cast<uint64_t*>(p); p += (size_t)a.list_cap * 8;
    c.B = reinterpret_cast<uint64_t*>(p); p += (size_t)a.list_cap * 8;
    c.todo_key = reinterpret_cast<uint64_t*>(p); p += HS_MAX_ROW * 8;
    c.hash = reinterpret_cast<uint32_t*>(p); p += (size_t)4 << a.hash_bits;
    c.todo_id = reinterpret_cast<uint32_t*>(p); p += HS_MAX_ROW * 4;
    c.pref_row = reinterpret_cast<uint32_t*>(p); p += 2 * HS_MAX_ROW * 4;
    c.pref_node = reinterpret_cast<uint32_t*>(p);
    c.hop = 0;
    c.s_len = &s_ints[0]; c.s_best = &s_ints[1]; c.s_best_next = &s_bn[0]; c.s_ntodo = &s_ints[3];
    c.s_hash_count = &s_ints[4]; c.s_flag = &s_ints[5]; c.s_nadmit = &s_ints[6];
    c.s_maxtodo = &s_maxtodo;
    c.hash_bits = a.hash_bits;
    c.hash_mask = (1u << a.hash_bits) - 1;
    c.hash_limit = (int)((15u << a.hash_bits) >> 4) - HS_MAX_ROW;
    c.n_dist = c.n_expand = c.n_overflow = 0;
    int lane = threadIdx.x & 31;
    int ng = V.ld >> 2;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_work = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        unsigned int q = s_work;
        if (q >= (unsigned)a.nq) break;

        const float* qsrc;
        uint32_t self = NIL;
        if (a.mode == 0) { qsrc = a.queries + (size_t)q * V.ld; c.qnorm = V.sim != SIM_DOT ? a.qnorms[q] : 0.0f; }
        else { self = a.nodes[q]; qsrc = V.vecs + (size_t)self * V.ld; c.qnorm = V.sim != SIM_DOT ? V.norms[self] : 0.0f; }
        for (int i = threadIdx.x; i < ng; i += blockDim.x) reinterpret_cast<float4*>(c.qvec)[i] = reinterpret_cast<const float4*>(qsrc)[i];
        __syncthreads();

        // entry point: similarity + single-entry list (search.rs:256-261)
        if (threadIdx.x < 32) {
            uint32_t ep = G.entry_node;
            float ab = warp_dot_t<NG>(reinterpret_cast<const float4*>(V.vecs + (size_t)ep * V.ld), reinterpret_cast<const float4*>(c.qvec), ng, lane);
            if (lane == 0) {
                c.A[0] = make_key(finish_similarity(V, ab, ep, c.qnorm), ep, 1);
                *c.s_len = 1;
                c.n_dist++;
            }
        }
        __syncthreads();

        int top = a.mode == 1 ? (int)G.level[self] : -1;
        for (int layer = (int)G.entry_layer; layer >= 0; --layer) {
            int ef;
            if (a.mode == 0) ef = layer == 0 ? a.ef0 : 1;
            else ef = layer <= top ? a.efC : 1;
            hs_reseed(c);
            hs_layer_search<NG, W, PAIR>(V, G, c, layer, ef);
            __syncthreads();
            if (a.mode == 1 && layer <= top && layer < HS_MAX_LAYERS) {
                int len = *c.s_len;
                uint64_t* dst = a.found + ((size_t)q * HS_MAX_LAYERS + layer) * a.efC;
                for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = c.A[i];
                if (threadIdx.x == 0) a.found_count[(size_t)q * HS_MAX_LAYERS + layer] = len;
            }
            // next layer's entry points: the list as it stands -- a single node above the insertion
            // layers and for queries (ef == 1, search.rs:323-328), all efC results once inside the
            // node's layers (build.rs:142-150).  Scores to the same query do not change, so they are
            // not recomputed; hs_reseed() marks them unexpanded and seeds the visited set.
            __syncthreads();
        }

        // a node that rises above the current top layer (only on graph reuse, nidx_vec_extend_hnsw) finds nothing up there
        if (a.mode == 1 && threadIdx.x == 0)
            for (int layer = (int)G.entry_layer + 1; layer <= top && layer < HS_MAX_LAYERS; ++layer) a.found_count[(size_t)q * HS_MAX_LAYERS + layer] = 0;

        if (a.mode == 0) hs_emit_results<NG, W, PAIR>(V, G, c, a, q);
    }
    // counters: n_dist lives in lane 0 of every warp, the rest in thread 0
    if (lane == 0 && c.n_dist) atomicAdd(&a.counters[0], c.n_dist);
    if (threadIdx.x == 0) {
        if (c.n_expand) atomicAdd(&a.counters[1], c.n_expand);
        if (c.n_overflow & 0xFFFFFFFFull) atomicAdd(&a.counters[2], c.n_overflow & 0xFFFFFFFFull);
        if (c.n_overflow >> 32) atomicAdd(&a.counters[3], c.n_overflow >> 32);
    }
}

}  // namespace nidx
