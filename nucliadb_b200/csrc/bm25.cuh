// nidx_b200 — K7: BM25 top-k over device-resident postings (sm_100a).
//
// Replaces the tantivy collector call of the reference:
//   nidx/nidx_text/src/reader.rs:432-435        TopDocs::with_limit(k+1).order_by_score() + Count
//   nidx/nidx_paragraph/src/reader.rs:290-292   same, OR of TermQuery(Basic) (keyword_parser.rs:62-67)
// with tantivy 0.26's BM25 (restated in oracle/bm25.hpp; parity unpinned):
//   score(doc) = sum over matching query terms of  idf_t * (1 + k1) * tf / (tf + k1 * (1 - b + b * fieldnorm / avg))
//
// One CTA per query walks the doc-id space in tiles of `tile` documents.  Per tile
//   1. every query term's posting cursor is advanced to the tile end (gallop + binary search);
//   2. pass 1: the tile's postings, flattened over terms, are scored and accumulated into a shared-memory
//      accumulator with integer atomics.  Contributions are converted to fixed point (2^-shift) so the sum
//      is independent of the order the atomics land in: equal scores stay bit-equal, which keeps the
//      (score desc, doc asc) tie order of TopDocs deterministic;
//   3. pass 2: the same postings are walked again; atomicExch(acc, 0) hands each touched document to
//      exactly one thread (and leaves the accumulator clean for the next tile), which offers it to a
//      block-wide streaming top-k.  No dense clear, no dense scan: work is proportional to postings.
// HBM traffic = the query's postings once (doc id + tf, 8 B; 4 B when tf is not needed) + 1 B fieldnorm
// gather per posting; everything else stays in shared memory / L1.
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace nidx {

constexpr int BM_THREADS = 256;
constexpr int BM_MAX_TERMS = 128;

struct TxtDev {
    uint32_t n_docs, n_terms;
    const uint64_t* term_off;
    const uint32_t* post_doc;
    const uint32_t* post_tf;
    const unsigned char* fieldnorm;
    const uint64_t* alive;
};

struct Bm25Args {
    const uint32_t* query_terms;
    const uint32_t* query_off;
    int nq;
    int mode, use_tf, k, cap, tile;
    const float* term_weight;   // [n_terms] idf * (1 + k1) from the collection statistics
    const float* norm_cache;    // [256] k1 * (1 - b + b * fieldnorm(id) / avg)
    int shift;                  // fixed point: 2^-shift
    uint64_t* out_keys;         // [nq][k] rank keys (score desc, doc asc), 0 = none
    unsigned long long* out_total;  // [nq] matching documents (Count collector)
};

__host__ __device__ __forceinline__ size_t bm_smem_bytes(int tile, int cap) {
    return (size_t)tile * 4 + (size_t)tile + (size_t)cap * 8 + BM_MAX_TERMS * (8 + 8 + 4 + 4) + 1024 + 64;
}

__global__ void __launch_bounds__(BM_THREADS) bm25_kernel(TxtDev T, Bm25Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    __shared__ int s_total;
    __shared__ unsigned long long s_hits;
    unsigned char* p = smem;
    uint64_t* tk_buf = reinterpret_cast<uint64_t*>(p); p += (size_t)a.cap * 8;
    uint64_t* cur = reinterpret_cast<uint64_t*>(p); p += BM_MAX_TERMS * 8;     // cursor per term (absolute posting index)
    uint64_t* tend = reinterpret_cast<uint64_t*>(p); p += BM_MAX_TERMS * 8;    // end of the term's postings
    uint32_t* acc = reinterpret_cast<uint32_t*>(p); p += (size_t)a.tile * 4;
    float* ncache = reinterpret_cast<float*>(p); p += 1024;
    int* pre = reinterpret_cast<int*>(p); p += BM_MAX_TERMS * 4;               // exclusive prefix of per-term counts in the tile
    float* tw = reinterpret_cast<float*>(p); p += BM_MAX_TERMS * 4;
    uint32_t* cnt32 = reinterpret_cast<uint32_t*>(p);                          // [tile/4] packed byte counters (AND)

    int q = blockIdx.x;
    const uint32_t* terms = a.query_terms + a.query_off[q];
    int nt = (int)(a.query_off[q + 1] - a.query_off[q]);
    if (nt > BM_MAX_TERMS) nt = BM_MAX_TERMS;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, a.k, a.cap);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ncache[i] = a.norm_cache[i];
    for (int i = threadIdx.x; i < a.tile; i += blockDim.x) acc[i] = 0;
    for (int i = threadIdx.x; i < a.tile / 4; i += blockDim.x) cnt32[i] = 0;
    bool missing = false;
    if (threadIdx.x < nt) {
        uint32_t t = terms[threadIdx.x];
        bool ok = t < T.n_terms;
        cur[threadIdx.x] = ok ? T.term_off[t] : 0;
        tend[threadIdx.x] = ok ? T.term_off[t + 1] : 0;
        tw[threadIdx.x] = ok ? a.term_weight[t] : 0.0f;
        missing = !ok || T.term_off[t] == T.term_off[t + 1];
    }
    if (threadIdx.x == 0) s_hits = 0;
    // an AND query with a term that has no postings matches nothing
    int any_missing = __syncthreads_or(missing);
    bool dead = (a.mode == 1 && any_missing) || nt == 0;
    const float scale = (float)(1u << a.shift);

    for (uint32_t lo = 0; lo < T.n_docs && !dead; lo += a.tile) {
        uint32_t hi = lo + a.tile < T.n_docs ? lo + a.tile : T.n_docs;
        // 1. advance cursors to the first posting with doc >= hi
        uint64_t my_begin = 0, my_end = 0;
        if (threadIdx.x < nt) {
            uint64_t b = cur[threadIdx.x], e = tend[threadIdx.x];
            my_begin = b;
            uint64_t step = 32, l = b, r = e;
            while (l + step < e && T.post_doc[l + step] < hi) { l += step; step <<= 1; }  // gallop
            r = l + step < e ? l + step : e;
            // invariant: every posting before l is < hi (or l == b); first posting >= hi lies in [l, r]
            while (l < r) {
                uint64_t m = (l + r) >> 1;
                if (T.post_doc[m] < hi) l = m + 1; else r = m;
            }
            my_end = l;
            pre[threadIdx.x] = (int)(my_end - my_begin);
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // tiny exclusive scan (nt <= 128)
            int run = 0;
            for (int t = 0; t < nt; ++t) { int c = pre[t]; pre[t] = run; run += c; }
            s_total = run;
        }
        __syncthreads();
        int total = s_total;
        // 2. pass 1: accumulate
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            int l = 0, r = nt - 1;  // last term with pre[t] <= i
            while (l < r) { int m = (l + r + 1) >> 1; if (pre[m] <= i) l = m; else r = m - 1; }
            uint64_t pi = cur[l] + (uint64_t)(i - pre[l]);
            uint32_t d = T.post_doc[pi];
            uint32_t tf = a.use_tf ? T.post_tf[pi] : 1u;
            float tff = (float)tf;
            float s = __fmul_rn(tw[l], __fdiv_rn(tff, __fadd_rn(tff, ncache[T.fieldnorm[d]])));
            uint32_t fx = (uint32_t)__float2uint_rn(__fmul_rn(s, scale));
            if (fx == 0) fx = 1;
            atomicAdd(&acc[d - lo], fx);
            if (a.mode == 1) atomicAdd(&cnt32[(d - lo) >> 2], 1u << (8 * ((d - lo) & 3)));
        }
        __syncthreads();
        // 3. pass 2: claim + offer (lock-step rounds, BlockTopK::offer synchronises)
        for (int base = 0; base < total; base += blockDim.x) {
            int i = base + threadIdx.x;
            uint64_t key = 0;
            if (i < total) {
                int l = 0, r = nt - 1;
                while (l < r) { int m = (l + r + 1) >> 1; if (pre[m] <= i) l = m; else r = m - 1; }
                uint64_t pi = cur[l] + (uint64_t)(i - pre[l]);
                uint32_t d = T.post_doc[pi];
                uint32_t v = atomicExch(&acc[d - lo], 0u);
                if (v != 0) {
                    bool match = true;
                    if (a.mode == 1) {
                        uint32_t sh = 8 * ((d - lo) & 3);
                        uint32_t c = (atomicAnd(&cnt32[(d - lo) >> 2], ~(0xFFu << sh)) >> sh) & 0xFFu;
                        match = (int)c == nt;
                    }
                    if (match && T.alive) match = (T.alive[d >> 6] >> (d & 63)) & 1;
                    if (match) {
                        atomicAdd(&s_hits, 1ull);
                        key = make_key(__fdiv_rn((float)v, scale), d, 0);
                    }
                }
            }
            tk.offer(key);
        }
        __syncthreads();
        if (threadIdx.x < nt) cur[threadIdx.x] = my_end;
        __syncthreads();
    }
    int c = tk.finish();
    uint64_t* out = a.out_keys + (size_t)q * a.k;
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) out[i] = i < c ? tk_buf[i] : 0;
    if (threadIdx.x == 0 && a.out_total) a.out_total[q] = s_hits;
}

// keys -> (doc, score, count) with the min_score cut applied after top-k (reader.rs:302-305).
__global__ void bm25_finish_kernel(const uint64_t* keys, int nq, int k, float min_score, uint32_t* out_docs, float* out_scores, int* out_counts) {
    int q = blockIdx.x;
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = keys[(size_t)q * k + i];
        bool ok = key != 0 && !(key_score(key) < min_score);
        out_docs[(size_t)q * k + i] = ok ? key_id(key) : NIL;
        out_scores[(size_t)q * k + i] = ok ? key_score(key) : 0.0f;
        if (ok) atomicAdd(&s_count, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = s_count;
}

}  // namespace nidx
