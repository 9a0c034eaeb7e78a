// nidx_b200 — K7: BM25 top-k over device-resident postings (sm_100a).
//
// Replaces the tantivy collector call of the reference:
//   nidx/nidx_text/src/reader.rs:432-435        TopDocs::with_limit(k+1).order_by_score() + Count
//   nidx/nidx_paragraph/src/reader.rs:290-292   same, OR of TermQuery(Basic) (keyword_parser.rs:62-67)
// with tantivy 0.26's BM25 (restated in oracle/bm25.hpp; parity unpinned):
//   score(doc) = sum over matching query terms of  idf_t * (1 + k1) * tf / (tf + k1 * (1 - b + b * fieldnorm / avg))
//
// Index-time layout (bm25_prepare_* kernels, once per segment):
//   * post_tfn[i] = tf << 8 | fieldnorm_id(doc): the document's length code travels with the posting, so
//     scoring never gathers fieldnorm[doc] at random (one 32-byte sector per posting otherwise);
//   * skip[row][t] = first posting of the term with doc >= t * BM_TILE, for every term with df >= BM_SKIP_DF:
//     a query term's slice of a doc tile is two adjacent loads instead of a search.
// Query time: one CTA per query walks the doc-id space in tiles of BM_TILE documents.  Per tile
//   1. per-term slice bounds (skip table; gallop + binary search for rare terms), prefetched one tile ahead;
//   2. pass 1: the tile's postings, flattened over terms, are scored and accumulated into a shared-memory
//      accumulator with integer atomics.  Contributions are fixed point (2^-shift), so the sum does not
//      depend on the order the atomics land in: equal scores stay bit-equal and the (score desc, doc asc)
//      tie order of TopDocs is deterministic;
//   3. pass 2: the same postings again; atomicExch(acc, 0) hands each touched document to exactly one thread
//      (and leaves the accumulator clean for the next tile), which appends it to a streaming top-k buffer if
//      it beats the running threshold.  Work is proportional to postings: no dense clear, no dense scan.
// HBM traffic = the query's postings once (8 B each); everything else stays in shared memory / L1 / L2.
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace nidx {

constexpr int BM_THREADS = 256;
constexpr int BM_MAX_TERMS = 128;
constexpr int BM_TILE = 12288;     // documents per tile (48 KB of u32 accumulators; 3 CTAs per SM)
constexpr int BM_SKIP_DF = 32;     // terms with at least this many postings get a skip row
constexpr int BM_ROUND = 2;        // postings per thread in flight / between two top-k capacity checks
constexpr int BM_TOUCH_CAP = 4096; // documents hit per tile that are tracked individually (else dense scan)

struct TxtDev {
    uint32_t n_docs, n_terms, n_tiles;
    const uint64_t* term_off;
    const uint32_t* post_doc;
    const uint32_t* post_tfn;        // tf << 8 | fieldnorm id
    const uint32_t* skip_row;        // [n_terms] row in skip[] or NIL
    const uint32_t* skip;            // [rows][n_tiles + 1] posting index relative to term_off[term]
    const uint64_t* alive;
};

struct Bm25Args {
    const uint32_t* query_terms;
    const uint32_t* query_off;
    int nq;
    int mode, use_tf, k, cap;
    const float* term_weight;   // [n_terms] idf * (1 + k1) from the collection statistics
    const float* norm_cache;    // [256] k1 * (1 - b + b * fieldnorm(id) / avg)
    int shift;                  // fixed point: 2^-shift
    int after_mode;             // search-after (nidx_paragraph reader.rs:379-392): 0 none, 1 Drop, 2 KeepAfter, 3 Keep
    float after_score;
    uint64_t after_docaddr, docaddr_base;
    uint64_t* out_keys;         // [nq][k] rank keys (score desc, doc asc), 0 = none
    unsigned long long* out_total;  // [nq] matching documents (Count collector)
};

// ---- index-time kernels ---------------------------------------------------------------------------------
__global__ void bm25_pack_tfn_kernel(const uint32_t* __restrict__ post_doc, const uint32_t* __restrict__ post_tf, const unsigned char* __restrict__ fieldnorm,
                                     uint64_t n_post, uint32_t* __restrict__ post_tfn) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t tf = post_tf ? post_tf[i] : 1u;
        if (tf > 0xFFFFFFu) tf = 0xFFFFFFu;
        post_tfn[i] = (tf << 8) | fieldnorm[post_doc[i]];
    }
}
// one thread per (skip row, tile boundary)
__global__ void bm25_build_skip_kernel(const uint64_t* __restrict__ term_off, const uint32_t* __restrict__ post_doc, const uint32_t* __restrict__ row_term,
                                       uint32_t n_rows, uint32_t n_tiles, uint32_t* __restrict__ skip) {
    uint64_t total = (uint64_t)n_rows * (n_tiles + 1);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t row = (uint32_t)(i / (n_tiles + 1)), t = (uint32_t)(i % (n_tiles + 1));
        uint32_t term = row_term[row];
        uint64_t b = term_off[term], e = term_off[term + 1];
        uint64_t bound = (uint64_t)t * BM_TILE;
        uint64_t l = b, r = e;
        while (l < r) {
            uint64_t m = (l + r) >> 1;
            if ((uint64_t)post_doc[m] < bound) l = m + 1; else r = m;
        }
        skip[i] = (uint32_t)(l - b);
    }
}

__host__ __device__ __forceinline__ size_t bm_smem_bytes(int cap, bool conj) {
    return (size_t)BM_TILE * 4 + (conj ? (size_t)BM_TILE : 0) + (size_t)cap * 8 + (size_t)BM_TOUCH_CAP * 2 +
           BM_MAX_TERMS * (8 + 8 + 2 * 8 + 2 * 4 + 4 + 4 + 4) + 1024 + 64;
}

// Software pipeline over tiles: while tile t is accumulated and collected out of registers / shared memory, the
// slices of tile t+1 are resolved (skip entries were requested two tiles earlier) and its postings are already in
// flight from HBM, so no global-memory latency sits on the per-tile critical path.
__global__ void __launch_bounds__(BM_THREADS) bm25_kernel(TxtDev T, Bm25Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    __shared__ int s_total[2], s_min_count[2], s_ntouched[2];   // touched counters alternate with the tile parity
    __shared__ unsigned long long s_hits;
    unsigned char* p = smem;
    uint64_t* tk_buf = reinterpret_cast<uint64_t*>(p); p += (size_t)a.cap * 8;
    uint64_t* tbase = reinterpret_cast<uint64_t*>(p); p += BM_MAX_TERMS * 8;       // term_off[term]
    uint64_t* tend = reinterpret_cast<uint64_t*>(p); p += BM_MAX_TERMS * 8;        // term_off[term + 1]
    uint64_t* cur = reinterpret_cast<uint64_t*>(p); p += 2 * BM_MAX_TERMS * 8;     // [2][terms] first posting of the tile (absolute)
    uint32_t* acc = reinterpret_cast<uint32_t*>(p); p += (size_t)BM_TILE * 4;
    float* ncache = reinterpret_cast<float*>(p); p += 1024;
    int* pre = reinterpret_cast<int*>(p); p += 2 * BM_MAX_TERMS * 4;               // [2][terms] exclusive prefix of per-term counts
    float* tw = reinterpret_cast<float*>(p); p += BM_MAX_TERMS * 4;
    uint32_t* srow = reinterpret_cast<uint32_t*>(p); p += BM_MAX_TERMS * 4;        // skip row or NIL
    int* cnt_t = reinterpret_cast<int*>(p); p += BM_MAX_TERMS * 4;
    unsigned short* touched = reinterpret_cast<unsigned short*>(p); p += (size_t)BM_TOUCH_CAP * 2;   // tile-relative ids of the docs hit in this tile
    unsigned char* cnt8 = p;                                                       // [BM_TILE] matched-term counters (AND only)

    int q = blockIdx.x;
    const uint32_t* terms = a.query_terms + a.query_off[q];
    int nt = (int)(a.query_off[q + 1] - a.query_off[q]);
    if (nt > BM_MAX_TERMS) nt = BM_MAX_TERMS;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, a.k, a.cap);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ncache[i] = a.norm_cache[i];
    for (int i = threadIdx.x; i < BM_TILE; i += blockDim.x) acc[i] = 0;
    if (a.mode == 1) for (int i = threadIdx.x; i < BM_TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(cnt8)[i] = 0;
    bool missing = false;
    unsigned int my_hits = 0;  // matching documents claimed by this thread (one shared atomic per warp at the end, not per hit)
    uint64_t my_next = 0;      // first posting of the next unresolved tile for my term
    uint32_t pf_end = 0;       // skip entry of the next unresolved tile's end, requested one resolve step ahead
    size_t my_skip = 0;
    if (threadIdx.x < nt) {
        uint32_t t = terms[threadIdx.x];
        bool ok = t < T.n_terms;
        uint64_t b = ok ? T.term_off[t] : 0, e = ok ? T.term_off[t + 1] : 0;
        tbase[threadIdx.x] = b;
        tend[threadIdx.x] = e;
        tw[threadIdx.x] = ok ? a.term_weight[t] : 0.0f;
        uint32_t row = ok ? T.skip_row[t] : NIL;
        srow[threadIdx.x] = row;
        missing = b == e;
        my_next = b;
        if (row != NIL) { my_skip = (size_t)row * (T.n_tiles + 1); pf_end = T.skip[my_skip + 1]; }
    }
    if (threadIdx.x == 0) { s_hits = 0; s_ntouched[0] = 0; s_ntouched[1] = 0; }
    int any_missing = __syncthreads_or(missing);   // an AND query with a term without postings matches nothing
    bool dead = (a.mode == 1 && any_missing) || nt == 0;
    const float scale = (float)(1u << a.shift);
    const int lane = threadIdx.x & 31;
    const uint32_t n_tiles = dead ? 0 : T.n_tiles;

    // resolve(tile, buf): slices of every term in `tile` -> cur[buf], pre[buf], s_total[buf], s_min_count[buf]
    auto resolve = [&](uint32_t tile, int buf) {
        uint32_t hi = (tile + 1) * BM_TILE < T.n_docs ? (tile + 1) * BM_TILE : T.n_docs;
        if (threadIdx.x < nt) {
            uint64_t bgn = my_next, e = tend[threadIdx.x], end;
            if (srow[threadIdx.x] != NIL) {
                end = tbase[threadIdx.x] + pf_end;
                if (tile + 2 <= T.n_tiles) pf_end = T.skip[my_skip + tile + 2];
            } else {                                                       // rare term: a few postings in total
                uint64_t l = bgn;
                while (l < e && T.post_doc[l] < hi) ++l;
                end = l;
            }
            cur[buf * BM_MAX_TERMS + threadIdx.x] = bgn;
            cnt_t[threadIdx.x] = (int)(end - bgn);
            my_next = end;
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // exclusive scan of the per-term counts by one warp (nt <= 128)
            int run = 0, mn = INT_MAX;
            for (int t0 = 0; t0 < nt; t0 += 32) {
                int t = t0 + threadIdx.x;
                int v = t < nt ? cnt_t[t] : 0;
                if (t < nt && v < mn) mn = v;
                int x = v;
                for (int off = 1; off < 32; off <<= 1) { int y = __shfl_up_sync(0xFFFFFFFFu, x, off); if ((int)threadIdx.x >= off) x += y; }
                if (t < nt) pre[buf * BM_MAX_TERMS + t] = run + x - v;
                run += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
            for (int off = 16; off >= 1; off >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, off));
            if (threadIdx.x == 0) { s_total[buf] = run; s_min_count[buf] = mn; }
        }
        __syncthreads();
    };
    // fetch(buf): this thread's first BM_ROUND postings of the tile resolved in `buf` -> registers (loads in flight)
    uint32_t d_n[BM_ROUND], tfn_n[BM_ROUND], d_c[BM_ROUND], tfn_c[BM_ROUND];
    int tl_n[BM_ROUND], tl_c[BM_ROUND];
    auto locate = [&](int buf, int i, int& l) -> uint64_t {
        const int* pr = pre + buf * BM_MAX_TERMS;
        int lo_ = 0, r = nt - 1;  // last term with pre[t] <= i
        while (lo_ < r) { int m = (lo_ + r + 1) >> 1; if (pr[m] <= i) lo_ = m; else r = m - 1; }
        l = lo_;
        return cur[buf * BM_MAX_TERMS + lo_] + (uint64_t)(i - pr[lo_]);
    };
    auto fetch = [&](int buf) {
        int total = s_total[buf];
        bool skip_tile = total == 0 || (a.mode == 1 && s_min_count[buf] == 0);
#pragma unroll
        for (int u = 0; u < BM_ROUND; ++u) {
            int i = u * BM_THREADS + threadIdx.x;
            tl_n[u] = -1;
            if (!skip_tile && i < total) {
                int l;
                uint64_t pi = locate(buf, i, l);
                d_n[u] = __ldg(T.post_doc + pi);
                tfn_n[u] = __ldg(T.post_tfn + pi);
                tl_n[u] = l;
            }
        }
    };
    auto accumulate = [&](uint32_t lo, uint32_t d, uint32_t tfn, int l, int par) {
        bool first = false;
        uint32_t off = 0;
        if (l >= 0) {
            float tff = a.use_tf ? (float)(tfn >> 8) : 1.0f;
            float s = __fmul_rn(tw[l], __fdiv_rn(tff, __fadd_rn(tff, ncache[tfn & 0xFFu])));
            uint32_t fx = (uint32_t)__float2uint_rn(__fmul_rn(s, scale));
            if (fx == 0) fx = 1;
            off = d - lo;
            first = atomicAdd(&acc[off], fx) == 0;   // fx >= 1, so a zero means nobody was here before
            if (a.mode == 1) atomicAdd(reinterpret_cast<uint32_t*>(cnt8) + (off >> 2), 1u << (8 * (off & 3)));
        }
        unsigned m = __ballot_sync(0xFFFFFFFFu, first);   // warp-aggregated append to the touched list
        if (m) {
            int basepos = 0;
            if (lane == 0) basepos = atomicAdd(&s_ntouched[par], __popc(m));
            basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
            int pos = basepos + __popc(m & ((1u << lane) - 1));
            if (first && pos < BM_TOUCH_CAP) touched[pos] = (unsigned short)off;
        }
    };

    if (n_tiles) { resolve(0, 0); fetch(0); }
#pragma unroll
    for (int u = 0; u < BM_ROUND; ++u) { d_c[u] = d_n[u]; tfn_c[u] = tfn_n[u]; tl_c[u] = tl_n[u]; }

    for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        int cb = tile & 1, nb = cb ^ 1;
        uint32_t lo = tile * BM_TILE;
        uint32_t hi = lo + BM_TILE < T.n_docs ? lo + BM_TILE : T.n_docs;
        if (tile + 1 < n_tiles) { resolve(tile + 1, nb); fetch(nb); }   // next tile's postings now in flight
        int total = s_total[cb];
        bool do_tile = !(total == 0 || (a.mode == 1 && s_min_count[cb] == 0));   // AND: some term has nothing in this tile
        if (do_tile) {
            // pass 1: score + accumulate (registers first, then whatever did not fit the prefetch window)
#pragma unroll
            for (int u = 0; u < BM_ROUND; ++u) accumulate(lo, d_c[u], tfn_c[u], tl_c[u], cb);
            for (int base = BM_THREADS * BM_ROUND; base < total; base += BM_THREADS) {
                int i = base + threadIdx.x, l = -1;
                uint32_t d = 0, tfn = 0;
                if (i < total) { uint64_t pi = locate(cb, i, l); d = __ldg(T.post_doc + pi); tfn = __ldg(T.post_tfn + pi); }
                accumulate(lo, d, tfn, l, cb);
            }
            __syncthreads();
            // pass 2: every touched document once -> count, reset, offer to the streaming top-k
            int ntouched = s_ntouched[cb];
            bool dense = ntouched > BM_TOUCH_CAP;      // list overflow: fall back to scanning the whole tile
            int work = dense ? (int)(hi - lo) : ntouched;
            for (int base = 0; base < work; base += BM_THREADS * BM_ROUND) {
#pragma unroll
                for (int u = 0; u < BM_ROUND; ++u) {
                    int j = base + u * BM_THREADS + threadIdx.x;
                    if (j < work) {
                        uint32_t off = dense ? (uint32_t)j : (uint32_t)touched[j];
                        uint32_t v = acc[off];
                        if (v != 0) {
                            acc[off] = 0;
                            bool match = true;
                            if (a.mode == 1) { match = (int)cnt8[off] == nt; cnt8[off] = 0; }
                            uint32_t doc = lo + off;
                            if (match && T.alive) match = (T.alive[doc >> 6] >> (doc & 63)) & 1;
                            if (match) {
                                my_hits++;
                                float score = __fdiv_rn((float)v, scale);
                                bool after = true;   // is_after(): strictly lower score, or an equal score that the tie break keeps
                                if (a.after_mode != 0) {
                                    uint32_t so = ordered_bits(score), ao = ordered_bits(a.after_score);
                                    after = so < ao || (so == ao && (a.after_mode == 3 || (a.after_mode == 2 && a.docaddr_base + doc > a.after_docaddr)));
                                }
                                uint64_t key = make_key(score, doc, 0);
                                if (after && key > tk_thr) tk_buf[atomicAdd(&tk_count, 1)] = key;
                            }
                        }
                    }
                }
                __syncthreads();
                if (tk_count > a.cap - BM_THREADS * BM_ROUND) tk.flush();
            }
            // reset after everyone has read it (the loop above synchronised at least once iff ntouched > 0); the next
            // user of this parity is tile + 2, behind the barriers of the next iteration's resolve()
            if (threadIdx.x == 0 && ntouched > 0) s_ntouched[cb] = 0;
        }
#pragma unroll
        for (int u = 0; u < BM_ROUND; ++u) { d_c[u] = d_n[u]; tfn_c[u] = tfn_n[u]; tl_c[u] = tl_n[u]; }
    }
    for (int off = 16; off >= 1; off >>= 1) my_hits += __shfl_xor_sync(0xFFFFFFFFu, my_hits, off);
    if (lane == 0 && my_hits) atomicAdd(&s_hits, (unsigned long long)my_hits);
    int c = tk.finish();
    uint64_t* out = a.out_keys + (size_t)q * a.k;
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) out[i] = i < c ? tk_buf[i] : 0;
    if (threadIdx.x == 0 && a.out_total) a.out_total[q] = s_hits;
}

// ---- term-major variant (queries with at most BM_TM_TERMS terms) ---------------------------------------------
// The flattened kernel above spends most of its instructions on bookkeeping that exists only to balance postings
// over threads (per-tile prefix scan, a binary search per posting and pass, four barriers per tile).  With ~12
// postings per (term, tile) a warp per term slice is balanced enough, and everything about a term -- cursor, skip
// prefetch, weight -- can live in the registers of ONE lane of the warp that owns it (term t belongs to warp
// t % 8, lane t / 8), broadcast by shuffle when the slice is processed: no shared-memory cursor arrays, no scan,
// no search, two barriers per tile (accumulate | collect).
constexpr int BM_TM_OWN = 8;                         // terms per warp
constexpr int BM_TM_TERMS = BM_TM_OWN * (BM_THREADS / 32);

__host__ __device__ __forceinline__ size_t bm_tm_smem_bytes(int cap, bool conj) {
    return (size_t)BM_TILE * 4 + (conj ? (size_t)BM_TILE : 0) + (size_t)cap * 8 + (size_t)BM_TOUCH_CAP * 2 + 1024 + 64;
}

__global__ void __launch_bounds__(BM_THREADS) bm25_tm_kernel(TxtDev T, Bm25Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    __shared__ int s_ntouched[2];
    __shared__ unsigned long long s_hits;
    unsigned char* p = smem;
    uint64_t* tk_buf = reinterpret_cast<uint64_t*>(p); p += (size_t)a.cap * 8;
    uint32_t* acc = reinterpret_cast<uint32_t*>(p); p += (size_t)BM_TILE * 4;
    float* ncache = reinterpret_cast<float*>(p); p += 1024;
    unsigned short* touched = reinterpret_cast<unsigned short*>(p); p += (size_t)BM_TOUCH_CAP * 2;
    unsigned char* cnt8 = p;   // [BM_TILE] matched-term counters (AND only)

    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* terms = a.query_terms + a.query_off[q];
    int nt = (int)(a.query_off[q + 1] - a.query_off[q]);
    if (nt > BM_TM_TERMS) nt = BM_TM_TERMS;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, a.k, a.cap);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ncache[i] = a.norm_cache[i];
    for (int i = threadIdx.x; i < BM_TILE; i += blockDim.x) acc[i] = 0;
    if (a.mode == 1) for (int i = threadIdx.x; i < BM_TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(cnt8)[i] = 0;

    // this lane's term (if any): t = warp + 8 * lane
    const int my_t = warp + (BM_THREADS / 32) * lane;
    const bool own = lane < BM_TM_OWN && my_t < nt;
    const int n_own = nt > warp ? (nt - warp + (BM_THREADS / 32) - 1) / (BM_THREADS / 32) : 0;   // terms owned by this warp (warp-uniform)
    uint64_t t_base = 0, t_end = 0, t_next = 0;
    size_t t_skip = 0;
    uint32_t pf_end = 0;
    bool has_skip = false;
    float t_w = 0.0f;
    bool missing = false;
    if (own) {
        uint32_t t = terms[my_t];
        bool ok = t < T.n_terms;
        t_base = ok ? T.term_off[t] : 0;
        t_end = ok ? T.term_off[t + 1] : 0;
        t_w = ok ? a.term_weight[t] : 0.0f;
        uint32_t row = ok ? T.skip_row[t] : NIL;
        has_skip = row != NIL;
        missing = t_base == t_end;
        t_next = t_base;
        if (has_skip) { t_skip = (size_t)row * (T.n_tiles + 1); pf_end = T.skip[t_skip + 1]; }
    }
    if (threadIdx.x == 0) { s_hits = 0; s_ntouched[0] = 0; s_ntouched[1] = 0; }
    int any_missing = __syncthreads_or(missing);
    const bool dead = (a.mode == 1 && any_missing) || nt == 0;
    const float scale = (float)(1u << a.shift);
    const uint32_t n_tiles = dead ? 0 : T.n_tiles;
    unsigned int my_hits = 0;

    for (uint32_t tile = 0; tile < n_tiles; ++tile) {
        const int cb = tile & 1;
        const uint32_t lo = tile * BM_TILE;
        const uint32_t hi = lo + BM_TILE < T.n_docs ? lo + BM_TILE : T.n_docs;
        // slice [sb, se) of this lane's term in the tile
        uint64_t sb = t_next, se = t_next;
        if (own) {
            if (has_skip) {
                se = t_base + pf_end;
                if (tile + 2 <= T.n_tiles) pf_end = T.skip[t_skip + tile + 2];
            } else {
                while (se < t_end && T.post_doc[se] < hi) ++se;
            }
            t_next = se;
        }
        bool do_tile = true;
        if (a.mode == 1) do_tile = __syncthreads_and(!own || se > sb);   // AND: some term has nothing in this tile
        if (do_tile) {
            // pass 1: first chunk of every owned term loaded up front (independent loads in flight), then accumulated
            uint32_t d0[BM_TM_OWN], f0[BM_TM_OWN];
#pragma unroll
            for (int j = 0; j < BM_TM_OWN; ++j) {
                uint64_t b = __shfl_sync(0xFFFFFFFFu, sb, j), e = __shfl_sync(0xFFFFFFFFu, se, j);
                d0[j] = 0; f0[j] = 0;
                if (j < n_own && b + lane < e) { d0[j] = __ldg(T.post_doc + b + lane); f0[j] = __ldg(T.post_tfn + b + lane); }
            }
#pragma unroll
            for (int j = 0; j < BM_TM_OWN; ++j) {
                if (j >= n_own) break;
                uint64_t b = __shfl_sync(0xFFFFFFFFu, sb, j), e = __shfl_sync(0xFFFFFFFFu, se, j);
                float w = __shfl_sync(0xFFFFFFFFu, t_w, j);
                for (uint64_t base = b; base < e; base += 32) {   // warp-uniform trip count
                    bool act = base + lane < e;
                    uint32_t d = d0[j], tfn = f0[j];
                    if (base != b && act) { d = __ldg(T.post_doc + base + lane); tfn = __ldg(T.post_tfn + base + lane); }
                    bool first = false;
                    uint32_t off = 0;
                    if (act) {
                        float tff = a.use_tf ? (float)(tfn >> 8) : 1.0f;
                        float sc = __fmul_rn(w, __fdiv_rn(tff, __fadd_rn(tff, ncache[tfn & 0xFFu])));
                        uint32_t fx = (uint32_t)__float2uint_rn(__fmul_rn(sc, scale));
                        if (fx == 0) fx = 1;
                        off = d - lo;
                        first = atomicAdd(&acc[off], fx) == 0;
                        if (a.mode == 1) atomicAdd(reinterpret_cast<uint32_t*>(cnt8) + (off >> 2), 1u << (8 * (off & 3)));
                    }
                    unsigned m = __ballot_sync(0xFFFFFFFFu, first);
                    if (m) {
                        int basepos = 0;
                        if (lane == 0) basepos = atomicAdd(&s_ntouched[cb], __popc(m));
                        basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
                        int pos = basepos + __popc(m & ((1u << lane) - 1));
                        if (first && pos < BM_TOUCH_CAP) touched[pos] = (unsigned short)off;
                    }
                }
            }
            __syncthreads();
            // pass 2: every touched document once
            int ntouched = s_ntouched[cb];
            bool dense = ntouched > BM_TOUCH_CAP;
            int work = dense ? (int)(hi - lo) : ntouched;
            for (int base = 0; base < work; base += BM_THREADS * BM_ROUND) {
#pragma unroll
                for (int u = 0; u < BM_ROUND; ++u) {
                    int j = base + u * BM_THREADS + threadIdx.x;
                    if (j < work) {
                        uint32_t off = dense ? (uint32_t)j : (uint32_t)touched[j];
                        uint32_t v = acc[off];
                        if (v != 0) {
                            acc[off] = 0;
                            bool match = true;
                            if (a.mode == 1) { match = (int)cnt8[off] == nt; cnt8[off] = 0; }
                            uint32_t doc = lo + off;
                            if (match && T.alive) match = (T.alive[doc >> 6] >> (doc & 63)) & 1;
                            if (match) {
                                my_hits++;
                                float score = __fdiv_rn((float)v, scale);
                                bool after = true;
                                if (a.after_mode != 0) {
                                    uint32_t so = ordered_bits(score), ao = ordered_bits(a.after_score);
                                    after = so < ao || (so == ao && (a.after_mode == 3 || (a.after_mode == 2 && a.docaddr_base + doc > a.after_docaddr)));
                                }
                                uint64_t key = make_key(score, doc, 0);
                                if (after && key > tk_thr) tk_buf[atomicAdd(&tk_count, 1)] = key;
                            }
                        }
                    }
                }
                __syncthreads();
                if (tk_count > a.cap - BM_THREADS * BM_ROUND) tk.flush();
            }
            if (threadIdx.x == 0 && ntouched > 0) s_ntouched[cb] = 0;   // next user of this parity: tile + 2, behind tile + 1's barrier
        }
    }
    for (int off = 16; off >= 1; off >>= 1) my_hits += __shfl_xor_sync(0xFFFFFFFFu, my_hits, off);
    if (lane == 0 && my_hits) atomicAdd(&s_hits, (unsigned long long)my_hits);
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
