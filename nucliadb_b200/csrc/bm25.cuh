// nidx_b200 — K7: BM25 top-k over device-resident postings (sm_100a).
//
// Replaces the tantivy collector call of the reference:
//   nidx/nidx_text/src/reader.rs:432-435        TopDocs::with_limit(k+1).order_by_score() + Count
//   nidx/nidx_paragraph/src/reader.rs:290-292   same, OR of TermQuery(Basic) (keyword_parser.rs:62-67)
// with tantivy 0.26's BM25 (restated in oracle/bm25.hpp; parity unpinned):
//   score(doc) = sum over matching query terms of  idf_t * (1 + k1) * tf / (tf + k1 * (1 - b + b * fieldnorm / avg))
//
// Index-time layout (once per segment):
//   * post[i] = (doc, tf << 8 | fieldnorm_id(doc)), ascending doc per term: ONE 8-byte record per posting, the document's
//     length code travels with it (no random gather of fieldnorm[doc]);
//   * skip[row][f] = first posting of the term with doc >= f * BM_FINE, for every term with df >= BM_SKIP_DF: the slice
//     of a term in ANY run of fine tiles is two table entries.
// Query time: one CTA per query.  A query's postings hit a tiny fraction of the documents (50 terms of df ~5 k over 5 M
// documents touch 5 %), so the accumulator is an open-addressing HASH TABLE in shared memory (document -> fixed-point
// score), not a dense array: with the same shared memory a tile covers ~10x more documents, the runs of one term inside
// a tile are ~10x longer (full warps, amortised bounds), and there are ~10x fewer tiles and barriers.  The tile span is
// chosen per query from its posting count (load factor ~0.4); a tile that would overfill the table falls back to single
// fine tiles (BM_FINE documents <= half the slots, whatever the posting count).  Per tile:
//   pass 1  the tile's postings, cut into 32-posting chunks of ONE term each (consecutive, coalesced 8-byte loads, four
//           chunks in flight per warp), are scored and added with shared-memory atomics (measured on B200: one
//           warp-wide ATOMS per ~4 clk per SM, scripts/ubench_smem.cu).  Contributions are fixed point (2^-shift): the sum
//           does not depend on the order, equal scores stay bit-equal and TopDocs' (score desc, doc asc) tie order is
//           deterministic.  The thread that claims a slot counts the document (Count collector); the thread whose add
//           carries a document's sum across the current top-k threshold records the slot as a candidate, so only
//           documents that can still enter the top-k are looked at again;
//   pass 2  candidates -> streaming top-k buffer (final sums, exact threshold test); dense 16-byte reset of the table,
//           overlapped with warp 0 resolving the next tile's slices (skip entries prefetched one tile ahead).
// HBM traffic = the query's postings once (8 B each) + one skip entry per (term, tile).
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace nidx {

constexpr int BM_THREADS = 256;
constexpr int BM_WARPS = BM_THREADS / 32;
constexpr int BM_MAX_TERMS = 128;
constexpr int BM_TPL = BM_MAX_TERMS / 32;  // query terms per lane of the resolving warp
constexpr int BM_FINE = 4096;              // skip-table granularity (documents)
constexpr int BM_SKIP_DF = 256;            // terms with at least this many postings get a skip row
constexpr int BM_CHUNK_CAP = 1024;         // 32-posting chunks per tile handled through the chunk map
constexpr int BM_UNROLL = 4;               // chunks in flight per warp
constexpr uint32_t BM_EMPTY = 0xFFFFFFFFu;

struct TxtDev {
    uint32_t n_docs, n_terms, n_fine;
    const uint64_t* term_off;
    const uint2* post;               // (doc, tf << 8 | fieldnorm id)
    const uint32_t* skip_row;        // [n_terms] row in skip[] or NIL
    const uint32_t* skip;            // [rows][n_fine + 1] posting index relative to term_off[term]
    const uint64_t* alive;
};

struct Bm25Args {
    const uint32_t* query_terms;
    const uint32_t* query_off;
    int nq;
    int k, cap;                 // cap: top-k buffer entries (power of two >= 2k, >= k + BM_THREADS)
    int hash_bits;              // accumulator table = 1 << hash_bits slots (>= 2 * BM_FINE)
    const float* term_weight;   // [n_terms] idf * (1 + k1) from the collection statistics
    const float* norm_cache;    // [256] k1 * (1 - b + b * fieldnorm(id) / avg)
    int shift;                  // fixed point: 2^-shift
    int after_mode;             // search-after (nidx_paragraph reader.rs:379-392): 0 none, 1 Drop, 2 KeepAfter, 3 Keep
    float after_score;
    uint64_t after_docaddr, docaddr_base;
    uint64_t* out_keys;         // [nq][k] rank keys (score desc, doc asc), 0 = none
    unsigned long long* out_total;  // [nq] matching documents (Count collector)
    unsigned int* error_flag;   // set if a table ever filled up (cannot happen by construction; checked by the host in debug runs)
};

// ---- index-time kernels ---------------------------------------------------------------------------------
__global__ void bm25_pack_kernel(const uint32_t* __restrict__ post_doc, const uint32_t* __restrict__ post_tf, const unsigned char* __restrict__ fieldnorm,
                                 uint64_t n_post, uint2* __restrict__ post) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t tf = post_tf ? post_tf[i] : 1u;
        if (tf > 0xFFFFFFu) tf = 0xFFFFFFu;
        uint32_t d = post_doc[i];
        post[i] = make_uint2(d, (tf << 8) | fieldnorm[d]);
    }
}
// one thread per (skip row, fine-tile boundary)
__global__ void bm25_build_skip_kernel(const uint64_t* __restrict__ term_off, const uint2* __restrict__ post, const uint32_t* __restrict__ row_term,
                                       uint32_t n_rows, uint32_t n_fine, uint32_t* __restrict__ skip) {
    uint64_t total = (uint64_t)n_rows * (n_fine + 1);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t row = (uint32_t)(i / (n_fine + 1)), t = (uint32_t)(i % (n_fine + 1));
        uint32_t term = row_term[row];
        uint64_t b = term_off[term], e = term_off[term + 1];
        uint64_t bound = (uint64_t)t * BM_FINE;
        uint64_t l = b, r = e;
        while (l < r) {
            uint64_t m = (l + r) >> 1;
            if ((uint64_t)post[m].x < bound) l = m + 1; else r = m;
        }
        skip[i] = (uint32_t)(l - b);
    }
}

__host__ __device__ __forceinline__ size_t bm_smem_bytes(int cap, int hash_bits, bool conj) {
    size_t slots = (size_t)1 << hash_bits;
    return (size_t)cap * 8 + slots * 4 * 2 + (conj ? slots : 0) /* byte counters */ + slots * 2 /* candidates */ + 2 * BM_MAX_TERMS * (8 + 4) /* run start, length */ +
           2 * (BM_MAX_TERMS + 1) * 4 /* chunk prefix */ + BM_MAX_TERMS * 4 /* weights */ + 1024 /* norm / ratio table */ + 2 * BM_CHUNK_CAP + 64;
}

__device__ __forceinline__ uint2 ldg_post(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// CONJ: nidx_text (all terms must match).  TF: real term frequencies (else IndexRecordOption::Basic, tf == 1).
template <bool CONJ, bool TF>
__global__ void __launch_bounds__(BM_THREADS) bm25_kernel(TxtDev T, Bm25Args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int tk_count;
    __shared__ uint64_t tk_thr;
    __shared__ int s_ncand, s_tk_snapshot, s_nchunks[2], s_ptile[2], s_minlen[2];
    __shared__ unsigned long long s_hits;
    const uint32_t S = 1u << a.hash_bits, smask = S - 1;
    unsigned char* p = smem;
    uint64_t* tk_buf = reinterpret_cast<uint64_t*>(p); p += (size_t)a.cap * 8;
    uint64_t* run_b = reinterpret_cast<uint64_t*>(p); p += 2 * BM_MAX_TERMS * 8;       // [2][terms] first posting of the tile (absolute)
    uint32_t* keys = reinterpret_cast<uint32_t*>(p); p += (size_t)S * 4;
    uint32_t* vals = reinterpret_cast<uint32_t*>(p); p += (size_t)S * 4;
    uint32_t* cnts = reinterpret_cast<uint32_t*>(p); if (CONJ) p += (size_t)S;         // matched-term counters, one BYTE per slot (AND only)
    uint32_t* run_len = reinterpret_cast<uint32_t*>(p); p += 2 * BM_MAX_TERMS * 4;     // [2][terms]
    uint32_t* pre = reinterpret_cast<uint32_t*>(p); p += 2 * (BM_MAX_TERMS + 1) * 4;   // [2][terms + 1] exclusive prefix of the runs' chunk counts
    float* tw = reinterpret_cast<float*>(p); p += BM_MAX_TERMS * 4;                    // weight * 2^shift is NOT folded: the oracle's order of operations is kept
    float* ntab = reinterpret_cast<float*>(p); p += 1024;                              // TF: norm cache; else 1 / (1 + norm) per fieldnorm id
    unsigned short* cand = reinterpret_cast<unsigned short*>(p); p += (size_t)S * 2;   // slots whose sum crossed the threshold in this tile
    unsigned char* chunk_run = p;                                                      // [2][BM_CHUNK_CAP] run of every chunk

    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* terms = a.query_terms + a.query_off[q];
    int nt = (int)(a.query_off[q + 1] - a.query_off[q]);
    if (nt > BM_MAX_TERMS) nt = BM_MAX_TERMS;
    BlockTopK tk;
    tk.init(tk_buf, &tk_count, &tk_thr, a.k, a.cap);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        float nc = a.norm_cache[i];
        ntab[i] = TF ? nc : __fdiv_rn(1.0f, __fadd_rn(1.0f, nc));   // tf == 1: tf / (tf + norm), the same two roundings as the division per posting
    }
    for (uint32_t i = threadIdx.x; i < S; i += blockDim.x) { keys[i] = BM_EMPTY; vals[i] = 0; if (CONJ && i < S / 4) cnts[i] = 0; }

    // ---- warp 0 owns the terms: lane l holds terms l, l + 32, ... (cursor, end, skip row, prefetched skip entry) ----
    uint64_t t_base[BM_TPL], t_end[BM_TPL], t_cur[BM_TPL];
    size_t t_skip[BM_TPL];
    uint32_t t_pf[BM_TPL];
    bool t_has[BM_TPL];
    bool missing = false;
    unsigned long long my_total = 0;
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < BM_TPL; ++j) {
            int i = lane + 32 * j;
            t_base[j] = t_end[j] = t_cur[j] = 0; t_skip[j] = 0; t_pf[j] = 0; t_has[j] = false;
            if (i < nt) {
                uint32_t t = terms[i];
                bool ok = t < T.n_terms;
                t_base[j] = ok ? T.term_off[t] : 0;
                t_end[j] = ok ? T.term_off[t + 1] : 0;
                t_cur[j] = t_base[j];
                tw[i] = ok ? a.term_weight[t] : 0.0f;
                uint32_t row = ok ? T.skip_row[t] : NIL;
                t_has[j] = row != NIL;
                if (t_has[j]) t_skip[j] = (size_t)row * (T.n_fine + 1);
                missing |= t_base[j] == t_end[j];
                my_total += t_end[j] - t_base[j];
            }
        }
        for (int off = 16; off >= 1; off >>= 1) my_total += __shfl_xor_sync(0xFFFFFFFFu, my_total, off);
    }
    if (threadIdx.x == 0) { s_hits = 0; s_ncand = 0; s_tk_snapshot = 0; s_ptile[0] = (int)min(my_total, (unsigned long long)INT_MAX); }
    int any_missing = __syncthreads_or(missing);   // an AND query with a term without postings matches nothing
    const bool dead = (CONJ && any_missing) || nt == 0;
    const float scale = (float)(1u << a.shift);
    const uint32_t n_fine = dead ? 0 : T.n_fine;
    // tile span (fine tiles): the query's postings spread evenly would fill 3/8 of the table per tile
    const uint32_t limit = S - S / 4;     // a tile with more postings than this is redone one fine tile at a time
    uint32_t m = 1;
    {
        unsigned long long P = (unsigned long long)(unsigned)s_ptile[0];
        unsigned long long target = (S / 8) * 3;
        if (P == 0) m = n_fine ? n_fine : 1;
        else {
            unsigned long long nf = n_fine ? n_fine : 1, mm = target * nf / P;
            if (mm < 1) mm = 1;
            if (mm > nf) mm = nf;
            m = (uint32_t)mm;
        }
    }
    __syncthreads();

    // resolve(f1, buf) by warp 0: slices of every term in fine tiles [cursor position, f1) -> run_b / run_len / pre / chunk_run of `buf`
    auto resolve = [&](uint32_t f1, uint32_t m_next, int buf) {
        uint32_t hi = (uint64_t)f1 * BM_FINE < T.n_docs ? f1 * BM_FINE : T.n_docs;
        uint32_t run = 0;
        int mn = INT_MAX, tot = 0;
#pragma unroll
        for (int j = 0; j < BM_TPL; ++j) {
            int i = lane + 32 * j;
            uint32_t len = 0;
            if (i < nt) {
                uint64_t bgn = t_cur[j], end;
                if (t_has[j]) {
                    end = t_base[j] + t_pf[j];
                } else {                                                   // rare term: a few postings in total
                    uint64_t l = bgn;
                    while (l < t_end[j] && T.post[l].x < hi) ++l;
                    end = l;
                }
                run_b[buf * BM_MAX_TERMS + i] = bgn;
                len = (uint32_t)(end - bgn);
                run_len[buf * BM_MAX_TERMS + i] = len;
                t_cur[j] = end;
                mn = min(mn, (int)len);
                tot += (int)len;
            }
            // exclusive scan of the chunk counts over the terms (term order = lane + 32 j: scan lanes, then carry `run`)
            uint32_t c = (len + 31) >> 5, x = c;
            for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off); if (lane >= off) x += y; }
            if (i < nt) pre[buf * (BM_MAX_TERMS + 1) + i] = run + x - c;
            uint32_t first = run + x - c;
            if (i < nt && first + c <= BM_CHUNK_CAP) for (uint32_t g = 0; g < c; ++g) chunk_run[buf * BM_CHUNK_CAP + first + g] = (unsigned char)i;
            run += __shfl_sync(0xFFFFFFFFu, x, 31);
        }
        for (int off = 16; off >= 1; off >>= 1) { mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, off)); tot += __shfl_xor_sync(0xFFFFFFFFu, tot, off); }
        if (lane == 0) { pre[buf * (BM_MAX_TERMS + 1) + nt] = run; s_nchunks[buf] = (int)run; s_ptile[buf] = tot; s_minlen[buf] = mn; }
        // request the skip entries of the tile after this one
        uint32_t f2 = f1 + m_next < T.n_fine ? f1 + m_next : T.n_fine;
#pragma unroll
        for (int j = 0; j < BM_TPL; ++j) if (t_has[j]) t_pf[j] = __ldg(T.skip + t_skip[j] + f2);
    };
    auto load_pf = [&](uint32_t f1) {   // synchronous (re)load of the skip entries for boundary f1
#pragma unroll
        for (int j = 0; j < BM_TPL; ++j) if (t_has[j]) t_pf[j] = __ldg(T.skip + t_skip[j] + f1);
    };

    // ---- one posting: score, hash insert, add, threshold crossing ---------------------------------------------------
    uint32_t thr_fx = 1;   // sums >= thr_fx may still enter the top-k (1 = everything that is touched)
    unsigned int my_hits = 0;
    auto posting = [&](bool act, uint2 pd, float w, uint32_t lo) {
        bool crossed = false;
        uint32_t h = 0;
        if (act) {
            float frac;
            if (TF) { float tff = (float)(pd.y >> 8); frac = __fdiv_rn(tff, __fadd_rn(tff, ntab[pd.y & 0xFFu])); }
            else frac = ntab[pd.y & 0xFFu];
            float s = __fmul_rn(w, frac);
            uint32_t fx = (uint32_t)__float2uint_rn(__fmul_rn(s, scale));
            if (fx == 0) fx = 1;
            uint32_t key = pd.x - lo;
            h = (key * 2654435761u) >> (32 - a.hash_bits);
            uint32_t probes = 0;
            while (true) {
                uint32_t old = atomicCAS(&keys[h], BM_EMPTY, key);
                if (old == BM_EMPTY) {   // claimed: this thread counts the document (OR: the Count collector)
                    if (!CONJ) my_hits += T.alive ? (unsigned)((T.alive[pd.x >> 6] >> (pd.x & 63)) & 1) : 1u;
                    break;
                }
                if (old == key) break;
                h = (h + 1) & smask;
                if (++probes > S) { atomicExch(a.error_flag, 1u); break; }
            }
            uint32_t oldv = atomicAdd(&vals[h], fx);
            if (CONJ) {
                uint32_t oldw = atomicAdd(&cnts[h >> 2], 1u << (8 * (h & 3)));   // byte counter (nt <= 128) inside its 32-bit word
                uint32_t oldc = (oldw >> (8 * (h & 3))) & 0xFFu;
                crossed = (int)(oldc + 1) == nt;     // the posting that completes the conjunction hands the document on
            } else {
                crossed = oldv < thr_fx && oldv + fx >= thr_fx;
            }
        }
        unsigned mk = __ballot_sync(0xFFFFFFFFu, crossed);
        if (mk) {
            int basepos = 0;
            if (lane == 0) basepos = atomicAdd(&s_ncand, __popc(mk));
            basepos = __shfl_sync(0xFFFFFFFFu, basepos, 0);
            if (crossed) cand[basepos + __popc(mk & ((1u << lane) - 1))] = (unsigned short)h;
        }
    };
    // a candidate slot -> top-k buffer (final sum; exact threshold, alive, search-after)
    auto offer_slot = [&](uint32_t h, uint32_t lo) {
        uint32_t v = vals[h], doc = lo + keys[h];
        bool match = true;
        if (T.alive) match = (T.alive[doc >> 6] >> (doc & 63)) & 1;
        if (CONJ && match) my_hits++;
        if (match) {
            float score = __fdiv_rn((float)v, scale);
            bool after = true;   // is_after(): strictly lower score, or an equal score that the tie break keeps
            if (a.after_mode != 0) {
                uint32_t so = ordered_bits(score), ao = ordered_bits(a.after_score);
                after = so < ao || (so == ao && (a.after_mode == 3 || (a.after_mode == 2 && a.docaddr_base + doc > a.after_docaddr)));
            }
            uint64_t key = make_key(score, doc, 0);
            if (after && key > tk_thr) tk_buf[atomicAdd(&tk_count, 1)] = key;
        }
    };

    uint32_t f0 = 0;
    int buf = 0;
    if (n_fine) {
        if (warp == 0) { load_pf(min(m, n_fine)); resolve(min(m, n_fine), m, 0); }
        __syncthreads();
    }
    while (f0 < n_fine) {
        uint32_t f1 = min(f0 + m, n_fine);
        if (s_ptile[buf] > (int)limit && f1 - f0 > 1) {   // (uniform) would overfill the table: redo from f0 one fine tile at a time
            __syncthreads();
            m = 1;
            if (warp == 0) {
#pragma unroll
                for (int j = 0; j < BM_TPL; ++j) { int i = lane + 32 * j; if (i < nt) t_cur[j] = run_b[buf * BM_MAX_TERMS + i]; }
                load_pf(f0 + 1);
                resolve(f0 + 1, 1, buf);
            }
            __syncthreads();
            continue;
        }
        const uint32_t lo = f0 * BM_FINE;
        const bool skip_tile = s_ptile[buf] == 0 || (CONJ && s_minlen[buf] == 0);   // AND: some term has nothing in this tile
        if (!skip_tile) {
            // ---- pass 1 ----
            const int nch = s_nchunks[buf];
            const uint32_t* prb = pre + buf * (BM_MAX_TERMS + 1);
            if (nch <= BM_CHUNK_CAP) {
                for (int g0 = warp; g0 < nch; g0 += BM_WARPS * BM_UNROLL) {
                    uint2 pd[BM_UNROLL];
                    float w[BM_UNROLL];
                    bool act[BM_UNROLL];
#pragma unroll
                    for (int u = 0; u < BM_UNROLL; ++u) {
                        int g = g0 + u * BM_WARPS;
                        act[u] = false; w[u] = 0.0f; pd[u] = make_uint2(0, 0);
                        if (g < nch) {
                            int r = chunk_run[buf * BM_CHUNK_CAP + g];
                            uint32_t within = (uint32_t)(g - (int)prb[r]) * 32u + lane;
                            act[u] = within < run_len[buf * BM_MAX_TERMS + r];
                            w[u] = tw[r];
                            if (act[u]) pd[u] = ldg_post(T.post + run_b[buf * BM_MAX_TERMS + r] + within);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < BM_UNROLL; ++u)
                        if (g0 + u * BM_WARPS < nch) posting(act[u], pd[u], w[u], lo);
                }
            } else {   // dense tile (more chunks than the map holds): the block walks each run
                for (int r = 0; r < nt; ++r) {
                    uint32_t len = run_len[buf * BM_MAX_TERMS + r];
                    uint64_t b = run_b[buf * BM_MAX_TERMS + r];
                    float w = tw[r];
                    for (uint32_t base = 0; base < len; base += BM_THREADS) {   // warp-uniform trip count
                        uint32_t i = base + threadIdx.x;
                        bool act = i < len;
                        uint2 pd = act ? ldg_post(T.post + b + i) : make_uint2(0, 0);
                        posting(act, pd, w, lo);
                    }
                }
            }
            __syncthreads();
            // ---- pass 2a: candidates -> top-k buffer; warp 0 resolves the next tile meanwhile ----
            // The branch below must be uniform: it is taken on the buffer fill recorded at the end of the previous tile's pass 2b
            // (pass 1 does not touch the buffer), never on tk_count itself, which the other warps are already incrementing.
            const int ncand = s_ncand;
            const int room = a.cap - a.k;
            if (s_tk_snapshot + ncand > a.cap) {     // (uniform) rare: the buffer cannot take them all at once -> rounds with a flush each
                for (int base = 0; base < ncand; base += room) {
                    tk.flush();
                    int end = min(ncand, base + room);
                    for (int i = base + threadIdx.x; i < end; i += BM_THREADS) offer_slot(cand[i], lo);
                }
                tk.flush();
                if (warp == 0 && f1 < n_fine) resolve(min(f1 + m, n_fine), m, buf ^ 1);
            } else if (warp == 0) {
                if (f1 < n_fine) resolve(min(f1 + m, n_fine), m, buf ^ 1);
            } else {
                for (int i = threadIdx.x - 32; i < ncand; i += BM_THREADS - 32) offer_slot(cand[i], lo);
            }
            __syncthreads();
            // ---- pass 2b: dense reset of the table (16-byte stores) ----
            if (tk_thr != 0) {   // threshold in fixed point, conservative (float(v) is within 2^-24 of v)
                float ts = key_score(tk_thr);
                float lowb = __fmul_rn(__fmul_rn(ts, scale), 0.9999990f);
                thr_fx = lowb >= 1.0f ? (uint32_t)lowb : 1u;
            }
            uint4 e4 = make_uint4(BM_EMPTY, BM_EMPTY, BM_EMPTY, BM_EMPTY), z4 = make_uint4(0, 0, 0, 0);
            for (uint32_t i = threadIdx.x; i < S / 4; i += BM_THREADS) {
                reinterpret_cast<uint4*>(keys)[i] = e4;
                reinterpret_cast<uint4*>(vals)[i] = z4;
                if (CONJ && i < S / 16) reinterpret_cast<uint4*>(cnts)[i] = z4;
            }
            if (threadIdx.x == 0) { s_ncand = 0; s_tk_snapshot = tk_count; }
            __syncthreads();
        } else {
            if (warp == 0 && f1 < n_fine) resolve(min(f1 + m, n_fine), m, buf ^ 1);
            __syncthreads();
        }
        f0 = f1;
        buf ^= 1;
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
