// nidx_b200 -- K7 variant: BM25 top-k with warp-chunked posting assignment (sm_100a).  See bm25.cuh for the layout, the
// reference citations and the two-pass tile scheme; this file only holds the variant kernel.  Selected with NIDX_B200_BM25=w
// (api.cu) until it has been measured against bm25_kernel on the 5 M-document workload.
#pragma once
#include "bm25.cuh"

namespace nidx {

// BW_THREADS x BW_ROUND = 512 postings per round of the CTA in both instantiations, so the streaming top-k buffer (cap >= k + 512)
// and with it the shared-memory footprint (about 74 KB at k = 100: 3 CTAs per SM) are those of bm25_kernel:
//   <128, 4>  4 warps: the per-tile bookkeeping is paid by half as many warps, 12 warps per SM;
//   <256, 2>  8 warps: bm25_kernel's shape, one binary search per lane and round instead of two.
__host__ __device__ __forceinline__ size_t bw_smem_bytes(int cap, bool conj) { return bm_smem_bytes(cap, conj) + 1024; }

// Same tile pipeline as bm25_kernel (tile t is accumulated and collected while the slices of tile t+1 are resolved and its
// first postings are in flight), with the per-posting overhead cut down:
//   * a warp owns a CONTIGUOUS run of the tile's flattened postings (BW_CHUNK per round) and its lanes stride through it,
//     so a lane finds its term with ONE binary search per round and then walks forward (consecutive slots are 32 postings
//     apart, a term slice averages more than that) instead of searching for every posting;
//   * tf == 1 (nidx_paragraph): idf*(1+k1) * 1/(1+norm[fieldnorm]) comes from a 256-entry table computed with the same
//     rounded operations at kernel start -- no division per posting; score = sum * 2^-shift instead of sum / 2^shift.
// All three are bit-identical to bm25_kernel's arithmetic.
template <int BW_THREADS, int BW_ROUND>
__global__ void __launch_bounds__(BW_THREADS) bm25_w_kernel(TxtDev T, Bm25Args a) {
    constexpr int BW_CHUNK = BW_ROUND * 32;    // contiguous flattened postings a warp owns per round
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
    float* ratio = reinterpret_cast<float*>(p); p += 1024;                         // tf == 1: 1 / (1 + ncache[fieldnorm id])
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
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        float nc = a.norm_cache[i];
        ncache[i] = nc;
        ratio[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, nc));   // the same two rounded operations bm25_kernel does per posting
    }
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
    const float inv_scale = __uint_as_float((uint32_t)(127 - a.shift) << 23);   // 2^-shift: x * inv_scale == x / scale, exactly
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
    // fetch(buf): this thread's first BW_ROUND postings of the tile resolved in `buf` -> registers (loads in flight)
    uint32_t d_n[BW_ROUND], tfn_n[BW_ROUND], d_c[BW_ROUND], tfn_c[BW_ROUND];
    int tl_n[BW_ROUND], tl_c[BW_ROUND];
    auto locate = [&](int buf, int i, int& l) -> uint64_t {
        const int* pr = pre + buf * BM_MAX_TERMS;
        int lo_ = 0, r = nt - 1;  // last term with pre[t] <= i
        while (lo_ < r) { int m = (lo_ + r + 1) >> 1; if (pr[m] <= i) lo_ = m; else r = m - 1; }
        l = lo_;
        return cur[buf * BM_MAX_TERMS + lo_] + (uint64_t)(i - pr[lo_]);
    };
    // slot u of this lane in the round starting at flattened index `base`: base + warp * BW_CHUNK + u * 32 + lane.  One binary
    // search for slot 0, then the term index only moves forward (walk): pre[] is non-decreasing and the slots ascend by 32.
    auto walk = [&](int buf, int i, int& l) -> uint64_t {
        const int* pr = pre + buf * BM_MAX_TERMS;
        while (l + 1 < nt && pr[l + 1] <= i) ++l;   // last term with pre[t] <= i
        return cur[buf * BM_MAX_TERMS + l] + (uint64_t)(i - pr[l]);
    };
    auto fetch = [&](int buf) {
        int total = s_total[buf];
        bool skip_tile = total == 0 || (a.mode == 1 && s_min_count[buf] == 0);
        int i0 = warp * BW_CHUNK + lane, l = 0;
        bool any = !skip_tile && i0 < total;
        if (any) locate(buf, i0, l);
#pragma unroll
        for (int u = 0; u < BW_ROUND; ++u) {
            int i = i0 + u * 32;
            tl_n[u] = -1;
            if (any && i < total) {
                uint64_t pi = walk(buf, i, l);
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
            float s;
            if (a.use_tf) {
                float tff = (float)(tfn >> 8);
                s = __fmul_rn(tw[l], __fdiv_rn(tff, __fadd_rn(tff, ncache[tfn & 0xFFu])));
            } else {
                s = __fmul_rn(tw[l], ratio[tfn & 0xFFu]);
            }
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
    for (int u = 0; u < BW_ROUND; ++u) { d_c[u] = d_n[u]; tfn_c[u] = tfn_n[u]; tl_c[u] = tl_n[u]; }

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
            for (int u = 0; u < BW_ROUND; ++u) accumulate(lo, d_c[u], tfn_c[u], tl_c[u], cb);
            for (int base = BW_THREADS * BW_ROUND; base < total; base += BW_THREADS * BW_ROUND) {
                int i0 = base + warp * BW_CHUNK + lane, l = 0;
                bool any = i0 < total;
                if (any) locate(cb, i0, l);
                uint32_t d_r[BW_ROUND], tfn_r[BW_ROUND];
                int tl_r[BW_ROUND];
#pragma unroll
                for (int u = 0; u < BW_ROUND; ++u) {            // all of the round's loads first ...
                    int i = i0 + u * 32;
                    tl_r[u] = -1; d_r[u] = 0; tfn_r[u] = 0;
                    if (any && i < total) { uint64_t pi = walk(cb, i, l); d_r[u] = __ldg(T.post_doc + pi); tfn_r[u] = __ldg(T.post_tfn + pi); tl_r[u] = l; }
                }
#pragma unroll
                for (int u = 0; u < BW_ROUND; ++u) accumulate(lo, d_r[u], tfn_r[u], tl_r[u], cb);   // ... then the (warp-synchronous) accumulation
            }
            __syncthreads();
            // pass 2: every touched document once -> count, reset, offer to the streaming top-k
            int ntouched = s_ntouched[cb];
            bool dense = ntouched > BM_TOUCH_CAP;      // list overflow: fall back to scanning the whole tile
            int work = dense ? (int)(hi - lo) : ntouched;
            for (int base = 0; base < work; base += BW_THREADS * BW_ROUND) {
#pragma unroll
                for (int u = 0; u < BW_ROUND; ++u) {
                    int j = base + u * BW_THREADS + threadIdx.x;
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
                                float score = __fmul_rn((float)v, inv_scale);
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
                if (tk_count > a.cap - BW_THREADS * BW_ROUND) tk.flush();
            }
            // reset after everyone has read it (the loop above synchronised at least once iff ntouched > 0); the next
            // user of this parity is tile + 2, behind the barriers of the next iteration's resolve()
            if (threadIdx.x == 0 && ntouched > 0) s_ntouched[cb] = 0;
        }
#pragma unroll
        for (int u = 0; u < BW_ROUND; ++u) { d_c[u] = d_n[u]; tfn_c[u] = tfn_n[u]; tl_c[u] = tl_n[u]; }
    }
    for (int off = 16; off >= 1; off >>= 1) my_hits += __shfl_xor_sync(0xFFFFFFFFu, my_hits, off);
    if (lane == 0 && my_hits) atomicAdd(&s_hits, (unsigned long long)my_hits);
    int c = tk.finish();
    uint64_t* out = a.out_keys + (size_t)q * a.k;
    for (int i = threadIdx.x; i < a.k; i += blockDim.x) out[i] = i < c ? tk_buf[i] : 0;
    if (threadIdx.x == 0 && a.out_total) a.out_total[q] = s_hits;
}


}  // namespace nidx
