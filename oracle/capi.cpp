// ORACLE — TEST INFRASTRUCTURE ONLY (see distance.hpp).  extern "C" surface for ctypes
// (oracle/oracle.py).  Built by oracle/Makefile into oracle/liboracle.so.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "bm25.hpp"
#include "distance.hpp"
#include "hnsw.hpp"
#include "rabitq.hpp"
#include "segment.hpp"

using namespace nidx_oracle;

// One visited-set scratch per OS thread, kept across calls: the reference's FxHashSet costs nothing to
// set up, so the epoch array must not be re-allocated (O(n) page faults) on every batch.
static Scratch& tls_scratch() {
    static thread_local Scratch sc;
    return sc;
}

static Data make_data(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim) {
    Data D;
    D.vecs = vecs; D.norms = norms; D.n = n; D.d = d; D.ld = ld; D.sim = sim;
    return D;
}
static GraphView make_view(uint32_t n, int M, int M0, const uint8_t* level, uint32_t entry_node, uint32_t entry_layer, uint32_t* adj0,
                           float* w0, const uint64_t* upper_off, uint32_t* adjU, float* wU) {
    GraphView g;
    g.n = n; g.M = M; g.M0 = M0; g.s0 = stride0_for(M0); g.su = strideU_for(M);
    g.level = level; g.entry_node = entry_node; g.entry_layer = entry_layer;
    g.adj0 = adj0; g.w0 = w0; g.upper_off = upper_off; g.adjU = adjU; g.wU = wU;
    return g;
}

extern "C" {

float oracle_dot(const float* a, const float* b, int d) { return dot_ordered(a, b, d); }
double oracle_dot_f64(const float* a, const float* b, int d) { return dot_f64(a, b, d); }
float oracle_norm(const float* a, int d) { return norm_ordered(a, d); }
float oracle_cosine(const float* a, const float* b, int d) {
    return cosine_from_parts(dot_ordered(a, b, d), norm_ordered(a, d), norm_ordered(b, d));
}
void oracle_normalize(const float* in, float* out, int d) { normalize_vector(in, out, d); }
void oracle_norms(const float* vecs, uint32_t n, int d, int ld, float* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = norm_ordered(vecs + (size_t)i * ld, d);
}
int oracle_stride0(int M0) { return stride0_for(M0); }
int oracle_strideU(int M) { return strideU_for(M); }
int oracle_prune_m(int m) { return prune_m(m); }
int oracle_use_hnsw(uint64_t total, uint64_t matching, uint64_t top_k, int has_rabitq, int M) {
    return use_hnsw(total, matching, top_k, has_rabitq != 0, (size_t)M) ? 1 : 0;
}
void oracle_assign_levels(uint32_t n, int M, uint64_t seed, uint8_t* level) { assign_levels(n, M, seed, level); }

// Exact scan (segment.rs:569-623) for nq queries; out_* are [nq][k], out_count[nq].
void oracle_brute_force(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim, const float* queries, int nq, int qld,
                        int k, float min_score, const uint64_t* alive_bits, uint32_t n_paragraphs, const uint32_t* first_vec,
                        const uint32_t* num_vec, uint32_t* out_ids, float* out_scores, int* out_count, int nthreads) {
    Data D = make_data(vecs, norms, n, d, ld, sim);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int qi = 0; qi < nq; ++qi) {
        Query q{queries + (size_t)qi * qld, sim != SIM_DOT ? norm_ordered(queries + (size_t)qi * qld, d) : 0.0f};
        auto r = brute_force_search(D, q, (size_t)k, min_score, alive_bits, n_paragraphs, first_vec, num_vec);
        out_count[qi] = (int)r.size();
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j] = j < (int)r.size() ? r[j].id : NIL;
            out_scores[(size_t)qi * k + j] = j < (int)r.size() ? r[j].score : 0.0f;
        }
    }
}

// HNSW search (search.rs:306-383) for nq queries on a flat graph.  counters[3] = n_dist, n_expand, n_edges_read (summed).
void oracle_hnsw_search(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim, int M, int M0, const uint8_t* level,
                        uint32_t entry_node, uint32_t entry_layer, const uint32_t* adj0, const uint64_t* upper_off, const uint32_t* adjU,
                        const float* queries, int nq, int qld, int k, int ef, float min_score, int with_duplicates, int multi_vector,
                        const uint64_t* filter_bits, const uint32_t* paragraph_of, uint32_t* out_ids, float* out_scores, int* out_count,
                        uint64_t* counters, int nthreads) {
    Data D = make_data(vecs, norms, n, d, ld, sim);
    GraphView G = make_view(n, M, M0, level, entry_node, entry_layer, const_cast<uint32_t*>(adj0), nullptr, upper_off,
                            const_cast<uint32_t*>(adjU), nullptr);
    int nt = nthreads > 0 ? nthreads : 1;
    std::vector<Counters> cnts(nt);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int qi = 0; qi < nq; ++qi) {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        Query q{queries + (size_t)qi * qld, sim != SIM_DOT ? norm_ordered(queries + (size_t)qi * qld, d) : 0.0f};
        NodeFilter f;
        f.filter_bits = filter_bits; f.paragraph_of = paragraph_of; f.with_duplicates = with_duplicates != 0; f.multi_vector = multi_vector != 0;
        auto r = hnsw_search(D, G, q, (size_t)k, ef, min_score, f, tls_scratch(), &cnts[t]);
        if (r.size() > (size_t)k) r.resize(k);  // segment.rs:555 .take(top_k)
        out_count[qi] = (int)r.size();
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j] = j < (int)r.size() ? r[j].id : NIL;
            out_scores[(size_t)qi * k + j] = j < (int)r.size() ? r[j].score : 0.0f;
        }
    }
    if (counters) {
        counters[0] = counters[1] = counters[2] = 0;
        for (auto& c : cnts) { counters[0] += c.n_dist; counters[1] += c.n_expand; counters[2] += c.n_edges_read; }
    }
}

// One layer_search (search.rs:242-304) for tests: returns up to k results, sorted desc.
int oracle_layer_search(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim, int M, int M0, const uint8_t* level,
                        const uint32_t* adj0, const uint64_t* upper_off, const uint32_t* adjU, const float* query, int layer, int k,
                        const uint32_t* eps, int n_eps, uint32_t* out_ids, float* out_scores) {
    Data D = make_data(vecs, norms, n, d, ld, sim);
    GraphView G = make_view(n, M, M0, level, 0, 0, const_cast<uint32_t*>(adj0), nullptr, upper_off, const_cast<uint32_t*>(adjU), nullptr);
    Scratch sc;
    Query q{query, sim != SIM_DOT ? norm_ordered(query, d) : 0.0f};
    std::vector<uint32_t> e(eps, eps + n_eps);
    auto r = layer_search(D, G, q, layer, (size_t)k, e, sc, nullptr);
    for (size_t i = 0; i < r.size(); ++i) { out_ids[i] = r[i].id; out_scores[i] = r[i].score; }
    return (int)r.size();
}

// build.rs:57-95 for tests.
int oracle_select_neighbours(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim, int k, const uint32_t* cand_ids,
                             const float* cand_scores, int n_cand, uint32_t* out_ids, float* out_scores) {
    Data D = make_data(vecs, norms, n, d, ld, sim);
    std::vector<Scored> c(n_cand);
    for (int i = 0; i < n_cand; ++i) c[i] = {cand_ids[i], cand_scores[i]};
    auto r = select_neighbours_heuristic(D, (size_t)k, c, nullptr);
    for (size_t i = 0; i < r.size(); ++i) { out_ids[i] = r[i].id; out_scores[i] = r[i].score; }
    return (int)r.size();
}

// Graph sizing helper: fills upper_off[n], returns number of upper rows; entry = lowest id of the top layer.
uint64_t oracle_graph_layout(uint32_t n, const uint8_t* level, uint64_t* upper_off, uint32_t* entry_node, uint32_t* entry_layer) {
    uint64_t rows = 0;
    uint32_t top = 0;
    for (uint32_t i = 0; i < n; ++i) { upper_off[i] = rows; rows += level[i]; if (level[i] > top) top = level[i]; }
    *entry_layer = top;
    *entry_node = 0;
    for (uint32_t i = 0; i < n; ++i) if (level[i] == top) { *entry_node = i; break; }
    return rows;
}

// Batch-synchronous build into caller-provided flat arrays (pre-filled with NIL / 0).
double oracle_hnsw_build(const float* vecs, const float* norms, uint32_t n, int d, int ld, int sim, int M, int M0, int efC,
                         const uint8_t* level, uint32_t entry_node, uint32_t entry_layer, uint32_t* adj0, float* w0, const uint64_t* upper_off,
                         uint32_t* adjU, float* wU, const uint32_t* order, const uint32_t* batch_ends, uint32_t n_batches, int nthreads,
                         uint64_t* counters) {
    Data D = make_data(vecs, norms, n, d, ld, sim);
    Params prm;
    prm.M = M; prm.M0 = M0; prm.efC = efC;
    GraphView G = make_view(n, M, M0, level, entry_node, entry_layer, adj0, w0, upper_off, adjU, wU);
    auto t0 = std::chrono::steady_clock::now();
    int nt = nthreads > 0 ? nthreads : 1;
    std::vector<Counters> cnts(nt);
    uint32_t begin = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        uint32_t end = batch_ends[b];
        std::vector<std::vector<std::vector<Scored>>> found(end - begin);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
        for (int64_t i = begin; i < (int64_t)end; ++i) {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            found[i - begin] = insert_search(D, G, prm, order[i], tls_scratch(), &cnts[t]);
        }
        std::vector<uint32_t> idx(end - begin);
        for (uint32_t i = 0; i < end - begin; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t c) { return order[begin + a] < order[begin + c]; });
        for (uint32_t i : idx)
            for (int l = 0; l < (int)found[i].size(); ++l) layer_insert(D, G, prm, order[begin + i], l, found[i][l], &cnts[0]);
        begin = end;
    }
    if (counters) {
        counters[0] = counters[1] = counters[2] = 0;
        for (auto& c : cnts) { counters[0] += c.n_dist; counters[1] += c.n_expand; counters[2] += c.n_edges_read; }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- Fssc (searcher.rs:150-199) as a tiny handle API ------------------------------------------
void* oracle_fssc_new(int size, int with_duplicates) { return new Fssc((size_t)size, with_duplicates != 0); }
void oracle_fssc_free(void* h) { delete (Fssc*)h; }
void oracle_fssc_add(void* h, const char* id, float score, uint32_t segment, uint32_t addr, const void* vec_bytes, int n_bytes) {
    ((Fssc*)h)->add(FsscItem{id, score, segment, addr}, std::string((const char*)vec_bytes, (size_t)n_bytes));
}
int oracle_fssc_result(void* h, uint32_t* segments, uint32_t* addrs, float* scores) {
    auto r = ((Fssc*)h)->sorted();
    for (size_t i = 0; i < r.size(); ++i) { segments[i] = r[i].segment; addrs[i] = r[i].addr; scores[i] = r[i].score; }
    return (int)r.size();
}

// ---- RaBitQ (rabitq.rs) ---------------------------------------------------------------------------
uint64_t oracle_rabitq_encoded_len(int d) { return rabitq_encoded_len(d); }
void oracle_rabitq_encode(const float* vecs, uint32_t n, int d, int ld, unsigned char* out, int nthreads) {
    size_t len = rabitq_encoded_len(d);
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) rabitq_encode(vecs + (size_t)i * ld, d, out + (size_t)i * len);
}
// estimate / error of every encoded vector for nq queries: out_* are [nq][n]
void oracle_rabitq_estimate(const unsigned char* enc, uint32_t n, int d, const float* queries, int nq, int qld, float* out_est, float* out_err,
                            int nthreads) {
    size_t len = rabitq_encoded_len(d);
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int qi = 0; qi < nq; ++qi) {
        RabitqQuery rq = RabitqQuery::from_vector(queries + (size_t)qi * qld, d);
        for (uint32_t v = 0; v < n; ++v) rq.similarity(enc + (size_t)v * len, out_est + (size_t)qi * n + v, out_err + (size_t)qi * n + v);
    }
}
// query side of the quantisation, for checking the kernel's preparation: planes [4][d/64] u64, scalars low, delta, sum_quantized
void oracle_rabitq_query(const float* q, int d, uint64_t* planes, float* low, float* delta, uint32_t* sum_quantized) {
    RabitqQuery rq = RabitqQuery::from_vector(q, d);
    for (int p = 0; p < 4; ++p) std::memcpy(planes + (size_t)p * (d / 64), rq.plane[p].data(), (size_t)d / 8);
    *low = rq.low; *delta = rq.delta; *sum_quantized = rq.sum_quantized;
}
// brute force with RaBitQ (segment.rs:581-608): estimate every vector, keep upper_bound >= min_score, rerank_top with exact dot
void oracle_rabitq_brute_force(const float* vecs, uint32_t n, int d, int ld, const unsigned char* enc, const float* queries, int nq, int qld, int k,
                               float min_score, uint32_t* out_ids, float* out_scores, int* out_count, uint64_t* out_exact_evals, int nthreads) {
    size_t len = rabitq_encoded_len(d);
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1)
    for (int qi = 0; qi < nq; ++qi) {
        const float* q = queries + (size_t)qi * qld;
        RabitqQuery rq = RabitqQuery::from_vector(q, d);
        std::vector<std::pair<uint32_t, float>> cand;
        for (uint32_t v = 0; v < n; ++v) {
            float est, err;
            rq.similarity(enc + (size_t)v * len, &est, &err);
            if (est + err >= min_score) cand.push_back({v, est + err});   // EstimatedScore::new_with_error: upper_bound = score + error
        }
        uint64_t evals = 0;
        auto r = rerank_top(cand, (size_t)k, min_score, [&](uint32_t v) { evals++; return dot_ordered(vecs + (size_t)v * ld, q, d); });
        if (out_exact_evals) out_exact_evals[qi] = evals;
        out_count[qi] = (int)r.size();
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j] = j < (int)r.size() ? r[j].id : NIL;
            out_scores[(size_t)qi * k + j] = j < (int)r.size() ? r[j].score : 0.0f;
        }
    }
}

// HNSW search with a quantised query (search.rs:306-383, RaBitQ branch).  Dot similarity only.  counters: [exact similarities,
// expansions, edges read, quantised estimates].
void oracle_hnsw_search_rabitq(const float* vecs, uint32_t n, int d, int ld, const unsigned char* enc, int M, int M0, const uint8_t* level,
                               uint32_t entry_node, uint32_t entry_layer, const uint32_t* adj0, const uint64_t* upper_off, const uint32_t* adjU,
                               const float* queries, int nq, int qld, int k, float min_score, int with_duplicates, const uint64_t* filter_bits,
                               uint32_t* out_ids, float* out_scores, int* out_count, uint64_t* counters, int nthreads) {
    Data D = make_data(vecs, nullptr, n, d, ld, SIM_DOT);
    GraphView G = make_view(n, M, M0, level, entry_node, entry_layer, const_cast<uint32_t*>(adj0), nullptr, upper_off,
                            const_cast<uint32_t*>(adjU), nullptr);
    int nt = nthreads > 0 ? nthreads : 1;
    std::vector<Counters> cnts(nt);
    std::vector<uint64_t> quant((size_t)nt * 8, 0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int qi = 0; qi < nq; ++qi) {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        Query q{queries + (size_t)qi * qld, 0.0f};
        NodeFilter f;
        f.filter_bits = filter_bits; f.with_duplicates = with_duplicates != 0;
        auto r = hnsw_search_rabitq(D, G, enc, q, (size_t)k, min_score, f, tls_scratch(), &cnts[t], &quant[(size_t)t * 8]);
        if (r.size() > (size_t)k) r.resize(k);
        out_count[qi] = (int)r.size();
        for (int j = 0; j < k; ++j) {
            out_ids[(size_t)qi * k + j] = j < (int)r.size() ? r[j].id : NIL;
            out_scores[(size_t)qi * k + j] = j < (int)r.size() ? r[j].score : 0.0f;
        }
    }
    if (counters) {
        counters[0] = counters[1] = counters[2] = counters[3] = 0;
        for (auto& c : cnts) { counters[0] += c.n_dist; counters[1] += c.n_expand; counters[2] += c.n_edges_read; }
        for (int t = 0; t < nt; ++t) counters[3] += quant[(size_t)t * 8];
    }
}

// ---- BM25 ---------------------------------------------------------------------------------------
uint8_t oracle_fieldnorm_to_id(uint32_t v) { return fieldnorm_to_id(v); }
uint32_t oracle_fieldnorm_id_to_value(uint32_t id) { return fieldnorm_id_to_value(id); }
float oracle_bm25_idf(uint64_t df, uint64_t n) { return bm25_idf(df, n); }
float oracle_bm25_term_score(uint64_t df, uint64_t n_docs, uint64_t total_tokens, uint32_t fieldnorm_id, uint32_t tf) {
    float cache[256];
    bm25_norm_cache((float)total_tokens / (float)n_docs, cache);
    return bm25_term_score(bm25_idf(df, n_docs) * (1.0f + BM25_K1), cache[fieldnorm_id], tf);
}
// nq queries, each query_terms[query_off[i] .. query_off[i+1]).  out_* are [nq][k].
void oracle_bm25_search(uint32_t n_docs, uint32_t n_terms, const uint64_t* term_off, const uint32_t* post_doc, const uint32_t* post_tf,
                        const uint8_t* fieldnorm_id, const uint64_t* alive_bits, uint64_t total_docs, uint64_t total_tokens,
                        const uint64_t* doc_freq, const uint32_t* query_terms, const uint32_t* query_off, int nq, int mode, int use_tf, int k,
                        uint32_t* out_docs, float* out_scores, int* out_count, uint64_t* out_total, int nthreads) {
    PostingsView P;
    P.n_docs = n_docs; P.n_terms = n_terms; P.term_off = term_off; P.doc = post_doc; P.tf = post_tf; P.fieldnorm_id = fieldnorm_id;
    P.alive_bits = alive_bits;
    Bm25Stats S;
    S.total_docs = total_docs; S.total_tokens = total_tokens; S.doc_freq = doc_freq;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int qi = 0; qi < nq; ++qi) {
        uint64_t total = 0;
        static thread_local Bm25Scratch scratch;
        auto r = bm25_search(P, S, query_terms + query_off[qi], (int)(query_off[qi + 1] - query_off[qi]), mode, use_tf != 0, (size_t)k, &total,
                             scratch);
        out_count[qi] = (int)r.size();
        if (out_total) out_total[qi] = total;
        for (int j = 0; j < k; ++j) {
            out_docs[(size_t)qi * k + j] = j < (int)r.size() ? r[j].doc : NIL;
            out_scores[(size_t)qi * k + j] = j < (int)r.size() ? r[j].score : 0.0f;
        }
    }
}

}  // extern "C"
