/* nidx_b200 — C ABI of the B200-native nidx search hot path (libnidx_b200.so).
 *
 * This is the drop-in boundary: every entry point replaces one Rust interface of the reference
 * (cited per function) and is what a cgo/FFI/ctypes binding on the reference side binds
 * (INTEGRATION.md shows the Rust `extern "C"` block).  Plain pointers and sizes only; the caller
 * owns every buffer it passes, the library owns the handles and all device memory.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NIDX_E* code on failure;
 *     nidx_last_error() returns the message of the calling thread's last failure
 *     (reference: anyhow::Error / VectorErr strings, nidx_vector/src/lib.rs:203-232).
 *   - `mem` arguments say where the caller's buffers live: NIDX_MEM_HOST (the library copies
 *     host<->device inside the call) or NIDX_MEM_DEVICE (pointers are device pointers on the
 *     index's GPU; nothing is copied).  `stream` is a cudaStream_t (NULL = default stream); calls
 *     with NIDX_MEM_DEVICE are asynchronous on it, calls with NIDX_MEM_HOST return after the
 *     results are in the host buffers.
 *   - search entry points are re-entrant: they may be called concurrently from many threads on
 *     one handle (reference: searchers are Sync+Send behind an Arc, index_cache.rs:41-47).
 *   - there is NO CPU fallback: without a CUDA device every call fails with NIDX_ENODEVICE.
 */
#ifndef NIDX_B200_H
#define NIDX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIDX_OK 0
#define NIDX_EINVAL (-1)     /* bad argument (VectorErr::InconsistentDimensions, ...) */
#define NIDX_ENODEVICE (-2)  /* no usable CUDA device */
#define NIDX_ECUDA (-3)      /* CUDA runtime error */
#define NIDX_EIO (-4)        /* segment file error */
#define NIDX_ESTATE (-5)     /* e.g. HNSW search on an index without a graph */
#define NIDX_EOVERFLOW (-6)  /* an internal bounded structure overflowed; results incomplete */

#define NIDX_MEM_HOST 0
#define NIDX_MEM_DEVICE 1

#define NIDX_SIM_DOT 0     /* config.rs:33-37 Similarity::Dot */
#define NIDX_SIM_COSINE 1  /* Similarity::Cosine */
#define NIDX_SIM_L2 2      /* extension (north_star): -|a - b|^2 as a similarity; the reference has no L2 (config.rs:33-37), so parity is
                              against the oracle's restatement and a float64 brute force only */

#define NIDX_METHOD_AUTO 0   /* segment.rs:538 use_hnsw() cost model decides */
#define NIDX_METHOD_HNSW 1   /* hnsw/search.rs:306-383 */
#define NIDX_METHOD_BRUTE 2  /* segment.rs:569-623 */
#define NIDX_METHOD_BRUTE_RABITQ 3 /* segment.rs:581-608 with a RaBitQ query: quantised scan + exact rerank (rabitq.rs:222-244) */
#define NIDX_METHOD_HNSW_RABITQ 4  /* hnsw/search.rs:306-383 with a RaBitQ query: the walk ranks by the estimate, layer 0 returns
                                      min(100 k, 2000) nodes, rerank_top + closest_up_nodes on exact similarities.  What AUTO takes
                                      when the segment carries codes (segment.rs:506-513) and the cost model picks the graph */

#define NIDX_NIL 0xFFFFFFFFu

const char* nidx_last_error(void);
/* Number of CUDA devices the library can use (0 => every other call fails). */
int nidx_device_count(void);
/* Kernel launches issued by this process through the library (bench.py's gpu_launches). */
uint64_t nidx_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Vector segment  (reference: nidx_vector OpenSegment + VectorConfig, segment.rs / config.rs)
 * ------------------------------------------------------------------------------------------ */
typedef struct nidx_vec_segment nidx_vec_segment;

typedef struct nidx_vec_config {
    int32_t dimension;        /* VectorType::DenseF32{dimension}, config.rs:102-124 */
    int32_t similarity;       /* NIDX_SIM_* */
    int32_t multi_vector;     /* VectorCardinality::Multi: one result per paragraph */
    int32_t m;                /* params.rs:40 M (also M_MAX);   0 => 30 */
    int32_t m0;               /* params.rs:34 M_MAX_0;          0 => 60 */
    int32_t ef_construction;  /* params.rs:43;                  0 => 100 */
    int32_t ef_search;        /* params.rs:46;                  0 => 30 */
    int32_t device;           /* CUDA device ordinal */
} nidx_vec_config;

/* segment::create's data-store half (segment.rs:199-239, data_store/v2.rs:54-80): take n vectors
 * ([n][ld] f32, ld >= dimension) and, optionally, the paragraph address of every vector
 * (vectors of one paragraph must be contiguous; NULL = one vector per paragraph).  Vectors are
 * re-laid out in HBM as [n][ld4] rows (16-byte aligned) and their norms are precomputed.
 * No graph yet: brute force works, HNSW needs nidx_vec_build_hnsw or nidx_vec_set_graph. */
int nidx_vec_create(const nidx_vec_config* cfg, const float* vectors, uint64_t n, int32_t ld, int mem, const uint32_t* paragraph_of,
                    nidx_vec_segment** out);

/* Open a segment directory written by the reference (or by nidx_vec_save): vectors.bin
 * (data_store/v2/vector_store.rs:33-68) + hnsw.graph (hnsw/disk/v2.rs:16-49) [+ hnsw.edges].
 * Replaces segment::open (segment.rs:39-90) for the hot path's needs. */
int nidx_vec_open(const nidx_vec_config* cfg, const char* dir, nidx_vec_segment** out);
/* Write vectors.bin / hnsw.graph / hnsw.edges in the reference's formats (DiskHnswV2::serialize_to,
 * hnsw/disk/v2.rs:213-218; VectorStoreWriter, vector_store.rs:113-146). */
int nidx_vec_save(nidx_vec_segment* seg, const char* dir);
void nidx_vec_close(nidx_vec_segment* seg);

uint64_t nidx_vec_len(const nidx_vec_segment* seg);
/* Device pointers of the resident data, for zero-copy consumers (bench, tests). */
const float* nidx_vec_device_vectors(const nidx_vec_segment* seg, int32_t* ld_out);

/* HnswBuilder (hnsw/build.rs:36-166) on the GPU: levels for all nodes first (initialize_graph),
 * then batch-synchronous insertion (DESIGN.md "build"): batches of at most max_batch nodes search
 * the frozen graph and are linked in ascending id.  seed: level RNG seed (reference uses 2). */
int nidx_vec_build_hnsw(nidx_vec_segment* seg, uint64_t seed, int32_t max_batch, void* stream);

/* The top layer of every node as HnswBuilder::initialize_graph draws them (build.rs:40,49-55,97-101: SmallRng seeded with
 * `seed` (the reference uses 2), level = round(-ln(u) / ln(M))), capped at the library's layer limit.  nidx_vec_build_hnsw uses
 * exactly these.  Pure host function: needs no device. */
int nidx_hnsw_levels(uint64_t n, int32_t m, uint64_t seed, uint8_t* out_level);

/* utils::normalize_vector (nidx_vector/src/utils.rs:20-23) for n rows of d floats ([n][ld], in place): x / sqrt(fold(acc + x*x)),
 * the fold sequential in f32 exactly as the reference's iterator (bit-identical results).  Used at index time when
 * VectorConfig.normalize_vectors is set (indexer.rs:94-146) and on the query (searcher.rs:246-252).  mem = NIDX_MEM_HOST copies
 * in and out and waits; NIDX_MEM_DEVICE works in place on `stream`. */
int nidx_normalize_vectors(int32_t device, float* vectors, uint64_t n, int32_t d, int32_t ld, int mem, void* stream);

/* The planner's cost model (use_hnsw, segment.rs:626-660): 1 if the HNSW walk is estimated cheaper than the exhaustive scan
 * for `matching_nodes` of `total_nodes` paragraphs passing the filter.  has_rabitq = the segment carries 1-bit codes.  m = the
 * graph's M (the reference's compile-time hnsw::M = 30).  Pure host function: needs no device.  nidx_vec_search applies it for
 * NIDX_METHOD_AUTO. */
int nidx_use_hnsw(uint64_t total_nodes, uint64_t matching_nodes, uint64_t top_k, int has_rabitq, int m);

/* merge_indexes' fast path (segment.rs:143-167): the first n_existing vectors of this segment already have a graph
 * (the largest input segment of a merge, without deletions) given in the flat layout below for n_existing nodes; only the
 * remaining vectors are inserted.  Levels of the new nodes come from a fresh RNG (HnswBuilder::new + initialize_graph with
 * skip_nodes = n_existing, build.rs:36-55); the entry point moves only if a higher layer appears (ram_hnsw.rs:99-107).
 * The edge similarities (w0, wU: the contents of hnsw.edges) are required -- the reverse-link prune ranks by them.  Links
 * in a layer > 0 to a node that is not in that layer are dropped first (fix_broken_graph, ram_hnsw.rs:118-123). */
int nidx_vec_extend_hnsw(nidx_vec_segment* seg, uint64_t n_existing, const uint8_t* level_existing, const uint32_t* adj0, const float* w0,
                         const uint32_t* adjU, const float* wU, uint32_t entry_node, uint32_t entry_layer, uint64_t seed, int32_t max_batch,
                         void* stream);

/* Flat graph import / export (the layout in DESIGN.md; the oracle uses the same one).
 * level[n] u8; adj0[n][s0] u32, adjU[rows][su] u32, NIDX_NIL padded; w0/wU edge similarities
 * (may be NULL on import: search does not need them, merge/build does). Host pointers. */
int nidx_vec_graph_dims(const nidx_vec_segment* seg, int32_t* s0, int32_t* su, uint64_t* upper_rows, uint32_t* entry_node,
                        uint32_t* entry_layer);
int nidx_vec_set_graph(nidx_vec_segment* seg, const uint8_t* level, const uint32_t* adj0, const float* w0, const uint32_t* adjU,
                       const float* wU);
int nidx_vec_get_graph(const nidx_vec_segment* seg, uint8_t* level, uint32_t* adj0, float* w0, uint32_t* adjU, float* wU);

/* OpenSegment::apply_deletions (segment.rs:428-445): bit per paragraph, 1 = alive; NULL = all alive. */
int nidx_vec_set_alive(nidx_vec_segment* seg, const uint64_t* alive_bits, int mem);

typedef struct nidx_vec_search_params {
    int32_t k;                /* VectorSearchRequest.result_per_page (request_types.rs:18-35) */
    int32_t ef;               /* layer-0 width = max(k, ef) (hnsw/search.rs:338-345); 0 => config */
    float min_score;          /* VectorSearchRequest.min_score */
    int32_t with_duplicates;  /* VectorSearchRequest.with_duplicates */
    int32_t method;           /* NIDX_METHOD_* */
    const uint64_t* filter_bits; /* filter formula evaluated to a bitset over paragraphs
                                    (segment.rs:516-534); NULL = no filter. Same `mem` as queries. */
    uint64_t filter_matching; /* number of set bits in filter_bits ∧ alive (segment.rs:531), only read
                                 by the NIDX_METHOD_AUTO cost model; 0 = unknown (count on device) */
} nidx_vec_search_params;

/* OpenSegment::search (segment.rs:477-567) for a batch of nq queries ([nq][ldq] f32).
 * out_ids/out_scores are [nq][k] (vector address + similarity, descending; NIDX_NIL padded),
 * out_counts[nq] the number of valid results per query. */
int nidx_vec_search(nidx_vec_segment* seg, const float* queries, int32_t nq, int32_t ldq, int mem, const nidx_vec_search_params* p,
                    uint32_t* out_ids, float* out_scores, int32_t* out_counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Filters on the device (reference: ParagraphInvertedIndexes, inverted_index/paragraph.rs:39-186 over fst_index.rs + map.rs)
 * ------------------------------------------------------------------------------------------ */
#define NIDX_INV_LABELS 0  /* label index: key = labels_key(label) (paragraph.rs:64-66), looked up by PREFIX (get_prefix) */
#define NIDX_INV_FIELDS 1  /* field index: key = FieldKey bytes (utils.rs:80-117), looked up EXACTLY (get) */
/* One inverted index of the segment: n_keys byte strings, strictly ascending in memcmp order (the fst's order), key i =
 * key_bytes[key_off[i] .. key_off[i + 1]), its paragraph addresses = postings[post_off[i] .. post_off[i + 1]).  The keys stay on the
 * host side of the library (the lookup is the fst's job: a binary search), the postings live in HBM.  Host pointers. */
int nidx_vec_set_inverted_index(nidx_vec_segment* seg, int32_t which, uint32_t n_keys, const uint8_t* key_bytes, const uint64_t* key_off,
                                const uint64_t* post_off, const uint32_t* postings);

#define NIDX_F_LABEL 0  /* AtomClause::Label: n = 1 key, every label key that starts with it (formula.rs:21, paragraph.rs:140-142) */
#define NIDX_F_KEYS 1   /* AtomClause::KeyPrefixSet: n field keys, each looked up exactly (paragraph.rs:143-147) */
#define NIDX_F_AND 2    /* CompoundClause And: intersection of the n operand nodes that follow */
#define NIDX_F_OR 3     /* CompoundClause Or: union */
#define NIDX_F_NOT 4    /* CompoundClause Not: complement of the INTERSECTION of its n operands (paragraph.rs:160-178) */
typedef struct nidx_filter_node {   /* a Formula / Clause tree in pre-order (formula.rs:40-100); a Formula with several clauses is an AND / OR root */
    int32_t kind;                   /* NIDX_F_* */
    int32_t n;                      /* LABEL: 1; KEYS: number of keys; AND / OR / NOT: number of operand subtrees that follow */
    const uint8_t* const* keys;     /* LABEL / KEYS: n byte strings (host pointers) */
    const uint32_t* key_len;
} nidx_filter_node;

/* ParagraphInvertedIndexes::filter + the intersection with the alive set (segment.rs:516-531): the formula's bitset over the
 * paragraphs, computed in HBM (postings -> bits, bit algebra, popcount).  out_bits (`mem`; (paragraphs + 63) / 64 words) may be NULL;
 * *out_matching = number of set bits (the reference's `bitset.iter().count()`). */
int nidx_vec_filter(nidx_vec_segment* seg, const nidx_filter_node* nodes, int32_t n_nodes, uint64_t* out_bits, int mem, uint64_t* out_matching, void* stream);

/* nidx_vec_search with the filter given as a formula: evaluated on the device and fed to the search without a host round trip
 * (p->filter_bits must be NULL; NIDX_METHOD_AUTO reads the match count back, 8 bytes, for the cost model as segment.rs:531 does). */
int nidx_vec_search_formula(nidx_vec_segment* seg, const float* queries, int32_t nq, int32_t ldq, int mem, const nidx_vec_search_params* p,
                            const nidx_filter_node* nodes, int32_t n_nodes, uint32_t* out_ids, float* out_scores, int32_t* out_counts, void* stream);

/* Searcher::_search's cross-segment / cross-shard top-k (searcher.rs:241-290 Fssc without the string
 * keys, shard_merge.rs:332-348): merge n_parts partial results [n_parts][nq][k] (score desc) into
 * [nq][k]; out_part[nq][k] receives the index of the part each winner came from.  part_stride = elements
 * between consecutive parts in ids / scores (0 = nq*k, i.e. dense), so an all-gather buffer can be merged in place.
 * Device pointers. */
int nidx_merge_topk(int32_t device, const uint32_t* ids, const float* scores, int32_t n_parts, int64_t part_stride, int32_t nq, int32_t k,
                    uint32_t* out_ids, float* out_scores, int32_t* out_part, void* stream);

/* Counters of the last HNSW search / build on this segment (for the roofline accounting,
 * SURVEY 8d): [0] similarity evaluations, [1] node expansions, [2] visited-set overflows.  Every search call counts into its
 * own workspace, so concurrent searches never mix their counts; "last" = the call that was issued last. */
int nidx_vec_counters(nidx_vec_segment* seg, uint64_t out[3]);
/* The same with the quantised walk's: [0] exact similarities computed, [1] expansions, [2] visited-set overflows, [3] closest_up
 * overflows, [4] RaBitQ estimates, [5] exact similarities the sequential rerank_top needed (<= the share of [0] spent there). */
int nidx_vec_counters_ex(nidx_vec_segment* seg, uint64_t out[6]);

/* RaBitQ 1-bit codes (vector_types/rabitq.rs; Dot similarity and dimension % 64 == 0 only, config.rs:170-173).
 * nidx_vec_rabitq_encode builds the reference's vectors.quant records ([f32 dot_quant_original][u32 sum_bits][dim/8 sign
 * bits], quant_vector_store.rs:29,57-60) in HBM; nidx_vec_rabitq_codes copies them out ([n][dim/8 + 8] bytes, host);
 * nidx_vec_rabitq_estimate evaluates QueryVector::similarity (estimate, error bound; rabitq.rs:202-218) of every stored
 * vector for nq queries into out_estimate / out_error [nq][n] (same `mem` as the queries).  NIDX_METHOD_BRUTE_RABITQ in
 * and NIDX_METHOD_HNSW_RABITQ in nidx_vec_search need the codes. */
int nidx_vec_rabitq_encode(nidx_vec_segment* seg, void* stream);
int nidx_vec_rabitq_codes(const nidx_vec_segment* seg, uint8_t* out_codes);
int nidx_vec_rabitq_estimate(nidx_vec_segment* seg, const float* queries, int32_t nq, int32_t ldq, int mem, float* out_estimate, float* out_error,
                             void* stream);

/* Device time (CUDA events on the caller's stream) of the dominant kernel of the last nidx_vec_search on
 * this segment: hnsw_search_kernel, or the first scan_scores_kernel of a brute-force call.  Diagnostics
 * for the roofline line of bench.py; not meaningful under concurrent searches. */
int nidx_vec_last_kernel_ms(nidx_vec_segment* seg, float* ms);

/* ------------------------------------------------------------------------------------------
 * Text segment: BM25 over device-resident postings
 * (reference: tantivy TopDocs::order_by_score called at nidx_text/src/reader.rs:433-435 and
 *  nidx_paragraph/src/reader.rs:290-292; statistics over the union of segments,
 *  nidx_tantivy/src/index_reader.rs:39-77)
 * ------------------------------------------------------------------------------------------ */
typedef struct nidx_txt_segment nidx_txt_segment;

#define NIDX_BM25_OR 0   /* nidx_paragraph keyword query: Occur::Should (keyword_parser.rs:62-67) */
#define NIDX_BM25_AND 1  /* nidx_text: QueryParser::set_conjunction_by_default (reader.rs:372-377) */

/* Postings in CSR form: term_off[n_terms+1], post_doc/post_tf[term_off[n_terms]] (doc ids ascending
 * per term), fieldnorm_id[n_docs] (tantivy's 1-byte fieldnorm code).  Host pointers; copied to HBM. */
int nidx_txt_create(int32_t device, uint32_t n_docs, uint32_t n_terms, const uint64_t* term_off, const uint32_t* post_doc,
                    const uint32_t* post_tf, const uint8_t* fieldnorm_id, nidx_txt_segment** out);
/* Collection statistics of the whole index (all segments, all GPUs): total docs, total tokens,
 * doc_freq[n_terms].  Defaults to the segment's own statistics. */
int nidx_txt_set_stats(nidx_txt_segment* seg, uint64_t total_docs, uint64_t total_tokens, const uint64_t* doc_freq);
int nidx_txt_set_alive(nidx_txt_segment* seg, const uint64_t* alive_bits);
void nidx_txt_close(nidx_txt_segment* seg);

typedef struct nidx_txt_search_params {
    int32_t k;       /* result_per_page + 1 in the reference (reader.rs:386-387) */
    int32_t mode;    /* NIDX_BM25_* */
    int32_t use_tf;  /* 0: IndexRecordOption::Basic (tf == 1), 1: real term frequencies */
    float min_score; /* results below are dropped after top-k (reader.rs:302-305) */
    /* nidx_paragraph search-after (reader.rs:350-392 build_topdocs_search_after_collector / is_after): documents that
     * are not "after" (after_score, tie break) are scored -inf by the reference's tweak_score; here they are counted
     * in out_total but never enter the top-k. */
    int32_t after_mode;      /* 0 = no search_after; 1 = SearchAfterTieBreak::Drop; 2 = KeepAfter(after_docaddr); 3 = Keep */
    float after_score;       /* SearchAfter.score */
    uint64_t after_docaddr;  /* KeepAfter payload */
    uint64_t docaddr_base;   /* segment_ord << 32: docaddr = docaddr_base + doc (reader.rs:366-371) */
} nidx_txt_search_params;

/* nq queries; query i is query_terms[query_off[i] .. query_off[i+1]) (term ids).
 * out_docs/out_scores [nq][k] (score desc, then doc asc), out_counts[nq], out_total[nq] = number of
 * matching documents (the Count collector, reader.rs:433).  `mem` applies to queries and outputs. */
int nidx_txt_search(nidx_txt_segment* seg, const uint32_t* query_terms, const uint32_t* query_off, int32_t nq, int mem,
                    const nidx_txt_search_params* p, uint32_t* out_docs, float* out_scores, int32_t* out_counts, uint64_t* out_total,
                    void* stream);

/* Device time (CUDA events on the caller's stream) of bm25_kernel in the last nidx_txt_search on this segment (bench roofline);
 * not meaningful under concurrent searches. */
int nidx_txt_last_kernel_ms(nidx_txt_segment* seg, float* ms);

/* ------------------------------------------------------------------------------------------
 * Segments sharded over the GPUs of one node: one process (or thread) per GPU, one segment each
 * (reference: the searcher's scatter-gather, nidx/src/searcher/grpc.rs:253-431, merged by
 *  shard_merge.rs:332-348 / 177-231; inside one index the cross-segment collection Fssc, nidx_vector/src/searcher.rs:150-199)
 * ------------------------------------------------------------------------------------------ */
typedef struct nidx_shard_comm nidx_shard_comm;

/* ncclGetUniqueId: rank 0 creates the 128-byte id and hands it to the other ranks by whatever channel the host has
 * (the reference's searchers already know each other through their gRPC addresses). */
int nidx_shard_unique_id(uint8_t out_id[128]);
/* ncclCommInitRank on `device`; collective: every rank of `world` must call it with the same id.  NCCL is bound at run time
 * (libnccl.so.2); without it these entry points fail with NIDX_ESTATE and everything else keeps working. */
int nidx_shard_init(const uint8_t unique_id[128], int32_t rank, int32_t world, int32_t device, nidx_shard_comm** out);
void nidx_shard_destroy(nidx_shard_comm* comm);

/* Paragraph keys for the cross-segment de-duplication: keys[p] identifies paragraph p's id across segments (the reference keys
 * Fssc by the paragraph id string, searcher.rs:62-90: pass a 64-bit hash of it).  NULL = (rank, paragraph address). Host pointer. */
int nidx_vec_set_paragraph_keys(nidx_vec_segment* seg, const uint64_t* keys);

/* OpenSegment::search on this rank's segment + exchange + merge, identical results on every rank.  Collective: every rank calls
 * it with the same queries, nq, k and dedup, in the same order (one call at a time per communicator).
 *   dedup = 0: the parts are shards -- merge_vector_responses (kmerge by score, shard_merge.rs:332-348);
 *   dedup = 1: the parts are segments of ONE index -- Fssc (searcher.rs:150-199): one entry per paragraph key, and with
 *              p->with_duplicates == 0 byte-identical vectors are suppressed across segments (by a 64-bit hash of the bytes).
 * out_ids are vector addresses local to the part in out_part[nq][k] (-1 = none); out_counts[nq] may be NULL.
 * `mem` applies to queries and outputs; device calls are asynchronous on `stream`. */
int nidx_vec_search_sharded(nidx_shard_comm* comm, nidx_vec_segment* seg, const float* queries, int32_t nq, int32_t ldq, int mem,
                            const nidx_vec_search_params* p, int32_t dedup, uint32_t* out_ids, float* out_scores, int32_t* out_part, int32_t* out_counts,
                            void* stream);
/* BM25 over a document-partitioned index (every part scores with the statistics of the whole index, nidx_txt_set_stats):
 * merge_document_responses' order (bm25 desc, part asc, doc asc; shard_merge.rs:227-231); out_total = Count over all parts. */
int nidx_txt_search_sharded(nidx_shard_comm* comm, nidx_txt_segment* seg, const uint32_t* query_terms, const uint32_t* query_off, int32_t nq, int mem,
                            const nidx_txt_search_params* p, uint32_t* out_docs, float* out_scores, int32_t* out_part, int32_t* out_counts, uint64_t* out_total,
                            void* stream);

/* ------------------------------------------------------------------------------------------
 * One shard search as ONE device-side plan + rank fusion on the device
 * (reference: run_index_searches, nidx/src/searcher/shard_search.rs:176-241 -- the paragraph, document and vector searches of a
 *  request run on scoped threads; the ranked lists are fused afterwards in Python,
 *  nucliadb/src/nucliadb/search/search/rank_fusion.py:78-186)
 * ------------------------------------------------------------------------------------------ */
/* Caller keys of a text segment's documents (for the paragraph index: the paragraph id, as a 64-bit hash or table index -- the
 * same key space as nidx_vec_set_paragraph_keys), used to match keyword and semantic results.  NULL = the document number. */
int nidx_txt_set_doc_keys(nidx_txt_segment* seg, const uint64_t* keys);

typedef struct nidx_rrf_source {
    const uint64_t* keys;     /* [nq][k] item keys, best first (every source sorted by its score, descending); ~0 = no item */
    const float* scores;      /* [nq][k] the source's scores: reported as they are when only one source has results (rank_fusion.py:86-89) */
    const int32_t* counts;    /* [nq] valid items per query; NULL = k minus trailing ~0 keys */
    int32_t k;
    double weight;            /* the retriever's boost w(r) (rank_fusion.py:133-141) */
} nidx_rrf_source;

/* ReciprocalRankFusion.fuse (rank_fusion.py:78-96, 143-186) for nq queries: score(d) = sum over the sources, in the order given, of
 * 1 / (k + rank) * weight in IEEE double arithmetic (bit-identical to the reference's Python floats); one output item per key (the
 * first occurrence), sorted by score descending, ties in first-insertion order (Python's stable sort).  Rows of out_* are
 * sum(k_i) long: out_refs = first occurrence's source << 28 | mask of contributing sources << 24 | its position in that source;
 * out_counts[nq] = number of fused items.  At most 4 sources. */
int nidx_rank_fusion_rrf(int32_t device, const nidx_rrf_source* sources, int32_t n_sources, int32_t nq, double k, int mem, uint64_t* out_keys,
                         double* out_scores, uint32_t* out_refs, int32_t* out_counts, void* stream);

typedef struct nidx_shard_search_request {
    int32_t nq;
    /* vectors_request (shard_search.rs:211-213): vec == NULL = not requested */
    nidx_vec_segment* vec; const float* queries; int32_t ldq; const nidx_vec_search_params* vec_params;
    const nidx_filter_node* formula; int32_t n_formula;     /* optional filter formula evaluated on the device (nidx_vec_search_formula) */
    /* paragraphs_request (shard_search.rs:189-191): the keyword search, BM25 over the paragraph index */
    nidx_txt_segment* par; const uint32_t* par_terms; const uint32_t* par_off; const nidx_txt_search_params* par_params;
    /* texts_request (shard_search.rs:185-187): BM25 over the document (field) index */
    nidx_txt_segment* doc; const uint32_t* doc_terms; const uint32_t* doc_off; const nidx_txt_search_params* doc_params;
    /* rank fusion of the paragraph (keyword) and vector (semantic) lists; rrf_k <= 0: none */
    double rrf_k, weight_keyword, weight_semantic;
    int32_t semantic_first;   /* order of the sources (the reference iterates a dict: insertion order decides ties and which item object survives) */
} nidx_shard_search_request;

typedef struct nidx_shard_search_response {
    uint32_t* vec_ids; float* vec_scores; int32_t* vec_counts;                          /* [nq][vec_params->k] as nidx_vec_search */
    uint32_t* par_docs; float* par_scores; int32_t* par_counts; uint64_t* par_total;    /* [nq][par_params->k] as nidx_txt_search */
    uint32_t* doc_docs; float* doc_scores; int32_t* doc_counts; uint64_t* doc_total;    /* [nq][doc_params->k] */
    uint64_t* fused_keys; double* fused_scores; uint32_t* fused_refs; int32_t* fused_counts;   /* [nq][kv + kp] as nidx_rank_fusion_rrf */
} nidx_shard_search_response;

/* run_index_searches for a batch of nq requests against one shard's indexes: the requested searches run concurrently on side
 * streams forked from `stream`, are joined back, and (rrf_k > 0, vector and paragraph requests present) the two ranked lists
 * are fused on the device -- keys through nidx_vec_set_paragraph_keys / nidx_txt_set_doc_keys.  `mem` applies to all inputs
 * and outputs; with host buffers the call returns when the results are in them (one synchronisation at the end). */
int nidx_shard_search(const nidx_shard_search_request* req, nidx_shard_search_response* resp, int mem, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NIDX_B200_H */
