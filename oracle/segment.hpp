// ORACLE — TEST INFRASTRUCTURE ONLY (see distance.hpp).
//
// CPU restatement of the per-segment and cross-segment logic around the two search paths:
//   nidx/nidx_vector/src/segment.rs:569-623   brute_force_search (best vector per alive paragraph,
//                                             keep iff upper_bound >= min_score, sort desc, take k)
//   nidx/nidx_vector/src/segment.rs:626-660   use_hnsw cost model
//   nidx/nidx_vector/src/searcher.rs:150-199  Fssc (cross-segment fixed-size sorted collection)
//   nidx/src/searcher/shard_merge.rs:332-348  merge_vector_responses (kmerge_by score >=)
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "hnsw.hpp"

namespace nidx_oracle {

// segment.rs:569-623, dense-f32 branch.  Paragraph p owns vectors [first_vec[p], first_vec[p]+num_vec[p]);
// first_vec == nullptr means one vector per paragraph (vector addr == paragraph addr).
// Returns (vector addr, score) of the best vector of each of the top-k paragraphs.
static inline std::vector<Scored> brute_force_search(const Data& D, const Query& q, size_t top_k, float min_score, const uint64_t* alive_bits,
                                                     uint32_t n_paragraphs, const uint32_t* first_vec, const uint32_t* num_vec) {
    std::vector<Scored> scored;
    for (uint32_t p = 0; p < n_paragraphs; ++p) {
        if (alive_bits && !((alive_bits[p >> 6] >> (p & 63)) & 1)) continue;  // bitset.iter()
        uint32_t v0 = first_vec ? first_vec[p] : p, nv = num_vec ? num_vec[p] : 1;
        bool have = false;
        Scored best{0, 0.0f};
        for (uint32_t v = v0; v < v0 + nv; ++v) {  // max_by total_cmp; later element wins ties in Rust's max_by
            float s = similarity(D.sim, D.vec(v), D.nrm(v), q.q, q.qnorm, D.d);
            if (!have || ordered_bits(s) >= ordered_bits(best.score)) { best = {v, s}; have = true; }
        }
        if (have && best.score >= min_score) scored.push_back(best);  // 594: >=
    }
    std::sort(scored.begin(), scored.end(), better);  // 611 sort_unstable_by desc; ties -> lower id
    if (scored.size() > top_k) scored.resize(top_k);
    return scored;
}

// segment.rs:626-660.
static inline bool use_hnsw(size_t total_nodes, size_t matching_nodes, size_t top_k, bool has_rabitq, size_t M = 30) {
    size_t full_cost, search_mult, rerank_mult;
    const size_t RERANKING_FACTOR = 100;  // rabitq.rs:34
    if (has_rabitq) { full_cost = 16; search_mult = RERANKING_FACTOR * 3 / 4; rerank_mult = RERANKING_FACTOR / 2; }
    else { full_cost = 1; search_mult = 1; rerank_mult = 0; }
    float l = std::log((float)total_nodes) - 2.0f;
    float hnsw_rq = l * l * std::log((float)top_k) * (float)search_mult;
    size_t hnsw_full = (top_k * rerank_mult) + (top_k * M * total_nodes / matching_nodes);
    size_t bf_rq = matching_nodes;
    size_t bf_full = top_k * rerank_mult;
    size_t hnsw_cost = (size_t)(hnsw_rq < 0 ? 0 : hnsw_rq) + hnsw_full * full_cost;  // `as usize` saturates negatives to 0
    size_t bf_cost = bf_rq + bf_full * full_cost;
    return hnsw_cost < bf_cost;
}

// searcher.rs:150-199.  Candidates are identified by a paragraph-id string (hash/eq on id) and
// carry the raw vector bytes for the with_duplicates=false test.
struct FsscItem {
    std::string id;
    float score;
    uint32_t segment, addr;
};
struct Fssc {
    size_t size;
    bool with_duplicates;
    std::vector<std::string> seen;   // vector bytes already offered
    std::vector<FsscItem> buff;      // set keyed by id
    Fssc(size_t size_, bool wd) : size(size_), with_duplicates(wd) {}
    void add(const FsscItem& cand, const std::string& vector_bytes) {
        if (!with_duplicates) {
            for (auto& s : seen) if (s == vector_bytes) return;
            seen.push_back(vector_bytes);
        }
        auto same = std::find_if(buff.begin(), buff.end(), [&](const FsscItem& b) { return b.id == cand.id; });
        if (buff.size() == size) {  // is_full
            // smallest element that is smaller than the candidate
            int pick = -1;
            for (int i = 0; i < (int)buff.size(); ++i)
                if (cand.score > buff[i].score && (pick < 0 || buff[i].score < buff[pick].score)) pick = i;
            if (pick >= 0) {
                buff.erase(buff.begin() + pick);
                // HashSet::insert does not replace an equal (same id) element that is still present
                same = std::find_if(buff.begin(), buff.end(), [&](const FsscItem& b) { return b.id == cand.id; });
                if (same == buff.end()) buff.push_back(cand);
            }
        } else if (same == buff.end()) {
            buff.push_back(cand);
        }
    }
    std::vector<FsscItem> sorted() const {  // From<Fssc> for Vec: sort desc by score
        std::vector<FsscItem> r = buff;
        std::stable_sort(r.begin(), r.end(), [](const FsscItem& a, const FsscItem& b) { return a.score > b.score; });
        return r;
    }
};

}  // namespace nidx_oracle
