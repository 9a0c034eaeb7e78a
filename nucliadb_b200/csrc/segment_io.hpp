// nidx_b200 — readers / writers for the reference's on-disk segment files (host side).
//
//   vectors.bin   nidx/nidx_vector/src/data_store/v2/vector_store.rs:33-68,113-146
//                 records [dim x f32 LE][paragraph_addr u32 LE], no padding for f32 (alignment 4).
//   hnsw.graph    nidx/nidx_vector/src/hnsw/disk/v2.rs:16-49 (format), 122-157 (serialize_node),
//                 176-211 (serialize_into), 248-312 (deserialize)
//   hnsw.edges    edge similarities as f32 LE in the order the edges appear in hnsw.graph.
// The in-memory side is the flat graph layout of DESIGN.md (adj0[n][s0], upper pool, NIL padded).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace segio {

struct FlatGraph {
    uint64_t n = 0;
    int s0 = 0, su = 0;
    uint32_t entry_node = 0, entry_layer = 0;
    uint64_t upper_rows = 0;
    std::vector<uint8_t> level;
    std::vector<uint32_t> adj0, adjU;
    std::vector<float> w0, wU;
};

inline bool read_file(const std::string& path, std::vector<unsigned char>& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = sz > 0 ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != out.size()) { err = "short read on " + path; return false; }
    return true;
}
inline bool write_file(const std::string& path, const void* data, size_t bytes, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot create " + path; return false; }
    size_t put = bytes ? fwrite(data, 1, bytes, f) : 0;
    fclose(f);
    if (put != bytes) { err = "short write on " + path; return false; }
    return true;
}

inline uint32_t rd32(const unsigned char* p) { uint32_t v; memcpy(&v, p, 4); return v; }

inline bool write_vectors_bin(const std::string& path, const float* vecs, uint64_t n, int d, int ld, const uint32_t* paragraph_of, std::string& err) {
    size_t rec = (size_t)d * 4 + 4;
    std::vector<unsigned char> buf((size_t)n * rec);
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(buf.data() + i * rec, vecs + (size_t)i * ld, (size_t)d * 4);
        memcpy(buf.data() + i * rec + (size_t)d * 4, &paragraph_of[i], 4);
    }
    return write_file(path, buf.data(), buf.size(), err);
}

// hnsw/disk/v2.rs:248-312.  A node is in layer l > 0 iff it has edges there (as the reference's
// deserialize decides); its level is the highest such layer.
inline bool parse_graph_v2(const std::vector<unsigned char>& g, const std::vector<unsigned char>& edges, uint64_t n, int s0, int su, int max_layers,
                           FlatGraph& out, std::string& err) {
    if (g.size() < 8 + 4 * n) { err = "file too short"; return false; }
    size_t end = g.size();
    out.n = n; out.s0 = s0; out.su = su;
    out.entry_layer = rd32(g.data() + end - 8);
    out.entry_node = rd32(g.data() + end - 4);
    out.level.assign(n, 0);
    struct Ref { uint32_t node, layer; size_t start; uint32_t count; };
    std::vector<Ref> refs;
    for (uint64_t node = 0; node < n; ++node) {
        size_t indexing_pos = end - (node + 3) * 4;
        size_t node_end = rd32(g.data() + indexing_pos);
        if (node_end > end) { err = "node offset out of range"; return false; }
        for (uint32_t layer = 0;; ++layer) {
            size_t layer_pos = node_end - (size_t)(layer + 1) * 4;
            if (layer_pos >= end) { err = "layer offset out of range"; return false; }
            size_t off = rd32(g.data() + layer_pos);
            if (off > node_end) { err = "edge offset out of range"; return false; }
            size_t start = node_end - off;
            uint32_t cnt = rd32(g.data() + start);
            size_t cnx_end = start + 4 + (size_t)cnt * 4;
            if (cnx_end > end) { err = "edge list out of range"; return false; }
            if (layer == 0 || cnt > 0) {
                if ((int)layer >= max_layers) { err = "too many layers"; return false; }
                if (cnt > (uint32_t)(layer == 0 ? s0 : su)) { err = "node has more edges than the configured M/M0 allows"; return false; }
                refs.push_back({(uint32_t)node, layer, start + 4, cnt});
                if (cnt > 0 && layer > out.level[node]) out.level[node] = (uint8_t)layer;
            }
            if (cnx_end == layer_pos) break;
        }
    }
    // A node is also in layer l if something links to it there (its own list may be empty), and the
    // entry point is in the entry layer even when it is alone in it.
    for (const Ref& r : refs)
        for (uint32_t e = 0; e < r.count; ++e) {
            uint32_t t = rd32(g.data() + r.start + (size_t)e * 4);
            if (t >= n) { err = "edge target out of range"; return false; }
            if (r.layer > out.level[t]) out.level[t] = (uint8_t)r.layer;
        }
    if (n && out.entry_node < n && (int)out.entry_layer < max_layers && out.entry_layer > out.level[out.entry_node])
        out.level[out.entry_node] = (uint8_t)out.entry_layer;
    if (n && (out.entry_node >= n || (int)out.entry_layer >= max_layers)) { err = "bad entry point"; return false; }
    std::vector<uint64_t> upper_off(n);
    uint64_t rows = 0;
    for (uint64_t i = 0; i < n; ++i) { upper_off[i] = rows; rows += out.level[i]; }
    out.upper_rows = rows;
    out.adj0.assign((size_t)n * s0, 0xFFFFFFFFu);
    out.adjU.assign((size_t)(rows ? rows : 1) * su, 0xFFFFFFFFu);
    bool have_w = !edges.empty();
    if (have_w) { out.w0.assign((size_t)n * s0, 0.0f); out.wU.assign((size_t)(rows ? rows : 1) * su, 0.0f); }
    size_t epos = 0;
    for (const Ref& r : refs) {  // refs are in file order == hnsw.edges order
        uint32_t* row = r.layer == 0 ? &out.adj0[(size_t)r.node * s0] : &out.adjU[(size_t)(upper_off[r.node] + r.layer - 1) * su];
        float* w = !have_w ? nullptr : (r.layer == 0 ? &out.w0[(size_t)r.node * s0] : &out.wU[(size_t)(upper_off[r.node] + r.layer - 1) * su]);
        for (uint32_t e = 0; e < r.count; ++e) {
            row[e] = rd32(g.data() + r.start + (size_t)e * 4);
            if (have_w) {
                if (epos + 4 > edges.size()) { err = "hnsw.edges too short"; return false; }
                memcpy(&w[e], edges.data() + epos, 4);
                epos += 4;
            }
        }
    }
    return true;
}

// hnsw/disk/v2.rs:122-157, 176-211.  num_layers = entry_layer + 1.
inline bool write_graph_v2(const std::string& graph_path, const std::string& edges_path, const FlatGraph& fg, std::string& err) {
    std::vector<unsigned char> g, e;
    auto put32 = [](std::vector<unsigned char>& v, uint32_t x) { unsigned char b[4]; memcpy(b, &x, 4); v.insert(v.end(), b, b + 4); };
    auto putf = [](std::vector<unsigned char>& v, float x) { unsigned char b[4]; memcpy(b, &x, 4); v.insert(v.end(), b, b + 4); };
    if (fg.n == 0) return write_file(graph_path, nullptr, 0, err) && write_file(edges_path, nullptr, 0, err);
    uint32_t num_layers = fg.entry_layer + 1;
    std::vector<uint64_t> upper_off(fg.n);
    uint64_t rows = 0;
    for (uint64_t i = 0; i < fg.n; ++i) { upper_off[i] = rows; rows += fg.level[i]; }
    std::vector<uint32_t> nodes_end;
    size_t pos_total = 0;
    for (uint64_t node = 0; node < fg.n; ++node) {
        std::vector<size_t> indexing(num_layers);
        size_t pos = 0;
        for (uint32_t layer = 0; layer < num_layers; ++layer) {
            uint32_t cnt = 0;
            const uint32_t* row = nullptr;
            const float* w = nullptr;
            int stride = layer == 0 ? fg.s0 : fg.su;
            if (layer <= fg.level[node]) {
                row = layer == 0 ? &fg.adj0[(size_t)node * fg.s0] : &fg.adjU[(size_t)(upper_off[node] + layer - 1) * fg.su];
                if (!fg.w0.empty()) w = layer == 0 ? &fg.w0[(size_t)node * fg.s0] : &fg.wU[(size_t)(upper_off[node] + layer - 1) * fg.su];
                while ((int)cnt < stride && row[cnt] != 0xFFFFFFFFu) ++cnt;
            }
            indexing[layer] = pos;
            put32(g, cnt);
            for (uint32_t k = 0; k < cnt; ++k) { put32(g, row[k]); putf(e, w ? w[k] : 0.0f); }
            pos += (size_t)(1 + cnt) * 4;
        }
        pos += (size_t)num_layers * 4;
        for (int layer = (int)num_layers - 1; layer >= 0; --layer) put32(g, (uint32_t)(pos - indexing[layer]));
        pos_total += pos;
        if (pos_total > 0xFFFFFFFFull) { err = "graph exceeds the 4 GiB u32 offset limit of hnsw.graph"; return false; }
        nodes_end.push_back((uint32_t)pos_total);
    }
    for (size_t i = nodes_end.size(); i-- > 0;) put32(g, nodes_end[i]);
    put32(g, fg.entry_layer);
    put32(g, fg.entry_node);
    return write_file(graph_path, g.data(), g.size(), err) && write_file(edges_path, e.data(), e.size(), err);
}

}  // namespace segio
