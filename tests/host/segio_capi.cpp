// Test shim (CPU suite): a C interface over the product's host-side segment file code, nucliadb_b200/csrc/segment_io.hpp,
// compiled with plain g++ so that the reader / writer of hnsw.graph, hnsw.edges and vectors.bin are exercised without a GPU.
#include "../../nucliadb_b200/csrc/segment_io.hpp"

extern "C" {

// flat graph -> hnsw.graph + hnsw.edges (DiskHnswV2::serialize_into)
int segio_write_graph(const char* graph_path, const char* edges_path, uint64_t n, int s0, int su, uint32_t entry_node, uint32_t entry_layer,
                      const uint8_t* level, const uint32_t* adj0, const float* w0, const uint32_t* adjU, const float* wU, uint64_t upper_rows) {
    segio::FlatGraph fg;
    fg.n = n; fg.s0 = s0; fg.su = su; fg.entry_node = entry_node; fg.entry_layer = entry_layer; fg.upper_rows = upper_rows;
    fg.level.assign(level, level + n);
    fg.adj0.assign(adj0, adj0 + n * s0);
    fg.adjU.assign(adjU, adjU + (upper_rows ? upper_rows : 1) * su);
    if (w0) { fg.w0.assign(w0, w0 + n * s0); fg.wU.assign(wU, wU + (upper_rows ? upper_rows : 1) * su); }
    std::string err;
    return segio::write_graph_v2(graph_path, edges_path, fg, err) ? 0 : -1;
}

// hnsw.graph (+ hnsw.edges) -> flat graph.  Call once with null outputs for the sizes, then with buffers.
int segio_parse_graph(const char* graph_path, const char* edges_path, uint64_t n, int s0, int su, int max_layers, uint32_t* entry, uint64_t* upper_rows,
                      uint8_t* level, uint32_t* adj0, float* w0, uint32_t* adjU, float* wU, char* err_out, int err_cap) {
    std::vector<unsigned char> g, e;
    std::string err;
    segio::FlatGraph fg;
    bool ok = segio::read_file(graph_path, g, err) && (!edges_path || segio::read_file(edges_path, e, err)) &&
              segio::parse_graph_v2(g, e, n, s0, su, max_layers, fg, err);
    if (!ok) { if (err_out) snprintf(err_out, err_cap, "%s", err.c_str()); return -1; }
    entry[0] = fg.entry_node; entry[1] = fg.entry_layer;
    *upper_rows = fg.upper_rows;
    if (level) memcpy(level, fg.level.data(), n);
    if (adj0) memcpy(adj0, fg.adj0.data(), fg.adj0.size() * 4);
    if (adjU) memcpy(adjU, fg.adjU.data(), fg.adjU.size() * 4);
    if (w0 && !fg.w0.empty()) memcpy(w0, fg.w0.data(), fg.w0.size() * 4);
    if (wU && !fg.wU.empty()) memcpy(wU, fg.wU.data(), fg.wU.size() * 4);
    return 0;
}

int segio_write_vectors(const char* path, const float* vecs, uint64_t n, int d, int ld, const uint32_t* paragraph_of) {
    std::string err;
    return segio::write_vectors_bin(path, vecs, n, d, ld, paragraph_of, err) ? 0 : -1;
}
}
