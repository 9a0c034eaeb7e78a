"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pure-Python restatement (small cases) of the reference's
segment files, used to check the product's C++ reader/writer (nucliadb_b200/csrc/segment_io.hpp):

* hnsw.graph / hnsw.edges   nidx/nidx_vector/src/hnsw/disk/v2.rs:16-49 (format), 122-157 serialize_node,
                            176-211 serialize_into, 159-174 get_out_edges, 236-244 get_node, 220-234 entrypoint
* vectors.bin               nidx/nidx_vector/src/data_store/v2/vector_store.rs:33-68,113-146

A graph here is ``layers[l][node] = [(target, weight), ...]`` (RAMHnsw, ram_hnsw.rs:31-69).
"""
from __future__ import annotations

import struct


def serialize_graph(layers, num_nodes, entry_node, entry_layer):
    """DiskHnswV2::serialize_into -> (graph bytes, edges bytes)."""
    if num_nodes == 0:
        return b"", b""
    g, e = bytearray(), bytearray()
    nodes_end, pos_total = [], 0
    num_layers = len(layers)
    for node in range(num_nodes):
        indexing, pos = {}, 0
        for layer in range(num_layers):
            edges = layers[layer].get(node, [])
            indexing[layer] = pos
            g += struct.pack("<I", len(edges))
            for to, w in edges:
                g += struct.pack("<I", to)
                e += struct.pack("<f", w)
            pos += (1 + len(edges)) * 4
        pos += num_layers * 4
        for layer in reversed(range(num_layers)):
            g += struct.pack("<I", pos - indexing[layer])
        pos_total += pos
        nodes_end.append(pos_total)
    for end in reversed(nodes_end):
        g += struct.pack("<I", end)
    g += struct.pack("<II", entry_layer, entry_node)
    return bytes(g), bytes(e)


def _u32(b, pos):
    return struct.unpack_from("<I", b, pos)[0]


def entrypoint(graph: bytes):
    return _u32(graph, len(graph) - 4), _u32(graph, len(graph) - 8)  # (node, layer)


def get_out_edges(graph: bytes, node: int, layer: int):
    """DiskHnswV2::get_node + get_out_edges."""
    indexing_end = len(graph) - 8
    node_end = _u32(graph, indexing_end - (node + 1) * 4)
    pos = node_end - (layer + 1) * 4
    start = node_end - _u32(graph, pos)
    n = _u32(graph, start)
    return [_u32(graph, start + 4 + 4 * i) for i in range(n)]


def write_vectors_bin(vectors, paragraph_of) -> bytes:
    out = bytearray()
    for v, p in zip(vectors, paragraph_of):
        out += struct.pack(f"<{len(v)}f", *[float(x) for x in v])
        out += struct.pack("<I", int(p))
    return bytes(out)
