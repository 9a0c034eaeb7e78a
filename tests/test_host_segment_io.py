"""CPU coverage of the product's host-side code that needs no device:
  * nucliadb_b200/csrc/segment_io.hpp (the reader / writer of the reference's hnsw.graph, hnsw.edges and vectors.bin), compiled
    with g++ through tests/host/segio_capi.cpp and compared byte for byte with oracle/disk_v2.py, the pure-Python restatement
    that test_oracle_golden.py pins to the reference's worked example (hnsw/disk/v2.rs:16-49) and hnsw_test (v2.rs:349-398);
  * nidx_hnsw_levels (build.rs:40,97-101), compared with the oracle's level draw."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from oracle import disk_v2
from nucliadb_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
NIL = 0xFFFFFFFF


@pytest.fixture(scope="module")
def segio(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("segio") / "segio_capi.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "segio_capi.cpp")], check=True)
    return C.CDLL(so)


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def random_graph(n, m, m0, seed):
    """A flat graph with the layout invariants of the library (left-packed rows, links inside the layer, the entry point on top)."""
    rng = np.random.default_rng(seed)
    g = O.Graph(n, m, m0, O.assign_levels(n, m, 2))
    for layer in range(int(g.level.max()) + 1):
        members = np.nonzero(g.level >= layer)[0]
        cap = m0 if layer == 0 else m
        for node in members:
            k = int(rng.integers(0, min(cap, len(members) - 1) + 1)) if len(members) > 1 else 0
            if layer > 0 and node == g.entry_node and layer == g.entry_layer:
                k = 0                                   # alone or not, the entry point may have no edges on top
            targets = rng.choice(members[members != node], k, replace=False) if k else np.zeros(0, np.int64)
            row, wrow = (g.adj0[node], g.w0[node]) if layer == 0 else (g.adjU[int(g.upper_off[node]) + layer - 1], g.wU[int(g.upper_off[node]) + layer - 1])
            row[:k], wrow[:k] = targets, rng.random(k, dtype=np.float32)
    return g


def as_layers(g):
    layers = []
    for layer in range(g.entry_layer + 1):
        cnx = {}
        for node in np.nonzero(g.level >= layer)[0]:
            row, wrow = (g.adj0[node], g.w0[node]) if layer == 0 else (g.adjU[int(g.upper_off[node]) + layer - 1], g.wU[int(g.upper_off[node]) + layer - 1])
            k = int((row != NIL).sum())
            cnx[int(node)] = [(int(t), float(w)) for t, w in zip(row[:k], wrow[:k])]
        layers.append(cnx)
    return layers


def write(segio, g, d):
    gp, ep = os.path.join(d, "hnsw.graph"), os.path.join(d, "hnsw.edges")
    rows = int(g.level.astype(np.int64).sum())
    rc = segio.segio_write_graph(gp.encode(), ep.encode(), C.c_uint64(g.n), C.c_int(g.adj0.shape[1]), C.c_int(g.adjU.shape[1]), C.c_uint32(g.entry_node),
                                 C.c_uint32(g.entry_layer), p(g.level), p(g.adj0), p(g.w0), p(g.adjU), p(g.wU), C.c_uint64(rows))
    assert rc == 0
    return gp, ep


def parse(segio, gp, ep, n, s0, su):
    entry = np.zeros(2, np.uint32)
    rows = C.c_uint64()
    err = C.create_string_buffer(256)
    rc = segio.segio_parse_graph(gp.encode(), ep.encode() if ep else None, C.c_uint64(n), C.c_int(s0), C.c_int(su), C.c_int(8), p(entry), C.byref(rows),
                                 None, None, None, None, None, err, 256)
    if rc:
        return None, err.value.decode()
    level = np.zeros(n, np.uint8)
    adj0, w0 = np.zeros((n, s0), np.uint32), np.zeros((n, s0), np.float32)
    adjU, wU = np.zeros((max(rows.value, 1), su), np.uint32), np.zeros((max(rows.value, 1), su), np.float32)
    assert segio.segio_parse_graph(gp.encode(), ep.encode() if ep else None, C.c_uint64(n), C.c_int(s0), C.c_int(su), C.c_int(8), p(entry), C.byref(rows),
                                   p(level), p(adj0), p(w0), p(adjU), p(wU), err, 256) == 0
    return dict(entry=entry, rows=rows.value, level=level, adj0=adj0, w0=w0, adjU=adjU, wU=wU), ""


@pytest.mark.parametrize("n,m,m0,seed", [(1, 4, 8, 0), (2, 4, 8, 1), (40, 4, 8, 2), (300, 3, 6, 3)])
def test_writer_equals_the_python_restatement_and_reader_round_trips(segio, tmp_path, n, m, m0, seed):
    g = random_graph(n, m, m0, seed)
    gp, ep = write(segio, g, str(tmp_path))
    want_graph, want_edges = disk_v2.serialize_graph(as_layers(g), n, g.entry_node, g.entry_layer)
    assert open(gp, "rb").read() == want_graph and open(ep, "rb").read() == want_edges            # DiskHnswV2::serialize_into, byte for byte
    got, err = parse(segio, gp, ep, n, g.adj0.shape[1], g.adjU.shape[1])
    assert got is not None, err
    # levels come back from where nodes have or receive edges (v2.rs:248-312): the same, except for isolated upper-layer nodes
    linked = np.zeros(n, np.uint8)
    for layer, cnx in enumerate(as_layers(g)):
        for node, edges in cnx.items():
            if edges and layer > linked[node]:
                linked[node] = layer
            for t, _ in edges:
                linked[t] = max(linked[t], layer)
    linked[g.entry_node] = max(linked[g.entry_node], g.entry_layer)
    assert (got["level"] == linked).all() and got["entry"].tolist() == [g.entry_node, g.entry_layer]
    assert (got["adj0"] == g.adj0).all() and np.array_equal(got["w0"], g.w0)
    for node in np.nonzero(linked > 0)[0]:
        for layer in range(1, int(linked[node]) + 1):
            src = g.adjU[int(g.upper_off[node]) + layer - 1]
            off = int(linked[:node].astype(np.int64).sum()) + layer - 1
            assert (got["adjU"][off] == src).all()
    for node in range(n):                                                                       # and the Python reader agrees
        assert disk_v2.get_out_edges(want_graph, node, 0) == [int(t) for t in g.adj0[node][g.adj0[node] != NIL]]


def test_reader_takes_the_reference_worked_example(segio, tmp_path):
    # hnsw/disk/v2.rs:16-49: node 0 with 5 edges in layer 0, 3 in layer 1, none in layer 2 (entry point (0, 2))
    layers = [{0: [(1, 0.1), (17, 0.2), (5433, 0.3), (45, 0.4), (667, 0.5)]}, {0: [(45, 1.0), (666, 2.0), (22, 3.0)]}, {}]
    graph, edges = disk_v2.serialize_graph(layers, 1, entry_node=0, entry_layer=2)
    gp, ep = tmp_path / "hnsw.graph", tmp_path / "hnsw.edges"
    gp.write_bytes(graph), ep.write_bytes(edges)
    got, err = parse(segio, str(gp), str(ep), 1, 8, 4)
    assert got is None and "out of range" in err            # the example's targets (17, 5433...) do not exist in a 1-node graph: rejected
    # the 3-node hnsw_test (v2.rs:349-398)
    cnx = [{0: [(1, 1.0)], 1: [(2, 2.0)], 2: [(0, 3.0)]}, {0: [(1, 4.0)], 1: [(2, 5.0)]}, {0: [(1, 6.0)]}]
    graph, edges = disk_v2.serialize_graph(cnx, 3, 0, 2)
    gp.write_bytes(graph), ep.write_bytes(edges)
    got, err = parse(segio, str(gp), str(ep), 3, 8, 4)
    assert got is not None, err
    assert got["level"].tolist() == [2, 2, 1] and got["entry"].tolist() == [0, 2]
    assert got["adj0"][:, 0].tolist() == [1, 2, 0] and got["w0"][:, 0].tolist() == [1.0, 2.0, 3.0]
    assert got["adjU"][0, 0] == 1 and got["adjU"][1, 0] == 1 and got["adjU"][2, 0] == 2 and got["wU"][1, 0] == 6.0
    # corrupt files are refused, not read out of bounds
    gp.write_bytes(graph[:-5])
    assert parse(segio, str(gp), str(ep), 3, 8, 4)[0] is None
    gp.write_bytes(graph)
    ep.write_bytes(edges[:-4])
    assert "too short" in parse(segio, str(gp), str(ep), 3, 8, 4)[1]
    assert "more edges" in parse(segio, str(gp), None, 3, 0, 4)[1]


def test_vectors_bin_records(segio, tmp_path):
    rng = np.random.default_rng(5)
    v = rng.standard_normal((7, 12)).astype(np.float32)      # ld 12 > d 10: the device-side row padding is dropped on disk
    par = np.asarray([0, 0, 1, 2, 2, 2, 3], np.uint32)
    path = str(tmp_path / "vectors.bin")
    assert segio.segio_write_vectors(path.encode(), p(v), C.c_uint64(7), C.c_int(10), C.c_int(12), p(par)) == 0
    assert open(path, "rb").read() == disk_v2.write_vectors_bin(v[:, :10], par)      # [dim x f32 LE][paragraph_addr u32 LE]


def test_level_draw_equals_the_oracle():
    L = _lib.load()
    for n, m, seed in ((0, 30, 2), (1, 30, 2), (5000, 30, 2), (5000, 16, 2), (3000, 4, 77)):
        out = np.zeros(max(n, 1), np.uint8)
        assert L.nidx_hnsw_levels(C.c_uint64(n), C.c_int32(m), C.c_uint64(seed), p(out)) == 0
        assert (out[:n] == O.assign_levels(n, m, seed)).all()
    assert L.nidx_hnsw_levels(C.c_uint64(4), C.c_int32(1), C.c_uint64(2), p(np.zeros(4, np.uint8))) != 0
