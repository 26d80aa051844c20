"""Tile-stream storage of the hot SpMV phases (engine.cuh TileStreamDev), checked on the CPU.

osqp_b200_stream_selftest builds the stream with the same host code osqp_setup uses and replays the device
reduction (quads, per-lane row-end flags, flag ranks, segmented scan, carries) lane by lane on the host -- no GPU, no
oracle involved; the expected product comes from scipy.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp


def _run(lib, M, x, grid, ngroups, paired=0):
    M = M.tocsr()
    M.sort_indices()
    rp = M.indptr.astype(np.int64)
    ci = M.indices.astype(np.int64)
    va = M.data.astype(np.float64)
    y = np.zeros(M.shape[0])
    pad = C.c_double()
    ip, fp = C.POINTER(C.c_longlong), C.POINTER(C.c_double)
    lib.osqp_b200_stream_selftest.restype = C.c_longlong
    rc = lib.osqp_b200_stream_selftest(
        C.c_longlong(M.shape[0]), C.c_longlong(M.shape[1]), rp.ctypes.data_as(ip), ci.ctypes.data_as(ip),
        va.ctypes.data_as(fp), x.ctypes.data_as(fp), C.c_longlong(grid), C.c_longlong(ngroups), C.c_longlong(paired),
        y.ctypes.data_as(fp), C.byref(pad))
    return rc, y, pad.value


@pytest.mark.parametrize("rows,cols,density,grid,ngroups", [
    (3000, 2000, 0.02, 8, 1),
    (3000, 2000, 0.02, 8, 2),
    (5000, 7000, 0.006, 16, 4),      # short row segments, many zero quads
    (1200, 30000, 0.004, 148, 2),
    (40000, 3000, 0.008, 148, 1),
    (257, 129, 0.3, 9, 3),           # groups do not divide the grid
])
def test_stream_matches_scipy(pkg, engine_lib, rows, cols, density, grid, ngroups):
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(rows + cols)
    M = sp.random(rows, cols, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    x = rng.standard_normal(cols)
    rc, y, pad = _run(lib, M, x, grid, ngroups)
    if rc == 2:
        pytest.skip("builder declined (padding) -- CSR path would be used")
    assert rc == 0
    ref = M @ x
    scale = np.abs(M) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y - ref) / scale) < 1e-14
    assert pad < 1.4  # stored entries (quad padding, zero quads) per non-zero


@pytest.mark.parametrize("rows,cols,density,grid", [(3000, 2000, 0.02, 8), (9000, 7000, 0.004, 148)])
def test_paired_stream_matches_scipy(pkg, engine_lib, rows, cols, density, grid):
    # cluster pairs: blocks 2p / 2p+1 stream the same row range of column groups 0 / 1
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(rows)
    M = sp.random(rows, cols, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    x = rng.standard_normal(cols)
    rc, y, pad = _run(lib, M, x, grid, 2, paired=1)
    assert rc == 0
    scale = np.abs(M) @ np.abs(x) + 1e-300
    assert np.max(np.abs(y - M @ x) / scale) < 1e-14


def test_stream_edge_cases(pkg, engine_lib):
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(5)
    # empty rows, one dense row, one-entry rows, rows shorter than a quad inside one lane
    rows, cols = 2000, 4096
    M = sp.random(rows, cols, density=0.01, random_state=rng, data_rvs=rng.standard_normal, format="lil")
    M[5, :] = 0
    M[6, :] = 0
    M[7, :] = rng.standard_normal(cols)
    for r in range(100, 400):
        M[r, :] = 0
        M[r, (r * 7) % cols] = 1.0 + r
    M = M.tocsr()
    x = rng.standard_normal(cols)
    for ngroups in (1, 2):
        rc, y, pad = _run(lib, M, x, 12, ngroups)
        assert rc in (0, 2)
        if rc == 0:
            np.testing.assert_allclose(y, M @ x, rtol=0, atol=1e-11)


def test_stream_declines_hopeless_padding(pkg, engine_lib):
    lib = pkg.load_library(engine_lib)
    M = sp.eye(20000, format="csr")
    rc, _, _ = _run(lib, M, np.ones(20000), 148, 2)
    assert rc == 2  # 8 stored entries per one-entry row: the builder must refuse


def test_dense_rows_are_split_into_pieces(pkg, engine_lib):
    # portfolio-like: a few rows with thousands of entries next to many short rows (BASELINE config 4: F' rows)
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(11)
    rows, cols = 3000, 20000
    M = sp.random(rows, cols, density=0.0005, random_state=rng, data_rvs=rng.standard_normal, format="lil")
    for r in (0, 17, 1500, 2999):
        M[r, :] = rng.standard_normal(cols) * (rng.random(cols) < 0.5)
    M = M.tocsr()
    x = rng.standard_normal(cols)
    for ngroups in (1, 2):
        rc, y, pad = _run(lib, M, x, 148, ngroups)
        assert rc == 0
        scale = np.abs(M) @ np.abs(x) + 1e-300
        assert np.max(np.abs(y - M @ x) / scale) < 1e-13
    rc, _, _ = _run(lib, M, x, 148, 2, paired=1)
    assert rc == 2  # split rows and cluster pairs do not combine: the builder falls back to the unpaired layout


def test_builder_is_independent_of_host_thread_count(pkg, engine_lib, monkeypatch):
    # osqp_setup splits its host-side index work over threads (row-range ownership): same stream, same result
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(21)
    M = sp.random(60000, 30000, density=0.002, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    x = rng.standard_normal(30000)
    ys = []
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("OSQP_B200_HOST_THREADS", threads)
        rc, y, _ = _run(lib, M, x, 148, 2)
        assert rc == 0
        ys.append(y)
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])


@pytest.mark.parametrize("lane_rows", ["1", "0"])
def test_both_layouts_at_bench_like_shape(pkg, engine_lib, monkeypatch, lane_rows):
    # the lane-row layout (one stream row per lane, sliced-ELL order; the default wherever its padding stays small) and
    # the scan layout (quads dealt across the lanes, segmented scan) must give the same product
    monkeypatch.setenv("OSQP_B200_LANE_ROWS", lane_rows)
    lib = pkg.load_library(engine_lib)
    rng = np.random.default_rng(5)
    M = sp.random(30000, 20000, density=0.0025, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    x = rng.standard_normal(20000)
    for ngroups, paired in ((1, 0), (2, 0), (2, 1), (4, 0)):
        rc, y, pad = _run(lib, M, x, 148, ngroups, paired)
        assert rc == 0
        scale = np.abs(M) @ np.abs(x) + 1e-300
        assert np.max(np.abs(y - M @ x) / scale) < 1e-14
        assert pad < 1.35
