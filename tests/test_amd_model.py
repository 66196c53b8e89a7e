"""CPU: the product's cooperative AMD source (parth_b200/csrc/amd_core.cuh) run under the host lane
model (tests/amd_model: one OS thread per lane, barrier-based collectives) against the oracle
restatement (oracle/amd_restated.c).  Warp widths 4, 8 and 32 put the chunk boundaries, the flattened
element build, the run-wise degree-list removal and the workspace compaction under very different
schedules; the result must be bit-identical every time."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from parth_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_ip = C.POINTER(C.c_int)


def _lib(width):
    out_dir = os.path.join(HERE, "amd_model", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libamd_model_w{width}.so")
    srcs = [os.path.join(HERE, "amd_model", "amd_model.cpp"), os.path.join(HERE, "amd_model", "lanes_host.h"),
            os.path.join(ROOT, "parth_b200", "csrc", "amd_core.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", f"-DPB_MODEL_WARP={width}",
                        "-I", os.path.join(ROOT, "parth_b200", "csrc"), srcs[0], "-o", out], check=True)
    lib = C.CDLL(out)
    lib.amd_model_order.argtypes = [C.c_int, _ip, _ip, _ip, C.c_int, C.c_int]
    assert lib.amd_model_warp() == width
    return lib


def _order(lib, Mp, Mi, threads, extra):
    Mp, Mi = np.ascontiguousarray(Mp, np.int32), np.ascontiguousarray(Mi, np.int32)
    n = len(Mp) - 1
    P = np.full(n, -7, np.int32)
    sym = lib.amd_model_order(n, Mp.ctypes.data_as(_ip), Mi.ctypes.data_as(_ip), P.ctypes.data_as(_ip), threads, extra)
    return P, sym


def _csr(n, r, c):
    _, Mp, Mi = synth.coo_to_csc(n, r, c)
    return Mp, Mi


def _rsym(rng, n, m):
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    k = r != c
    return _csr(n, np.concatenate([r[k], c[k]]), np.concatenate([c[k], r[k]]))


def _rasym(rng, n, m):
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    k = r != c
    return _csr(n, r[k], c[k])


def _grid2d(k):
    idx = np.arange(k * k).reshape(k, k)
    r, c = [], []
    for a, b in ((idx[:-1, :], idx[1:, :]), (idx[:, :-1], idx[:, 1:])):
        r += [a.ravel(), b.ravel()]
        c += [b.ravel(), a.ravel()]
    return _csr(k * k, np.concatenate(r), np.concatenate(c))


def _kuhn(k):
    idx = np.arange(k ** 3).reshape(k, k, k)
    r, c = [], []
    for d in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]:
        a = idx[:k - d[0], :k - d[1], :k - d[2]]
        b = idx[d[0]:, d[1]:, d[2]:]
        r += [a.ravel(), b.ravel()]
        c += [b.ravel(), a.ravel()]
    return _csr(k ** 3, np.concatenate(r), np.concatenate(c))


def _graphs(seed):
    rng = np.random.default_rng(seed)
    g = [synth.grid_graph(k) for k in (2, 3, 5, 8, 11)] + [_grid2d(24), _kuhn(7)]
    g += [_rsym(rng, int(rng.integers(1, 300)), int(rng.integers(0, 900))) for _ in range(14)]
    g += [_rasym(rng, int(rng.integers(2, 200)), int(rng.integers(1, 900))) for _ in range(6)]   # general A + A' build
    g.append(_rsym(rng, 400, 30000))                                                            # dense rows
    hub = np.arange(1, 200)
    g.append(_csr(200, np.concatenate([np.zeros(199, np.int64), hub]), np.concatenate([hub, np.zeros(199, np.int64)])))
    g.append((np.zeros(6, np.int32), np.zeros(0, np.int32)))                                    # no edges: identity
    return g


@pytest.mark.parametrize("width", [4, 8, 32])
def test_cooperative_amd_is_bit_exact_under_the_lane_model(port, width):
    lib = _lib(width)
    for gi, (Mp, Mi) in enumerate(_graphs(40 + width)):
        want = port.amd(Mp, Mi)
        # extra = 0: the smallest legal work-space (compaction runs often); large: it never runs
        for extra in (0, 1 << 20):
            got, _ = _order(lib, Mp, Mi, 2 * width, extra)
            assert np.array_equal(got, want), (width, gi, len(Mp) - 1, extra)


def test_symmetric_rows_take_the_parallel_list_build(port):
    lib = _lib(8)
    _, sym = _order(lib, *synth.grid_graph(4), 16, 0)
    assert sym == 1
    rng = np.random.default_rng(3)
    _, sym = _order(lib, *_rasym(rng, 50, 200), 16, 0)
    assert sym == 0
