"""GPU: examples/example_creator (the reference's tests/example_creator.cpp:61-164 against the drop-in class) and the
tree fixture on disk: the tool's cold tree, saved with ParthAPI::saveTree, is loaded here into the
CUDA path and into the oracle; both must report the update the tool printed for each of its five scenarios."""
import os
import re
import subprocess

import numpy as np
import pytest

from parth_b200 import synth

pytestmark = pytest.mark.gpu


def _read_mtx(path):
    with open(path) as f:
        assert f.readline().startswith("%%MatrixMarket matrix coordinate pattern general")
        n, _, nnz = (int(x) for x in f.readline().split())
        rc = np.loadtxt(f, dtype=np.int64).reshape(-1, 2) - 1
    assert len(rc) == nnz
    return synth.coo_to_csc(n, rc[:, 0], rc[:, 1])


def test_example_creator_and_tree_fixture(gpu_api, port, tmp_path):
    from parth_b200 import api, build
    from oracle import cpu_parth
    exe = build.build_example("example_creator")
    out = str(tmp_path)
    tree_file = os.path.join(out, "tree.bin")
    n, L = 16, 5
    r = subprocess.run([exe, "-o", out, "-g", str(n), "-l", str(L), "-e", "3", "-s", "5", "--save-tree", tree_file],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    said = {int(m.group(1)): (int(m.group(2)), float(m.group(3)))
            for m in re.finditer(r"test (\d+): changes (\d+) reuse ([0-9.]+)", r.stdout)}
    assert sorted(said) == [0, 1, 2, 3, 4]
    t = api.load_tree_file(tree_file)
    assert (t["M_n"], t["levels"], t["nodes"]) == (n ** 3, L, 2 ** (L + 1) - 1)
    assert sorted(t["dofs"].tolist()) == list(range(n ** 3))
    N0, Ap0, Ai0 = _read_mtx(os.path.join(out, "original_matrix.mtx"))
    gMp, gMi = synth.grid_graph(n)
    assert np.array_equal(t["Mp"], gMp) and np.array_equal(t["Mi"], gMi)   # the snapshot graph travels with the tree
    tree = cpu_parth.Tree(t["M_n"], L, t["nodes"], t["meta"], t["node_ptr"], t["dofs"], t["labels"])
    # scenario 1 (ancestor / descendant pairs) must be a full-reuse update; scenario 0 touches the root's children
    assert said[1][1] == 1.0 and said[1][0] == 0
    for k in range(5):
        N, Ap, Ai = _read_mtx(os.path.join(out, f"{k}_matrix.mtx"))
        g = gpu_api.ParthAPI(); g.verbose = False; g.ND_levels = L; g.reorder_type = gpu_api.AMD
        o = port.CpuParth("port"); o.set_options(reorder_type=1, nd_levels=L)
        g.import_tree(tree, t["Mp"], t["Mi"]); o.import_tree(tree, t["Mp"], t["Mi"])
        g.setMatrix(N, Ap, Ai, 1); o.setMatrix(N, Ap, Ai, 1)
        st, perm = g.computePermutation(1)
        o.computePermutation(1)
        assert g.getNumChanges() == o.getNumChanges() == said[k][0], k
        assert np.array_equal(g.changed_edges(), o.changed_edges())
        if o.getReuse() > 0:
            assert g.getReuse() == o.getReuse()
        assert abs(g.getReuse() - said[k][1]) < 1e-8, k
        assert sorted(perm.tolist()) == list(range(N))
    # --load-tree: a second run starts from the fixture instead of dissecting again and reports the same updates
    r2 = subprocess.run([exe, "-o", out, "-g", str(n), "-l", str(L), "-e", "3", "-s", "5", "--load-tree", tree_file],
                        capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    said2 = {int(m.group(1)): (int(m.group(2)), float(m.group(3)))
             for m in re.finditer(r"test (\d+): changes (\d+) reuse ([0-9.]+)", r2.stdout)}
    assert said2 == said


def test_e2e_update_example_runs_pageable_and_pinned(gpu_api):
    """examples/e2e_update: warm updates through PARTH::ParthAPI with std::vector inputs, pageable and page-locked
    (parth_b200_pin_host); the permutation it ends with is valid and the update really finds its change"""
    import json
    from parth_b200 import build
    exe = build.build_example("e2e_update")
    r = subprocess.run([exe, "40", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["valid_permutation"] and out["N"] == 64000 and out["ms_per_update_pageable"] > 0
    assert out["ms_per_update_pinned"] > 0  # pinning the caller's vectors worked
