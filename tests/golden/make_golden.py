"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libparth_ref.so,
compiled from /root/reference by oracle/Makefile) in the build container.

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

The reference is non-deterministic with more than one OpenMP thread, so the
script insists on OMP_NUM_THREADS=1.  METIS results depend on the METIS build, so trees are
FIXTURES (inputs to both sides), never golden values; everything stored as "expected" is a
deterministic function of (tree fixture, graph), produced by the reference's own code.

Outputs
  matrices.npz   structural patterns of reference tests/matrices/*.mtx (original + added entries)
  build.npz      ParthAPI::computeMeshFromMatrix results (hashes, sizes) for those and for grids
  tree_fish.npz  METIS tree fixture: reference cold computePermutation on original_matrix
  tree_g12.npz   METIS tree fixture of a 12^3 grid (4 levels)
  scenarios.npz  expected results of warm updates on injected trees (changed edges, dirty sets,
                 reuse ratio, permutations, exported trees) for every scenario in scenario_list()
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_parth as cp  # noqa: E402
from parth_b200 import synth  # noqa: E402
from tests import scenario_defs as sd  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MTX = "/root/reference/tests/matrices"


def main():
    assert os.environ.get("OMP_NUM_THREADS") == "1", "run with OMP_NUM_THREADS=1"
    cp.build("reference")
    # ---- matrices
    N, Ap, Ai = synth.read_mtx_pattern(f"{REF_MTX}/original_matrix.mtx")
    base = set((Ap.searchsorted(np.arange(len(Ai)), side="right") - 1).astype(np.int64) * N + Ai)
    out = {"N": np.int32(N), "Ap": Ap, "Ai": Ai}
    mats = {"original": (N, Ap, Ai)}
    for k in "01234":
        n2, ap2, ai2 = synth.read_mtx_pattern(f"{REF_MTX}/{k}_matrix.mtx")
        col = np.repeat(np.arange(n2, dtype=np.int64), np.diff(ap2))
        keys = col * N + ai2
        new = np.array(sorted(set(keys.tolist()) - base), dtype=np.int64)
        assert len(set(base) - set(keys.tolist())) == 0
        out["added_" + k] = np.stack([new % N, new // N], axis=1).astype(np.int32)
        mats[k] = (n2, ap2, ai2)
    np.savez_compressed(os.path.join(HERE, "matrices.npz"), **out)

    # ---- build goldens
    b = {}
    ref = cp.CpuParth("reference")

    def add_build(name, N, Ap, Ai, dim):
        ref.setMatrix(N, Ap, Ai, dim)
        Mp, Mi = ref.mesh()
        b[name + "_hash_Mp"] = np.array(synth.hexhash(Mp))
        b[name + "_hash_Mi"] = np.array(synth.hexhash(Mi))
        b[name + "_M_n"] = np.int32(len(Mp) - 1)
        b[name + "_nnz"] = np.int32(len(Mi))

    for k, (n2, ap2, ai2) in mats.items():
        add_build(k, n2, ap2, ai2, 1)
    for name, (N2, Ap2, Ai2, dim) in sd.build_cases().items():
        add_build(name, N2, Ap2, Ai2, dim)
    np.savez_compressed(os.path.join(HERE, "build.npz"), **b)

    # ---- METIS tree fixtures (reference cold start)
    def cold_tree(N, Ap, Ai, levels, path):
        r = cp.CpuParth("reference")
        r.set_options(reorder_type=0, nd_levels=levels)
        r.setMatrix(N, Ap, Ai, 1)
        st, perm = r.computePermutation(1)
        assert st == 0 and sorted(perm.tolist()) == list(range(N))
        t = r.export_tree()
        t.save(path)
        return t

    cold_tree(*mats["original"], 8, os.path.join(HERE, "tree_fish.npz"))
    Ng, Apg, Aig = synth.poisson3d(12)
    cold_tree(Ng, Apg, Aig, 4, os.path.join(HERE, "tree_g12.npz"))

    # ---- scenarios
    s = {}
    for name, sc in sd.scenario_list().items():
        res = sd.run_scenario(cp, "reference", sc)
        for k, v in res.items():
            s[name + "/" + k] = v
        print(name, "status", int(res["status"]), "changes", int(res["num_changes"]), "reuse", float(res["reuse"]),
              "fine", res["dirty_fine"].tolist(), "coarse", res["dirty_coarse"].tolist())
    np.savez_compressed(os.path.join(HERE, "scenarios.npz"), **s)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
