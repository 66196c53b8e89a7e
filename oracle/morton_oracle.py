"""oracle/morton_oracle.py -- TEST INFRASTRUCTURE (never imported by the product path).

CPU restatement (numpy) of ReorderingType MORTON_CODE.  PARITY UNPINNED: the reference declares the type
(/root/reference/src/Parth/include/ParthTypes.h:11) but HMD::Permute only prints "Morton code is not implemented
yet" and returns with perm unset (src/Parth/HMD.cpp:82-83); ParthAPI has no coordinate input.  There is no
reference output, golden vector or test for this path, so the oracle states the textbook definition the device
kernels (parth_b200/csrc/morton.cu) follow, and the tests pin it with definition-level properties (Z-order of a
2x2x2 / 4x4x4 lattice, a bit-by-bit interleave, sortedness, permutation validity):

  * per axis: lo = min, ext = max - lo over the finite-or-infinite, non-NaN coordinates;
    q = 0 if ext <= 0 or x is NaN, else min(2^bits - 1, floor((x - lo) / ext * 2^bits)) in IEEE double arithmetic,
    one rounding per operation, evaluated left to right (numpy does exactly that; the device uses __dsub_rn /
    __ddiv_rn / __dmul_rn so no FMA contraction can change a bit);
  * key = sum over bit i of  x_i << (3i+2) | y_i << (3i+1) | z_i << 3i   (x most significant);
  * order = ascending (key, point id): perm[new] = old;
  * a separator-tree node is ordered by the global rank of its DOFs (HMD::Permute's contract, HMD.cpp:77-87:
    perm[new] = old LOCAL id, local ids = rank of the DOF inside the node's ascending DOF list).
"""
import numpy as np


def quantise(xyz, bits=21):
    xyz = np.asarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    q = np.zeros((n, 3), dtype=np.uint64)
    scale = np.float64(2.0 ** bits)
    qmax = np.uint64((1 << bits) - 1)
    for a in range(3):
        x = xyz[:, a]
        ok = ~np.isnan(x)
        if not ok.any():
            continue
        lo = x[ok].min()
        with np.errstate(invalid="ignore", over="ignore"):
            ext = x[ok].max() - lo
            if not (ext > 0.0):
                continue
            t = (x - lo) / ext * scale
            big = t >= scale
            tt = np.where(ok & ~big, t, 0.0)
            tt = np.where(np.isnan(tt), 0.0, tt)  # inf - inf inside an infinite box
            qa = tt.astype(np.uint64)
        qa[big & ok] = qmax
        q[:, a] = qa
    return q


def interleave(q):
    """bit-by-bit definition (deliberately not the magic-number spread the device uses)"""
    key = np.zeros(q.shape[0], dtype=np.uint64)
    for i in range(21):
        for a in range(3):
            key |= ((q[:, a] >> np.uint64(i)) & np.uint64(1)) << np.uint64(3 * i + 2 - a)
    return key


def keys(xyz, bits=21):
    return interleave(quantise(xyz, bits))


def order(xyz, bits=21):
    """perm[new] = old, ascending (key, point id); also returns the keys"""
    k = keys(xyz, bits)
    return np.argsort(k, kind="stable").astype(np.int32), k


def node_order(rank, dofs):
    """local permutation of a node whose ascending DOF list is dofs: perm[new] = old local id"""
    return np.argsort(rank[np.asarray(dofs)], kind="stable").astype(np.int32)
