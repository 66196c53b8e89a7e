"""oracle/cpu_parth.py -- TEST INFRASTRUCTURE (ctypes binding of oracle/cpu_parth.h).

Loads either CPU implementation of the path behind the same C ABI:
  kind="port"       oracle/liboracle_parth.so   (plain-C restatement)
  kind="reference"  oracle/_ref/libparth_ref.so (the unmodified reference, compiled by oracle/Makefile)
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {
    "port": os.path.join(HERE, "liboracle_parth.so"),
    "reference": os.path.join(HERE, "_ref", "libparth_ref.so"),
}
NEEDS_REDISSECT = -10
NO_TREE = -11
UNSUPPORTED = -12

_ip = C.POINTER(C.c_int)
_libs = {}


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_ip)


def build(kind="port", quiet=True):
    """Compile the requested library with oracle/Makefile (no-op if up to date)."""
    target = "port" if kind == "port" else "ref"
    r = subprocess.run(["make", "-C", HERE, target], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


def available(kind):
    return os.path.exists(PATHS[kind])


def load(kind="port"):
    if kind in _libs:
        return _libs[kind]
    path = PATHS[kind]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle {'port' if kind == 'port' else 'ref'}`")
    # RTLD_LOCAL + distinct files: both kinds can live in one process
    lib = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    lib.cpu_parth_kind.restype = C.c_char_p
    lib.cpu_parth_create.restype = C.c_void_p
    lib.cpu_parth_destroy.argtypes = [C.c_void_p]
    lib.cpu_parth_set_options.argtypes = [C.c_void_p] + [C.c_int] * 7
    lib.cpu_parth_set_matrix.argtypes = [C.c_void_p, C.c_int, _ip, _ip, C.c_int, _ip, C.c_int]
    lib.cpu_parth_set_mesh.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _ip, C.c_int]
    lib.cpu_parth_mesh_dims.argtypes = [C.c_void_p, _ip, _ip]
    lib.cpu_parth_mesh_copy.argtypes = [C.c_void_p, _ip, _ip]
    lib.cpu_parth_compute_permutation.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int]
    lib.cpu_parth_get_reuse.argtypes = [C.c_void_p]
    lib.cpu_parth_get_reuse.restype = C.c_double
    lib.cpu_parth_get_num_changes.argtypes = [C.c_void_p]
    lib.cpu_parth_get_changed_edges.argtypes = [C.c_void_p, _ip, C.c_int]
    lib.cpu_parth_get_dirty.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    lib.cpu_parth_get_timers.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.cpu_parth_get_mesh_perm.argtypes = [C.c_void_p, _ip, C.c_int]
    lib.cpu_parth_tree_dims.argtypes = [C.c_void_p, _ip, _ip, _ip]
    lib.cpu_parth_tree_export.argtypes = [C.c_void_p] + [_ip] * 6
    lib.cpu_parth_tree_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [_ip] * 6
    lib.cpu_parth_integrate.argtypes = [C.c_void_p]
    lib.cpu_parth_detect.argtypes = [C.c_void_p]
    lib.cpu_parth_permute_fine.argtypes = [C.c_void_p, _ip, C.c_int]
    lib.cpu_parth_submesh.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]
    lib.cpu_parth_map_mesh_perm.argtypes = [C.c_void_p, _ip, C.c_int, C.c_int, _ip]
    lib.cpu_parth_amd.argtypes = [C.c_int, _ip, _ip, _ip]
    lib.cpu_parth_etree.argtypes = [C.c_int, _ip, _ip, _ip, _ip]
    lib.cpu_parth_colcount.argtypes = [C.c_int, _ip, _ip, _ip, _ip, _ip]
    lib.cpu_parth_colcount.restype = C.c_int64
    if hasattr(lib, "cpu_parth_supernodes"):  # port only
        lib.cpu_parth_supernodes.argtypes = [C.c_int, _ip, _ip, _ip, _ip, _ip, C.POINTER(C.c_int64)]
    _libs[kind] = lib
    return lib


class Tree:
    """Flat tree fixture (layout documented in cpu_parth.h / include/parth_b200.h)."""

    FIELDS = ("meta", "node_ptr", "dofs", "labels")

    def __init__(self, M_n, levels, nodes, meta, node_ptr, dofs, labels, dof2node=None, mesh_perm=None):
        self.M_n, self.levels, self.nodes = int(M_n), int(levels), int(nodes)
        self.meta = _i32(meta).reshape(self.nodes, 7)
        self.node_ptr = _i32(node_ptr)
        self.dofs = _i32(dofs)
        self.labels = _i32(labels)
        self.dof2node = None if dof2node is None else _i32(dof2node)
        self.mesh_perm = None if mesh_perm is None else _i32(mesh_perm)

    def save(self, path):
        np.savez_compressed(path, hdr=np.array([self.M_n, self.levels, self.nodes], np.int32),
                            meta=self.meta, node_ptr=self.node_ptr, dofs=self.dofs, labels=self.labels)

    @staticmethod
    def load(path):
        z = np.load(path)
        M_n, levels, nodes = (int(x) for x in z["hdr"])
        return Tree(M_n, levels, nodes, z["meta"], z["node_ptr"], z["dofs"], z["labels"])

    def node_dofs(self, i):
        return self.dofs[self.node_ptr[i]:self.node_ptr[i + 1]]


class CpuParth:
    """Python face of one cpu_parth handle; method names follow PARTH::ParthAPI."""

    def __init__(self, kind="port"):
        self.kind = kind
        self.lib = load(kind)
        self.h = C.c_void_p(self.lib.cpu_parth_create())
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                self.lib.cpu_parth_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_options(self, reorder_type=0, nd_levels=8, lagging=True, aggressive=False, relocate_lvl=2,
                    lag_lvl=5, verbose=False):
        self.lib.cpu_parth_set_options(self.h, int(reorder_type), int(nd_levels), int(lagging),
                                       int(aggressive), int(relocate_lvl), int(lag_lvl), int(verbose))

    def setMatrix(self, N, Ap, Ai, dim=1, new_to_old=None):
        Ap, Ai = _i32(Ap), _i32(Ai)
        m = None if new_to_old is None or len(new_to_old) == 0 else _i32(new_to_old)
        return self.lib.cpu_parth_set_matrix(self.h, int(N), _p(Ap), _p(Ai), int(dim), _p(m),
                                             0 if m is None else len(m))

    def setMesh(self, n, Mp, Mi, new_to_old=None):
        Mp, Mi = _i32(Mp), _i32(Mi)
        m = None if new_to_old is None or len(new_to_old) == 0 else _i32(new_to_old)
        return self.lib.cpu_parth_set_mesh(self.h, int(n), _p(Mp), _p(Mi), _p(m), 0 if m is None else len(m))

    def mesh(self):
        n, nnz = C.c_int(), C.c_int()
        self.lib.cpu_parth_mesh_dims(self.h, C.byref(n), C.byref(nnz))
        Mp = np.empty(n.value + 1, np.int32)
        Mi = np.empty(max(nnz.value, 0), np.int32)
        self.lib.cpu_parth_mesh_copy(self.h, _p(Mp), _p(Mi))
        return Mp, Mi

    def computePermutation(self, dim=3):
        """returns (status, perm); status < 0 is one of the CPU_PARTH_* codes"""
        n, nnz = C.c_int(), C.c_int()
        self.lib.cpu_parth_mesh_dims(self.h, C.byref(n), C.byref(nnz))
        out = np.empty(max(n.value * dim, 1), np.int32)
        r = self.lib.cpu_parth_compute_permutation(self.h, int(dim), _p(out), out.size)
        if r < 0:
            return r, None
        return 0, out[:r]

    def getReuse(self):
        return float(self.lib.cpu_parth_get_reuse(self.h))

    def getNumChanges(self):
        return int(self.lib.cpu_parth_get_num_changes(self.h))

    def changed_edges(self):
        n = self.getNumChanges()
        out = np.empty((max(n, 1), 2), np.int32)
        self.lib.cpu_parth_get_changed_edges(self.h, _p(out), n)
        return out[:n]

    def dirty(self):
        M_n, lv, nodes = self.tree_dims()
        f = np.empty(max(nodes, 1), np.int32)
        c = np.empty(max(nodes, 1), np.int32)
        nf, nc = C.c_int(), C.c_int()
        self.lib.cpu_parth_get_dirty(self.h, _p(f), C.byref(nf), _p(c), C.byref(nc))
        return f[:nf.value].copy(), c[:nc.value].copy()

    def timers(self):
        t = (C.c_double * 5)()
        self.lib.cpu_parth_get_timers(self.h, t)
        return dict(zip(("init", "integrate", "detect", "compute", "map"), list(t)))

    def mesh_perm(self):
        M_n, _, _ = self.tree_dims()
        out = np.empty(max(M_n, 1), np.int32)
        r = self.lib.cpu_parth_get_mesh_perm(self.h, _p(out), out.size)
        return out[:max(r, 0)]

    def tree_dims(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.cpu_parth_tree_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export_tree(self):
        M_n, lv, nodes = self.tree_dims()
        meta = np.empty((nodes, 7), np.int32)
        node_ptr = np.empty(nodes + 1, np.int32)
        dofs = np.empty(M_n, np.int32)
        labels = np.empty(M_n, np.int32)
        d2n = np.empty(M_n, np.int32)
        mp = np.empty(M_n, np.int32)
        self.lib.cpu_parth_tree_export(self.h, _p(meta), _p(node_ptr), _p(dofs), _p(labels), _p(d2n), _p(mp))
        return Tree(M_n, lv, nodes, meta, node_ptr, dofs, labels, d2n, mp)

    def import_tree(self, tree, Mp_prev, Mi_prev):
        Mp_prev, Mi_prev = _i32(Mp_prev), _i32(Mi_prev)
        return self.lib.cpu_parth_tree_import(self.h, tree.M_n, tree.levels, tree.nodes, _p(tree.meta),
                                              _p(tree.node_ptr), _p(tree.dofs), _p(tree.labels),
                                              _p(Mp_prev), _p(Mi_prev))

    def integrate(self):
        return self.lib.cpu_parth_integrate(self.h)

    def detect(self):
        return self.lib.cpu_parth_detect(self.h)

    def permute_fine(self, nodes):
        nodes = _i32(nodes)
        return self.lib.cpu_parth_permute_fine(self.h, _p(nodes), len(nodes))

    def submesh(self, node, coarse=False):
        n, nnz = C.c_int(), C.c_int()
        self.lib.cpu_parth_submesh(self.h, int(node), int(coarse), C.byref(n), C.byref(nnz), None, None, None)
        Mp = np.empty(n.value + 1, np.int32)
        Mi = np.empty(max(nnz.value, 1), np.int32)
        l2g = np.empty(max(n.value, 1), np.int32)
        self.lib.cpu_parth_submesh(self.h, int(node), int(coarse), C.byref(n), C.byref(nnz), _p(Mp), _p(Mi), _p(l2g))
        return Mp, Mi[:nnz.value], l2g[:n.value]

    def mapMeshPermToMatrixPerm(self, mesh_perm, dim):
        mesh_perm = _i32(mesh_perm)
        out = np.empty(len(mesh_perm) * dim, np.int32)
        self.lib.cpu_parth_map_mesh_perm(self.h, _p(mesh_perm), len(mesh_perm), int(dim), _p(out))
        return out


def amd(Ap, Ai, kind="port"):
    Ap, Ai = _i32(Ap), _i32(Ai)
    n = len(Ap) - 1
    P = np.empty(max(n, 1), np.int32)
    load(kind).cpu_parth_amd(n, _p(Ap), _p(Ai), _p(P))
    return P[:n]


def etree(Ap, Ai, perm, kind="port"):
    Ap, Ai, perm = _i32(Ap), _i32(Ai), _i32(perm)
    n = len(Ap) - 1
    parent = np.empty(max(n, 1), np.int32)
    r = load(kind).cpu_parth_etree(n, _p(Ap), _p(Ai), _p(perm), _p(parent))
    assert r == 0
    return parent[:n]


def colcount(Ap, Ai, perm, parent, kind="port"):
    Ap, Ai, perm, parent = _i32(Ap), _i32(Ai), _i32(perm), _i32(parent)
    n = len(Ap) - 1
    cc = np.empty(max(n, 1), np.int32)
    lnz = load(kind).cpu_parth_colcount(n, _p(Ap), _p(Ai), _p(perm), _p(parent), _p(cc))
    return int(lnz), cc[:n]


def supernodes(parent, colcount):
    """(f4) fundamental supernodes (port restatement): dict(super, supermap, sparent, nsuper, ssize, xsize)"""
    parent, colcount = _i32(parent), _i32(colcount)
    n = len(parent)
    sup, smap, spar = np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.int32)
    sizes = (C.c_int64 * 2)()
    k = load("port").cpu_parth_supernodes(n, _p(parent), _p(colcount), _p(sup), _p(smap), _p(spar), sizes)
    return dict(super=sup[:k + 1], supermap=smap[:n], sparent=spar[:k], nsuper=k, ssize=int(sizes[0]), xsize=int(sizes[1]))
