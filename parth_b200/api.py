"""Python host mirror of PARTH::ParthAPI over the C ABI (include/parth_b200.h).

Same method names, argument meaning and defaults as the reference class
(reference src/Parth/include/ParthAPI.h:16-145), so the parity tests read like the reference's own
callers (examples/api_demos/quick_start.cpp:72-118).  This module only marshals numpy arrays /
device pointers into the C entry points; there is no compute here and no CPU fallback: loading
fails loudly when the CUDA extension is missing, and creating a handle fails without a GPU.
"""
import ctypes as C
import weakref
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libparth_b200.so")

OK = 0
NEEDS_REDISSECT = 100
NEEDS_EXCHANGE = 101
METIS, AMD, MORTON_CODE = 0, 1, 2
COARSE_REPAIR, COARSE_CALLBACK, COARSE_REPORT, COARSE_RELOCATE = 0, 1, 2, 3

_ip = C.POINTER(C.c_int)
_lib = None


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("build_ms", "detect_ms", "integrate_ms", "dirty_ms", "reorder_ms",
                                          "assemble_ms", "expand_ms", "symbolic_ms", "build_kernel_ms")] + [
        ("kernels_launched", C.c_longlong), ("build_path", C.c_int),
        ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("integrate_passes", C.c_int), ("integrate_sweeps", C.c_int),
        ("symbolic_etree_ms", C.c_double), ("diff_kernel_ms", C.c_double), ("symbolic_path", C.c_int),
        ("build_fused_detect", C.c_int), ("proof_kernel_ms", C.c_double), ("build_deferred", C.c_int),
        ("expand_speculative", C.c_int), ("morton_passes", C.c_int), ("morton_ms", C.c_double),
        ("supernodes_ms", C.c_double)]


DECOMPOSER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _ip, _ip, C.c_int, C.c_int, _ip, _ip, _ip, _ip)

# every symbol include/parth_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "parth_b200_version": (C.c_char_p, []),
    "parth_b200_last_error": (C.c_char_p, [_vp]),
    "parth_b200_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "parth_b200_destroy": (C.c_int, [_vp]),
    "parth_b200_clear": (C.c_int, [_vp]),
    "parth_b200_set_options": (C.c_int, [_vp] + [C.c_int] * 7),
    "parth_b200_set_coarse_policy": (C.c_int, [_vp, C.c_int]),
    "parth_b200_set_assume_symmetric": (C.c_int, [_vp, C.c_int]),
    "parth_b200_set_matrix": (C.c_int, [_vp, C.c_int, _ip, _ip, C.c_int]),
    "parth_b200_set_matrix_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int]),
    "parth_b200_set_mesh": (C.c_int, [_vp, C.c_int, _ip, _ip]),
    "parth_b200_set_mesh_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int]),
    "parth_b200_set_new_to_old_map": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_mesh_dims": (C.c_int, [_vp, _ip, _ip]),
    "parth_b200_get_mesh": (C.c_int, [_vp, _ip, _ip]),
    "parth_b200_mesh_dev": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "parth_b200_import_tree": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int] + [_ip] * 6),
    "parth_b200_tree_dims": (C.c_int, [_vp, _ip, _ip, _ip]),
    "parth_b200_export_tree": (C.c_int, [_vp] + [_ip] * 6),
    "parth_b200_set_decomposer": (C.c_int, [_vp, DECOMPOSER, _vp]),
    "parth_b200_use_builtin_decomposer": (C.c_int, [_vp, C.c_int]),
    "parth_b200_compute_permutation": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_compute_permutation_dev": (C.c_int, [_vp, _vp, C.c_int]),
    "parth_b200_accept_snapshot": (C.c_int, [_vp]),
    "parth_b200_get_dof_maps": (C.c_int, [_vp, _ip, _ip, _ip, _ip, _ip, _ip]),
    "parth_b200_get_snapshot": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "parth_b200_map_mesh_perm_to_matrix_perm": (C.c_int, [_vp, _ip, C.c_int, _ip, C.c_int]),
    "parth_b200_map_mesh_perm_to_matrix_perm_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int]),
    "parth_b200_get_reuse": (C.c_double, [_vp]),
    "parth_b200_get_num_changes": (C.c_int, [_vp]),
    "parth_b200_get_changed_edges": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_get_dirty_nodes": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "parth_b200_get_timers": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "parth_b200_get_mesh_perm": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_mesh_perm_dev": (C.c_int, [_vp, C.POINTER(_vp)]),
    "parth_b200_stage_integrate": (C.c_int, [_vp]),
    "parth_b200_stage_detect": (C.c_int, [_vp, _ip]),
    "parth_b200_stage_permute_fine": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_stage_submesh": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]),
    "parth_b200_amd_batch": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip]),
    "parth_b200_set_coordinates": (C.c_int, [_vp, _vp, C.c_int, C.c_longlong, C.c_longlong, C.c_int]),
    "parth_b200_set_coordinates_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_longlong, C.c_longlong, C.c_int]),
    "parth_b200_morton_order": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_morton_order_dev": (C.c_int, [_vp, C.POINTER(_vp), _ip]),
    "parth_b200_morton_keys": (C.c_int, [_vp, _vp, C.c_int]),
    "parth_b200_set_new_to_old_map_dev": (C.c_int, [_vp, _vp, C.c_int]),
    "parth_b200_pin_host": (C.c_int, [_vp, C.c_size_t]),
    "parth_b200_unpin_host": (C.c_int, [_vp]),
    "parth_b200_compute_permutation_block": (C.c_int, [_vp, _ip, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "parth_b200_symbolic": (C.c_int, [_vp, _ip, _ip, _ip, C.POINTER(C.c_int64)]),
    "parth_b200_symbolic_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "parth_b200_supernodes": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip, _ip, _ip, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64)]),
    "parth_b200_supernodes_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int64)]),
    "parth_b200_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "parth_b200_synchronize": (C.c_int, [_vp]),
    "parth_b200_stream": (_vp, [_vp]),
    "parth_b200_set_shard": (C.c_int, [_vp, C.c_int, C.c_int]),
    "parth_b200_exchange_plan": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "parth_b200_pack_labels_dev": (C.c_int, [_vp, _ip, C.c_int, _vp]),
    "parth_b200_unpack_labels_dev": (C.c_int, [_vp, _ip, C.c_int, _vp]),
    "parth_b200_finish_permutation": (C.c_int, [_vp, _ip, C.c_int]),
    "parth_b200_finish_permutation_dev": (C.c_int, [_vp, _vp, C.c_int]),
    "parth_b200_comm_unique_id": (C.c_int, [_vp]),
    "parth_b200_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "parth_b200_comm_init_local": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "parth_b200_comm_info": (C.c_int, [_vp, _ip, _ip, C.POINTER(C.c_char_p)]),
    "parth_b200_block_range": (C.c_int, [_vp, C.c_int, C.c_int, _ip, _ip]),
    "parth_b200_set_matrix_block": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "parth_b200_set_matrix_block_dev": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int]),
}


TREE_MAGIC = 0x54324250  # "PB2T"


def save_tree_file(path, M_n, levels, nodes, meta, node_ptr, dofs, labels, Mp, Mi):
    """the separator-tree fixture format of PARTH::ParthAPI::saveTree (include/parth/ParthAPI.h): int32 little-endian
    "PB2T", 1, M_n, levels, nodes, nnz, meta[nodes*7], node_ptr[nodes+1], dofs[M_n], labels[M_n], Mp[M_n+1], Mi[nnz]"""
    Mp, Mi = _i32(Mp), _i32(Mi)
    nnz = int(Mp[-1])
    with open(path, "wb") as f:
        np.array([TREE_MAGIC, 1, M_n, levels, nodes, nnz], "<i4").tofile(f)
        for a, n in ((meta, nodes * 7), (node_ptr, nodes + 1), (dofs, M_n), (labels, M_n), (Mp, M_n + 1), (Mi, nnz)):
            a = np.ascontiguousarray(np.asarray(a).reshape(-1)[:n], "<i4")
            if len(a) != n:
                raise ValueError("save_tree_file: array shorter than the header says")
            a.tofile(f)


def load_tree_file(path):
    with open(path, "rb") as f:
        hdr = np.fromfile(f, "<i4", 6)
        if len(hdr) != 6 or hdr[0] != TREE_MAGIC or hdr[1] != 1:
            raise ValueError("not a parth_b200 tree fixture: " + path)
        M_n, levels, nodes, nnz = (int(x) for x in hdr[2:])
        out = dict(M_n=M_n, levels=levels, nodes=nodes)
        for k, n in (("meta", nodes * 7), ("node_ptr", nodes + 1), ("dofs", M_n), ("labels", M_n), ("Mp", M_n + 1), ("Mi", nnz)):
            a = np.fromfile(f, "<i4", n)
            if len(a) != n:
                raise ValueError("truncated tree fixture: " + path)
            out[k] = a.astype(np.int32)
        out["meta"] = out["meta"].reshape(nodes, 7)
    return out


class ParthError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"parth_b200 status {code}: {msg}")
        self.code = code


def load_library():
    """dlopen the in-tree CUDA extension; no fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m parth_b200.build` "
                          "(the product path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_ip)


class ParthAPI:
    """Drop-in for PARTH::ParthAPI; one instance owns one device handle."""

    def __init__(self, device=-1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.parth_b200_create(C.byref(h), int(device))
        if rc != 0:
            raise ParthError(rc, self.lib.parth_b200_last_error(None).decode())
        self.h = h
        # option fields of the reference class (ParthAPI.h:19-25, Integrator.h:24-28)
        self.reorder_type = METIS
        self.verbose = True
        self.activate_aggressive_reuse = False
        self.lagging = True
        self.ND_levels = 8
        self.num_cores = 10
        self.relocate_lvl = 2
        self.lag_lvl = 5
        self._decomposer_ref = None
        # COARSE_REPORT: the C ABI leaves the snapshot where it was after PARTH_B200_NEEDS_REDISSECT.  The parity
        # tests compare multi-update sequences with the oracle port, which snapshots after reporting (the reference
        # always finishes the update), so this mirror accepts the snapshot on the caller's behalf by default.
        self.report_advances_snapshot = True
        self._push_options()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.parth_b200_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers
    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise ParthError(rc, self.lib.parth_b200_last_error(self.h).decode())
        return rc

    def _push_options(self):
        # the public fields may be written directly (like the reference's): push them when they changed
        opts = (int(self.reorder_type), int(self.ND_levels), int(self.lagging), int(self.activate_aggressive_reuse),
                int(self.relocate_lvl), int(self.lag_lvl), int(self.verbose))
        if opts != getattr(self, "_pushed", None):
            self._check(self.lib.parth_b200_set_options(self.h, *opts))
            self._pushed = opts

    # ---- options (ParthAPI.cpp:13-45)
    def setReorderingType(self, t):
        self.reorder_type = int(t); self._push_options()

    def getReorderingType(self):
        return self.reorder_type

    def setVerbose(self, v):
        self.verbose = bool(v); self._push_options()

    def getVerbose(self):
        return self.verbose

    def setNDLevels(self, n):
        self.ND_levels = int(n); self._push_options()

    def getNDLevels(self):
        return self.ND_levels

    def setNumberOfCores(self, n):
        self.num_cores = int(n)  # stored and never used, like the reference (ParthAPI.cpp:34)

    def getNumberOfCores(self):
        return self.num_cores

    def setRelocatedLevel(self, l):
        self.relocate_lvl = int(l); self._push_options()

    def getRelocatedLevel(self):
        return self.relocate_lvl

    def setLagLevel(self, l):
        self.lag_lvl = int(l); self._push_options()

    def getLagLevel(self):
        return self.lag_lvl

    def setCoarsePolicy(self, p):
        self._check(self.lib.parth_b200_set_coarse_policy(self.h, int(p)))

    def setAssumeSymmetric(self, on):
        self._check(self.lib.parth_b200_set_assume_symmetric(self.h, int(on)))

    def useBuiltinDecomposer(self, on=True):
        self._check(self.lib.parth_b200_use_builtin_decomposer(self.h, int(on)))

    def setShard(self, rank, world):
        self._check(self.lib.parth_b200_set_shard(self.h, int(rank), int(world)))

    def clearParth(self):
        self._check(self.lib.parth_b200_clear(self.h))

    # ---- one mesh over several GPUs: row blocks (include/parth_b200.h, stage_sharded.cu)
    @staticmethod
    def commUniqueId():
        """128 bytes from rank 0 (NCCL unique id) to hand to every rank's commInit"""
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.parth_b200_comm_unique_id(buf)
        if rc != 0:
            raise ParthError(rc, lib.parth_b200_last_error(None).decode())
        return buf.raw

    def commInit(self, unique_id, rank, world):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.parth_b200_comm_init(self.h, buf, int(rank), int(world)))

    @staticmethod
    def commInitLocal(apis):
        """the ranks are the given ParthAPI objects of this process, one host thread each (test plumbing)"""
        lib = load_library()
        arr = (C.c_void_p * len(apis))(*[a.h for a in apis])
        rc = lib.parth_b200_comm_init_local(arr, len(apis))
        if rc != 0:
            raise ParthError(rc, "comm_init_local failed")

    def commInfo(self):
        r, w, k = C.c_int(), C.c_int(), C.c_char_p()
        self._check(self.lib.parth_b200_comm_info(self.h, C.byref(r), C.byref(w), C.byref(k)))
        return r.value, w.value, (k.value or b"").decode()

    def blockRange(self, N, nnz):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_block_range(self.h, int(N), int(nnz), C.byref(a), C.byref(b)))
        return a.value, b.value

    def setMatrixBlock(self, N, nnz, Ap, Ai, dim=1):
        """Ap / Ai: whole host arrays (int32, contiguous) of which only this rank's block is read"""
        self._push_options()
        Ap, Ai = _i32(Ap), _i32(Ai)
        self._check(self.lib.parth_b200_set_matrix_block(self.h, int(N), int(nnz), C.c_void_p(Ap.ctypes.data),
                                                         C.c_void_p(Ai.ctypes.data), int(dim)))

    def setMatrixBlockPtr(self, N, nnz, Ap_addr, Ai_addr, dim=1):
        """host addresses in GLOBAL indexing (a caller that holds only its block passes them biased by -c0 / -Ap[c0])"""
        self._push_options()
        self._check(self.lib.parth_b200_set_matrix_block(self.h, int(N), int(nnz), C.c_void_p(Ap_addr),
                                                         C.c_void_p(Ai_addr), int(dim)))

    def setMatrixBlockDev(self, N, nnz, dAp_ptr, dAi_ptr, q0, q1, dim=1):
        self._push_options()
        self._check(self.lib.parth_b200_set_matrix_block_dev(self.h, int(N), int(nnz), C.c_void_p(dAp_ptr),
                                                             C.c_void_p(dAi_ptr), int(dim), int(q0), int(q1)))

    # ---- inputs
    def setMatrix(self, N, Ap, Ai, dim=1, new_to_old=None):
        self._push_options()
        Ap, Ai = _i32(Ap), _i32(Ai)
        self._check(self.lib.parth_b200_set_matrix(self.h, int(N), _p(Ap), _p(Ai), int(dim)))
        self.setNewToOldDOFMap(new_to_old)

    def setMatrixDev(self, N, dAp_ptr, dAi_ptr, nnz, dim=1):
        self._push_options()
        self._check(self.lib.parth_b200_set_matrix_dev(self.h, int(N), C.c_void_p(dAp_ptr), C.c_void_p(dAi_ptr),
                                                       int(nnz), int(dim)))

    def setMesh(self, n, Mp, Mi, new_to_old=None):
        self._push_options()
        Mp, Mi = _i32(Mp), _i32(Mi)
        self._check(self.lib.parth_b200_set_mesh(self.h, int(n), _p(Mp), _p(Mi)))
        self.setNewToOldDOFMap(new_to_old)

    def setMeshDev(self, n, dMp_ptr, dMi_ptr, nnz):
        self._push_options()
        self._check(self.lib.parth_b200_set_mesh_dev(self.h, int(n), C.c_void_p(dMp_ptr), C.c_void_p(dMi_ptr), int(nnz)))

    def setNewToOldDOFMap(self, new_to_old):
        if new_to_old is None or len(new_to_old) == 0:
            self._check(self.lib.parth_b200_set_new_to_old_map(self.h, None, 0))
        else:
            m = _i32(new_to_old)
            self._check(self.lib.parth_b200_set_new_to_old_map(self.h, _p(m), len(m)))

    def setNewToOldDOFMapDev(self, dmap_ptr, length):
        self._check(self.lib.parth_b200_set_new_to_old_map_dev(self.h, C.c_void_p(dmap_ptr), int(length)))

    # ---- mesh / tree access
    def mesh_dims(self):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_mesh_dims(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def mesh(self):
        n, nnz = self.mesh_dims()
        Mp = np.empty(n + 1, np.int32)
        Mi = np.empty(max(nnz, 1), np.int32)
        self._check(self.lib.parth_b200_get_mesh(self.h, _p(Mp), _p(Mi)))
        return Mp, Mi[:nnz]

    def import_tree(self, tree, Mp_prev=None, Mi_prev=None):
        Mp_prev = None if Mp_prev is None else _i32(Mp_prev)
        Mi_prev = None if Mi_prev is None else _i32(Mi_prev)
        meta, node_ptr, dofs, labels = _i32(tree.meta), _i32(tree.node_ptr), _i32(tree.dofs), _i32(tree.labels)
        self._check(self.lib.parth_b200_import_tree(self.h, tree.M_n, tree.levels, tree.nodes, _p(meta), _p(node_ptr),
                                                    _p(dofs), _p(labels), _p(Mp_prev), _p(Mi_prev)))

    def tree_dims(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_tree_dims(self.h, C.byref(a), C.byref(b), C.byref(c)), allow=(-3,))
        return a.value, b.value, c.value

    def export_tree(self):
        M_n, lv, nodes = self.tree_dims()
        out = dict(M_n=M_n, levels=lv, nodes=nodes,
                   meta=np.empty((nodes, 7), np.int32), node_ptr=np.empty(nodes + 1, np.int32),
                   dofs=np.empty(max(M_n, 1), np.int32), labels=np.empty(max(M_n, 1), np.int32),
                   dof2node=np.empty(max(M_n, 1), np.int32), mesh_perm=np.empty(max(M_n, 1), np.int32))
        self._check(self.lib.parth_b200_export_tree(self.h, _p(out["meta"]), _p(out["node_ptr"]), _p(out["dofs"]),
                                                    _p(out["labels"]), _p(out["dof2node"]), _p(out["mesh_perm"])))
        for k in ("dofs", "labels", "dof2node", "mesh_perm"):
            out[k] = out[k][:M_n]
        return out

    def setDecomposer(self, fn):
        """fn(M_n, Mp, Mi, levels, nodes) -> (meta, node_ptr, dofs, labels) numpy arrays"""
        def tramp(user, M_n, Mp, Mi, levels, nodes, meta, node_ptr, dofs, labels):
            try:
                mp = np.ctypeslib.as_array(Mp, (M_n + 1,))
                mi = np.ctypeslib.as_array(Mi, (max(int(mp[M_n]), 1),))[:int(mp[M_n])]
                m, npt, d, l = fn(M_n, mp, mi, levels, nodes)
                np.ctypeslib.as_array(meta, (nodes * 7,))[:] = _i32(m).reshape(-1)
                np.ctypeslib.as_array(node_ptr, (nodes + 1,))[:] = _i32(npt)
                np.ctypeslib.as_array(dofs, (M_n,))[:] = _i32(d)
                np.ctypeslib.as_array(labels, (M_n,))[:] = _i32(l)
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return -5
        self._decomposer_ref = DECOMPOSER(tramp)
        self._check(self.lib.parth_b200_set_decomposer(self.h, self._decomposer_ref, None))

    # ---- the path
    def computePermutation(self, dim=3, out=None):
        """returns (status, perm): status 0, or NEEDS_REDISSECT under COARSE_REPORT.
        `out`: optional caller-owned int32 buffer of at least M_n*dim entries (e.g. pinned memory)"""
        self._push_options()
        n, _ = self.mesh_dims()
        if out is not None:
            if out.dtype != np.int32 or not out.flags.c_contiguous or out.size < n * dim:
                raise ValueError("out must be a contiguous int32 array of at least M_n*dim entries")
            perm = out
        else:
            perm = np.empty(max(n * dim, 1), np.int32)
        rc = self._check(self.lib.parth_b200_compute_permutation(self.h, _p(perm), int(dim)), allow=(NEEDS_REDISSECT, NEEDS_EXCHANGE))
        if rc == NEEDS_REDISSECT and self.report_advances_snapshot:
            self.acceptSnapshot()
        return rc, perm[:n * dim]

    def computePermutationBlock(self, dim=3, out=None):
        """sharded mesh: this rank's block of the expanded permutation -> (status, first, block)"""
        self._push_options()
        n, _ = self.mesh_dims()
        if out is None:
            out = np.empty(max(n * dim, 1), np.int32)
        first, count = C.c_int(), C.c_int()
        rc = self._check(self.lib.parth_b200_compute_permutation_block(self.h, _p(out), int(dim), C.byref(first),
                                                                       C.byref(count)), allow=(NEEDS_REDISSECT,))
        return rc, first.value, out[:count.value]

    def computePermutationDev(self, dperm_ptr, dim=3):
        self._push_options()
        rc = self._check(self.lib.parth_b200_compute_permutation_dev(self.h, C.c_void_p(dperm_ptr), int(dim)),
                         allow=(NEEDS_REDISSECT, NEEDS_EXCHANGE))
        if rc == NEEDS_REDISSECT and self.report_advances_snapshot:
            self.acceptSnapshot()
        return rc

    def preparedUpdateDev(self, N, dAp_ptr, dAi_ptr, nnz, dim, dperm_ptr, out_dim):
        """One resident update (setMatrixDev + computePermutationDev on fixed device arrays) as a closure whose ctypes
        arguments are built once.  The generic methods cost ~25 us of Python per update (option push, argument
        conversion) -- more than the device idles between two updates -- a hot loop over a fixed set of buffers
        binds them once instead.  Options are pushed when the closure is made; same status handling as
        computePermutationDev.  Returns the closure; calling it returns the status of the update."""
        self._push_options()
        set_fn, cp_fn = self.lib.parth_b200_set_matrix_dev, self.lib.parth_b200_compute_permutation_dev
        a_set = (self.h, C.c_int(int(N)), C.c_void_p(dAp_ptr), C.c_void_p(dAi_ptr), C.c_int(int(nnz)), C.c_int(int(dim)))
        a_cp = (self.h, C.c_void_p(dperm_ptr), C.c_int(int(out_dim)))

        me = weakref.ref(self)  # the closure may be stored on the object itself: no reference cycle

        def run():
            rc = set_fn(*a_set)
            if rc:
                me()._check(rc)
            rc = cp_fn(*a_cp)
            if rc:
                me()._check(rc, allow=(NEEDS_REDISSECT, NEEDS_EXCHANGE))
                if rc == NEEDS_REDISSECT and me().report_advances_snapshot:
                    me().acceptSnapshot()
            return rc
        return run

    def dof_maps(self):
        """(old_to_new_map, added_dofs, removed_dofs) as ParthAPI::computeOldToNewDOFMap leaves them"""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_get_dof_maps(self.h, C.byref(a), None, C.byref(b), None, C.byref(c), None))
        o2n, add, rem = (np.empty(max(x.value, 1), np.int32) for x in (a, b, c))
        self._check(self.lib.parth_b200_get_dof_maps(self.h, C.byref(a), _p(o2n), C.byref(b), _p(add), C.byref(c), _p(rem)))
        return o2n[:a.value], add[:b.value], rem[:c.value]

    def snapshot(self):
        """(Mp_prev, Mi_prev) of the handle"""
        m, z = C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_get_snapshot(self.h, C.byref(m), C.byref(z), None, None))
        Mp, Mi = np.zeros(m.value + 1, np.int32), np.empty(max(z.value, 1), np.int32)
        self._check(self.lib.parth_b200_get_snapshot(self.h, C.byref(m), C.byref(z), _p(Mp), _p(Mi)))
        return Mp, Mi[:z.value]

    def acceptSnapshot(self):
        """after NEEDS_REDISSECT: advance the snapshot to the current graph (parth_b200_accept_snapshot)"""
        self._check(self.lib.parth_b200_accept_snapshot(self.h))

    # ---- multi-GPU exchange (see parth_b200/sharded.py)
    def exchange_plan(self):
        n = C.c_int()
        self._check(self.lib.parth_b200_exchange_plan(self.h, C.byref(n), None, None, None))
        nodes, owner, sizes = (np.empty(max(n.value, 1), np.int32) for _ in range(3))
        self._check(self.lib.parth_b200_exchange_plan(self.h, C.byref(n), _p(nodes), _p(owner), _p(sizes)))
        return nodes[:n.value].copy(), owner[:n.value].copy(), sizes[:n.value].copy()

    def pack_labels_dev(self, nodes, dbuf_ptr):
        nodes = _i32(nodes)
        self._check(self.lib.parth_b200_pack_labels_dev(self.h, _p(nodes), len(nodes), C.c_void_p(dbuf_ptr)))

    def unpack_labels_dev(self, nodes, dbuf_ptr):
        nodes = _i32(nodes)
        self._check(self.lib.parth_b200_unpack_labels_dev(self.h, _p(nodes), len(nodes), C.c_void_p(dbuf_ptr)))

    def finishPermutationDev(self, dperm_ptr, dim=3):
        self._check(self.lib.parth_b200_finish_permutation_dev(self.h, C.c_void_p(dperm_ptr), int(dim)))

    def finishPermutation(self, dim=3, out=None):
        n, _ = self.mesh_dims()
        perm = out if out is not None else np.empty(max(n * dim, 1), np.int32)
        self._check(self.lib.parth_b200_finish_permutation(self.h, _p(perm), int(dim)))
        return perm[:n * dim]

    def mapMeshPermToMatrixPerm(self, mesh_perm, dim=-1):
        mesh_perm = _i32(mesh_perm)
        d = dim
        if d == -1:
            raise ValueError("pass dim explicitly from Python")
        out = np.empty(len(mesh_perm) * d, np.int32)
        self._check(self.lib.parth_b200_map_mesh_perm_to_matrix_perm(self.h, _p(mesh_perm), len(mesh_perm), _p(out), int(d)))
        return out

    def mapMeshPermToMatrixPermDev(self, dmesh_perm_ptr, n, dout_ptr, dim):
        self._check(self.lib.parth_b200_map_mesh_perm_to_matrix_perm_dev(self.h, C.c_void_p(dmesh_perm_ptr), int(n),
                                                                         C.c_void_p(dout_ptr), int(dim)))

    def getReuse(self):
        return float(self.lib.parth_b200_get_reuse(self.h))

    def getNumChanges(self):
        return int(self.lib.parth_b200_get_num_changes(self.h))

    def changed_edges(self):
        n = self.getNumChanges()
        out = np.empty((max(n, 1), 2), np.int32)
        self._check(min(0, self.lib.parth_b200_get_changed_edges(self.h, _p(out), n)))
        return out[:n]

    def dirty(self):
        _, _, nodes = self.tree_dims()
        f = np.empty(max(nodes, 1), np.int32)
        c = np.empty(max(nodes, 1), np.int32)
        nf, nc = C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_get_dirty_nodes(self.h, _p(f), C.byref(nf), _p(c), C.byref(nc)))
        return f[:nf.value].copy(), c[:nc.value].copy()

    def timers(self):
        t = (C.c_double * 5)()
        self._check(self.lib.parth_b200_get_timers(self.h, t))
        return dict(zip(("init", "integrate", "detect", "compute", "map"), list(t)))

    def mesh_perm(self):
        M_n, _, _ = self.tree_dims()
        out = np.empty(max(M_n, 1), np.int32)
        r = self.lib.parth_b200_get_mesh_perm(self.h, _p(out), out.size)
        if r < 0:
            self._check(r)
        return out[:r]

    def stats(self):
        s = Stats()
        self._check(self.lib.parth_b200_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def synchronize(self):
        self._check(self.lib.parth_b200_synchronize(self.h))

    def stream(self):
        return self.lib.parth_b200_stream(self.h)

    # ---- stage-wise
    def stage_integrate(self):
        self._push_options()
        self._check(self.lib.parth_b200_stage_integrate(self.h))

    def stage_detect(self):
        self._push_options()
        n = C.c_int()
        self._check(self.lib.parth_b200_stage_detect(self.h, C.byref(n)))
        return n.value

    def stage_permute_fine(self, nodes):
        self._push_options()
        nodes = _i32(nodes)
        self._check(self.lib.parth_b200_stage_permute_fine(self.h, _p(nodes), len(nodes)))

    def stage_submesh(self, node, coarse=False):
        n, nnz = C.c_int(), C.c_int()
        self._check(self.lib.parth_b200_stage_submesh(self.h, int(node), int(coarse), C.byref(n), C.byref(nnz), None, None, None))
        Mp = np.empty(n.value + 1, np.int32)
        Mi = np.empty(max(nnz.value, 1), np.int32)
        l2g = np.empty(max(n.value, 1), np.int32)
        self._check(self.lib.parth_b200_stage_submesh(self.h, int(node), int(coarse), C.byref(n), C.byref(nnz), _p(Mp), _p(Mi), _p(l2g)))
        return Mp, Mi[:nnz.value], l2g[:n.value]

    def amd_batch(self, graphs):
        """graphs: list of (Mp, Mi); returns list of permutations"""
        gp = np.zeros(len(graphs) + 1, np.int32)
        for i, (mp, _) in enumerate(graphs):
            gp[i + 1] = gp[i] + len(mp) - 1
        Mp = np.concatenate([_i32(mp) for mp, _ in graphs]) if graphs else np.zeros(1, np.int32)
        Mi = np.concatenate([_i32(mi) for _, mi in graphs] + [np.zeros(1, np.int32)])
        perm = np.empty(max(int(gp[-1]), 1), np.int32)
        self._check(self.lib.parth_b200_amd_batch(self.h, len(graphs), _p(gp), _p(Mp), _p(Mi), _p(perm)))
        return [perm[gp[i]:gp[i + 1]].copy() for i in range(len(graphs))]

    # ---- ReorderingType MORTON_CODE (extension: the reference leaves it unimplemented, HMD.cpp:82-83)
    def setCoordinates(self, xyz, bits=21):
        """xyz: float64 array [n, 3] (C order: interleaved points; Fortran order: an Eigen column-major V)"""
        xyz = np.asarray(xyz, dtype=np.float64)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError("setCoordinates: xyz must be [n, 3]")
        n = xyz.shape[0]
        if xyz.flags.f_contiguous and not xyz.flags.c_contiguous:
            ps, cs = 1, n
        else:
            xyz = np.ascontiguousarray(xyz)
            ps, cs = 3, 1
        self._xyz_keep = xyz
        self._check(self.lib.parth_b200_set_coordinates(self.h, C.c_void_p(xyz.ctypes.data), n, ps, cs, int(bits)))

    def setCoordinatesDev(self, dxyz_ptr, n, point_stride=3, comp_stride=1, bits=21):
        self._check(self.lib.parth_b200_set_coordinates_dev(self.h, C.c_void_p(dxyz_ptr), int(n), int(point_stride),
                                                            int(comp_stride), int(bits)))

    def mortonOrder(self):
        """global Morton order of the points given to setCoordinates: perm[new] = old"""
        d, n = C.c_void_p(), C.c_int()
        self._check(self.lib.parth_b200_morton_order_dev(self.h, C.byref(d), C.byref(n)))
        perm = np.empty(max(n.value, 1), np.int32)
        self._check(self.lib.parth_b200_morton_order(self.h, _p(perm), len(perm)))
        return perm[:n.value]

    def mortonOrderDev(self):
        d, n = C.c_void_p(), C.c_int()
        self._check(self.lib.parth_b200_morton_order_dev(self.h, C.byref(d), C.byref(n)))
        return d.value, n.value

    def mortonKeys(self):
        """the sorted keys (same order as mortonOrder)"""
        d, n = C.c_void_p(), C.c_int()
        self._check(self.lib.parth_b200_morton_order_dev(self.h, C.byref(d), C.byref(n)))
        keys = np.empty(max(n.value, 1), np.uint64)
        self._check(self.lib.parth_b200_morton_keys(self.h, C.c_void_p(keys.ctypes.data), len(keys)))
        return keys[:n.value]

    def symbolicDev(self, dperm_ptr, dparent_ptr, dcolcount_ptr):
        """stage (d) with device pointers (0 = NULL): returns nnz(L); etree / column counts stay on the device"""
        lnz = C.c_int64()
        self._check(self.lib.parth_b200_symbolic_dev(self.h, C.c_void_p(dperm_ptr or None), C.c_void_p(dparent_ptr or None),
                                                     C.c_void_p(dcolcount_ptr or None), C.byref(lnz)))
        return int(lnz.value)

    def supernodes(self, parent, colcount):
        """(f4) fundamental supernodes of L from etree + column counts (host arrays in and out):
        returns dict(super, supermap, sparent, nsuper, ssize, xsize)"""
        parent, colcount = _i32(parent), _i32(colcount)
        n = len(parent)
        sup = np.empty(n + 1, np.int32)
        smap = np.empty(max(n, 1), np.int32)
        spar = np.empty(max(n, 1), np.int32)
        ns, a, b = C.c_int(), C.c_int64(), C.c_int64()
        self._check(self.lib.parth_b200_supernodes(self.h, n, _p(parent), _p(colcount), _p(sup), _p(smap), _p(spar),
                                                   C.byref(ns), C.byref(a), C.byref(b)))
        k = ns.value
        return dict(super=sup[:k + 1], supermap=smap[:n], sparent=spar[:k], nsuper=k, ssize=int(a.value), xsize=int(b.value))

    def supernodesDev(self, n, dparent_ptr, dcolcount_ptr, dsuper_ptr=0, dsupermap_ptr=0, dsparent_ptr=0):
        """device pointers (0 = NULL for the outputs): returns (nsuper, ssize, xsize)"""
        ns, a, b = C.c_int(), C.c_int64(), C.c_int64()
        self._check(self.lib.parth_b200_supernodes_dev(self.h, int(n), C.c_void_p(dparent_ptr), C.c_void_p(dcolcount_ptr),
                                                       C.c_void_p(dsuper_ptr or None), C.c_void_p(dsupermap_ptr or None),
                                                       C.c_void_p(dsparent_ptr or None), C.byref(ns), C.byref(a), C.byref(b)))
        return ns.value, int(a.value), int(b.value)

    def symbolic(self, perm=None):
        n, _ = self.mesh_dims()
        parent = np.empty(max(n, 1), np.int32)
        cc = np.empty(max(n, 1), np.int32)
        lnz = C.c_int64()
        p = None if perm is None else _i32(perm)
        self._check(self.lib.parth_b200_symbolic(self.h, _p(p), _p(parent), _p(cc), C.byref(lnz)))
        return parent[:n], cc[:n], int(lnz.value)
