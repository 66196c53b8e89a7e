"""Synthetic inputs for the five BASELINE.json configs, plus small host utilities.

Nothing here is on the hot path: these build the CSC matrices / mesh graphs / separator-tree
fixtures that tests and bench.py feed to the library.  Generators follow the shapes named in
SURVEY.md section 8(d):
  * poisson3d       7-point grid, index z + y*n + x*n^2, full symmetric pattern with diagonal
                    (shape of compute3DPossion, reference src/utils/Parth_utils.cpp:40-85)
  * kron_blocks     graph (x) dim x dim full blocks  (dim > 1 inputs of setMatrix)
  * geometric_tree  recursive-coordinate-bisection separator tree in the flat fixture layout
                    (stand-in for the METIS tree where METIS cannot run)
"""
import numpy as np

FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
_MASK = (1 << 64) - 1


def fnv1a64(a):
    """FNV-1a-64 over an int32 stream, one xor-multiply per element."""
    a = np.ascontiguousarray(a, dtype=np.int32).view(np.uint32)
    x = FNV_OFFSET
    for v in a.tolist():
        x ^= v
        x = (x * FNV_PRIME) & _MASK
    return x


def hexhash(a):
    return "%016x" % fnv1a64(a)


def read_mtx_pattern(path):
    """Matrix-Market coordinate file -> (N, Ap, Ai) CSC pattern, columns sorted by row."""
    with open(path) as f:
        header = f.readline()
        sym = "symmetric" in header.lower()
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        nr, nc, nnz = (int(t) for t in line.split())
        data = np.loadtxt(f, usecols=(0, 1), dtype=np.int64, ndmin=2)
    r = data[:, 0] - 1
    c = data[:, 1] - 1
    if sym:
        off = r != c
        r, c = np.concatenate([r, c[off]]), np.concatenate([c, r[off]])
    return coo_to_csc(nc, r, c)


def coo_to_csc(n, rows, cols):
    key = cols.astype(np.int64) * n + rows.astype(np.int64)
    key = np.unique(key)
    c = (key // n).astype(np.int32)
    r = (key % n).astype(np.int32)
    Ap = np.zeros(n + 1, np.int64)
    Ap[1:] = np.cumsum(np.bincount(c, minlength=n))
    return int(n), Ap.astype(np.int32), r


def poisson3d(n, diagonal=True, lower_only=False):
    """Full symmetric 7-point pattern on an n^3 grid -> (N, Ap, Ai), columns ascending."""
    N = n * n * n
    ind = np.arange(N, dtype=np.int64)
    z = ind % n
    y = (ind // n) % n
    x = ind // (n * n)
    offs = [(-n * n, x > 0), (-n, y > 0), (-1, z > 0)]
    if diagonal:
        offs.append((0, np.ones(N, bool)))
    if not lower_only:
        offs += [(1, z < n - 1), (n, y < n - 1), (n * n, x < n - 1)]
    if lower_only:
        # "lower triangle + diagonal" in CSC: rows >= col, i.e. the positive offsets
        offs = ([(0, np.ones(N, bool))] if diagonal else []) + [(1, z < n - 1), (n, y < n - 1), (n * n, x < n - 1)]
    valid = np.stack([m for _, m in offs], axis=1)
    cand = np.stack([ind + o for o, _ in offs], axis=1)
    Ap = np.zeros(N + 1, np.int64)
    Ap[1:] = np.cumsum(valid.sum(axis=1))
    Ai = cand[valid].astype(np.int32)
    return N, Ap.astype(np.int32), Ai


def kron_blocks(Ap, Ai, dim):
    """pattern (x) full dim x dim blocks -> CSC with N*dim columns (rows ascending)."""
    n = len(Ap) - 1
    deg = np.diff(Ap).astype(np.int64)
    deg_b = np.repeat(deg * dim, dim)
    Bp = np.zeros(n * dim + 1, np.int64)
    Bp[1:] = np.cumsum(deg_b)
    # column c*dim+d holds rows r*dim+e for every r in col c, e in 0..dim-1, ascending
    rows = (Ai.astype(np.int64)[:, None] * dim + np.arange(dim)[None, :]).reshape(-1)
    starts = Ap[:-1].astype(np.int64) * dim
    out = np.empty(int(Bp[-1]), np.int32)
    col_of_entry = np.repeat(np.arange(n), deg * dim)
    local = np.arange(len(rows)) - np.repeat(starts, deg * dim)
    for d in range(dim):
        pos = Bp[col_of_entry * dim + d] + local
        out[pos] = rows
    return n * dim, Bp.astype(np.int32), out


def mesh_from_pattern(Ap, Ai, dim=1):
    """Small-case numpy restatement of the CSC->mesh graph build for generators that need a graph
    (NOT the oracle: tests use oracle/ for parity)."""
    N = len(Ap) - 1
    M = N // dim
    cols = []
    rows = []
    for c in range(0, N, dim):
        seg = Ai[Ap[c]:Ap[c + 1]:dim] // dim
        rows.append(seg)
        cols.append(np.full(len(seg), c // dim, np.int64))
    r = np.concatenate(rows).astype(np.int64)
    c = np.concatenate(cols)
    off = r != c
    r, c = r[off], c[off]
    _, Mp, Mi = coo_to_csc(M, np.concatenate([r, c]), np.concatenate([c, r]))
    return Mp, Mi


def grid_graph(n):
    """mesh graph (no diagonal) of the n^3 7-point grid, built directly."""
    _, Ap, Ai = poisson3d(n, diagonal=False)
    return Ap, Ai


def add_edges(Mp, Mi, edges):
    """insert undirected edges (array [k,2]) into a sorted CSR graph, rows stay sorted/unique."""
    M = len(Mp) - 1
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    e = e[e[:, 0] != e[:, 1]]
    src = np.repeat(np.arange(M, dtype=np.int64), np.diff(Mp))
    r = np.concatenate([src, e[:, 0], e[:, 1]])
    c = np.concatenate([Mi.astype(np.int64), e[:, 1], e[:, 0]])
    _, P, I = coo_to_csc(M, c, r)  # coo_to_csc sorts by (second arg major) -> pass swapped
    return P, I


def remove_edges(Mp, Mi, edges):
    M = len(Mp) - 1
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    src = np.repeat(np.arange(M, dtype=np.int64), np.diff(Mp))
    key = src * M + Mi.astype(np.int64)
    kill = np.concatenate([e[:, 0] * M + e[:, 1], e[:, 1] * M + e[:, 0]])
    keep = ~np.isin(key, kill)
    Pn = np.zeros(M + 1, np.int64)
    Pn[1:] = np.cumsum(np.bincount(src[keep], minlength=M))
    return Pn.astype(np.int32), Mi[keep].astype(np.int32)


def graph_to_matrix(Mp, Mi, diagonal=True):
    """mesh graph -> full symmetric CSC pattern with diagonal (dim = 1 matrix)."""
    M = len(Mp) - 1
    src = np.repeat(np.arange(M, dtype=np.int64), np.diff(Mp))
    r = Mi.astype(np.int64)
    c = src
    if diagonal:
        d = np.arange(M, dtype=np.int64)
        r = np.concatenate([r, d])
        c = np.concatenate([c, d])
    return coo_to_csc(M, r, c)


def geometric_tree(n, levels):
    """Recursive coordinate bisection of the n^3 grid: bisect the longest
    axis at lo + len/2 with a single-plane separator; leaf when level == levels or len < 3;
    identity labels.  Returns the flat fixture arrays (meta, node_ptr, dofs, labels)."""
    nodes = 2 ** (levels + 1) - 1
    meta = np.full((nodes, 7), -1, np.int32)
    dof_lists = [None] * nodes
    N = n * n * n

    def box_ids(lo, hi):
        xs = np.arange(lo[0], hi[0], dtype=np.int64)
        ys = np.arange(lo[1], hi[1], dtype=np.int64)
        zs = np.arange(lo[2], hi[2], dtype=np.int64)
        ids = (xs[:, None, None] * n * n + ys[None, :, None] * n + zs[None, None, :]).reshape(-1)
        return ids.astype(np.int32)

    # iterative DFS; post-order offsets computed afterwards
    stack = [(0, -1, 0, (0, 0, 0), (n, n, n))]
    while stack:
        nid, parent, lvl, lo, hi = stack.pop()
        lens = [hi[a] - lo[a] for a in range(3)]
        ax = int(np.argmax(lens))
        ln = lens[ax]
        if lvl == levels or ln < 3:
            ids = box_ids(lo, hi)
            dof_lists[nid] = ids
            meta[nid] = (-1, -1, nid, parent, 0, lvl, len(ids))
            continue
        s = lo[ax] + ln // 2
        slo, shi = list(lo), list(hi)
        slo[ax], shi[ax] = s, s + 1
        ids = box_ids(slo, shi)
        dof_lists[nid] = ids
        lhi = list(hi); lhi[ax] = s
        rlo = list(lo); rlo[ax] = s + 1
        meta[nid] = (2 * nid + 1, 2 * nid + 2, nid, parent, 0, lvl, len(ids))
        stack.append((2 * nid + 1, nid, lvl + 1, tuple(lo), tuple(lhi)))
        stack.append((2 * nid + 2, nid, lvl + 1, tuple(rlo), tuple(hi)))

    sizes = np.array([0 if d is None else len(d) for d in dof_lists], np.int64)
    node_ptr = np.zeros(nodes + 1, np.int32)
    node_ptr[1:] = np.cumsum(sizes)
    dofs = np.concatenate([np.sort(d) for d in dof_lists if d is not None]).astype(np.int32)
    # concatenation above skips empty nodes but node_ptr accounts for them with zero length
    labels = np.concatenate([np.arange(len(d), dtype=np.int32) for d in dof_lists if d is not None])
    assert len(dofs) == N
    set_postorder_offsets(meta, sizes)
    return meta, node_ptr, dofs, labels


def set_postorder_offsets(meta, sizes):
    """offset(node) = first + |left subtree| + |right subtree| (reference HMD.cpp:303-337,681-697)."""
    nodes = len(meta)
    sub = np.zeros(nodes, np.int64)
    for i in range(nodes - 1, -1, -1):
        if meta[i, 2] == -1:
            continue
        s = sizes[i]
        if meta[i, 0] != -1:
            s += sub[meta[i, 0]]
        if meta[i, 1] != -1:
            s += sub[meta[i, 1]]
        sub[i] = s
    first = np.zeros(nodes, np.int64)
    for i in range(nodes):
        if meta[i, 2] == -1:
            continue
        l, r = meta[i, 0], meta[i, 1]
        ls = sub[l] if l != -1 else 0
        rs = sub[r] if r != -1 else 0
        if l != -1:
            first[l] = first[i]
        if r != -1:
            first[r] = first[i] + ls
        meta[i, 4] = first[i] + ls + rs
    return sub


def subtree_sizes(meta, node_ptr):
    sizes = np.diff(node_ptr).astype(np.int64)
    nodes = len(meta)
    sub = np.zeros(nodes, np.int64)
    for i in range(nodes - 1, -1, -1):
        if meta[i, 2] == -1:
            continue
        s = sizes[i]
        for ch in (meta[i, 0], meta[i, 1]):
            if ch != -1:
                s += sub[ch]
        sub[i] = s
    return sub


def assemble_mesh_perm(meta, node_ptr, dofs, labels):
    """perm[label + offset(node)] = dof (reference HMD.cpp:348-363), numpy version for fixtures."""
    M = len(dofs)
    node_of = np.repeat(np.arange(len(meta)), np.diff(node_ptr))
    perm = np.empty(M, np.int32)
    perm[labels + meta[node_of, 4]] = dofs
    return perm


def dof_to_node(meta, node_ptr, dofs):
    node_of = np.repeat(np.arange(len(meta)), np.diff(node_ptr)).astype(np.int32)
    out = np.empty(len(dofs), np.int32)
    out[dofs] = node_of
    return out


def lattice_pattern(n, dirs, diagonal=True, dtype=np.int32):
    """Symmetric pattern on an n^3 lattice: vertex (x,y,z) ~ (x,y,z) +/- d for every d in `dirs`
    (index z + y*n + x*n^2).  Returns (N, Ap, Ai) with ascending columns.  Memory-lean: one offset
    at a time, so it also serves the 512^3 config."""
    N = n * n * n
    ind = np.arange(N, dtype=np.int64)
    z = (ind % n).astype(np.int32)
    y = ((ind // n) % n).astype(np.int32)
    x = (ind // (n * n)).astype(np.int32)
    offs = []
    for (dx, dy, dz) in dirs:
        for s in (-1, 1):
            offs.append((s * (dz + dy * n + dx * n * n), s * dx, s * dy, s * dz))
    if diagonal:
        offs.append((0, 0, 0, 0))
    offs.sort()
    masks = []
    cnt = np.zeros(N, np.int32)
    for (o, dx, dy, dz) in offs:
        m = np.ones(N, bool)
        for c, d in ((x, dx), (y, dy), (z, dz)):
            if d > 0:
                m &= c < n - d
            elif d < 0:
                m &= c >= -d
        masks.append(m)
        cnt += m
    Ap = np.zeros(N + 1, np.int64)
    np.cumsum(cnt, out=Ap[1:])
    Ai = np.empty(int(Ap[-1]), dtype)
    pos = Ap[:-1].copy()
    for (o, _, _, _), m in zip(offs, masks):
        Ai[pos[m]] = (ind[m] + o).astype(dtype)
        pos[m] += 1
    return N, Ap.astype(np.int32) if Ap[-1] < 2 ** 31 else Ap, Ai


KUHN_DIRS = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
GRID_DIRS = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
