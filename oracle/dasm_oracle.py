"""CPU oracle (numpy) for the dealii-asm smoother hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`dealii-asm_b200/`, `include/`,
`drivers/`) may import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do.

It restates, for structured quad/hex meshes with an arbitrary smooth map, the algorithm of
the reference (file:line into /root/reference):

* 1-D FE_Q(k) Gauss-Lobatto Lagrange basis at QGauss(k+1) points           [deal.II FE_Q/QGauss]
* weak Laplacian cell integral  evaluate(grad) -> J^-1 J^-T detJ w -> integrate(grad)
                                                      include/operator.h:866-875, 1335-1351
* "merged" coefficients  JxW * J^-1 J^-T                 include/operator.h:674-711
* harmonic cell / patch extents                         include/grid_tools.h:11-138
* element-centred patch DoF lists with overlap          include/dof_tools.h:9-137
* 1-D patch mass/stiffness with overlap and BC rows     [deal.II TensorProductMatrixCreator::
                                                          create_laplace_tensor_product_matrix],
                                                          call site include/matrix_free.h:350-363
* FDM apply_inverse (S x S x S) L^-1 (S x S x S)^T      [deal.II TensorProductMatrixSymmetricSum],
                                                          call site include/matrix_free.h:1046-1052
* weights 1/valence, 1/sqrt(valence), RAS              include/matrix_free.h:536-712
* additive Schwarz application with pre/post/symm/none   include/matrix_free.h:1007-1260
* Jacobi (inverse diagonal, |d|<=1e-10 -> 1)             include/operator.h:1512-1524
* PreconditionChebyshev vmult/step/estimate_eigenvalues  [deal.II], configured at
                                                          include/precondition.templates.h:89-158
* SolverCG / SolverGMRES iteration counts                element_centered_preconditioners_01.cc:108-263

The deal.II pieces are NOT in /root/reference (un-vendored dependency, no version pin: the
README clones master, API use places it at ~9.5-pre); they are restated from the published
algorithm and pinned by the reference's own golden outputs
(tests/element_centered_preconitioners/small/*.output, indices_overlap_01.output,
subdivided_hyper_cube_balanced_01.output, tridiagonal_01.output) in tests/test_oracle_golden.py.

Array conventions: cell-local tensors are stored [cell, (z,) y, x] (x fastest =
deal.II "lexicographic"); vectors are flat float64 arrays indexed by global DoF number.
"""
from __future__ import annotations

import numpy as np

INVALID = np.uint32(0xFFFFFFFF)


# --------------------------------------------------------------------------------------
# 1-D basis
# --------------------------------------------------------------------------------------
def gauss_points(n):
    """QGauss(n) on [0,1] (points ascending, weights)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_points(n):
    """QGaussLobatto(n) points on [0,1] = support points of FE_Q(n-1)."""
    if n == 1:
        return np.array([0.5])
    if n == 2:
        return np.array([0.0, 1.0])
    # interior points: roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    r = np.sort(np.polynomial.legendre.legroots(dc).real)
    # Newton polish
    for _ in range(3):
        f = np.polynomial.legendre.legval(r, dc)
        df = np.polynomial.legendre.legval(r, np.polynomial.legendre.legder(dc))
        r = r - f / df
    x = np.concatenate([[-1.0], r, [1.0]])
    x = 0.5 * (x - x[::-1])  # symmetrise
    return 0.5 * (x + 1.0)


def lagrange(nodes, x):
    """values V[q,i]=l_i(x_q) and derivatives D[q,i]=l_i'(x_q) of the Lagrange basis on nodes."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(nodes)
    V = np.ones((len(x), n))
    D = np.zeros((len(x), n))
    for i in range(n):
        denom = 1.0
        for j in range(n):
            if j != i:
                denom *= nodes[i] - nodes[j]
        for q in range(len(x)):
            v = 1.0
            for j in range(n):
                if j != i:
                    v *= x[q] - nodes[j]
            V[q, i] = v / denom
            s = 0.0
            for m in range(n):
                if m == i:
                    continue
                p = 1.0
                for j in range(n):
                    if j != i and j != m:
                        p *= x[q] - nodes[j]
                s += p
            D[q, i] = s / denom
    return V, D


class Basis1D:
    """FE_Q(k) (Gauss-Lobatto nodes) evaluated at QGauss(k+1)."""

    def __init__(self, k):
        self.k = k
        self.n = k + 1
        self.nodes = gauss_lobatto_points(k + 1)
        self.qp, self.qw = gauss_points(k + 1)
        self.N, self.D = lagrange(self.nodes, self.qp)  # [q,i]
        # collocation derivative in the Gauss-point Lagrange basis  Dq[q,p] = lg_p'(x_q)
        _, self.Dq = lagrange(self.qp, self.qp)

    def reference_mass_stiffness(self):
        """[deal.II internal::create_reference_mass_and_stiffness_matrices] on the unit cell."""
        M = self.N.T @ (self.qw[:, None] * self.N)
        K = self.D.T @ (self.qw[:, None] * self.D)
        return M, K


# --------------------------------------------------------------------------------------
# mesh-size knob  (include/grid_generator.h:107-156)
# --------------------------------------------------------------------------------------
def decompose_for_subdivided_hyper_cube_balanced(dim, s):
    """returns (n_refine, subdivisions[dim]); reference include/grid_generator.h:107-135."""
    n_refine = s // 6
    remainder = s % 6
    sub = [1] * dim
    if remainder == 1 and s > 1:
        sub[0] = 3
        sub[1] = 2
        if dim > 2:
            sub[2] = 2
        n_refine -= 1
    if remainder == 2:
        sub[0] = 2
    elif remainder == 3:
        sub[0] = 3
    elif remainder == 4:
        sub[0] = 2
        sub[1] = 2
    elif remainder == 5:
        sub[0] = 3
        sub[1] = 2
    return n_refine, sub


# --------------------------------------------------------------------------------------
# structured mesh
# --------------------------------------------------------------------------------------
class StructuredMesh:
    """Tensor-product topology quad/hex mesh of the unit box [0,L]^dim with an optional smooth map.

    n_cells[d]  cells per direction, periodic[d] per direction; non-periodic boundaries are
    homogeneous Dirichlet when dirichlet=True (boundary id 1 in the reference driver,
    element_centered_preconditioners_01.cc:410-413) else natural (Neumann).
    `mapfun(X)` maps reference-box coordinates [..., dim] -> physical; the geometry used by all
    integrals is its degree-`mapping_degree` Lagrange interpolant per cell (MappingQCache,
    matrix_free_loop_08.likwid.cc:180-199).
    cell_order: processing order (list of lexicographic cell ids); numbering is first-touch in
    that order.
    """

    def __init__(self, dim, n_cells, periodic, dirichlet=True, mapfun=None, mapping_degree=2,
                 lengths=None, cell_order=None):
        self.dim = dim
        self.n_cells = tuple(int(c) for c in n_cells)
        self.periodic = tuple(bool(p) for p in periodic)
        self.dirichlet = bool(dirichlet)
        self.mapfun = mapfun
        self.mapping_degree = mapping_degree
        self.lengths = tuple(lengths) if lengths is not None else (1.0,) * dim
        self.C = int(np.prod(self.n_cells))
        if cell_order is None:
            cell_order = np.arange(self.C)
        self.cell_order = np.asarray(cell_order, dtype=np.int64)

    def cell_ijk(self, lex):
        out = []
        for d in range(self.dim):
            out.append(lex % self.n_cells[d])
            lex = lex // self.n_cells[d]
        return tuple(out)

    def cell_lex(self, ijk):
        lex = 0
        for d in reversed(range(self.dim)):
            lex = lex * self.n_cells[d] + ijk[d]
        return lex

    def neighbor(self, ijk, d, side):
        """lexicographic ijk of the face neighbour or None (side 0: low, 1: high)."""
        v = list(ijk)
        v[d] += -1 if side == 0 else 1
        if v[d] < 0 or v[d] >= self.n_cells[d]:
            if not self.periodic[d]:
                return None
            v[d] %= self.n_cells[d]
        return tuple(v)

    def cell_points(self, ijk, ref):
        """physical points of reference points ref[..., dim] in cell ijk using the Q_m interpolant."""
        m = self.mapping_degree
        nodes = gauss_lobatto_points(m + 1)
        h = [self.lengths[d] / self.n_cells[d] for d in range(self.dim)]
        # support points
        grids = np.meshgrid(*[(ijk[d] + nodes) * h[d] for d in reversed(range(self.dim))], indexing="ij")
        X = np.stack([g for g in reversed(grids)], axis=-1)  # [(z,)y,x,dim]
        if self.mapfun is not None:
            X = self.mapfun(X)
        ref = np.asarray(ref)
        # interpolate
        out = np.zeros(ref.shape[:-1] + (self.dim,))
        Vs = [lagrange(nodes, ref[..., d].reshape(-1))[0] for d in range(self.dim)]  # [P, m+1]
        P = Vs[0].shape[0]
        if self.dim == 2:
            val = np.einsum("pj,pi,jid->pd", Vs[1], Vs[0], X)
        else:
            val = np.einsum("pk,pj,pi,kjid->pd", Vs[2], Vs[1], Vs[0], X)
        return val.reshape(out.shape)

    def jacobians(self, basis):
        """J[c, q(lex), d, e] = d x_d / d xi_e at the QGauss(n)^dim points, via the Q_m interpolant."""
        m = self.mapping_degree
        nodes = gauss_lobatto_points(m + 1)
        V, D = lagrange(nodes, basis.qp)  # [q, m+1]
        nq = basis.n
        dim = self.dim
        h = [self.lengths[d] / self.n_cells[d] for d in range(dim)]
        J = np.zeros((self.C, nq ** dim, dim, dim))
        for c in range(self.C):
            ijk = self.cell_ijk(c)
            grids = np.meshgrid(*[(ijk[d] + nodes) * h[d] for d in reversed(range(dim))], indexing="ij")
            X = np.stack([g for g in reversed(grids)], axis=-1)
            if self.mapfun is not None:
                X = self.mapfun(X)
            for e in range(dim):
                mats = [D if d == e else V for d in range(dim)]
                if dim == 2:
                    g = np.einsum("bj,ai,jid->bad", mats[1], mats[0], X)
                else:
                    g = np.einsum("ck,bj,ai,kjid->cbad", mats[2], mats[1], mats[0], X)
                J[c, :, :, e] = g.reshape(nq ** dim, dim)
        return J


def sine_map(X):
    """matrix_free_loop_08.likwid.cc:185-199 (use cartesian mesh = false)."""
    dim = X.shape[-1]
    out = X.copy()
    for d in range(dim):
        out[..., d] = X[..., d] + np.sin(2 * np.pi * X[..., (d + 1) % dim]) * np.sin(np.pi * X[..., d]) * 0.1
    return out


def kershaw_map(epsy, epsz):
    """include/kershaw.h:4-80."""

    def right(eps, x):
        return np.where(x <= 0.5, (2 - eps) * x, 1 + eps * (x - 1))

    def left(eps, x):
        return 1 - right(eps, 1 - x)

    def step(x):
        return np.where(x <= 0, 0.0, np.where(x >= 1, 1.0, ((6 * x - 15) * x + 10) * x * x * x))

    def f(X):
        x = X[..., 0]
        y = X[..., 1]
        z = X[..., 2] if X.shape[-1] > 2 else 0 * x
        layer = (x * 6.0).astype(np.int64)
        layer = np.minimum(layer, 5)  # x == 1
        lam = (x - layer / 6.0) * 6
        Yo = np.zeros_like(x)
        Zo = np.zeros_like(x)
        for lay in range(6):
            sel = layer == lay
            if not np.any(sel):
                continue
            if lay == 0:
                Y = left(epsy, y)
                Z = left(epsz, z)
            elif lay in (1, 4):
                s = step(lam)
                Y = (1 - s) * left(epsy, y) + s * right(epsy, y)
                Z = (1 - s) * left(epsz, z) + s * right(epsz, z)
            elif lay == 2:
                s = step(lam / 2)
                Y = (1 - s) * right(epsy, y) + s * left(epsy, y)
                Z = (1 - s) * right(epsz, z) + s * left(epsz, z)
            elif lay == 3:
                s = step((1 + lam) / 2)
                Y = (1 - s) * right(epsy, y) + s * left(epsy, y)
                Z = (1 - s) * right(epsz, z) + s * left(epsz, z)
            else:
                Y = right(epsy, y)
                Z = right(epsz, z)
            Yo = np.where(sel, Y, Yo)
            Zo = np.where(sel, Z, Zo)
        out = X.copy()
        out[..., 1] = Yo
        if X.shape[-1] > 2:
            out[..., 2] = Zo
        return out

    return f


# --------------------------------------------------------------------------------------
# DoF numbering
# --------------------------------------------------------------------------------------
def _entity_of(idx, k):
    return 0 if idx == 0 else (2 if idx == k else 1)


def brick_major_order(n_cells, brick=(4, 4, 4)):
    """processing order of the B200 library: bricks of `brick` cells in lexicographic brick order
    (x fastest), cells lexicographic inside a brick.  Returns lexicographic cell ids."""
    dim = len(n_cells)
    nc = list(n_cells) + [1] * (3 - dim)
    B = list(brick[:dim]) + [1] * (3 - dim)
    order = []
    for bz in range(0, nc[2], B[2]):
        for by in range(0, nc[1], B[1]):
            for bx in range(0, nc[0], B[0]):
                for z in range(bz, min(bz + B[2], nc[2])):
                    for y in range(by, min(by + B[1], nc[1])):
                        for x in range(bx, min(bx + B[0], nc[0])):
                            order.append((z * nc[1] + y) * nc[0] + x)
    return np.array(order, dtype=np.int64)


LEX_FLAG = 0x80000000  # start index of an entity stored in a "lex" brick (first DoF of the entity, box strides)


def lex_bricks(mesh, brick=(4, 4, 4)):
    """keys (reversed brick coordinates) of the bricks numbered lexicographically: full 4^dim bricks with a
    neighbour cell across every face (interior or periodic), hence without constrained DoFs."""
    dim = mesh.dim
    nc = mesh.n_cells
    out = set()
    if any(b != 4 for b in brick[:dim]):
        return out
    for key in np.ndindex(*[(nc[d] + 3) // 4 for d in reversed(range(dim))]):
        org = [key[dim - 1 - d] * 4 for d in range(dim)]
        ok = True
        for d in range(dim):
            if org[d] + 4 > nc[d]:
                ok = False
            elif not mesh.periodic[d] and (org[d] == 0 or org[d] + 4 == nc[d]):
                ok = False
        if ok:
            out.add(tuple(key))
    return out


def number_dofs_owner_cell(mesh, k, brick=(4, 4, 4), lex=True):
    """Native "brick-grouped owner-cell" numbering of the B200 library, restated independently.

    Every mesh entity (3^dim per cell: vertex / line / quad / hex interior in the lexicographic
    3x3(x3) layout of vector_access_reduced.h:30-164) is owned by the cell for which it is a lower
    entity (code 0 or 1 per direction); code 2 is owned only by the last cell of a non-periodic
    direction.  Cells are grouped in bricks of `brick` cells (brick-major order = mesh.cell_order).

    "Lex" bricks (lex_bricks above) come first: brick i owns the box [0, 4k)^dim of its tile, numbered
    lexicographically (x fastest) from i (4k)^dim; the start index of an entity of such a brick is the index of
    its first DoF with LEX_FLAG set, and the DoFs of the entity are expanded with the box strides 1, 4k, (4k)^2.
    Then the other bricks: per brick, first the owned entities that do not lie on a lower brick face with a
    neighbour cell across it ("private"), cell by cell in lexicographic entity order, then the shared
    ones in the same order; DoFs are lexicographic and contiguous inside an entity.  Entities on a Dirichlet
    boundary are numbered too (they exist in the vector, like constrained DoFs in deal.II) and flagged
    constrained.

    returns cell_dofs [C, n^dim] (uint32, by lexicographic cell id), n_dofs, constrained [n_dofs]
            bool, compressed [C, 3^dim] entity start indices (plain: constrained keep their index)
    """
    dim = mesh.dim
    n = k + 1
    nc = mesh.n_cells
    km1 = k - 1
    size = [2 * nc[d] + (0 if mesh.periodic[d] else 1) for d in range(dim)]
    ent_start = -np.ones(size[::-1], dtype=np.int64)  # [(z,)y,x] slot lattice
    next_dof = 0
    constrained_ranges = []

    def ent_size(ee):
        cnt = 1
        for d in range(dim):
            if ee[d] == 1:
                cnt *= km1
        return cnt

    # group the processing order into bricks (cells of a brick are consecutive in mesh.cell_order)
    bricks = {}
    for c in mesh.cell_order:
        ijk = mesh.cell_ijk(int(c))
        key = tuple(ijk[d] // brick[d] for d in reversed(range(dim)))
        bricks.setdefault(key, []).append(int(c))
    lexset = lex_bricks(mesh, brick) if lex else set()
    box = (4 * k) ** dim
    for key in sorted(bricks.keys()):
        if key not in lexset:
            continue
        org = [key[dim - 1 - d] * 4 for d in range(dim)]
        # slots 2 org .. 2 org + 7 per direction: even slot 2c -> box coordinate c k, odd slot 2c+1 -> c k + 1
        for loc in np.ndindex(*([8] * dim)):  # loc[0] is the slowest direction
            first = 0
            slot = []
            for d in reversed(range(dim)):
                sl = loc[dim - 1 - d]
                first = first * (4 * k) + (sl // 2) * k + (sl % 2)
                slot.append((2 * org[d] + sl))
            ent_start[tuple(slot)] = (next_dof + first) | LEX_FLAG
        next_dof += box
    for key in sorted(bricks.keys()):
        if key in lexset:
            continue
        cells = bricks[key]
        org = [key[dim - 1 - d] * brick[d] for d in range(dim)]
        for pass_ in (0, 1):
            for c in cells:
                ijk = mesh.cell_ijk(c)
                for e in range(3 ** dim):
                    ee = [(e // 3 ** d) % 3 for d in range(dim)]
                    owned = all(ee[d] < 2 or (not mesh.periodic[d] and ijk[d] == nc[d] - 1) for d in range(dim))
                    if not owned:
                        continue
                    shared = any(ee[d] == 0 and ijk[d] == org[d] and mesh.neighbor(ijk, d, 0) is not None for d in range(dim))
                    if int(shared) != pass_:
                        continue
                    slot = [2 * ijk[d] + ee[d] for d in range(dim)]
                    cnt = ent_size(ee)
                    ent_start[tuple(slot[::-1])] = next_dof
                    if mesh.dirichlet and cnt > 0 and any(
                            (not mesh.periodic[d]) and (slot[d] == 0 or slot[d] == size[d] - 1) for d in range(dim)):
                        constrained_ranges.append((next_dof, next_dof + cnt))
                    next_dof += cnt
    assert np.all(ent_start >= 0)
    compressed = np.zeros((mesh.C, 3 ** dim), dtype=np.uint32)
    for c in range(mesh.C):
        ijk = mesh.cell_ijk(c)
        for e in range(3 ** dim):
            ee = [(e // 3 ** d) % 3 for d in range(dim)]
            slot = [(2 * ijk[d] + ee[d]) % size[d] for d in range(dim)]
            compressed[c, e] = ent_start[tuple(slot[::-1])]
    cell_dofs = expand_compressed(compressed, k, dim)
    constrained = np.zeros(next_dof, dtype=bool)
    for a0, b0 in constrained_ranges:
        constrained[a0:b0] = True
    return cell_dofs, next_dof, constrained, compressed


def expand_compressed(compressed, k, dim):
    """27 (9) entity start indices -> n^dim lexicographic indices per cell; the standard-orientation
    case of vector_access_reduced.h:267-405 (vertex idx, line idx+i, quad idx+j*(k-1)+i, ...); entities of lex
    bricks (LEX_FLAG) use the strides 1, 4k, (4k)^2 of the brick's box instead."""
    n = k + 1
    C = compressed.shape[0]
    out = np.zeros((C,) + (n,) * dim, dtype=np.uint32)
    km1 = k - 1
    for e in range(3 ** dim):
        ee = [(e // 3 ** d) % 3 for d in range(dim)]
        sl = []
        shape = []
        for d in range(dim):
            if ee[d] == 0:
                sl.append(slice(0, 1))
                shape.append(1)
            elif ee[d] == 2:
                sl.append(slice(k, k + 1))
                shape.append(1)
            else:
                sl.append(slice(1, k))
                shape.append(km1)
        cnt = int(np.prod(shape))
        if cnt == 0:
            continue
        off = np.arange(cnt, dtype=np.int64).reshape(shape[::-1])
        # offsets with the box strides: sum_d i_d (4k)^d
        grids = np.meshgrid(*[np.arange(shape[d], dtype=np.int64) for d in reversed(range(dim))], indexing="ij")
        off_lex = np.zeros(shape[::-1], dtype=np.int64)
        for j, d in enumerate(reversed(range(dim))):
            off_lex += grids[j] * (4 * k) ** d
        st = compressed[:, e].astype(np.int64).reshape((C,) + (1,) * dim)
        invalid = st == INVALID
        is_lex = (~invalid) & ((st & LEX_FLAG) != 0)
        val = np.where(is_lex, (st & (LEX_FLAG - 1)) + off_lex[None], st + off[None])
        val = np.where(invalid, np.int64(INVALID), val)
        out[(slice(None),) + tuple(sl[::-1])] = val.astype(np.uint32)
    return out.reshape(C, n ** dim)


def dealii_numbering_2d(n_refine, k):
    """deal.II default DoFHandler::distribute_dofs numbering on GridGenerator::hyper_cube +
    refine_global(n_refine), 2-D: active cells in Morton (z-)order; per cell new vertex DoFs
    (vertices 0..3), then new line DoFs (lines 0:left 1:right 2:bottom 3:top), then the quad.
    Pinned by /root/reference/indices_overlap_01.output.  Returns (cell_dofs [C,n^2] by
    lexicographic cell id, n_dofs, morton_order, boundary mask)."""
    nc = 2 ** n_refine
    n = k + 1
    C = nc * nc

    def morton(idx):
        x = y = 0
        for b in range(n_refine):
            q = (idx >> (2 * (n_refine - 1 - b))) & 3
            x = (x << 1) | (q & 1)
            y = (y << 1) | (q >> 1)
        return x, y

    vert = {}
    linex = {}  # horizontal lines keyed (i, jv)
    liney = {}  # vertical lines keyed (iv, j)
    cell_dofs = np.zeros((C, n * n), dtype=np.uint32)
    nxt = 0
    order = []
    for m in range(C):
        i, j = morton(m)
        lex = j * nc + i
        order.append(lex)
        loc = np.zeros((n, n), dtype=np.int64)
        for v, (a, b) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
            key = (i + a, j + b)
            if key not in vert:
                vert[key] = nxt
                nxt += 1
            loc[b * k, a * k] = vert[key]
        for l in range(4):
            if l < 2:  # vertical line at x = i + l
                key = (i + l, j)
                if key not in liney:
                    liney[key] = nxt
                    nxt += k - 1
                for t in range(k - 1):
                    loc[1 + t, l * k] = liney[key] + t
            else:  # horizontal line at y = j + (l-2)
                key = (i, j + l - 2)
                if key not in linex:
                    linex[key] = nxt
                    nxt += k - 1
                for t in range(k - 1):
                    loc[(l - 2) * k, 1 + t] = linex[key] + t
        for b in range(k - 1):
            for a in range(k - 1):
                loc[1 + b, 1 + a] = nxt
                nxt += 1
        cell_dofs[lex] = loc.reshape(-1)
    boundary = np.zeros(nxt, dtype=bool)
    for c in range(C):
        i, j = c % nc, c // nc
        loc = cell_dofs[c].reshape(n, n)
        if i == 0:
            boundary[loc[:, 0]] = True
        if i == nc - 1:
            boundary[loc[:, k]] = True
        if j == 0:
            boundary[loc[0, :]] = True
        if j == nc - 1:
            boundary[loc[k, :]] = True
    return cell_dofs, nxt, np.array(order), boundary


# --------------------------------------------------------------------------------------
# Laplace operator
# --------------------------------------------------------------------------------------
def _apply_1d(M, u, axis):
    """contract matrix M[q,i] with tensor u along `axis` (axis counted in [cell,(z,)y,x] layout)."""
    out = np.tensordot(u, M, axes=([axis], [1]))  # contracted axis goes last
    return np.moveaxis(out, -1, axis)


def merged_coefficients(J, basis, dim):
    """G[c,q,d,e] = JxW * (J^-1 J^-T)   (include/operator.h:674-711; same quantity deal.II
    applies in submit_gradient(get_gradient), operator.h:866-875)."""
    w1 = basis.qw
    if dim == 2:
        w = np.einsum("b,a->ba", w1, w1).reshape(-1)
    else:
        w = np.einsum("c,b,a->cba", w1, w1, w1).reshape(-1)
    Jinv = np.linalg.inv(J)
    det = np.linalg.det(J)
    G = np.einsum("cqde,cqfe->cqdf", Jinv, Jinv) * (det * w[None, :])[..., None, None]
    return G


def construct_q_coefficients(mesh, basis):
    """mapping type "construct q" (include/operator.h:712-746, 1221-1333): the cell stores the coordinates of its quadrature
    points; the Jacobian at a point is the collocation derivative (Lagrange basis through the Gauss points) of that coordinate
    field, then G = JxW J^-1 J^-T as in merged_coefficients.  Returns G[c, q, d, e]."""
    dim, n = mesh.dim, basis.n
    qp = basis.qp
    grids = np.meshgrid(*[qp] * dim, indexing="ij")               # [(z,) y, x]
    ref = np.stack([g for g in reversed(grids)], axis=-1)         # [..., (x, y(, z))]
    J = np.zeros((mesh.C, n ** dim, dim, dim))
    for c in range(mesh.C):
        X = mesh.cell_points(mesh.cell_ijk(c), ref)               # [(z,) y, x, dim]
        for e in range(dim):                                       # derivative along direction e = contraction of axis dim-1-e
            dX = np.moveaxis(np.tensordot(X, basis.Dq, axes=([dim - 1 - e], [1])), -1, dim - 1 - e)
            J[c, :, :, e] = dX.reshape(n ** dim, dim)
    return merged_coefficients(J, basis, dim)


class LaplaceOperator:
    """Matrix-free weak Laplacian, the restatement of LaplaceOperatorMatrixFree::vmult
    (include/operator.h:1353-1430).  Constrained (homogeneous Dirichlet) DoFs are read as zero and
    not written; with `copy_constrained` (the pre/post-hook variant, matrix_free_internal.h:226-255)
    dst[i] = src[i] on them, otherwise they stay zero."""

    def __init__(self, dim, k, cell_dofs, n_dofs, constrained, G, dtype=np.float64):
        self.dim = dim
        self.k = k
        self.n = k + 1
        self.basis = Basis1D(k)
        self.cell_dofs = cell_dofs.astype(np.int64)
        self.n_dofs = n_dofs
        self.constrained = constrained
        self.dtype = dtype
        self.G = G.astype(dtype)
        self.N = self.basis.N.astype(dtype)
        self.Dq = self.basis.Dq.astype(dtype)
        valid = self.cell_dofs != int(INVALID)
        self.mask = (valid & ~constrained[np.where(valid, self.cell_dofs, 0)]).astype(dtype)
        self.idx = np.where(valid, self.cell_dofs, 0)

    def cell_apply(self, ul):
        """ul [C, n^dim] -> local result [C, n^dim]."""
        dim, n = self.dim, self.n
        C = ul.shape[0]
        u = ul.reshape((C,) + (n,) * dim)
        # interpolate to Gauss points (collocation), axis numbering: x is last
        for d in range(dim):
            u = _apply_1d(self.N, u, u.ndim - 1 - d)
        grads = [_apply_1d(self.Dq, u, u.ndim - 1 - d).reshape(C, -1) for d in range(dim)]
        g = np.stack(grads, axis=-1)  # [C,q,dim]
        f = np.einsum("cqde,cqe->cqd", self.G, g)
        r = np.zeros_like(u)
        for d in range(dim):
            fd = f[..., d].reshape(u.shape)
            r = r + _apply_1d(self.Dq.T, fd, u.ndim - 1 - d)
        for d in range(dim):
            r = _apply_1d(self.N.T, r, r.ndim - 1 - d)
        return r.reshape(C, -1)

    def vmult(self, x, copy_constrained=False):
        x = np.asarray(x, dtype=self.dtype)
        ul = x[self.idx] * self.mask
        rl = self.cell_apply(ul) * self.mask
        y = np.zeros(self.n_dofs, dtype=self.dtype)
        np.add.at(y, self.idx.reshape(-1), rl.reshape(-1))
        if copy_constrained:
            y[self.constrained] = x[self.constrained]
        return y

    def diagonal(self):
        """MatrixFreeTools::compute_diagonal restated: unit-vector applications per cell."""
        C = self.cell_dofs.shape[0]
        nloc = self.n ** self.dim
        d = np.zeros(self.n_dofs, dtype=np.float64)
        for i in range(nloc):
            e = np.zeros((C, nloc), dtype=self.dtype)
            e[:, i] = 1.0
            r = self.cell_apply(e * self.mask)[:, i] * self.mask[:, i]
            np.add.at(d, self.idx[:, i], r)
        return d

    def inverse_diagonal(self):
        """include/operator.h:1512-1524: 1/d, entries with |d| <= 1e-10 -> 1 ; constrained -> 1."""
        d = self.diagonal()
        d[self.constrained] = 1.0
        out = np.where(np.abs(d) > 1e-10, 1.0 / np.where(np.abs(d) > 1e-10, d, 1.0), 1.0)
        return out.astype(self.dtype)

    def dense(self):
        A = np.zeros((self.n_dofs, self.n_dofs))
        for i in range(self.n_dofs):
            e = np.zeros(self.n_dofs)
            e[i] = 1
            A[:, i] = self.vmult(e)
        return A


# --------------------------------------------------------------------------------------
# harmonic extents   (include/grid_tools.h:11-138)
# --------------------------------------------------------------------------------------
def harmonic_cell_extents(mesh, basis):
    dim = mesh.dim
    qp, qw = basis.qp, basis.qw
    ext = np.zeros((mesh.C, dim))
    if dim == 2:
        tq = qp[:, None]
        tw = qw
    else:
        a, b = np.meshgrid(qp, qp, indexing="xy")
        tq = np.stack([a.reshape(-1), b.reshape(-1)], axis=-1)
        tw = np.einsum("b,a->ba", qw, qw).reshape(-1)
    for c in range(mesh.C):
        ijk = mesh.cell_ijk(c)
        for d in range(dim):
            p0 = np.zeros((len(tw), dim))
            p1 = np.zeros((len(tw), dim))
            others = [e for e in range(dim) if e != d]
            for t, e in enumerate(others):
                p0[:, e] = tq[:, t]
                p1[:, e] = tq[:, t]
            p0[:, d] = 0.0
            p1[:, d] = 1.0
            x0 = mesh.cell_points(ijk, p0)
            x1 = mesh.cell_points(ijk, p1)
            ext[c, d] = np.sum(np.linalg.norm(x0 - x1, axis=-1) * tw)
    return ext


def harmonic_patch_extents(mesh, basis):
    """result[c,d,:] = (left neighbour extent or 0, own, right neighbour extent or 0)."""
    if hasattr(mesh, "harmonic_patch_extents"):
        return mesh.harmonic_patch_extents(basis)   # unstructured meshes: through the faces (UnstructuredMesh below)
    ext = harmonic_cell_extents(mesh, basis)
    out = np.zeros((mesh.C, mesh.dim, 3))
    for c in range(mesh.C):
        ijk = mesh.cell_ijk(c)
        for d in range(mesh.dim):
            out[c, d, 1] = ext[c, d]
            for side in (0, 1):
                nb = mesh.neighbor(ijk, d, side)
                if nb is not None:
                    out[c, d, 2 * side] = ext[mesh.cell_lex(nb), d]
    return out


# --------------------------------------------------------------------------------------
# FDM patch matrices + eigen-decomposition
# --------------------------------------------------------------------------------------
INTERNAL, DIRICHLET, NEUMANN = 0, 1, 2


def laplace_tensor_product_matrix_1d(M_ref, K_ref, extent, btype, n_overlap):
    """[deal.II TensorProductMatrixCreator::create_laplace_tensor_product_matrix], one direction.
    extent = (h_left, h, h_right), btype = (left, right) in {INTERNAL, DIRICHLET, NEUMANN}."""
    n = M_ref.shape[0]
    m = n - 2 + 2 * n_overlap
    M = np.zeros((m, m))
    K = np.zeros((m, m))
    o = n_overlap - 1
    M[o:o + n, o:o + n] = M_ref * extent[1]
    K[o:o + n, o:o + n] = K_ref / extent[1]
    if btype[0] == INTERNAL:
        assert extent[0] > 0
        M[:n_overlap, :n_overlap] += M_ref[n - n_overlap:, n - n_overlap:] * extent[0]
        K[:n_overlap, :n_overlap] += K_ref[n - n_overlap:, n - n_overlap:] / extent[0]
    elif btype[0] == DIRICHLET:
        i0 = n_overlap - 1
        M[i0, :] = 0
        M[:, i0] = 0
        K[i0, :] = 0
        K[:, i0] = 0
    if btype[1] == INTERNAL:
        assert extent[2] > 0
        s = n_overlap + n - 2
        M[s:s + n_overlap, s:s + n_overlap] += M_ref[:n_overlap, :n_overlap] * extent[2]
        K[s:s + n_overlap, s:s + n_overlap] += K_ref[:n_overlap, :n_overlap] / extent[2]
    elif btype[1] == DIRICHLET:
        i0 = n_overlap + n - 2
        M[i0, :] = 0
        M[:, i0] = 0
        K[i0, :] = 0
        K[:, i0] = 0
    return M, K


def generalized_eig(M, K):
    """K s = lambda M s with S^T M S = I; rows/cols with zero mass diagonal (constrained) are
    excluded: eigenvalue 1, zero eigenvector row/column [deal.II
    TensorProductMatrixSymmetricSum spectral assembly, recollection].  Returns S[m,m] (columns =
    eigenvectors), lam[m]."""
    from scipy.linalg import eigh

    m = M.shape[0]
    free = np.array([M[i, i] != 0.0 for i in range(m)])
    idx = np.nonzero(free)[0]
    S = np.zeros((m, m))
    lam = np.ones(m)
    if len(idx):
        w, v = eigh(K[np.ix_(idx, idx)], M[np.ix_(idx, idx)])
        S[np.ix_(idx, idx)] = v
        lam[idx] = w
    return S, lam


def patch_dof_indices(mesh, k, cell_dofs, n_overlap):
    """element-centred patch index lists, include/dof_tools.h:9-137 (return_all = true; INVALID
    outside the domain).  patch size 1-D m = k - 1 + 2 n_overlap."""
    dim = mesh.dim
    n = k + 1
    m = k - 1 + 2 * n_overlap
    out = np.full((mesh.C, m ** dim), int(INVALID), dtype=np.int64)
    if n_overlap == 1:
        return cell_dofs.astype(np.int64).copy()

    def translate(i):
        if i < n_overlap - 1:
            return 0, k + 1 - n_overlap + i
        elif i < k + n_overlap:
            return 1, i - (n_overlap - 1)
        else:
            return 2, i - (n_overlap + k - 1)

    tr = [translate(i) for i in range(m)]
    for c in range(mesh.C):
        ijk = mesh.cell_ijk(c)
        loc = np.full((m,) * dim, int(INVALID), dtype=np.int64)
        for pidx in np.ndindex(*(m,) * dim):  # pidx = ((z,)y,x)
            p = pidx[::-1]  # (x,y,z)
            cur = ijk
            ok = True
            li = []
            for d in range(dim):
                which, l = tr[p[d]]
                li.append(l)
                if which != 1:
                    nb = mesh.neighbor(cur, d, 0 if which == 0 else 1)
                    if nb is None:
                        ok = False
                        break
                    cur = nb
            if not ok:
                continue
            lex = 0
            for d in reversed(range(dim)):
                lex = lex * n + li[d]
            loc[pidx] = cell_dofs[mesh.cell_lex(cur), lex]
        out[c] = loc.reshape(-1)
    return out


def vertex_patch_dof_indices(mesh, k, cell_dofs):
    """vertex-patch index lists, include/dof_tools.h:206-300 with the cell selection of
    collect_cells_for_vertex_patch (include/matrix_free.h:1490-1509): the patch of a cell is the
    (2k-1)^dim interior of the 2^dim cells above it; all INVALID if one of them does not exist."""
    dim = mesh.dim
    n = k + 1
    m = 2 * k - 1
    out = np.full((mesh.C, m ** dim), int(INVALID), dtype=np.int64)
    for c in range(mesh.C):
        ijk = mesh.cell_ijk(c)
        cells = []
        ok = True
        for q in range(2 ** dim):
            cur = ijk
            for d in range(dim):
                if (q >> d) & 1:
                    cur = mesh.neighbor(cur, d, 1) if cur is not None else None
                    if cur is None:
                        break
            if cur is None:
                ok = False
                break
            cells.append(mesh.cell_lex(cur))
        if not ok:
            continue
        loc = np.zeros((m,) * dim, dtype=np.int64)
        for pidx in np.ndindex(*(m,) * dim):
            p = pidx[::-1]
            q = 0
            lex = 0
            ls = []
            for d in range(dim):
                g = p[d] + 1
                half = 1 if g > k else 0
                q |= half << d
                ls.append(g - half * k)
            for d in reversed(range(dim)):
                lex = lex * n + ls[d]
            loc[pidx] = cell_dofs[cells[q], lex]
        out[c] = loc.reshape(-1)
    return out


def vertex_patch_matrix_1d(M_ref, K_ref, h0, h1):
    """include/tensor_product_matrix_creator.h:7-61: two cells of extents h0, h1 glued at the patch
    vertex, outermost nodes dropped."""
    n = M_ref.shape[0]
    m = 2 * (n - 1) - 1
    M = np.zeros((m, m))
    K = np.zeros((m, m))
    M[:n - 1, :n - 1] += M_ref[1:, 1:] * h0
    K[:n - 1, :n - 1] += K_ref[1:, 1:] / h0
    M[n - 2:, n - 2:] += M_ref[:n - 1, :n - 1] * h1
    K[n - 2:, n - 2:] += K_ref[:n - 1, :n - 1] / h1
    return M, K


class FDMPreconditioner:
    """ASPoissonPreconditioner restated (include/matrix_free.h:73-1364), element-centred patches.

    weight_type in {none, pre, post, symm, ras}.  Weights are 1/valence (pre/post), 1/sqrt(valence)
    (symm), RAS 0/1 ownership by smallest processing index of the cells whose closure contains the
    DoF (matrix_free.h:536-673).
    """

    def __init__(self, mesh, k, cell_dofs, n_dofs, constrained, n_overlap=1, weight_type="symm",
                 dtype=np.float64, cell_rank=None, element_centric=True):
        self.mesh = mesh
        self.k = k
        self.dim = mesh.dim
        self.n_overlap = n_overlap
        self.weight_type = weight_type
        self.dtype = dtype
        self.n_dofs = n_dofs
        self.constrained = constrained
        basis = Basis1D(k)
        M_ref, K_ref = basis.reference_mass_stiffness()
        self.element_centric = element_centric
        self.m = (k - 1 + 2 * n_overlap) if element_centric else (2 * k - 1)
        m, dim = self.m, self.dim
        ext = harmonic_patch_extents(mesh, basis)
        self.extents = ext
        idx = patch_dof_indices(mesh, k, cell_dofs, n_overlap) if element_centric else vertex_patch_dof_indices(mesh, k, cell_dofs)
        valid = idx != int(INVALID)
        valid &= ~constrained[np.where(valid, idx, 0)]
        self.idx = np.where(valid, idx, 0)
        self.mask = valid.astype(dtype)
        # 1-D eigen decompositions per cell and direction
        self.S = np.zeros((mesh.C, dim, m, m))
        self.lam = np.zeros((mesh.C, dim, m))
        cache = {}
        for c in range(mesh.C):
            ijk = mesh.cell_ijk(c)
            for d in range(dim):
                bt = []
                for side in (0, 1):
                    if mesh.neighbor(ijk, d, side) is not None:
                        bt.append(INTERNAL)
                    else:
                        bt.append(DIRICHLET if mesh.dirichlet else NEUMANN)
                if not element_centric:
                    h0 = ext[c, d, 1] if ext[c, d, 1] != 0 else 1.0   # collect_patch_extend, matrix_free.h:1511-1524
                    h1 = ext[c, d, 2] if ext[c, d, 2] != 0 else 1.0
                    key = ("v", h0, h1)
                    if key not in cache:
                        cache[key] = generalized_eig(*vertex_patch_matrix_1d(M_ref, K_ref, h0, h1))
                    self.S[c, d], self.lam[c, d] = cache[key]
                    continue
                key = (tuple(ext[c, d]), tuple(bt))
                if key not in cache:
                    M, K = laplace_tensor_product_matrix_1d(M_ref, K_ref, ext[c, d], bt, n_overlap)
                    cache[key] = generalized_eig(M, K)
                self.S[c, d], self.lam[c, d] = cache[key]
        self.n_instances_1d = len(cache)
        # weights
        val = np.zeros(n_dofs)
        np.add.at(val, self.idx.reshape(-1), self.mask.reshape(-1).astype(np.float64))
        self.valence = val
        if weight_type == "ras":
            rank = cell_rank if cell_rank is not None else np.argsort(mesh.cell_order)  # processing index
            owner = np.full(n_dofs, np.iinfo(np.int64).max, dtype=np.int64)
            core = cell_dofs.astype(np.int64)
            cv = (core != int(INVALID))
            for c in range(mesh.C):
                ii = core[c][cv[c]]
                owner[ii] = np.minimum(owner[ii], rank[c])
            # local 0/1 weights per patch entry (core entries owned by this cell)
            wl = np.zeros_like(self.mask, dtype=np.float64)
            pidx_core = self._core_positions()
            for c in range(mesh.C):
                sel = pidx_core
                wl[c, sel] = (owner[self.idx[c, sel]] == rank[c]).astype(np.float64)
            self.w_local = (wl * self.mask).astype(dtype)
            self.weights = None
        elif weight_type == "none":
            self.weights = None
            self.w_local = None
        else:
            with np.errstate(divide="ignore"):
                w = np.where(val == 0, 0.0, 1.0 / (np.sqrt(val) if weight_type == "symm" else val))
            self.weights = w.astype(dtype)
            self.w_local = self.weights[self.idx]
        self.Sd = self.S.astype(dtype)
        lam_sum = self._lam_sum()
        self.inv_lam = (1.0 / lam_sum).astype(dtype)

    def _core_positions(self):
        m, dim, k, no = self.m, self.dim, self.k, self.n_overlap
        core1 = [(no - 1 <= i) and (i < k + no) for i in range(m)]
        sel = []
        for p in np.ndindex(*(m,) * dim):
            sel.append(all(core1[i] for i in p))
        return np.array(sel)

    def _lam_sum(self):
        lam = self.lam
        if self.dim == 2:
            return (lam[:, 1, :, None] + lam[:, 0, None, :]).reshape(lam.shape[0], -1)
        return (lam[:, 2, :, None, None] + lam[:, 1, None, :, None] + lam[:, 0, None, None, :]).reshape(lam.shape[0], -1)

    def apply_inverse(self, rl):
        """rl [C, m^dim] -> z local; (S2 x S1 x S0) diag^-1 (S2 x S1 x S0)^T."""
        C, m, dim = rl.shape[0], self.m, self.dim
        S = self.Sd
        t = rl.reshape((C,) + (m,) * dim)
        if dim == 2:
            t = np.einsum("cia,cjb,cji->cba", S[:, 0], S[:, 1], t)
            t = t * self.inv_lam.reshape(t.shape)
            t = np.einsum("cia,cjb,cba->cji", S[:, 0], S[:, 1], t)
        else:
            t = np.einsum("cia,cjb,ckd,ckji->cdba", S[:, 0], S[:, 1], S[:, 2], t)
            t = t * self.inv_lam.reshape(t.shape)
            t = np.einsum("cia,cjb,ckd,cdba->ckji", S[:, 0], S[:, 1], S[:, 2], t)
        return t.reshape(C, -1)

    def vmult(self, r):
        r = np.asarray(r, dtype=self.dtype)
        rl = r[self.idx] * self.mask
        wt = self.weight_type
        if wt in ("pre", "symm"):
            rl = rl * self.w_local
        zl = self.apply_inverse(rl)
        if wt in ("post", "symm", "ras"):
            zl = zl * self.w_local
        zl = zl * self.mask
        z = np.zeros(self.n_dofs, dtype=self.dtype)
        np.add.at(z, self.idx.reshape(-1), zl.reshape(-1))
        return z

    def is_symmetric(self):
        return self.weight_type in ("none", "symm")


class JacobiPreconditioner:
    def __init__(self, op):
        self.inv_diag = op.inverse_diagonal()
        self.dtype = op.dtype

    def vmult(self, r):
        return (self.inv_diag * np.asarray(r, dtype=self.dtype)).astype(self.dtype)

    def is_symmetric(self):
        return True


# --------------------------------------------------------------------------------------
# Chebyshev smoother  [deal.II PreconditionChebyshev, recollection of ~9.5]
# --------------------------------------------------------------------------------------
class Chebyshev:
    def __init__(self, op, precon, degree=3, smoothing_range=20.0, polynomial_type="1st kind",
                 ev_algorithm=None, eig_cg_n_iterations=40):
        self.op = op
        self.P = precon
        self.degree = degree
        self.smoothing_range = smoothing_range
        self.poly = polynomial_type
        if ev_algorithm is None:
            ev_algorithm = "lanczos" if precon.is_symmetric() else "power iteration"
        self.ev_algorithm = ev_algorithm
        self.n_it = eig_cg_n_iterations
        self.dtype = op.dtype
        self.max_ev = None

    def initial_guess(self):
        n = self.op.n_dofs
        v = (np.arange(n) % 11).astype(self.dtype)
        v = v - v.mean(dtype=self.dtype)
        v[self.op.constrained] = 0
        return v.astype(self.dtype)

    def estimate_eigenvalues(self):
        v = self.initial_guess()
        A = lambda x: self.op.vmult(x)
        if self.ev_algorithm == "power iteration":
            lam = 0.0
            v = v / np.linalg.norm(v)
            for _ in range(self.n_it):
                w = self.P.vmult(A(v))
                lam = float(np.dot(v.astype(np.float64), w.astype(np.float64)))
                v = (w / np.linalg.norm(w)).astype(self.dtype)
            lam = abs(lam)
            self.min_ev = lam
            self.max_ev = 1.2 * lam
        else:
            evs = lanczos_cg_eigenvalues(A, self.P.vmult, v, self.n_it, 1e-10, self.dtype)
            self.min_ev = evs[0]
            self.max_ev = 1.2 * evs[-1]
        alpha = self.max_ev / self.smoothing_range if self.smoothing_range > 1 else min(0.9 * self.max_ev, self.min_ev)
        if self.poly == "4th kind":
            self.delta = self.max_ev
            self.theta = self.max_ev
        else:
            self.delta = (self.max_ev - alpha) * 0.5
            self.theta = (self.max_ev + alpha) * 0.5
        return self.min_ev, self.max_ev

    def set_eigenvalues(self, max_ev, min_ev=None):
        self.max_ev = max_ev
        self.min_ev = min_ev if min_ev is not None else max_ev / 1.2
        alpha = self.max_ev / self.smoothing_range if self.smoothing_range > 1 else min(0.9 * self.max_ev, self.min_ev)
        if self.poly == "4th kind":
            self.delta = self.theta = self.max_ev
        else:
            self.delta = (self.max_ev - alpha) * 0.5
            self.theta = (self.max_ev + alpha) * 0.5

    def coefficients(self):
        """list of (f1, f2) for iteration indices 0/1, 2, 3, ..."""
        out = [(0.0, (4.0 / (3.0 * self.theta)) if self.poly == "4th kind" else 1.0 / self.theta)]
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return out
        sigma = self.theta / self.delta
        rho_old = 1.0 / sigma
        for j in range(self.degree - 1):
            if self.poly == "4th kind":
                f1 = (2 * j + 1.0) / (2 * j + 5.0)
                f2 = (8 * j + 12.0) / (self.theta * (2 * j + 5.0))
            else:
                rho = 1.0 / (2.0 * sigma - rho_old)
                f1 = rho * rho_old
                f2 = 2.0 * rho / self.delta
                rho_old = rho
            out.append((f1, f2))
        return out

    def _run(self, x, b, first_is_step):
        if self.max_ev is None:
            self.estimate_eigenvalues()
        dt = self.dtype
        co = self.coefficients()
        b = np.asarray(b, dtype=dt)
        f2 = dt(co[0][1]) if dt != np.float64 else co[0][1]
        if first_is_step:
            t = b - self.op.vmult(x, copy_constrained=True)
            x_new = x + f2 * self.P.vmult(t)
        else:
            x_new = f2 * self.P.vmult(b)
        x_old, x = x, x_new.astype(dt)
        for (f1, f2) in co[1:]:
            t = b - self.op.vmult(x, copy_constrained=True)
            z = self.P.vmult(t)
            x_new = x + f1 * (x - x_old) + f2 * z
            x_old, x = x, x_new.astype(dt)
        return x

    def vmult(self, b):
        return self._run(np.zeros(self.op.n_dofs, dtype=self.dtype), b, False)

    def step(self, x, b):
        return self._run(np.asarray(x, dtype=self.dtype), b, True)


def lanczos_cg_eigenvalues(A, Pinv, b, n_it, tol, dtype=np.float64):
    """Ritz values of P^-1 A from preconditioned CG coefficients (deal.II SolverCG eigenvalue
    signal): tridiagonal T with diag 1/alpha_k + beta_{k-1}/alpha_{k-1}, offdiag sqrt(beta_k)/alpha_k."""
    x = np.zeros_like(b)
    r = b.copy()
    z = Pinv(r)
    p = z.copy()
    rz = float(np.dot(r.astype(np.float64), z.astype(np.float64)))
    alphas, betas = [], []
    r0 = np.linalg.norm(r)
    for it in range(n_it):
        Ap = A(p)
        pAp = float(np.dot(p.astype(np.float64), Ap.astype(np.float64)))
        if pAp == 0:
            break
        alpha = rz / pAp
        x = x + dtype(alpha) * p
        r = r - dtype(alpha) * Ap
        alphas.append(alpha)
        if np.linalg.norm(r) < tol * r0 or np.linalg.norm(r) < 1e-300:
            break
        z = Pinv(r)
        rz_new = float(np.dot(r.astype(np.float64), z.astype(np.float64)))
        beta = rz_new / rz
        betas.append(beta)
        rz = rz_new
        p = z + dtype(beta) * p
    kk = len(alphas)
    T = np.zeros((kk, kk))
    for i in range(kk):
        T[i, i] = 1.0 / alphas[i] + (betas[i - 1] / alphas[i - 1] if i > 0 else 0.0)
        if i + 1 < kk:
            T[i, i + 1] = T[i + 1, i] = np.sqrt(betas[i]) / alphas[i]
    return np.sort(np.linalg.eigvalsh(T))


# --------------------------------------------------------------------------------------
# Krylov solvers for iteration-count goldens
# --------------------------------------------------------------------------------------
def solve_cg(A, Pinv, b, rel_tol=1e-2, abs_tol=1e-10, max_it=1000):
    """deal.II SolverCG + ReductionControl: stop when ||r|| <= max(abs_tol, rel_tol*||r0||)."""
    x = np.zeros_like(b)
    r = b.copy()
    r0 = np.linalg.norm(r)
    target = max(abs_tol, rel_tol * r0)
    if r0 <= target:
        return x, 0
    z = Pinv(r)
    p = z.copy()
    rz = np.dot(r, z)
    for it in range(1, max_it + 1):
        Ap = A(p)
        alpha = rz / np.dot(p, Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        if np.linalg.norm(r) <= target:
            return x, it
        z = Pinv(r)
        rz_new = np.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, max_it


def solve_gmres(A, Pinv, b, rel_tol=1e-2, abs_tol=1e-10, max_it=1000, restart=30, right=True):
    """GMRES(30) with right preconditioning (element_centered_preconditioners_01.cc:150-176 sets
    right_preconditioning = true); iteration count = number of Arnoldi steps until the residual
    estimate drops below max(abs_tol, rel_tol * ||r0||)."""
    n = len(b)
    x = np.zeros(n)
    r = b - A(x)
    beta0 = np.linalg.norm(r)
    target = max(abs_tol, rel_tol * beta0)
    if beta0 <= target:
        return x, 0
    its = 0
    while its < max_it:
        r = b - A(x)
        beta = np.linalg.norm(r)
        V = [r / beta]
        Z = []
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        done = False
        kk = 0
        for j in range(restart):
            zj = Pinv(V[j])
            Z.append(zj)
            w = A(zj)
            for i in range(j + 1):
                H[i, j] = np.dot(w, V[i])
                w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j] if H[j + 1, j] > 0 else w)
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            d = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / d, H[j + 1, j] / d
            H[j, j] = d
            H[j + 1, j] = 0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            its += 1
            kk = j + 1
            if abs(g[j + 1]) <= target or its >= max_it:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:kk, :kk]), g[:kk])
        for i in range(kk):
            x = x + y[i] * Z[i]
        if done:
            break
    return x, its


def thomas_solve(a, b, c, d):
    """tridiagonal solve (include/preconditioners.h:420-526 TriDiagonalMatrixView); a: sub, b: diag, c: super."""
    n = len(b)
    cp = np.zeros(n)
    dp = np.zeros(n)
    cp[0] = c[0] / b[0]
    dp[0] = d[0] / b[0]
    for i in range(1, n):
        den = b[i] - a[i] * cp[i - 1]
        cp[i] = (c[i] / den) if i < n - 1 else 0.0
        dp[i] = (d[i] - a[i] * dp[i - 1]) / den
    x = np.zeros(n)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


# --------------------------------------------------------------------------------------
# Multigrid: two-level transfer and V-cycle
#   [deal.II MGTwoLevelTransfer / MGTransferGlobalCoarsening, Multigrid, PreconditionMG as set up by the reference in
#    include/multigrid.h:260-465 and element_centered_preconditioners_01.cc:540-740]
# --------------------------------------------------------------------------------------
def transfer_matrix_1d(k_fine, k_coarse, child=None):
    """Values of the coarse 1-D Lagrange basis (Gauss-Lobatto nodes, degree k_coarse) at the fine nodes:
    child = None: same cell, fine degree k_fine (polynomial transfer); child = 0 / 1: the fine cell is the lower / upper half
    of the coarse cell (geometric 2:1 transfer).  [n_fine, n_coarse]"""
    xc = gauss_lobatto_points(k_coarse + 1)
    xf = gauss_lobatto_points(k_fine + 1)
    if child is not None:
        xf = 0.5 * (xf + child)
    P = lagrange(xc, xf)[0]
    P[np.abs(P) < 1e-15] = 0.0
    return P


class TwoLevelTransfer:
    """Prolongation = embedding of the coarse finite element space into the fine one, cell by cell, weighted with the
    inverse valence of the fine DoFs (so that the contributions of all cells sharing a DoF sum to its interpolated value);
    restriction = its transpose.  Constrained (homogeneous Dirichlet) DoFs are read as zero and not written
    (MGTwoLevelTransfer::prolongate_and_add / restrict_and_add).
    Geometric transfer: mesh_f has twice the cells of mesh_c per direction, same degree; polynomial transfer: same mesh."""

    def __init__(self, mesh_f, op_f, mesh_c, op_c, parent=None):
        """parent (unstructured meshes): parent[fine cell] = coarse cell | child position << 28, None on the same mesh"""
        dim = op_f.dim
        self.dim, self.op_f, self.op_c = dim, op_f, op_c
        unstructured = not hasattr(mesh_f, "n_cells")
        self.geometric = (parent is not None) if unstructured else tuple(mesh_f.n_cells) != tuple(mesh_c.n_cells)
        if self.geometric:
            assert op_f.k == op_c.k and (unstructured or all(f == 2 * c for f, c in zip(mesh_f.n_cells, mesh_c.n_cells)))
            self.P1 = [transfer_matrix_1d(op_f.k, op_c.k, a) for a in (0, 1)]
        else:
            self.P1 = [transfer_matrix_1d(op_f.k, op_c.k)] * 2
        # parent cell of every fine cell and the child position per direction
        self.parent = np.zeros(mesh_f.C, dtype=np.int64)
        self.child = np.zeros((mesh_f.C, dim), dtype=np.int64)
        for c in range(mesh_f.C):
            ijk = mesh_f.cell_ijk(c)
            if unstructured:
                self.parent[c] = (int(parent[c]) & 0x0FFFFFFF) if self.geometric else c
                if self.geometric:
                    self.child[c] = [(int(parent[c]) >> (28 + d)) & 1 for d in range(dim)]
            elif self.geometric:
                self.parent[c] = mesh_c.cell_lex(tuple(i // 2 for i in ijk))
                self.child[c] = [i % 2 for i in ijk]
            else:
                self.parent[c] = c
        val = np.zeros(op_f.n_dofs)
        np.add.at(val, op_f.idx.reshape(-1), (op_f.cell_dofs != int(INVALID)).astype(np.float64).reshape(-1))
        self.w = 1.0 / np.maximum(val, 1.0)

    def _local(self, c, transpose, u):
        n_f, n_c = self.op_f.n, self.op_c.n
        u = u.reshape((n_c if not transpose else n_f,) * self.dim)
        for d in range(self.dim):
            P = self.P1[self.child[c, d]]
            u = _apply_1d(P.T if transpose else P, u[None], u.ndim - d)[0]
        return u.reshape(-1)

    def prolongate_and_add(self, dst, src):
        of, oc = self.op_f, self.op_c
        src = np.asarray(src, dtype=np.float64)
        add = np.zeros(of.n_dofs)
        for c in range(of.idx.shape[0]):
            p = self.parent[c]
            loc = self._local(c, False, src[oc.idx[p]] * oc.mask[p])
            np.add.at(add, of.idx[c], loc * self.w[of.idx[c]] * of.mask[c])
        return (np.asarray(dst, dtype=np.float64) + add).astype(of.dtype)

    def restrict_and_add(self, dst, src):
        of, oc = self.op_f, self.op_c
        src = np.asarray(src, dtype=np.float64)
        add = np.zeros(oc.n_dofs)
        for c in range(of.idx.shape[0]):
            p = self.parent[c]
            loc = self._local(c, True, src[of.idx[c]] * self.w[of.idx[c]] * of.mask[c])
            np.add.at(add, oc.idx[p], loc * oc.mask[p])
        return (np.asarray(dst, dtype=np.float64) + add).astype(oc.dtype)


class Multigrid:
    """V-cycle preconditioner (deal.II Multigrid::level_v_step through PreconditionMG::vmult):
         level 0:  x_0 = coarse.vmult(d_0)
         level l:  x_l = S_l.vmult(d_l);  t = d_l - A_l x_l;  d_{l-1} = R t;  recurse;  x_l += P x_{l-1};  x_l = S_l.step(x_l, d_l)
    ops[l], smoothers[l] (smoothers[0] = coarse-grid solver), transfers[l] between levels l and l - 1 (l >= 1).  The level
    number type is that of the level operators (float in the reference's matrix-free set-up,
    element_centered_preconditioners_01.cc:787-792); the outer vectors are double."""

    def __init__(self, ops, smoothers, transfers):
        self.ops, self.smoothers, self.transfers = ops, smoothers, transfers
        self.L = len(ops) - 1

    def _level(self, l, d):
        if l == 0:
            return self.smoothers[0].vmult(d)
        op = self.ops[l]
        x = self.smoothers[l].vmult(d)
        t = (d - op.vmult(x)).astype(op.dtype)
        t[op.constrained] = 0
        dc = self.transfers[l].restrict_and_add(np.zeros(self.ops[l - 1].n_dofs, dtype=self.ops[l - 1].dtype), t)
        xc = self._level(l - 1, dc)
        x = self.transfers[l].prolongate_and_add(x, xc)
        return self.smoothers[l].step(x, d)

    def vmult(self, src):
        op = self.ops[self.L]
        d = np.asarray(src).astype(op.dtype)
        return self._level(self.L, d).astype(np.float64)


# --------------------------------------------------------------------------------------
# Exact-block additive Schwarz  (include/preconditioners.h:528-605 RestrictedMatrixView, 744-813 RestrictedPreconditioner;
# include/restrictors.h:48-338 ElementCenteredRestrictor)
# --------------------------------------------------------------------------------------
class ExactBlockASM(FDMPreconditioner):
    """Same patches and weights as the FDM preconditioner, but the block of a patch is the inverse of the restriction
    R_c A R_c^T of the assembled operator matrix (gauss_jordan in the reference) instead of the fast-diagonalisation
    approximation.  `A` : dense matrix of the operator (LaplaceOperator.dense())."""

    def __init__(self, A, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.block_inv = []
        for c in range(self.idx.shape[0]):
            sel = self.mask[c] > 0
            ii = self.idx[c][sel]
            B = np.eye(self.idx.shape[1])
            if ii.size:
                B[np.ix_(sel, sel)] = np.linalg.inv(A[np.ix_(ii, ii)])
            self.block_inv.append(B.astype(self.dtype))

    def apply_inverse(self, rl):
        return np.stack([self.block_inv[c] @ rl[c] for c in range(rl.shape[0])])


# --------------------------------------------------------------------------------------
# Orientation-aware compressed access (include/reduced_access.h:66-152, 286-770): the 3^dim start indices of a cell plus one
# orientation word (dim = 2: 4 line bits; dim = 3: 12 line bits + 6 x 3 quad bits).  Pinned by reduced_access_01.result (2-D) and
# reduced_access_02.result (3-D) - line flips and quad flag 1 (transposed quad); the other quad flags follow deal.II's
# ShapeInfo::compute_orientation_table as recollected and are NOT pinned by a golden.
# --------------------------------------------------------------------------------------
def orientation_table(n):
    """[8, n*n]: position inside a quad of n x n interior DoFs for the 8 (orientation, flip, rotation) combinations; row 0 is the
    standard orientation, row 1 the transposed quad (pinned), rows 2..7 unpinned."""
    t = np.zeros((8, n * n), dtype=np.int64)
    for j in range(n):
        for i in range(n):
            q = i + j * n
            t[0, q] = i + j * n
            t[1, q] = j + i * n
            t[2, q] = j + (n - 1 - i) * n
            t[3, q] = i + (n - 1 - j) * n
            t[4, q] = (n - 1 - i) + (n - 1 - j) * n
            t[5, q] = (n - 1 - j) + (n - 1 - i) * n
            t[6, q] = (n - 1 - j) + i * n
            t[7, q] = (n - 1 - i) + j * n
    return t


def compress_orientation(orientations, do_post=False):
    """include/reduced_access.h:66-152 (orientations: 4 line flags in 2-D; 12 line flags + 6 quad flags in 3-D, deal.II order)."""
    o = list(orientations)
    word = 0
    if len(o) == 4:
        order = [2, 3, 0, 1] if do_post else [2, 0, 1, 3]
        for s, i in enumerate(order):
            word += o[i] << s
    elif len(o) == 18:
        if do_post:
            shift = 0
            for i in (2, 3, 6, 7, 0, 1, 4, 5, 8, 9, 10, 11):
                word |= o[i] << shift
                shift += 1
            for i in range(6):
                word |= o[12 + i] << shift
                shift += 3
        else:
            tab = [(1, o[2]), (1, o[0]), (3, o[16]), (1, o[1]), (1, o[3]),
                   (1, o[8]), (3, o[14]), (1, o[9]), (3, o[12]), (3, o[13]), (1, o[10]), (3, o[15]), (1, o[11]),
                   (1, o[6]), (1, o[4]), (3, o[17]), (1, o[5]), (1, o[7])]
            s = 0
            for width, val in tab:
                word += val << s
                s += width
    else:
        raise NotImplementedError
    return word


def adjust_for_orientation(dim, degree, local, orientation, table, integrate=False):
    """include/reduced_access.h:528-702: in-place reorientation of the (degree+1)^dim local values that were gathered in the
    standard layout (`orientation` = post word of compress_orientation)."""
    if dim == 1 or orientation == 0:
        return local
    np_, np2 = degree + 1, (degree + 1) ** 2
    n_lines = 4 if dim == 2 else 12
    v = local
    if dim == 2 or (orientation & 0b111111111111):
        for l in range(n_lines):
            if orientation & 1:
                stride = np_ ** (l // (2 if dim == 2 else 4))
                if dim == 2:
                    begin = [1, degree * degree + degree + 1, degree + 1, 2 * degree + 1][l]
                else:
                    begin = [1, np_ * degree + 1, np2 * degree + 1, np2 * degree + np_ * degree + 1, np_, np_ + degree,
                             np2 * degree + np_, np2 * degree + np_ + degree, np2, np2 + degree, np2 + np_ * degree,
                             np2 + np_ * degree + degree][l]
                for i0 in range((degree - 1) // 2):
                    a, b = begin + i0 * stride, begin + (degree - 2 - i0) * stride
                    v[a], v[b] = v[b], v[a]
            orientation >>= 1
    elif dim == 3:
        orientation >>= 12
    if dim == 3 and orientation != 0:
        for q in range(6):
            flag = orientation & 0b111
            if flag != 0:
                d = q // 2
                stride0 = np_ if d == 0 else 1
                stride1 = np_ if d == 2 else np2
                begin = 0 if q % 2 == 0 else (np_ ** (q // 2)) * degree
                temp = [None] * ((degree - 1) ** 2)
                i = 0
                for i1 in range(1, degree):
                    for i0 in range(1, degree):
                        if integrate:
                            j = table[flag][(i0 - 1) + (i1 - 1) * (degree - 1)]
                            i0_, i1_ = (j % (degree - 1)) + 1, (j // (degree - 1)) + 1
                            temp[i] = v[begin + i0_ * stride0 + i1_ * stride1]
                        else:
                            temp[table[flag][i]] = v[begin + i0 * stride0 + i1 * stride1]
                        i += 1
                i = 0
                for i1 in range(1, degree):
                    for i0 in range(1, degree):
                        v[begin + i0 * stride0 + i1 * stride1] = temp[i]
                        i += 1
            orientation >>= 3
    return v


def gather_post(global_vector, dim, degree, dofs_of_cell, orientation, table):
    """include/reduced_access.h:704-770: standard expansion of the 3^dim start indices, then adjust_for_orientation."""
    comp = np.asarray(dofs_of_cell, dtype=np.uint32).reshape(1, -1)
    idx = expand_compressed(comp, degree, dim)[0].astype(np.int64)
    local = [global_vector[i] for i in idx]
    return adjust_for_orientation(dim, degree, local, orientation, table, False)


def _rotr32(x, r):
    r %= 32
    return ((x >> r) | (x << (32 - r))) & 0xFFFFFFFF


def gather_oriented(global_vector, dim, degree, dofs_of_cell, orientation_in, table):
    """include/reduced_access.h:286-526: gather with the orientation applied on the fly (`orientation_in` = pre word)."""
    g, d = global_vector, list(dofs_of_cell)
    out = []
    if dim == 2:
        orientation, offset, compressed = orientation_in, 0, 0
        for j in range(degree + 1):
            ind = d[compressed * 3: compressed * 3 + 3]
            if orientation and (orientation & 1) and (j == 0 or j == degree):
                out.append(g[ind[0]])
                out += [g[ind[1] + (degree - 2 - i)] for i in range(degree - 1)]
                out.append(g[ind[2]])
            elif orientation and (orientation & 0b11) and 0 < j < degree:
                out.append(g[ind[0] + (degree - 2 - offset)] if orientation & 0b01 else g[ind[0] + offset])
                out += [g[ind[1] + offset * (degree - 1) + i] for i in range(degree - 1)]
                out.append(g[ind[2] + (degree - 2 - offset)] if orientation & 0b10 else g[ind[2] + offset])
            else:
                out.append(g[ind[0] + offset])
                out += [g[ind[1] + offset * (degree - 1) + i] for i in range(degree - 1)]
                out.append(g[ind[2] + offset])
            if j == 0 or j == degree - 1:
                compressed += 1
                offset = 0
                orientation = orientation >> 1 if j == 0 else orientation >> 2
            else:
                offset += 1
        return out
    orientation = orientation_in
    o_ptr, compressed_k, offset_k = orientation, 0, 0
    for k in range(degree + 1):
        compressed_j = offset_j = 0
        for j in range(degree + 1):
            offset = ((degree - 1) if compressed_j == 1 else 1) * offset_k + offset_j
            ind = d[3 * (compressed_k * 3 + compressed_j): 3 * (compressed_k * 3 + compressed_j) + 3]
            k_end, j_end = (k == 0 or k == degree), (j == 0 or j == degree)
            if orientation != 0 and (o_ptr & 1) and k_end and j_end:
                out.append(g[ind[0]])
                out += [g[ind[1] + (degree - 2 - i)] for i in range(degree - 1)]
                out.append(g[ind[2]])
            elif orientation != 0 and (o_ptr & 0b11111) and ((k_end and 0 < j < degree) or (0 < k < degree and j_end)):
                jk = j if k_end else k
                out.append(g[ind[0] + (degree - 1 - jk)] if o_ptr & 0b00001 else g[ind[0] + (jk - 1)])
                qf = (o_ptr >> 1) & 0b111
                for i in range(degree - 1):
                    out.append(g[ind[1] + (table[qf][(degree - 1) * (jk - 1) + i] if qf != 0 else (degree - 1) * (jk - 1) + i)])
                out.append(g[ind[2] + (degree - 1 - jk)] if o_ptr & 0b10000 else g[ind[2] + (jk - 1)])
            elif orientation != 0 and (o_ptr & 0b111111) and 0 < k < degree and 0 < j < degree:
                q0, q1 = o_ptr & 0b111, (o_ptr >> 3) & 0b111
                out.append(g[ind[0] + (table[q0][offset] if q0 != 0 else offset)])
                out += [g[ind[1] + offset * (degree - 1) + i] for i in range(degree - 1)]
                out.append(g[ind[2] + (table[q1][offset] if q1 != 0 else offset)])
            else:
                out.append(g[ind[0] + offset])
                out += [g[ind[1] + offset * (degree - 1) + i] for i in range(degree - 1)]
                out.append(g[ind[2] + offset])
            if j == 0 or j == degree - 1:
                compressed_j += 1
                offset_j = 0
            else:
                offset_j += 1
            if k_end:
                if j_end:
                    o_ptr = _rotr32(o_ptr, 1)
                elif j == degree - 1:
                    o_ptr = _rotr32(o_ptr, 5)
            else:
                if j_end:
                    o_ptr = _rotr32(o_ptr, 5)
                elif j == degree - 1:
                    o_ptr = _rotr32(o_ptr, 6)
            if 0 < k < degree - 1 and j == degree:
                o_ptr = _rotr32(o_ptr, 32 - 16)
        if k == 0 or k == degree - 1:
            compressed_k += 1
            offset_k = 0
        else:
            offset_k += 1
    return out


# --------------------------------------------------------------------------------------
# Unstructured all-hex meshes (the ball of element_centered_preconditioners_01.cc:398-402): entity connectivity, the packed
# orientation word of a cell and the 3^3 compressed indices that deal.II's DoFHandler / ConstraintInfoReduced::initialize
# (include/vector_access_reduced.h:30-164, include/reduced_access.h:154-285) provide in the reference, restated for a mesh given by
# arrays.  parity unpinned against deal.II's own numbering (a permutation of it); the expansion is the pinned
# adjust_for_orientation above.
# --------------------------------------------------------------------------------------
class UnstructuredMesh:
    """vertices [V, 3]; cells [C, 8] vertex numbers in lexicographic order; support [C, 27, 3] support points of the triquadratic
    cell map (lexicographic; None: trilinear from the vertices).  An entity (line, quad) takes its frame from the first cell that
    contains it: a line is flipped in a cell that runs it the other way, the code f of a quad is the row of orientation_table
    under which the entity's DoFs appear in the cell's face layout."""
    dim = 3

    def __init__(self, vertices, cells, support=None, dirichlet=True):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.C = self.cells.shape[0]
        self.dirichlet = bool(dirichlet)
        self.cell_order = np.arange(self.C)
        if support is None:
            r = np.array([0.0, 0.5, 1.0])
            w = np.stack([1 - r, r], axis=-1)                                       # [3, 2]
            Xv = self.vertices[self.cells].reshape(self.C, 2, 2, 2, 3)              # [c, z, y, x, :]
            support = np.einsum("kz,jy,ix,nzyxd->nkjid", w, w, w, Xv).reshape(self.C, 27, 3)
        self.support = np.asarray(support, dtype=np.float64).reshape(self.C, 27, 3)
        self._connect()

    # local description of the 27 entities: e = ex + 3 ey + 9 ez, component 1 = "runs along this direction"
    @staticmethod
    def _corners(e):
        c = (e % 3, (e // 3) % 3, e // 9)
        free = [d for d in range(3) if c[d] == 1]
        out = []
        for p in range(2 ** len(free)):
            bits = [cc // 2 for cc in c]
            for t, d in enumerate(free):
                bits[d] = (p >> t) & 1
            out.append(bits[0] + 2 * bits[1] + 4 * bits[2])
        return out

    @staticmethod
    def _line_number(e):
        ex, ey, ez = e % 3, (e // 3) % 3, e // 9
        if ex == 1:
            return (ey == 2) + 2 * (ez == 2)
        if ey == 1:
            return 4 + (ex == 2) + 2 * (ez == 2)
        return 8 + (ex == 2) + 2 * (ey == 2)

    @staticmethod
    def _quad_number(e):
        c = (e % 3, (e // 3) % 3, e // 9)
        d = [i for i in range(3) if c[i] != 1][0]
        return 2 * d + (c[d] == 2)

    def _connect(self):
        t2 = orientation_table(2)                       # corner version of the 8 codes: entity corner q sits at local corner t2[f][q]
        lines, quads = {}, {}
        self.entity = np.zeros((self.C, 27), dtype=np.int64)
        self.orientation = np.zeros(self.C, dtype=np.int64)
        self.quad_cells = []
        for c in range(self.C):
            cv = self.cells[c]
            word = 0
            for e in range(27):
                loc = [int(cv[v]) for v in self._corners(e)]
                if len(loc) == 1:
                    self.entity[c, e] = loc[0]
                elif len(loc) == 2:
                    key = (min(loc), max(loc))
                    if key not in lines:
                        lines[key] = (len(lines), loc[0])
                    self.entity[c, e] = lines[key][0]
                    if lines[key][1] != loc[0]:
                        word |= 1 << self._line_number(e)
                elif len(loc) == 4:
                    key = tuple(sorted(loc))
                    if key not in quads:
                        quads[key] = (len(quads), loc)
                        self.quad_cells.append([c])
                    else:
                        self.quad_cells[quads[key][0]].append(c)
                    qid, frame = quads[key]
                    self.entity[c, e] = qid
                    code = [f for f in range(8) if all(loc[t2[f][q]] == frame[q] for q in range(4))]
                    word |= code[0] << (12 + 3 * self._quad_number(e))
                else:
                    self.entity[c, e] = c
            self.orientation[c] = word
        self.n_lines, self.n_quads = len(lines), len(quads)
        V = self.vertices.shape[0]
        self.quad_bnd = np.array([len(q) == 1 for q in self.quad_cells])
        self.vertex_bnd = np.zeros(V, dtype=bool)
        self.line_bnd = np.zeros(self.n_lines, dtype=bool)
        for c in range(self.C):
            for fe in (12, 14, 10, 16, 4, 22):
                if self.quad_bnd[self.entity[c, fe]]:
                    fc = (fe % 3, (fe // 3) % 3, fe // 9)
                    d = [i for i in range(3) if fc[i] != 1][0]
                    for e in range(27):
                        ec = (e % 3, (e // 3) % 3, e // 9)
                        if ec[d] != fc[d]:
                            continue
                        nd = sum(1 for x in ec if x == 1)
                        if nd == 0:
                            self.vertex_bnd[self.entity[c, e]] = True
                        elif nd == 1:
                            self.line_bnd[self.entity[c, e]] = True

    def number_dofs(self, k):
        """returns (cell_dofs [C, n^3] oriented addresses with INVALID on constrained entities, n_dofs, constrained mask,
        compressed [C, 27] with constrained entities INVALID, compressed_plain)."""
        V = self.vertices.shape[0]
        pl, pq, ph = k - 1, (k - 1) ** 2, (k - 1) ** 3
        line0 = V
        quad0 = line0 + self.n_lines * pl
        hex0 = quad0 + self.n_quads * pq
        n_dofs = hex0 + self.C * ph
        comp_plain = np.zeros((self.C, 27), dtype=np.int64)
        bnd = np.zeros((self.C, 27), dtype=bool)
        for e in range(27):
            nd = (e % 3 == 1) + ((e // 3) % 3 == 1) + (e // 9 == 1)
            ids = self.entity[:, e]
            if nd == 0:
                comp_plain[:, e], bnd[:, e] = ids, self.vertex_bnd[ids]
            elif nd == 1:
                comp_plain[:, e], bnd[:, e] = line0 + ids * pl, self.line_bnd[ids]
            elif nd == 2:
                comp_plain[:, e], bnd[:, e] = quad0 + ids * pq, self.quad_bnd[ids]
            else:
                comp_plain[:, e] = hex0 + ids * ph
        comp = np.where(bnd & self.dirichlet, int(INVALID), comp_plain)
        constrained = np.zeros(n_dofs, dtype=bool)
        if self.dirichlet:
            constrained[np.nonzero(self.vertex_bnd)[0]] = True
            for l in np.nonzero(self.line_bnd)[0]:
                constrained[line0 + l * pl: line0 + (l + 1) * pl] = True
            for q in np.nonzero(self.quad_bnd)[0]:
                constrained[quad0 + q * pq: quad0 + (q + 1) * pq] = True
        std = expand_compressed(comp.astype(np.uint32), k, 3).astype(np.int64)
        table = orientation_table(k - 1)
        cell_dofs = np.zeros_like(std)
        for c in range(self.C):
            cell_dofs[c] = adjust_for_orientation(3, k, list(std[c]), int(self.orientation[c]), table, False)
        return cell_dofs, n_dofs, constrained, comp.astype(np.uint32), comp_plain.astype(np.uint32)

    # ---- geometry (MappingQCache(2) data = the support points) ------------------------------------------------------------------
    def cell_ijk(self, c):
        return c

    def neighbor(self, c, d, side):
        """anything but None when the face 2 d + side is interior (FDMPreconditioner only asks whether there is a neighbour)."""
        fe = (12, 14, 10, 16, 4, 22)[2 * d + side]
        return None if self.quad_bnd[self.entity[c, fe]] else True

    def cell_points(self, c, ref):
        nodes = np.array([0.0, 0.5, 1.0])
        ref = np.asarray(ref)
        Vs = [lagrange(nodes, ref[..., d].reshape(-1))[0] for d in range(3)]
        X = self.support[c].reshape(3, 3, 3, 3)
        return np.einsum("pk,pj,pi,kjid->pd", Vs[2], Vs[1], Vs[0], X).reshape(ref.shape[:-1] + (3,))

    def jacobians(self, basis):
        nodes = np.array([0.0, 0.5, 1.0])
        V, D = lagrange(nodes, basis.qp)
        nq = basis.n
        J = np.zeros((self.C, nq ** 3, 3, 3))
        X = self.support.reshape(self.C, 3, 3, 3, 3)
        for e in range(3):
            mats = [D if d == e else V for d in range(3)]
            g = np.einsum("ck,bj,ai,nkjid->ncbad", mats[2], mats[1], mats[0], X)
            J[:, :, :, e] = g.reshape(self.C, nq ** 3, 3)
        return J

    def harmonic_patch_extents(self, basis):
        """include/grid_tools.h:54-138: every cell adds its extent normal to a face to that face; a neighbour's extent is the
        face total minus the own one (0 on the boundary)."""
        ext = harmonic_cell_extents(self, basis)
        face = np.zeros(self.n_quads)
        fes = (12, 14, 10, 16, 4, 22)
        for c in range(self.C):
            for d in range(3):
                face[self.entity[c, fes[2 * d]]] += ext[c, d]
                face[self.entity[c, fes[2 * d + 1]]] += ext[c, d]
        out = np.zeros((self.C, 3, 3))
        for c in range(self.C):
            for d in range(3):
                out[c, d, 1] = ext[c, d]
                for side in (0, 1):
                    q = self.entity[c, fes[2 * d + side]]
                    out[c, d, 2 * side] = 0.0 if self.quad_bnd[q] else face[q] - ext[c, d]
        return out


# --------------------------------------------------------------------------------------
# Power kernel (power_kernel_01.likwid.cc): dst_0 = A src, dst_1 = M dst_0 with the second operator applied per wave of cells
# --------------------------------------------------------------------------------------
def mass_vmult(op, x, JxW):
    """process_batch_post of power_kernel_01.likwid.cc:423-440 summed over the cells: evaluate(values), submit_value(get_value),
    integrate(values), distribute_local_to_global; `op` supplies the index lists and the basis, JxW [C, n^dim] the weights."""
    x = np.asarray(x, dtype=op.dtype)
    C = op.idx.shape[0]
    u = (x[op.idx] * op.mask).reshape((C,) + (op.n,) * op.dim)
    for d in range(op.dim):
        u = _apply_1d(op.N, u, u.ndim - 1 - d)
    u = u * np.asarray(JxW, dtype=op.dtype).reshape(u.shape)
    for d in range(op.dim):
        u = _apply_1d(op.N.T, u, u.ndim - 1 - d)
    y = np.zeros(op.n_dofs, dtype=op.dtype)
    np.add.at(y, op.idx.reshape(-1), (u.reshape(C, -1) * op.mask).reshape(-1))
    return y


def determine_pre_post(cell_vertices, cell_granularity, batch_size=1, track_individual_cell=True):
    """power_kernel_01.likwid.cc:122-260: cells are processed in waves of `cell_granularity`; a vertex remembers the first / last wave
    that touches it; an entity (a cell, or a batch of `batch_size` consecutive cells) may be pre-processed in the first wave and
    post-processed in the last wave over its vertices.  Returns (pre_indices, pre_ptr, post_indices, post_ptr)."""
    cv = np.asarray(cell_vertices)
    C = cv.shape[0]
    g = C if cell_granularity <= 0 else cell_granularity
    wave = np.arange(C) // g
    n_waves = int(wave[-1]) + 1
    first = np.full(int(cv.max()) + 1, np.iinfo(np.int64).max, dtype=np.int64)
    last = np.zeros(int(cv.max()) + 1, dtype=np.int64)
    for c in range(C):
        first[cv[c]] = np.minimum(first[cv[c]], wave[c])
        last[cv[c]] = np.maximum(last[cv[c]], wave[c])
    ent = np.arange(C) if track_individual_cell else np.arange(C) // batch_size
    n_ent = int(ent[-1]) + 1
    mn = np.full(n_ent, np.iinfo(np.int64).max, dtype=np.int64)
    mx = np.zeros(n_ent, dtype=np.int64)
    for c in range(C):
        mn[ent[c]] = min(mn[ent[c]], first[cv[c]].min())
        mx[ent[c]] = max(mx[ent[c]], last[cv[c]].max())

    def process(ids):
        temp = [[] for _ in range(n_waves)]
        for i, w in enumerate(ids):
            temp[w].append(i)
        ptr = np.cumsum([0] + [len(t) for t in temp])
        return np.array([i for t in temp for i in t], dtype=np.int64), ptr

    return process(mn) + process(mx)
