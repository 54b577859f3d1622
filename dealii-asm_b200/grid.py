"""Unstructured all-hex meshes as arrays for LaplaceOperatorMatrixFree.from_arrays (dasm_op_create_unstructured).

hyper_ball(): the ball of BASELINE configs[3] (reference: GridGenerator::hyper_ball_balanced + refine_global + MappingQCache(2),
element_centered_preconditioners_01.cc:398-402, 415-425).  deal.II is not vendored in the reference, so the generator is this
library's: the same coarse topology as hyper_ball_balanced in 3-D (32 cells: a 2 x 2 x 2 inner block and 4 cells over each of its
6 faces, outer faces on the sphere), every refinement splits a cell into 8 and new points follow the coarse cell's map (inner
cells: trilinear; outer cells: linear blend between the inner face and the radial projection of the outer face), and every cell
carries the 27 support points of its triquadratic map.  Cell frames differ between the pieces, so lines and quads of
neighbouring cells meet with non-standard orientation (the case ConstraintInfoReduced's orientation word exists for).

rotated_cube(): an n^3 box whose cells get pseudo-random right-handed local frames: exercises all line flips and all 8 quad codes.
"""
import itertools

import numpy as np


def _dedup(points, tol=1e-10):
    """unique rows of points [N, 3] up to `tol` -> (unique [V, 3], inverse [N]) with vertices in order of first appearance.
    Every coordinate axis is clustered on its own (sorted values, a new cluster where the gap exceeds tol), a vertex is the triple of
    its cluster numbers: no rounding boundary can split two evaluations of the same point."""
    keys = np.zeros(points.shape, dtype=np.int64)
    for d in range(points.shape[1]):
        order = np.argsort(points[:, d], kind="stable")
        v = points[order, d]
        cluster = np.concatenate([[0], np.cumsum(np.diff(v) > tol)])
        keys[order, d] = cluster
    _, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)                    # unique ids sorted by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return points[first[order]], rank[inv.reshape(-1)]


def _ball_coarse(radius, a, b, c):
    """list of maps ref [N, 3] -> x [N, 3], one per coarse cell"""
    mag = {0: 0.0, 1: a, 2: b, 3: c}

    def lattice(i, j, l):
        return radius * np.array([i, j, l], dtype=np.float64) * mag[abs(i) + abs(j) + abs(l)]

    def sphere(i, j, l):
        v = np.array([i, j, l], dtype=np.float64)
        return radius * v / np.linalg.norm(v)

    maps = []

    def trilinear(corners):
        corners = np.asarray(corners)            # [8, 3] lexicographic

        def f(ref):
            w = [np.stack([1 - ref[:, d], ref[:, d]], axis=-1) for d in range(3)]
            return np.einsum("pk,pj,pi,kjid->pd", w[2], w[1], w[0], corners.reshape(2, 2, 2, 3))
        return f

    for l0, j0, i0 in itertools.product((-1, 0), repeat=3):
        maps.append(trilinear([lattice(i0 + di, j0 + dj, l0 + dl) for dl in (0, 1) for dj in (0, 1) for di in (0, 1)]))

    def shell(inner, outer):
        inner, outer = np.asarray(inner).reshape(2, 2, 3), np.asarray(outer).reshape(2, 2, 3)

        def f(ref):
            w = [np.stack([1 - ref[:, d], ref[:, d]], axis=-1) for d in range(2)]
            xi = np.einsum("pj,pi,jid->pd", w[1], w[0], inner)
            xo = np.einsum("pj,pi,jid->pd", w[1], w[0], outer)
            xo = radius * xo / np.linalg.norm(xo, axis=-1, keepdims=True)
            z = ref[:, 2:3]
            return (1 - z) * xi + z * xo
        return f

    for d in range(3):
        for s in (-1, 1):
            d1, d2 = ((d + 1) % 3, (d + 2) % 3) if s == 1 else ((d + 2) % 3, (d + 1) % 3)   # right-handed (xi, eta, radial)
            for b0, a0 in itertools.product((-1, 0), repeat=2):
                inner, outer = [], []
                for db in (0, 1):
                    for da in (0, 1):
                        idx = [0, 0, 0]
                        idx[d], idx[d1], idx[d2] = s, a0 + da, b0 + db
                        inner.append(lattice(*idx))
                        outer.append(sphere(*idx))
                maps.append(shell(inner, outer))
    return maps


def hyper_ball(n_refinements=0, radius=1.0, a=0.528, b=0.4533, c=0.3752):
    """returns dict(vertices [V, 3], cells [C, 8] uint32, support [C, 27, 3]); C = 32 * 8^n_refinements"""
    maps = _ball_coarse(radius, a, b, c)
    m = 2 ** n_refinements
    nodes = np.array([0.0, 0.5, 1.0])
    # reference support points of all children of one coarse cell: [m^3, 27, 3]
    child = np.array([(i, j, l) for l in range(m) for j in range(m) for i in range(m)], dtype=np.float64)
    loc = np.array([(x, y, z) for z in nodes for y in nodes for x in nodes])
    ref = ((child[:, None, :] + loc[None, :, :]) / m).reshape(-1, 3)
    support = np.concatenate([f(ref).reshape(m ** 3, 27, 3) for f in maps], axis=0)
    corner = [0, 2, 6, 8, 18, 20, 24, 26]
    vertices, inv = _dedup(support[:, corner, :].reshape(-1, 3))
    return dict(vertices=np.ascontiguousarray(vertices), cells=inv.reshape(-1, 8).astype(np.uint32), support=np.ascontiguousarray(support))


def _rotations():
    """the 24 right-handed frames: (perm, signs) with local axis e -> global axis perm[e], direction signs[e]"""
    out = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            M = np.zeros((3, 3))
            for e in range(3):
                M[perm[e], e] = signs[e]
            if np.linalg.det(M) > 0:
                out.append((perm, signs))
    return out


def rotated_cube(n=2, seed=0, mapfun=None, rotate=True):
    """n^3 cells of the unit cube; cell c uses a pseudo-random right-handed local frame (rotate=False: the standard one).
    mapfun(X [..., 3]) deforms the geometry (evaluated at the 27 support points of every cell)."""
    rng = np.random.default_rng(seed)
    rots = _rotations()
    nodes = np.array([0.0, 0.5, 1.0])
    loc = np.array([(x, y, z) for z in nodes for y in nodes for x in nodes])   # local reference points
    support = []
    for l, j, i in itertools.product(range(n), repeat=3):
        perm, signs = rots[rng.integers(len(rots))] if rotate else ((0, 1, 2), (1, 1, 1))
        g = np.zeros_like(loc)
        for e in range(3):
            g[:, perm[e]] = loc[:, e] if signs[e] == 1 else 1 - loc[:, e]
        support.append((g + np.array([i, j, l])) / n)
    support = np.array(support)
    corner = [0, 2, 6, 8, 18, 20, 24, 26]
    vertices, inv = _dedup(support[:, corner, :].reshape(-1, 3))
    if mapfun is not None:
        support = mapfun(support)
        vertices = mapfun(vertices)
    return dict(vertices=np.ascontiguousarray(vertices), cells=inv.reshape(-1, 8).astype(np.uint32), support=np.ascontiguousarray(support))


def ball_parents(n_refinements):
    """parent[fine cell] = coarse cell | child position << 28 between hyper_ball(n_refinements) and hyper_ball(n_refinements - 1)
    (dasm_transfer_create_unstructured): the children of a cell keep its frame"""
    mf = 2 ** n_refinements
    mc = mf // 2
    f = np.arange(32 * mf ** 3, dtype=np.int64)
    q, r = f // mf ** 3, f % mf ** 3
    i, j, l = r % mf, (r // mf) % mf, r // (mf * mf)
    parent = q * mc ** 3 + ((l // 2) * mc + j // 2) * mc + i // 2
    code = (i & 1) | ((j & 1) << 1) | ((l & 1) << 2)
    return (parent | (code << 28)).astype(np.uint32)
