"""CPU test (no device): the library's DoF numbering (dasm_mesh_host_numbering) against the oracle's independent
restatement, bit-exact, on meshes with lex bricks (full 4x4x4 bricks with neighbours across all faces, numbered as
lexicographic boxes so that the TMA engine can move them), irregular bricks and Dirichlet boundaries.

The 27 start indices per cell are the standard-orientation layout of include/vector_access_reduced.h:30-164."""
import numpy as np
import pytest

import dasm_oracle as o
from __graft_entry__ import load_package

CASES = [
    ((8, 8, 8), (1, 1, 1), True, 4),      # 8 lex bricks
    ((8, 4, 12), (1, 1, 1), True, 3),     # lex bricks, one brick in y (its own neighbour)
    ((12, 12, 12), (0, 0, 0), True, 2),   # Dirichlet: one interior lex brick, 26 irregular ones
    ((9, 8, 6), (1, 1, 1), True, 2),      # truncated bricks next to lex bricks
    ((12, 8, 8), (0, 1, 1), False, 3),    # natural boundary in x
    ((4, 3, 5), (1, 1, 1), True, 1),      # no lex brick
    ((8, 8, 4), (1, 1, 1), True, 5),
]


@pytest.mark.parametrize("nc,periodic,dirichlet,k", CASES)
def test_numbering_matches_oracle(nc, periodic, dirichlet, k):
    pkg = load_package()
    mesh = pkg.Mesh(None, nc, periodic=periodic, dirichlet=dirichlet)
    nb = mesh.host_numbering(k)
    om = o.StructuredMesh(3, nc, periodic, dirichlet=dirichlet)
    om.cell_order = o.brick_major_order(nc)
    cd, nd, con, comp = o.number_dofs_owner_cell(om, k)
    assert nb["n_owned"] == nd and nb["n_ghost"] == 0
    assert np.array_equal(nb["cidx_plain"], comp[om.cell_order])
    # every DoF is referenced, the expansion is a bijection onto [0, nd) over the cells' unique DoFs
    idx = o.expand_compressed(comp, k, 3).astype(np.int64)
    assert idx.min() == 0 and idx.max() == nd - 1 and len(np.unique(idx)) == nd
    n_lex = len(o.lex_bricks(om))
    lex_entries = (comp != o.INVALID) & ((comp & o.LEX_FLAG) != 0)
    assert lex_entries.any() == (n_lex > 0)
    if n_lex:
        # the lex boxes are the first n_lex blocks of 64 k^3 DoFs
        assert (comp[lex_entries] & (o.LEX_FLAG - 1)).max() < n_lex * 64 * k ** 3
        rest = comp[~lex_entries & (comp != o.INVALID)]
        assert rest.size == 0 or rest.min() >= n_lex * 64 * k ** 3


def test_lex_box_is_lexicographic():
    """the DoF at tile point (X, Y, Z) of lex brick i is i (4k)^3 + X + 4k Y + 16 k^2 Z."""
    k, nc = 3, (8, 8, 8)
    om = o.StructuredMesh(3, nc, (1, 1, 1))
    om.cell_order = o.brick_major_order(nc)
    cd, nd, con, comp = o.number_dofs_owner_cell(om, k)
    n = k + 1
    cd = cd.reshape(-1, n, n, n)
    for c in range(om.C):
        ijk = om.cell_ijk(c)
        b = (ijk[2] // 4 * 2 + ijk[1] // 4) * 2 + ijk[0] // 4
        for z in range(k):
            for y in range(k):
                for x in range(k):
                    X, Y, Z = (ijk[0] % 4) * k + x, (ijk[1] % 4) * k + y, (ijk[2] % 4) * k + z
                    assert cd[c, z, y, x] == b * (4 * k) ** 3 + X + 4 * k * (Y + 4 * k * Z)
