"""helpers shared by the GPU parity tests, smoke() and bench.py's cpu_baseline: build the oracle's view of a
problem that was set up through the C ABI."""
import numpy as np

import dasm_oracle as o


def oracle_mesh(pmesh, brick=(4, 4, 4), mapping_degree=2):
    mapfun = None
    if pmesh.map_kind == "sine":
        mapfun = o.sine_map
    elif pmesh.map_kind == "kershaw":
        mapfun = o.kershaw_map(pmesh.map_params[0], pmesh.map_params[1])
    mesh = o.StructuredMesh(3, pmesh.n_cells_dir, pmesh.periodic, dirichlet=pmesh.dirichlet, mapfun=mapfun,
                            lengths=pmesh.length, mapping_degree=mapping_degree)
    mesh.cell_order = o.brick_major_order(pmesh.n_cells_dir, brick)
    return mesh


def oracle_problem(pkg, pmesh, op, n_overlap=1, weight_type="symm", dtype=np.float64, check_indices=True, with_fdm=True,
                   mapping_degree=2, element_centric=True):
    """returns (oracle LaplaceOperator, oracle FDMPreconditioner) in the library's DoF numbering; the numbering
    itself is recomputed by the oracle and compared bit-exactly with the library's compressed indices."""
    k = op.degree
    mesh = oracle_mesh(pmesh, mapping_degree=mapping_degree)
    cd, nd, con, comp = o.number_dofs_owner_cell(mesh, k)
    if check_indices:
        lib_comp = op.compressed_indices(plain=True)  # rows in processing order
        assert nd == op.n_dofs()
        assert np.array_equal(lib_comp, comp[mesh.cell_order]), "compressed DoF indices differ from the oracle"
        lib_con = np.sort(op.constrained_dofs())
        assert np.array_equal(lib_con, np.nonzero(con)[0].astype(np.uint32))
    b = o.Basis1D(k)
    G = o.merged_coefficients(mesh.jacobians(b), b, 3)
    oop = o.LaplaceOperator(3, k, cd, nd, con, G, dtype=dtype)
    oP = None
    if with_fdm:
        lex_rank = np.arange(mesh.C)  # RAS ownership by lexicographic cell id
        oP = o.FDMPreconditioner(mesh, k, cd, nd, con, n_overlap, weight_type, dtype=dtype, cell_rank=lex_rank,
                                 element_centric=element_centric)
    return oop, oP
