"""CPU test of the even-odd blocks the warp-specialised kernels use for every 1-D contraction (EOMat / mat_vec in
dealii-asm_b200/csrc/kernels_fast.cuh): the host packers of libdasm (through the host-only hook dasm_test_eo_pack) and a
numpy restatement of the device-side mat_vec reproduce the dense products for centrosymmetric matrices (mass /
stiffness), forward and backward eigenvector matrices with even-first ordering."""
import ctypes

import numpy as np
import pytest

from __graft_entry__ import load_package


def eo_apply(P, Q, v, pre, post):
    """restatement of mat_vec<n, T, PRE, POST, false>"""
    n = len(v)
    m, h = (n + 1) // 2, n // 2
    if pre:
        e = np.array([v[i] + v[n - 1 - i] for i in range(h)] + ([v[h]] if m > h else []))
        o = np.array([v[i] - v[n - 1 - i] for i in range(h)])
    else:
        e, o = v[:m].copy(), v[m:].copy()
    p, q = P @ e, (Q @ o if h else np.zeros(0))
    if not post:
        return np.concatenate([p, q])
    r = np.zeros(n)
    for a in range(h):
        r[a] = p[a] + q[a]
        r[n - 1 - a] = p[a] - q[a]
    if m > h:
        r[h] = p[h]
    return r


def pack(pkg, kind, A):
    n = A.shape[0]
    m, h = (n + 1) // 2, n // 2
    P, Q = np.zeros(m * m), np.zeros(max(h * h, 1))
    Ac = np.ascontiguousarray(A, dtype=np.float64)
    rc = pkg.lib().dasm_test_eo_pack(n, kind, Ac.ctypes.data_as(ctypes.c_void_p), P.ctypes.data_as(ctypes.c_void_p),
                                     Q.ctypes.data_as(ctypes.c_void_p))
    return rc, P.reshape(m, m), Q[:h * h].reshape(h, h)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_even_odd_blocks(n):
    pkg = load_package()
    rng = np.random.default_rng(n)
    J = np.eye(n)[::-1]
    m = (n + 1) // 2
    # centrosymmetric (and symmetric, like the 1-D mass / stiffness matrices)
    B = rng.uniform(-1, 1, (n, n))
    A = B + B.T
    A = A + J @ A @ J
    v = rng.uniform(-1, 1, n)
    rc, P, Q = pack(pkg, 0, A)
    assert rc == 0
    assert np.allclose(eo_apply(P, Q, v, True, True), A @ v, rtol=1e-14, atol=1e-14)
    # eigenvectors of a centrosymmetric symmetric matrix are even or odd: order the even ones first
    _, S = np.linalg.eigh(A + 3 * n * np.eye(n))
    even = [a for a in range(n) if np.allclose(J @ S[:, a], S[:, a], atol=1e-10)]
    odd = [a for a in range(n) if np.allclose(J @ S[:, a], -S[:, a], atol=1e-10)]
    assert len(even) == m and len(even) + len(odd) == n
    S = S[:, even + odd]
    w = np.abs(rng.uniform(0.5, 1, n))
    w = 0.5 * (w + w[::-1])  # symmetric weights
    fwd = (S * w[:, None]).T          # u = S^T diag(w) v      rows: eigen index
    bwd = w[:, None] * S              # y = diag(w) S u        rows: nodal index
    rc, P, Q = pack(pkg, 1, fwd)
    assert rc == 0
    assert np.allclose(eo_apply(P, Q, v, True, False), fwd @ v, rtol=1e-13, atol=1e-13)
    rc, P, Q = pack(pkg, 2, bwd)
    assert rc == 0
    assert np.allclose(eo_apply(P, Q, v, False, True), bwd @ v, rtol=1e-13, atol=1e-13)
    # a matrix without the symmetry is rejected (the caller then uses the brick kernels)
    rc, _, _ = pack(pkg, 0, B)
    assert rc != 0
