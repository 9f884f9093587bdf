"""Deterministic synthetic matrices for BASELINE.json's configs.
Grid node (x,y,z) -> index (z*g + y)*g + x (x fastest), matching `nd_grid_order`.
No stored zeros (the reference counts stored zeros as structural entries)."""
import numpy as np
import scipy.sparse as sp


def _tridiag(g, lo, d, up):
    return sp.diags([np.full(g - 1, lo), np.full(g, d), np.full(g - 1, up)], [-1, 0, 1], format="csc")


def _clean(a):
    a = sp.csc_matrix(a)
    a.eliminate_zeros(); a.sum_duplicates(); a.sort_indices()
    return a


def laplacian2d(g):
    """cfg1: 5-point Laplacian, A = I (x) T + T (x) I, T = tridiag(-1,2,-1); n = g^2."""
    T = _tridiag(g, -1.0, 2.0, -1.0); E = sp.identity(g, format="csc")
    return _clean(sp.kron(E, T) + sp.kron(T, E))


def laplacian3d(g):
    """cfg2/cfg4: 7-point Laplacian, diagonal 6, six -1 neighbours, Dirichlet truncation; n = g^3."""
    T = _tridiag(g, -1.0, 2.0, -1.0); E = sp.identity(g, format="csc")
    return _clean(sp.kron(sp.kron(E, E), T) + sp.kron(sp.kron(E, T), E) + sp.kron(sp.kron(T, E), E))


def convdiff3d(g, c=(0.5, 0.25, 0.125)):
    """cfg3: upwind convection-diffusion, diagonal 6+cx+cy+cz, west/south/down -(1+c), east/north/up -1.
    Row- and column-diagonally dominant, so the reference's in-supernode pivoting picks ipiv[k] = k."""
    E = sp.identity(g, format="csc")
    Tx = _tridiag(g, -(1.0 + c[0]), 2.0 + c[0], -1.0)
    Ty = _tridiag(g, -(1.0 + c[1]), 2.0 + c[1], -1.0)
    Tz = _tridiag(g, -(1.0 + c[2]), 2.0 + c[2], -1.0)
    return _clean(sp.kron(sp.kron(E, E), Tx) + sp.kron(sp.kron(E, Ty), E) + sp.kron(sp.kron(Tz, E), E))


def elasticity27(g, scale=1.0):
    """cfg5: 27-point stencil (diag 26, 26 neighbours -1) (x) M3, 3 dof per node; n = 3 g^3."""
    B = _tridiag(g, 1.0, 1.0, 1.0)
    N = sp.kron(sp.kron(B, B), B, format="csc")                # 27-point neighbourhood incl. self
    S = sp.identity(g ** 3, format="csc") * 27.0 - N           # diag 26, neighbours -1
    M3 = sp.csc_matrix(np.array([[2.0, 0.5, 0.5], [0.5, 2.0, 0.5], [0.5, 0.5, 2.0]]))
    return _clean(sp.kron(S, M3) * scale)


def pivoting_stress(n, density, seed):
    """`sprand(n,n,p)+I`-like unsymmetric matrices that force real pivoting
    (test/test_structunsymm.jl:60-90), with a fixed NumPy seed."""
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=density, random_state=rng, format="csc", data_rvs=rng.random)
    return _clean(a + sp.identity(n, format="csc"))


def rhs_for(a, x=None):
    """b = A x*, x* = 1..n (what `makerhs!` does by default, SpkProblem.jl:408-412)."""
    n = a.shape[0]
    x = np.arange(1, n + 1, dtype=np.float64) if x is None else x
    return a @ x
