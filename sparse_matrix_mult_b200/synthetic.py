"""Synthetic inputs of the five BASELINE.json configs, with reduced
variants for parity tests.  numpy/scipy only, fixed seeds: the same matrices come out on every box.

  cfg1  A 10,000^2, density 1e-3, A*A -> sparse
  cfg2  A 20,000 x 50,000, density 5e-4, A*A^T -> dense symmetric
  cfg3  H 5,000 x 100,000 density 1e-3, Q banded 100,000^2 (65 diagonals) -> triple product, upper
  cfg4  R-MAT scale 20, edge factor 16, A*A -> sparse   (cfg4r: scale 16, comparable with the reference)
  cfg5  H 40,000 x 1,000,000 density 2e-4, Q banded 1M^2 -> triple product, upper
"""
import numpy as np
import scipy.sparse as sp


def random_csr(rows, cols, density, seed=1234):
    return sp.random(rows, cols, density=density, format='csr', random_state=np.random.default_rng(seed),
                     dtype=np.float64)


def banded_cov(n, half_width=32, length=8.0):
    """Symmetric banded covariance: offsets -half..+half, value exp(-|o|/length)."""
    offs = list(range(-half_width, half_width + 1))
    return sp.diags([np.full(n - abs(o), np.exp(-abs(o) / length)) for o in offs], offs, format='csr',
                    dtype=np.float64)


def rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=20):
    """R-MAT graph adjacency (2^scale vertices, edge_factor * 2^scale edges, duplicates summed), uniform values."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    ne = edge_factor * n
    rows = np.zeros(ne, dtype=np.int64)
    cols = np.zeros(ne, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for _ in range(scale):
        r = rng.random(ne)
        rows = (rows << 1) | (r >= ab)                                # quadrants c, d -> lower half
        cols = (cols << 1) | (((r >= a) & (r < ab)) | (r >= abc))      # quadrants b, d -> right half
    vals = rng.random(ne)
    m = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()      # sums duplicates, sorts columns
    m.indices = m.indices.astype(np.int32)
    m.indptr = m.indptr.astype(np.int32)
    return m


def workload(name):
    """-> dict(kind, a, b, kwargs) where kwargs are those of sparse_matrix_multiply."""
    if name == "cfg1":
        a = random_csr(10_000, 10_000, 1e-3)
        return dict(kind="sparse", a=a, b=a, kwargs=dict(output_format="sparse", symmetric=False))
    if name == "cfg1s":                                     # reduced
        a = random_csr(2_000, 2_000, 4e-3)
        return dict(kind="sparse", a=a, b=a, kwargs=dict(output_format="sparse", symmetric=False))
    if name == "cfg2":
        a = random_csr(20_000, 50_000, 5e-4)
        return dict(kind="dense", a=a, b=sp.csr_matrix(a.T), kwargs=dict(output_format="dense", symmetric=True))
    if name == "cfg2s":
        a = random_csr(1_500, 4_000, 5e-3)
        return dict(kind="dense", a=a, b=sp.csr_matrix(a.T), kwargs=dict(output_format="dense", symmetric=True))
    if name == "cfg3":
        h = random_csr(5_000, 100_000, 1e-3)
        return dict(kind="triple", a=h, b=banded_cov(100_000), kwargs=dict(use_triple_product=True, compute_full_matrix=0))
    if name == "cfg3s":
        h = random_csr(400, 8_000, 5e-3)
        return dict(kind="triple", a=h, b=banded_cov(8_000), kwargs=dict(use_triple_product=True, compute_full_matrix=0))
    if name == "cfg4":
        a = rmat(20)
        return dict(kind="sparse", a=a, b=a, kwargs=dict(output_format="sparse", symmetric=False))
    if name.startswith("cfg4r"):                           # cfg4r (scale 16) or cfg4r<scale>
        scale = int(name[5:]) if len(name) > 5 else 16
        a = rmat(scale)
        return dict(kind="sparse", a=a, b=a, kwargs=dict(output_format="sparse", symmetric=False))
    if name == "cfg5":
        h = random_csr(40_000, 1_000_000, 2e-4)
        return dict(kind="triple", a=h, b=banded_cov(1_000_000), kwargs=dict(use_triple_product=True, compute_full_matrix=0))
    if name == "cfg5s":
        h = random_csr(3_000, 60_000, 2e-3)
        return dict(kind="triple", a=h, b=banded_cov(60_000), kwargs=dict(use_triple_product=True, compute_full_matrix=0))
    raise KeyError(name)


def csr_bytes(x):
    return 12 * x.nnz + 4 * (x.shape[0] + 1)
