"""sparse_matrix_mult_b200 -- B200 (sm_100a) drop-in for sparse_matrix_mult.sparse_matrix_multiply.

Mirrors /root/reference/sparse_matrix_mult/__init__.py:1-3: one public symbol.
"""
from .matrix_ops import sparse_matrix_multiply

__all__ = ['sparse_matrix_multiply']
