"""oracle -- CPU checkers for the CUDA path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  sparse_matrix_mult_b200 never does.

  oracle.port  -- ctypes view of liboracle.so (spgemm_oracle.c, the plain-C restatement)
  oracle.ref   -- ctypes view of oracle/_ref/*.so (the reference's own code, built by `make ref`)
"""
