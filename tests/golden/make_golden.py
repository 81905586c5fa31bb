#!/usr/bin/env python
"""Regenerate tests/golden/reference_vectors.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's Python wrapper (/root/reference/sparse_matrix_mult/matrix_ops.py), which loads
the reference's shipped libsparse_x86_64.so, calls its public `sparse_matrix_multiply` on the inputs built
by tests/golden/cases.py, and stores inputs (as CSR arrays) and outputs.  The .npz is what travels to the
GPU box; /root/reference never does.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import cases  # noqa: E402

with redirect_stdout(io.StringIO()):          # the reference prints loader chatter at import
    from sparse_matrix_mult import sparse_matrix_multiply as ref_multiply  # noqa: E402


def main():
    store = {}
    names = []
    for name, a, b, kwargs in cases.all_cases():
        with redirect_stdout(io.StringIO()):
            out = ref_multiply(a, b, **kwargs)
        names.append(name)
        if isinstance(out, np.ndarray):
            store[name + "/dense"] = out
        else:
            out = csr_matrix(out)
            store[name + "/indptr"] = out.indptr
            store[name + "/indices"] = out.indices      # first-touch (unsorted) order, as returned
            store[name + "/data"] = out.data
            store[name + "/shape"] = np.array(out.shape)
    store["__names__"] = np.array(names)
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(names)} cases, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
