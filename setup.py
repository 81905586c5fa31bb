"""Packaging of the B200 path: `python setup.py build_ext --inplace` / `pip install .`

Counterpart of the reference's setup.py (/root/reference/setup.py:131-178), whose custom build_ext shells out to g++
and drops libsparse_<arch>.so into sparse_matrix_mult/lib/.  Here build_ext runs nvcc (sm_100a only, through
sparse_matrix_mult_b200/Makefile) and drops libspgemm_b200.so into sparse_matrix_mult_b200/lib/, where
MatrixOpsLibrary finds it (matrix_ops.py).  The library name deliberately does not match the reference loader's
`libsparse*.so` pattern (matrix_ops.py:118), so both packages can sit side by side.
"""
import os
import shutil
import subprocess

from setuptools import Extension, find_packages, setup
from setuptools.command.build_ext import build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sparse_matrix_mult_b200")


class BuildCudaLibrary(build_ext):
    """Compiles csrc/*.cu into lib/libspgemm_b200.so with the package's Makefile (nvcc + sm_100a, -lineinfo)."""

    def run(self):
        nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            raise RuntimeError("nvcc not found (set NVCC=...): libspgemm_b200 is CUDA-only, there is no CPU build")
        jobs = str(os.cpu_count() or 4)
        subprocess.check_call(["make", "-C", PKG, "-j", jobs, f"NVCC={nvcc}"])
        lib = os.path.join(PKG, "lib", "libspgemm_b200.so")
        if not os.path.exists(lib):
            raise RuntimeError(f"{lib} was not produced")
        # non-inplace builds: copy the library next to the built package
        if not self.inplace:
            dst = os.path.join(self.build_lib, "sparse_matrix_mult_b200", "lib")
            os.makedirs(dst, exist_ok=True)
            shutil.copy2(lib, dst)


setup(
    name="sparse_matrix_mult_b200",
    version="0.2",
    description="B200 (sm_100a) CUDA implementation of sparse_matrix_mult.sparse_matrix_multiply",
    packages=find_packages(include=["sparse_matrix_mult_b200", "sparse_matrix_mult_b200.*"]),
    include_package_data=True,
    package_data={"sparse_matrix_mult_b200": ["lib/*.so"], "": ["include/*.h"]},
    install_requires=["numpy", "scipy"],
    extras_require={"test": ["pytest"], "multi_process": ["torch"]},
    ext_modules=[Extension("sparse_matrix_mult_b200._native", sources=[])],     # placeholder: triggers build_ext
    cmdclass={"build_ext": BuildCudaLibrary},
)
