"""Builds oracle/_ref/libsparse_ref_omp_patched.so: the reference's own src/*.cpp with the repairs of SURVEY.md
Appendix B applied, compiled with setup.py:142-165's flags (-O3 -fPIC -std=c++11 -fopenmp).  TEST INFRASTRUCTURE ONLY.

Today's src/ cannot run its sparse-output path; the shipped binary can, but it is serial.  This
build gives the reference's multi-threaded sparse_nosym / sparse_sym a working body so that bench.py can time the
reference on every host core.  tests/test_oracle.py shows it bit-identical to the shipped binary.

No reference source is copied into the repository: the sources are read where they lie under /root/reference,
edited IN MEMORY by the anchored substitutions below (each asserts how many places it hit), written to a temporary
directory, compiled, and the temporary directory is removed.

Repairs (Appendix B numbering):
  1  src/sparsework.cpp:45,190       the position-marker array must start at -1, not 0
  2  src/sparsework.cpp:84-102,231-249   after the pool grows, the values block has to move to its new offset
  3  src/sparsework.cpp:135-148,286-299  before the final shrink, the values block has to move down to its final offset
  4  src/sparse_sparse_sparse.cpp:149,290  partial results live in one pool: free its base, not three interior pointers
  5  src/sparse_sparse_sparse.cpp:94     USE_OPENMP -> _OPENMP (else every thread runs every partition)
  6  src/sparse_sparse_sparse.cpp:191    `threads` is uninitialised without OpenMP
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "libsparse_ref_omp_patched.so")


def sub(text, pattern, repl, count, what):
    new, n = re.subn(pattern, repl, text, flags=re.S)
    if n != count:
        raise SystemExit(f"build_patched_ref: {what}: expected {count} site(s), found {n} -- the reference changed")
    return new


def patch_sparsework(t):
    # 1: markers start at -1
    t = sub(t, r"\(int\*\)calloc\(\(size_t\)matrixb->cols, sizeof\(int\)\);",
            "(int*)malloc((size_t)matrixb->cols * sizeof(int));\n"
            "    if (workArray != NULL) memset(workArray, -1, (size_t)matrixb->cols * sizeof(int));", 2, "repair 1")
    # 2: remember the old capacity, move the values block after the pointers are re-derived
    t = sub(t, r"estimated_nzmax \*= 2;", "size_t old_cap__ = estimated_nzmax;\n estimated_nzmax *= 2;", 2, "repair 2a")
    t = sub(t, r"(memory_pool = new_memory;.*?estimated_nzmax \* sizeof\(int\)\);)",
            r"\1\n memmove(matrixc->values, memory_pool + (size_t)(local_rows + 1) * sizeof(int) + old_cap__ * sizeof(int),"
            r" (size_t)matrixc->nzmax * sizeof(double));", 2, "repair 2b")
    # 3: move the values block down to where the shrunk pool will expect it, then shrink
    t = sub(t, r"(char\* final_memory = \(char\*\)realloc\(memory_pool, final_size\);)",
            r"{ double* dst__ = (double*)(memory_pool + (size_t)(local_rows + 1) * sizeof(int) + (size_t)matrixc->nzmax * sizeof(int));"
            r" memmove(dst__, matrixc->values, (size_t)matrixc->nzmax * sizeof(double)); matrixc->values = dst__; }\n    \1",
            2, "repair 3")
    return t


def patch_driver(t):
    t = sub(t, r"destroy_sparsemat\(&dimensions\[i\]\);", "free(dimensions[i].rowPtr);", 2, "repair 4")
    t = sub(t, r"#ifdef USE_OPENMP", "#ifdef _OPENMP", 1, "repair 5")
    t = sub(t, r"//threads = 1;", "threads = 1;", 1, "repair 6")
    return t


def main():
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        raise SystemExit(f"build_patched_ref: {src} not found (the GPU box only uses the prebuilt file)")
    tmp = tempfile.mkdtemp(prefix="ref_patched_")
    try:
        files = []
        for name in sorted(os.listdir(src)):
            if not name.endswith(".cpp"):
                continue
            text = open(os.path.join(src, name)).read()
            if name == "sparsework.cpp":
                text = patch_sparsework(text)
            elif name == "sparse_sparse_sparse.cpp":
                text = patch_driver(text)
            path = os.path.join(tmp, name)
            with open(path, "w") as f:
                f.write(text)
            files.append(path)
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O3", "-fPIC", "-std=c++11", "-fopenmp", "-w", f"-I{REF}/include", "-shared",
                               *files, "-o", OUT])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("built", OUT)


if __name__ == "__main__":
    sys.exit(main())
