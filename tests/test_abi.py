"""The C-ABI library loads and exports every symbol include/ligero_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ligero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ligero_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ligero_b200.h but not exported"


def test_python_binding_covers_header():
    from ligero_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA devices, creating a context must fail loudly."""
    import torch
    from ligero_b200 import Context, LigeroB200Error
    if torch.cuda.is_available():
        return
    try:
        Context(0)
    except LigeroB200Error as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context(0) succeeded without a GPU")


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under ligero_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "ligero_b200")
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|ligero_ref|cref)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{os.path.join(dirpath, f)} references the oracle"


def test_nothing_at_run_time_reads_the_reference_tree():
    """/root/reference does not exist on the GPU box: the product, the GPU tests, the scripts, smoke() and bench.py must
    not name it (the fixture generator and the CPU-only reference-facts test, which skips without it, may)."""
    import glob
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py"), os.path.join(ROOT, "tests", "util.py")]
    files += glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py"))
    for top in ("ligero_b200", "scripts"):
        for dirpath, _, names in os.walk(os.path.join(ROOT, top)):
            files += [os.path.join(dirpath, f) for f in names if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".sh", ".inc"))]
    assert len(files) > 30
    for p in files:
        assert "/root/reference" not in open(p).read(), f"{p} names /root/reference"


def test_header_is_valid_c_and_the_c_host_example_links():
    """include/ligero_b200.h compiled as C11 by gcc, and tests/c/mgpu_prove.c (a C host driving lg_mgpu_prove) and
    tests/c/new_prove_check.c (LigeroCircuit::new / prove / verify from C) linked against the in-tree library; without a GPU the program stops at lg_ctx_create with LG_ERR_CUDA -- no CPU fallback."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "ligero_b200")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "mgpu_prove")
        r = subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                            os.path.join(root, "tests", "c", "mgpu_prove.c"), "-o", exe, "-L", libdir, "-lligero_b200",
                            f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exe2 = os.path.join(tmp, "new_prove_check")
        r = subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                            os.path.join(root, "tests", "c", "new_prove_check.c"), "-o", exe2, "-L", libdir, "-lligero_b200",
                            f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        import torch
        if not torch.cuda.is_available():
            run = subprocess.run([exe, "1", "6"], capture_output=True, text=True)
            assert run.returncode != 0 and "lg_ctx_create" in run.stderr
            run = subprocess.run([exe2, "8"], capture_output=True, text=True)
            assert run.returncode != 0 and "lg_ctx_create" in run.stderr
