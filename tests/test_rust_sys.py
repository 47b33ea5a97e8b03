"""The Rust FFI crate (source only: no Rust toolchain in this image) is generated from the C header: it must be current and
declare every entry point the header declares, with the same arity."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_bindings_are_current_and_complete():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_sys.py"), "--check"]).returncode == 0, \
        "rust/ligero-b200-sys/src/lib.rs is stale: run scripts/gen_rust_sys.py"
    header = open(os.path.join(ROOT, "include", "ligero_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    c_protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(lg_[a-z0-9_]+)\s*\(([^;{]*)\)\s*;", header)}
    rs = open(os.path.join(ROOT, "rust", "ligero-b200-sys", "src", "lib.rs")).read()
    rs_protos = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (lg_[a-z0-9_]+)\(([^)]*)\)", rs)}
    assert sorted(c_protos) == sorted(rs_protos)
    for name, params in c_protos.items():
        n_c = 0 if params.strip() in ("", "void") else params.count(",") + 1
        n_rs = 0 if not rs_protos[name].strip() else rs_protos[name].count(",") + 1
        assert n_c == n_rs, name
    # const-correctness spot checks
    assert "pub fn lg_commit(ctx: *mut LgCtx, preenc_u: *const u64, rows: usize, k: usize, rho_inv: u32, root_out: *mut u8, " \
           "out: *mut *mut LgMatrix) -> c_int;" in rs
    assert "labels: *const *const c_char" in rs and "shard_u: *const *mut c_void" in rs
    assert "pub fn lg_last_error(ctx: *const LgCtx) -> *const c_char;" in rs
