#!/usr/bin/env bash
# Turn "parity unpinned" into a checked fact on any machine with cargo and network access to crates.io / GitHub:
#
#     rust/ligero-parity-dump/run.sh [path-to-a-checkout-of-NP-Eng/ligero]
#
# Copies the reference (default: clones it) into a scratch directory, adds parity_dump.rs as a #[cfg(test)] child
# module of `ligero` (the proof fields and the test circuits are private to the crate, so the dump has to be compiled
# inside it), runs it with the deterministic test RNG and leaves tests/golden/rust_parity_dump.json in THIS repo.
# Then:   python -m pytest tests/test_golden.py -k rust_parity      (oracle vs the Rust reference, no GPU needed)
#         python -m pytest tests/test_gpu_host_driver.py -m gpu -k rust_parity   (GPU prover vs the Rust reference)
# On a mismatch the test reports which of the recalled arkworks conventions (oracle `Formats`, lg_ctx_set_formats) to flip.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
repo="$(cd "$here/../.." && pwd)"
work="$(mktemp -d)"
if [ $# -ge 1 ]; then cp -r "$1" "$work/ligero"; else git clone --depth 1 https://github.com/NP-Eng/ligero "$work/ligero"; fi
cp "$here/parity_dump.rs" "$work/ligero/src/ligero/parity_dump.rs"
grep -q "mod parity_dump" "$work/ligero/src/ligero/mod.rs" || printf '\n#[cfg(test)]\nmod parity_dump;\n' >> "$work/ligero/src/ligero/mod.rs"
cd "$work/ligero"
DETERMINISTIC_TEST_RNG=1 LIGERO_PARITY_DUMP="$repo/tests/golden/rust_parity_dump.json" \
  cargo test --release parity_dump -- --nocapture
echo "wrote $repo/tests/golden/rust_parity_dump.json"
