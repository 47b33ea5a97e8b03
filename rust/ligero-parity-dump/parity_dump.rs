//! Parity dump for ligero_b200 (https://github.com/NP-Eng/ligero, `src/ligero/parity_dump.rs`).
//!
//! This file is NOT part of the reference.  `run.sh` (next to it) copies it into a checkout of the
//! reference as a `#[cfg(test)]` child module of `ligero` -- it has to live there: the proof structs'
//! fields are private (`src/ligero/mod.rs:96-144`) and the circuit generators the reference's own tests use
//! are `#[cfg(test)]` items (`src/arithmetic_circuit/tests.rs`).  It runs exactly the reference's end-to-end
//! tests (`src/ligero/tests.rs:144-243` lemniscate / determinant, `365-415` poseidon) with the deterministic
//! test RNG and writes, per circuit, everything ligero_b200's oracle and GPU prover must reproduce bit for
//! bit: shapes, `u_root`, `preenc_u_lc`, both test polynomials, the three sets of opened indices, and the
//! whole proof in the byte layout of `lg_proof_serialize` (include/ligero_b200.h).
//!
//!     DETERMINISTIC_TEST_RNG=1 cargo test --release parity_dump -- --nocapture
//!     -> target/ligero_parity_dump.json   (tests/test_golden.py reads it via LIGERO_PARITY_DUMP=...)
use std::{fmt::Write as _, str::FromStr};

use ark_bn254::Fr;
use ark_crypto_primitives::{
    merkle_tree::{Config, Path},
    sponge::poseidon::PoseidonSponge,
};
use ark_ff::{BigInteger, PrimeField};
use ark_poly::DenseUVPolynomial;
use ark_poly_commit::test_sponge;
use ark_relations::r1cs::ConstraintSystem;
use ark_serialize::CanonicalSerialize;
use itertools::Itertools;

use super::{types::LigeroMTTestParams, LigeroCircuit, LigeroProof};
use crate::{
    arithmetic_circuit::{
        tests::{generate_3_by_3_determinant_circuit, generate_lemniscate_circuit},
        ArithmeticCircuit,
    },
    reader::read_constraint_system,
    DEFAULT_SECURITY_LEVEL,
};

fn hex(bytes: &[u8]) -> String {
    bytes.iter().fold(String::new(), |mut s, b| {
        write!(s, "{b:02x}").unwrap();
        s
    })
}

/// canonical little-endian 32 bytes, as `lg_proof_serialize` writes an `Fr`
fn fr_bytes(x: &Fr) -> Vec<u8> {
    let mut v = x.into_bigint().to_bytes_le();
    v.resize(32, 0);
    v
}

fn fr_hex_be(x: &Fr) -> String {
    let h = hex(&x.into_bigint().to_bytes_be());
    let t = h.trim_start_matches('0');
    format!("0x{}", if t.is_empty() { "0" } else { t })
}

fn put_u64(out: &mut Vec<u8>, v: u64) {
    out.extend_from_slice(&v.to_le_bytes());
}

fn put_frs(out: &mut Vec<u8>, v: &[Fr]) {
    put_u64(out, v.len() as u64);
    v.iter().for_each(|x| out.extend_from_slice(&fr_bytes(x)));
}

/// a digest (`Vec<u8>` for `TestMerkleTreeParams`) as CanonicalSerialize writes it: u64 length + bytes.
/// Going through `serialize_compressed` keeps this honest if the digest type ever changes.
fn put_digest<D: CanonicalSerialize>(out: &mut Vec<u8>, d: &D) {
    d.serialize_compressed(&mut *out).unwrap();
}

fn put_opened<C: Config>(out: &mut Vec<u8>, columns: &[Vec<Fr>], paths: &[Path<C>]) {
    put_u64(out, columns.len() as u64);
    columns.iter().for_each(|c| put_frs(out, c));
    put_u64(out, paths.len() as u64);
    for p in paths {
        put_digest(out, &p.leaf_sibling_hash);
        put_u64(out, p.auth_path.len() as u64);
        p.auth_path.iter().for_each(|d| put_digest(out, d));
        put_u64(out, p.leaf_index as u64);
    }
}

/// the layout of `lg_proof_serialize`:
///   u_root | preenc_u_lc, columns, paths | linear poly coeffs, columns, paths | quadratic poly coeffs, columns, paths
fn wire_bytes<C: Config>(p: &LigeroProof<Fr, C>) -> Vec<u8> {
    let mut out = Vec::new();
    put_digest(&mut out, &p.u_root);
    put_frs(&mut out, &p.interleaved_proof.preenc_u_lc);
    put_opened(&mut out, &p.interleaved_proof.columns, &p.interleaved_proof.paths);
    put_frs(&mut out, p.linear_constraints_proof.polynomial.coeffs());
    put_opened(&mut out, &p.linear_constraints_proof.columns, &p.linear_constraints_proof.paths);
    put_frs(&mut out, p.quadratic_constraints_proof.polynomial.coeffs());
    put_opened(&mut out, &p.quadratic_constraints_proof.columns, &p.quadratic_constraints_proof.paths);
    out
}

fn frs_json(v: &[Fr]) -> String {
    format!("[{}]", v.iter().map(|x| format!("\"{}\"", fr_hex_be(x))).join(","))
}

fn indices_json<C: Config>(paths: &[Path<C>]) -> String {
    format!("[{}]", paths.iter().map(|p| p.leaf_index.to_string()).join(","))
}

fn dump_one(name: &str, circuit: ArithmeticCircuit<Fr>, outputs: Vec<usize>, vars: Vec<(usize, Fr)>) -> String {
    let ligero = LigeroCircuit::new(circuit, outputs, DEFAULT_SECURITY_LEVEL);
    let sponge: PoseidonSponge<Fr> = test_sponge();
    let mt_params = LigeroMTTestParams::new();
    let proof = ligero.prove(vars, &mt_params, &mut sponge.clone());
    let mut root = Vec::new();
    put_digest(&mut root, &proof.u_root);
    let blob = wire_bytes(&proof);
    let entry = format!(
        "\"{name}\":{{\"m\":{},\"k\":{},\"n\":{},\"t\":{},\"u_root\":\"{}\",\"preenc_u_lc\":{},\"linear_poly\":{},\
         \"quadratic_poly\":{},\"interleaved_indices\":{},\"linear_indices\":{},\"quadratic_indices\":{},\
         \"proof_len\":{},\"proof_hex\":\"{}\"}}",
        ligero.m,
        ligero.k,
        ligero.n,
        ligero.t,
        hex(&root[8..]),
        frs_json(&proof.interleaved_proof.preenc_u_lc),
        frs_json(proof.linear_constraints_proof.polynomial.coeffs()),
        frs_json(proof.quadratic_constraints_proof.polynomial.coeffs()),
        indices_json(&proof.interleaved_proof.paths),
        indices_json(&proof.linear_constraints_proof.paths),
        indices_json(&proof.quadratic_constraints_proof.paths),
        blob.len(),
        hex(&blob),
    );
    assert!(ligero.verify(proof, &mt_params, &mut sponge.clone()));
    entry
}

#[test]
fn parity_dump() {
    assert_eq!(
        std::env::var("DETERMINISTIC_TEST_RNG").as_deref(),
        Ok("1"),
        "run with DETERMINISTIC_TEST_RNG=1: test_sponge() draws its round constants from ark_std::test_rng()"
    );
    let mut entries = Vec::new();

    // src/ligero/tests.rs:197-209
    let c = generate_lemniscate_circuit();
    let out = c.last();
    entries.push(dump_one("lemniscate", c, vec![out], vec![(1, Fr::from(8)), (2, Fr::from(4))]));

    // src/ligero/tests.rs:211-243
    let c = generate_3_by_3_determinant_circuit();
    let out = c.last();
    let vals = [2i64, 0, -1, 3, 5, 2, -4, 1, 4, 13];
    let vars = vals.iter().enumerate().map(|(i, v)| (i + 1, Fr::from(*v))).collect_vec();
    entries.push(dump_one("determinant", c, vec![out], vars));

    // src/ligero/tests.rs:365-415 (needs the circom fixtures of the checkout: run from the crate root)
    let cs: ConstraintSystem<Fr> =
        read_constraint_system("circom/poseidon/poseidon.r1cs", "circom/poseidon/poseidon_js/poseidon.wasm");
    let (circuit, outputs) = ArithmeticCircuit::from_constraint_system(&cs);
    let witness: Vec<Fr> =
        serde_json::from_str::<Vec<String>>(&std::fs::read_to_string("circom/poseidon/witness.json").unwrap())
            .unwrap()
            .iter()
            .map(|s| Fr::from_str(s).unwrap())
            .collect();
    let vars = witness.into_iter().enumerate().skip(1).collect_vec();
    entries.push(dump_one("poseidon", circuit, outputs, vars));

    let json = format!("{{\"source\":\"NP-Eng/ligero via rust/ligero-parity-dump\",{}}}\n", entries.join(","));
    let path = std::env::var("LIGERO_PARITY_DUMP").unwrap_or_else(|_| "target/ligero_parity_dump.json".into());
    std::fs::write(&path, json).unwrap();
    println!("wrote {path}");
}
