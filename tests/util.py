"""Shared helpers for the parity tests (host-side glue only; no arithmetic of the product lives here)."""
from __future__ import annotations

import numpy as np

from ligero_b200 import fr_to_limbs, limbs_to_fr
from oracle import ligero_oracle as O

P = O.P


def csc_right_block(a: "O.SparseMatrix", mk: int):
    """CSC of the right-hand mk-column block of the oracle's A = [[I, -(Px;Py;Pz)], [0, Padd]]."""
    cols = [[] for _ in range(mk)]
    consts = {}
    table = []
    for i, row in enumerate(a.rows):
        for v, j in row:
            if j < 3 * mk:
                assert j == i and v == 1      # identity block
                continue
            if v == 1:
                vid = 0
            elif v == P - 1:
                vid = 1
            else:
                if v not in consts:
                    consts[v] = len(table) + 2
                    table.append(v)
                vid = consts[v]
            cols[j - 3 * mk].append((i, vid))
    col_ptr = np.zeros(mk + 1, dtype=np.uint32)
    row_idx, val_id = [], []
    for c, ent in enumerate(cols):
        for i, vid in ent:
            row_idx.append(i)
            val_id.append(vid)
        col_ptr[c + 1] = len(row_idx)
    return col_ptr, np.array(row_idx, dtype=np.uint32), np.array(val_id, dtype=np.uint32), (fr_to_limbs(table) if table else None)


def flat_limbs(rows):
    return fr_to_limbs([x for r in rows for x in r])


def gpu_prove(ctx, lc: "O.LigeroCircuit", preenc_u, sponge: "O.PoseidonSponge"):
    """Drive the GPU primitives through the phase order of prove_inner (src/ligero/mod.rs:457-578) with the
    host-side Fiat-Shamir sponge, returning an oracle-shaped LigeroProof for comparison."""
    m, k, n, t = lc.m, lc.k, lc.n, lc.t
    cm = ctx.commit(flat_limbs(preenc_u), 4 * m, k, O.RHO_INV)
    cons = ctx.constraints(m * k, *csc_right_block(lc.a, m * k))
    try:
        sponge.absorb_bytes(cm.root)

        def open_cols():
            seed = sponge.squeeze_bytes(32)
            idx = ctx.expand_indices(seed, n, t)
            cols, sib, auth = cm.open(idx)
            columns = [limbs_to_fr(cols[q]) for q in range(len(idx))]
            paths = [O.MerklePath(bytes(sib[q]), [bytes(x) for x in auth[q]], int(idx[q])) for q in range(len(idx))]
            return O.OpenedColumns(columns, paths)

        seed = sponge.squeeze_bytes(32)
        r_int = ctx.expand_fr(seed, 4 * m)
        lc_vec = limbs_to_fr(cm.row_combine(r_int))
        sponge.absorb_field_elements(lc_vec)
        inter = open_cols()
        seed_l = sponge.squeeze_bytes(32)
        lin = limbs_to_fr(cm.linear_test(cons, seed=seed_l))
        sponge.absorb_field_elements(lin)
        lin_open = open_cols()
        seed_q = sponge.squeeze_bytes(32)
        r_q = ctx.expand_fr(seed_q, m)
        quad = limbs_to_fr(cm.quadratic_test(r_q))
        sponge.absorb_field_elements(quad)
        quad_open = open_cols()
        return O.LigeroProof(cm.root, lc_vec, inter, lin, lin_open, quad, quad_open)
    finally:
        cons.free()
        cm.free()


def proofs_equal(a: "O.LigeroProof", b: "O.LigeroProof") -> bool:
    def oc_eq(x, y):
        return x.columns == y.columns and [(p.leaf_sibling_hash, p.auth_path, p.leaf_index) for p in x.paths] == \
            [(p.leaf_sibling_hash, p.auth_path, p.leaf_index) for p in y.paths]
    return (a.u_root == b.u_root and a.preenc_u_lc == b.preenc_u_lc and oc_eq(a.interleaved, b.interleaved)
            and a.linear_poly == b.linear_poly and oc_eq(a.linear, b.linear)
            and a.quadratic_poly == b.quadratic_poly and oc_eq(a.quadratic, b.quadratic))


def repeated_squaring_r1cs(steps: int = 10, x: int = 3):
    """R1CS of circom/repeated_squaring_10.circom (lines 19-29: tmp0 <== x*x; tmp_i <== tmp_{i-1}^2; y <== tmp9), built by
    hand because the reference ships no .r1cs for it: wires [1, y, x, tmp0 .. tmp_{steps-2}] as circom orders
    them with the linear constraint y = tmp_{steps-1} substituted away, and circom's sign convention (-a) * b = -c, as in
    the reference's multiplication.r1cs / cube.r1cs.  Returns (A, B, C rows as [(coeff, wire)], n_wires, witness)."""
    n_wires = 3 + steps - 1
    wire_of = lambda i: 3 + i if i < steps - 1 else 1          # tmp_i (the last one is the output y)
    a, b, c = [], [], []
    prev = 2                                                    # x
    vals = {0: 1, 2: x % P}
    for i in range(steps):
        a.append([(P - 1, prev)])
        b.append([(1, prev)])
        c.append([(P - 1, wire_of(i))])
        vals[wire_of(i)] = vals[prev] * vals[prev] % P
        prev = wire_of(i)
    return a, b, c, n_wires, [vals[w] for w in range(n_wires)]
