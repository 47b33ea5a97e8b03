#!/usr/bin/env python3
"""Generator for ligero_b200/csrc/fr_shoup_body.inc: the straight-line body of

    fr_mul_shoup_raw(t, y, w, p):   t = y*w - qhat*r   in [0, 2r)   for any y < 2^256 * (1 - 2^-29), w < r,
                                    p = floor(w * 2^256 / r)        (precomputed per table entry)

as carry-chain PTX statements (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32.X) with a
plain-C emulation of every instruction emitted next to it, so host unit tests execute the same algorithm.

Three truncated products (8 x 32-bit limbs, W = 2^32):
  H    = sum_{i+j >= 6} y_i p_j W^(i+j)           43 wide MACs   qhat = floor(H / W^8)  (in {Q-1, Q})
  T    = sum_{i+j <= 7} y_i w_j W^(i+j)           28 wide + 8 low MACs            \\ one accumulator,
       + sum_{i+j <= 7} qhat_i nr_j W^(i+j)       28 wide + 8 low MACs            /  mod W^8,  nr = W^8 - r
Each product keeps two accumulators (even / odd positions) so every 64-bit partial product lands on an
aligned limb pair and a row is ONE carry chain per accumulator.

    python scripts/gen_fr_shoup.py   # rewrites the .inc; the file is committed
"""
import os

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
NR = (1 << 256) - R
NR_LIMBS = [(NR >> (32 * i)) & 0xFFFFFFFF for i in range(8)]

out = []          # lines of the .inc
declared = set()


class Chain:
    """one asm statement: a list of (op, dst, srcs)"""

    def __init__(self):
        self.ins = []

    def add(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))


def is_const(s):
    return s[0].isdigit()


def emit_chain(ch, init):
    """init: set of variables holding a value before this statement (updated)."""
    if not ch.ins:
        return
    outs, ins_ = [], []
    written = set()
    for op, dst, srcs in ch.ins:
        for s in srcs:
            if is_const(s):
                continue
            if s in written or s in outs:
                continue          # produced inside this statement, or already an in/out operand
            if s not in ins_:
                ins_.append(s)
        if dst not in outs:
            outs.append(dst)
        written.add(dst)
    # a dst that is also read before being written in this statement is read-write
    rw = set()
    seen_w = set()
    for op, dst, srcs in ch.ins:
        for s in srcs:
            if s in outs and s not in seen_w:
                rw.add(s)
        seen_w.add(dst)
    for v in rw:
        assert v in init, f"{v} read before initialised"
    ins_ = [s for s in ins_ if s not in outs]
    for s in ins_:
        assert s in init, f"{s} read before initialised"
    names = outs + ins_
    idx = {v: i for i, v in enumerate(names)}

    def ref(s):
        return str(int(s.rstrip('u'), 16) if s.startswith('0x') else int(s)) if is_const(s) else f"%{idx[s]}"

    ptx = []
    for op, dst, srcs in ch.ins:
        ptx.append(f"{op}.u32 {ref(dst)}, " + ", ".join(ref(s) for s in srcs) + ";")
    for v in outs:
        if v not in declared:
            declared.add(v)
    cons_out = ", ".join(f'"{"+r" if v in rw else "=r"}"({v})' for v in outs)
    cons_in = ", ".join(f'"r"({v})' for v in ins_)
    out.append("#ifdef __CUDA_ARCH__")
    body = ' "\n      "'.join(" ".join(ptx[i:i + 2]) for i in range(0, len(ptx), 2))
    out.append(f'  asm("{body}"\n      : {cons_out}\n      : {cons_in});')
    out.append("#else")
    out.append("  {")
    out.append("    uint32_t cf = 0; (void)cf;")
    for op, dst, srcs in ch.ins:
        a = [s if not is_const(s) else (s if s.endswith('u') else s + 'u') for s in srcs]
        base = op.split('.')
        name = base[0]
        flags = base[1:]
        if name in ('mul',):
            out.append(f"    {dst} = emu::{flags[0]}({a[0]}, {a[1]});")
            continue
        if name in ('mad', 'madc'):
            half = flags[0]
            cc = 'cc' in flags
            cin = 'cf' if name == 'madc' else '0u'
            out.append(f"    {{ uint32_t c_ = {cin}; {dst} = emu::addc(emu::{half}({a[0]}, {a[1]}), {a[2]}, c_); {'cf = c_;' if cc else ''} }}")
        elif name in ('add', 'addc'):
            cc = 'cc' in flags
            cin = 'cf' if name == 'addc' else '0u'
            out.append(f"    {{ uint32_t c_ = {cin}; {dst} = emu::addc({a[0]}, {a[1]}, c_); {'cf = c_;' if cc else ''} }}")
        elif name in ('sub', 'subc'):
            cc = 'cc' in flags
            cin = 'cf' if name == 'subc' else '0u'
            out.append(f"    {{ uint32_t c_ = {cin}; {dst} = emu::subb({a[0]}, {a[1]}, c_); {'cf = c_;' if cc else ''} }}")
        else:
            raise ValueError(op)
    out.append("  }")
    out.append("#endif")
    for v in outs:
        init.add(v)


def mac_chain(acc, pairs, a, bs, init, top_limit, tail=None):
    """pairs: list of positions p (contiguous, step 2); product a*bs[k] is added into (acc[p], acc[p+1]).
    tail: optional (position, b): low half only, added into acc[position] (ends the chain, carry dropped).
    After the last pair the carry goes into acc[p_last+2] if that limb exists (<= top_limit) and a carry is possible."""
    ch = Chain()
    first = True
    last_pair_fresh = False
    for p, b in zip(pairs, bs):
        lo, hi = f"{acc}{p}", f"{acc}{p + 1}"
        lo_init, hi_init = lo in init, hi in init
        fresh = (not lo_init) and (not hi_init)
        if first and fresh:
            # no carry in, nothing to add: plain product (later pairs of a fresh row stay carry-free too)
            ch.add("mul.lo", lo, a, b)
            ch.add("mul.hi", hi, a, b)
            emit_chain(ch, init)
            ch = Chain()
            last_pair_fresh = True
            continue
        ch.add("mad.lo.cc" if first else "madc.lo.cc", lo, a, b, lo if lo_init else "0")
        ch.add("madc.hi.cc", hi, a, b, hi if hi_init else "0")
        first = False
        last_pair_fresh = fresh
    if tail is not None:
        tp, tb = tail
        t = f"{acc}{tp}"
        if first:
            if t in init:
                ch.add("mad.lo", t, a, tb, t)
            else:
                ch.add("mul.lo", t, a, tb)
        else:
            ch.add("madc.lo", t, a, tb, t if t in init else "0")
    elif not first:
        nxt = pairs[-1] + 2
        if nxt <= top_limit and not last_pair_fresh:
            t = f"{acc}{nxt}"
            ch.add("addc", t, t if t in init else "0", "0")
        else:
            # drop the carry flag: rewrite the last instruction without .cc
            op, dst, srcs = ch.ins[-1]
            ch.ins[-1] = (op.replace(".cc", ""), dst, srcs)
    emit_chain(ch, init)


def main():
    init = set()
    for i in range(8):
        init.update({f"y{i}", f"w{i}", f"p{i}"})
    out.append("// GENERATED by scripts/gen_fr_shoup.py -- do not edit.  Body of fr_mul_shoup_raw (see fr.cuh).")
    out.append("// inputs: y0..y7 (any value < 2^256(1-2^-29)), w0..w7 (w < r), p0..p7 (floor(w 2^256 / r)); output t0..t7")
    names = [f"he{i}" for i in range(6, 16)] + [f"ho{i}" for i in range(7, 16)] + [f"q{i}" for i in range(8)] + \
            [f"e{i}" for i in range(8)] + [f"o{i}" for i in range(1, 8)] + ["hx"]
    out.append("  uint32_t " + ", ".join(names) + ";")
    # ---- H = sum_{i+j>=6} y_i p_j W^(i+j): accumulators he (even positions) / ho (odd positions) ----
    for i in range(8):
        prods = [(j, i + j) for j in range(max(0, 6 - i), 8)]
        ev = [(j, p) for j, p in prods if p % 2 == 0]
        od = [(j, p) for j, p in prods if p % 2 == 1]
        mac_chain("he", [p for _, p in ev], f"y{i}", [f"p{j}" for j, _ in ev], init, 15)
        mac_chain("ho", [p for _, p in od], f"y{i}", [f"p{j}" for j, _ in od], init, 15)
    # qhat = (he + ho) >> 256 : limb 7 only feeds the carry
    ch = Chain()
    ch.add("add.cc", "hx", "he7", "ho7")
    for i in range(8):
        a, b = f"he{8 + i}", f"ho{8 + i}"
        a = a if a in init else "0"
        b = b if b in init else "0"
        ch.add("addc.cc" if i < 7 else "addc", f"q{i}", a, b)
    emit_chain(ch, init)
    split = len(out)
    # ---- T = (y*w + qhat*nr) mod W^8: accumulators e (pairs 0,2,4,6) / o (pairs 1,3,5 and limb 7) ----
    for (a_name, b_of) in (("y", lambda j: f"w{j}"), ("q", lambda j: "0x%08xu" % NR_LIMBS[j])):
        for i in range(8):
            prods = [(j, i + j) for j in range(0, 8 - i)]
            ev = [(j, p) for j, p in prods if p % 2 == 0]
            od = [(j, p) for j, p in prods if p % 2 == 1 and p < 7]
            tl = [(j, p) for j, p in prods if p == 7]
            mac_chain("e", [p for _, p in ev], f"{a_name}{i}", [b_of(j) for j, _ in ev], init, 7)
            mac_chain("o", [p for _, p in od], f"{a_name}{i}", [b_of(j) for j, _ in od], init, 7,
                      tail=(7, b_of(tl[0][0])) if tl else None)
    ch = Chain()
    for i in range(1, 8):
        op = "add.cc" if i == 1 else ("addc.cc" if i < 7 else "addc")
        ch.add(op, f"e{i}", f"e{i}", f"o{i}")
    emit_chain(ch, init)
    # two files: the quotient part needs only (y, p), the remainder part only (y, w, qhat) -- a caller may load the
    # two halves of a table entry separately to keep fewer registers live
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ligero_b200", "csrc")
    for name, lines in (("fr_shoup_body_q.inc", out[:split]), ("fr_shoup_body_t.inc", [out[0]] + out[split:])):
        path = os.path.join(base, name)
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", os.path.normpath(path), len(lines), "lines")


if __name__ == "__main__":
    main()
