#!/usr/bin/env python3
"""Generate ripp_b200/csrc/pvm_programs.cuh: lane-parallel schedules of the Jacobian point formulas.

The "PVM" (pvm.cuh) runs a point operation on a group of 8 lanes: every step is either a MUL step
(each lane one Fq Montgomery product) or a LIN step (each lane one  r = a +- b +- c), operands and
results in the group's shared-memory register file of Fq values.  This script expands the formulas
(dbl-2009-l, madd-2007-bl, over Fq for G1 and over Fq2 = Fq[u]/(u^2+1) for G2) into Fq-level
dataflow, list-schedules them into homogeneous steps and allocates registers.

    python tools/gen_pvm.py > ripp_b200/csrc/pvm_programs.cuh
"""
import sys

LANES = 8
ZERO, JUNK = 0, 1          # reserved registers
FIRST_FREE = 2


class Graph:
    def __init__(self):
        self.nodes = []        # (kind, a, b, c, sb, sc) ; kind in {"in", "mul", "lin"}
        self.pinned = {}       # node -> register (inputs)

    def inp(self, reg):
        self.nodes.append(("in", None, None, None, 0, 0))
        self.pinned[len(self.nodes) - 1] = reg
        return len(self.nodes) - 1

    def mul(self, a, b):
        self.nodes.append(("mul", a, b, None, 0, 0))
        return len(self.nodes) - 1

    def lin(self, a, b=None, sb=0, c=None, sc=0):
        self.nodes.append(("lin", a, b, c, sb, sc))
        return len(self.nodes) - 1


class Fq:
    """Field ops over Fq nodes."""

    def __init__(self, g):
        self.g = g

    def mul(self, a, b):
        return self.g.mul(a, b)

    def sqr(self, a):
        return self.g.mul(a, a)

    def add(self, a, b):
        return self.g.lin(a, b, 0)

    def sub(self, a, b):
        return self.g.lin(a, b, 1)

    def lin3(self, a, b, sb, c, sc):
        return self.g.lin(a, b, sb, c, sc)

    def one_inputs(self, regs):
        return self.g.inp(regs[0])

    width = 1


class Fq2:
    """Field ops over pairs of Fq nodes (u^2 = -1)."""

    def __init__(self, g):
        self.g = g

    def mul(self, a, b):
        g = self.g
        t0, t1 = g.mul(a[0], b[0]), g.mul(a[1], b[1])
        s, u = g.lin(a[0], a[1], 0), g.lin(b[0], b[1], 0)
        t2 = g.mul(s, u)
        return (g.lin(t0, t1, 1), g.lin(t2, t0, 1, t1, 1))

    def sqr(self, a):
        g = self.g
        s, d = g.lin(a[0], a[1], 0), g.lin(a[0], a[1], 1)
        t = g.mul(a[0], a[1])
        return (g.mul(s, d), g.lin(t, t, 0))

    def add(self, a, b):
        return tuple(self.g.lin(x, y, 0) for x, y in zip(a, b))

    def sub(self, a, b):
        return tuple(self.g.lin(x, y, 1) for x, y in zip(a, b))

    def lin3(self, a, b, sb, c, sc):
        return tuple(self.g.lin(x, y, sb, z, sc) for x, y, z in zip(a, b, c))

    def one_inputs(self, regs):
        return (self.g.inp(regs[0]), self.g.inp(regs[1]))

    width = 2


def jac_dbl(F, X, Y, Z):
    A, B = F.sqr(X), F.sqr(Y)
    C = F.sqr(B)
    t2 = F.sqr(F.add(X, B))
    D0 = F.lin3(t2, A, 1, C, 1)
    D = F.add(D0, D0)
    E = F.lin3(A, A, 0, A, 0)
    Fv = F.sqr(E)
    X3 = F.lin3(Fv, D, 1, D, 1)
    YZ = F.mul(Y, Z)
    Z3 = F.add(YZ, YZ)
    m = F.mul(E, F.sub(D, X3))
    C2 = F.add(C, C)
    C4 = F.add(C2, C2)
    Y3 = F.lin3(m, C4, 1, C4, 1)
    return X3, Y3, Z3


def jac_madd(F, X, Y, Z, X2, Y2):
    Z1Z1 = F.sqr(Z)
    U2 = F.mul(X2, Z1Z1)
    S2 = F.mul(Y2, F.mul(Z, Z1Z1))
    H = F.sub(U2, X)
    HH = F.sqr(H)
    I2 = F.add(HH, HH)
    I = F.add(I2, I2)
    J = F.mul(H, I)
    r0 = F.sub(S2, Y)
    r = F.add(r0, r0)
    V = F.mul(X, I)
    rr = F.sqr(r)
    t1 = F.lin3(rr, J, 1, V, 1)
    X3 = F.sub(t1, V)
    m = F.mul(r, F.sub(V, X3))
    YJ = F.mul(Y, J)
    Y3 = F.lin3(m, YJ, 1, YJ, 1)
    Z3 = F.lin3(F.sqr(F.add(Z, H)), Z1Z1, 1, HH, 1)
    return X3, Y3, Z3, H  # H == 0 flags the exceptional cases (P = +-Q); written to a fixed register


def flatten(v):
    return list(v) if isinstance(v, tuple) else [v]


def schedule(g, outputs, out_regs, nreg_limit=120):
    """List-schedule into homogeneous steps; returns (steps, kinds, nregs)."""
    n = len(g.nodes)
    users = [[] for _ in range(n)]
    for i, (k, a, b, c, sb, sc) in enumerate(g.nodes):
        for o in (a, b, c):
            if o is not None:
                users[o].append(i)
    # keep only nodes needed by outputs
    need = set()
    stack = list(outputs)
    while stack:
        x = stack.pop()
        if x in need:
            continue
        need.add(x)
        k, a, b, c, _, _ = g.nodes[x]
        stack.extend(o for o in (a, b, c) if o is not None)
    # priority = longest path to an output (muls weigh 4, lins 1)
    prio = [0] * n
    for i in reversed(range(n)):
        if i not in need:
            continue
        w = 4 if g.nodes[i][0] == "mul" else 1
        prio[i] = w + max([prio[u] for u in users[i] if u in need], default=0)
    done_step = {i: -1 for i in need if g.nodes[i][0] == "in"}
    remaining = [i for i in sorted(need) if g.nodes[i][0] != "in"]
    steps, kinds = [], []
    while remaining:
        for kind in ("lin", "mul"):
            progressed = True
            while progressed:
                s = len(steps)
                ready = [i for i in remaining if g.nodes[i][0] == kind and
                         all(o is None or (o in done_step and done_step[o] < s) for o in g.nodes[i][1:4])]
                if not ready:
                    break
                ready.sort(key=lambda i: -prio[i])
                take = ready[:LANES]
                steps.append(take)
                kinds.append(kind)
                for i in take:
                    done_step[i] = s
                    remaining.remove(i)
                progressed = kind == "lin"  # drain all ready LIN ops before the next MUL step
    # final LIN step(s): move outputs into their pinned registers
    moves = list(zip(outputs, out_regs))
    # register allocation
    last_use = {}
    for s, take in enumerate(steps):
        for i in take:
            for o in g.nodes[i][1:4]:
                if o is not None:
                    last_use[o] = s
    final_step = len(steps)
    for o, _ in moves:
        last_use[o] = final_step
    reg = dict(g.pinned)
    reserved = set(g.pinned.values()) | {ZERO, JUNK} | set(out_regs)
    free = [r for r in range(FIRST_FREE, nreg_limit) if r not in reserved]
    release_at = {}
    maxreg = max(reserved)
    for s, take in enumerate(steps):
        for r in release_at.pop(s, []):
            free.append(r)
        free.sort()
        for i in take:
            reg[i] = free.pop(0)
            maxreg = max(maxreg, reg[i])
        for i in list(reg):
            if i in g.pinned:
                continue
            if last_use.get(i, -1) == s and i not in [o for o, _ in moves]:
                release_at.setdefault(s + 1, []).append(reg[i])
                last_use[i] = -2
    # encode
    words, kind_bits = [], []

    def enc(dst, a, b, c, sb, sc):
        return dst | (a << 7) | (b << 14) | (c << 21) | (sb << 28) | (sc << 29)

    for s, take in enumerate(steps):
        row = []
        for i in take:
            k, a, b, c, sb, sc = g.nodes[i]
            if k == "mul":
                row.append(enc(reg[i], reg[a], reg[b], ZERO, 0, 0))
            else:
                row.append(enc(reg[i], reg[a], reg[b] if b is not None else ZERO, reg[c] if c is not None else ZERO, sb, sc))
        while len(row) < LANES:
            row.append(enc(JUNK, ZERO, ZERO, ZERO, 0, 0))
        words.append(row)
        kind_bits.append(1 if kinds[s] == "mul" else 0)
    for k0 in range(0, len(moves), LANES):
        row = [enc(dst, reg[o], ZERO, ZERO, 0, 0) for o, dst in moves[k0:k0 + LANES]]
        while len(row) < LANES:
            row.append(enc(JUNK, ZERO, ZERO, ZERO, 0, 0))
        words.append(row)
        kind_bits.append(0)
    return words, kind_bits, maxreg + 1


def build(field, op):
    g = Graph()
    F = (Fq if field == "fq" else Fq2)(g)
    w = F.width
    # pinned layout: acc X, Y, Z then addend X2, Y2, then the H flag register
    base = FIRST_FREE
    regs = {name: [base + w * k + j for j in range(w)] for k, name in enumerate(("X", "Y", "Z", "X2", "Y2", "H"))}
    X, Y, Z = (F.one_inputs(regs[nm]) for nm in ("X", "Y", "Z"))
    if op == "dbl":
        outs = jac_dbl(F, X, Y, Z)
        out_regs = regs["X"] + regs["Y"] + regs["Z"]
    else:
        X2, Y2 = F.one_inputs(regs["X2"]), F.one_inputs(regs["Y2"])
        outs = jac_madd(F, X, Y, Z, X2, Y2)
        out_regs = regs["X"] + regs["Y"] + regs["Z"] + regs["H"]
    flat = [x for o in outs for x in flatten(o)]
    # avoid clobbering pinned inputs early: outputs are produced in temporaries and moved at the end
    return schedule(g, flat, out_regs), regs


def main():
    print("// GENERATED by tools/gen_pvm.py -- do not edit.")
    print("// Lane-parallel schedules of the Jacobian point formulas for the PVM interpreter (pvm.cuh).")
    print("// word = dst | a << 7 | b << 14 | c << 21 | sb << 28 | sc << 29 ; MUL step: dst = a * b ;")
    print("// LIN step: dst = a + (sb ? -b : b) + (sc ? -c : c) ; register 0 = zero, 1 = junk.")
    print("#pragma once")
    print("#include <stdint.h>")
    print("namespace ripp { namespace pvm {")
    print("struct Program { const uint32_t* words; const uint8_t* is_mul; int nsteps; int nregs; };")
    maxregs = 0
    for field in ("fq", "fq2"):
        for op in ("dbl", "madd"):
            (words, kinds, nregs), regs = build(field, op)
            maxregs = max(maxregs, nregs)
            name = "%s_%s" % (field.upper(), op.upper())
            flat = ", ".join("0x%08xu" % w for row in words for w in row)
            print("// %s: %d steps (%d MUL), %d registers" % (name, len(words), sum(kinds), nregs))
            print("RIPP_PVM_CONST uint32_t %s_WORDS[%d] = {%s};" % (name, len(words) * LANES, flat))
            print("RIPP_PVM_CONST uint8_t %s_KINDS[%d] = {%s};" % (name, len(kinds), ", ".join(str(k) for k in kinds)))
            print("static constexpr int %s_NSTEPS = %d;" % (name, len(words)))
    w1 = {nm: FIRST_FREE + k for k, nm in enumerate(("X", "Y", "Z", "X2", "Y2", "H"))}
    w2 = {nm: FIRST_FREE + 2 * k for k, nm in enumerate(("X", "Y", "Z", "X2", "Y2", "H"))}
    for nm in w1:
        print("static constexpr int FQ_REG_%s = %d, FQ2_REG_%s = %d;" % (nm, w1[nm], nm, w2[nm]))
    print("static constexpr int NREGS = %d;" % maxregs)
    print("}}  // namespace ripp::pvm")


if __name__ == "__main__":
    main()
