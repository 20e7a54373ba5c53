#!/usr/bin/env python
"""Generate csrc/taps_generated.inc: the visit loops of spread_rows.cu as inline PTX.

Why generated PTX: the tile accumulators live in statically indexed registers, so a visit must jump
to straight-line code specialised for its x offset.  nvcc lowers a many-way C++ `switch` to a
compare/branch tree (17 issue slots per visit, measured with ncu); PTX `brx.idx` gives one indexed
branch.  The whole loop over a run of staged visits is one asm block: with the loop in C++ and only
the taps in asm, nvcc re-assigned the 32 accumulator pairs between the asm instances of the
pipelined loop and paid ~9 register moves per visit (ncu, round 1).  The FMAs are Blackwell's
packed `fma.rn.f32x2` (SASS FFMA2).

Register layout of a warp's tile (2 grid rows x 16 cells, lane = coil):
    acc[r*16 + c*8 + j]   r = row (0, 1), c = 0 real / 1 imaginary part,
                          j = cell pair: 64-bit register = (cell 2j, cell 2j+1)
A visit is staged in shared memory as a 48-byte packet (read with warp-broadcast loads)
    P_0..P_3 : pair-packed x weights, P_q = (w[2q - par], w[2q + 1 - par]), par = x offset & 1
    s0, s1   : row scales wy[dy] wz[dz], wy[dy + 1] wz[dz]
    idx      : floor(x offset / 2) + W // 2  (>= 0): which cell pairs j = idx - W//2 + q are hit
    s        : sorted point index
and, for the spreader, the point's coil values in a [visit][32 coils] (re, im) array.

spread:  acc[r][c][j] += P_q * (a, a),  a = (re | im of this lane's coil value) * s_r
interp:  kt[s][lane] += sum_r s_r * sum_{q, lanes of the pair} acc[r][c][j] * P_q   (red.global.add.v2.f32)

The loop is software-pipelined by hand over two register sets (A, B): the packet (and value) of
visit k + 1 is loaded before the taps of visit k execute.  See `gen_loop` for the run structure.
"""
import sys

NJ = 8  # cell pairs per row


def cases(W):
    jb0 = W // 2
    return jb0, NJ + jb0  # idx offset, number of cases (jb = -jb0 .. NJ-1)


def load_set(X, off_pk, off_vb, pred, spread):
    """Load visit packet (+ coil value) into register set X; under `pred` (no further visit) only mark it."""
    L = []
    if pred:
        L.append(f"@{pred} mov.u32 i{X}, 0xffffffff;")
    p = f"@!{pred} " if pred else ""
    L += [f"{p}ld.shared.v2.b64 {{P{X}0, P{X}1}}, [pk+{off_pk}];",
          f"{p}ld.shared.v2.b64 {{P{X}2, P{X}3}}, [pk+{off_pk + 16}];",
          f"{p}ld.shared.v4.b32 {{s{X}0, s{X}1, i{X}, n{X}}}, [pk+{off_pk + 32}];"]
    if spread:
        L.append(f"{p}ld.shared.b64 v{X}, [vb+{off_vb}];")
    return L


def taps_spread(W, X, c):
    """acc += P (x) A for tap-kernel case c, visit in register set X (straight-line code)."""
    jb0, _ = cases(W)
    NP = W // 2 + 1
    jb = c - jb0
    L = [f"mov.b64 {{vx, vy}}, v{X};",
         f"mul.f32 t0, vx, s{X}0;", f"mul.f32 t1, vy, s{X}0;",
         f"mul.f32 t2, vx, s{X}1;", f"mul.f32 t3, vy, s{X}1;",
         "mov.b64 A0, {t0, t0};", "mov.b64 A1, {t1, t1};", "mov.b64 A2, {t2, t2};", "mov.b64 A3, {t3, t3};"]
    for q in range(NP):
        j = jb + q
        if j < 0 or j >= NJ:
            continue
        for r in range(2):
            for ri in range(2):
                a = r * 16 + ri * 8 + j
                L.append(f"fma.rn.f32x2 %{a}, P{X}{q}, A{r * 2 + ri}, %{a};")
    return L


def taps_interp(W, X, c):
    """kt[s] += sum of the taps of case c read from the tile registers, visit in register set X."""
    jb0, _ = cases(W)
    NP = W // 2 + 1
    jb = c - jb0
    L = []
    # the four partial sums are independent chains: interleave them so that no instruction waits
    # on the one issued just before it
    first = [True] * 4
    for q in range(NP):
        j = jb + q
        if j < 0 or j >= NJ:
            continue
        for r in range(2):
            for ri in range(2):
                o = r * 2 + ri
                a = r * 16 + ri * 8 + j
                if first[o]:
                    L.append(f"mul.rn.f32x2 S{o}, %{a}, P{X}{q};")
                    first[o] = False
                else:
                    L.append(f"fma.rn.f32x2 S{o}, %{a}, P{X}{q}, S{o};")
    # row scales applied on the packed pairs, then one horizontal add per component
    L += [f"mov.b64 A0, {{s{X}0, s{X}0}};", f"mov.b64 A1, {{s{X}1, s{X}1}};",
          "mul.rn.f32x2 S0, S0, A0;", "mul.rn.f32x2 S1, S1, A0;",
          "fma.rn.f32x2 S0, S2, A1, S0;", "fma.rn.f32x2 S1, S3, A1, S1;",
          "mov.b64 {lo, hi}, S0;", "add.f32 t0, lo, hi;",
          "mov.b64 {lo, hi}, S1;", "add.f32 t1, lo, hi;",
          f"mad.wide.u32 addr, n{X}, 256, ktl;",
          "red.global.add.v2.f32 [addr], {t0, t1};"]
    return L


def gen_loop(W, spread):
    """One asm block that consumes a run of `n` staged visits.

    The stream builder groups the visits of a tile by tap-kernel case, so consecutive visits mostly
    share their case: after one indexed branch (`brx.idx`) the visits of a run execute a two-visit
    loop of straight-line code -- one taken branch per two visits -- and only a change of case goes
    back through the dispatcher.  Two register sets (A, B) alternate: the packet of the next visit
    is loaded before the taps of the current one execute.
    """
    name = f"rows_loop_{'spread' if spread else 'interp'}_w{W}"
    taps = taps_spread if spread else taps_interp
    _, ncase = cases(W)
    step = ["add.u32 pk, pk, 48;"] + (["add.u32 vb, vb, 256;"] if spread else []) + ["add.s32 n, n, -1;"]
    step2 = ["add.u32 pk, pk, 96;"] + (["add.u32 vb, vb, 512;"] if spread else []) + ["add.s32 n, n, -2;"]
    body = ["{",
            ".reg .b64 PA<4>, PB<4>, vA, vB, A<4>, S<4>, addr, ktl;",
            ".reg .b32 sA<2>, sB<2>, iA, iB, nA, nB, vx, vy, lo, hi, t<4>, pk, vb, n;",
            ".reg .pred p1, p2, p3, p4, p5;",
            "mov.u32 pk, %32;", "mov.u32 n, %33;"]
    body.append("mov.u32 vb, %34;" if spread else "mov.u64 ktl, %34;")
    body += load_set("A", 0, 0, None, spread)
    for X, Y in (("A", "B"), ("B", "A")):
        body += [f"D{X}:", "setp.lt.s32 p5, n, 1;", "@p5 bra.uni DONE;",
                 f"tbl{X}: .branchtargets " + ", ".join(f"R{X}{i}" for i in range(ncase)) + ";",
                 f"brx.idx.uni i{X}, tbl{X};"]
        for c in range(ncase):
            body += [f"R{X}{c}:", "setp.lt.s32 p1, n, 2;"]
            body += load_set(Y, 48, 256, "p1", spread)
            body += taps(W, X, c)
            body += [f"setp.ne.u32 p2, i{Y}, {c};", f"@p2 bra.uni X{X}{c};", "setp.lt.s32 p3, n, 3;"]
            body += load_set(X, 96, 512, "p3", spread)
            body += taps(W, Y, c)
            body += step2
            body += [f"setp.eq.u32 p4, i{X}, {c};", f"@p4 bra.uni R{X}{c};", f"bra.uni D{X};",
                     f"X{X}{c}:"]
            body += step
            body += [f"bra.uni D{Y};"]
    body += ["DONE:", "}"]
    L = [f"__device__ __forceinline__ void {name}(",
         "    unsigned long long (&acc)[32], unsigned pk, int n, "
         + ("unsigned vb) {" if spread else "const void* ktl) {"),
         "  asm volatile("]
    L += [f'      "{b}\\n"' for b in body]
    L.append("      : " + ", ".join(f'"+l"(acc[{i}])' for i in range(32)))
    L.append('      : "r"(pk), "r"(n), ' + ('"r"(vb)' if spread else '"l"(ktl)'))
    L.append('      : "memory");')
    L.append("}")
    return "\n".join(L)


def main():
    out = ["// GENERATED by tools/gen_taps.py -- do not edit.  See that script for the layout.",
           "#pragma once", ""]
    for W in (4, 5, 6, 7):
        out += [gen_loop(W, True), "", gen_loop(W, False), ""]
    path = sys.argv[1] if len(sys.argv) > 1 else "mrinufft_b200/csrc/taps_generated.inc"
    open(path, "w").write("\n".join(out))


if __name__ == "__main__":
    main()
