#!/usr/bin/env python
"""Generate csrc/taps_generated.inc: the visit loops of spread_rows.cu as inline PTX.

Why generated PTX: the tile accumulators live in statically indexed registers, so a visit must jump
to straight-line code specialised for its x offset.  nvcc lowers a many-way C++ `switch` to a
compare/branch tree (17 issue slots per visit, measured with ncu); PTX `brx.idx` gives one indexed
branch.  The whole loop over a run of staged visits is one asm block: with the loop in C++ and only
the taps in asm, nvcc re-assigned the 32 accumulator pairs between the asm instances of the
pipelined loop and paid ~9 register moves per visit (ncu, round 1).  The FMAs are Blackwell's
packed `fma.rn.f32x2` (SASS FFMA2).

Register layout of a lane's share of the tile (2 grid rows x 16 cells):
    acc[r*16 + c*8 + j]   r = row (0, 1), c = 0 real / 1 imaginary part,
                          j = cell pair: 64-bit register = (cell 2j, cell 2j+1)

Coil classes.  TC = 32: lane = coil, the warp's tile is 2 rows.  TC < 32 (batches of at most TC coils):
lane = (row group g, coil t), g = lane / TC, t = lane % TC; the G = 32 / TC groups hold GY row pairs x
GZ planes of a larger tile, so a point visits fewer tiles and no lane idles for want of coils.

A visit is staged in shared memory as a packet (read with warp-broadcast loads; the packets of a block of 32
visits are stored piece-major, 16-byte piece q of visit k at 512 q + 16 k, see `fa`)
    TC = 32 (48 bytes):  P_0..P_3 | s0, s1, idx, s
    TC < 32 (32 + ESZ):  P_0..P_3 | idx, s | wy[2 GY] | wz[GZ] | padding
    (TC < 32 with room in the entry -- 3-D class 16, the 2-D classes: the 2 G products wy[2 gy + r] wz[gz]
     take the place of the windows and a lane loads its pair s_0, s_1 directly)
        P_q      pair-packed x weights, P_q = (w[2q - par], w[2q + 1 - par]), par = x offset & 1
        s0, s1   row scales wy[dy] wz[dz], wy[dy + 1] wz[dz]; for TC < 32 a lane forms its own pair from
                 the windows: s_r = wy[2 gy + r] * wz[gz]  (zero outside the point's footprint)
        idx      floor(x offset / 2) + W // 2  (>= 0): which cell pairs j = idx - W//2 + q are hit
        s        sorted point index
and, for the spreader, the point's coil values in a [visit][TC coils] (re, im) array.

spread:  acc[r][c][j] += P_q * (a, a),  a = (re | im of this lane's coil value) * s_r
interp:  kt[s][t] += sum_r s_r * sum_{q, lanes of the pair} acc[r][c][j] * P_q   (red.global.add.v2.f32;
         for TC < 32 every lane parks its partial sum in shared memory and the kernel adds the G row groups
         of a coil after the run, see `taps_interp`)

The loop is software-pipelined by hand over two register sets (A, B): the packet (and value) of
visit k + 1 is loaded before the taps of visit k execute.  See `gen_loop` for the run structure.
"""
import sys

NJ = 8  # cell pairs per row
DEFER_INTERP_TAIL = True  # see gen_loop


def obuf_stride(TC):
    """Bytes per visit of the interpolator's partial-sum buffer: 32 lanes x 8 plus a skew that keeps the
    column reads of the reduction (half a warp = 16 / TC visits x TC coils) off each other's banks.
    Mirrors `Cls::OBS` in rows_common.cuh."""
    return 256 if TC >= 16 else 256 + 8 * TC


def class_geometry(dim, TC):
    """(GY row pairs, GZ planes) of the tile of coil class TC: mirrors `Cls<DIM, TC>` in spread_rows.cu."""
    G = 32 // TC
    GZ = (8 if G >= 32 else 4 if G >= 8 else 2 if G >= 2 else 1) if dim == 3 else 1
    return G // GZ, GZ


def entry_bytes(dim, TC):
    if TC == 32:
        return 16
    GY, GZ = class_geometry(dim, TC)
    return (8 + 8 * GY + 4 * GZ + 15) // 16 * 16


def cases(W):
    jb0 = W // 2
    return jb0, NJ + jb0  # idx offset, number of cases (jb = -jb0 .. NJ-1)


class Layout:
    def __init__(self, dim, TC):
        self.dim, self.TC = dim, TC
        self.esz = entry_bytes(dim, TC)
        self.pkt = 32 + self.esz
        self.vstride = TC * 8
        self.generic = TC != 32
        # the entry holds the lanes' row scales themselves (mirrors `Cls::PROD` in rows_common.cuh)
        self.prod = self.generic and 8 + 8 * (32 // TC) <= self.esz
        self.obs = obuf_stride(TC)
        # the row groups' partial sums go to k-space directly, one red per lane and visit like class 32
        # (mirrors `Cls::DIRECT`): with two groups parking and re-reading them costs more than the second red
        self.direct = self.generic and TC == 16
        self.parked = self.generic and not self.direct


def fa(b):
    """Address of packet byte b inside a staged block: the packets are piece-major, 16-byte piece q of visit k at
    512 q + 16 k (mirrors `Cls::fa` in rows_common.cuh)."""
    return (b // 16) * 512 + b % 16


def load_set(lay, X, k, pred, spread):
    """Load visit packet k (relative to pk) (+ coil value) into register set X; under `pred` (no further
    visit) only mark it."""
    off_pk, off_vb = k * 16, k * lay.vstride
    L = []
    if pred:
        L.append(f"@{pred} mov.u32 i{X}, 0xffffffff;")
    p = f"@!{pred} " if pred else ""
    L += [f"{p}ld.shared.v2.b64 {{P{X}0, P{X}1}}, [pk+{off_pk + fa(0)}];",
          f"{p}ld.shared.v2.b64 {{P{X}2, P{X}3}}, [pk+{off_pk + fa(16)}];"]
    if lay.generic:
        L += [f"{p}ld.shared.v2.b32 {{i{X}, n{X}}}, [pk+{off_pk + fa(32)}];",
              f"{p}ld.shared.v2.b32 {{s{X}0, s{X}1}}, [pky+{off_pk}];"]
        if not lay.prod:
            L.append(f"{p}ld.shared.b32 z{X}, [pkz+{off_pk}];")
    else:
        L.append(f"{p}ld.shared.v4.b32 {{s{X}0, s{X}1, i{X}, n{X}}}, [pk+{off_pk + fa(32)}];")
    if spread:
        L.append(f"{p}ld.shared.b64 v{X}, [vb+{off_vb}];")
    return L


def lane_scales(lay, X):
    """TC < 32: the lane's two row scales from its window entries."""
    if not lay.generic or lay.prod:
        return []
    return [f"mul.f32 s{X}0, s{X}0, z{X};", f"mul.f32 s{X}1, s{X}1, z{X};"]


def taps_spread(lay, W, X, c, slot=0):
    """acc += P (x) A for tap-kernel case c, visit in register set X (straight-line code)."""
    jb0, _ = cases(W)
    NP = W // 2 + 1
    jb = c - jb0
    L = lane_scales(lay, X)
    L += [f"mov.b64 {{vx, vy}}, v{X};",
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


def taps_interp_sums(lay, W, X, c, S="S"):
    """Partial sums of a visit: S0..S3 = sum over the tap pairs of case c of tile registers x x-weights
    (S0 / S1: real / imaginary part of row 0, S2 / S3: row 1), visit in register set X."""
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
                    L.append(f"mul.rn.f32x2 {S}{o}, %{a}, P{X}{q};")
                    first[o] = False
                else:
                    L.append(f"fma.rn.f32x2 {S}{o}, %{a}, P{X}{q}, {S}{o};")
    return L


def interp_tail(lay, s0, s1, z, n, slot, S="S"):
    """Row scales applied on the packed pairs, one horizontal add per component, and the result on its way:
    a red to k-space (class 32) or a store into the partial-sum buffer (smaller classes).  `s0, s1, z, n`:
    registers that hold the visit's scales / z window entry / point index."""
    L = []
    if lay.generic and not lay.prod:
        L += [f"mul.f32 q0, {s0}, {z};", f"mul.f32 q1, {s1}, {z};"]
        s0, s1 = "q0", "q1"
    L += [f"mov.b64 A0, {{{s0}, {s0}}};", f"mov.b64 A1, {{{s1}, {s1}}};",
          f"mul.rn.f32x2 {S}0, {S}0, A0;", f"mul.rn.f32x2 {S}1, {S}1, A0;",
          f"fma.rn.f32x2 {S}0, {S}2, A1, {S}0;", f"fma.rn.f32x2 {S}1, {S}3, A1, {S}1;",
          f"mov.b64 {{lo, hi}}, {S}0;", "add.f32 t0, lo, hi;",
          f"mov.b64 {{lo, hi}}, {S}1;", "add.f32 t1, lo, hi;"]
    if lay.parked:
        # the G row groups of a coil each hold a partial sum: parked in shared memory ([visit][lane], rows of
        # `obuf_stride` bytes), summed and added to k-space by the kernel after the run (one compact loop
        # instead of shuffles + a predicated red in every tap-kernel case: the visit loops stay small
        # enough for the instruction cache)
        L += [f"st.shared.v2.f32 [ob+{slot * lay.obs}], {{t0, t1}};"]
    else:
        L += [f"mad.wide.u32 addr, {n}, {lay.TC * 8}, ktl;",
              "red.global.add.v2.f32 [addr], {t0, t1};"]
    return L


def taps_interp(lay, W, X, c, slot):
    """kt[s] += sum of the taps of case c read from the tile registers, visit in register set X."""
    return taps_interp_sums(lay, W, X, c) + interp_tail(lay, f"s{X}0", f"s{X}1", f"z{X}", f"n{X}", slot)


def gen_loop(W, spread, dim=3, TC=32):
    """One asm block that consumes a run of `n` staged visits.

    The stream builder groups the visits of a tile by tap-kernel case, so consecutive visits mostly
    share their case: after one indexed branch (`brx.idx`) the visits of a run execute a two-visit
    loop of straight-line code -- one taken branch per two visits -- and only a change of case goes
    back through the dispatcher.  Two register sets (A, B) alternate: the packet of the next visit
    is loaded before the taps of the current one execute.
    """
    lay = Layout(dim, TC)
    name = f"rows_loop_{'spread' if spread else 'interp'}_w{W}" + (f"_d{dim}c{TC}" if lay.generic else "")
    taps = taps_spread if spread else taps_interp
    _, ncase = cases(W)

    def step(k):
        s = [f"add.u32 pk, pk, {k * 16};"]
        if lay.generic:
            s += [f"add.u32 pky, pky, {k * 16};", f"add.u32 pkz, pkz, {k * 16};"]
        if spread:
            s.append(f"add.u32 vb, vb, {k * lay.vstride};")
        elif lay.parked:
            s.append(f"add.u32 ob, ob, {k * lay.obs};")
        s.append(f"add.s32 n, n, -{k};")
        return s

    body = ["{",
            ".reg .b64 PA<4>, PB<4>, vA, vB, A<4>, S<4>, U<4>, addr, ktl;",
            ".reg .b32 sA<2>, sB<2>, iA, iB, nA, nB, zA, zB, vx, vy, lo, hi, t<4>, q<2>, e<4>, pk, pky, pkz, vb, ob, n;",
            ".reg .pred p1, p2, p3, p4, p5;",
            "mov.u32 pk, %32;", "mov.u32 n, %33;"]
    if spread:
        body.append("mov.u32 vb, %34;")
    elif lay.parked:
        body.append("mov.u32 ob, %34;")
    else:
        body.append("mov.u64 ktl, %34;")
    if lay.generic:
        body += ["add.u32 pky, pk, %35;", "add.u32 pkz, pk, %36;"]
    body += load_set(lay, "A", 0, None, spread)
    # One order of the register sets (A then B).  A run that leaves a case after an odd number of visits
    # moves set B into set A (the already loaded next visit) and goes back through the dispatcher: ~10 moves
    # per change of case instead of a second copy of all the tap kernels with the roles of A and B swapped
    # -- the visit loops are half the size, which the instruction cache (32 KB) rewards.
    moves = [f"mov.b64 PA{q}, PB{q};" for q in range(4)] + ["mov.b32 sA0, sB0;", "mov.b32 sA1, sB1;",
                                                             "mov.b32 iA, iB;", "mov.b32 nA, nB;"]
    if lay.generic and not lay.prod:
        moves.append("mov.b32 zA, zB;")
    if spread:
        moves.append("mov.b64 vA, vB;")
    X, Y = "A", "B"
    body += ["DA:", "setp.lt.s32 p5, n, 1;", "@p5 bra.uni DONE;",
             "tblA: .branchtargets " + ", ".join(f"RA{i}" for i in range(ncase)) + ";",
             "brx.idx.uni iA, tblA;"]
    # (class 32 only: 21.2 -> 20.5 ms at cfg-C; in the smaller classes the extra live registers spill)
    defer = (not spread) and DEFER_INTERP_TAIL and (not lay.generic or lay.TC == 16)
    for c in range(ncase):
        body += [f"RA{c}:", "setp.lt.s32 p1, n, 2;"]
        body += load_set(lay, Y, 1, "p1", spread)
        if defer:
            # The tail of visit A (scales -> horizontal add -> red / store: a chain of dependent fixed-latency
            # instructions) is emitted BEHIND the branch that ends A's basic block, in the block of visit B's
            # taps, so that ptxas can interleave the two; the values it needs are copied out of set A first,
            # because the block also reloads that set for the visit after next.
            body += taps_interp_sums(lay, W, X, c, "S")
            body += [f"setp.ne.u32 p2, i{Y}, {c};", "@p2 bra.uni TA;", "setp.lt.s32 p3, n, 3;",
                     "mov.b32 e0, sA0;", "mov.b32 e1, sA1;", "mov.b32 e2, nA;"] + (["mov.b32 e3, zA;"] if lay.generic and not lay.prod else [])
            body += load_set(lay, X, 2, "p3", spread)
            body += interp_tail(lay, "e0", "e1", "e3", "e2", 0, "S")
            body += taps_interp_sums(lay, W, Y, c, "U")
            body += interp_tail(lay, f"s{Y}0", f"s{Y}1", f"z{Y}", f"n{Y}", 1, "U")
        else:
            body += taps(lay, W, X, c, 0)
            body += [f"setp.ne.u32 p2, i{Y}, {c};", "@p2 bra.uni XA;", "setp.lt.s32 p3, n, 3;"]
            body += load_set(lay, X, 2, "p3", spread)
            body += taps(lay, W, Y, c, 1)
        body += step(2)
        body += [f"setp.eq.u32 p4, i{X}, {c};", f"@p4 bra.uni RA{c};", "bra.uni DA;"]
    if defer:
        body += ["TA:"] + interp_tail(lay, "sA0", "sA1", "zA", "nA", 0, "S")
    body += ["XA:"] + step(1) + moves + ["bra.uni DA;"]
    body += ["DONE:", "}"]
    args = "unsigned vb" if spread else ("unsigned ob" if lay.parked else "const void* ktl")
    if lay.generic:
        args += ", unsigned yo, unsigned zo"
    L = [f"__device__ __forceinline__ void {name}(",
         f"    unsigned long long (&acc)[32], unsigned pk, int n, {args}) {{",
         "  asm volatile("]
    L += [f'      "{b}\\n"' for b in body]
    L.append("      : " + ", ".join(f'"+l"(acc[{i}])' for i in range(32)))
    ins = '"r"(pk), "r"(n), ' + ('"r"(vb)' if spread else ('"r"(ob)' if lay.parked else '"l"(ktl)'))
    if lay.generic:
        ins += ', "r"(yo), "r"(zo)'
    L.append("      : " + ins)
    L.append('      : "memory");')
    L.append("}")
    return "\n".join(L)


CLASSES = {3: (16, 8, 4, 2, 1), 2: (16, 8)}  # coil classes below 32 per dimension


def main():
    """Without arguments: csrc/taps_generated.inc (TC = 32) and csrc/taps_generated_d{dim}c{TC}.inc for the
    smaller coil classes (one file per class: each is compiled in its own translation unit)."""
    base = sys.argv[1] if len(sys.argv) > 1 else "mrinufft_b200/csrc"
    head = ["// GENERATED by tools/gen_taps.py -- do not edit.  See that script for the layout.",
            "#pragma once", ""]
    out = list(head)
    for W in (4, 5, 6, 7):
        out += [gen_loop(W, True), "", gen_loop(W, False), ""]
    open(f"{base}/taps_generated.inc", "w").write("\n".join(out))
    for dim, tcs in CLASSES.items():
        for TC in tcs:
            out = list(head)
            for W in (4, 5, 6, 7):
                out += [gen_loop(W, True, dim, TC), "", gen_loop(W, False, dim, TC), ""]
            open(f"{base}/taps_generated_d{dim}c{TC}.inc", "w").write("\n".join(out))


if __name__ == "__main__":
    main()
