#!/usr/bin/env python
"""Generate the inline-PTX "tile interpreter" inner loops of caffe_escoin_b200/csrc/tile_kernel.cuh.

Why generated PTX: the hot loop of the direct sparse convolution must keep an OT x TY x TX block of output
accumulators in REGISTERS and apply each nonzero weight (oc, ic, kh, kw) to it with the input patch also in
registers (one shared-memory load per several FMAs instead of one per FMA, which is what caps the reference's
sconv_shm kernel).  Which accumulators a nonzero touches is data dependent, and registers cannot be indexed
dynamically, so the kernel is a threaded-code interpreter: every (oc_local, kh, kw) combination has its own
straight-line handler (TY*TX FMAs on fixed registers) and the pruned weights are compiled, at plan time, into
a byte-code stream of {weight, handler id} records that each warp walks with a `brx.idx` indirect branch.
CUDA C++ cannot express this (a `switch` is lowered to a compare tree), hence one generated asm block per
tile shape.

PAIR = 2 variants process two images per lane with the packed `fma.rn.f32x2` (FFMA2, new on sm_100): the two
images are interleaved as float2 in shared memory, so a 128-bit shared load yields two register pairs and one
issue slot does two FMAs -- the dispatch overhead then hides under the FMA pipe instead of competing with it.

Usage: python tools/gen_interp.py   (writes caffe_escoin_b200/csrc/generated/interp_v<k>.inc + variant_list.inc)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUTDIR = os.path.join(ROOT, "caffe_escoin_b200", "csrc", "generated")

# (OT, TY, TX, KH, KW, S, PAIR, NCW, NLW): channel block, tile, kernel taps, stride, images per lane,
# compute warps, loader warps.  8 warps/CTA (2 per SM sub-partition) allow 255 registers; 12 allow 168.
VARIANTS = [
    # the first-generation indirect-branch interpreter, kept for comparison (winners of the first B200 sweep)
    (2, 7, 4, 3, 3, 1, 1, 12, 4, 152),
    (4, 4, 4, 3, 3, 1, 2, 8, 4, 232),
    (4, 4, 4, 5, 5, 1, 1, 10, 2),
]


# ---- third generation ("sieve"): no indirect branch at all.  The handlers sit in canonical (oc_local, kh, kw)
# order and every input channel of a channel block carries a bit mask of the handlers that have a nonzero; the
# warp falls through the handler chain and skips the absent ones with warp-uniform *direct* forward branches
# (whole channel / whole kernel row first, then single taps).  Weights are a plain fp32 stream consumed in order.
# Same tuple layout as VARIANTS.
SIEVE = [
    # Trailing letter = patch-load plan the variant is compiled for: "a" aligned body (rows staged by TMA),
    # "b" patch aligned (rows staged by the cp.async loader)
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "a"),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "b"),
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "a"),
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (2, 7, 4, 3, 3, 1, 1, 12, 4, 152, "a"),
    (2, 7, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (6, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (4, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (8, 4, 4, 3, 3, 1, 1, 8, 4, 232, "b"),
    # 5x5 stride 1
    (4, 7, 4, 5, 5, 1, 1, 8, 4, 232, "a"),
    (4, 7, 4, 5, 5, 1, 1, 8, 4, 232, "b"),
    (4, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b"),
    (6, 4, 4, 5, 5, 1, 1, 8, 4, 232, "b"),
    # 1x1
    (8, 2, 4, 1, 1, 1, 1, 16, 4, 104, "b"),
    (6, 7, 4, 1, 1, 1, 1, 8, 4, 232, "b"),
    # 3x3 stride 2
    (4, 4, 4, 3, 3, 2, 1, 8, 4, 232, "b"),
    (8, 2, 4, 3, 3, 2, 1, 12, 4, 152, "b"),
]
# rolling-row-predicate editions (appended after ROWS so earlier ids stay put)
SIEVE_R = [
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "a", 1),
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "b", 1),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "a", 1),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "b", 1),
    # TMA rows, halo columns by warp shuffle; weights prefetched one tap ahead (one MOV less per tap, +1.5 % here)
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "s", 1, 1),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "s", 1, 1),
    (4, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b", 1),
    (3, 4, 4, 3, 3, 1, 1, 16, 4, 104, "b", 1),   # four warps per scheduler; 384 channels = 8 exact groups of 16 x 3
    (6, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b", 1),
    (5, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b", 1),   # 512 channels / 5 -> 9 channel-block groups: 288 units = 2 full waves of 148
    (4, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b", 1),
    # 5x5 with fewer output channels per lane: 100 handlers of o4 are ~35 KB of code and the instruction cache thrashes
    # (no_instruction = 4 stall cycles per issue on AlexNet conv2); o2 / o3 halve it at the price of more patch loads
    (2, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b", 1),
    (3, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b", 1),
    # stride 2
    (4, 4, 4, 3, 3, 2, 1, 8, 4, 232, "b", 1),
    (4, 2, 4, 3, 3, 2, 1, 12, 4, 152, "b", 1),
]


def load_plans(PC, PAIR, KW):
    per_vec = 4 // PAIR
    PADL = (KW - 1) // 2
    loads = []
    pos = 0
    while pos < PC:
        rem = PC - pos
        if PAIR == 1 and rem <= 2:
            loads.append((pos, 2, pos * 4)); pos += 2
        elif PAIR == 2 and rem == 1:
            loads.append((pos, 1, pos * 8)); pos += 1
        else:
            loads.append((pos, per_vec, pos * 4 * PAIR)); pos += per_vec
    loads_a = []
    if PAIR == 1:
        c = -PADL
        while c < PC - PADL:
            rem = PC - PADL - c
            n = 4 if (c % 4 == 0 and rem >= 4) else 2 if (c % 2 == 0 and rem >= 2) else 1
            loads_a.append((c + PADL, n, c * 4))
            c += n
    return loads, loads_a, pos, PADL


def emit_patch_loads(a, PLAN, PR, XW, TX, PADL, loads, loads_a, edge_operand=None):
    """Patch rows -> registers.  PLAN "a": aligned body + scalar halo loads (TMA rows); "b": patch aligned (cp.async rows);
    "s": TMA rows, but the halo columns come from the neighbouring lanes' bodies by warp shuffle (a 4-byte load with a
    16-byte lane stride is a 4-way bank conflict: 12 LSU wavefronts per row against 4 + 2 here); the lanes of a tile
    row are consecutive lanes of one warp."""
    if PLAN != "s":
        plan = loads_a if PLAN == "a" else loads
        for r in range(PR):
            for (po, cnt, byte) in plan:
                b = r * XW + po
                sgn = "+%d" % byte
                if cnt == 4:
                    a("ld.shared.v4.f32 {x%d, x%d, x%d, x%d}, [ad%d%s];" % (b, b + 1, b + 2, b + 3, r, sgn))
                elif cnt == 2:
                    a("ld.shared.v2.f32 {x%d, x%d}, [ad%d%s];" % (b, b + 1, r, sgn))
                else:
                    a("ld.shared.f32 x%d, [ad%d%s];" % (b, r, sgn))
        return
    assert TX % 4 == 0 and PADL <= TX
    for r in range(PR):
        for q in range(TX // 4):
            b = r * XW + PADL + 4 * q
            a("ld.shared.v4.f32 {x%d, x%d, x%d, x%d}, [ad%d+%d];" % (b, b + 1, b + 2, b + 3, r, 16 * q))
    # source lanes come with the lane word: the neighbours, or -- for the first / last tile of a row -- a spare lane
    # parked on the row's TMA zero-filled halo block, so no select is needed
    for r in range(PR):
        for j in range(PADL):
            a("mov.b32 hb, x%d;" % (r * XW + TX + j))                 # left neighbour's body column TX-PADL+j
            a("shfl.sync.idx.b32 hb, hb, sl, 0x1f, 0xffffffff;")
            a("mov.b32 x%d, hb;" % (r * XW + j))
            a("mov.b32 hb, x%d;" % (r * XW + PADL + j))               # right neighbour's body column j
            a("shfl.sync.idx.b32 hb, hb, sr, 0x1f, 0xffffffff;")
            a("mov.b32 x%d, hb;" % (r * XW + PADL + TX + j))


def emit_edge_preds(a, operand):
    """lane word >> 20: bits 0-4 = lane holding the left neighbour's body, bits 5-9 = the right neighbour's"""
    a(".reg .b32 hb, sl, sr;")
    a("and.b32 sl, %%%d, 31;" % operand)
    a("shr.u32 sr, %%%d, 5;" % operand)
    a("and.b32 sr, sr, 31;")


def sieve_layout(OT, KH, KW):
    """bit position of handler (o, kh, kw): whole output channels per 32-bit mask word"""
    OPW = max(1, 32 // (KH * KW))
    NW = (OT + OPW - 1) // OPW
    return OPW, NW


def gen_sieve(vid, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS=0, PLAN="b", ROLL=0, PF=2):
    assert PAIR == 1
    NACC = OT * TY * TX
    PR = (TY - 1) * S + KH
    PC = (TX - 1) * S + KW
    loads, loads_a, XW, PADL = load_plans(PC, PAIR, KW)
    NX = PR * XW
    NC = OT * KH * KW
    NR = OT * KH                      # kernel rows (oc_local, kh)
    OPW, NW = sieve_layout(OT, KH, KW)
    HB = 4 * (1 + NW)                 # header bytes: {plane offset | END, mask words}
    NOPS = NACC
    L = []
    a = L.append
    a("{")
    a(".reg .f32 x<%d>, w, wn, wn2;" % NX)
    a(".reg .b32 pc, pw, off, noff, t, t2, ad<%d>, m<%d>, nm<%d>;" % (PR, NW, NW))
    a(".reg .pred p, pr<3>, pt<%d>;" % KW)
    if PLAN == "s":
        emit_edge_preds(a, NOPS + 3)
    # Stream (4-byte words): H_0 | H_1 W_0.. | H_2 W_1.. | ... | H_END W_(n-1).. ; H = {plane byte offset, masks};
    # the header of step i+1 precedes the weights of step i, so it is fetched a whole step ahead.
    # pc walks the headers and stays provably warp-uniform (advanced by the popcount of the masks), so the masks end
    # up in uniform registers and every skip is a BRA.U.  pw walks the weights; it is pc + (lane_base >> 31), i.e. the
    # same address, but not provably uniform: otherwise ptxas routes every weight LDS -> R2UR -> FFMA(UR) and the
    # conversion sits on the critical path of each tap.  Weights are prefetched two taps ahead (w <- wn <- wn2 <- LDS).
    a("mov.u32 pc, %%%d;" % NOPS)
    a("shr.u32 pw, %%%d, 31;" % (NOPS + 1))
    a("add.u32 pw, pw, pc;")
    a("ld.shared.b32 off, [pc];")
    for k in range(NW):
        a("ld.shared.b32 m%d, [pc+%d];" % (k, 4 + 4 * k))
    # every lane holds the same header; the broadcast shuffles only tell ptxas so when it cannot prove it itself
    a("shfl.sync.idx.b32 off, off, 0, 0x1f, 0xffffffff;")
    for k in range(NW):
        a("shfl.sync.idx.b32 m%d, m%d, 0, 0x1f, 0xffffffff;" % (k, k))
    a("add.u32 pc, pc, %d;" % HB)
    a("add.u32 pw, pw, %d;" % HB)
    a("SLOOP:")
    a("setp.eq.u32 p, off, 0xffffffff;")
    a("@p bra.uni SDONE;")
    a("ld.shared.b32 noff, [pc];")
    for k in range(NW):
        a("ld.shared.b32 nm%d, [pc+%d];" % (k, 4 + 4 * k))
    a("ld.shared.f32 wn, [pw+%d];" % HB)
    if PF == 2:
        a("ld.shared.f32 wn2, [pw+%d];" % (HB + 4))
    a("add.u32 pw, pw, %d;" % HB)      # pw -> the weight held in wn
    # pc -> header of step i+2 = past this step's weights
    a("popc.b32 t, m0;")
    for k in range(1, NW):
        a("popc.b32 t2, m%d;" % k)
        a("add.u32 t, t, t2;")
    a("shl.b32 t, t, 2;")
    a("add.u32 pc, pc, t;")
    a("add.u32 pc, pc, %d;" % HB)
    a("shfl.sync.idx.b32 noff, noff, 0, 0x1f, 0xffffffff;")
    for k in range(NW):
        a("shfl.sync.idx.b32 nm%d, nm%d, 0, 0x1f, 0xffffffff;" % (k, k))

    def rowmask(r):
        o, kh = r // KH, r % KH
        return o // OPW, ((1 << KW) - 1) << ((o % OPW) * KH * KW + kh * KW)

    def set_row_pred(r):
        wd, msk = rowmask(r)
        a("and.b32 t, m%d, 0x%x;" % (wd, msk))
        a("setp.ne.u32 pr%d, t, 0;" % (r % 3))
    if ROLL:
        # rolling row predicates, two rows ahead of their branch (a predicate consumed right after it is produced
        # stalls the branch for the uniform-datapath latency)
        set_row_pred(0)
        if NR > 1:
            set_row_pred(1)
    a("add.u32 ad0, %%%d, off;" % (NOPS + 1))
    for r in range(1, PR):
        a("add.u32 ad%d, ad%d, %%%d;" % (r, r - 1, NOPS + 2))
    emit_patch_loads(a, PLAN, PR, XW, TX, PADL, loads, loads_a)

    def emit_taps(o, kh, single_test_done):
        wd, ob = o // OPW, (o % OPW) * KH * KW
        if not single_test_done:
            for kw in range(KW):
                a("and.b32 t, m%d, 0x%x;" % (wd, 1 << (ob + kh * KW + kw)))
                a("setp.ne.u32 pt%d, t, 0;" % kw)
        for kw in range(KW):
            c = (o * KH + kh) * KW + kw
            if not single_test_done:
                a("@!pt%d bra.uni SH%dE;" % (kw, c))
            a("mov.f32 w, wn;")
            if PF == 2:
                a("mov.f32 wn, wn2;")
                a("ld.shared.f32 wn2, [pw+8];")
            else:
                a("ld.shared.f32 wn, [pw+4];")
            a("add.u32 pw, pw, 4;")
            for ty in range(TY):
                for tx in range(TX):
                    acc = (o * TY + ty) * TX + tx
                    xi = (ty * S + kh) * XW + tx * S + kw
                    a("fma.rn.f32 %%%d, w, x%d, %%%d;" % (acc, xi, acc))
            a("SH%dE:" % c)
    if ROLL:
        for r in range(NR):
            a("SRR%d:" % r)
            if r + 2 < NR:
                set_row_pred(r + 2)
            a("@!pr%d bra.uni SRR%d;" % (r % 3, r + 1))
            emit_taps(r // KH, r % KH, KW == 1)
        a("SRR%d:" % NR)
    else:
        for o in range(OT):
            wd, ob = o // OPW, (o % OPW) * KH * KW
            if KH * KW > 1:
                a("and.b32 t, m%d, 0x%x;" % (wd, ((1 << (KH * KW)) - 1) << ob))
                a("setp.eq.u32 p, t, 0;")
                a("@p bra.uni SO%dE;" % o)
            for kh in range(KH):
                if KW > 1 and KH > 1:
                    a("and.b32 t, m%d, 0x%x;" % (wd, ((1 << KW) - 1) << (ob + kh * KW)))
                    a("setp.eq.u32 p, t, 0;")
                    a("@p bra.uni SR%d_%dE;" % (o, kh))
                emit_taps(o, kh, False)
                if KW > 1 and KH > 1:
                    a("SR%d_%dE:" % (o, kh))
            if KH * KW > 1:
                a("SO%dE:" % o)
    a("mov.b32 off, noff;")
    for k in range(NW):
        a("mov.b32 m%d, nm%d;" % (k, k))
    a("bra.uni SLOOP;")
    a("SDONE:")
    a("}")
    body = "\n".join('      "%s\\n\\t"' % s for s in L)
    ops_out = ", ".join('"+f"(acc[%d])' % i for i in range(NOPS))
    name = "s%s%s%s_o%d_y%d_x%d_k%dx%d_s%d_w%d%s" % (PLAN, "r" if ROLL else "", "1" if PF == 1 else "", OT, TY, TX, KH, KW, S, NCW,
                                                  "_r%d" % CREGS if CREGS else "")
    src = []
    src.append("// ---- sieve variant %s: %d accumulator registers, %d patch registers, %d handlers, %d mask words ----"
               % (name, NACC, NX, NC, NW))
    src.append("template <> struct Interp<%d> {" % vid)
    src.append("  static constexpr int OT = %d, TY = %d, TX = %d, KH = %d, KW = %d, S = %d, PAIR = %d;"
               % (OT, TY, TX, KH, KW, S, PAIR))
    src.append("  static constexpr int NACC = %d, NC = %d, PR = %d, PC = %d, XW = %d, PADL = %d;" % (NOPS, NC, PR, PC, XW, PADL))
    NTW = NCW + (4 if CREGS else NLW)
    src.append("  static constexpr int NCW = %d, NLW = %d, NTW = %d, CREGS = %d, MODE = %d, SHFL = %d;" % (NCW, NLW, NTW, CREGS, 2 if PLAN == "b" else 1, 1 if PLAN == "s" else 0))
    src.append('  static constexpr const char *name() { return "sconv_tile_%s"; }' % name)
    src.append("  __device__ __forceinline__ static void run(float (&acc)[%d], unsigned prog, unsigned lane_base," % NOPS)
    src.append("                                             unsigned pitch_bytes, unsigned edge) {")
    src.append("    asm volatile(")
    src.append(body)
    src.append("      : %s" % ops_out)
    src.append('      : "r"(prog), "r"(lane_base), "r"(pitch_bytes), "r"(edge)')
    src.append('      : "memory");')
    src.append("  }")
    src.append("};")
    return "\n".join(src)


# ---- backward weight through the same machinery ("W" variants).  The lane keeps its OT x TY x TX tile of TOP DIFF in
# registers (read-only operands) and walks the forward sieve stream of the block: for every nonzero tap it forms
# sum_t dy[o][t] * x[t + (kh, kw)] over its tile (four independent FMA chains) and stores the partial to the warp's
# scratch row P[tap ordinal][lane] in shared memory.  The caller then sums each row over the lanes and adds it to the
# weight gradient (one atomic per tap, unit and chunk).  Weights in the stream are skipped (pc advances by popcount).
# tuple: (OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS, PLAN)
BWDW = [
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "s"),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "s"),
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "a"),
    (3, 7, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "a"),
    (4, 7, 4, 3, 3, 1, 1, 8, 4, 232, "b"),
    (4, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (6, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (5, 4, 4, 3, 3, 1, 1, 12, 4, 152, "b"),
    (4, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b"),
    (8, 2, 4, 1, 1, 1, 1, 16, 4, 104, "b"),
    (2, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b"),
    (3, 4, 4, 5, 5, 1, 1, 12, 4, 152, "b"),
]


def gen_bwdw(vid, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS, PLAN):
    assert PAIR == 1
    NACC = OT * TY * TX
    PR = (TY - 1) * S + KH
    PC = (TX - 1) * S + KW
    loads, loads_a, XW, PADL = load_plans(PC, PAIR, KW)
    NX = PR * XW
    NC = OT * KH * KW
    NR = OT * KH
    OPW, NW = sieve_layout(OT, KH, KW)
    HB = 4 * (1 + NW)
    NOPS = NACC
    L = []
    a = L.append
    a("{")
    a(".reg .f32 x<%d>, s<4>, r<2>;" % NX)
    a(".reg .b32 pc, pp, off, noff, t, t2, ad<%d>, m<%d>, nm<%d>;" % (PR, NW, NW))
    a(".reg .pred p, pr<3>, pt<%d>;" % KW)
    if PLAN == "s":
        emit_edge_preds(a, 1 + NOPS + 4)
    # operands: %0 = taps executed (out), %1.. = dy tile, then prog, lane_base, pitch_bytes, scratch (lane's column)
    I0 = 1
    a("mov.u32 pc, %%%d;" % (I0 + NOPS))
    # the combine (3 adds + store) of a tap is deferred to the head of the NEXT executed tap, where it overlaps that
    # tap's FMAs instead of stalling on its own dependent adds; row 0 of the scratch is a dummy for the first "pending"
    a("mov.u32 pp, %%%d;" % (I0 + NOPS + 3))
    for k in range(4):
        a("mov.f32 s%d, 0f00000000;" % k)
    a("ld.shared.b32 off, [pc];")
    for k in range(NW):
        a("ld.shared.b32 m%d, [pc+%d];" % (k, 4 + 4 * k))
    a("shfl.sync.idx.b32 off, off, 0, 0x1f, 0xffffffff;")
    for k in range(NW):
        a("shfl.sync.idx.b32 m%d, m%d, 0, 0x1f, 0xffffffff;" % (k, k))
    a("add.u32 pc, pc, %d;" % HB)
    a("WLOOP:")
    a("setp.eq.u32 p, off, 0xffffffff;")
    a("@p bra.uni WDONE;")
    a("ld.shared.b32 noff, [pc];")
    for k in range(NW):
        a("ld.shared.b32 nm%d, [pc+%d];" % (k, 4 + 4 * k))
    a("popc.b32 t, m0;")
    for k in range(1, NW):
        a("popc.b32 t2, m%d;" % k)
        a("add.u32 t, t, t2;")
    a("shl.b32 t, t, 2;")
    a("add.u32 pc, pc, t;")
    a("add.u32 pc, pc, %d;" % HB)
    a("shfl.sync.idx.b32 noff, noff, 0, 0x1f, 0xffffffff;")
    for k in range(NW):
        a("shfl.sync.idx.b32 nm%d, nm%d, 0, 0x1f, 0xffffffff;" % (k, k))

    def rowmask(r):
        o, kh = r // KH, r % KH
        return o // OPW, ((1 << KW) - 1) << ((o % OPW) * KH * KW + kh * KW)

    def set_row_pred(r):
        wd, msk = rowmask(r)
        a("and.b32 t, m%d, 0x%x;" % (wd, msk))
        a("setp.ne.u32 pr%d, t, 0;" % (r % 3))
    set_row_pred(0)
    if NR > 1:
        set_row_pred(1)
    a("add.u32 ad0, %%%d, off;" % (I0 + NOPS + 1))
    for r in range(1, PR):
        a("add.u32 ad%d, ad%d, %%%d;" % (r, r - 1, I0 + NOPS + 2))
    emit_patch_loads(a, PLAN, PR, XW, TX, PADL, loads, loads_a)
    for r in range(NR):
        o, kh = r // KH, r % KH
        wd, ob = o // OPW, (o % OPW) * KH * KW
        a("WRR%d:" % r)
        if r + 2 < NR:
            set_row_pred(r + 2)
        a("@!pr%d bra.uni WRR%d;" % (r % 3, r + 1))
        if KW > 1:
            for kw in range(KW):
                a("and.b32 t, m%d, 0x%x;" % (wd, 1 << (ob + kh * KW + kw)))
                a("setp.ne.u32 pt%d, t, 0;" % kw)
        for kw in range(KW):
            c = (o * KH + kh) * KW + kw
            if KW > 1:
                a("@!pt%d bra.uni WH%dE;" % (kw, c))
            a("add.rn.f32 r0, s0, s1;")
            a("add.rn.f32 r1, s2, s3;")
            a("add.rn.f32 r0, r0, r1;")
            a("st.shared.f32 [pp], r0;")
            a("add.u32 pp, pp, 128;")
            n = 0
            for ty in range(TY):
                for tx in range(TX):
                    acc = (o * TY + ty) * TX + tx
                    xi = (ty * S + kh) * XW + tx * S + kw
                    if n < 4:
                        a("mul.rn.f32 s%d, %%%d, x%d;" % (n, I0 + acc, xi))
                    else:
                        a("fma.rn.f32 s%d, %%%d, x%d, s%d;" % (n % 4, I0 + acc, xi, n % 4))
                    n += 1
            a("WH%dE:" % c)
    a("WRR%d:" % NR)
    a("mov.b32 off, noff;")
    for k in range(NW):
        a("mov.b32 m%d, nm%d;" % (k, k))
    a("bra.uni WLOOP;")
    a("WDONE:")
    a("add.rn.f32 r0, s0, s1;")
    a("add.rn.f32 r1, s2, s3;")
    a("add.rn.f32 r0, r0, r1;")
    a("st.shared.f32 [pp], r0;")
    a("sub.u32 pp, pp, %%%d;" % (I0 + NOPS + 3))
    a("shr.u32 %0, pp, 7;")
    a("}")
    body = "\n".join('      "%s\\n\\t"' % s for s in L)
    ops_in = ", ".join('"f"(acc[%d])' % i for i in range(NOPS))
    name = "w%s_o%d_y%d_x%d_k%dx%d_s%d_w%d%s" % (PLAN, OT, TY, TX, KH, KW, S, NCW, "_r%d" % CREGS if CREGS else "")
    src = []
    src.append("// ---- backward-weight variant %s: %d top-diff registers, %d patch registers, %d handlers ----"
               % (name, NACC, NX, NC))
    src.append("template <> struct Interp<%d> {" % vid)
    src.append("  static constexpr int OT = %d, TY = %d, TX = %d, KH = %d, KW = %d, S = %d, PAIR = %d;"
               % (OT, TY, TX, KH, KW, S, PAIR))
    src.append("  static constexpr int NACC = %d, NC = %d, PR = %d, PC = %d, XW = %d, PADL = %d;" % (NOPS, NC, PR, PC, XW, PADL))
    NTW = NCW + (4 if CREGS else NLW)
    src.append("  static constexpr int NCW = %d, NLW = %d, NTW = %d, CREGS = %d, MODE = %d, SHFL = %d;" % (NCW, NLW, NTW, CREGS, 6 if PLAN == "b" else 5, 1 if PLAN == "s" else 0))
    src.append('  static constexpr const char *name() { return "sconv_tile_%s"; }' % name)
    src.append("  // returns the number of taps executed; their partials are in scratch rows 1 .. ntaps (row 0 is a dummy)")
    src.append("  __device__ __forceinline__ static unsigned run_w(const float (&acc)[%d], unsigned prog, unsigned lane_base," % NOPS)
    src.append("                                                   unsigned pitch_bytes, unsigned scratch, unsigned edge) {")
    src.append("    unsigned ntaps;")
    src.append("    asm volatile(")
    src.append(body)
    src.append('      : "=r"(ntaps)')
    src.append("      : %s," % ops_in)
    src.append('        "r"(prog), "r"(lane_base), "r"(pitch_bytes), "r"(scratch), "r"(edge)')
    src.append('      : "memory");')
    src.append("    return ntaps;")
    src.append("  }")
    src.append("  __device__ __forceinline__ static void run(float (&)[%d], unsigned, unsigned, unsigned, unsigned) {}" % NOPS)
    src.append("};")
    return "\n".join(src)


# ---- fourth generation ("rows"): the unit of control flow is a kernel ROW (oc_local, kh).  Empty rows are skipped
# with a warp-uniform direct branch; inside a nonempty row the KW taps are straight-line code, each FMA guarded by
# a predicate (absent taps issue but do not execute).  A taken branch costs ~50 cycles of front-end redirect, a
# predicated-off instruction one issue slot, so this trades redirects for issue slots -- which the packed
# fma.rn.f32x2 (PAIR 2: two images per lane) has to spare: 2 FMAs per lane per slot.
# Stream (16-byte quads): H_0 | H_1 R_0.. | H_2 R_1.. | ... | H_END R_(n-1).. ; H = {plane byte offset | END, mask
# words, pad}, R = the KW weights of one nonempty row (zeros for absent taps), rows in (oc_local, kh) order.
# tuple: (OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS, PLAN, TAP)  TAP: "p" predicated, "j" branch per tap, "d" dense
ROWS = [
    # 3x3, two images per lane (FFMA2), patch-aligned rows
    (4, 4, 4, 3, 3, 1, 2, 8, 4, 232, "b", "j"),
    (2, 4, 4, 5, 5, 1, 2, 8, 4, 232, "b", "j"),
]


def rows_layout(OT, KH, KW):
    OPW, NW = sieve_layout(OT, KH, KW)
    HQ = 1 if NW <= 3 else 2          # header quads
    WQ = 1 if KW <= 4 else 2          # quads per row of weights
    return OPW, NW, HQ, WQ


def gen_rows(vid, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS, PLAN, TAP):
    NACC = OT * TY * TX
    PR = (TY - 1) * S + KH
    PC = (TX - 1) * S + KW
    loads, loads_a, XW, PADL = load_plans(PC, PAIR, KW)
    assert PLAN == "b" or PAIR == 1
    NX = PR * XW
    NC = OT * KH * KW
    OPW, NW, HQ, WQ = rows_layout(OT, KH, KW)
    HB, WB = 16 * HQ, 16 * WQ
    NOPS = NACC * PAIR
    L = []
    a = L.append
    a("{")
    if PAIR == 1:
        a(".reg .f32 x<%d>, ww<%d>;" % (NX, KW))
    else:
        a(".reg .b64 x<%d>, a<%d>, ww<%d>;" % (NX, NACC, KW))
        for i in range(NACC):
            a("mov.b64 a%d, {%%%d, %%%d};" % (i, 2 * i, 2 * i + 1))
    a(".reg .f32 wn<%d>;" % (4 * WQ))
    # m<k>: warp-uniform copies of the mask words (row / channel branches); mv<k>: the same words as plain per-lane
    # registers for the tap predicates -- a predicate ptxas can prove uniform is turned back into a branch
    # (loaded through an address ptxas cannot prove uniform: pc + (lane_base >> 31), the shift is always 0)
    a(".reg .b32 pc, pcl, off, t, ad<%d>, m<%d>, mv<%d>, nv<%d>, nh<%d>, nu<%d>;" % (PR, NW, NW, 4 * HQ, 4 * HQ, NW + 1))
    a("shr.u32 pcl, %%%d, 31;" % (NOPS + 1))
    a(".reg .pred p, q<%d>;" % KW)

    def ld_quads(regs, base, nq, ptr="pc"):
        for k in range(nq):
            a("ld.shared.v4.b32 {%s}, [%s+%d];" % (", ".join("%s%d" % (regs, 4 * k + j) for j in range(4)), ptr, base + 16 * k))

    def ld_quads_f(regs, base, nq):
        for k in range(nq):
            a("ld.shared.v4.f32 {%s}, [pc+%d];" % (", ".join("%s%d" % (regs, 4 * k + j) for j in range(4)), base + 16 * k))
    a("mov.u32 pc, %%%d;" % NOPS)
    a("add.u32 pcl, pcl, pc;")
    ld_quads("nh", 0, HQ)
    if TAP == "p":
        ld_quads("nv", 0, HQ, "pcl")
    a("shfl.sync.idx.b32 off, nh0, 0, 0x1f, 0xffffffff;")
    for k in range(NW):
        a("shfl.sync.idx.b32 m%d, nh%d, 0, 0x1f, 0xffffffff;" % (k, k + 1))
        if TAP == "p":
            a("mov.b32 mv%d, nv%d;" % (k, k + 1))
    a("add.u32 pc, pc, %d;" % HB)
    a("add.u32 pcl, pcl, %d;" % HB)
    a("RLOOP:")
    a("setp.eq.u32 p, off, 0xffffffff;")
    a("@p bra.uni RDONE;")
    ld_quads("nh", 0, HQ)
    if TAP == "p":
        ld_quads("nv", 0, HQ, "pcl")
    ld_quads_f("wn", HB, WQ)
    a("add.u32 pc, pc, %d;" % HB)
    a("add.u32 pcl, pcl, %d;" % HB)
    a("add.u32 ad0, %%%d, off;" % (NOPS + 1))
    for r in range(1, PR):
        a("add.u32 ad%d, ad%d, %%%d;" % (r, r - 1, NOPS + 2))
    plan = loads_a if PLAN == "a" else loads
    for r in range(PR):
        for (po, cnt, byte) in plan:
            b = r * XW + po
            sgn = "+%d" % byte
            if PAIR == 1 and cnt == 4:
                a("ld.shared.v4.f32 {x%d, x%d, x%d, x%d}, [ad%d%s];" % (b, b + 1, b + 2, b + 3, r, sgn))
            elif PAIR == 1 and cnt == 2:
                a("ld.shared.v2.f32 {x%d, x%d}, [ad%d%s];" % (b, b + 1, r, sgn))
            elif PAIR == 1:
                a("ld.shared.f32 x%d, [ad%d%s];" % (b, r, sgn))
            elif cnt == 2:
                a("ld.shared.v2.b64 {x%d, x%d}, [ad%d%s];" % (b, b + 1, r, sgn))
            else:
                a("ld.shared.b64 x%d, [ad%d%s];" % (b, r, sgn))
    # the next header, made warp-uniform early (it is only consumed at the end of the step)
    for k in range(NW + 1):
        a("shfl.sync.idx.b32 nu%d, nh%d, 0, 0x1f, 0xffffffff;" % (k, k))
    for o in range(OT):
        wd, ob = o // OPW, (o % OPW) * KH * KW
        if KH > 1:
            a("and.b32 t, m%d, 0x%x;" % (wd, ((1 << (KH * KW)) - 1) << ob))
            a("setp.eq.u32 p, t, 0;")
            a("@p bra.uni RO%dE;" % o)
        for kh in range(KH):
            a("and.b32 t, m%d, 0x%x;" % (wd, ((1 << KW) - 1) << (ob + kh * KW)))
            a("setp.eq.u32 p, t, 0;")
            a("@p bra.uni RR%d_%dE;" % (o, kh))
            for kw in range(KW):
                if PAIR == 1:
                    a("mov.f32 ww%d, wn%d;" % (kw, kw))
                else:
                    a("mov.b64 ww%d, {wn%d, wn%d};" % (kw, kw, kw))
            ld_quads_f("wn", WB, WQ)
            a("add.u32 pc, pc, %d;" % WB)
            a("add.u32 pcl, pcl, %d;" % WB)
            for kw in range(KW):
                c = (o * KH + kh) * KW + kw
                guard = ""
                if TAP in ("p", "j"):
                    a("and.b32 t, %s%d, 0x%x;" % ("mv" if TAP == "p" else "m", wd, 1 << (ob + kh * KW + kw)))
                    a("setp.ne.u32 q%d, t, 0;" % kw)
                if TAP == "j":
                    a("@!q%d bra.uni RT%dE;" % (kw, c))
                elif TAP == "p":
                    guard = "@q%d " % kw
                for ty in range(TY):
                    for tx in range(TX):
                        acc = (o * TY + ty) * TX + tx
                        xi = (ty * S + kh) * XW + tx * S + kw
                        if PAIR == 1:
                            a("%sfma.rn.f32 %%%d, ww%d, x%d, %%%d;" % (guard, acc, kw, xi, acc))
                        else:
                            a("%sfma.rn.f32x2 a%d, ww%d, x%d, a%d;" % (guard, acc, kw, xi, acc))
                if TAP == "j":
                    a("RT%dE:" % c)
            a("RR%d_%dE:" % (o, kh))
        if KH > 1:
            a("RO%dE:" % o)
    a("mov.b32 off, nu0;")
    for k in range(NW):
        a("mov.b32 m%d, nu%d;" % (k, k + 1))
        if TAP == "p":
            a("mov.b32 mv%d, nv%d;" % (k, k + 1))
    a("bra.uni RLOOP;")
    a("RDONE:")
    if PAIR == 2:
        for i in range(NACC):
            a("mov.b64 {%%%d, %%%d}, a%d;" % (2 * i, 2 * i + 1, i))
    a("}")
    body = "\n".join('      "%s\\n\\t"' % s for s in L)
    ops_out = ", ".join('"+f"(acc[%d])' % i for i in range(NOPS))
    name = "r%s%s_o%d_y%d_x%d_k%dx%d_s%d_p%d_w%d%s" % (PLAN, TAP, OT, TY, TX, KH, KW, S, PAIR, NCW,
                                                      "_r%d" % CREGS if CREGS else "")
    src = []
    src.append("// ---- rows variant %s: %d accumulator registers, %d patch registers%s, %d rows, %d mask words ----"
               % (name, NOPS, NX * PAIR, "" if PAIR == 1 else " (64-bit pairs)", OT * KH, NW))
    src.append("template <> struct Interp<%d> {" % vid)
    src.append("  static constexpr int OT = %d, TY = %d, TX = %d, KH = %d, KW = %d, S = %d, PAIR = %d;"
               % (OT, TY, TX, KH, KW, S, PAIR))
    src.append("  static constexpr int NACC = %d, NC = %d, PR = %d, PC = %d, XW = %d, PADL = %d;" % (NOPS, NC, PR, PC, XW, PADL))
    NTW = NCW + (4 if CREGS else NLW)
    src.append("  static constexpr int NCW = %d, NLW = %d, NTW = %d, CREGS = %d, MODE = %d;"
               % (NCW, NLW, NTW, CREGS, 3 if PLAN == "a" else 4))
    src[-1] = src[-1].replace(";", ", SHFL = 0;", 1) if "SHFL" not in src[-1] else src[-1]
    src.append('  static constexpr const char *name() { return "sconv_tile_%s"; }' % name)
    src.append("  __device__ __forceinline__ static void run(float (&acc)[%d], unsigned prog, unsigned lane_base," % NOPS)
    src.append("                                             unsigned pitch_bytes, unsigned /*plan_a*/) {")
    src.append("    asm volatile(")
    src.append(body)
    src.append("      : %s" % ops_out)
    src.append('      : "r"(prog), "r"(lane_base), "r"(pitch_bytes)')
    src.append('      : "memory");')
    src.append("  }")
    src.append("};")
    return "\n".join(src)


def gen_variant(vid, OT, TY, TX, KH, KW, S, PAIR, NCW, NLW, CREGS=0):
    NACC = OT * TY * TX               # accumulator registers (32-bit for PAIR 1, 64-bit pairs for PAIR 2)
    PR = (TY - 1) * S + KH            # patch rows held in registers
    PC = (TX - 1) * S + KW            # patch cols needed
    per_vec = 4 // PAIR               # positions per 128-bit shared load
    # per-row load plans (register offset in the row, positions, byte offset from the lane base).
    # plan B ("patch aligned", handler NC): the lane base is the patch's first column, 16-byte aligned; 128-bit
    #   loads with a 64-bit tail.  Used when rows are staged element-wise (the halo is pad_w columns wide).
    # plan A ("aligned body", handler NC+1, PAIR 1 only): the lane base is the tile's first OUTPUT column; data
    #   column 0 sits on a 16-byte boundary -- what TMA staging produces -- and the PADL halo columns to its left and
    #   the tail are fetched with naturally aligned narrower loads.
    PADL = (KW - 1) // 2
    loads = []
    pos = 0
    while pos < PC:
        rem = PC - pos
        if PAIR == 1 and rem <= 2:
            loads.append((pos, 2, pos * 4)); pos += 2
        elif PAIR == 2 and rem == 1:
            loads.append((pos, 1, pos * 8)); pos += 1
        else:
            loads.append((pos, per_vec, pos * 4 * PAIR)); pos += per_vec
    loads_a = []
    if PAIR == 1:
        c = -PADL
        while c < PC - PADL:
            rem = PC - PADL - c
            n = 4 if (c % 4 == 0 and rem >= 4) else 2 if (c % 2 == 0 and rem >= 2) else 1
            loads_a.append((c + PADL, n, c * 4))
            c += n
    XW = pos                          # positions per patch row actually loaded
    NX = PR * XW
    NC = OT * KH * KW
    L = []
    a = L.append
    a("{")
    NOPS = NACC * PAIR                # "+f" operands
    if PAIR == 1:
        a(".reg .f32 x<%d>;" % NX)
    else:
        a(".reg .b64 x<%d>, ww, a<%d>;" % (NX, NACC))
        for i in range(NACC):
            a("mov.b64 a%d, {%%%d, %%%d};" % (i, 2 * i, 2 * i + 1))
    a(".reg .b32 wcur, c1, c2, cj, pc, ad<%d>;" % PR)
    # Threaded code.  Record i = {payload_i, handler_{i+2}}: a record carries its own payload and the id of the
    # handler that runs two records later.  Register protocol at a handler's entry: wcur = payload of the record
    # being executed (its shared-memory load may still be in flight: it was issued just before the branch and
    # overlaps the branch's own bubble), c1 = id of the next handler (ready, so the jump-table lookup that brx.idx
    # expands to starts at the top of the handler and hides behind the FMAs), c2 = id of the handler after that
    # (in flight), pc -> the next record.  No payload register rotation; one move per record.
    a("mov.u32 pc, %%%d;" % NOPS)
    a("ld.shared.v2.b32 {cj, c1}, [pc];")         # header: {handler_0, handler_1}
    a("ld.shared.v2.b32 {wcur, c2}, [pc+8];")     # record 0: {payload_0, handler_2}
    a("add.u32 pc, pc, 16;")
    targets = ["HF%d" % c for c in range(NC)] + ["HLOAD", "HLOADA" if PAIR == 1 else "HLOAD", "HEND"]
    a("TL: .branchtargets %s;" % ", ".join(targets))
    a("brx.idx.uni cj, TL;")
    adv = ["mov.b32 c1, c2;", "ld.shared.v2.b32 {wcur, c2}, [pc];", "add.u32 pc, pc, 8;", "brx.idx.uni cj, TL;"]
    for c in range(NC):
        o, kh, kw = c // (KH * KW), (c // KW) % KH, c % KW
        a("HF%d:" % c)
        a("mov.b32 cj, c1;")
        if PAIR == 2:
            a("mov.b64 ww, {wcur, wcur};")
        for ty in range(TY):
            for tx in range(TX):
                acc = (o * TY + ty) * TX + tx
                xi = (ty * S + kh) * XW + tx * S + kw
                if PAIR == 1:
                    a("fma.rn.f32 %%%d, wcur, x%d, %%%d;" % (acc, xi, acc))
                else:
                    a("fma.rn.f32x2 a%d, ww, x%d, a%d;" % (acc, xi, acc))
        L.extend(adv)
    for label, plan in (("HLOAD", loads), ("HLOADA", loads_a)):
        if not plan:
            continue
        a(label + ":")
        a("mov.b32 cj, c1;")
        # wcur = byte offset of the input-channel plane inside the staged chunk; %NOPS+1 = lane base address (shared,
        # bytes); %NOPS+2 = row pitch in bytes
        a("add.u32 ad0, %%%d, wcur;" % (NOPS + 1))
        for r in range(1, PR):
            a("add.u32 ad%d, ad%d, %%%d;" % (r, r - 1, NOPS + 2))
        for r in range(PR):
            for (po, cnt, byte) in plan:
                b = r * XW + po
                sgn = "+%d" % byte
                if PAIR == 1 and cnt == 4:
                    a("ld.shared.v4.f32 {x%d, x%d, x%d, x%d}, [ad%d%s];" % (b, b + 1, b + 2, b + 3, r, sgn))
                elif PAIR == 1 and cnt == 2:
                    a("ld.shared.v2.f32 {x%d, x%d}, [ad%d%s];" % (b, b + 1, r, sgn))
                elif PAIR == 1:
                    a("ld.shared.f32 x%d, [ad%d%s];" % (b, r, sgn))
                elif cnt == 2:
                    a("ld.shared.v2.b64 {x%d, x%d}, [ad%d%s];" % (b, b + 1, r, sgn))
                else:
                    a("ld.shared.b64 x%d, [ad%d%s];" % (b, r, sgn))
        L.extend(adv)
    a("HEND:")
    if PAIR == 2:
        for i in range(NACC):
            a("mov.b64 {%%%d, %%%d}, a%d;" % (2 * i, 2 * i + 1, i))
    a("}")
    body = "\n".join('      "%s\\n\\t"' % s for s in L)
    ops_out = ", ".join('"+f"(acc[%d])' % i for i in range(NOPS))
    name = "o%d_y%d_x%d_k%dx%d_s%d_p%d_w%d%s" % (OT, TY, TX, KH, KW, S, PAIR, NCW, "_r%d" % CREGS if CREGS else "")
    src = []
    src.append("// ---- variant %s: %d accumulator registers%s, %d patch registers, %d handlers ----"
               % (name, NACC, "" if PAIR == 1 else " (64-bit pairs)", NX, NC + 2))
    src.append("template <> struct Interp<%d> {" % vid)
    src.append("  static constexpr int OT = %d, TY = %d, TX = %d, KH = %d, KW = %d, S = %d, PAIR = %d;"
               % (OT, TY, TX, KH, KW, S, PAIR))
    src.append("  static constexpr int NACC = %d, NC = %d, PR = %d, PC = %d, XW = %d, PADL = %d;  // NACC fp32 accumulators"
               % (NOPS, NC, PR, PC, XW, PADL))
    # CREGS > 0: the loader warps form a warpgroup of their own (4 warps, NLW of them active) that gives its
    # registers back with setmaxnreg.dec and the compute warpgroups grow to CREGS with setmaxnreg.inc
    NTW = NCW + (4 if CREGS else NLW)
    src.append("  static constexpr int NCW = %d, NLW = %d, NTW = %d, CREGS = %d, MODE = 0, SHFL = 0;" % (NCW, NLW, NTW, CREGS))
    src.append('  static constexpr const char *name() { return "sconv_tile_%s"; }' % name)
    src.append("  __device__ __forceinline__ static void run(float (&acc)[%d], unsigned prog, unsigned lane_base,"
               % NOPS)
    src.append("                                             unsigned pitch_bytes, unsigned /*plan_a*/) {")
    src.append("    asm volatile(")
    src.append(body)
    src.append("      : %s" % ops_out)
    src.append('      : "r"(prog), "r"(lane_base), "r"(pitch_bytes)')
    src.append('      : "memory");')
    src.append("  }")
    src.append("};")
    return "\n".join(src)


def main():
    os.makedirs(OUTDIR, exist_ok=True)
    head = ["// GENERATED by tools/gen_interp.py -- do not edit.  Inline-PTX threaded-code interpreter for one tile shape.",
            "// Record stream (shared memory): a header {handler_0, handler_1} then 8-byte records {u32 payload_i, u32 handler_(i+2)};",
            "// handler < NC: FMA (payload = fp32 weight), handler == NC / NC+1: LOAD patch, patch-aligned / aligned-body plan (payload = byte",
            "// offset of the channel plane), NC+2: end of segment.",
            "#pragma once",
            "template <int VID> struct Interp;"]
    ALL = ([(v, 0) for v in VARIANTS] + [(v, 1 if v[10] == "a" else 2) for v in SIEVE] +
           [(v, 3 if v[10] == "a" else 4) for v in ROWS] + [(v, 2 if v[10] == "b" else 1) for v in SIEVE_R] +
           [(v, 6 if v[10] == "b" else 5) for v in BWDW])
    for i, (v, mode) in enumerate(ALL):
        path = os.path.join(OUTDIR, "interp_v%d.inc" % i)
        txt = "\n".join(head + [(gen_variant, gen_sieve, gen_sieve, gen_rows, gen_rows, gen_bwdw, gen_bwdw)[mode](i, *v)]) + "\n"
        if not os.path.exists(path) or open(path).read() != txt:
            open(path, "w").write(txt)
    lst = os.path.join(OUTDIR, "variant_list.inc")
    txt = "// GENERATED by tools/gen_interp.py\n#define ESCORT_NUM_VARIANTS %d\n#define ESCORT_VARIANT_LIST(X) \\\n" % len(ALL)
    def row(i, v, mode):
        cregs = v[9] if len(v) > 9 else 0
        ntw = v[7] + (4 if cregs else v[8])
        shfl = 1 if (len(v) > 10 and v[10] == "s") else 0
        return "  X(%d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d)" % ((i,) + tuple(v[:9]) + (ntw, mode, shfl))
    txt += " \\\n".join(row(i, v, mode) for i, (v, mode) in enumerate(ALL)) + "\n"
    if not os.path.exists(lst) or open(lst).read() != txt:
        open(lst, "w").write(txt)
    print("generated %d variants in %s" % (len(ALL), OUTDIR))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--count":
        print(len(VARIANTS) + len(SIEVE) + len(ROWS) + len(SIEVE_R) + len(BWDW))
    else:
        main()
