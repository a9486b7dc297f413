"""Test helper: execute a compiled tape (as the CUDA tape kernel would) with Python big ints.

Used ONLY by the CPU test-suite to check the host-side trace compiler and slot allocator
(csrc/tracer.hpp, csrc/tape.hpp) without a GPU.  It is not importable from the product package
and is not a fallback: the product path has no CPU execution.
"""
from oracle import fr_model as M

(T_NOP, T_INPUT, T_ADD, T_SUB, T_MUL, T_DIV, T_IDIV, T_MOD, T_POW, T_SHL, T_SHR, T_BAND, T_BOR, T_BXOR, T_BNOT,
 T_LT, T_LE, T_GT, T_GE, T_EQ, T_NEQ, T_LAND, T_LOR, T_EQZ, T_SEL, T_FAIL_IF, T_FAIL_NE, T_BITC, T_LUT, T_INV,
 T_CADD, T_DOT, T_LD, T_ST, T_STC, T_ICADD, T_IADD, T_ISEL, T_IBIT, T_IFAIL_NE, T_ISUM, T_LUTG, T_IBITG, T_FILL, T_RNE, T_INPUT_BIT, T_ISUMT, T_INBITG) = range(48)
F_ADDEND = 32        # T_DOT: field b is an addend
F_ADDEND2 = 16       # T_ISUM / T_ISUMT: a second addend in field c
ST_SPECULATION = 6
F_CHECK = 128        # T_ADD / T_SUB / T_MUL / T_DOT of a fused R1CS check: compare the result with slot dst, c = constraint
F_RING = 64          # T_LD: value comes from ring entry b (requested LD_RING reloads earlier)
NO_ROW = 0xFFFFFFFF
F_CZERO = 16         # T_SEL: third operand is the constant 0
F_STORE = 8          # flag bit 3: the result is also stored to value-store row c
R = 1 << 256

BIN = {T_ADD: M.add, T_SUB: M.sub, T_MUL: M.mul, T_POW: M.pow_, T_SHL: M.shl, T_SHR: M.shr, T_BAND: M.band,
       T_BOR: M.bor, T_BXOR: M.bxor, T_LT: M.lt, T_LE: M.leq, T_GT: M.gt, T_GE: M.geq, T_EQ: M.eq, T_NEQ: M.neq,
       T_LAND: M.land, T_LOR: M.lor}


BSLOT = 0x40000000      # operand field: a slot of the bit file (value typed 0/1)
BSLOT_DST = 0x8000      # dst field: idem
ROW_BIT = 0x80000000    # row field: a bit row


def run_tape(tape, consts_mont, layout, inputs, want_first_bad=False):
    """layout: WitnessCalculator.layout() (slot files, typed rows, wire -> row map)
    -> (witness: canonical value of every wire (None = its row was never written), status)
    want_first_bad: also the first violated constraint of a tape with a fused R1CS check (T_RNE), 0xffffffff if none"""
    first_bad = 0xFFFFFFFF
    consts = [M.from_mont(c) for c in consts_mont]
    slots = [None] * layout["n_slots"]
    bslots = [None] * layout["n_bslots"]
    frows = [None] * layout["n_frows"]
    brows = [None] * layout["n_brows"]
    ring = [None] * 8     # reload stream (entries b < LD_RING): a snapshot of the row at request time, so that a stale request is caught
    status = 0
    import numpy as np
    words = np.ascontiguousarray(tape).view(np.uint32).reshape(-1, 4)

    def get_slot(code):
        if code & BSLOT:
            v = bslots[code & 0xFFFF]
            assert v in (0, 1), "bit slot holds %r" % (v,)
        else:
            v = slots[code]
        assert v is not None, "read of an empty slot"
        return v

    def set_slot(dst, v):
        if dst & BSLOT_DST:
            assert v in (0, 1), "a value typed 0/1 is %r" % (v,)
            bslots[dst & 0x7FFF] = v
        else:
            slots[dst] = v

    def set_row(row, v):
        if row & ROW_BIT:
            assert v in (0, 1), "bit row written with %r" % (v,)
            brows[row & ~ROW_BIT] = v
        else:
            frows[row] = v

    def get_row(row):
        v = brows[row & ~ROW_BIT] if row & ROW_BIT else frows[row]
        assert v is not None, "load of an unwritten row"
        return v

    iconsts = layout.get("iconsts", [])

    class Int(int):
        """a raw 64-bit integer in a field slot: must never be read as a field element"""

    def get_int(idx, is_const):
        if is_const:
            return iconsts[idx]
        v = get_slot(idx)
        assert idx & BSLOT or isinstance(v, Int), "integer operation reads a field value"
        return int(v)

    pc = 0
    while pc < len(tape):
        ins = tape[pc]
        pc += 1
        op, flags, dst, a, b, c = (int(ins["op"]), int(ins["flags"]), int(ins["dst"]), int(ins["a"]), int(ins["b"]),
                                   int(ins["c"]))

        def operand(idx, bit):
            if flags & bit:
                return consts[idx]
            v = get_slot(idx)
            assert not isinstance(v, Int), "field operation reads a raw integer"
            return v

        res = None
        if op == T_DOT:
            # a terms follow as (constant index, slot) pairs, two per 16-byte record
            acc = 0
            for j in range(a):
                rec = words[pc + j // 2]
                cidx, slot = (int(rec[2]), int(rec[3])) if j & 1 else (int(rec[0]), int(rec[1]))
                acc += consts[cidx] * get_slot(slot)
            if flags & F_ADDEND:
                acc += operand(b, 2)
            pc += (a + 1) // 2
            res = acc % M.Q
        elif op == T_ISUM:
            # addend + sum_j (bit_j << shift_j): a terms follow, four (bit slot | shift << 16) per 16-byte record
            acc = get_int(b, flags & 2) if flags & F_ADDEND else 0
            if flags & F_ADDEND2:
                acc += get_int(c, flags & 4)
            for j in range(a):
                t = int(words[pc + j // 4][j % 4])
                v = bslots[t & 0xFFFF]
                assert v in (0, 1), "T_ISUM term is not a 0/1 value"
                acc += v << (t >> 16)
            pc += (a + 3) // 4
            assert acc < 1 << 62 and not dst & BSLOT_DST
            res = Int(acc)
        elif op == T_ISUMT:
            # T_ISUM with its terms dealt into layers of distinct shifts: a layers of 32 words, word l = bit slot | base << 16 of
            # the term whose shift is base + l (slot 0xffff: none)
            acc = get_int(b, flags & 2) if flags & F_ADDEND else 0
            if flags & F_ADDEND2:
                acc += get_int(c, flags & 4)
            for layer in range(a):
                for l in range(32):
                    t = int(words[pc + layer * 8 + l // 4][l % 4])
                    if t & 0xFFFF == 0xFFFF:
                        continue
                    v = bslots[t & 0xFFFF]
                    assert v in (0, 1), "T_ISUMT term is not a 0/1 value"
                    acc += v << ((t >> 16) + l)
            pc += a * 8
            assert acc < 1 << 62 and not dst & BSLOT_DST
            res = Int(acc)
        elif op in (T_LUTG, T_IBITG, T_INBITG):
            # a members, one record each; every member reads its operands before any result is written
            outs = []
            src = get_int(b, 0) if op == T_IBITG else None
            for m in range(a):
                rec = [int(x) for x in words[pc + m]]
                if op == T_LUTG:
                    sl = [rec[0] & 0xFFFF, rec[0] >> 16, rec[1] & 0xFFFF]
                    idx = 0
                    for i in range((rec[2] >> 8) & 0xFF):
                        v = bslots[sl[i]]
                        assert v in (0, 1), "T_LUTG input is not a 0/1 value"
                        idx |= v << i
                    outs.append(((rec[2] >> idx) & 1, rec[1] >> 16, rec[3]))
                elif op == T_INBITG:
                    v = int(inputs[c + m])          # main input c + m taken as a bit (speculative typing)
                    if v > 1:
                        status = ST_SPECULATION
                    outs.append((v & 1, rec[0] & 0xFFFF, rec[3]))
                else:
                    outs.append(((src >> (c + m)) & 1, rec[0] & 0xFFFF, rec[3]))
            for v, d, row in outs:
                bslots[d] = v
                if row != NO_ROW:
                    assert row & ROW_BIT
                    set_row(row, v)
            pc += a
            continue
        elif op == T_FILL:
            assert a in (0, 0xFFFFFFFF) and c & ROW_BIT
            for k in range(b):
                set_row(c + k, 1 if a else 0)
            continue
        elif op == T_INPUT:
            res = inputs[a] % M.Q
        elif op == T_CADD:
            # a + (b != 0 ? constant c : 0)
            x = operand(a, 1)
            res = (x + consts[c]) % M.Q if get_slot(b) != 0 else x
        elif op == T_LUT:
            # boolean function of up to three bit slots: a = slots 0 | 1 << 16; b = table | number of inputs << 8 | slot 2 << 16
            sl = [a & 0xFFFF, a >> 16, b >> 16]
            idx = 0
            for i in range((b >> 8) & 0xFF):
                v = bslots[sl[i]]
                assert v in (0, 1), "T_LUT input is not a 0/1 value"
                idx |= v << i
            res = (b >> idx) & 1
            assert dst & BSLOT_DST
        elif op == T_BITC:
            # bit b of the RAW (Montgomery) limbs of field slot a
            assert not (a & BSLOT) and slots[a] is not None and not isinstance(slots[a], Int)
            res = ((slots[a] * R % M.Q) >> b) & 1
        elif op == T_FAIL_NE:
            if status == 0 and operand(a, 1) != operand(b, 2):
                status = c
            continue
        elif op == T_INPUT_BIT:
            # speculative typing: the input must be literally 0 or 1, else the witness is marked for the generic program
            v = int(inputs[a])
            if v > 1:
                status = ST_SPECULATION
            res = v & 1
        elif op == T_RNE:
            if operand(a, 1) != operand(b, 2):
                first_bad = min(first_bad, c)
            continue
        elif op == T_IFAIL_NE:
            if status == 0 and get_int(a, flags & 1) != get_int(b, flags & 2):
                status = c
            continue
        elif op == T_ICADD:
            res = get_int(a, flags & 1)
            if int(get_slot(b)) != 0:
                res += iconsts[c]
            assert res < 1 << 62
            res = Int(res)
        elif op == T_IADD:
            res = Int(get_int(a, flags & 1) + get_int(b, flags & 2))
            assert res < 1 << 62
        elif op == T_ISEL:
            cond = consts[a] if flags & 1 else int(get_slot(a))
            res = Int(get_int(b, flags & 2) if cond != 0 else get_int(c, flags & 4))
        elif op == T_IBIT:
            res = (get_int(a, 0) >> b) & 1
        elif op == T_LD:
            v = get_row(c)
            if dst & BSLOT_DST:
                assert c & ROW_BIT and not flags & F_RING
                bslots[dst & 0x7FFF] = v
                continue
            assert not c & ROW_BIT
            if flags & F_RING:
                assert ring[b] is not None and ring[b][0] == c, "ring entry does not hold the expected row"
                slots[dst] = ring[b][1]
            else:
                slots[dst] = v
            assert slots[dst] == v, "streamed reload is stale"
            ring[b] = None
            if a != NO_ROW:
                assert not a & ROW_BIT and frows[a] is not None, "request of an unwritten row"
                ring[b] = (a, frows[a])
            continue
        elif op == T_ST:
            assert bool(a & BSLOT) == bool(c & ROW_BIT), "store between a slot and a row of different types"
            set_row(c, get_slot(a))
            continue
        elif op == T_STC:
            set_row(c, consts[a])
            continue
        elif op == T_FAIL_IF:
            if status == 0 and operand(a, 1) != 0:
                status = c
            continue
        elif op == T_SEL:
            x, y = operand(a, 1), operand(b, 2)
            z = 0 if flags & F_CZERO else operand(c, 4)
            res = y if x != 0 else z
        elif op in (T_BNOT, T_EQZ, T_INV):
            x = operand(a, 1)
            res = M.bnot(x) if op == T_BNOT else int(x == 0) if op == T_EQZ else pow(x, M.Q - 2, M.Q)
        elif op == T_DIV:
            x, y = operand(a, 1), operand(b, 2)
            res = 0 if y == 0 else M.div(x, y)
        elif op in (T_IDIV, T_MOD):
            x, y = operand(a, 1), operand(b, 2)
            if y == 0:          # the failure is raised by the FAIL_IF the tracer puts in front of the operation
                res = 0
            else:
                res = x // y if op == T_IDIV else x % y
        elif op in BIN:
            res = BIN[op](operand(a, 1), operand(b, 2))
        else:
            raise ValueError("bad tape op %d" % op)
        if flags & F_CHECK:
            # last instruction of a constraint of the fused R1CS check: compare with slot dst, constraint c; nothing is written
            assert op in (T_ADD, T_SUB, T_MUL, T_DOT) and not flags & F_STORE
            if res != get_slot(dst):
                first_bad = min(first_bad, c)
            continue
        set_slot(dst, res)
        if flags & F_STORE:
            assert op not in (T_CADD,) and (op != T_SEL or flags & F_CZERO)
            assert bool(dst & BSLOT_DST) == bool(c & ROW_BIT), "fused store between a slot and a row of different types"
            set_row(c, res)
    witness = [(brows[loc & ~ROW_BIT] if loc & ROW_BIT else frows[loc]) for loc in layout["wire_loc"]]
    return (witness, status, first_bad) if want_first_bad else (witness, status)
