"""Named inputs: `<circuit>.dat` + input JSON -> the positional input rows the C ABI takes.

Mirrors what the reference's generated calculator does before `run(ctx)`:
  * loadCircuit reads the input hash map from `<circuit>.dat`            (common/main.cpp:22-124, App. A.3)
  * loadJson qualifies nested names (`a.b[i].c`), parses numbers          (common/main.cpp:144-284)
  * setInputSignal looks the FNV-1a hash up with linear probing, checks sizes and double assignment
                                                                           (common/calcwit.cpp:17-24, 51-97)
Errors the reference reports with assert / runtime_error are raised as InputError with the same message text.
This is host-side glue above the C ABI (no arithmetic happens here; values >= q are reduced on the device like
Fr_str2element does).
"""
from __future__ import annotations

import json
import re
import struct

from .formats import fnv1a


class InputError(ValueError):
    pass


class InputMap:
    """The input hash map of a `.dat` file.  The file does not carry its own section sizes (the generated C++ does,
    circuit.rs:463-497); the map size is the power of two >= 256 for which the witness-to-signal list that follows
    matches the program's `%%witness` list."""

    def __init__(self, dat_bytes: bytes, witness: list[int], input_start: int, n_inputs: int):
        self.input_start, self.n_inputs = input_start, n_inputs
        want = b"".join(struct.pack("<Q", s) for s in witness)
        size = 256
        table = None
        while 24 * size + len(want) <= len(dat_bytes):
            if dat_bytes[24 * size:24 * size + len(want)] == want:
                table = [struct.unpack_from("<QQQ", dat_bytes, 24 * k) for k in range(size)]
                break
            size *= 2
        if table is None:
            raise InputError("the .dat file does not belong to this program (witness list not found)")
        self.table = table

    @classmethod
    def from_files(cls, dat_path, wc):
        """wc: engine.WitnessCalculator (for the witness list and the main-input range)"""
        with open(dat_path, "rb") as f:
            data = f.read()
        return cls(data, wc.witness_signals(), 1 + wc.n_outputs, wc.n_inputs)

    def position(self, h):          # calcwit.cpp:51-69
        n = len(self.table)
        pos = h % n
        for _ in range(n):
            eh, sid, _sz = self.table[pos]
            if eh == h:
                return pos
            if sid == 0:
                raise InputError("Signal not found")
            pos = (pos + 1) % n
        raise InputError("Signals not found")


class SymInputMap(InputMap):
    """The same lookup built from a `.sym` file (`circom --sym`: one `#s,#w,#c,name` line per signal,
    constraint_writers/src/sym_writer.rs:4-14) instead of the `.dat` hash map: the main component's input signals are
    grouped by name without their trailing indices (`main.in[3]` -> key `in`, as loadJson addresses them,
    main.cpp:241-284), first signal and element count per key."""

    def __init__(self, sym_entries, input_start: int, n_inputs: int):
        self.input_start, self.n_inputs = input_start, n_inputs
        groups = {}
        for label, _w, _c, name in sym_entries:
            if not (input_start <= label < input_start + n_inputs) or not name.startswith("main."):
                continue
            key = re.sub(r"(\[\d+\])+$", "", name[5:])
            first, count = groups.get(key, (label, 0))
            groups[key] = (min(first, label), count + 1)
        if sum(c for _f, c in groups.values()) != n_inputs:
            raise InputError("the .sym file does not belong to this program (main inputs not found)")
        self.table = [(fnv1a(k), first, count) for k, (first, count) in groups.items()]
        self._pos = {h: i for i, (h, _f, _c) in enumerate(self.table)}

    @classmethod
    def from_files(cls, sym_path, wc):
        from . import formats
        return cls(formats.read_sym(sym_path), 1 + wc.n_outputs, wc.n_inputs)

    def position(self, h):
        if h not in self._pos:
            raise InputError("Signal not found")
        return self._pos[h]


class NamedInputMap(SymInputMap):
    """The same lookup from the program's own `;;%%main_input <name> <first signal> <size>` lines (a compiler that carries
    patches/main_input_directive.rs.diff): neither the .dat nor the .sym is needed."""

    def __init__(self, main_inputs, input_start: int, n_inputs: int):
        self.input_start, self.n_inputs = input_start, n_inputs
        if not main_inputs or sum(sz for _n, _s, sz in main_inputs) != n_inputs:
            raise InputError("the program text does not name its main inputs")
        self.table = [(fnv1a(name), start, size) for name, start, size in main_inputs]
        self._pos = {h: i for i, (h, _f, _c) in enumerate(self.table)}

    @classmethod
    def from_program(cls, wc):
        return cls(wc.main_inputs(), 1 + wc.n_outputs, wc.n_inputs)


Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _parse_number(val):             # json2FrElements, main.cpp:144-190; -> the value mod q, as Fr_str2element (mpz_fdiv_r) leaves it
    if isinstance(val, bool):
        raise InputError("Invalid JSON type")
    if isinstance(val, str):
        s, base = val, 10
        p = val[:2]
        if p in ("0b", "0B"):
            s, base = val[2:], 2
        elif p in ("0o", "0O"):
            s, base = val[2:], 8
        elif p in ("0x", "0X"):
            s, base = val[2:], 16
        digits = "0123456789abcdef"[:base] if base == 16 else "0123456789"[:base]
        # check_valid_number (main.cpp:126-142): digits of the base only -- a sign is not a digit, so "-5" is rejected
        if any(ch not in digits + (digits.upper() if base == 16 else "") for ch in s):
            raise InputError("Invalid number in JSON input: %s" % val)
        return int(s, base) % Q if s else 0
    # JSON numbers go through a double printed with no decimals (main.cpp:167-172); the sign survives and Fr_str2element
    # reduces with a floor modulus: -5 -> q - 5
    if isinstance(val, int):
        return (int(format(float(val), ".0f")) if abs(val) >= 1 << 53 else val) % Q
    if isinstance(val, float):
        if val != val or val in (float("inf"), float("-inf")):
            raise InputError("Invalid number in JSON input: %s" % val)
        return int(format(val, ".0f")) % Q
    raise InputError("Invalid JSON type")


def _flatten(val, out):
    if isinstance(val, list):
        for v in val:
            _flatten(v, out)
    else:
        out.append(_parse_number(val))


def _elem_type(prefix, v):          # check_type, main.cpp:192-207
    if not isinstance(v, list):     # nlohmann value_t: unsigned / negative integers and floats are distinct types
        if isinstance(v, dict):
            return "object"
        if isinstance(v, bool):
            return "boolean"
        if isinstance(v, int):
            return "number_unsigned" if v >= 0 else "number_integer"
        return {float: "number_float", str: "string", type(None): "null"}.get(type(v), type(v).__name__)
    if not v:
        return "null"
    t = _elem_type(prefix, v[0])
    for x in v[1:]:
        if _elem_type(prefix, x) != t:
            raise InputError("Types are not the same in the the key %s" % prefix)
    return t


def qualify(prefix, v, out):        # qualify_input / qualify_input_list, main.cpp:209-239
    if isinstance(v, list):
        if v and _elem_type(prefix, v) == "object":
            def walk(pfx, x):
                if isinstance(x, list):
                    for i, y in enumerate(x):
                        walk("%s[%d]" % (pfx, i), y)
                else:
                    qualify(pfx, x, out)
            walk(prefix, v)
        else:
            out[prefix] = v
    elif isinstance(v, dict):
        for k, x in v.items():
            qualify(k if not prefix else prefix + "." + k, x, out)
    else:
        out[prefix] = v


def row_from_json(imap: InputMap, doc) -> list[int]:
    """One input object -> the main inputs in signal order (loadJson + setInputSignal)."""
    flat = {}
    qualify("", doc, flat)
    row = [None] * imap.n_inputs
    for name, val in flat.items():
        h = fnv1a(name)
        try:
            pos = imap.position(h)
        except InputError as e:
            raise InputError("Error loading signal %s: %s" % (name, e)) from None
        _eh, sid, size = imap.table[pos]
        vals = []
        _flatten(val, vals)
        if len(vals) < size:
            raise InputError("Error loading signal %s: Not enough values" % name)
        if len(vals) > size:
            raise InputError("Error loading signal %s: Too many values" % name)
        for i, x in enumerate(vals):
            k = sid + i - imap.input_start
            if not 0 <= k < imap.n_inputs:
                raise InputError("Error setting signal: %s" % name)
            if row[k] is not None:
                raise InputError("Error setting signal: %s\nSignal assigned twice: %d" % (name, sid + i))
            row[k] = x
    if any(x is None for x in row):
        raise InputError("Not all inputs have been set. Only %d out of %d" % (sum(x is not None for x in row), imap.n_inputs))
    return row


def rows_from_json_text(imap: InputMap, text: str) -> list[list[int]]:
    """A JSON object (one witness, the reference's input.json) or an array of objects (a batch)."""
    doc = json.loads(text)
    docs = doc if isinstance(doc, list) else [doc]
    return [row_from_json(imap, d) for d in docs]
