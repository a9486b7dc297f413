"""Python readers/writers for the reference's binary file formats (little-endian).

Host-side helpers for tests, fixtures and the bench harness; the product's own loaders are
in csrc/ (C++) behind the C ABI.  Layouts follow the reference writers:

  .r1cs  constraint_writers/src/r1cs_writer.rs:6-14,49-91,245-341; section order on disk is
         constraints(2) -> header(1) -> wire2label(3) (constraint_list/src/r1cs_porting.rs:19-53)
  .wtns  code_producers/src/c_elements/common/main.cpp:286-332
  .dat   code_producers/src/c_elements/c_code_generator.rs:511-615,754-804
  .sym   constraint_writers/src/sym_writer.rs:4-14
"""
from __future__ import annotations

import struct

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


# ------------------------------------------------------------------ .r1cs
def _lc_bytes(lc, fs):
    # term order = lexicographic order of the wire id's minimal LE byte string (r1cs_writer.rs:59-60)
    def key(w):
        b = w.to_bytes(8, "little").rstrip(b"\0")
        return b if b else b"\0"
    out = [struct.pack("<I", len(lc))]
    for w in sorted(lc, key=key):
        out.append(struct.pack("<I", w))
        out.append(int(lc[w]).to_bytes(fs, "little"))
    return b"".join(out)


def write_r1cs(path, constraints, n_wires, n_pub_out, n_pub_in, n_prv_in, wire2label, prime=BN254_R, n_labels=None):
    """constraints: list of (A, B, C) dicts {wire id -> canonical coefficient}."""
    fs = ((prime.bit_length() + 63) // 64) * 8
    body = b"".join(_lc_bytes(a, fs) + _lc_bytes(b, fs) + _lc_bytes(c, fs) for (a, b, c) in constraints)
    header = (struct.pack("<I", fs) + prime.to_bytes(fs, "little") +
              struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in,
                          n_labels if n_labels is not None else len(wire2label), len(constraints)))
    w2l = b"".join(struct.pack("<Q", x) for x in wire2label)
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        for typ, payload in ((2, body), (1, header), (3, w2l)):
            f.write(struct.pack("<IQ", typ, len(payload)))
            f.write(payload)


def read_r1cs(path):
    """-> dict(prime, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints,
    constraints [(A,B,C) dicts], wire2label).  Scans the section table like r1cs_reader.rs:459-476."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"r1cs":
        raise ValueError("not an r1cs file")
    version, nsec = struct.unpack_from("<II", data, 4)
    if version != 1:
        raise ValueError("unsupported r1cs version %d" % version)
    pos = 12
    sections = {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections[typ] = (pos, size)
        pos += size
    hp, _hs = sections[1]
    fs, = struct.unpack_from("<I", data, hp)
    prime = int.from_bytes(data[hp + 4:hp + 4 + fs], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_cons = struct.unpack_from("<IIIIQI", data, hp + 4 + fs)
    cp, _cs = sections[2]
    cons = []
    p = cp
    for _ in range(n_cons):
        triple = []
        for _k in range(3):
            n, = struct.unpack_from("<I", data, p)
            p += 4
            lc = {}
            for _t in range(n):
                w, = struct.unpack_from("<I", data, p)
                lc[w] = int.from_bytes(data[p + 4:p + 4 + fs], "little")
                p += 4 + fs
            triple.append(lc)
        cons.append(tuple(triple))
    w2l = []
    if 3 in sections:
        wp, ws = sections[3]
        w2l = list(struct.unpack_from("<%dQ" % (ws // 8), data, wp))
    return dict(prime=prime, field_size=fs, n_wires=n_wires, n_pub_out=n_pub_out, n_pub_in=n_pub_in,
                n_prv_in=n_prv_in, n_labels=n_labels, n_constraints=n_cons, constraints=cons, wire2label=w2l)


# ------------------------------------------------------------------ .wtns
def wtns_bytes(values, prime=BN254_R):
    n8 = ((prime.bit_length() + 63) // 64) * 8
    out = [b"wtns", struct.pack("<II", 2, 2),
           struct.pack("<IQ", 1, 8 + n8), struct.pack("<I", n8), prime.to_bytes(n8, "little"),
           struct.pack("<I", len(values)),
           struct.pack("<IQ", 2, n8 * len(values))]
    out += [int(v).to_bytes(n8, "little") for v in values]
    return b"".join(out)


def write_wtns(path, values, prime=BN254_R):
    with open(path, "wb") as f:
        f.write(wtns_bytes(values, prime))


def read_wtns(path_or_bytes):
    data = path_or_bytes
    if not isinstance(data, (bytes, bytearray)):
        with open(path_or_bytes, "rb") as f:
            data = f.read()
    if data[:4] != b"wtns":
        raise ValueError("not a wtns file")
    version, nsec = struct.unpack_from("<II", data, 4)
    pos = 12
    prime, n8, vals = None, None, None
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", data, pos)
        pos += 12
        if typ == 1:
            n8, = struct.unpack_from("<I", data, pos)
            prime = int.from_bytes(data[pos + 4:pos + 4 + n8], "little")
            nvars, = struct.unpack_from("<I", data, pos + 4 + n8)
        elif typ == 2:
            vals = [int.from_bytes(data[pos + i * n8:pos + (i + 1) * n8], "little") for i in range(size // n8)]
        pos += size
    assert len(vals) == nvars
    return dict(version=version, prime=prime, n8=n8, values=vals)


# ------------------------------------------------------------------ .dat
def fnv1a(s):
    """64-bit FNV-1a (calcwit.cpp:17-24; code_producers/src/components/mod.rs:46-51)."""
    h = 0xCBF29CE484222325
    for ch in s.encode():
        h ^= ch
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def dat_bytes(main_inputs, witness2signal, constants, prime=BN254_R, io_map=None):
    """main_inputs: list of (qualified name, first signal id, size).  constants: canonical ints.
    io_map: {template instance id: [(offset, dims, element size)] by signal code} (only circuits with mixed component
    arrays have one, build.rs:520-552); no bus section."""
    n = len(main_inputs)
    size = 256
    while size < n:
        size *= 2
    table = [(0, 0, 0)] * size
    for name, start, sz in main_inputs:       # c_code_generator.rs:511-539 (open addressing, linear probe)
        h = fnv1a(name)
        pos = h % size
        while table[pos][1] != 0:
            pos = (pos + 1) % size
        table[pos] = (h, start, sz)
    out = [struct.pack("<QQQ", *e) for e in table]
    out += [struct.pack("<Q", s) for s in witness2signal]
    nbits = ((prime.bit_length() + 63) // 64) * 64
    R = 1 << nbits
    for c in constants:                        # c_code_generator.rs:552-615: {i32 short, u32 type, n*R mod q}
        sv = c - prime if c > prime // 2 else c
        if -(1 << 31) <= sv <= (1 << 31) - 1:
            short, typ = sv, 0x40000000
        else:
            short, typ = 0, 0xC0000000
        out.append(struct.pack("<iI", short, typ) + ((c * R) % prime).to_bytes(nbits // 8, "little"))
    if io_map:                                 # c_code_generator.rs:617-674 (BTreeMap: ascending ids)
        ids = sorted(io_map)
        out += [struct.pack("<I", i) for i in ids]
        for i in ids:
            out.append(struct.pack("<I", len(io_map[i])))
            for offset, dims, size in io_map[i]:
                out.append(struct.pack("<II", offset, max(len(dims) - 1, 0)))
                out += [struct.pack("<I", d) for d in dims[1:]]
                out.append(struct.pack("<II", size, 0))
    return b"".join(out)


# ------------------------------------------------------------------ batch witness container helpers
def pack_inputs(rows, n_inputs):
    """rows: iterable of per-witness lists of canonical ints -> bytes B x n_inputs x 32 (LE)."""
    out = bytearray()
    for r in rows:
        assert len(r) == n_inputs
        for v in r:
            out += int(v).to_bytes(32, "little")
    return bytes(out)


# ---- .sym (constraint_writers/src/sym_writer.rs:4-14; mkdocs/docs/circom-language/formats/sym.md) -------------------
def write_sym(path, entries):
    """entries: iterable of (signal label, witness position or -1, component number, qualified name)"""
    with open(path, "w") as f:
        for s, w, c, name in entries:
            f.write("%d,%d,%d,%s\n" % (s, w, c, name))


def read_sym(path):
    out = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            s, w, c, name = line.split(",", 3)
            out.append((int(s), int(w), int(c), name))
    return out
