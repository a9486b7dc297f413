"""Drive the stand-in compiler: description -> (.cvm, .r1cs[, .cpp + .dat]) in the reference's formats."""
from __future__ import annotations

import os

from circom_cvm_b200 import formats

from .emit_cvm import emit_cvm
from .execute import build_program, flatten_constraints, simplify_o1
from .translate import compile_program


class Artifact:
    pass


def compile_circuit(main_fn, args=(), public=(), functions=(), name=None, o1=True, constraint_assert_disabled=False):
    prog = build_program(main_fn, args, public, functions)
    main = prog.main
    n_signals = main.n_signals + 1                       # + constant-one signal 0 (build.rs:339)
    cons = flatten_constraints(prog)
    n_main_io = 1 + main.n_out + main.n_in
    if o1:
        cons, witness = simplify_o1(cons, n_signals, protect=n_main_io)
    else:
        witness = list(range(n_signals))
    sig2wit = {s: w for w, s in enumerate(witness)}
    r1cs_cons = [tuple({sig2wit[s]: v for s, v in lc.items()} for lc in abc) for abc in cons]
    compiled = compile_program(prog, constraint_assert_disabled)
    compiled.witness = witness
    art = Artifact()
    art.name = name or main_fn.__name__
    art.prog, art.compiled = prog, compiled
    art.witness = witness
    art.n_signals = n_signals
    art.constraints = r1cs_cons
    art.n_wires = len(witness)
    art.n_pub_out = main.n_out
    art.n_pub_in = getattr(prog, "n_pub_in", 0)
    art.n_prv_in = main.n_in - art.n_pub_in
    art.n_inputs = main.n_in
    art.n_outputs = main.n_out
    art.input_start = 1 + main.n_out                     # c_elements/mod.rs:154-156
    art.main_inputs = [(s.name, 1 + s.offset, s.size) for s in main.tmpl.signals if s.xtype == "in"]
    art.main_outputs = [(s.name, 1 + s.offset, s.size) for s in main.tmpl.signals if s.xtype == "out"]
    art.cvm = emit_cvm(compiled)
    return art


def faithful_cvm(art):
    """The same program printed exactly as the fork's --cvm emitters print it, defects included (emit_cvm.py): no component
    creation, literal addresses assigned to, the defective array-equality and multi-element return shapes.  Loadable
    together with the generated C++ (emit_cpp), which carries the component creation."""
    return emit_cvm(art.compiled, faithful=True)


def _indices(dims, flat):
    out = []
    for d in reversed(dims):
        out.append(flat % d)
        flat //= d
    return "".join("[%d]" % i for i in reversed(out))


def sym_entries(art):
    """(#s, #w, #c, name) for every signal, as `circom --sym` lists them (constraint_writers/src/sym_writer.rs:4-14,
    mkdocs/docs/circom-language/formats/sym.md): signal label, witness position or -1, component number, qualified name."""
    sig2wit = {s: w for w, s in enumerate(art.witness)}
    out = []

    def walk(inst, base_sig, cmp_id, prefix):
        for s in inst.tmpl.signals:
            for flat in range(s.size):
                label = base_sig + s.offset + flat
                out.append((label, sig2wit.get(label, -1), cmp_id, "%s.%s%s" % (prefix, s.name, _indices(s.dims, flat))))
        for cmpsym, flat, sub, sig_off, cmp_off in inst.subs:
            walk(sub, base_sig + sig_off, cmp_id + 1 + cmp_off, "%s.%s%s" % (prefix, cmpsym.name, _indices(cmpsym.dims, flat)))

    walk(art.prog.main, 1, 0, "main")
    out.sort()
    return out


def write_artifact(art, outdir, with_cpp=False):
    os.makedirs(outdir, exist_ok=True)
    base = os.path.join(outdir, art.name)
    with open(base + ".cvm", "w") as f:
        f.write(art.cvm)
    formats.write_r1cs(base + ".r1cs", art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in,
                       art.witness, n_labels=art.n_signals)
    formats.write_sym(base + ".sym", sym_entries(art))
    paths = {"cvm": base + ".cvm", "r1cs": base + ".r1cs", "sym": base + ".sym"}
    if with_cpp:
        from .emit_cpp import emit_cpp
        with open(base + ".cpp", "w") as f:
            f.write(emit_cpp(art))
        consts = sorted(art.compiled.constants, key=art.compiled.constants.get)
        with open(base + ".dat", "wb") as f:
            f.write(formats.dat_bytes(art.main_inputs, art.witness, consts, io_map=art.compiled.io_map))
        paths.update(cpp=base + ".cpp", dat=base + ".dat")
    return paths
