"""Process-level drop-in for the reference's generated witness calculator (common/main.cpp:334-371):

    python -m circom_cvm_b200 <circuit.cvm> <input.json> <output.wtns> [--r1cs <circuit.r1cs>] [--sym <circuit.sym>]
                              [--cpp <circuit.cpp>]

reads `<circuit>.dat` next to the program (as the reference reads `<argv0>.dat`) - or, with --sym or when there is no
.dat, the `circom --sym` symbol file, or the program's own `;;%%main_input` lines - to resolve the input names, takes the same input.json and
writes the same bytes to output.wtns -- computed on the GPU.  input.json may also be an array of input objects:
the batch goes through one kernel launch and `<output>` gets one file per witness (`out.wtns`, `out.1.wtns`, ...).
Exit code 1 with the reference's message on a failing assert (the reference aborts).
"""
import os
import sys


def main(argv):
    args, opts, k = [], {}, 1
    while k < len(argv):
        if argv[k] in ("--r1cs", "--sym", "--cpp") and k + 1 < len(argv):
            opts[argv[k]] = argv[k + 1]
            k += 2
        elif argv[k].startswith("--"):
            args = []
            break
        else:
            args.append(argv[k])
            k += 1
    if len(args) != 3:
        print("Usage: python -m circom_cvm_b200 <circuit.cvm> <input.json> <output.wtns> [--r1cs <file>] [--sym <file>] [--cpp <file>]",
              file=sys.stderr)
        return 1
    from . import engine as E
    from .inputs import InputError, InputMap, NamedInputMap, SymInputMap, rows_from_json_text
    cvm, jin, wout = args
    r1cs_path = opts.get("--r1cs")
    # the generated C++ of the same compile (component creation for an unpatched emitter; section sizes of the .dat) and
    # the .dat (io-map of mixed component arrays) are picked up next to the program when they are there
    stem = os.path.splitext(cvm)[0]
    cpp = opts.get("--cpp") or (stem + ".cpp" if os.path.exists(stem + ".cpp") else None)
    dat = stem + ".dat" if cpp and os.path.exists(stem + ".dat") else None
    wc = E.WitnessCalculator(cvm_path=cvm, cpp_path=cpp, dat_path=dat)
    try:
        sym_path = opts.get("--sym") or (stem + ".sym" if not os.path.exists(stem + ".dat") and os.path.exists(stem + ".sym") else None)
        if sym_path:
            imap = SymInputMap.from_files(sym_path, wc)
        elif os.path.exists(stem + ".dat"):
            imap = InputMap.from_files(stem + ".dat", wc)
        else:                               # the program text itself (;;%%main_input lines)
            imap = NamedInputMap.from_program(wc)
        with open(jin) as f:
            rows = rows_from_json_text(imap, f.read())
    except InputError as e:
        print(str(e), file=sys.stderr)
        return 1
    if r1cs_path:
        wt, st, bad = wc.calculate_checked(rows, E.R1cs(r1cs_path))
    else:
        wt, st = wc.calculate(rows)
        bad = None
    rc = 0
    for k in range(len(rows)):
        if st[k] != 0:
            print("witness %d: failed assert / toInt / division (status %d)" % (k, int(st[k])), file=sys.stderr)
            rc = 1
            continue
        if bad is not None and bad[k] != E.NO_BAD:
            print("witness %d: constraint %d is not satisfied" % (k, int(bad[k])), file=sys.stderr)
            rc = 1
        base, ext = os.path.splitext(wout)
        wc.write_wtns(wout if k == 0 else "%s.%d%s" % (base, k, ext), wt[k])
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv))
