"""Lower the circuit description to the reference's bucket IR.

Mirrors compiler/src/intermediate_representation/translate.rs: frame layout (template
arguments, then parameters, then declared vars: translate.rs:1932-1935), constants
initialised by Store buckets at the top of the body (translate.rs:232-298), component
creation hoisted before the first statement (translate.rs:1948), `<==` as Store + Assert(Eq)
(translate.rs:583-734), address arithmetic as AddAddress/MulAddress/ToAddress folded when
constant (translate.rs:1533-1648, ir_processing/reduce_stack.rs:28-55), and the
Last/NoLast/Unknown input-status analysis (ir_processing/build_inputs_info.rs).
Fixture tooling only.
"""
from __future__ import annotations

from .dsl import (Assert, Bin, CallE, Constrain, Function, If, Loop, NewCmp, Num, Ref, Return, Set, SigSet, Un, prod)
from .execute import CircuitError


# ------------------------------------------------------------------ IR
class Value:
    def __init__(self, kind, value): self.kind, self.value = kind, value     # 'u32' | 'ff'


class Load:
    def __init__(self, atype, loc, cmp=None, size=1):
        self.atype, self.loc, self.cmp, self.size = atype, loc, cmp, size      # 'var' | 'sig' | 'sub'


class Compute:
    def __init__(self, op, args): self.op, self.args = op, args


class Mapped:
    """LocationRule::Mapped (location_rule.rs:41-45): a signal of a component of a MIXED array (its instances are not
    all the same template instance), addressed through the io-map at run time: signal code + one index list."""
    def __init__(self, code, indexes, ndims): self.code, self.indexes, self.ndims = code, indexes, ndims


class Store:
    def __init__(self, atype, loc, cmp, src, size, line, cmp_name=None):
        self.atype, self.loc, self.cmp, self.src, self.size, self.line = atype, loc, cmp, src, size, line
        self.cmp_name = cmp_name
        self.status, self.needs_dec = None, False      # for 'sub' destinations
        self.sub_header = None                         # header of the sub-component's template (C++ run call)


class LoopB:
    def __init__(self, cond, body, line): self.cond, self.body, self.line = cond, body, line


class BranchB:
    def __init__(self, cond, then, other, line): self.cond, self.then, self.other, self.line = cond, then, other, line


class AssertB:
    def __init__(self, expr, line): self.expr, self.line = expr, line


class CreateCmpB:
    def __init__(self, **kw): self.__dict__.update(kw)


class CallB:
    def __init__(self, symbol, args, dest, line, arena):
        self.symbol, self.args, self.dest, self.line, self.arena = symbol, args, dest, line, arena


class ReturnB:
    def __init__(self, value, size, line): self.value, self.size, self.line = value, size, line


class TemplateCode:
    def __init__(self): pass


class FunctionCode:
    def __init__(self): pass


# ------------------------------------------------------------------ helpers
def u32(v):
    return Value("u32", v)


def fold_addr(terms):
    """AddAddress over terms with constant folding (reduce_stack.rs)."""
    const = 0
    rest = []
    for t in terms:
        if isinstance(t, Value) and t.kind == "u32":
            const += t.value
        else:
            rest.append(t)
    if not rest:
        return u32(const)
    acc = rest[0]
    for t in rest[1:]:
        acc = Compute("add_addr", [acc, t])
    if const:
        acc = Compute("add_addr", [acc, u32(const)])
    return acc


def io_signals(inst):
    """outputs then inputs in declaration order = the order of TemplateInstance::wires, whose positions are the signal
    codes (translate.rs:77-86; build.rs:531-552 keeps the non-intermediate ones)."""
    return [s for xt in ("out", "in") for s in inst.tmpl.signals if s.xtype == xt]


def signal_code(inst, name):
    for k, s in enumerate(io_signals(inst)):
        if s.name == name:
            return k
    raise CircuitError("%s is not an input/output of %s" % (name, inst.name))


class _Translator:
    def __init__(self, prog, body_owner, inst=None):
        self.prog, self.owner, self.inst = prog, body_owner, inst
        self.code = []
        self.max_arena = 0

    # ---- frame layout
    def layout_vars(self):
        off = 0
        for v in self.owner.vars:          # params were declared first by construction
            v.offset = off
            off += max(v.size, 1) if v.dims == () else v.size
        return off

    # ---- addresses
    def address(self, base, dims, idx):
        terms = []
        for k, i in enumerate(idx):
            stride = prod(dims[k + 1:])
            if isinstance(i, Num):
                terms.append(u32(stride * i.v))
            else:
                terms.append(Compute("mul_addr", [u32(stride), Compute("to_addr", [self.expr(i)])]))
        terms.append(u32(base))
        return fold_addr(terms), prod(dims[len(idx):])

    def location(self, ref):
        """-> (atype, loc, cmp, size, cmpsym, subinst)"""
        sym = ref.sym
        if sym.kind == "var":
            loc, n = self.address(sym.offset, sym.dims, ref.idx)
            return "var", loc, None, n, None, None
        if sym.kind == "sig":
            loc, n = self.address(sym.offset, sym.dims, ref.idx)
            return "sig", loc, None, n, None, None
        cmp_loc, cn = self.address(sym.offset, sym.dims, ref.idx)
        assert cn == 1
        subs = {id(s): s for s in sym.instances.values()}
        if len(subs) != 1:
            return self.mapped_location(ref, cmp_loc, list(subs.values()))
        sub = next(iter(subs.values()))
        ss = sub.sigsym[ref.sig]
        loc, n = self.address(ss.offset, ss.dims, ref.sigidx)
        return "sub", loc, cmp_loc, n, sym, (sub, ss)

    def mapped_location(self, ref, cmp_loc, subs):
        """translate.rs:1009-1041 (ClusterType::Mixed): the offset, the dimensions and the size of the signal come from
        the io-map entry of whichever template instance sits in the slot (build.rs:520-552)."""
        sym = ref.sym
        infos = [(signal_code(s, ref.sig), s.sigsym[ref.sig]) for s in subs]
        code, first = infos[0]
        if any(c != code or len(ss.dims) != len(first.dims) or ss.xtype != first.xtype for c, ss in infos):
            raise CircuitError("instances of the mixed array %s disagree on signal %s" % (sym.name, ref.sig))
        sizes = {prod(ss.dims[len(ref.sigidx):]) for _c, ss in infos}
        if len(sizes) != 1:
            raise CircuitError("copy of %s.%s whose length differs between instances (SizeOption::Multiple) is not "
                               "supported by the generator" % (sym.name, ref.sig))
        idx = []
        for i in ref.sigidx:
            idx.append(u32(i.v) if isinstance(i, Num) else Compute("to_addr", [self.expr(i)]))
        return "sub", Mapped(code, idx, len(first.dims)), cmp_loc, sizes.pop(), sym, (None, first)

    # ---- expressions
    def expr(self, e):
        if isinstance(e, Num):
            return Value("ff", e.v)
        if isinstance(e, Ref):
            atype, loc, cmp, n, _c, _s = self.location(e)
            return Load(atype, loc, cmp, n)
        if isinstance(e, Bin):
            return Compute(e.op, [self.expr(e.a), self.expr(e.b)])
        if isinstance(e, Un):
            return Compute(e.op, [self.expr(e.a)])
        if isinstance(e, CallE):
            raise CircuitError("function calls are only allowed as a whole right-hand side")
        raise CircuitError("bad expression")

    # ---- statements
    def store(self, dst, src_ir, line, src_size=1):
        atype, loc, cmp, n, csym, subinfo = self.location(dst)
        size = min(n, src_size) if src_size > 1 else (n if n == 1 else n)
        if n > 1 and src_size == 1:
            raise CircuitError("scalar stored into an array")
        st = Store(atype, loc, cmp, src_ir, size, line, cmp_name=csym.name if csym else None)
        if atype == "sub":
            sub, ss = subinfo
            if ss.xtype != "in":
                raise CircuitError("assignment to a non-input signal of a sub-component")
            st.sub_header = sub.header if sub is not None else None      # None: through _functionTable[templateId]
        return st

    def call(self, dst, e, line):
        fn = e.fn
        assert isinstance(fn, Function)
        fcode = self.prog.function_code(fn)
        args = []
        for a, p in zip(e.args, fn.params):
            if p.size > 1 or p.dims != ():
                assert isinstance(a, Ref)
                atype, loc, cmp, n, _c, _s = self.location(a)
                assert n == p.size, "argument size mismatch"
                args.append((Load(atype, loc, cmp, n), n))
            else:
                args.append((self.expr(a), 1))
        atype, loc, cmp, n, csym, subinfo = self.location(dst)
        dest = Store(atype, loc, cmp, None, n, line, cmp_name=csym.name if csym else None)
        if atype == "sub":
            dest.sub_header = subinfo[0].header if subinfo[0] is not None else None
        self.max_arena = max(self.max_arena, fcode.arena)
        return CallB(fcode.header, args, dest, line, fcode.arena)

    def block(self, stmts, out):
        for st in stmts:
            self.stmt(st, out)

    def stmt(self, st, out):
        if isinstance(st, Set) or isinstance(st, SigSet):
            if isinstance(st.src, CallE):
                out.append(self.call(st.dst, st.src, st.line))
            else:
                src = self.expr(st.src)
                ssize = src.size if isinstance(src, Load) else 1
                out.append(self.store(st.dst, src, st.line, ssize))
            if isinstance(st, SigSet) and st.constrain and not self.prog.constraint_assert_disabled:
                l, r = self.expr(st.dst), self.expr(st.src)
                n = l.size if isinstance(l, Load) else 1
                out.append(AssertB(Compute("eq", [l, r]) if n == 1 else Compute(("eq", n), [l, r]), st.line))
        elif isinstance(st, Constrain):
            if not self.prog.constraint_assert_disabled:
                out.append(AssertB(Compute("eq", [self.expr(st.l), self.expr(st.r)]), st.line))
        elif isinstance(st, Loop):
            body = []
            self.block(st.body, body)
            out.append(LoopB(self.expr(st.cond), body, st.line))
        elif isinstance(st, If):
            then, other = [], []
            self.block(st.then, then)
            self.block(st.other, other)
            out.append(BranchB(self.expr(st.cond), then, other, st.line))
        elif isinstance(st, Assert):
            out.append(AssertB(self.expr(st.e), st.line))
        elif isinstance(st, NewCmp):
            pass        # creation is hoisted (translate.rs:1948)
        elif isinstance(st, Return):
            v = self.expr(st.e)
            out.append(ReturnB(v, v.size if isinstance(v, Load) else 1, st.line))
        else:
            raise CircuitError("bad statement")


def _collect_values(node, acc):
    """gather ff constants (the reference interns them in a table: constant_tracking/src/lib.rs)."""
    if isinstance(node, Value):
        if node.kind == "ff":
            acc.setdefault(node.value, len(acc))
    elif isinstance(node, Mapped):
        for a in node.indexes: _collect_values(a, acc)
    elif isinstance(node, Load):
        _collect_values(node.loc, acc)
        if node.cmp is not None: _collect_values(node.cmp, acc)
    elif isinstance(node, Compute):
        for a in node.args: _collect_values(a, acc)
    elif isinstance(node, Store):
        _collect_values(node.loc, acc)
        if node.cmp is not None: _collect_values(node.cmp, acc)
        if node.src is not None: _collect_values(node.src, acc)
    elif isinstance(node, LoopB):
        _collect_values(node.cond, acc)
        for s in node.body: _collect_values(s, acc)
    elif isinstance(node, BranchB):
        _collect_values(node.cond, acc)
        for s in node.then + node.other: _collect_values(s, acc)
    elif isinstance(node, AssertB):
        _collect_values(node.expr, acc)
    elif isinstance(node, CallB):
        for a, _n in node.args: _collect_values(a, acc)
        _collect_values(node.dest, acc)
    elif isinstance(node, ReturnB):
        _collect_values(node.value, acc)


# ------------------------------------------------------------------ input status analysis
def build_inputs_info(code):
    """Last / NoLast / Unknown + needs_decrement, walking each list backwards
    (ir_processing/build_inputs_info.rs:31-47, 150-284)."""
    status = {}          # key -> [needs_decrement, found_last]
    unknown_names = set()

    def visit_list(lst, inside_loop):
        level = set()
        for node in reversed(lst):
            if isinstance(node, BranchB):
                visit_list(node.then, True)
                visit_list(node.other, True)
            elif isinstance(node, LoopB):
                visit_list(node.body, True)
            elif isinstance(node, Store) and node.atype == "sub":
                visit_store(node, level, inside_loop)
            elif isinstance(node, CallB) and node.dest.atype == "sub":
                visit_store(node.dest, level, inside_loop)

    def visit_store(st, level, inside_loop):
        if isinstance(st.cmp, Value):
            key = "cmp_%d" % st.cmp.value
            if key in status:
                info = status[key]
                if info[1]:
                    st.status, st.needs_dec = "nolast", info[0]
                elif key in level:
                    st.status, st.needs_dec = "nolast", True
                else:
                    st.status, st.needs_dec = "unknown", True
                    if not inside_loop:
                        info[1] = True
                    else:
                        level.add(key)
            else:
                if st.cmp_name not in unknown_names:
                    if inside_loop:
                        st.status, st.needs_dec = "unknown", True
                        status[key] = [True, False]
                        level.add(key)
                    else:
                        st.status, st.needs_dec = "last", False
                        status[key] = [False, True]
                else:
                    st.status, st.needs_dec = "unknown", True
                    status[key] = [True, not inside_loop]
                    if inside_loop:
                        level.add(key)
        else:
            st.status, st.needs_dec = "unknown", True
            unknown_names.add(st.cmp_name)

    visit_list(code, False)


# ------------------------------------------------------------------ drivers
def _position(dims, flat):
    out = []
    for d in reversed(dims):
        out.append(flat % d)
        flat //= d
    return "".join("[%d]" % i for i in reversed(out))


def translate_template(prog, inst):
    t = inst.tmpl
    tr = _Translator(prog, t, inst)
    frame = tr.layout_vars()
    code = []
    # constants: one Store per template-argument element (translate.rs:232-298)
    for v in t.vars:
        if v.is_param:
            for k, val in enumerate(v.init):
                code.append(Store("var", u32(v.offset + k), None, Value("ff", val), 1, 0))
    # component creation (translate.rs:365-442, uniform clusters)
    for c in t.components:
        subs = [s for s in inst.subs if s[0] is c]
        first = subs[0]
        kinds = {id(s[2]) for s in subs}
        if len(kinds) != 1:             # mixed cluster: one bucket per position (translate.rs:444-507)
            for (_c, flat, sub, sig_off, cmp_off) in subs:
                prog.templates_in_mixed.add(sub.id)
                code.append(CreateCmpB(
                    line=0, symbol=sub.header, template_id=sub.id, name=c.name + _position(c.dims, flat),
                    slot=c.offset + flat, signal_offset=sig_off, signal_offset_jump=0,
                    component_offset=cmp_off, component_offset_jump=0,
                    number_of_cmp=1, dimensions=list(c.dims), has_inputs=sub.n_in > 0))
            continue
        sub = first[2]
        code.append(CreateCmpB(
            line=0, symbol=sub.header, template_id=sub.id, name=c.name, slot=c.offset,
            signal_offset=first[3], signal_offset_jump=sub.n_signals,
            component_offset=first[4], component_offset_jump=sub.n_components,
            number_of_cmp=c.size, dimensions=list(c.dims), has_inputs=sub.n_in > 0))
    tr.block(t.body, code)
    build_inputs_info(code)
    out = TemplateCode()
    out.inst, out.header, out.name, out.id = inst, inst.header, inst.name, inst.id
    out.code = code
    out.frame = frame
    out.max_arena = tr.max_arena
    out.n_out, out.n_in, out.n_mid = inst.n_out, inst.n_in, inst.n_mid
    out.n_slots = inst.n_slots
    out.inputs = [s for s in t.signals if s.xtype == "in"]
    out.outputs = [s for s in t.signals if s.xtype == "out"]
    return out


def translate_function(prog, fn, fid):
    tr = _Translator(prog, fn)
    frame = tr.layout_vars()
    out = FunctionCode()
    out.header = "%s_%d" % (fn.name, fid)
    out.name = fn.name
    out.params = list(fn.params)
    out.returns = tuple(fn.returns)
    out.arena = frame
    prog._fcodes[fn.name] = out        # registered before the body so recursion resolves
    code = []
    tr.block(fn.body, code)
    out.code = code
    out.max_arena = tr.max_arena
    return out


class Compiled:
    """Everything the emitters need."""


def compile_program(prog, constraint_assert_disabled=False):
    prog.constraint_assert_disabled = constraint_assert_disabled
    prog._fcodes = {}
    prog.templates_in_mixed = set()

    def function_code(fn):
        if fn.name not in prog._fcodes:
            translate_function(prog, fn, len(prog._fcodes))
        return prog._fcodes[fn.name]

    prog.function_code = function_code
    for f in prog.functions.values():
        function_code(f)
    out = Compiled()
    out.prog = prog
    out.templates = [translate_template(prog, inst) for inst in prog.order]
    out.functions = list(prog._fcodes.values())
    consts = {}
    for t in out.templates:
        for n in t.code: _collect_values(n, consts)
    for f in out.functions:
        for n in f.code: _collect_values(n, consts)
    out.constants = consts          # value -> index in the constant table
    out.main = prog.main
    # io-map (build.rs:520-552): template instance id -> [(offset, dims, element size)] indexed by signal code
    out.io_map = {i: [(sg.offset, tuple(sg.dims), 1) for sg in io_signals(prog.order[i])]
                  for i in sorted(prog.templates_in_mixed)}
    return out
