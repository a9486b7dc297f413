"""Print the bucket IR as Circom-Virtual-Machine text, in the shapes of the reference's
`impl WriteCVM` emitters (file:line cited per function).  Fixture tooling only.

Two modes.  faithful=True prints EXACTLY what the fork's emitters print, defects included: nothing for
component creation, literal addresses assigned to in copy loops, the array-equality loop that increments values as
addresses and returns a register nobody writes, the first element's value as the operand of a multi-element return.
The product's parser (csrc/cvm_parse.hpp) and the oracle honour each of these; component creation then comes from the
generated .cpp (cvmgpu_program_load_with_cpp).  The default mode keeps the departures the first round introduced:
  * CreateCmp emits nothing in the reference (create_component_bucket.rs:356-360), which makes
    a .cvm file non-executable on its own.  We emit one extension line per bucket,
        ;;%%create_cmp <slot> $<header> <sig_off> <sig_jump> <cmp_off> <cmp_jump> <n>
    -- a `;;` comment to any other consumer, a directive to ours.
  * copy loops increment their address operands textually in the reference, so literal
    addresses would be assigned to (`i64.5 = i64.add i64.5 i64.1`, store_bucket.rs:1026-1028);
    we first move literal addresses into fresh registers.
  * array equality (`===` on arrays) is emitted as a counted loop over address registers, see array_eq().
  * the io-map that "mapped" accesses (mixed component arrays) read is not in the fork's .cvm at all (circuit.rs:577-621);
    we emit one `;;%%io_map <template id> <n> {offset len lengths[1..] size busId}*` line per template instance.
  * the main component's input names are only in the .dat / .sym; we emit `;;%%main_input <name> <first signal> <size>`.
  * multi-element `return` passes the *address* register as documented in
    mkdocs/docs/circom-language/formats/circom-virtual-machine.md:201-203 (the emitter at
    return_bucket.rs:131 loads the first element instead, which cannot work).
"""
from __future__ import annotations

from .dsl import P
from .translate import (AssertB, BranchB, CallB, Compute, CreateCmpB, Load, LoopB, Mapped, ReturnB, Store, Value)

OPNAME = {
    "mul": ("ff.mul", "MUL"), "div": ("ff.div", "DIV"), "add": ("ff.add", "ADD"), "sub": ("ff.sub", "SUB"),
    "pow": ("ff.pow", "POW"), "idiv": ("ff.idiv", "INT_DIV"), "mod": ("ff.rem", "MOD"),
    "shl": ("ff.shl", "SHIFT_L"), "shr": ("ff.shr", "SHIFT_R"), "leq": ("ff.le", "LESSER_EQ"),
    "geq": ("ff.ge", "GREATER_EQ"), "lt": ("ff.lt", "LESSER"), "gt": ("ff.gt", "GREATER"),
    "neq": ("ff.neq", "NOT_EQ"), "lor": ("ff.or", "BOOL_OR"), "land": ("ff.and", "BOOL_AND"),
    "bor": ("ff.bor", "BITOR"), "band": ("ff.band", "BITAND"), "bxor": ("ff.bxor", "BITXOR"),
    "lnot": ("ff.eqz", "BOOL_NOT"), "bnot": ("ff.bnot", "COMPLEMENT"),
    "to_addr": ("ff.wrap_i64", "TO_ADDRESS"), "mul_addr": ("i64.mul", "MUL_ADDRESS"),
    "add_addr": ("i64.add", "ADD_ADDRESS"), "eq": ("ff.eq", ""),
}


def declare_variable(dims):        # cvm_code_generator.rs:1770-1783
    return "ff %d %s" % (len(dims), " ".join(str(d) for d in dims))


class CvmEmitter:
    def __init__(self, compiled, faithful=False):
        self.faithful = faithful
        self.c = compiled
        self.var_no = 0
        self.out = []

    def fresh(self):                # cvm_elements/mod.rs:202-206
        v = "x_%d" % self.var_no
        self.var_no += 1
        return v

    # ---- expressions -> (instructions, result operand)
    def value(self, n):             # value_bucket.rs:96-104
        return [], ("i64.%d" % n.value) if n.kind == "u32" else ("ff.%d" % n.value)

    def expr(self, n):
        if isinstance(n, Value):
            return self.value(n)
        if isinstance(n, Load):
            return self.load(n)
        if isinstance(n, Compute):
            return self.compute(n)
        raise TypeError(n)

    def location(self, atype, loc, cmp):        # location_rule.rs:66-85 (Indexed)
        if isinstance(loc, Mapped):
            return self.mapped(loc, cmp)
        ins, vloc = self.expr(loc)
        if atype == "sub":
            ins2, vcmp = self.expr(cmp)
            return ins + ins2, (vcmp, vloc)
        return ins, (None, vloc)

    def mapped(self, loc, cmp):                 # location_rule.rs:86-171 (Mapped; no bus accesses)
        ins = [";; is subcomponent mapped"]
        i2, vcmp = self.expr(cmp)
        ins += i2
        tid = self.fresh()
        ins.append("%s = get_template_id %s" % (tid, vcmp))
        sp = self.fresh()
        ins.append("%s = get_template_signal_position %s %d" % (sp, tid, loc.code))
        if not loc.indexes:
            return ins, (vcmp, sp)
        i3, prev = self.expr(loc.indexes[0])
        ins += i3
        for i in range(1, len(loc.indexes)):
            dimi = self.fresh()
            ins.append("%s =  get_template_signal_dimension %s %d %d" % (dimi, tid, loc.code, i))
            i4, vidx = self.expr(loc.indexes[i])
            ins += i4
            curmul = self.fresh()
            ins.append("%s = i64.mul %s %s" % (curmul, prev, dimi))
            cursize = self.fresh()
            ins.append("%s = i64.add %s %s" % (cursize, curmul, vidx))
            prev = cursize
        diff = loc.ndims - len(loc.indexes)
        if diff > 0:
            # as printed: diff-1 factors starting at dimension <number of accesses> (= 1 without buses), :135-140
            for i in range(diff - 1):
                dimi = self.fresh()
                ins.append("%s =  get_template_signal_dimension %s %d %d" % (dimi, tid, loc.code, 1 + i))
                cursize = self.fresh()
                ins.append("%s = i64.mul %s %s" % (cursize, prev, dimi))
                prev = cursize
        vsize = self.fresh()
        ins.append("%s =  get_template_signal_size %s %d" % (vsize, tid, loc.code))
        finalsize = self.fresh()
        ins.append("%s = i64.mul %s %s" % (finalsize, prev, vsize))
        access = self.fresh()
        ins.append("%s = i64.add %s %s" % (access, sp, prev))
        ins.append(";; end of load bucket")
        return ins, (vcmp, access)

    def load(self, n):              # load_bucket.rs:459-485
        ins = [";; load bucket"]
        i2, (vcmp, vloc) = self.location(n.atype, n.loc, n.cmp)
        ins += i2
        ins.append(";; end of load bucket")
        res = self.fresh()
        ins.append(self._get(n.atype, res, vcmp, vloc))
        return ins, res

    @staticmethod
    def _get(atype, res, vcmp, vloc):
        if atype == "var":
            return "%s = ff.load %s" % (res, vloc)
        if atype == "sig":
            return "%s = get_signal %s" % (res, vloc)
        return "%s = get_cmp_signal %s %s" % (res, vcmp, vloc)

    def compute(self, n):           # compute_bucket.rs:471-620
        op = n.op
        if isinstance(op, tuple):
            return self.array_eq(n, op[1])
        ins = [";; compute bucket"]
        vres = []
        for a in n.args:
            i2, r = self.expr(a)
            ins += i2
            vres.append(r)
        if op == "neg":
            ins.append(";; OP(PREFIX_SUB)")
            res = self.fresh()
            ins.append("%s = ff.sub 0 %s" % (res, vres[0]))
        else:
            mnem, name = OPNAME[op]
            ins.append(";; OP(%s)" % name)
            res = self.fresh()
            ins.append("%s = %s %s" % (res, mnem, " ".join(vres)))
        ins.append(";; end of compute bucket")
        return ins, res

    def array_eq(self, n, size):
        """Element-wise equality of two arrays (Eq(n>1)).  The fork's emitter for this case
        (compute_bucket.rs:538-586) loads the FIRST ELEMENT VALUES and then increments those values as if
        they were addresses, so its output cannot be executed.  We emit what the C++ twin computes
        (compute_bucket.rs:430-455: a loop of Fr_eq over both arrays): a counted loop over address
        registers that ANDs the per-element results, without the data-dependent early exit."""
        if self.faithful:
            # compute_bucket.rs:471-489, 538-586 verbatim: the operands are the LOADED first elements; `res` (returned) is
            # allocated first and never written, the loop writes the third fresh variable
            ins = [";; compute bucket"]
            vres = []
            for a in n.args:
                i2, r = self.expr(a)
                ins += i2
                vres.append(r)
            ins.append(";; OP(EQ)")
            res = self.fresh()
            counter = self.fresh()
            res2 = self.fresh()
            ins.append("%s = i64.%d" % (counter, size))
            ins += ["loop", "if %s " % counter, "%s = ff.eq %s" % (res2, " ".join(vres)), "if %s " % res2,
                    "%s = i64.sub %s i64.1" % (counter, counter), "%s = i64.add %s i64.1" % (vres[0], vres[0]),
                    "%s = i64.add %s i64.1" % (vres[1], vres[1]), "continue", "end", "end", "break", "end",
                    ";; end of compute bucket"]
            return ins, res
        ins = [";; compute bucket"]
        locs = []
        for a in n.args:
            assert isinstance(a, Load)
            i2, (vcmp, vloc) = self.location(a.atype, a.loc, a.cmp)
            ins += i2
            r = self.fresh()
            ins.append("%s = %s" % (r, vloc))
            locs.append((a.atype, vcmp, r))
        ins.append(";; OP()")
        counter, acc = self.fresh(), self.fresh()
        ins.append("%s = i64.%d" % (counter, size))
        ins.append("%s = ff.1" % acc)
        ins += ["loop", "if %s " % counter]
        vals = []
        for (atype, vcmp, r) in locs:
            v = self.fresh()
            ins.append(self._get(atype, v, vcmp, r))
            vals.append(v)
        e = self.fresh()
        ins.append("%s = ff.eq %s %s" % (e, vals[0], vals[1]))
        ins.append("%s = ff.and %s %s" % (acc, acc, e))
        ins.append("%s = i64.sub %s i64.1" % (counter, counter))
        for (_a, _c, r) in locs:
            ins.append("%s = i64.add %s i64.1" % (r, r))
        ins += ["continue", "end", "break", "end", ";; end of compute bucket"]
        return ins, acc

    # ---- statements
    @staticmethod
    def _set_cmp(st, vcmp, vloc, vsrc, final=True):
        # store_bucket.rs:875-897 ; names cvm_code_generator.rs:97-119
        if st.status == "nolast":
            return ("set_cmp_input_cnt %s %s %s" if st.needs_dec else "set_cmp_input %s %s %s") % (vcmp, vloc, vsrc)
        if st.status == "last":
            return "set_cmp_input_run %s %s %s" % (vcmp, vloc, vsrc)
        return "set_cmp_input_cnt_check %s %s %s" % (vcmp, vloc, vsrc)

    def store(self, st):            # store_bucket.rs:817-1045
        ins = [";; store bucket. Line %d" % st.line]
        if st.size == 1:
            ins.append(";; getting src")
            i2, vsrc = self.expr(st.src)
            ins += i2
            ins.append(";; getting dest")
            i3, (vcmp, vloc) = self.location(st.atype, st.loc, st.cmp)
            ins += i3
            if st.atype == "var":
                ins.append("ff.store %s %s" % (vloc, vsrc))
            elif st.atype == "sig":
                ins.append("set_signal %s %s" % (vloc, vsrc))
            else:
                ins.append(self._set_cmp(st, vcmp, vloc, vsrc))
        else:
            src = st.src
            assert isinstance(src, Load), "multi-element store needs a load source"
            ins.append(";; getting src")
            i2, (scmp, sloc) = self.location(src.atype, src.loc, src.cmp)
            ins += i2
            src_value = self.fresh()
            get_src = self._get(src.atype, src_value, scmp, sloc)
            ins.append(";; getting dest")
            i3, (vcmp, vloc) = self.location(st.atype, st.loc, st.cmp)
            ins += i3
            counter = self.fresh()
            ins += self._copy_loop(st, counter, st.size, get_src, src_value, sloc, vcmp, vloc)
        ins.append(";; end of store bucket")
        return ins

    def _copy_loop(self, st, counter, n, get_src, src_value, sloc, vcmp, vloc, call=False):
        """counter loop of single-element copies; the last element of a Last/Unknown sub-component
        input is peeled out so that only it can trigger the run (store_bucket.rs:944-1035)."""
        ins = []
        # The reference emitter increments its address operands textually, so a literal address
        # yields `i64.5 = i64.add i64.5 i64.1` (store_bucket.rs:1026-1028).  We materialise literal
        # addresses into fresh registers first, which is what that code means.
        if not self.faithful and not sloc.startswith("x_"):
            r = self.fresh()
            ins.append("%s = %s" % (r, sloc))
            get_src = get_src.replace(" " + sloc, " " + r) if get_src.endswith(" " + sloc) else get_src
            sloc = r
        if not self.faithful and not vloc.startswith("x_"):
            r = self.fresh()
            ins.append("%s = %s" % (r, vloc))
            vloc = r
        last_out, last_ins = False, []
        if st.atype == "var":
            set_dest = "ff.store %s %s" % (vloc, src_value)
        elif st.atype == "sig":
            set_dest = "set_signal %s %s" % (vloc, src_value)
        elif st.status == "nolast":
            set_dest = self._set_cmp(st, vcmp, vloc, src_value)
        elif st.status == "last":
            last_out = True
            set_dest = "set_cmp_input %s %s %s" % (vcmp, vloc, src_value)
            last_ins = [get_src, "set_cmp_input_run %s %s %s" % (vcmp, vloc, src_value)]
        else:
            last_out = True
            set_dest = "set_cmp_input_cnt %s %s %s" % (vcmp, vloc, src_value)
            last_ins = [get_src, "set_cmp_input_cnt_check %s %s %s" % (vcmp, vloc, src_value)]
        if call and self.faithful:       # call_bucket.rs:960-972: the full size, then one subtracted for a peeled last element
            ins.append("%s = i64.%d" % (counter, n))
            if last_out:
                ins.append("%s = i64.sub %s i64.1" % (counter, counter))
        else:
            ins.append("%s = i64.%d" % (counter, n - 1 if last_out else n))
        ins += ["loop", "if %s " % counter, get_src, set_dest,
                "%s = i64.sub %s i64.1" % (counter, counter),
                "%s = i64.add %s i64.1" % (sloc, sloc),
                "%s = i64.add %s i64.1" % (vloc, vloc),
                "continue", "end", "break", "end"]
        ins += last_ins
        return ins

    def call(self, n):              # call_bucket.rs:849-1003
        ins = [";; start of call bucket"]
        params = ""
        for k, (a, size) in enumerate(n.args):
            if size > 1:
                i2, (acmp, aloc) = self.location(a.atype, a.loc, a.cmp)
                ins += i2
                if a.atype == "var":
                    params += " i64.memory(%s,%d)" % (aloc, size)
                elif a.atype == "sig":
                    params += " signal(%s,%d)" % (aloc, size)
                else:
                    params += " subcmpsignal(%s,%s,%d)" % (acmp, aloc, size)
            else:
                i2, r = self.expr(a)
                ins += i2
                params += " %s" % r
            ins.append("// end copying argument %d" % k)
        d = n.dest
        call_dest = self.fresh()
        i3, (vcmp, vloc) = self.location(d.atype, d.loc, d.cmp)
        ins.append("%s = spr" % call_dest)
        ins.append("ff.call $%s %s i64.%d %s" % (n.symbol, call_dest, d.size, params))
        ins += i3
        src_value = self.fresh()
        get_src = "%s = ff.load %s" % (src_value, call_dest)
        counter = self.fresh()
        # call_bucket.rs:960-990: counter = size, minus one when the last element is peeled
        loop = self._copy_loop(d, counter, d.size, get_src, src_value, call_dest, vcmp, vloc, call=True)
        ins += loop
        ins.append("// end call bucket")
        return ins

    def stmt(self, n):
        if isinstance(n, Store):
            return self.store(n)
        if isinstance(n, LoopB):    # loop_bucket.rs:97-120
            ins = [";; loop bucket. Line %d" % n.line, "loop"]
            i2, vcond = self.expr(n.cond)
            ins += i2
            ins.append("if %s" % vcond)
            for s in n.body:
                ins += self.stmt(s)
            ins += ["continue", "end", "end", ";; end of loop bucket"]
            return ins
        if isinstance(n, BranchB):  # branch_bucket.rs:126-168
            ins = [";; branch bucket"]
            if n.then:
                i2, vcond = self.expr(n.cond)
                ins += i2
                ins.append("if %s" % vcond)
                for s in n.then:
                    ins += self.stmt(s)
                if n.other:
                    ins.append("else")
                    for s in n.other:
                        ins += self.stmt(s)
                ins.append("end")
            elif n.other:
                i2, vcond = self.expr(n.cond)
                ins += i2
                res = self.fresh()
                ins.append("%s = ff.eqz %s" % (res, vcond))
                ins.append("if %s" % res)
                for s in n.other:
                    ins += self.stmt(s)
                ins.append("end")
            ins.append(";; end of branch bucket")
            return ins
        if isinstance(n, AssertB):  # assert_bucket.rs:88-107
            ins = [";; assert bucket"]
            i2, avar = self.expr(n.expr)
            ins += i2
            cvar = self.fresh()
            ins += ["%s = ff.eqz %s" % (cvar, avar), "if %s" % cvar, "error 0", "end", ";; end of assert bucket"]
            return ins
        if isinstance(n, CreateCmpB) and self.faithful:
            return []                   # create_component_bucket.rs:356-360
        if isinstance(n, CreateCmpB):
            return [";;%%%%create_cmp %d $%s %d %d %d %d %d" % (
                n.slot, n.symbol, n.signal_offset, n.signal_offset_jump, n.component_offset,
                n.component_offset_jump, n.number_of_cmp)]
        if isinstance(n, CallB):
            return self.call(n)
        if isinstance(n, ReturnB):  # return_bucket.rs:125-147
            ins = ["// return bucket"]
            if n.size == 1:
                i2, src = self.expr(n.value)
                ins += i2
                ins.append("return %s 1" % src)
            else:
                if self.faithful:       # return_bucket.rs:131: the value bucket is evaluated, i.e. the first element is loaded
                    i2, aloc = self.expr(n.value)
                else:
                    i2, (_c, aloc) = self.location(n.value.atype, n.value.loc, n.value.cmp)
                ins += i2
                vcond, final = self.fresh(), self.fresh()
                ins += ["%s = i64.le %d destination_size" % (vcond, n.size), "if %s" % vcond,
                        "%s = %d" % (final, n.size), "else", "%s = destination_size" % final, "end",
                        "return %s %s" % (aloc, final)]
            return ins
        raise TypeError(n)

    # ---- file
    def emit(self):
        c = self.c
        main = c.main
        total_signals = main.n_signals + 1                     # build.rs:242
        heap = 3 * main.n_components                           # build.rs:241 (no extra indexes needed)
        o = self.out
        o += [";; Prime value", "%%%%prime %d" % P, "\n"]
        o += [";; Memory of signals", "%%%%signals %d" % total_signals, "\n"]
        o += [";; Heap of components", "%%%%components_heap %d" % heap, "\n"]
        o += [";; Types (for each field we store name type offset size nDims dims)", "\n"]
        o += [";; Main template", "%%%%start %s" % main.header, "\n"]
        o += [";; Component creation mode (implicit/explicit)", "%%components explicit", "\n"]
        o += [";; Witness (signal list)", "%%witness" + "".join(" %d" % s for s in c.witness), "\n"]
        if not self.faithful:       # extension (a comment to other consumers): the main-input name table of the .dat
            for sg in main.tmpl.signals:
                if sg.xtype == "in":
                    o.append(";;%%%%main_input %s %d %d" % (sg.name, 1 + sg.offset, sg.size))
        if not self.faithful:       # extension (a comment to other consumers): the .dat io-map records as text
            for tid, defs in sorted(getattr(c, "io_map", {}).items()):
                rec = "".join(" %d %d%s %d 0" % (off, max(len(dims) - 1, 0), "".join(" %d" % d for d in dims[1:]), size)
                              for off, dims, size in defs)
                o.append(";;%%%%io_map %d %d%s" % (tid, len(defs), rec))
        for f in c.functions:      # function.rs:137-168
            ins = "".join(" " + declare_variable(p.dims) for p in f.params)
            o.append("%%%%function %s [%s] [%s]" % (f.header, declare_variable(f.returns), ins))
            o.append("local.memory %d" % f.arena)
            for s in f.code:
                o += self.stmt(s)
        for t in c.templates:      # template.rs:158-208 (bracket 1 = Input wires, bracket 2 = Output wires)
            b1 = "".join(" " + declare_variable(s.dims) for s in t.inputs)
            b2 = "".join(" " + declare_variable(s.dims) for s in t.outputs)
            o.append("%%%%template %s [%s] [%s] [%d] [%d]" % (
                t.header, b1, b2, t.n_out + t.n_in + t.n_mid, t.n_slots))
            for s in t.code:
                o += self.stmt(s)
            o.append("")
        return "\n".join(o) + "\n"


def emit_cvm(compiled, faithful=False):
    return CvmEmitter(compiled, faithful).emit()
