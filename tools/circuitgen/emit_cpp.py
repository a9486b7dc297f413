"""Print the bucket IR as C++ in the shapes of the reference's `impl WriteC` emitters, so that the same
program can be linked against the REFERENCE's own runtime (common/main.cpp, calcwit.cpp, generic/fr.cpp)
and run as the reference would run it.  This is how program-level parity with the reference runtime is
pinned (tests/test_ref_runtime.py) and how the CPU baseline is produced (bench.py).  Fixture tooling only.

Shapes followed (compiler/src/...):
  file skeleton, get_* functions, run()     circuit_design/circuit.rs:424-567
  T_create / T_run                          circuit_design/template.rs:225-407
  functions                                 circuit_design/function.rs:94-135
  Value / Load / Compute / Store / ...      intermediate_representation/{value,load,compute,store,branch,loop,
                                            assert,call,return,create_component}_bucket.rs  (impl WriteC)
"""
from __future__ import annotations

from .translate import (AssertB, BranchB, CallB, Compute, CreateCmpB, Load, LoopB, Mapped, ReturnB, Store, Value)

FR_OP = {"add": "Fr_add", "div": "Fr_div", "mul": "Fr_mul", "sub": "Fr_sub", "pow": "Fr_pow", "idiv": "Fr_idiv",
         "mod": "Fr_mod", "shl": "Fr_shl", "shr": "Fr_shr", "leq": "Fr_leq", "geq": "Fr_geq", "lt": "Fr_lt",
         "gt": "Fr_gt", "eq": "Fr_eq", "neq": "Fr_neq", "lor": "Fr_lor", "land": "Fr_land", "bor": "Fr_bor",
         "band": "Fr_band", "bxor": "Fr_bxor", "neg": "Fr_neg", "lnot": "Fr_lnot", "bnot": "Fr_bnot"}


class CppEmitter:
    def __init__(self, art):
        self.art = art
        self.c = art.compiled
        self.max_aux = 0

    # ---- expressions -> (prologue lines, C expression).  `depth` = first free expaux slot.
    def expr(self, n, depth):
        if isinstance(n, Value):
            if n.kind == "u32":
                return [], str(n.value)
            return [], "&circuitConstants[%d]" % self.c.constants[n.value]      # value_bucket.rs:77-93
        if isinstance(n, Load):
            return self.location(n.atype, n.loc, n.cmp, depth)
        if isinstance(n, Compute):
            return self.compute(n, depth)
        raise TypeError(n)

    def location(self, atype, loc, cmp, depth):           # load_bucket.rs:249-457
        if isinstance(loc, Mapped):
            p2, c = self.expr(cmp, depth)
            pro, idx = self.mapped_load(loc, c, depth)
            return p2 + pro, "&ctx->signalValues[ctx->componentMemory[mySubcomponents[%s]].signalStart + %s]" % (c, idx)
        pro, idx = self.expr(loc, depth)
        if atype == "var":
            return pro, "&lvar[%s]" % idx
        if atype == "sig":
            return pro, "&signalValues[mySignalStart + %s]" % idx
        p2, c = self.expr(cmp, depth)
        return pro + p2, "&ctx->signalValues[ctx->componentMemory[mySubcomponents[%s]].signalStart + %s]" % (c, idx)

    def mapped_load(self, loc, cmp_expr, depth):           # load_bucket.rs:262-318 (no bus accesses)
        cur_def = "ctx->templateInsId2IOSignalInfo[ctx->componentMemory[mySubcomponents[%s]].templateId].defs[%d]" % (cmp_expr, loc.code)
        access = cur_def + ".offset"
        pro = []
        if loc.indexes:
            p0, map_index = self.expr(loc.indexes[0], depth)
            pro += p0
            for i in range(1, len(loc.indexes)):
                pi, e = self.expr(loc.indexes[i], depth)
                pro += pi
                map_index = "(%s)*(%s.lengths[%d])+%s" % (map_index, cur_def, i - 1, e)
            if loc.ndims - len(loc.indexes) > 0:
                pro.append("//There is a difference %d;" % (loc.ndims - len(loc.indexes)))
                for i in range(len(loc.indexes), loc.ndims):
                    map_index = "%s*%s.lengths[%d]" % (map_index, cur_def, i - 1)
            access = "%s+(%s)*%s.size" % (access, map_index, cur_def)
        return pro, access

    def mapped_store(self, loc, cmp_expr):                  # store_bucket.rs:500-566 (no bus accesses)
        tid = "ctx->componentMemory[mySubcomponents[%s]].templateId" % cmp_expr
        access = "ctx->templateInsId2IOSignalInfo[%s].defs[%d].offset" % (tid, loc.code)
        pro = []
        if loc.indexes:
            pro += ["{", "uint map_accesses_aux[1];", "{",
                    "IOFieldDef *cur_def = &(ctx->templateInsId2IOSignalInfo[%s].defs[%d]);" % (tid, loc.code),
                    "{", "uint map_index_aux[%d];" % len(loc.indexes)]
            p0, e0 = self.expr(loc.indexes[0], 0)
            pro += p0
            pro.append("map_index_aux[0]=%s;" % e0)
            map_index = "map_index_aux[0]"
            for i in range(1, len(loc.indexes)):
                pi, e = self.expr(loc.indexes[i], 0)
                pro += pi
                pro.append("map_index_aux[%d]=%s;" % (i, e))
                map_index = "(%s)*cur_def->lengths[%d]+map_index_aux[%d]" % (map_index, i - 1, i)
            if loc.ndims - len(loc.indexes) > 0:
                pro.append("//There is a difference %d;" % (loc.ndims - len(loc.indexes)))
                for i in range(len(loc.indexes), loc.ndims):
                    map_index = "%s*cur_def->lengths[%d]" % (map_index, i - 1)
            pro.append("map_accesses_aux[0] = %s*cur_def->size;" % map_index)
            pro += ["}", "}"]
            access += "+map_accesses_aux[0]"
        return pro, access, bool(loc.indexes)

    def sub_dest(self, d, out):
        """destination inside a sub-component (cmp_index_ref already declared) -> (C expression, close the extra block?)"""
        if isinstance(d.loc, Mapped):
            pro, idx, opened = self.mapped_store(d.loc, "cmp_index_ref")
        else:
            pro, idx = self.expr(d.loc, 0)
            opened = False
        out += pro
        return "&ctx->signalValues[ctx->componentMemory[mySubcomponents[cmp_index_ref]].signalStart + %s]" % idx, opened

    def compute(self, n, depth):                          # compute_bucket.rs:314-469
        op = n.op
        multi = isinstance(op, tuple)
        key = op[0] if multi else op
        if key in ("add_addr", "mul_addr", "to_addr"):
            pro, ops = [], []
            for a in n.args:
                p, e = self.expr(a, depth)
                pro += p
                ops.append(e)
            if key == "add_addr":
                return pro, "(%s + %s)" % (ops[0], ops[1])
            if key == "mul_addr":
                return pro, "(%s * %s)" % (ops[0], ops[1])
            return pro, "Fr_toInt(%s)" % ops[0]
        self.max_aux = max(self.max_aux, depth + 1)
        pro, ops = [], []
        for k, a in enumerate(n.args):
            p, e = self.expr(a, depth + 1 + k)
            pro += p
            ops.append(e)
        res = "&expaux[%d]" % depth
        pro.append("%s(%s);" % (FR_OP[key], ",".join([res] + ops)))
        if multi:                                          # array equality: compute_bucket.rs:375-407
            pro.append("{ uint index_multiple_eq = 1;")
            pro.append("while(index_multiple_eq < %d && Fr_isTrue(%s)) {" % (op[1], res))
            pro.append("Fr_eq(%s,%s + index_multiple_eq,%s + index_multiple_eq);" % (res, ops[0], ops[1]))
            pro.append("index_multiple_eq++;")
            pro.append("} }")
        return pro, res

    # ---- statements
    def trigger(self, st, cmp_expr, size):
        """inputCounter bookkeeping and the synchronous run of the sub-component (store_bucket.rs:662-800)."""
        counter = "ctx->componentMemory[mySubcomponents[%s]].inputCounter" % cmp_expr
        out = []
        if st.status == "nolast":
            if st.needs_dec:
                out.append("%s -= %d;" % (counter, size))
                out.append("assert(%s > 0);" % counter)
            return out
        if st.sub_header is None:        # mapped destination: through the table of run functions (store_bucket.rs:706-710)
            call = "(*_functionTable[ctx->componentMemory[mySubcomponents[%s]].templateId])(mySubcomponents[%s],ctx);" % (cmp_expr, cmp_expr)
        else:
            call = "%s_run(mySubcomponents[%s],ctx);" % (st.sub_header, cmp_expr)
        if st.status == "unknown":
            out.append("if(!(%s -= %d)){" % (counter, size))
            out.append(call)
            out.append("}")
        else:
            if st.needs_dec:
                out.append("%s -= %d;" % (counter, size))
                out.append("assert(!(%s));" % counter)
            out.append(call)
        return out

    def store(self, st):                                   # store_bucket.rs:444-814
        out = ["{"]
        cmp_expr = None
        if st.atype == "sub":
            p, c = self.expr(st.cmp, 0)
            out += p
            out.append("uint cmp_index_ref = %s;" % c)
            cmp_expr = "cmp_index_ref"
            dest, opened = self.sub_dest(st, out)
        else:
            opened = False
            pro, dest = self.location(st.atype, st.loc, None, 0)
            out += pro
        out.append("PFrElement aux_dest = %s;" % dest)
        out.append("// load src")
        p, src = self.expr(st.src, 0)
        out += p
        out.append("// end load src")
        if st.size > 1:
            out.append("Fr_copyn(aux_dest,%s,%d);" % (src, st.size))
        else:
            out.append("Fr_copy(aux_dest,%s);" % src)
        if st.atype == "sub":
            out += self.trigger(st, cmp_expr, st.size)
        if opened:
            out.append("}")             # the map_accesses_aux block (store_bucket.rs:510,565 + :809)
        out.append("}")
        return out

    def call(self, n):                                     # call_bucket.rs:465-847
        out = ["{", "FrElement lvarcall[%d];" % max(1, n.arena)]
        pos = 0
        for a, size in n.args:
            p, src = self.expr(a, 0)
            out += p
            if size > 1:
                out.append("Fr_copyn(&lvarcall[%d],%s,%d);" % (pos, src, size))
            else:
                out.append("Fr_copy(&lvarcall[%d],%s);" % (pos, src))
            pos += size
        d = n.dest
        cmp_expr = None
        if d.atype == "sub":
            p, c = self.expr(d.cmp, 0)
            out += p
            out.append("uint cmp_index_ref = %s;" % c)
            cmp_expr = "cmp_index_ref"
            dest, opened = self.sub_dest(d, out)
        else:
            opened = False
            pro, dest = self.location(d.atype, d.loc, None, 0)
            out += pro
        out.append("%s(ctx,lvarcall,myId,%s,%d);" % (n.symbol, dest, d.size))
        if d.atype == "sub":
            out += self.trigger(d, cmp_expr, d.size)
        if opened:
            out.append("}")
        out.append("}")
        return out

    def stmt(self, n, name):
        if isinstance(n, Store):
            return self.store(n)
        if isinstance(n, LoopB):                           # loop_bucket.rs:78-94
            p, c = self.expr(n.cond, 0)
            out = list(p)
            out.append("while(Fr_isTrue(%s)){" % c)
            for s in n.body:
                out += self.stmt(s, name)
            out += p                                       # the condition is re-evaluated at the end of the body
            out.append("}")
            return out
        if isinstance(n, BranchB):                         # branch_bucket.rs:101-124
            p, c = self.expr(n.cond, 0)
            out = list(p)
            out.append("if(Fr_isTrue(%s)){" % c)
            for s in n.then:
                out += self.stmt(s, name)
            out.append("}else{")
            for s in n.other:
                out += self.stmt(s, name)
            out.append("}")
            return out
        if isinstance(n, AssertB):                         # assert_bucket.rs:71-86
            p, c = self.expr(n.expr, 0)
            out = ["{"] + p
            out.append('if (!Fr_isTrue(%s)) std::cout << "Failed assert in template/function " << myTemplateName << '
                       '" line %d. " <<  "Followed trace of components: " << ctx->getTrace(myId) << std::endl;' % (c, n.line))
            out.append("assert(Fr_isTrue(%s));" % c)
            out.append("}")
            return out
        if isinstance(n, CreateCmpB):                      # create_component_bucket.rs:206-354
            out = ["{"]
            if n.number_of_cmp > 1:
                out.append("uint aux_create = %d;" % n.slot)
                out.append("int aux_cmp_num = %d+ctx_index+1;" % n.component_offset)
                out.append("uint csoffset = mySignalStart+%d;" % n.signal_offset)
                dims = n.dimensions or [n.number_of_cmp]
                out.append("uint aux_dimensions[%d] = {%s};" % (len(dims), ",".join(str(d) for d in dims)))
                out.append("for (uint i = 0; i < %d; i++) {" % n.number_of_cmp)
                out.append('std::string new_cmp_name = "%s"+ctx->generate_position_array(aux_dimensions, %d, i);'
                           % (n.name, len(dims)))
                out.append("%s_create(csoffset,aux_cmp_num,ctx,new_cmp_name,myId);" % n.symbol)
                out.append("mySubcomponents[aux_create+ i] = aux_cmp_num;")
                out.append("csoffset += %d ;" % n.signal_offset_jump)
                out.append("aux_cmp_num += %d;" % n.component_offset_jump)
                out.append("}")
            else:
                out.append('std::string new_cmp_name = "%s";' % n.name)
                out.append("%s_create(mySignalStart+%d,%d+ctx_index+1,ctx,new_cmp_name,myId);"
                           % (n.symbol, n.signal_offset, n.component_offset))
                out.append("mySubcomponents[%d] = %d+ctx_index+1;" % (n.slot, n.component_offset))
            out.append("}")
            return out
        if isinstance(n, CallB):
            return self.call(n)
        if isinstance(n, ReturnB):                         # return_bucket.rs:98-122
            p, src = self.expr(n.value, 0)
            out = list(p)
            if n.size == 1:
                out.append("Fr_copy(destination,%s);" % src)
            else:
                out.append("Fr_copyn(destination,%s,std::min(%d,destination_size));" % (src, n.size))
            out.append("return;")
            return out
        raise TypeError(n)

    # ---- units
    def function(self, f):
        self.max_aux = 0
        body = []
        for s in f.code:
            body += self.stmt(s, f.name)
        head = ["void %s(Circom_CalcWit* ctx,FrElement* lvar,uint componentFather,FrElement* destination,int destination_size){"
                % f.header,
                "FrElement* circuitConstants = ctx->circuitConstants;",
                "FrElement expaux[%d];" % max(1, self.max_aux),
                'std::string myTemplateName = "%s";' % f.name,
                "u64 myId = componentFather;"]
        return head + body + ["}", ""]

    def template(self, t):
        self.max_aux = 0
        body = []
        for s in t.code:
            body += self.stmt(s, t.name)
        n_in = t.n_in
        create = [
            "void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather){" % t.header,
            "ctx->componentMemory[coffset].templateId = %d;" % t.id,
            'ctx->componentMemory[coffset].templateName = "%s";' % t.name,
            "ctx->componentMemory[coffset].signalStart = soffset;",
            "ctx->componentMemory[coffset].inputCounter = %d;" % n_in,
            "ctx->componentMemory[coffset].componentName = componentName;",
            "ctx->componentMemory[coffset].idFather = componentFather;",
            ("ctx->componentMemory[coffset].subcomponents = new uint[%d]{0};" % t.n_slots) if t.n_slots > 0
            else "ctx->componentMemory[coffset].subcomponents = new uint[0];",
        ]
        if n_in == 0:
            create.append("%s_run(coffset,ctx);" % t.header)
        create.append("}")
        run = [
            "void %s_run(uint ctx_index,Circom_CalcWit* ctx){" % t.header,
            "FrElement* circuitConstants = ctx->circuitConstants;",
            "FrElement* signalValues = ctx->signalValues;",
            "FrElement expaux[%d];" % max(1, self.max_aux),
            "FrElement lvar[%d];" % max(1, t.frame),
            "u64 mySignalStart = ctx->componentMemory[ctx_index].signalStart;",
            "std::string myTemplateName = ctx->componentMemory[ctx_index].templateName;",
            "std::string myComponentName = ctx->componentMemory[ctx_index].componentName;",
            "u64 myFather = ctx->componentMemory[ctx_index].idFather;",
            "u64 myId = ctx_index;",
            "u32* mySubcomponents = ctx->componentMemory[ctx_index].subcomponents;",
        ]
        release = [
            "for (uint i = 0; i < %d; i++){" % t.n_slots,
            "uint index_subc = ctx->componentMemory[ctx_index].subcomponents[i];",
            "if (index_subc != 0)release_memory_component(ctx,index_subc);",
            "}",
            "}", ""]
        return create + [""] + run + body + release

    def emit(self):
        art, c = self.art, self.c
        main = c.main
        o = ["#include <stdio.h>", "#include <iostream>", "#include <assert.h>", "#include <algorithm>",
             '#include "circom.hpp"', '#include "calcwit.hpp"']
        for t in c.templates:
            o.append("void %s_create(uint soffset,uint coffset,Circom_CalcWit* ctx,std::string componentName,uint componentFather);" % t.header)
            o.append("void %s_run(uint ctx_index,Circom_CalcWit* ctx);" % t.header)
        for f in c.functions:
            o.append("void %s(Circom_CalcWit* ctx,FrElement* lvar,uint componentFather,FrElement* destination,int destination_size);" % f.header)
        table = ",\n".join("%s_run" % t.header for t in c.templates)
        o.append("Circom_TemplateFunction _functionTable[%d] = { \n%s };" % (len(c.templates), table))
        o.append("Circom_TemplateFunction _functionTableParallel[%d] = { \n%s };" % (len(c.templates), ",\n".join("NULL" for _ in c.templates)))
        n_in_map = 256
        while n_in_map < len(art.main_inputs):
            n_in_map *= 2
        o += ["uint get_main_input_signal_start() {return %d;}\n" % (main.n_out + 1),
              "uint get_main_input_signal_no() {return %d;}\n" % main.n_in,
              "uint get_total_signal_no() {return %d;}\n" % (main.n_signals + 1),
              "uint get_number_of_components() {return %d;}\n" % main.n_components,
              "uint get_size_of_input_hashmap() {return %d;}\n" % n_in_map,
              "uint get_size_of_witness() {return %d;}\n" % len(art.witness),
              "uint get_size_of_constants() {return %d;}\n" % len(c.constants),
              "uint get_size_of_io_map() {return %d;}\n" % len(c.io_map),
              "uint get_size_of_bus_field_map() {return 0;}\n"]
        o += ["void release_memory_component(Circom_CalcWit* ctx, uint pos) {", "if (pos != 0){",
              "if(ctx->componentMemory[pos].subcomponents)", "delete []ctx->componentMemory[pos].subcomponents;",
              "}", "}", ""]
        o.append("// function declarations")
        for f in c.functions:
            o += self.function(f)
        o.append("// template declarations")
        for t in c.templates:
            o += self.template(t)
        o += ["void run(Circom_CalcWit* ctx){", '%s_create(1,0,ctx,"main",0);' % main.header]
        if main.n_in > 0:
            o.append("%s_run(0,ctx);" % main.header)
        o += ["}", ""]
        return "\n".join(o)


def emit_cpp(art):
    return CppEmitter(art).emit()
