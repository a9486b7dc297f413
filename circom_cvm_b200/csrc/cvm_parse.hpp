// Parser for the Circom-Virtual-Machine text format emitted by the reference's --cvm backend.
//
// Grammar follows the emitters, not the (partly outdated) format document:
//   header directives   compiler/src/circuit_design/circuit.rs:577-621,
//                       code_producers/src/cvm_elements/cvm_code_generator.rs:1785-1888
//   %%template line     compiler/src/circuit_design/template.rs:158-208 (bracket 1 = Input wires,
//                       bracket 2 = Output wires: compiler/src/circuit_design/build.rs:87-104)
//   %%function line     compiler/src/circuit_design/function.rs:137-168
//   instructions        `impl WriteCVM` in compiler/src/intermediate_representation/*_bucket.rs,
//                       mnemonics cvm_code_generator.rs:26-283
// Extension (not in the fork, which emits nothing for component creation:
// create_component_bucket.rs:356-360):   ;;%%create_cmp <slot> $<header> <sig_off> <sig_jump> <cmp_off> <cmp_jump> <n>
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "fr.cuh"

namespace cvm {

enum Op : uint16_t {
    OP_MOV, OP_FF_ADD, OP_FF_SUB, OP_FF_MUL, OP_FF_DIV, OP_FF_IDIV, OP_FF_REM, OP_FF_POW, OP_FF_SHL, OP_FF_SHR,
    OP_FF_BAND, OP_FF_BOR, OP_FF_BXOR, OP_FF_BNOT, OP_FF_LT, OP_FF_LE, OP_FF_GT, OP_FF_GE, OP_FF_EQ, OP_FF_NEQ,
    OP_FF_AND, OP_FF_OR, OP_FF_EQZ, OP_FF_WRAP_I64,
    OP_I64_ADD, OP_I64_SUB, OP_I64_MUL, OP_I64_LT, OP_I64_LE, OP_I64_GT, OP_I64_GE, OP_I64_EQ, OP_I64_NEQ,
    OP_FF_LOAD, OP_FF_STORE, OP_GET_SIGNAL, OP_SET_SIGNAL, OP_GET_CMP_SIGNAL,
    OP_SET_CMP_INPUT, OP_SET_CMP_INPUT_CNT, OP_SET_CMP_INPUT_RUN, OP_SET_CMP_INPUT_CNT_CHECK,
    OP_LOOP, OP_IF, OP_ELSE, OP_END, OP_BREAK, OP_CONTINUE, OP_ERROR, OP_CALL, OP_RETURN, OP_CREATE_CMP,
    OP_MAPPED_UNSUPPORTED
};

enum OperandKind : uint8_t { K_REG, K_I64, K_FF, K_SPR, K_ARG_MEM, K_ARG_SIG, K_ARG_SUBSIG };

struct Operand {
    uint8_t kind = K_I64;
    int64_t val = 0;      // K_REG: register index; K_I64: literal; K_FF: index into Program::ffconst
    // call arguments of the form i64.memory(a,n) / signal(a,n) / subcmpsignal(c,a,n)
    uint8_t akind = K_I64, ckind = K_I64;
    int64_t aval = 0, cval = 0, n = 0;
};

struct Ins {
    uint16_t op;
    int32_t dst = -1;             // register index
    std::vector<Operand> args;
    int32_t target = -1;          // OP_CALL: code index; OP_CREATE_CMP: code index
    int32_t line = 0;
    // control-flow links (indices into Code::ins)
    int32_t m_else = -1, m_end = -1, m_loop = -1;
    // OP_CREATE_CMP payload
    int64_t cc[6] = {0, 0, 0, 0, 0, 0};  // slot, sig_off, sig_jump, cmp_off, cmp_jump, n
    bool scalar_return = false;
};

struct Code {
    std::string header;
    bool is_function = false;
    int64_t n_inputs = 0, n_outputs = 0, n_signals = 0, n_subcmps = 0, local_memory = 0;
    std::vector<Ins> ins;
    std::unordered_map<std::string, int> regmap;
    int nregs = 0;
    int reg_destination = -1, reg_destination_size = -1;
};

struct Program {
    std::vector<Code> codes;
    std::unordered_map<std::string, int> code_index;
    std::vector<fr::Fr> ffconst;                      // canonical values of ff.<n> literals
    std::unordered_map<std::string, int> ffconst_index;
    int64_t n_signals = 0;
    int start = -1;
    std::string start_name;
    std::vector<int64_t> witness;
};

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// decimal string -> canonical Fr (reduced mod q by the caller being < q; literals emitted by the compiler are)
inline fr::Fr fr_from_decimal(const std::string &s) {
    fr::Fr r = fr::zero();
    for (char ch : s) {
        if (ch < '0' || ch > '9') throw ParseError("bad field literal: " + s);
        uint64_t carry = (uint64_t)(ch - '0');
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)r.v[i] * 10 + carry;
            r.v[i] = (uint32_t)t;
            carry = t >> 32;
        }
        if (carry) throw ParseError("field literal does not fit 256 bits: " + s);
    }
    while (fr::geq_raw(r, fr::modulus())) {
        fr::Fr t;
        fr::sub_raw(t, r, fr::modulus());
        r = t;
    }
    return r;
}

inline std::vector<std::string> split_ws(const std::string &line) {
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r')) i++;
        size_t j = i;
        while (j < n && !(line[j] == ' ' || line[j] == '\t' || line[j] == '\r')) j++;
        if (j > i) out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

inline bool is_int_token(const std::string &t) {
    if (t.empty()) return false;
    size_t i = (t[0] == '-') ? 1 : 0;
    if (i >= t.size()) return false;
    for (; i < t.size(); i++)
        if (t[i] < '0' || t[i] > '9') return false;
    return true;
}

class Parser {
  public:
    Program prog;

    void parse_file(const std::string &path) {
        std::ifstream f(path);
        if (!f) throw ParseError("cannot open " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        parse_text(ss.str());
    }

    void parse_text(const std::string &text) {
        size_t pos = 0;
        int lineno = 0;
        std::vector<std::pair<int, std::string>> pending_calls;  // resolved after all units are known
        while (pos <= text.size()) {
            size_t e = text.find('\n', pos);
            if (e == std::string::npos) e = text.size();
            std::string line = text.substr(pos, e - pos);
            pos = e + 1;
            lineno++;
            size_t b = line.find_first_not_of(" \t\r");
            if (b == std::string::npos) continue;
            line = line.substr(b);
            if (line.rfind(";;%%create_cmp", 0) == 0) {
                parse_create(line, lineno);
                continue;
            }
            if (line.rfind(";;", 0) == 0 || line.rfind("//", 0) == 0) continue;
            if (line.rfind("%%", 0) == 0) {
                parse_directive(line, lineno);
                continue;
            }
            parse_instruction(line, lineno);
        }
        finish();
    }

  private:
    Code *cur = nullptr;
    std::vector<std::pair<std::pair<int, int>, std::string>> fixups;  // ((code, ins), name)

    static int64_t dims_size(const std::string &payload) {
        auto t = split_ws(payload);
        int64_t total = 0;
        size_t k = 0;
        while (k + 1 < t.size()) {
            int nd = atoi(t[k + 1].c_str());
            int64_t size = 1;
            for (int d = 0; d < nd; d++) size *= atoll(t[k + 2 + d].c_str());
            total += size;
            k += 2 + nd;
        }
        return total;
    }

    static std::vector<std::string> brackets(const std::string &line) {
        std::vector<std::string> out;
        size_t p = 0;
        while ((p = line.find('[', p)) != std::string::npos) {
            size_t q = line.find(']', p);
            if (q == std::string::npos) break;
            out.push_back(line.substr(p + 1, q - p - 1));
            p = q + 1;
        }
        return out;
    }

    void new_code(const std::string &header, bool is_function) {
        if (prog.code_index.count(header)) throw ParseError("duplicate unit " + header);
        prog.code_index[header] = (int)prog.codes.size();
        prog.codes.emplace_back();
        cur = &prog.codes.back();
        cur->header = header;
        cur->is_function = is_function;
        cur->reg_destination = reg("destination");
        cur->reg_destination_size = reg("destination_size");
    }

    void parse_directive(const std::string &line, int lineno) {
        auto t = split_ws(line);
        const std::string &d = t[0];
        if (d == "%%prime") {
            fr::Fr p = fr::zero();
            // must be BN254: compare decimal text
            if (t.size() < 2 ||
                t[1] != "21888242871839275222246405745257275088548364400416034343698204186575808495617")
                throw ParseError("unsupported prime (only bn128 is implemented)");
            (void)p;
        } else if (d == "%%signals") {
            prog.n_signals = atoll(t.at(1).c_str());
        } else if (d == "%%start") {
            prog.start_name = t.at(1);
        } else if (d == "%%witness") {
            for (size_t i = 1; i < t.size(); i++) prog.witness.push_back(atoll(t[i].c_str()));
        } else if (d == "%%template") {
            new_code(t.at(1), false);
            auto br = brackets(line);
            if (br.size() < 4) throw ParseError("malformed %%template at line " + std::to_string(lineno));
            cur->n_inputs = dims_size(br[0]);
            cur->n_outputs = dims_size(br[1]);
            cur->n_signals = atoll(br[2].c_str());
            cur->n_subcmps = atoll(br[3].c_str());
        } else if (d == "%%function") {
            new_code(t.at(1), true);
        } else if (d == "%%components_heap" || d == "%%components" || d == "%%type") {
            // layout hints only
        } else {
            throw ParseError("unknown directive " + d + " at line " + std::to_string(lineno));
        }
    }

    int reg(const std::string &name) {
        auto it = cur->regmap.find(name);
        if (it != cur->regmap.end()) return it->second;
        int id = cur->nregs++;
        cur->regmap[name] = id;
        return id;
    }

    int ffconst(const std::string &dec) {
        auto it = prog.ffconst_index.find(dec);
        if (it != prog.ffconst_index.end()) return it->second;
        int id = (int)prog.ffconst.size();
        prog.ffconst.push_back(fr_from_decimal(dec));
        prog.ffconst_index[dec] = id;
        return id;
    }

    void simple_operand(const std::string &tok, uint8_t &kind, int64_t &val) {
        if (tok.rfind("i64.", 0) == 0 && is_int_token(tok.substr(4))) {
            kind = K_I64;
            val = atoll(tok.c_str() + 4);
        } else if (tok.rfind("ff.", 0) == 0 && is_int_token(tok.substr(3))) {
            kind = K_FF;
            val = ffconst(tok.substr(3));
        } else if (tok.rfind("i64", 0) == 0 && is_int_token(tok.substr(3))) {
            kind = K_I64;  // emitter defect: "i64<n>" without the dot (cvm_code_generator.rs:275,278)
            val = atoll(tok.c_str() + 3);
        } else if (is_int_token(tok)) {
            kind = K_I64;  // bare integers: `ff.sub 0 x`, `x = 3` (compute_bucket.rs:606, return_bucket.rs:138)
            val = atoll(tok.c_str());
        } else if (tok == "spr") {
            kind = K_SPR;
            val = 0;
        } else {
            kind = K_REG;
            val = reg(tok);
        }
    }

    Operand operand(const std::string &tok) {
        Operand o;
        static const char *forms[3] = {"i64.memory(", "signal(", "subcmpsignal("};
        for (int f = 0; f < 3; f++) {
            size_t L = strlen(forms[f]);
            if (tok.rfind(forms[f], 0) == 0 && tok.back() == ')') {
                std::string inner = tok.substr(L, tok.size() - L - 1);
                std::vector<std::string> parts;
                size_t p = 0;
                while (true) {
                    size_t c = inner.find(',', p);
                    parts.push_back(inner.substr(p, c == std::string::npos ? std::string::npos : c - p));
                    if (c == std::string::npos) break;
                    p = c + 1;
                }
                o.kind = (f == 0) ? K_ARG_MEM : (f == 1 ? K_ARG_SIG : K_ARG_SUBSIG);
                if (f == 2) {
                    if (parts.size() != 3) throw ParseError("bad call argument " + tok);
                    simple_operand(parts[0], o.ckind, o.cval);
                    simple_operand(parts[1], o.akind, o.aval);
                    o.n = atoll(parts[2].c_str());
                } else {
                    if (parts.size() != 2) throw ParseError("bad call argument " + tok);
                    simple_operand(parts[0], o.akind, o.aval);
                    o.n = atoll(parts[1].c_str());
                }
                return o;
            }
        }
        simple_operand(tok, o.kind, o.val);
        return o;
    }

    void parse_create(const std::string &line, int lineno) {
        if (!cur) throw ParseError("create_cmp outside a template");
        auto t = split_ws(line);
        if (t.size() < 8) throw ParseError("malformed create_cmp at line " + std::to_string(lineno));
        Ins in;
        in.op = OP_CREATE_CMP;
        in.line = lineno;
        in.cc[0] = atoll(t[1].c_str());
        for (int k = 0; k < 5; k++) in.cc[1 + k] = atoll(t[3 + k].c_str());
        std::string name = t[2];
        if (!name.empty() && name[0] == '$') name = name.substr(1);
        fixups.push_back({{(int)(cur - &prog.codes[0]), (int)cur->ins.size()}, name});
        cur->ins.push_back(in);
    }

    static const std::unordered_map<std::string, uint16_t> &optable() {
        static const std::unordered_map<std::string, uint16_t> t = {
            {"ff.add", OP_FF_ADD}, {"ff.sub", OP_FF_SUB}, {"ff.mul", OP_FF_MUL}, {"ff.div", OP_FF_DIV},
            {"ff.idiv", OP_FF_IDIV}, {"ff.rem", OP_FF_REM}, {"ff.pow", OP_FF_POW}, {"ff.shl", OP_FF_SHL},
            {"ff.shr", OP_FF_SHR}, {"ff.band", OP_FF_BAND}, {"ff.bor", OP_FF_BOR}, {"ff.bxor", OP_FF_BXOR},
            {"ff.bnot", OP_FF_BNOT}, {"ff.lt", OP_FF_LT}, {"ff.le", OP_FF_LE}, {"ff.gt", OP_FF_GT},
            {"ff.ge", OP_FF_GE}, {"ff.eq", OP_FF_EQ}, {"ff.neq", OP_FF_NEQ}, {"ff.and", OP_FF_AND},
            {"ff.or", OP_FF_OR}, {"ff.eqz", OP_FF_EQZ}, {"ff.wrap_i64", OP_FF_WRAP_I64},
            {"i64.add", OP_I64_ADD}, {"i64.sub", OP_I64_SUB}, {"i64.mul", OP_I64_MUL}, {"i64.lt", OP_I64_LT},
            {"i64.le", OP_I64_LE}, {"i64.gt", OP_I64_GT}, {"i64.ge", OP_I64_GE}, {"i64.eq", OP_I64_EQ},
            {"i64.neq", OP_I64_NEQ}, {"ff.load", OP_FF_LOAD}, {"get_signal", OP_GET_SIGNAL},
            {"get_cmp_signal", OP_GET_CMP_SIGNAL},
            {"get_template_id", OP_MAPPED_UNSUPPORTED}, {"get_template_signal_position", OP_MAPPED_UNSUPPORTED},
            {"get_template_signal_size", OP_MAPPED_UNSUPPORTED},
            {"get_template_signal_dimension", OP_MAPPED_UNSUPPORTED},
            {"get_template_signal_type", OP_MAPPED_UNSUPPORTED}, {"get_bus_signal_position", OP_MAPPED_UNSUPPORTED},
            {"get_bus_signal_size", OP_MAPPED_UNSUPPORTED}, {"get_bus_signal_dimension", OP_MAPPED_UNSUPPORTED},
            {"get_bus_signal_type", OP_MAPPED_UNSUPPORTED},
        };
        return t;
    }
    static const std::unordered_map<std::string, uint16_t> &stmttable() {
        static const std::unordered_map<std::string, uint16_t> t = {
            {"ff.store", OP_FF_STORE}, {"set_signal", OP_SET_SIGNAL}, {"set_cmp_input", OP_SET_CMP_INPUT},
            {"set_cmp_input_cnt", OP_SET_CMP_INPUT_CNT}, {"set_cmp_input_run", OP_SET_CMP_INPUT_RUN},
            {"set_cmp_input_cnt_check", OP_SET_CMP_INPUT_CNT_CHECK}, {"loop", OP_LOOP}, {"if", OP_IF},
            {"else", OP_ELSE}, {"end", OP_END}, {"break", OP_BREAK}, {"continue", OP_CONTINUE},
            {"error", OP_ERROR}, {"ff.call", OP_CALL}, {"return", OP_RETURN},
        };
        return t;
    }

    void parse_instruction(const std::string &line, int lineno) {
        auto t = split_ws(line);
        if (t.empty()) return;
        if (!cur) throw ParseError("instruction outside a unit at line " + std::to_string(lineno));
        if (t[0] == "local.memory") {
            cur->local_memory = atoll(t.at(1).c_str());
            return;
        }
        Ins in;
        in.line = lineno;
        if (t.size() >= 3 && t[1] == "=") {
            const std::string &d = t[0];
            if (d.rfind("i64.", 0) == 0 || d.rfind("ff.", 0) == 0)
                throw ParseError("line " + std::to_string(lineno) + ": assignment to the literal operand '" + d +
                                 "' (reference emitter defect, store_bucket.rs:1026-1028); re-emit with "
                                 "addresses held in registers");
            in.dst = reg(d);
            if (t.size() == 3) {
                in.op = OP_MOV;
                in.args.push_back(operand(t[2]));
            } else {
                auto it = optable().find(t[2]);
                if (it == optable().end())
                    throw ParseError("line " + std::to_string(lineno) + ": unknown operation " + t[2]);
                in.op = it->second;
                for (size_t k = 3; k < t.size(); k++) in.args.push_back(operand(t[k]));
            }
        } else {
            auto it = stmttable().find(t[0]);
            if (it == stmttable().end())
                throw ParseError("line " + std::to_string(lineno) + ": unknown instruction " + t[0]);
            in.op = it->second;
            size_t first = 1;
            if (in.op == OP_CALL) {
                std::string name = t.at(1);
                if (!name.empty() && name[0] == '$') name = name.substr(1);
                fixups.push_back({{(int)(cur - &prog.codes[0]), (int)cur->ins.size()}, name});
                first = 2;
            }
            if (in.op == OP_RETURN) in.scalar_return = (t.size() >= 3 && t[2] == "1");
            for (size_t k = first; k < t.size(); k++) in.args.push_back(operand(t[k]));
        }
        cur->ins.push_back(in);
    }

    void finish() {
        for (auto &fx : fixups) {
            auto it = prog.code_index.find(fx.second);
            if (it == prog.code_index.end()) throw ParseError("unknown template/function " + fx.second);
            prog.codes[fx.first.first].ins[fx.first.second].target = it->second;
        }
        auto it = prog.code_index.find(prog.start_name);
        if (it == prog.code_index.end()) throw ParseError("missing or unknown %%start");
        prog.start = it->second;
        for (auto &c : prog.codes) link(c);
    }

    static void link(Code &c) {
        struct Fr_ { int kind, start, els; };
        std::vector<Fr_> st;
        for (int pc = 0; pc < (int)c.ins.size(); pc++) {
            Ins &in = c.ins[pc];
            switch (in.op) {
                case OP_IF: st.push_back({0, pc, -1}); break;
                case OP_LOOP: st.push_back({1, pc, -1}); break;
                case OP_ELSE:
                    if (st.empty() || st.back().kind != 0) throw ParseError("else without if in " + c.header);
                    st.back().els = pc;
                    break;
                case OP_END: {
                    if (st.empty()) throw ParseError("unbalanced end in " + c.header);
                    Fr_ f = st.back();
                    st.pop_back();
                    c.ins[f.start].m_end = pc;
                    c.ins[f.start].m_else = f.els;
                    if (f.els >= 0) c.ins[f.els].m_end = pc;
                    in.m_loop = (f.kind == 1) ? f.start : -1;
                    break;
                }
                case OP_BREAK:
                case OP_CONTINUE: {
                    int k = (int)st.size() - 1;
                    while (k >= 0 && st[k].kind != 1) k--;
                    if (k < 0) throw ParseError("break/continue outside a loop in " + c.header);
                    in.m_loop = st[k].start;
                    break;
                }
                default: break;
            }
        }
        if (!st.empty()) throw ParseError("unbalanced control flow in " + c.header);
    }
};

}  // namespace cvm
