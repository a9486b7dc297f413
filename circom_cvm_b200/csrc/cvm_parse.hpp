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
// Component creation: the fork emits nothing for it (create_component_bucket.rs:356-360).  Two sources are accepted:
//   * the extension line  ;;%%create_cmp <slot> $<header> <sig_off> <sig_jump> <cmp_off> <cmp_jump> <n>
//     (a comment to any other consumer; patches/create_component_bucket.rs.diff makes the emitter print it), or
//   * the generated <circuit>.cpp of the same compile: recover_creates() reads the `<Sub>_create(...)` blocks that
//     `impl WriteC for CreateCmpBucket` prints (create_component_bucket.rs:206-354) and places them at the top of the
//     template's body (creation does not depend on signal values).
//
// Emitter defects that are honoured rather than rejected:
//   * copy loops increment their address operands textually, so a literal address is assigned to
//     (`i64.5 = i64.add i64.5 i64.1`, store_bucket.rs:1026-1028, call_bucket.rs:985-987): inside the loop that does this
//     the token is a register initialised to its literal value at loop entry (fix_literal_registers);
//   * a multi-element `return` passes the VALUE of the first element (return_bucket.rs:131 evaluates the load): the
//     address of that load is used (link_returns);
//   * array equality (compute_bucket.rs:538-586) compares the first elements and then increments those VALUES as if they
//     were addresses, and leaves its result in a register nobody reads while the consumer reads one that is never
//     written: the whole shape is recognised and replaced by OP_ARRAY_EQ, which computes what the C++ twin does
//     (compute_bucket.rs:375-407: the conjunction of the element-wise Fr_eq) into the register the consumer reads.
#pragma once
#include <algorithm>
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
    OP_MAPPED_UNSUPPORTED,
    // "mapped" accesses (location_rule.rs:86-171): a signal of a component in a MIXED array is addressed through the io-map
    OP_GET_TEMPLATE_ID,      // dst = template-instance id of sub-component args[0]
    OP_GET_TMPL_SIG_POS,     // dst = io_map[args[0]].defs[args[1]].offset
    OP_GET_TMPL_SIG_DIM,     // dst = io_map[args[0]].defs[args[1]].lengths[args[2]]   (args[2] >= 1)
    OP_GET_TMPL_SIG_SIZE,    // dst = io_map[args[0]].defs[args[1]].size
    OP_ARRAY_EQ      // dst = AND_k (A[k] == B[k]); args = {addr A, cmp A, addr B, cmp B}; cc = {load op A, load op B, n}
};

enum OperandKind : uint8_t { K_REG, K_I64, K_FF, K_SPR, K_ARG_MEM, K_ARG_SIG, K_ARG_SUBSIG };

struct Operand {
    uint8_t kind = K_I64;
    int64_t val = 0;      // K_REG: register index; K_I64: literal; K_FF: index into Program::ffconst
    // call arguments of the form i64.memory(a,n) / signal(a,n) / subcmpsignal(c,a,n)
    uint8_t akind = K_I64, ckind = K_I64;
    int64_t aval = 0, cval = 0, n = 0;
    // literal tokens (ids into Parser::lit_tokens) behind val / aval / cval: a literal that a copy loop assigns to
    // becomes a register inside that loop
    int32_t tok = -1, atok = -1, ctok = -1;
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
    std::vector<int64_t> positions;   // OP_CREATE_CMP of an array with undefined positions: the defined ones (else empty)
    bool scalar_return = false;
    int32_t lit_dst = -1;         // the destination is a literal token (id): resolved by fix_literal_registers
    bool ret_from_load = false;   // OP_RETURN of several elements whose operand is the VALUE loaded from args[2] (an address)
};

struct Code {
    std::string header;
    bool is_function = false;
    int64_t n_inputs = 0, n_outputs = 0, n_signals = 0, n_subcmps = 0, local_memory = 0;
    int64_t template_id = -1;     // <name>_<id> (template headers: build.rs:50, executed_template.rs); what _create stores
    std::vector<Ins> ins;
    std::unordered_map<std::string, int> regmap;
    int nregs = 0;
    int reg_destination = -1, reg_destination_size = -1;
};

// one input/output signal of a template instance that sits in a mixed component array (IODef, build.rs:531-552; the
// .dat record c_code_generator.rs:617-674 and its reader main.cpp:59-92 drop lengths[0], which no address needs)
struct IoDef {
    int64_t offset = 0;
    std::vector<int64_t> lengths_tail;   // lengths[1..]
    int64_t size = 1;
    int64_t bus_id = 0;
};

struct MainInput {
    std::string name;        // as loadJson addresses it (qualified for buses), without indices
    int64_t start = 0, size = 0;   // first signal, number of elements
};

struct Program {
    std::vector<MainInput> main_inputs;   // ;;%%main_input lines (the name table the .dat carries, c_code_generator.rs:511-539)
    std::unordered_map<int64_t, std::vector<IoDef>> io_map;   // template-instance id -> defs indexed by signal code
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
            if (line.rfind(";;%%io_map", 0) == 0) {
                parse_io_map(line, lineno);
                continue;
            }
            if (line.rfind(";;%%main_input", 0) == 0) {   // ;;%%main_input <name> <first signal> <size>
                auto t = split_ws(line);
                if (t.size() != 4) throw ParseError("malformed main_input at line " + std::to_string(lineno));
                MainInput mi;
                mi.name = t[1];
                mi.start = atoll(t[2].c_str());
                mi.size = atoll(t[3].c_str());
                prog.main_inputs.push_back(mi);
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
            size_t us = cur->header.rfind('_');
            if (us != std::string::npos && is_int_token(cur->header.substr(us + 1)) && cur->header[us + 1] != '-')
                cur->template_id = atoll(cur->header.c_str() + us + 1);
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

    std::vector<std::string> lit_tokens;
    std::unordered_map<std::string, int> lit_index;
    int lit_id(const std::string &tok) {
        auto it = lit_index.find(tok);
        if (it != lit_index.end()) return it->second;
        lit_tokens.push_back(tok);
        lit_index[tok] = (int)lit_tokens.size() - 1;
        return (int)lit_tokens.size() - 1;
    }
    static bool is_literal_token(const std::string &tok) {
        return (tok.rfind("i64.", 0) == 0 && is_int_token(tok.substr(4))) || (tok.rfind("ff.", 0) == 0 && is_int_token(tok.substr(3)));
    }

    void simple_operand(const std::string &tok, uint8_t &kind, int64_t &val, int32_t *tokid = nullptr) {
        if (tokid && is_literal_token(tok)) *tokid = lit_id(tok);
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
                    simple_operand(parts[0], o.ckind, o.cval, &o.ctok);
                    simple_operand(parts[1], o.akind, o.aval, &o.atok);
                    o.n = atoll(parts[2].c_str());
                } else {
                    if (parts.size() != 2) throw ParseError("bad call argument " + tok);
                    simple_operand(parts[0], o.akind, o.aval, &o.atok);
                    o.n = atoll(parts[1].c_str());
                }
                return o;
            }
        }
        simple_operand(tok, o.kind, o.val, &o.tok);
        return o;
    }

    // ;;%%io_map <template id> <n defs> { <offset> <n> <n lengths (dimensions 1..)> <size> <bus id> }*
    // -- the record of the .dat io-map (c_code_generator.rs:617-674) as text; a comment to any other consumer
    void parse_io_map(const std::string &line, int lineno) {
        auto t = split_ws(line);
        if (t.size() < 3) throw ParseError("malformed io_map at line " + std::to_string(lineno));
        const int64_t tid = atoll(t[1].c_str()), n = atoll(t[2].c_str());
        std::vector<IoDef> defs;
        size_t k = 3;
        for (int64_t d = 0; d < n; d++) {
            if (k + 2 > t.size()) throw ParseError("malformed io_map at line " + std::to_string(lineno));
            IoDef def;
            def.offset = atoll(t[k].c_str());
            const int64_t len = atoll(t[k + 1].c_str());
            k += 2;
            if (len < 0 || k + (size_t)len + 2 > t.size()) throw ParseError("malformed io_map at line " + std::to_string(lineno));
            for (int64_t i = 0; i < len; i++) def.lengths_tail.push_back(atoll(t[k++].c_str()));
            def.size = atoll(t[k++].c_str());
            def.bus_id = atoll(t[k++].c_str());
            defs.push_back(def);
        }
        prog.io_map[tid] = std::move(defs);
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
        for (size_t k = 8; k < t.size(); k++)
            if (t[k] != "|") in.positions.push_back(atoll(t[k].c_str()));
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
            {"get_template_id", OP_GET_TEMPLATE_ID}, {"get_template_signal_position", OP_GET_TMPL_SIG_POS},
            {"get_template_signal_size", OP_GET_TMPL_SIG_SIZE},
            {"get_template_signal_dimension", OP_GET_TMPL_SIG_DIM},
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
            if (is_literal_token(d)) in.lit_dst = lit_id(d);   // emitter defect, store_bucket.rs:1026-1028: see fix_literal_registers
            else in.dst = reg(d);
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
        for (auto &c : prog.codes) {
            cur = &c;
            fix_array_eq(c);
            fix_literal_registers(c);
            link_returns(c);
            link(c);
        }
    }

    // index of the innermost `loop` whose body holds instruction pc (-1: none), and of its `end`
    static void loop_ranges(const Code &c, std::vector<int> &loop_of, std::vector<int> &end_of) {
        loop_of.assign(c.ins.size(), -1);
        end_of.assign(c.ins.size(), -1);
        std::vector<std::pair<int, int>> st;   // (kind, start)
        for (int pc = 0; pc < (int)c.ins.size(); pc++) {
            int inner = -1;
            for (int k = (int)st.size() - 1; k >= 0; k--)
                if (st[k].first == 1) { inner = st[k].second; break; }
            loop_of[pc] = inner;
            const uint16_t op = c.ins[pc].op;
            if (op == OP_IF) st.push_back({0, pc});
            else if (op == OP_LOOP) st.push_back({1, pc});
            else if (op == OP_END) {
                if (st.empty()) throw ParseError("unbalanced end in " + c.header);
                if (st.back().first == 1) end_of[st.back().second] = pc;
                st.pop_back();
            }
        }
    }

    // `i64.5 = i64.add i64.5 i64.1` in a copy loop (store_bucket.rs:1016-1035, call_bucket.rs:975-994).  The loop's shape is
    //     loop / if cnt / GET src_location / SET dest_location / cnt = i64.sub cnt i64.1 /
    //     src_location = i64.add src_location i64.1 / dest_location = i64.add dest_location i64.1 / continue / end / break / end
    //     [GET src_location / set_cmp_input_{run,cnt_check} c dest_location v]        <- the peeled last element
    // and the two increments mean "the address operand of GET" and "the address operand of SET" -- by position, not by
    // token (both may be the same literal, e.g. variable 0 copied to signal 0).  A literal address becomes a fresh register
    // that is set to the literal right before `loop`, is used by GET / SET / the peeled pair, and nowhere else.
    void fix_literal_registers(Code &c) {
        bool any = false;
        for (const Ins &in : c.ins) any = any || in.lit_dst >= 0;
        if (!any) return;
        auto get_addr_arg = [](uint16_t op) -> int {
            return (op == OP_FF_LOAD || op == OP_GET_SIGNAL) ? 0 : op == OP_GET_CMP_SIGNAL ? 1 : -1;
        };
        auto set_addr_arg = [](uint16_t op) -> int {
            if (op == OP_FF_STORE || op == OP_SET_SIGNAL) return 0;
            if (op == OP_SET_CMP_INPUT || op == OP_SET_CMP_INPUT_CNT || op == OP_SET_CMP_INPUT_RUN || op == OP_SET_CMP_INPUT_CNT_CHECK) return 1;
            return -1;
        };
        struct Init { int loop, reg, tok; };
        std::vector<Init> inits;
        std::vector<Ins> &I = c.ins;
        for (int L = 0; L + 10 < (int)I.size(); L++) {
            if (I[L].op != OP_LOOP || I[L + 1].op != OP_IF) continue;
            const int ga = get_addr_arg(I[L + 2].op), sa = set_addr_arg(I[L + 3].op);
            if (ga < 0 || sa < 0 || I[L + 4].op != OP_I64_SUB || I[L + 5].op != OP_I64_ADD || I[L + 6].op != OP_I64_ADD) continue;
            if (I[L + 7].op != OP_CONTINUE || I[L + 8].op != OP_END || I[L + 9].op != OP_BREAK || I[L + 10].op != OP_END) continue;
            const int end = L + 10;
            const bool peeled = end + 2 < (int)I.size() && I[end + 1].op == I[L + 2].op &&
                                (I[end + 2].op == OP_SET_CMP_INPUT_RUN || I[end + 2].op == OP_SET_CMP_INPUT_CNT_CHECK);
            for (int k = 0; k < 2; k++) {
                Ins &inc = I[L + 5 + k];
                if (inc.lit_dst < 0) continue;
                Ins &user = I[L + 2 + k];
                const int ai = k ? sa : ga;
                if ((int)user.args.size() <= ai || user.args[ai].tok != inc.lit_dst || inc.args.at(0).tok != inc.lit_dst)
                    throw ParseError("line " + std::to_string(inc.line) + ": copy loop increments the literal '" + lit_tokens[inc.lit_dst] +
                                     "' that is not the address of its " + (k ? "store" : "load"));
                const int r = reg(lit_tokens[inc.lit_dst] + (k ? "@dst" : "@src") + std::to_string(inc.line));
                auto to_reg = [&](Operand &o) { o.kind = K_REG; o.val = r; o.tok = -1; };
                to_reg(user.args[ai]);
                to_reg(inc.args[0]);
                inc.dst = r;
                if (peeled) {
                    Ins &p = I[end + 1 + k];
                    const int pi = k ? 1 : ga;
                    if ((int)p.args.size() > pi && p.args[pi].tok == inc.lit_dst) to_reg(p.args[pi]);
                }
                inits.push_back({L, r, inc.lit_dst});
                inc.lit_dst = -1;
            }
        }
        for (const Ins &in : I)
            if (in.lit_dst >= 0)
                throw ParseError("line " + std::to_string(in.line) + ": assignment to the literal operand '" + lit_tokens[in.lit_dst] +
                                 "' outside a loop of the copy-loop shape (store_bucket.rs:1016-1035)");
        // initialisations, inserted from the back so that earlier loop indices stay valid
        std::stable_sort(inits.begin(), inits.end(), [](const Init &x, const Init &y) { return x.loop > y.loop; });
        for (const Init &p : inits) {
            Ins mv;
            mv.op = OP_MOV;
            mv.dst = p.reg;
            mv.line = I[p.loop].line;
            Operand o;
            uint8_t kind;
            int64_t val;
            simple_operand(lit_tokens[p.tok], kind, val);
            o.kind = kind;
            o.val = val;
            mv.args.push_back(o);
            I.insert(I.begin() + p.loop, mv);
        }
    }

    // nearest instruction before pc that assigns register r (-1: none)
    static int def_before(const Code &c, int pc, int r) {
        for (int k = pc - 1; k >= 0; k--)
            if (c.ins[k].dst == r) return k;
        return -1;
    }

    // multi-element return whose operand is the loaded first element (return_bucket.rs:131): keep the load's address
    static void link_returns(Code &c) {
        for (int pc = 0; pc < (int)c.ins.size(); pc++) {
            Ins &in = c.ins[pc];
            if (in.op != OP_RETURN || in.scalar_return || in.args.empty() || in.args[0].kind != K_REG) continue;
            const int d = def_before(c, pc, (int)in.args[0].val);
            if (d >= 0 && c.ins[d].op == OP_FF_LOAD) {
                in.ret_from_load = true;
                in.args.resize(2);
                in.args.push_back(c.ins[d].args.at(0));
            }
        }
    }

    // the emitter's array-equality shape (compute_bucket.rs:538-586) -> OP_ARRAY_EQ
    void fix_array_eq(Code &c) {
        auto is_reg = [](const Operand &o, int r) { return o.kind == K_REG && o.val == r; };
        auto is_lit1 = [](const Operand &o) { return o.kind == K_I64 && o.val == 1; };
        for (int i = 0; i + 12 < (int)c.ins.size(); i++) {
            std::vector<Ins> &I = c.ins;
            if (I[i].op != OP_MOV || I[i].args.size() != 1 || I[i].args[0].kind != K_I64) continue;
            const int cnt = I[i].dst;
            if (I[i + 1].op != OP_LOOP || I[i + 2].op != OP_IF || !is_reg(I[i + 2].args.at(0), cnt)) continue;
            if (I[i + 3].op != OP_FF_EQ || I[i + 3].args.size() != 2 || I[i + 3].args[0].kind != K_REG || I[i + 3].args[1].kind != K_REG) continue;
            const int ra = (int)I[i + 3].args[0].val, rb = (int)I[i + 3].args[1].val, r2 = I[i + 3].dst;
            if (I[i + 4].op != OP_IF || !is_reg(I[i + 4].args.at(0), r2)) continue;
            if (I[i + 5].op != OP_I64_SUB || I[i + 5].dst != cnt || !is_reg(I[i + 5].args.at(0), cnt) || !is_lit1(I[i + 5].args.at(1))) continue;
            if (I[i + 6].op != OP_I64_ADD || I[i + 6].dst != ra || !is_reg(I[i + 6].args.at(0), ra) || !is_lit1(I[i + 6].args.at(1))) continue;
            if (I[i + 7].op != OP_I64_ADD || I[i + 7].dst != rb || !is_reg(I[i + 7].args.at(0), rb) || !is_lit1(I[i + 7].args.at(1))) continue;
            if (I[i + 8].op != OP_CONTINUE || I[i + 9].op != OP_END || I[i + 10].op != OP_END || I[i + 11].op != OP_BREAK || I[i + 12].op != OP_END) continue;
            const int da = def_before(c, i, ra), db = def_before(c, i, rb);
            auto is_load = [](uint16_t op) { return op == OP_FF_LOAD || op == OP_GET_SIGNAL || op == OP_GET_CMP_SIGNAL; };
            if (da < 0 || db < 0 || !is_load(I[da].op) || !is_load(I[db].op))
                throw ParseError("line " + std::to_string(I[i].line) + ": array-equality loop (compute_bucket.rs:538-586) whose operands are not loads");
            // the consumer reads the register allocated two before the one the loop writes (fresh_var order: result, counter,
            // inner result; cvm_elements/mod.rs:202-206)
            std::string n2;
            for (auto &kv : c.regmap)
                if (kv.second == r2) n2 = kv.first;
            if (n2.rfind("x_", 0) != 0 || !is_int_token(n2.substr(2)) || atoll(n2.c_str() + 2) < 2)
                throw ParseError("line " + std::to_string(I[i].line) + ": array-equality loop with an unexpected result register " + n2);
            Ins eq;
            eq.op = OP_ARRAY_EQ;
            eq.line = I[i].line;
            eq.dst = reg("x_" + std::to_string(atoll(n2.c_str() + 2) - 2));
            auto addr_of = [&](int d, Operand &addr, Operand &cmp) {
                if (I[d].op == OP_GET_CMP_SIGNAL) { cmp = I[d].args.at(0); addr = I[d].args.at(1); }
                else addr = I[d].args.at(0);
            };
            Operand aa, ca, ab, cb;
            addr_of(da, aa, ca);
            addr_of(db, ab, cb);
            eq.args = {aa, ca, ab, cb};
            eq.cc[0] = I[da].op;
            eq.cc[1] = I[db].op;
            eq.cc[2] = I[i].args[0].val;
            I.erase(I.begin() + i, I.begin() + i + 13);
            I.insert(I.begin() + i, eq);
        }
    }

  public:
    // The io-map of a <circuit>.dat (only present when the circuit has mixed component arrays): the file carries no section
    // sizes, the generated C++ does (`uint get_size_of_*() {return N;}`, circuit.rs:481-497), so both are needed.  Layout:
    // hash map (24 B entries), witness list (u64), constants (40 B), then M template-instance ids (u32) and per id
    // { u32 n ; n x { u32 offset, u32 len, len x u32 lengths[1..], u32 size, u32 busId } }   (c_code_generator.rs:617-674,
    // reader main.cpp:59-92).  Returns the number of template instances read.
    size_t read_dat_io_map(const std::string &cpp, const unsigned char *dat, size_t dat_len) {
        auto getter = [&](const char *name, int64_t &out) -> bool {
            std::string key = std::string("uint ") + name + "() {return ";
            size_t p = cpp.find(key);
            if (p == std::string::npos) return false;
            out = atoll(cpp.c_str() + p + key.size());
            return true;
        };
        int64_t n_hash = 0, n_wit = 0, n_const = 0, n_io = 0;
        if (!getter("get_size_of_io_map", n_io) || n_io == 0) return 0;
        if (!getter("get_size_of_input_hashmap", n_hash) || !getter("get_size_of_witness", n_wit) ||
            !getter("get_size_of_constants", n_const))
            throw ParseError("the generated C++ does not define the get_size_of_* functions the .dat layout depends on");
        size_t pos = (size_t)n_hash * 24 + (size_t)n_wit * 8 + (size_t)n_const * 40;
        auto u32_at = [&](size_t at) -> uint32_t {
            if (at + 4 > dat_len) throw ParseError("the .dat file ends inside its io-map");
            return (uint32_t)dat[at] | (uint32_t)dat[at + 1] << 8 | (uint32_t)dat[at + 2] << 16 | (uint32_t)dat[at + 3] << 24;
        };
        std::vector<uint32_t> ids;
        for (int64_t i = 0; i < n_io; i++, pos += 4) ids.push_back(u32_at(pos));
        for (uint32_t id : ids) {
            const uint32_t n = u32_at(pos);
            pos += 4;
            std::vector<IoDef> defs;
            for (uint32_t j = 0; j < n; j++) {
                IoDef d;
                d.offset = u32_at(pos);
                const uint32_t len = u32_at(pos + 4);
                pos += 8;
                for (uint32_t k = 0; k < len; k++, pos += 4) d.lengths_tail.push_back(u32_at(pos));
                d.size = u32_at(pos);
                d.bus_id = u32_at(pos + 4);
                pos += 8;
                defs.push_back(std::move(d));
            }
            prog.io_map[id] = std::move(defs);
        }
        return ids.size();
    }

    // Component creation recovered from the generated <circuit>.cpp (see the header comment).  Call after parse_text /
    // before the program is used: it (re)links the units it touches.
    void recover_creates(const std::string &cpp) {
        size_t pos = 0;
        Code *unit = nullptr;
        std::vector<Ins> found;
        auto flush = [&]() {
            if (unit && !found.empty()) {
                bool has = false;
                for (const Ins &in : unit->ins) has = has || in.op == OP_CREATE_CMP;
                if (!has) {
                    unit->ins.insert(unit->ins.begin(), found.begin(), found.end());
                    link(*unit);
                }
            }
            found.clear();
        };
        auto num_after = [](const std::string &line, const std::string &key, int64_t &out) -> bool {
            size_t p = line.find(key);
            if (p == std::string::npos) return false;
            out = atoll(line.c_str() + p + key.size());
            return true;
        };
        int64_t multi_slot = -1, multi_cmp = 0, multi_sig = 0, multi_n = 0, sig_jump = 0, cmp_jump = 0;
        Code *create_unit = nullptr;
        std::string multi_sym;
        std::vector<int64_t> multi_positions;
        bool in_multi = false;
        while (pos < cpp.size()) {
            size_t e = cpp.find('\n', pos);
            if (e == std::string::npos) e = cpp.size();
            std::string line = cpp.substr(pos, e - pos);
            pos = e + 1;
            size_t b = line.find_first_not_of(" \t\r");
            if (b == std::string::npos) continue;
            line = line.substr(b);
            if (line.rfind("void ", 0) == 0) {
                flush();
                unit = nullptr;
                in_multi = false;
                size_t q = line.find("_run(uint ctx_index");
                if (q != std::string::npos) {
                    auto it = prog.code_index.find(line.substr(5, q - 5));
                    if (it != prog.code_index.end() && !prog.codes[it->second].is_function) unit = &prog.codes[it->second];
                }
                create_unit = nullptr;
                q = line.find("_create(uint soffset");
                if (q != std::string::npos) {
                    auto it = prog.code_index.find(line.substr(5, q - 5));
                    if (it != prog.code_index.end() && !prog.codes[it->second].is_function) create_unit = &prog.codes[it->second];
                }
                continue;
            }
            if (create_unit) {       // ctx->componentMemory[coffset].templateId = <id>;   (template.rs:243-249)
                int64_t id;
                if (num_after(line, "].templateId = ", id)) create_unit->template_id = id;
            }
            if (!unit) continue;
            int64_t v;
            if (line.rfind("uint aux_create = ", 0) == 0) {
                in_multi = true;
                multi_positions.clear();
                multi_slot = atoll(line.c_str() + 18);
                multi_sym.clear();
                sig_jump = cmp_jump = 0;
                continue;
            }
            if (in_multi && line.rfind("uint aux_positions", 0) == 0) {
                // uint aux_positions [N]= {a,b,c};   (create_component_bucket.rs:260)
                multi_positions.clear();
                size_t q = line.find('{');
                while (q != std::string::npos && q + 1 < line.size() && line[q + 1] != '}') {
                    multi_positions.push_back(atoll(line.c_str() + q + 1));
                    q = line.find(',', q + 1);
                }
                continue;
            }
            if (in_multi) {
                if (line.rfind("int aux_cmp_num = ", 0) == 0) multi_cmp = atoll(line.c_str() + 18);
                else if (num_after(line, "uint csoffset = mySignalStart+", v)) multi_sig = v;
                else if (num_after(line, "for (uint i = 0; i < ", v)) multi_n = v;
                else if (num_after(line, "for (uint i_aux = 0; i_aux < ", v)) multi_n = v;
                else if (num_after(line, "csoffset += ", v)) sig_jump = v;
                else if (num_after(line, "aux_cmp_num += ", v)) {
                    cmp_jump = v;
                    Ins in;
                    in.op = OP_CREATE_CMP;
                    in.cc[0] = multi_slot; in.cc[1] = multi_sig; in.cc[2] = sig_jump; in.cc[3] = multi_cmp; in.cc[4] = cmp_jump; in.cc[5] = multi_n;
                    auto it = prog.code_index.find(multi_sym);
                    if (it == prog.code_index.end()) throw ParseError("unknown template " + multi_sym + " in the generated C++");
                    in.target = it->second;
                    in.positions = multi_positions;
                    found.push_back(in);
                    in_multi = false;
                } else {
                    size_t q = line.find("_create(csoffset,aux_cmp_num,");
                    if (q != std::string::npos) multi_sym = line.substr(0, q);
                }
                continue;
            }
            size_t q = line.find("_create(mySignalStart+");
            if (q != std::string::npos) {
                // <Sym>_create(mySignalStart+<sig>,<cmp>+ctx_index+1,ctx,new_cmp_name,myId);  followed by mySubcomponents[<slot>] = ...
                Ins in;
                in.op = OP_CREATE_CMP;
                in.cc[1] = atoll(line.c_str() + q + 22);
                size_t comma = line.find(',', q);
                in.cc[3] = comma == std::string::npos ? 0 : atoll(line.c_str() + comma + 1);
                in.cc[5] = 1;
                auto it = prog.code_index.find(line.substr(0, q));
                if (it == prog.code_index.end()) throw ParseError("unknown template " + line.substr(0, q) + " in the generated C++");
                in.target = it->second;
                in.cc[0] = -1;
                found.push_back(in);
                continue;
            }
            if (!found.empty() && found.back().cc[0] == -1 && num_after(line, "mySubcomponents[", v)) found.back().cc[0] = v;
        }
        flush();
    }

  private:

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
