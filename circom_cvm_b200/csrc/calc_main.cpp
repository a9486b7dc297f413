// cvmgpu_calc -- native host program above the C ABI, with the process interface of the reference's generated
// witness calculator (code_producers/src/c_elements/common/main.cpp:334-371):
//
//     cvmgpu_calc <circuit.cvm> <input.json> <output.wtns> [--r1cs <circuit.r1cs>] [--sym <circuit.sym>]
//
// reads <circuit>.dat next to the program for the input hash map (as the reference reads <argv0>.dat,
// main.cpp:22-124), takes the same input.json (main.cpp:144-284: nested names qualified as a.b[i].c, numbers as
// decimal / 0x / 0b / 0o strings or JSON numbers, size and double-assignment checks of calcwit.cpp:51-97) and writes
// the same bytes to output.wtns (main.cpp:286-332) -- computed on the GPU through libcvmgpu.so.  input.json may
// also be an array of input objects: one launch, one file per witness (out.wtns, out.1.wtns, ...).
// Exit code 1 with a message where the reference aborts.  There is no CPU fallback: without a CUDA device the
// library call fails and so does this program.
#include <array>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cvmgpu.h"

namespace {

struct Fail : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- a small JSON reader (what loadJson needs)
struct JVal {
    enum Kind { NUL, BOOL, NUM_U, NUM_I, NUM_F, STR, ARR, OBJ } kind = NUL;
    std::string text;                                   // STR: contents; NUM_*: the literal
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;      // insertion order is irrelevant: keys are hashed
};

struct JParser {
    const std::string &s;
    size_t i = 0;
    explicit JParser(const std::string &t) : s(t) {}
    void ws() { while (i < s.size() && isspace((unsigned char)s[i])) i++; }
    [[noreturn]] void bad(const char *what) { throw Fail(std::string("invalid JSON input: ") + what); }
    JVal value() {
        ws();
        if (i >= s.size()) bad("unexpected end");
        char c = s[i];
        JVal v;
        if (c == '{') {
            v.kind = JVal::OBJ;
            i++;
            ws();
            if (i < s.size() && s[i] == '}') { i++; return v; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') bad("object key expected");
                std::string k = str();
                ws();
                if (i >= s.size() || s[i] != ':') bad("':' expected");
                i++;
                v.obj.emplace_back(k, value());
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == '}') { i++; break; }
                bad("',' or '}' expected");
            }
        } else if (c == '[') {
            v.kind = JVal::ARR;
            i++;
            ws();
            if (i < s.size() && s[i] == ']') { i++; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == ']') { i++; break; }
                bad("',' or ']' expected");
            }
        } else if (c == '"') {
            v.kind = JVal::STR;
            v.text = str();
        } else if (c == 't' && s.compare(i, 4, "true") == 0) { v.kind = JVal::BOOL; v.text = "true"; i += 4; }
        else if (c == 'f' && s.compare(i, 5, "false") == 0) { v.kind = JVal::BOOL; v.text = "false"; i += 5; }
        else if (c == 'n' && s.compare(i, 4, "null") == 0) { v.kind = JVal::NUL; i += 4; }
        else if (c == '-' || isdigit((unsigned char)c)) {
            size_t b = i;
            bool flt = false;
            if (s[i] == '-') i++;
            while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || s[i] == '+' || s[i] == '-')) {
                if (s[i] == '.' || s[i] == 'e' || s[i] == 'E') flt = true;
                i++;
            }
            v.text = s.substr(b, i - b);
            v.kind = flt ? JVal::NUM_F : (v.text[0] == '-' ? JVal::NUM_I : JVal::NUM_U);   // nlohmann's three number types
        } else bad("unexpected character");
        return v;
    }
    std::string str() {
        std::string out;
        i++;   // opening quote
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                if (++i >= s.size()) bad("bad escape");
                switch (s[i]) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': {
                        if (i + 4 >= s.size()) bad("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16);
                        i += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += s[i]; break;   // \" \\ \/
                }
                i++;
            } else out += s[i++];
        }
        if (i >= s.size()) bad("unterminated string");
        i++;
        return out;
    }
};

// ---------------------------------------------------------------- numbers -> 32-byte little-endian
// digits in `base` -> the value mod q as 32 bytes LE, like Fr_str2element (bn128/fr.cpp:56-62: mpz_set_str + mpz_fdiv_r):
// any number of digits; `negative` (only JSON numbers can be: they go through a double, main.cpp:167-172) gives q - value
void parse_number(const std::string &digits, int base, const std::string &shown, uint8_t out[32], bool negative = false) {
    static const uint32_t Q[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t limb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    (void)shown;
    auto geq_q = [&]() {
        if (limb[8]) return true;
        for (int k = 7; k >= 0; k--)
            if (limb[k] != Q[k]) return limb[k] > Q[k];
        return true;
    };
    auto sub_q = [&]() {
        uint64_t borrow = 0;
        for (int k = 0; k < 9; k++) {
            uint64_t d = (uint64_t)limb[k] - (k < 8 ? Q[k] : 0u) - borrow;
            limb[k] = (uint32_t)d;
            borrow = (d >> 32) & 1;
        }
    };
    for (char ch : digits) {
        int d;
        if (ch >= '0' && ch <= '9') d = ch - '0';
        else if (ch >= 'a' && ch <= 'f') d = ch - 'a' + 10;
        else if (ch >= 'A' && ch <= 'F') d = ch - 'A' + 10;
        else d = 99;
        if (d >= base) throw Fail("Invalid number in JSON input: " + shown);
        uint64_t carry = (uint64_t)d;
        for (int k = 0; k < 9; k++) {
            uint64_t t = (uint64_t)limb[k] * (uint64_t)base + carry;
            limb[k] = (uint32_t)t;
            carry = t >> 32;
        }
        while (geq_q()) sub_q();     // value < q before the step, so < 16 q + 16 after it
    }
    if (negative) {
        bool zero = true;
        for (int k = 0; k < 8; k++) zero = zero && limb[k] == 0;
        if (!zero) {
            uint64_t borrow = 0;
            for (int k = 0; k < 8; k++) {
                uint64_t d = (uint64_t)Q[k] - limb[k] - borrow;
                limb[k] = (uint32_t)d;
                borrow = (d >> 32) & 1;
            }
        }
    }
    memcpy(out, limb, 32);
}

void json_number(const JVal &v, uint8_t out[32]) {   // json2FrElements, main.cpp:144-190
    if (v.kind == JVal::STR) {
        const std::string &t = v.text;
        std::string p = t.substr(0, 2);
        if (p == "0b" || p == "0B") return parse_number(t.substr(2), 2, t, out);
        if (p == "0o" || p == "0O") return parse_number(t.substr(2), 8, t, out);
        if (p == "0x" || p == "0X") return parse_number(t.substr(2), 16, t, out);
        return parse_number(t, 10, t, out);
    }
    if (v.kind == JVal::NUM_U || v.kind == JVal::NUM_I || v.kind == JVal::NUM_F) {
        // the reference goes through a double: stream << std::fixed << std::setprecision(0) << val.get<double>()
        char buf[400];
        snprintf(buf, sizeof buf, "%.0f", strtod(v.text.c_str(), nullptr));
        std::string t = buf;
        const bool neg = !t.empty() && t[0] == '-';
        if (t.find_first_not_of("-0123456789") != std::string::npos) throw Fail("Invalid number in JSON input: " + t);   // inf / nan
        return parse_number(neg ? t.substr(1) : t, 10, t, out, neg);
    }
    throw Fail("Invalid JSON type");
}

void flatten(const JVal &v, std::vector<std::array<uint8_t, 32>> &out) {
    if (v.kind == JVal::ARR) {
        for (const JVal &e : v.arr) flatten(e, out);
    } else {
        std::array<uint8_t, 32> b;
        json_number(v, b.data());
        out.push_back(b);
    }
}

int elem_type(const std::string &prefix, const JVal &v) {   // check_type, main.cpp:192-207
    if (v.kind != JVal::ARR) return (int)v.kind;
    if (v.arr.empty()) return (int)JVal::NUL;
    int t = elem_type(prefix, v.arr[0]);
    for (size_t k = 1; k < v.arr.size(); k++)
        if (elem_type(prefix, v.arr[k]) != t) throw Fail("Types are not the same in the the key " + prefix);
    return t;
}

void qualify(const std::string &prefix, const JVal &v, std::map<std::string, const JVal *> &out);
void qualify_list(const std::string &prefix, const JVal &v, std::map<std::string, const JVal *> &out) {
    if (v.kind == JVal::ARR) {
        for (size_t k = 0; k < v.arr.size(); k++) qualify_list(prefix + "[" + std::to_string(k) + "]", v.arr[k], out);
    } else qualify(prefix, v, out);
}
void qualify(const std::string &prefix, const JVal &v, std::map<std::string, const JVal *> &out) {   // main.cpp:209-239
    if (v.kind == JVal::ARR) {
        if (!v.arr.empty() && elem_type(prefix, v) == (int)JVal::OBJ) qualify_list(prefix, v, out);
        else out[prefix] = &v;
    } else if (v.kind == JVal::OBJ) {
        for (const auto &kv : v.obj) qualify(prefix.empty() ? kv.first : prefix + "." + kv.first, kv.second, out);
    } else out[prefix] = &v;
}

uint64_t fnv1a(const std::string &s) {   // calcwit.cpp:17-24
    uint64_t h = 0xCBF29CE484222325ull;
    for (unsigned char c : s) {
        h ^= c;
        h *= 0x100000001B3ull;
    }
    return h;
}

// ---------------------------------------------------------------- <circuit>.dat: the input hash map (App. A.3)
struct HashEntry {
    uint64_t hash, signalid, size;
};
struct InputMap {
    std::vector<HashEntry> table;
    uint64_t input_start = 0, n_inputs = 0;
    size_t position(uint64_t h) const {   // calcwit.cpp:51-69
        size_t n = table.size(), pos = (size_t)(h % n);
        for (size_t k = 0; k < n; k++) {
            if (table[pos].hash == h) return pos;
            if (table[pos].signalid == 0) throw Fail("Signal not found");
            pos = (pos + 1) % n;
        }
        throw Fail("Signals not found");
    }
};

InputMap load_map(const std::string &dat_path, const uint64_t *witness, uint32_t n_wires, uint64_t input_start, uint64_t n_inputs) {
    std::ifstream f(dat_path, std::ios::binary);
    if (!f) throw Fail("cannot open " + dat_path);
    std::vector<char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    // the file does not carry its section sizes (the generated C++ does, circuit.rs:463-497): the map size is the power of
    // two >= 256 after which the program's witness-to-signal list follows
    InputMap m;
    m.input_start = input_start;
    m.n_inputs = n_inputs;
    for (size_t size = 256; 24 * size + 8 * (size_t)n_wires <= d.size(); size *= 2) {
        if (memcmp(d.data() + 24 * size, witness, 8 * (size_t)n_wires) == 0) {
            m.table.resize(size);
            memcpy(m.table.data(), d.data(), 24 * size);
            return m;
        }
    }
    throw Fail("the .dat file does not belong to this program (witness list not found)");
}

// The same table built from a `circom --sym` file (one "#s,#w,#c,name" line per signal,
// constraint_writers/src/sym_writer.rs:4-14): the main component's input signals grouped by name without their trailing
// indices ("main.in[3]" -> key "in", as loadJson addresses them), first signal and element count per key.
InputMap map_from_groups(const std::map<std::string, std::pair<uint64_t, uint64_t>> &groups, uint64_t input_start, uint64_t n_inputs) {
    InputMap m;
    m.input_start = input_start;
    m.n_inputs = n_inputs;
    size_t size = 256;
    while (size < 2 * groups.size()) size *= 2;
    m.table.assign(size, HashEntry{0, 0, 0});
    for (const auto &kv : groups) {
        const uint64_t h = fnv1a(kv.first);
        size_t pos = (size_t)(h % size);
        while (m.table[pos].signalid != 0) pos = (pos + 1) % size;
        m.table[pos] = HashEntry{h, kv.second.first, kv.second.second};
    }
    return m;
}

InputMap load_sym_map(const std::string &sym_path, uint64_t input_start, uint64_t n_inputs) {
    std::ifstream f(sym_path);
    if (!f) throw Fail("cannot open " + sym_path);
    std::map<std::string, std::pair<uint64_t, uint64_t>> groups;   // key -> (first signal, count)
    std::string line;
    uint64_t total = 0;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        size_t c1 = line.find(','), c2 = line.find(',', c1 + 1), c3 = line.find(',', c2 + 1);
        if (c1 == std::string::npos || c2 == std::string::npos || c3 == std::string::npos) throw Fail("malformed .sym line: " + line);
        const uint64_t label = strtoull(line.substr(0, c1).c_str(), nullptr, 10);
        std::string name = line.substr(c3 + 1);
        if (label < input_start || label >= input_start + n_inputs || name.compare(0, 5, "main.") != 0) continue;
        name = name.substr(5);
        while (!name.empty() && name.back() == ']') {   // strip trailing [i] groups
            size_t open = name.rfind('[');
            if (open == std::string::npos) break;
            name.erase(open);
        }
        auto it = groups.find(name);
        if (it == groups.end()) groups.emplace(name, std::make_pair(label, (uint64_t)1));
        else { it->second.first = std::min(it->second.first, label); it->second.second++; }
        total++;
    }
    if (total != n_inputs) throw Fail("the .sym file does not belong to this program (main inputs not found)");
    return map_from_groups(groups, input_start, n_inputs);
}

// the program's own `;;%%main_input <name> <first signal> <size>` lines (cvmgpu_program_main_inputs)
InputMap load_named_map(const cvmgpu_program *prog, uint64_t input_start, uint64_t n_inputs) {
    const char *text = nullptr;
    size_t len = 0;
    cvmgpu_program_main_inputs(prog, &text, &len);
    std::map<std::string, std::pair<uint64_t, uint64_t>> groups;
    std::istringstream ss(std::string(text ? text : "", len));
    std::string name;
    uint64_t start, size, total = 0;
    while (ss >> name >> start >> size) {
        groups[name] = std::make_pair(start, size);
        total += size;
    }
    if (groups.empty() || total != n_inputs)
        throw Fail("no <circuit>.dat next to the program, no --sym file, and the program text does not name its main inputs");
    return map_from_groups(groups, input_start, n_inputs);
}

void row_from_json(const InputMap &m, const JVal &doc, uint8_t *row) {   // loadJson + setInputSignal
    std::map<std::string, const JVal *> flat;
    qualify("", doc, flat);
    std::vector<char> set(m.n_inputs, 0);
    uint64_t n_set = 0;
    for (const auto &kv : flat) {
        size_t pos;
        try {
            pos = m.position(fnv1a(kv.first));
        } catch (const Fail &e) {
            throw Fail("Error loading signal " + kv.first + ": " + e.what());
        }
        std::vector<std::array<uint8_t, 32>> vals;
        flatten(*kv.second, vals);
        if (vals.size() < m.table[pos].size) throw Fail("Error loading signal " + kv.first + ": Not enough values");
        if (vals.size() > m.table[pos].size) throw Fail("Error loading signal " + kv.first + ": Too many values");
        for (size_t k = 0; k < vals.size(); k++) {
            uint64_t si = m.table[pos].signalid + k;
            if (si < m.input_start || si - m.input_start >= m.n_inputs) throw Fail("Error setting signal: " + kv.first);
            uint64_t idx = si - m.input_start;
            if (set[idx]) throw Fail("Error setting signal: " + kv.first + "\nSignal assigned twice: " + std::to_string(si));
            set[idx] = 1;
            n_set++;
            memcpy(row + 32 * idx, vals[k].data(), 32);
        }
    }
    if (n_set != m.n_inputs)
        throw Fail("Not all inputs have been set. Only " + std::to_string(n_set) + " out of " + std::to_string(m.n_inputs));
}

std::string slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Fail("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string strip_ext(const std::string &p) {
    size_t dot = p.find_last_of('.'), slash = p.find_last_of('/');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return p;
    return p.substr(0, dot);
}

}  // namespace

int main(int argc, char *argv[]) {
    std::vector<std::string> pos;
    std::string r1cs_path, sym_path, cpp_path;
    for (int k = 1; k < argc; k++) {
        std::string a = argv[k];
        if (a == "--r1cs" && k + 1 < argc) r1cs_path = argv[++k];
        else if (a == "--sym" && k + 1 < argc) sym_path = argv[++k];
        else if (a == "--cpp" && k + 1 < argc) cpp_path = argv[++k];
        else pos.push_back(a);
    }
    if (pos.size() != 3) {
        fprintf(stderr, "Usage: %s <circuit.cvm> <input.json> <output.wtns> [--r1cs <circuit.r1cs>] [--sym <circuit.sym>] "
                        "[--cpp <circuit.cpp>]\n", argv[0]);
        return 1;
    }
    cvmgpu_program *prog = nullptr;
    cvmgpu_r1cs *r1 = nullptr;
    int rc = 1;
    try {
        // the generated C++ of the same compile (component creation for an unpatched emitter; section sizes of the .dat)
        // and the .dat (io-map of mixed component arrays) are picked up next to the program when they are there
        const std::string stem = strip_ext(pos[0]);
        if (cpp_path.empty() && std::ifstream(stem + ".cpp").good()) cpp_path = stem + ".cpp";
        const std::string dat_path = stem + ".dat";
        const bool have_dat = !cpp_path.empty() && std::ifstream(dat_path).good();
        if (cvmgpu_program_load_files(pos[0].c_str(), cpp_path.empty() ? nullptr : cpp_path.c_str(),
                                      have_dat ? dat_path.c_str() : nullptr, 0, &prog) != CVMGPU_OK)
            throw Fail(cvmgpu_last_error());
        cvmgpu_program_info info;
        info.struct_size = sizeof(info);
        if (cvmgpu_program_info_get(prog, &info) != CVMGPU_OK) throw Fail(cvmgpu_last_error());
        const uint64_t *witness = nullptr;
        uint32_t n_wit = 0;
        cvmgpu_program_witness(prog, &witness, &n_wit);
        InputMap m = !sym_path.empty()                  ? load_sym_map(sym_path, 1 + info.n_outputs, info.n_inputs)
                     : std::ifstream(dat_path).good() ? load_map(dat_path, witness, n_wit, 1 + info.n_outputs, info.n_inputs)
                                                      : load_named_map(prog, 1 + info.n_outputs, info.n_inputs);
        std::string text = slurp(pos[1]);
        JParser jp(text);
        JVal doc = jp.value();
        std::vector<const JVal *> docs;
        if (doc.kind == JVal::ARR) {
            for (const JVal &e : doc.arr) docs.push_back(&e);
        } else docs.push_back(&doc);
        const uint64_t B = docs.size();
        std::vector<uint8_t> inputs((size_t)B * info.n_inputs * 32, 0);
        for (uint64_t b = 0; b < B; b++) row_from_json(m, *docs[b], inputs.data() + (size_t)b * info.n_inputs * 32);
        std::vector<uint8_t> wtns((size_t)B * info.n_wires * 32);
        std::vector<uint32_t> status(B, 0), bad(B, 0xffffffffu);
        if (!r1cs_path.empty() && cvmgpu_r1cs_load(r1cs_path.c_str(), &r1) != CVMGPU_OK) throw Fail(cvmgpu_last_error());
        if (cvmgpu_witness_batch_checked(prog, r1, inputs.data(), B, wtns.data(), status.data(), r1 ? bad.data() : nullptr) != CVMGPU_OK)
            throw Fail(cvmgpu_last_error());
        rc = 0;
        std::string base = strip_ext(pos[2]), ext = pos[2].substr(base.size());
        for (uint64_t b = 0; b < B; b++) {
            if (status[b] != 0) {
                fprintf(stderr, "witness %llu: failed assert / toInt / division (status %u)\n", (unsigned long long)b, status[b]);
                rc = 1;
                continue;
            }
            if (r1 && bad[b] != 0xffffffffu) {
                fprintf(stderr, "witness %llu: constraint %u is not satisfied\n", (unsigned long long)b, bad[b]);
                rc = 1;
            }
            std::string out = b == 0 ? pos[2] : base + "." + std::to_string(b) + ext;
            if (cvmgpu_wtns_write(out.c_str(), wtns.data() + (size_t)b * info.n_wires * 32, info.n_wires) != CVMGPU_OK)
                throw Fail(cvmgpu_last_error());
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        rc = 1;
    }
    if (r1) cvmgpu_r1cs_free(r1);
    if (prog) cvmgpu_program_free(prog);
    return rc;
}
