// TEST / BASELINE INFRASTRUCTURE ONLY.  main() for circuit binaries built from the REFERENCE runtime
// (code_producers/src/c_elements/common/{main,calcwit}.cpp + generic/fr.cpp, compiled by oracle/build_ref.py).
//
//   <circuit> <input.json> <output.wtns>      the reference's own main(), unchanged (common/main.cpp:334-371;
//                                             its main symbol renamed by objcopy)
//   <circuit> --bench <seconds> <seed> [mode [file]]
//                                             time run(ctx) only, one witness at a time; inputs: mode 0 random field
//                                             elements, 1 random bits, 2 rows of <file> (n_inputs x 32-byte LE
//                                             canonical values each, main-input signal order), cycled;
//                                             prints {"witnesses": n, "witnesses_per_s": x}
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "calcwit.hpp"
#include "circom.hpp"

extern "C" int circom_reference_main(int argc, char *argv[]);
Circom_Circuit *loadCircuit(std::string const &datFileName);

static uint64_t rng_state;
static uint64_t next64() {   // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char *argv[]) {
    if (argc >= 2 && strcmp(argv[1], "--bench") == 0) {
        double seconds = argc > 2 ? atof(argv[2]) : 5.0;
        rng_state = argc > 3 ? strtoull(argv[3], 0, 10) : 1;
        int bits_only = argc > 4 ? atoi(argv[4]) : 0;     // 1: inputs are bits (SHA-256); 2: rows of a file
        Circom_Circuit *circuit = loadCircuit(std::string(argv[0]) + ".dat");
        uint n_in = get_main_input_signal_no();
        uint start = get_main_input_signal_start();
        std::vector<unsigned char> pool;
        size_t pool_rows = 0;
        if (bits_only == 2) {
            FILE *f = argc > 5 ? fopen(argv[5], "rb") : nullptr;
            if (!f) { fprintf(stderr, "cannot open the input pool\n"); return 2; }
            unsigned char buf[4096];
            size_t got;
            while ((got = fread(buf, 1, sizeof buf, f)) > 0) pool.insert(pool.end(), buf, buf + got);
            fclose(f);
            pool_rows = pool.size() / ((size_t)n_in * 32);
            if (pool_rows == 0) { fprintf(stderr, "empty input pool\n"); return 2; }
        }
        // input hash map entries give (hash, first signal, size) for every main input
        uint map_n = get_size_of_input_hashmap();
        double busy = 0;
        uint64_t n = 0;
        auto t_begin = std::chrono::steady_clock::now();
        while (true) {
            Circom_CalcWit *ctx = new Circom_CalcWit(circuit);
            uint remaining = n_in;
            for (uint e = 0; e < map_n; e++) {
                HashSignalInfo &h = circuit->InputHashMap[e];
                if (h.signalid == 0) continue;
                for (uint i = 0; i < h.signalsize; i++) {
                    FrElement v;
                    if (bits_only == 2) {
                        const unsigned char *src = &pool[(((size_t)n % pool_rows) * n_in + (h.signalid - start + i)) * 32];
                        v.type = Fr_LONG;
                        v.shortVal = 0;
                        memcpy(v.longVal, src, 32);
                    } else if (bits_only) {
                        v.type = Fr_SHORT;
                        v.shortVal = (int32_t)(next64() & 1);
                    } else {
                        v.type = Fr_LONG;
                        v.shortVal = 0;
                        for (int k = 0; k < 4; k++) v.longVal[k] = next64();
                        v.longVal[3] &= 0x1fffffffffffffffull;   // < 2^253 < q: canonical
                    }
                    remaining--;
                    if (remaining == 0) {
                        auto t0 = std::chrono::steady_clock::now();
                        ctx->setInputSignal(h.hash, i, v);        // the last input triggers run(ctx) (calcwit.cpp:71-97)
                        auto t1 = std::chrono::steady_clock::now();
                        busy += std::chrono::duration<double>(t1 - t0).count();
                    } else {
                        ctx->setInputSignal(h.hash, i, v);
                    }
                }
            }
            (void)start;
            n++;
            delete[] ctx->signalValues;        // ~Circom_CalcWit frees nothing (calcwit.cpp:47-49)
            delete[] ctx->componentMemory;
            delete ctx;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() >= seconds) break;
        }
        printf("{\"witnesses\": %llu, \"witnesses_per_s\": %.3f, \"busy_s\": %.3f}\n", (unsigned long long)n, n / busy, busy);
        return 0;
    }
    return circom_reference_main(argc, argv);
}
