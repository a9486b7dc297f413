/* Plain-C caller of libcvmgpu.so: what a Rust `extern "C"` / cgo / JNI binding sees.  Compiled by tests/test_abi.py with
 * gcc against include/cvmgpu.h; checks the size-versioned info structs and runs one small batch (on a box without a GPU
 * the compute call must fail with CVMGPU_ERR_CUDA -- there is no CPU fallback).
 * usage: abi_smoke <circuit.cvm> <a> <b>      (Multiplier2: witness = [1, a*b, a, b]) */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cvmgpu.h"

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    printf("sizeof_program_info %zu\n", sizeof(cvmgpu_program_info));
    printf("sizeof_r1cs_info %zu\n", sizeof(cvmgpu_r1cs_info));
    printf("offsetof_n_wires %zu\n", offsetof(cvmgpu_program_info, n_wires));
    printf("offsetof_tape_len %zu\n", offsetof(cvmgpu_program_info, tape_len));
    printf("offsetof_tape_int %zu\n", offsetof(cvmgpu_program_info, tape_int));
    printf("offsetof_bound_table_constraints %zu\n", offsetof(cvmgpu_r1cs_info, bound_table_constraints));
    cvmgpu_program *p = NULL;
    if (cvmgpu_program_load(argv[1], 0, &p) != CVMGPU_OK) {
        fprintf(stderr, "load: %s\n", cvmgpu_last_error());
        return 1;
    }
    /* a caller built against an OLDER header passes a smaller struct: only that many bytes may be written */
    struct { cvmgpu_program_info info; unsigned char guard[64]; } box;
    memset(&box, 0xAB, sizeof(box));
    box.info.struct_size = (uint32_t)offsetof(cvmgpu_program_info, tape_lut);
    if (cvmgpu_program_info_get(p, &box.info) != CVMGPU_OK) return 1;
    const unsigned char *tail = (const unsigned char *)&box.info + offsetof(cvmgpu_program_info, tape_lut);
    for (size_t k = 0; k < sizeof(box) - offsetof(cvmgpu_program_info, tape_lut); k++)
        if (tail[k] != 0xAB) {
            fprintf(stderr, "info_get wrote past struct_size\n");
            return 1;
        }
    cvmgpu_program_info info;
    memset(&info, 0, sizeof(info));
    if (cvmgpu_program_info_get(p, &info) == CVMGPU_OK) {       /* struct_size not set: must be refused, not guessed */
        fprintf(stderr, "info_get accepted struct_size = 0\n");
        return 1;
    }
    info.struct_size = sizeof(info);
    if (cvmgpu_program_info_get(p, &info) != CVMGPU_OK || info.struct_size != sizeof(info)) return 1;
    printf("n_wires %u n_inputs %u tape_len %llu\n", info.n_wires, info.n_inputs, (unsigned long long)info.tape_len);
    unsigned char in[64], out[4 * 32];
    uint32_t status = 99;
    memset(in, 0, sizeof(in));
    unsigned long a = strtoul(argv[2], NULL, 10), b = strtoul(argv[3], NULL, 10);
    memcpy(in, &a, sizeof(a));            /* little-endian canonical field elements */
    memcpy(in + 32, &b, sizeof(b));
    int rc = cvmgpu_witness_batch(p, in, 1, out, &status);
    if (rc == CVMGPU_OK) {
        unsigned long long c = 0;
        memcpy(&c, out + 32, 8);
        printf("batch ok status %u product %llu\n", status, c);
    } else {
        printf("batch rc %d %s\n", rc, cvmgpu_last_error());
    }
    rc = cvmgpu_witness_batch_multi(p, NULL, in, 1, 1u, 0, info.n_wires, out, &status, NULL);
    printf("multi rc %d\n", rc);
    cvmgpu_program_free(p);
    return 0;
}
