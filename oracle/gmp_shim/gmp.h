/* Prototype-only stand-in for <gmp.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  This image ships the GMP runtime
 * (/usr/lib/x86_64-linux-gnu/libgmp.so.10, GMP 6.3.0) but not its development
 * header.  The reference's field arithmetic (code_producers/src/c_elements/
 * generic/fr.cpp) only needs the declarations below, which follow GMP's
 * documented public ABI (struct layout of mpz_t, the __gmpz_/__gmpn_ symbol
 * prefix).  Nothing here is used by the CUDA product path.
 */
#ifndef ORACLE_GMP_SHIM_H
#define ORACLE_GMP_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long int mp_limb_t;
typedef long int mp_limb_signed_t;
typedef long int mp_size_t;
typedef unsigned long int mp_bitcnt_t;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

/* ---- mpn layer ---- */
#define mpn_add __gmpn_add
#define mpn_add_1 __gmpn_add_1
#define mpn_add_n __gmpn_add_n
#define mpn_addmul_1 __gmpn_addmul_1
#define mpn_and_n __gmpn_and_n
#define mpn_cmp __gmpn_cmp
#define mpn_com __gmpn_com
#define mpn_copyi __gmpn_copyi
#define mpn_ior_n __gmpn_ior_n
#define mpn_lshift __gmpn_lshift
#define mpn_mul_1 __gmpn_mul_1
#define mpn_rshift __gmpn_rshift
#define mpn_sub_1 __gmpn_sub_1
#define mpn_sub_n __gmpn_sub_n
#define mpn_xor_n __gmpn_xor_n
#define mpn_zero_p __gmpn_zero_p

mp_limb_t __gmpn_add(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
void __gmpn_and_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
int __gmpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_com(mp_ptr, mp_srcptr, mp_size_t);
void __gmpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
void __gmpn_ior_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
mp_limb_t __gmpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t __gmpn_mul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
mp_limb_t __gmpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
mp_limb_t __gmpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
void __gmpn_xor_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
int __gmpn_zero_p(mp_srcptr, mp_size_t);

/* ---- mpz layer ---- */
#define mpz_add __gmpz_add
#define mpz_clear __gmpz_clear
#define mpz_export __gmpz_export
#define mpz_fdiv_q __gmpz_fdiv_q
#define mpz_fdiv_r __gmpz_fdiv_r
#define mpz_fits_sint_p __gmpz_fits_sint_p
#define mpz_get_si __gmpz_get_si
#define mpz_get_str __gmpz_get_str
#define mpz_import __gmpz_import
#define mpz_init __gmpz_init
#define mpz_init_set_si __gmpz_init_set_si
#define mpz_init_set_str __gmpz_init_set_str
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_invert __gmpz_invert
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_powm __gmpz_powm
#define mpz_set_si __gmpz_set_si
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_sub __gmpz_sub

void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_clear(mpz_ptr);
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void __gmpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
int __gmpz_fits_sint_p(mpz_srcptr);
signed long int __gmpz_get_si(mpz_srcptr);
char *__gmpz_get_str(char *, int, mpz_srcptr);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void __gmpz_init(mpz_ptr);
void __gmpz_init_set_si(mpz_ptr, signed long int);
int __gmpz_init_set_str(mpz_ptr, const char *, int);
void __gmpz_init_set_ui(mpz_ptr, unsigned long int);
int __gmpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
void __gmpz_set_si(mpz_ptr, signed long int);
size_t __gmpz_sizeinbase(mpz_srcptr, int);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_GMP_SHIM_H */
