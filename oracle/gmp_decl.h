/* Hand-written prototypes for the handful of GMP 6.x entry points the CPU oracle uses.
 * The image ships libgmp.so.10 (GMP 6.3.0 — the backend of the reference's default
 * `curv-kzen/rust-gmp-kzen` feature, /root/reference/Cargo.toml:29) but not gmp.h.
 * TEST INFRASTRUCTURE ONLY (see oracle/gg20_oracle.py header). */
#ifndef ORACLE_GMP_DECL_H
#define ORACLE_GMP_DECL_H
#include <stddef.h>
typedef unsigned long mp_limb_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t* _mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct* mpz_ptr;
typedef const __mpz_struct* mpz_srcptr;
void __gmpz_init(mpz_ptr);
void __gmpz_clear(mpz_ptr);
void __gmpz_set(mpz_ptr, mpz_srcptr);
void __gmpz_set_ui(mpz_ptr, unsigned long);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void*);
void* __gmpz_export(void*, size_t*, int, size_t, int, size_t, mpz_srcptr);
void __gmpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
int __gmpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
size_t __gmpz_sizeinbase(mpz_srcptr, int);
extern const char* const __gmp_version;
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_powm __gmpz_powm
#define mpz_invert __gmpz_invert
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_mod __gmpz_mod
#define mpz_tdiv_q __gmpz_tdiv_q
#define mpz_gcd __gmpz_gcd
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#endif
