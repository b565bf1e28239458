// Host harness for tests/test_field_inverse_host.py: the 31-steps-at-a-time binary GCD inversion of csrc/field.cuh (the same
// code the device runs) against Fermat (a^(p-2)) and the plain binary Euclid loop, for Fq and Fr.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "field.cuh"
using namespace kzg;
template <class P> int run(const char* name, int n) {
    std::mt19937_64 rng(12345);
    int bad = 0;
    for (int t = 0; t < n; t++) {
        Fp<P> a;
        for (int i = 0; i < 8; i++) a.l[i] = (uint32_t)rng();
        if (t % 7 == 1) for (int i = 2; i < 8; i++) a.l[i] = 0;         // short
        if (t % 7 == 2) { for (int i = 0; i < 8; i++) a.l[i] = 0; a.l[t % 8] = 1u << (t % 32); }  // power of two
        if (t % 7 == 3) { for (int i = 0; i < 8; i++) a.l[i] = P::mod(i); a.l[0] -= 1 + (t % 5); }  // p - small
        if (t == 0) { for (int i = 0; i < 8; i++) a.l[i] = 0; a.l[0] = 1; }
        a.l[7] &= 0x3fffffffu;
        while (fp_geq_mod<P>(a.l)) a.l[7] >>= 1;
        Fp<P> x = fp_inv(a), y = fp_inv_fermat(a), z = fp_inv_euclid(a);
        if (!fp_eq(x, y) || !fp_eq(z, y)) { bad++; if (bad < 5) printf("%s mismatch at %d\n", name, t); }
        if (!fp_is_zero(a) && !fp_eq(fp_mul(a, x), fp_one<P>())) { bad++; }
    }
    printf("%s: %d cases, %d bad\n", name, n, bad);
    return bad;
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 20000;
    int bad = run<FqP>("Fq", n) + run<FrP>("Fr", n);
    Fr z = fp_zero<FrP>();
    if (!fp_is_zero(fp_inv(z))) bad++;
    return bad ? 1 : 0;
}
