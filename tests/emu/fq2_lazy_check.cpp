// TEST INFRASTRUCTURE (host emulation of the carry-chain primitives):  Fq2::mul_lazy == Fq2::mul_karatsuba, Fq::mul_wide + redc_wide == Fq::mul
#define G16_EMU 1
#include "fq2.cuh"
#include <cstdio>
#include <random>
using namespace g16;
int main() {
    std::mt19937_64 rng(7);
    auto rnd = [&](bool edge) {
        Fq x;
        for (int i = 0; i < 12; ++i) x.l[i] = edge ? (rng() & 1 ? 0xffffffffu : 0u) : (uint32_t)rng();
        x.l[11] &= 0x0fffffffu;   // < 2^380 < q
        return Fq::to_mont(x);
    };
    int bad = 0;
    for (int it = 0; it < 20000; ++it) {
        bool e = it % 7 == 0;
        Fq2 a{rnd(e), rnd(it % 5 == 0)}, b{rnd(it % 3 == 0), rnd(e)};
        if (it == 1) a = Fq2::zero();
        if (it == 2) { a.c0 = Fq::neg(Fq::one()); a.c1 = Fq::neg(Fq::one()); b = a; }   // q - 1 everywhere
        if (it == 3) { a.c0 = Fq::zero(); b.c1 = Fq::neg(Fq::one()); }
        Fq2 x = Fq2::mul_lazy(a, b), y = Fq2::mul_karatsuba(a, b);
        if (!(x == y)) ++bad;
        uint32_t t[24];
        Fq::mul_wide(a.c0.l, b.c1.l, t);
        if (!(Fq::redc_wide(t) == Fq::mul(a.c0, b.c1))) ++bad;
    }
    printf("mismatches: %d\n", bad);
    return bad != 0;
}
