// Host prototype (scaffold for the next round, NOT part of the library): BLS12-381 Fq Montgomery multiplication on
// 8 x 48-bit limbs held in doubles, every limb product formed on the FP64 pipe the way Emmart et al. do it,
//     hi = fma_rz(a, b, 2^100)                 -> 2^100 + floor(a b / 2^48) 2^48      (ulp of that binade is 2^48)
//     lo = fma_rz(a, b, (2^100 + 2^52) - hi)   -> 2^52 + (a b mod 2^48)
// and the raw 64-bit patterns of hi / lo accumulated as integers (their exponent fields are constants that come off
// at the end of a column).  R = 2^384 as in the 12 x 32-bit representation, so the results are the same integers as
// Fq::mul's -- which is what main() checks, on random and edge operands.  tools/dfma_peak.cu measures the inner step
// on B200 at 9.06 T limb products/s (bound by its two DFMAs; the DADD and the two integer adds ride along), i.e. a
// ceiling of 70 G Fq-mul/s at 128 products per multiplication against 30.4 G/s for the IMAD.WIDE implementation.
//
//   g++ -O2 -std=c++17 -frounding-math -I ../csrc -o dpf_mul_prototype dpf_mul_prototype.cpp && ./dpf_mul_prototype
#define G16_EMU 1
#include <cfenv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include "fp.cuh"

using namespace g16;
typedef unsigned long long u64;
typedef unsigned __int128 u128;

static const double C1 = 1267650600228229401496703205376.0;                       // 2^100
static const double C2 = 1267650600228229401496703205376.0 + 4503599627370496.0;   // 2^100 + 2^52
static const u64 MASK48 = (1ull << 48) - 1;

static u64 bits_of(double d) { u64 u; memcpy(&u, &d, 8); return u; }

// 12 x 32-bit little-endian limbs <-> 8 x 48-bit limbs
static void to48(const uint32_t *l, u64 *o) {
    for (int i = 0; i < 8; ++i) {
        int bit = 48 * i, w = bit >> 5, sh = bit & 31;
        u128 v = 0;
        for (int k = 0; k < 3 && w + k < 12; ++k) v |= (u128)l[w + k] << (32 * k);
        o[i] = (u64)(v >> sh) & MASK48;
    }
}
static void from48(const u64 *a, uint32_t *l) {
    memset(l, 0, 48);
    for (int i = 0; i < 8; ++i) {
        int bit = 48 * i, w = bit >> 5, sh = bit & 31;
        u128 v = (u128)a[i] << sh;
        for (int k = 0; k < 3 && w + k < 12; ++k) l[w + k] |= (uint32_t)(v >> (32 * k));
    }
}

struct Dpf {
    double q[8];   // modulus limbs
    u64 ninv;      // -q^-1 mod 2^48
    u64 k_hi, k_lo;   // exponent patterns of the hi / lo results
    Dpf() {
        uint32_t m[12];
        for (int i = 0; i < 12; ++i) m[i] = FqParams::MOD(i);
        u64 q48[8];
        to48(m, q48);
        for (int i = 0; i < 8; ++i) q[i] = (double)q48[i];
        u64 inv = 1;   // Newton: q0 * inv = 1 mod 2^64
        for (int k = 0; k < 6; ++k) inv *= 2 - q48[0] * inv;
        ninv = (0 - inv) & MASK48;
        k_hi = bits_of(C1);
        k_lo = bits_of(4503599627370496.0);
    }
    // one limb product into the (hi, lo) integer accumulators of its column
    static inline void product(double a, double b, u64 &acc_hi, u64 &acc_lo) {
        double hi = std::fma(a, b, C1);          // rounding mode: toward zero (set in main)
        double sub = C2 - hi;
        double lo = std::fma(a, b, sub);
        acc_hi += bits_of(hi);
        acc_lo += bits_of(lo);
    }
    // product-scanning Montgomery multiplication: 64 + 64 limb products
    void mul(const u64 *a48, const u64 *b48, u64 *r48) const {
        double a[8], b[8], m[8];
        for (int i = 0; i < 8; ++i) { a[i] = (double)a48[i]; b[i] = (double)b48[i]; }
        u64 carry = 0, prev_hi = 0;   // prev_hi: sum of the hi parts of the previous column
        u64 out[8];
        for (int k = 0; k < 16; ++k) {
            u64 acc_hi = 0, acc_lo = 0;
            int terms = 0;
            for (int i = (k < 8 ? 0 : k - 7); i <= (k < 8 ? k : 7); ++i) {
                product(a[i], b[k - i], acc_hi, acc_lo);
                ++terms;
            }
            for (int i = (k < 8 ? 0 : k - 7); i <= 7 && i < k; ++i) {   // m_i known for i < k
                if (k - i > 7) continue;
                product(m[i], q[k - i], acc_hi, acc_lo);
                ++terms;
            }
            u64 col = (acc_lo - (u64)terms * k_lo) + prev_hi + carry;
            u64 hi_sum = acc_hi - (u64)terms * k_hi;
            if (k < 8) {
                u64 mk = ((col & MASK48) * ninv) & MASK48;
                m[k] = (double)mk;
                u64 h2 = 0, l2 = 0;
                product(m[k], q[0], h2, l2);
                col += l2 - k_lo;
                hi_sum += h2 - k_hi;
                // col is now divisible by 2^48
                carry = col >> 48;
            } else {
                out[k - 8] = col & MASK48;
                carry = col >> 48;
            }
            prev_hi = hi_sum;
        }
        // value = out + (prev_hi + carry) * 2^(48 * 8) would exceed 384 bits only if the result >= 2^384: it is < 2q
        // conditional subtraction of q
        u64 qq[8];
        for (int i = 0; i < 8; ++i) qq[i] = (u64)q[i];
        u64 t[8];
        long long borrow = 0;
        for (int i = 0; i < 8; ++i) {
            long long d = (long long)out[i] - (long long)qq[i] - borrow;
            borrow = d < 0;
            t[i] = (u64)(d + (borrow ? (long long)(1ull << 48) : 0)) & MASK48;
        }
        bool ge = !borrow || (prev_hi + carry) != 0;
        for (int i = 0; i < 8; ++i) r48[i] = ge ? t[i] : out[i];
    }
    // operand scanning: all 64 products of a * b first (independent of the reduction), then eight reduction rows, each
    // adding m_k * q to the columns above it; only m_k and the carry stay on the critical path
    void mul_rows(const u64 *a48, const u64 *b48, u64 *r48) const {
        double a[8], b[8];
        for (int i = 0; i < 8; ++i) { a[i] = (double)a48[i]; b[i] = (double)b48[i]; }
        u64 hi[16] = {0}, lo[16] = {0};
        int cnt[16] = {0};
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) { product(a[i], b[j], hi[i + j], lo[i + j]); ++cnt[i + j]; }
        u64 carry = 0, out[8];
        for (int k = 0; k < 16; ++k) {
            u64 col = (lo[k] - (u64)cnt[k] * k_lo) + (k ? hi[k - 1] - (u64)cnt[k - 1] * k_hi : 0) + carry;
            if (k < 8) {
                u64 mk = ((col & MASK48) * ninv) & MASK48;
                double m = (double)mk;
                for (int j = 0; j < 8; ++j) { product(m, q[j], hi[k + j], lo[k + j]); ++cnt[k + j]; }
                // the j = 0 product landed in this column after `col` was formed: add its low part
                u64 h0 = 0, l0 = 0;
                product(m, q[0], h0, l0);
                col += l0 - k_lo;
            } else {
                out[k - 8] = col & MASK48;
            }
            carry = col >> 48;
        }
        u64 top = (hi[15] - (u64)cnt[15] * k_hi) + carry;
        u64 qq[8], t[8];
        for (int i = 0; i < 8; ++i) qq[i] = (u64)q[i];
        long long borrow = 0;
        for (int i = 0; i < 8; ++i) {
            long long d = (long long)out[i] - (long long)qq[i] - borrow;
            borrow = d < 0;
            t[i] = (u64)(d + (borrow ? (long long)(1ull << 48) : 0)) & MASK48;
        }
        bool ge = !borrow || top != 0;
        for (int i = 0; i < 8; ++i) r48[i] = ge ? t[i] : out[i];
    }
};

int main() {
    std::fesetround(FE_TOWARDZERO);
    Dpf dpf;
    std::mt19937_64 rng(11);
    int bad = 0;
    for (int it = 0; it < 200000; ++it) {
        Fq x, y;
        for (int i = 0; i < 12; ++i) {
            bool e = it % 11 == 0;
            x.l[i] = e ? ((rng() & 1) ? 0xffffffffu : 0u) : (uint32_t)rng();
            y.l[i] = (it % 13 == 0) ? ((rng() & 1) ? 0xffffffffu : 0u) : (uint32_t)rng();
        }
        x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu;     // < 2^380 < q
        if (it == 1) x = Fq::zero();
        if (it == 2) { x = Fq::neg(Fq::one()); y = x; }
        if (it == 3) { for (int i = 0; i < 12; ++i) x.l[i] = FqParams::MOD(i); x.l[0] -= 1; y = x; }   // q - 1
        Fq want = Fq::mul(x, y);
        u64 a48[8], b48[8], r48[8];
        to48(x.l, a48); to48(y.l, b48);
        dpf.mul(a48, b48, r48);
        Fq got;
        from48(r48, got.l);
        if (!(got == want)) ++bad;
        dpf.mul_rows(a48, b48, r48);
        from48(r48, got.l);
        if (!(got == want)) ++bad;
    }
    printf("dpf mismatches: %d\n", bad);
    return bad != 0;
}
