// Test hooks: element-wise field and group operations exposed through the C ABI so that the
// parity tests can hit the device arithmetic directly (edge operands, exceptional group cases).
#pragma once
#include "msm_kernels.cuh"
#include "fixed_base_kernels.cuh"

namespace g16 {

struct DebugFqOp {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
        Fq x, y, r;
        for (int j = 0; j < 12; ++j) { x.l[j] = a[12 * i + j]; y.l[j] = b ? b[12 * i + j] : 0u; }
        switch (op) {
            case 0: r = Fq::mul(x, y); break;
            case 1: r = Fq::add(x, y); break;
            case 2: r = Fq::sub(x, y); break;
            case 3: r = Fq::inv(x); break;
            case 4: r = Fq::sqr(x); break;
            case 5: r = Fq::neg(x); break;
            case 6: r = Fq::mul_dual(x, y, Fq::sqr(x), Fq::add(x, y)); break;   // (x y + x^2 (x + y)) R^-1
            case 7: r = Fq::mul_quad(x, y, Fq::sqr(x), Fq::add(x, y), y, y, Fq::neg(x), x); break;   // ... + y^2 - x^2
            case 8: r = Fq::mul_diff(x, y, Fq::sqr(x), Fq::add(x, y)); break;   // x y - x^2 (x + y)
            default: r = Fq::zero();
        }
        for (int j = 0; j < 12; ++j) out[12 * i + j] = r.l[j];
    }
};

struct DebugFrFromMont {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *a, uint32_t *out) {
        uint32_t k[8];
        load_scalar(a, i, true, k);
        for (int j = 0; j < 8; ++j) out[8 * i + j] = k[j];
    }
};

// out = affine(P + Q) via from_affine(P) then the mixed addition of Q
template <class F>
struct DebugAdd {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t i, const uint32_t *p, const uint32_t *q, uint32_t *out) {
        Affine<F> a = load_affine<F>(p, i), b = load_affine<F>(q, i);
        XYZZ<F> acc = XYZZ<F>::from_affine(a);
        xyzz_madd(acc, b.x, b.y);
        store_affine<F>(out, i, xyzz_to_affine(acc));
    }
};

}  // namespace g16
