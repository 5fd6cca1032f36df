// Wire format of group elements: ark `CanonicalSerialize` / `CanonicalDeserialize` as ark-bls12-381 0.4.0
// implements it for G1Affine / G2Affine (the Zcash / IETF encoding), one thread per point.
//
// The reference derives it for `Proof` (/root/reference/crates/groth16-core/src/lib.rs:27-36: a || b || c,
// 192 bytes compressed, 384 uncompressed); a key file (the CLI's `save_proving_key` is a stub,
// crates/groth16-cli/src/lib.rs:157-168) would hold `Vec<G1Affine>` / `Vec<G2Affine>` of CRS length in the
// same encoding -- hence batch kernels.
//
//   Fq         48 bytes, big-endian, canonical (not Montgomery)
//   G1         compressed: x;  uncompressed: x || y
//   G2         compressed: x.c1 || x.c0;  uncompressed: x.c1 || x.c0 || y.c1 || y.c0
//   flags      top three bits of byte 0: 0x80 compressed, 0x40 infinity (all other bits zero),
//              0x20 (compressed, finite) y is the larger of {y, -y} (Fq2 ordered by c1, then c0)
//
// Decoding follows read_g1_compressed / read_g1_uncompressed + `deserialize_with_mode`: wrong compression
// flag -> UnexpectedFlags; infinity flag -> identity; coordinate >= q or no square root -> InvalidData;
// with validation on, a point outside the prime-order subgroup (or, for uncompressed input, off the curve)
// -> InvalidData.  The square root only has to be *a* root: the sign flag picks the result.
#pragma once
#include "msm_kernels.cuh"

namespace g16 {

constexpr uint32_t WIRE_OK = 0, WIRE_INVALID_DATA = 1, WIRE_UNEXPECTED_FLAGS = 2;

struct WireConst {
    G16_HD static constexpr uint32_t SQRT_EXP(int i) {   // (q + 1) / 4
        constexpr uint32_t m[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                    0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};
        return m[i];
    }
    G16_HD static constexpr uint32_t HALF_Q(int i) {     // (q - 1) / 2
        constexpr uint32_t m[12] = {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u,
                                    0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u};
        return m[i];
    }
    G16_HD static constexpr uint32_t B4(int i) {         // 4 in Montgomery form
        constexpr uint32_t m[12] = {0x000cfff3u, 0xaa270000u, 0xfc34000au, 0x53cc0032u, 0x6b0a807fu, 0x478fe97au,
                                    0xe6ba24d7u, 0xb1d37ebeu, 0xbf78ab2fu, 0x8ec9733bu, 0x3d83de7eu, 0x09d64551u};
        return m[i];
    }
};

// ---- field helpers -----------------------------------------------------------------------------------
G16_HD Fq fq_const_b4() {
    Fq r;
    for (int i = 0; i < 12; ++i) r.l[i] = WireConst::B4(i);
    return r;
}
// a^((q+1)/4): the square root of a when a is a square (q = 3 mod 4)
G16_HD Fq fq_sqrt_candidate(const Fq &a) {
    Fq r = Fq::one();
    for (int i = 11; i >= 0; --i) {
        uint32_t w = WireConst::SQRT_EXP(i);
        for (int b = 31; b >= 0; --b) {
            r = Fq::sqr(r);
            if ((w >> b) & 1u) r = Fq::mul(r, a);
        }
    }
    return r;
}
G16_HD bool fq_sqrt(const Fq &a, Fq &root) {
    root = fq_sqrt_candidate(a);
    return Fq::sqr(root) == a;
}
// canonical value > (q - 1) / 2
G16_HD bool fq_lex_largest(const Fq &mont) {
    Fq c = Fq::from_mont(mont);
    uint32_t h[12];
    for (int i = 0; i < 12; ++i) h[i] = WireConst::HALF_Q(i);
    return !Fq::geq_raw(h, c.l);   // !(half >= c)
}
// canonical limbs < q
G16_HD bool fq_is_canonical(const uint32_t *l) {
    uint32_t m[12];
    for (int i = 0; i < 12; ++i) m[i] = FqParams::MOD(i);
    return !Fq::geq_raw(l, m);
}

// 48 big-endian bytes <-> 12 little-endian u32 limbs (the byte buffers are 4-byte aligned: 48 | offsets)
G16_HD uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
G16_HD void fq_put_be(uint32_t *dst_words, const Fq &canonical) {
    for (int j = 0; j < 12; ++j) dst_words[j] = bswap32(canonical.l[11 - j]);
}
G16_HD void fq_get_be(const uint32_t *src_words, uint32_t *limbs_out) {
    for (int j = 0; j < 12; ++j) limbs_out[11 - j] = bswap32(src_words[j]);
}

// ---- per-field adapters ----------------------------------------------------------------------------------
template <class F> struct Wire;
template <> struct Wire<Fq> {
    static constexpr int FQS = 1;   // Fq elements per coordinate
    G16_HD static void put(uint32_t *dst, const Fq &v) { fq_put_be(dst, Fq::from_mont(v)); }
    // returns false when the value is not canonical; `mask` clears the three flag bits first
    G16_HD static bool get(const uint32_t *src, bool mask, Fq &v) {
        uint32_t l[12];
        fq_get_be(src, l);
        if (mask) l[11] &= 0x1fffffffu;
        if (!fq_is_canonical(l)) return false;
        Fq c;
        for (int i = 0; i < 12; ++i) c.l[i] = l[i];
        v = Fq::to_mont(c);
        return true;
    }
    G16_HD static bool lex_largest(const Fq &y) { return fq_lex_largest(y); }
    G16_HD static Fq curve_b() { return fq_const_b4(); }
    G16_HD static bool sqrt(const Fq &a, Fq &root) { return fq_sqrt(a, root); }
};
template <> struct Wire<Fq2> {
    static constexpr int FQS = 2;
    G16_HD static void put(uint32_t *dst, const Fq2 &v) {   // c1 first
        fq_put_be(dst, Fq::from_mont(v.c1));
        fq_put_be(dst + 12, Fq::from_mont(v.c0));
    }
    G16_HD static bool get(const uint32_t *src, bool mask, Fq2 &v) {
        return Wire<Fq>::get(src, mask, v.c1) & Wire<Fq>::get(src + 12, false, v.c0);
    }
    G16_HD static bool lex_largest(const Fq2 &y) { return y.c1.is_zero() ? fq_lex_largest(y.c0) : fq_lex_largest(y.c1); }
    G16_HD static Fq2 curve_b() { return Fq2{fq_const_b4(), fq_const_b4()}; }   // 4 (1 + u)
    // complex method for Fq[u] / (u^2 + 1), q = 3 mod 4
    G16_HD static bool sqrt(const Fq2 &a, Fq2 &root) {
        if (a.c1.is_zero()) {
            Fq s;
            if (fq_sqrt(a.c0, s)) { root = Fq2{s, Fq::zero()}; return true; }
            if (fq_sqrt(Fq::neg(a.c0), s)) { root = Fq2{Fq::zero(), s}; return true; }
            return false;   // unreachable: -1 is a non-residue, so a0 or -a0 is a square
        }
        Fq alpha;
        if (!fq_sqrt(Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1)), alpha)) return false;   // norm not a square: no root
        // 1/2 in Montgomery form = (q + 1) / 2 * R ... simpler: halve through the inverse of 2
        Fq two = Fq::dbl(Fq::one());
        Fq half = Fq::inv(two);
        Fq delta = Fq::mul(Fq::add(a.c0, alpha), half);
        Fq x0;
        if (!fq_sqrt(delta, x0)) {
            delta = Fq::mul(Fq::sub(a.c0, alpha), half);
            if (!fq_sqrt(delta, x0)) return false;
        }
        Fq x1 = Fq::mul(a.c1, Fq::inv(Fq::dbl(x0)));
        root = Fq2{x0, x1};
        return Fq2::sqr(root) == a;
    }
};

// y^2 == x^3 + b
template <class F>
G16_HD bool on_curve(const Affine<F> &p) {
    F rhs = F::add(F::mul(F::sqr(p.x), p.x), Wire<F>::curve_b());
    return F::sqr(p.y) == rhs;
}
// r * P == O, P finite and on the curve (left-to-right double and add over the 255 bits of r)
template <class F>
G16_HD bool in_subgroup(const Affine<F> &p) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int i = 7; i >= 0; --i) {
        uint32_t w = FrParams::MOD(i);
        for (int b = 31; b >= 0; --b) {
            xyzz_dbl_call(acc);
            if ((w >> b) & 1u) xyzz_madd_call(acc, p.x, p.y);
        }
    }
    return acc.is_inf();
}

// ---- kernels -------------------------------------------------------------------------------------------------
// pts: packed device points ((0,0) = infinity); out: n x (compressed ? 1 : 2) x FQS x 48 bytes
template <class F>
struct PointEncode {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *pts, uint32_t compressed, uint32_t *out) {
        constexpr int CW = 12 * Wire<F>::FQS;   // words per coordinate
        Affine<F> p = load_affine<F>(pts, i);
        uint32_t *dst = out + i * (compressed ? CW : 2 * CW);
        bool inf = p.is_inf();
        Wire<F>::put(dst, p.x);                  // infinity: all zero
        if (!compressed) Wire<F>::put(dst + CW, p.y);
        uint32_t flags = (compressed ? 0x80u : 0u) | (inf ? 0x40u : 0u);
        if (compressed && !inf && Wire<F>::lex_largest(p.y)) flags |= 0x20u;
        dst[0] |= flags;                         // byte 0 = low byte of the first (byte-swapped) word
    }
};

// bytes -> packed device points + per-point status (WIRE_*)
template <class F>
struct PointDecode {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t i, const uint32_t *in, uint32_t compressed, uint32_t validate, uint32_t *pts,
                           uint8_t *status) {
        constexpr int CW = 12 * Wire<F>::FQS;
        const uint32_t *src = in + i * (compressed ? CW : 2 * CW);
        uint32_t flags = src[0] & 0xe0u;
        Affine<F> p = Affine<F>::inf();
        uint32_t st = WIRE_OK;
        if (((flags & 0x80u) != 0) != (compressed != 0)) st = WIRE_UNEXPECTED_FLAGS;
        else if (flags & 0x40u) { /* identity */ }
        else if (!Wire<F>::get(src, true, p.x)) st = WIRE_INVALID_DATA;
        else if (compressed) {
            F rhs = F::add(F::mul(F::sqr(p.x), p.x), Wire<F>::curve_b());
            if (!Wire<F>::sqrt(rhs, p.y)) st = WIRE_INVALID_DATA;
            else if (Wire<F>::lex_largest(p.y) != ((flags & 0x20u) != 0)) p.y = F::neg(p.y);
        } else {
            if (!Wire<F>::get(src + CW, false, p.y)) st = WIRE_INVALID_DATA;
            else if (validate && !on_curve(p)) st = WIRE_INVALID_DATA;
        }
        if (st == WIRE_OK && validate && !p.is_inf() && !in_subgroup(p)) st = WIRE_INVALID_DATA;
        if (st != WIRE_OK) p = Affine<F>::inf();
        store_affine_pt<F>(pts, i, p);
        status[i] = (uint8_t)st;
    }
};

}  // namespace g16
