// NOT BUILT.  Measured-and-lost field routines, moved out of csrc/fp.cuh (members of `template <class P> struct Fp`):
//   sqr_sos              dedicated Montgomery squaring: -7 % IMAD.WIDE, +5 % kernel time (profiles/README.md run 7)
//   mul_wide / redc_wide product-then-reduce for lazy reduction in Fq2: -12.5 % IMAD.WIDE, +17 % kernel time
// Kept as a record of what was tried; see git history (round 1) for the versions that compiled and their host checks.
#if 0
    // ---- Montgomery squaring -------------------------------------------------------------------------
    // a^2 = 2 * sum_{i<j} a_i a_j B^(i+j) + sum_i a_i^2 B^(2i): N(N-1)/2 + N wide MADs for the product
    // instead of N^2, then T * R^-1 = redc(T_lo) + T_hi with a product-free CIOS reduction (N^2 wide
    // MADs).  222 instead of 288 IMAD.WIDE for Fq: a squaring costs ~0.78 of a multiplication on the
    // integer-multiply pipe (the extra shifts/adds run on the ALU pipe, which has headroom).
    //
    // MEASURED (B200, profiles/README.md run 7): although it retires 7 % fewer IMAD.WIDE in the bucket
    // accumulation kernel, the longer dependent structure (product -> doubling -> reduction, ~900 extra ALU
    // instructions) made that kernel 5 % SLOWER (72.7 -> 76.7 ms at 2^24), so `sqr` below stays `mul(a, a)`
    // and this routine is kept (and parity-tested in emulation) as `sqr_sos` for future scheduling work.
    G16_HD static Fp sqr(const Fp &a) { return mul(a, a); }
    G16_MUL_HD static Fp sqr_sos(const Fp &a) {
        // 1. cross products, kept in two 64-bit aligned accumulators:
        //    ce[k] holds word k (i + j even), co[k] holds word k + 1 (i + j odd)
        uint32_t ce[2 * N], co[2 * N];
#pragma unroll
        for (int k = 0; k < 2 * N; ++k) { ce[k] = 0; co[k] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; ++i) {
            // odd-sum chain: j = i+1, i+3, ...  -> co index i + j - 1
            {
                int j = i + 1;
                co[i + j - 1] = mad_lo_cc(a.l[i], a.l[j], co[i + j - 1]);
                co[i + j] = madc_hi_cc(a.l[i], a.l[j], co[i + j]);
#pragma unroll
                for (j = i + 3; j < N; j += 2) {
                    co[i + j - 1] = madc_lo_cc(a.l[i], a.l[j], co[i + j - 1]);
                    co[i + j] = madc_hi_cc(a.l[i], a.l[j], co[i + j]);
                }
                // j is now the first index past the chain: the carry lands on the next (small) word
                co[i + j - 1] = addc(co[i + j - 1], 0u);
            }
            // even-sum chain: j = i+2, i+4, ...  -> ce index i + j
            if (i + 2 < N) {
                int j = i + 2;
                ce[i + j] = mad_lo_cc(a.l[i], a.l[j], ce[i + j]);
                ce[i + j + 1] = madc_hi_cc(a.l[i], a.l[j], ce[i + j + 1]);
#pragma unroll
                for (j = i + 4; j < N; j += 2) {
                    ce[i + j] = madc_lo_cc(a.l[i], a.l[j], ce[i + j]);
                    ce[i + j + 1] = madc_hi_cc(a.l[i], a.l[j], ce[i + j + 1]);
                }
                ce[i + j] = addc(ce[i + j], 0u);
            }
        }
        // 2. flat cross sum c = ce + (co << 32), doubled, plus the squares on the even columns
        uint32_t t[2 * N];
        t[0] = ce[0];
        t[1] = add_cc(ce[1], co[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; ++k) t[k] = addc_cc(ce[k], co[k - 1]);
        t[2 * N - 1] = addc(ce[2 * N - 1], co[2 * N - 2]);
#pragma unroll
        for (int k = 2 * N - 1; k > 0; --k) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
        t[0] <<= 1;
        t[0] = mad_lo_cc(a.l[0], a.l[0], t[0]);
        t[1] = madc_hi_cc(a.l[0], a.l[0], t[1]);
#pragma unroll
        for (int i = 1; i < N; ++i) {
            t[2 * i] = madc_lo_cc(a.l[i], a.l[i], t[2 * i]);
            t[2 * i + 1] = madc_hi_cc(a.l[i], a.l[i], t[2 * i + 1]);
        }
        // 3. redc(t_lo): CIOS rounds without a product term (same even/odd accumulator as mul)
        uint32_t ev[N], od[N];
#pragma unroll
        for (int k = 0; k < N; ++k) { ev[k] = t[k]; od[k] = 0; }
        redc_round<true>(ev, od);
        redc_round<false>(od, ev);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            redc_round<false>(ev, od);
            redc_round<false>(od, ev);
        }
        Fp r;
        r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(ev[i], od[i + 1]);
        r.l[N - 1] = addc(ev[N - 1], 0u);
        final_sub(r.l);                       // redc(t_lo) <= p
        // 4. + t_hi (< p), one more conditional subtraction
        r.l[0] = add_cc(r.l[0], t[N]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(r.l[i], t[N + i]);
        r.l[N - 1] = addc(r.l[N - 1], t[2 * N - 1]);
        final_sub(r.l);
        return r;
    }
    // ---- wide product and separate reduction (lazy reduction in Fq2, see fq2.cuh) ---------------------------------
    // t[0 .. 2N) = a * b, no reduction; a, b < 2^(32 N).  Same two 64-bit aligned accumulators as the squaring above:
    // ce[k] holds word k of the partial products with i + j even, co[k] word k + 1 of those with i + j odd.
    G16_HD static void mul_wide(const uint32_t *a, const uint32_t *b, uint32_t *t) {
        uint32_t ce[2 * N], co[2 * N];
#pragma unroll
        for (int k = 0; k < 2 * N; ++k) { ce[k] = 0; co[k] = 0; }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            // j of the parity of i: i + j even -> ce[i + j]
            {
                int j = i & 1;
                ce[i + j] = mad_lo_cc(a[j], b[i], ce[i + j]);
                ce[i + j + 1] = madc_hi_cc(a[j], b[i], ce[i + j + 1]);
#pragma unroll
                for (j += 2; j < N; j += 2) {
                    ce[i + j] = madc_lo_cc(a[j], b[i], ce[i + j]);
                    ce[i + j + 1] = madc_hi_cc(a[j], b[i], ce[i + j + 1]);
                }
                if (i + j < 2 * N) ce[i + j] = addc(ce[i + j], 0u);   // j = first index past the chain
            }
            // j of the other parity: i + j odd -> co[i + j - 1]
            {
                int j = (i & 1) ^ 1;
                co[i + j - 1] = mad_lo_cc(a[j], b[i], co[i + j - 1]);
                co[i + j] = madc_hi_cc(a[j], b[i], co[i + j]);
#pragma unroll
                for (j += 2; j < N; j += 2) {
                    co[i + j - 1] = madc_lo_cc(a[j], b[i], co[i + j - 1]);
                    co[i + j] = madc_hi_cc(a[j], b[i], co[i + j]);
                }
                if (i + j - 1 < 2 * N) co[i + j - 1] = addc(co[i + j - 1], 0u);
            }
        }
        t[0] = ce[0];
        t[1] = add_cc(ce[1], co[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; ++k) t[k] = addc_cc(ce[k], co[k - 1]);
        t[2 * N - 1] = addc(ce[2 * N - 1], co[2 * N - 2]);
    }
    // T * R^-1 mod p for a 2N-word T < p * R, fully reduced: redc of the low half (product-free CIOS rounds) + high half
    G16_HD static Fp redc_wide(const uint32_t *t) {
        uint32_t ev[N], od[N];
#pragma unroll
        for (int k = 0; k < N; ++k) { ev[k] = t[k]; od[k] = 0; }
        redc_round<true>(ev, od);
        redc_round<false>(od, ev);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            redc_round<false>(ev, od);
            redc_round<false>(od, ev);
        }
        Fp r;
        r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(ev[i], od[i + 1]);
        r.l[N - 1] = addc(ev[N - 1], 0u);
        final_sub(r.l);                       // redc(t_lo) <= p
        r.l[0] = add_cc(r.l[0], t[N]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(r.l[i], t[N + i]);
        r.l[N - 1] = addc(r.l[N - 1], t[2 * N - 1]);
        final_sub(r.l);
        return r;
    }
    // x -= y over 2N words; returns the borrow (all ones when x < y)
    G16_HD static uint32_t sub_wide(uint32_t *x, const uint32_t *y) {
        x[0] = sub_cc(x[0], y[0]);
#pragma unroll
        for (int i = 1; i < 2 * N; ++i) x[i] = subc_cc(x[i], y[i]);
        return subc(0u, 0u);
    }
    // a + b without reduction (a, b < p: the sum fits the N words because p has spare top bits)
    G16_HD static void add_noreduce(const uint32_t *a, const uint32_t *b, uint32_t *r) {
        r[0] = add_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r[i] = addc_cc(a[i], b[i]);
        r[N - 1] = addc(a[N - 1], b[N - 1]);
    }
    // one reduction round on value = E + O * 2^32; on exit E[0] == 0 (mod 2^32), caller swaps roles
    template <bool FIRST>
    G16_HD static void redc_round(uint32_t *E, uint32_t *O) {
        if (!FIRST) {
            E[0] = add_cc(E[0], O[1]);
#pragma unroll
            for (int j = 0; j < N - 2; ++j) O[j] = addc_cc(O[j + 2], 0u);
            O[N - 2] = addc(0u, 0u);
            O[N - 1] = 0u;
        }
        uint32_t m = E[0] * P::NINV;
        O[0] = mad_lo_cc(P::MOD(1), m, O[0]);
        O[1] = madc_hi_cc(P::MOD(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            O[j] = madc_lo_cc(P::MOD(j + 1), m, O[j]);
            O[j + 1] = madc_hi_cc(P::MOD(j + 1), m, O[j + 1]);
        }
        E[0] = mad_lo_cc(P::MOD(0), m, E[0]);
        E[1] = madc_hi_cc(P::MOD(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = madc_lo_cc(P::MOD(j), m, E[j]);
            E[j + 1] = madc_hi_cc(P::MOD(j), m, E[j + 1]);
        }
        O[N - 1] = addc(O[N - 1], 0u);
    }

#endif

// Fq2::mul_lazy (member of struct Fq2), same status:
#if 0
    // Karatsuba with lazy reduction: 3 wide products, 2 Montgomery reductions (720 instead of 864 wide MADs).
    //   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 = a0 b1 + a1 b0 < 2 q^2 < q R
    //   c0 = a0 b0 - a1 b1 (+ q R when negative)                  < q R
    G16_HD static Fq2 mul_lazy(const Fq2 &a, const Fq2 &b) {
        constexpr int N = Fq::N;
        uint32_t v0[2 * N], v1[2 * N], sw[2 * N], sa[N], sb[N];
        Fq::mul_wide(a.c0.l, b.c0.l, v0);
        Fq::mul_wide(a.c1.l, b.c1.l, v1);
        Fq::add_noreduce(a.c0.l, a.c1.l, sa);
        Fq::add_noreduce(b.c0.l, b.c1.l, sb);
        Fq::mul_wide(sa, sb, sw);
        Fq::sub_wide(sw, v0);
        Fq::sub_wide(sw, v1);
        uint32_t borrow = Fq::sub_wide(v0, v1);
        v0[N] = add_cc(v0[N], FqParams::MOD(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) v0[N + i] = addc_cc(v0[N + i], FqParams::MOD(i) & borrow);
        v0[2 * N - 1] = addc(v0[2 * N - 1], FqParams::MOD(N - 1) & borrow);
        return Fq2{Fq::redc_wide(v0), Fq::redc_wide(sw)};
    }
#endif
