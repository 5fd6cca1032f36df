/* CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Jacobian short-Weierstrass (a = 0) group law + the ark-ec 0.4.2 variable-base MSM,
 * instantiated twice by cpu_msm.c (FE = fq_t -> G1, FE = fq2_t -> G2).
 *
 * Restates the published algorithm of ark-ec 0.4.2 (third-party, not vendored; pinned in
 * /root/reference/Cargo.lock:101-102) that the reference calls at
 * crates/groth16-core/src/lib.rs:282,296 (`VariableBaseMSM::msm` -> msm_bigint_wnaf),
 * :285,299 (`into_affine`) and crates/groth16-setup/src/lib.rs:166-241 (`Projective * Fr`,
 * double-and-add).  SURVEY.md App. A lists the algorithm this follows.
 *
 * Required macros: FE (field element type), F(name) (field fn), C(name) (curve fn prefix).
 */

typedef struct { FE x, y; int inf; } C(aff_t);
typedef struct { FE x, y, z; } C(jac_t);

static inline void C(jac_set_zero)(C(jac_t) *p) { F(set_one)(&p->x); F(set_one)(&p->y); F(set_zero)(&p->z); }
static inline int C(jac_is_zero)(const C(jac_t) *p) { return F(is_zero)(&p->z); }

/* dbl-2009-l, a = 0 (ark `double_in_place`) */
static inline void C(jac_double)(C(jac_t) *p) {
    if (C(jac_is_zero)(p)) return;
    FE a, b, c, d, e, f, t;
    F(sqr)(&a, &p->x); F(sqr)(&b, &p->y); F(sqr)(&c, &b);
    F(add)(&t, &p->x, &b); F(sqr)(&t, &t); F(sub)(&t, &t, &a); F(sub)(&t, &t, &c); F(dbl)(&d, &t);
    F(dbl)(&e, &a); F(add)(&e, &e, &a);
    F(sqr)(&f, &e);
    F(mul)(&t, &p->y, &p->z); F(dbl)(&p->z, &t);
    F(dbl)(&t, &d); F(sub)(&p->x, &f, &t);
    F(sub)(&t, &d, &p->x); F(mul)(&t, &e, &t);
    F(dbl)(&c, &c); F(dbl)(&c, &c); F(dbl)(&c, &c);
    F(sub)(&p->y, &t, &c);
}

/* madd-2007-bl (ark `AddAssign<&Affine>`): no-op for an infinity base, P+P -> double */
static inline void C(jac_madd)(C(jac_t) *p, const C(aff_t) *q) {
    if (q->inf) return;
    if (C(jac_is_zero)(p)) { p->x = q->x; p->y = q->y; F(set_one)(&p->z); return; }
    FE z1z1, u2, s2, h, hh, i, j, r, v, t;
    F(sqr)(&z1z1, &p->z);
    F(mul)(&u2, &q->x, &z1z1);
    F(mul)(&s2, &q->y, &p->z); F(mul)(&s2, &s2, &z1z1);
    if (F(eq)(&p->x, &u2) && F(eq)(&p->y, &s2)) { C(jac_double)(p); return; }
    F(sub)(&h, &u2, &p->x);
    F(sqr)(&hh, &h);
    F(dbl)(&i, &hh); F(dbl)(&i, &i);
    F(mul)(&j, &h, &i);
    F(sub)(&r, &s2, &p->y); F(dbl)(&r, &r);
    F(mul)(&v, &p->x, &i);
    /* Z3 = (Z1+H)^2 - Z1Z1 - HH */
    F(add)(&t, &p->z, &h); F(sqr)(&t, &t); F(sub)(&t, &t, &z1z1); F(sub)(&p->z, &t, &hh);
    /* X3 = r^2 - J - 2V */
    F(sqr)(&t, &r); F(sub)(&t, &t, &j); F(sub)(&t, &t, &v); F(sub)(&p->x, &t, &v);
    /* Y3 = r (V - X3) - 2 Y1 J */
    F(mul)(&j, &p->y, &j); F(dbl)(&j, &j);
    F(sub)(&t, &v, &p->x); F(mul)(&t, &r, &t); F(sub)(&p->y, &t, &j);
}
static inline void C(jac_msub)(C(jac_t) *p, const C(aff_t) *q) {
    C(aff_t) n = *q; F(neg)(&n.y, &q->y); C(jac_madd)(p, &n);
}

/* add-2007-bl (ark `AddAssign<&Projective>`) */
static inline void C(jac_add)(C(jac_t) *p, const C(jac_t) *q) {
    if (C(jac_is_zero)(p)) { *p = *q; return; }
    if (C(jac_is_zero)(q)) return;
    FE z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    F(sqr)(&z1z1, &p->z); F(sqr)(&z2z2, &q->z);
    F(mul)(&u1, &p->x, &z2z2); F(mul)(&u2, &q->x, &z1z1);
    F(mul)(&s1, &p->y, &q->z); F(mul)(&s1, &s1, &z2z2);
    F(mul)(&s2, &q->y, &p->z); F(mul)(&s2, &s2, &z1z1);
    if (F(eq)(&u1, &u2) && F(eq)(&s1, &s2)) { C(jac_double)(p); return; }
    F(sub)(&h, &u2, &u1);
    F(dbl)(&i, &h); F(sqr)(&i, &i);
    F(mul)(&j, &h, &i);
    F(sub)(&r, &s2, &s1); F(dbl)(&r, &r);
    F(mul)(&v, &u1, &i);
    /* Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) H */
    F(add)(&t, &p->z, &q->z); F(sqr)(&t, &t); F(sub)(&t, &t, &z1z1); F(sub)(&t, &t, &z2z2); F(mul)(&p->z, &t, &h);
    F(sqr)(&t, &r); F(sub)(&t, &t, &j); F(sub)(&t, &t, &v); F(sub)(&p->x, &t, &v);
    F(mul)(&s1, &s1, &j); F(dbl)(&s1, &s1);
    F(sub)(&t, &v, &p->x); F(mul)(&t, &r, &t); F(sub)(&p->y, &t, &s1);
}

/* `into_affine`: z = 0 -> {0, 0, infinity}; else one inversion */
static inline void C(jac_to_affine)(C(aff_t) *r, const C(jac_t) *p) {
    if (C(jac_is_zero)(p)) { F(set_zero)(&r->x); F(set_zero)(&r->y); r->inf = 1; return; }
    FE zi, zi2;
    F(inv)(&zi, &p->z); F(sqr)(&zi2, &zi);
    F(mul)(&r->x, &p->x, &zi2);
    F(mul)(&zi2, &zi2, &zi); F(mul)(&r->y, &p->y, &zi2);
    r->inf = 0;
}

/* `Projective * Fr` = mul_bigint: MSB-first double-and-add without leading zeros */
static inline void C(mul_bigint)(C(jac_t) *out, const C(aff_t) *base, const uint64_t k[4]) {
    C(jac_set_zero)(out);
    int top = 255;
    while (top >= 0 && !((k[top >> 6] >> (top & 63)) & 1)) --top;
    for (int i = top; i >= 0; --i) {
        C(jac_double)(out);
        if ((k[i >> 6] >> (i & 63)) & 1) C(jac_madd)(out, base);
    }
}

/* ark-ec 0.4.2 msm_bigint_wnaf; `threads` > 1 processes windows in
 * parallel, which is what ark's `parallel` feature would do (the reference builds without it). */
static int C(msm_bigint_wnaf)(C(jac_t) *result, const C(aff_t) *bases, const uint64_t *bigints /* n x 4 */,
                             size_t size, int threads) {
    C(jac_set_zero)(result);
    if (size == 0) return 0;
    int c = size < 32 ? 3 : (int)ln_without_floats(size) + 2;
    const int num_bits = 255;
    int digits_count = (num_bits + c - 1) / c;
    int64_t *digits = (int64_t *)malloc(sizeof(int64_t) * size * (size_t)digits_count);
    if (!digits) return -1;
    for (size_t i = 0; i < size; ++i) make_digits(digits + i * (size_t)digits_count, bigints + 4 * i, c, num_bits);
    C(jac_t) *window_sums = (C(jac_t) *)malloc(sizeof(C(jac_t)) * (size_t)digits_count);
    int failed = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (int w = 0; w < digits_count; ++w) {
        size_t nb = (size_t)1 << c;
        C(jac_t) *buckets = (C(jac_t) *)malloc(sizeof(C(jac_t)) * nb);
        if (!buckets) { failed = 1; continue; }
        for (size_t b = 0; b < nb; ++b) C(jac_set_zero)(&buckets[b]);
        for (size_t i = 0; i < size; ++i) {
            int64_t d = digits[i * (size_t)digits_count + (size_t)w];
            if (d > 0) C(jac_madd)(&buckets[d - 1], &bases[i]);
            else if (d < 0) C(jac_msub)(&buckets[-d - 1], &bases[i]);
        }
        C(jac_t) running, res;
        C(jac_set_zero)(&running); C(jac_set_zero)(&res);
        for (size_t b = nb; b-- > 0;) { C(jac_add)(&running, &buckets[b]); C(jac_add)(&res, &running); }
        window_sums[w] = res;
        free(buckets);
    }
    if (!failed) {
        C(jac_t) total; C(jac_set_zero)(&total);
        for (int w = digits_count - 1; w >= 1; --w) {
            C(jac_add)(&total, &window_sums[w]);
            for (int k = 0; k < c; ++k) C(jac_double)(&total);
        }
        *result = window_sums[0];
        C(jac_add)(result, &total);
    }
    free(window_sums); free(digits);
    return failed ? -1 : 0;
}
