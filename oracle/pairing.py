"""BLS12-381 ate pairing in plain Python -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used only by oracle/groth16_ref.py to restate `Bls12_381::multi_pairing(..).is_zero()` of
/root/reference/crates/groth16-core/src/lib.rs:352,429 (the verifier stays on the CPU; SURVEY.md 2
marks it out of scope for the engine, the oracle needs it to say whether a proof verifies).

Fq12 is represented as Fq[w]/(w^12 - 2 w^6 + 2)  (w^6 = 1 + u, u^2 = -1), the textbook construction;
the Miller loop runs over |x| = 0xd201000000010000 with affine line functions on the untwisted point.
The final exponent is (q^12 - 1)/r, so `pairing_product_is_one` is exactly "the product of pairings is
the identity of GT", whatever sign convention the Miller loop uses.
"""
from bls12_381 import Q, R, BLS_X

DEG = 12
# w^12 = 2 w^6 - 2
_MOD_TAIL = {0: -2, 6: 2}


def f12(coeffs):
    c = list(coeffs) + [0] * (DEG - len(coeffs))
    return [x % Q for x in c]


ONE = f12([1])
ZERO = f12([0])


def add(a, b): return [(x + y) % Q for x, y in zip(a, b)]
def sub(a, b): return [(x - y) % Q for x, y in zip(a, b)]
def neg(a): return [(-x) % Q for x in a]
def scal(a, k): return [x * k % Q for x in a]


def mul(a, b):
    t = [0] * (2 * DEG - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(2 * DEG - 2, DEG - 1, -1):
        v = t[k]
        if v:
            t[k - 6] += 2 * v      # w^k = w^(k-12) * (2 w^6 - 2)
            t[k - 12] -= 2 * v
    return [x % Q for x in t[:DEG]]


def _deg(p):
    d = len(p) - 1
    while d > 0 and p[d] == 0:
        d -= 1
    return d


def _poly_div(a, b):
    """quotient of polynomial division over Fq"""
    a = list(a)
    db = _deg(b)
    out = [0] * len(a)
    inv_lead = pow(b[db], -1, Q)
    for i in range(_deg(a) - db, -1, -1):
        c = a[db + i] * inv_lead % Q
        out[i] = c
        if c:
            for j in range(db + 1):
                a[i + j] = (a[i + j] - c * b[j]) % Q
    return out[:_deg(out) + 1]


def inv(a):
    """extended Euclid in Fq[w] against the modulus polynomial"""
    lm, hm = [1] + [0] * DEG, [0] * (DEG + 1)
    low = list(a) + [0]
    high = [(-_MOD_TAIL.get(i, 0)) % Q for i in range(DEG)] + [1]   # w^12 - 2 w^6 + 2
    while _deg(low):
        r = _poly_div(high, low)
        r += [0] * (DEG + 1 - len(r))
        nm, new = list(hm), list(high)
        for i in range(DEG + 1):
            for j in range(DEG + 1 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                new[i + j] = (new[i + j] - low[i] * r[j]) % Q
        lm, low, hm, high = nm, new, lm, low
    k = pow(low[0], -1, Q)
    return [x * k % Q for x in lm[:DEG]]


def power(a, e):
    out = ONE
    for bit in bin(e)[2:]:
        out = mul(out, out)
        if bit == '1':
            out = mul(out, a)
    return out


W = f12([0, 1])
W2_INV = inv(mul(W, W))
W3_INV = inv(mul(mul(W, W), W))


def embed_fq2(a):
    c0, c1 = a
    return f12([(c0 - c1) % Q, 0, 0, 0, 0, 0, c1])     # u = w^6 - 1


def untwist(Qp):
    """E'(Fq2) -> E(Fq12)"""
    return (mul(embed_fq2(Qp[0]), W2_INV), mul(embed_fq2(Qp[1]), W3_INV))


def _line(P1, P2, T):
    """line through P1, P2 (or tangent) evaluated at T; all in E(Fq12) affine"""
    x1, y1 = P1; x2, y2 = P2; xt, yt = T
    if x1 != x2:
        m = mul(sub(y2, y1), inv(sub(x2, x1)))
        return sub(mul(m, sub(xt, x1)), sub(yt, y1))
    if y1 == y2:
        m = mul(scal(mul(x1, x1), 3), inv(scal(y1, 2)))
        return sub(mul(m, sub(xt, x1)), sub(yt, y1))
    return sub(xt, x1)


def _dbl(P):
    x, y = P
    m = mul(scal(mul(x, x), 3), inv(scal(y, 2)))
    nx = sub(mul(m, m), scal(x, 2))
    return (nx, sub(mul(m, sub(x, nx)), y))


def _add(P1, P2):
    x1, y1 = P1; x2, y2 = P2
    if x1 == x2:
        return _dbl(P1) if y1 == y2 else None
    m = mul(sub(y2, y1), inv(sub(x2, x1)))
    nx = sub(sub(mul(m, m), x1), x2)
    return (nx, sub(mul(m, sub(x1, nx)), y1))


def miller_loop(P, Qp):
    """P in G1 (affine ints) or None, Qp in G2 (affine Fq2 tuples) or None"""
    if P is None or Qp is None:
        return ONE
    T = (f12([P[0]]), f12([P[1]]))
    Qe = untwist(Qp)
    Rp = Qe
    f = ONE
    for bit in bin(BLS_X)[3:]:
        f = mul(mul(f, f), _line(Rp, Rp, T))
        Rp = _dbl(Rp)
        if bit == '1':
            f = mul(f, _line(Rp, Qe, T))
            Rp = _add(Rp, Qe)
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 in GT  (`multi_pairing(..).is_zero()` in ark's additive notation)"""
    f = ONE
    for P, Qp in pairs:
        f = mul(f, miller_loop(P, Qp))
    return power(f, FINAL_EXP) == ONE
