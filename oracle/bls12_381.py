"""Exact big-integer model of BLS12-381 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The arithmetic that the reference's hot path runs lives in third-party crates that
are NOT vendored under /root/reference (Cargo.lock pins: ark-ec 0.4.2, ark-ff 0.4.2,
ark-bls12-381 0.4.0, ark-serialize 0.4.2, ark-poly 0.4.2).  The reference's own call
sites are crates/groth16-core/src/lib.rs:282,296 (VariableBaseMSM::msm),
:285,299 (into_affine) and crates/groth16-setup/src/lib.rs:166-171,189,198,205,216,
227,239 (Projective * Fr).  This file restates the *published mathematics* of the
curve (an MSM result is a well defined group element, so any exact implementation
gives the bit-exact canonical affine answer) and ark's in-memory representation
(Montgomery form, little-endian u64 limbs).

PARITY UNPINNED by the reference: it has no golden vector, fixed seed or serialised
proof for this path.  The pins used instead are the public
BLS12-381 known answers checked in tests/test_oracle_kat.py (generator encodings,
2G, r*G = O, Montgomery constants) and three-way agreement between this model, the
C restatement in oracle/cpu_msm.c and the CUDA engine.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------- constants
Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# |x| of the BLS parameter (x is negative): used by the pairing only
BLS_X = 0xD201000000010000

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    (
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)

FQ_LIMBS64 = 6
FR_LIMBS64 = 4
FQ_R = (1 << 384) % Q          # Montgomery radix for Fq (ark-ff MontBackend, 6 x u64)
FR_R = (1 << 256) % R          # Montgomery radix for Fr (4 x u64)
FQ_RINV = pow(FQ_R, -1, Q)
FR_RINV = pow(FR_R, -1, R)
FQ_R2 = FQ_R * FQ_R % Q
FR_R2 = FR_R * FR_R % R
FQ_NINV64 = (-pow(Q, -1, 1 << 64)) % (1 << 64)
FR_NINV64 = (-pow(R, -1, 1 << 64)) % (1 << 64)
FQ_NINV32 = FQ_NINV64 & 0xFFFFFFFF
FR_NINV32 = FR_NINV64 & 0xFFFFFFFF


# ----------------------------------------------------------------------------- field ops
class FqOps:
    """Fq as python ints in [0, Q)."""
    zero = 0
    one = 1
    @staticmethod
    def add(a, b): return (a + b) % Q
    @staticmethod
    def sub(a, b): return (a - b) % Q
    @staticmethod
    def neg(a): return (-a) % Q
    @staticmethod
    def mul(a, b): return a * b % Q
    @staticmethod
    def sqr(a): return a * a % Q
    @staticmethod
    def inv(a): return pow(a, -1, Q)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def small(k): return k % Q


class Fq2Ops:
    """Fq2 = Fq[u]/(u^2+1) as tuples (c0, c1)."""
    zero = (0, 0)
    one = (1, 0)
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
    @staticmethod
    def neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
    @staticmethod
    def mul(a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
    @staticmethod
    def sqr(a):
        return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)
    @staticmethod
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return (a[0] * d % Q, (-a[1]) * d % Q)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def small(k): return (k % Q, 0)


class Curve:
    """Short Weierstrass y^2 = x^3 + b over field F (a = 0).  Affine points are (x, y)
    tuples or None for the identity; Jacobian points are (X, Y, Z) with Z = 0 identity."""
    def __init__(self, F, b, gen, name):
        self.F, self.b, self.gen, self.name = F, b, gen, name

    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    # -- affine (slow, obviously correct)
    def add(self, P, Q_):
        F = self.F
        if P is None: return Q_
        if Q_ is None: return P
        x1, y1 = P; x2, y2 = Q_
        if x1 == x2:
            if y1 == y2 and not F.is_zero(y1):
                lam = F.mul(F.mul(F.small(3), F.sqr(x1)), F.inv(F.add(y1, y1)))
            else:
                return None
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # -- Jacobian
    def to_jac(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def jac_is_zero(self, P):
        return self.F.is_zero(P[2])

    def jac_double(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        t = F.sub(F.sub(F.sqr(F.add(X, B)), A), C)
        D = F.add(t, t)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def jac_add(self, P, Q_):
        F = self.F
        if F.is_zero(P[2]): return Q_
        if F.is_zero(Q_[2]): return P
        X1, Y1, Z1 = P; X2, Y2, Z2 = Q_
        Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2); S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jac_double(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1); Rr = F.sub(S2, S1)
        HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jac_to_affine(self, P):
        F = self.F
        if F.is_zero(P[2]):
            return None
        zi = F.inv(P[2]); zi2 = F.sqr(zi)
        return (F.mul(P[0], zi2), F.mul(P[1], F.mul(zi2, zi)))

    def mul(self, P, k):
        """k*P for any integer k (reduced mod R: all points used here lie in the r-torsion)."""
        k %= R
        if P is None or k == 0:
            return None
        acc = (self.F.one, self.F.one, self.F.zero)
        base = self.to_jac(P)
        for bit in bin(k)[2:]:
            acc = self.jac_double(acc)
            if bit == '1':
                acc = self.jac_add(acc, base)
        return self.jac_to_affine(acc)

    def msm_naive(self, points, scalars):
        """sum s_i * P_i, exactly what crates/groth16-core/src/lib.rs:275-300 returns
        (affine; identity for empty input, lib.rs:276-278)."""
        if len(points) != len(scalars):
            raise ValueError("length mismatch")
        acc = (self.F.one, self.F.one, self.F.zero)
        for P, s in zip(points, scalars):
            T = self.mul(P, s)
            if T is not None:
                acc = self.jac_add(acc, self.to_jac(T))
        return self.jac_to_affine(acc)

    def msm_pippenger(self, points, scalars, c=None):
        """Bucket method in Python (unsigned windows) for mid-size cross checks."""
        n = len(points)
        if len(scalars) != n:
            raise ValueError("length mismatch")
        if n == 0:
            return None
        if c is None:
            c = max(2, min(16, n.bit_length() - 2))
        F = self.F
        zero = (F.one, F.one, F.zero)
        nwin = (255 + c - 1) // c
        total = zero
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jac_double(total)
            buckets = {}
            for P, s in zip(points, scalars):
                if P is None:
                    continue
                d = ((s % R) >> (w * c)) & ((1 << c) - 1)
                if d:
                    b = buckets.get(d)
                    buckets[d] = self.to_jac(P) if b is None else self.jac_add(b, self.to_jac(P))
            running = zero; acc = zero
            if buckets:
                for d in range(max(buckets), 0, -1):
                    b = buckets.get(d)
                    if b is not None:
                        running = self.jac_add(running, b)
                    acc = self.jac_add(acc, running)
            total = self.jac_add(total, acc)
        return self.jac_to_affine(total)


G1 = Curve(FqOps, 4, G1_GEN, "G1")
G2 = Curve(Fq2Ops, (4, 4), G2_GEN, "G2")


# ----------------------------------------------------------------------------- ark in-memory layouts
def int_to_limbs64(x, n):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs64_to_int(limbs):
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def fq_to_mont(x):
    """canonical int -> ark Fp384 in-memory limbs (Montgomery, LE u64 x 6)."""
    return int_to_limbs64(x * FQ_R % Q, FQ_LIMBS64)


def fq_from_mont(limbs):
    return limbs64_to_int(limbs) * FQ_RINV % Q


def fr_to_mont(x):
    return int_to_limbs64((x % R) * FR_R % R, FR_LIMBS64)


def fr_from_mont(limbs):
    return limbs64_to_int(limbs) * FR_RINV % R


def g1_to_mont(P):
    """-> (12 u64 limbs x||y, infinity flag).  Identity is {x:0, y:0, infinity:true}."""
    if P is None:
        return [0] * 12, 1
    return fq_to_mont(P[0]) + fq_to_mont(P[1]), 0


def g1_from_mont(limbs, inf):
    if inf:
        return None
    return (fq_from_mont(limbs[0:6]), fq_from_mont(limbs[6:12]))


def g2_to_mont(P):
    """-> (24 u64 limbs x.c0||x.c1||y.c0||y.c1, infinity flag)."""
    if P is None:
        return [0] * 24, 1
    (x0, x1), (y0, y1) = P
    return fq_to_mont(x0) + fq_to_mont(x1) + fq_to_mont(y0) + fq_to_mont(y1), 0


def g2_from_mont(limbs, inf):
    if inf:
        return None
    f = [fq_from_mont(limbs[6 * i:6 * i + 6]) for i in range(4)]
    return ((f[0], f[1]), (f[2], f[3]))


# ----------------------------------------------------------------------------- Zcash / IETF encoding
# ark-bls12-381 0.4.0 serialises G1/G2 with this format; Proof
# (crates/groth16-core/src/lib.rs:27-36, derive CanonicalSerialize) = a || b || c.
HALF_Q = (Q - 1) // 2


def _fq_lex_largest(y):
    return y > HALF_Q


def _fq2_lex_largest(y):
    return _fq_lex_largest(y[1]) if y[1] != 0 else _fq_lex_largest(y[0])


def g1_compress(P):
    if P is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(P[0].to_bytes(48, "big"))
    b[0] |= 0x80
    if _fq_lex_largest(P[1]):
        b[0] |= 0x20
    return bytes(b)


def g1_uncompressed(P):
    if P is None:
        return bytes([0x40]) + bytes(95)
    return P[0].to_bytes(48, "big") + P[1].to_bytes(48, "big")


def g2_compress(P):
    if P is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = P
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80
    if _fq2_lex_largest(y):
        b[0] |= 0x20
    return bytes(b)


def g2_uncompressed(P):
    if P is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = P
    return b"".join(v.to_bytes(48, "big") for v in (x1, x0, y1, y0))


def fq_sqrt(a):
    # q = 3 mod 4
    s = pow(a, (Q + 1) // 4, Q)
    return s if s * s % Q == a % Q else None


def g1_decompress(b):
    flags = b[0] >> 5
    if flags & 2:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    y = fq_sqrt((x * x * x + 4) % Q)
    if y is None:
        raise ValueError("not on curve")
    if _fq_lex_largest(y) != bool(flags & 1):
        y = Q - y
    return (x, y)


def fq2_sqrt(a):
    """a root of a in Fq[u]/(u^2+1) or None (complex method, q = 3 mod 4)."""
    a0, a1 = a[0] % Q, a[1] % Q
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fq_sqrt((-a0) % Q)
        return None if s is None else (0, s)
    alpha = fq_sqrt((a0 * a0 + a1 * a1) % Q)
    if alpha is None:
        return None
    half = pow(2, -1, Q)
    for delta in ((a0 + alpha) * half % Q, (a0 - alpha) * half % Q):
        x0 = fq_sqrt(delta)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, Q) % Q
            if ((x0 * x0 - x1 * x1) % Q, 2 * x0 * x1 % Q) == (a0, a1):
                return (x0, x1)
    return None


class WireError(ValueError):
    """ark_serialize::SerializationError: .kind is "InvalidData" or "UnexpectedFlags"."""
    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


def _read_fq(b, mask):
    v = int.from_bytes(b, "big")
    if mask:
        v &= (1 << 381) - 1
    if v >= Q:
        raise WireError("InvalidData")
    return v


def _in_subgroup(curve, P):
    """r * P == O by plain double-and-add (Curve.mul reduces the scalar mod r, which would beg the question)."""
    acc = (curve.F.one, curve.F.one, curve.F.zero)
    base = curve.to_jac(P)
    for bit in bin(R)[2:]:
        acc = curve.jac_double(acc)
        if bit == '1':
            acc = curve.jac_add(acc, base)
    return curve.jac_is_zero(acc)


def point_deserialize(group, b, compressed=True, validate=True):
    """ark-bls12-381 0.4.0 read_g{1,2}_{compressed,uncompressed} + deserialize_with_mode [ark-memory]:
    compression flag must match, infinity flag -> identity, coordinates canonical, y from the curve equation
    with the sign flag; validate = Validate::Yes adds the subgroup check (and, here, the curve equation for
    uncompressed input).  Returns the affine point or None (identity)."""
    k = 1 if group == "g1" else 2
    flags = b[0] >> 5
    if bool(flags & 4) != bool(compressed):
        raise WireError("UnexpectedFlags")
    if flags & 2:
        return None
    curve = G1 if group == "g1" else G2
    if group == "g1":
        x = _read_fq(b[0:48], True)
        rhs = (x * x * x + 4) % Q
    else:
        x1 = _read_fq(b[0:48], True)
        x0 = _read_fq(b[48:96], False)
        x = (x0, x1)
        xx = Fq2Ops.mul(Fq2Ops.mul(x, x), x)
        rhs = ((xx[0] + 4) % Q, (xx[1] + 4) % Q)
    if compressed:
        y = fq_sqrt(rhs) if group == "g1" else fq2_sqrt(rhs)
        if y is None:
            raise WireError("InvalidData")
        largest = _fq_lex_largest(y) if group == "g1" else _fq2_lex_largest(y)
        if largest != bool(flags & 1):
            y = (Q - y) % Q if group == "g1" else ((Q - y[0]) % Q, (Q - y[1]) % Q)
    else:
        o = 48 * k
        if group == "g1":
            y = _read_fq(b[o:o + 48], False)
            ok = y * y % Q == rhs
        else:
            y1 = _read_fq(b[o:o + 48], False)
            y0 = _read_fq(b[o + 48:o + 96], False)
            y = (y0, y1)
            ok = Fq2Ops.mul(y, y) == rhs
        if validate and not ok:
            raise WireError("InvalidData")
    P = (x, y)
    if validate and not _in_subgroup(curve, P):
        raise WireError("InvalidData")
    return P


def proof_bytes(a, b, c, compressed=True):
    if compressed:
        return g1_compress(a) + g2_compress(b) + g1_compress(c)
    return g1_uncompressed(a) + g2_uncompressed(b) + g1_uncompressed(c)


# ----------------------------------------------------------------------------- deterministic inputs
M64 = (1 << 64) - 1


class SplitMix64:
    """Counter based PRNG shared by oracle, C baseline and the CUDA input generators
    (scalars seed 0x5eed0000+log2N, points seed 0xba5e0000+log2N)."""
    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)


def random_fr(rng, bits=255):
    """uniform in [0, r) (bits=255) or in [0, 2^bits) by rejection on 4 limbs."""
    while True:
        v = 0
        for i in range(4):
            v |= rng.next() << (64 * i)
        v &= (1 << bits) - 1
        if v < R:
            return v


def fr_rand_from_limbs(limbs):
    """ark Fr::rand as used for r, s at crates/groth16-core/src/lib.rs:152-153: four
    next_u64 limbs, top limb masked to 255 bits, rejected if >= r, and the limbs ARE the
    Montgomery representation ([ark-memory])."""
    v = limbs64_to_int(limbs) & ((1 << 255) - 1)
    if v >= R:
        return None
    return v * FR_RINV % R
