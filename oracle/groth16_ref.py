"""Reference-semantics model of setup / prove / verify -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Line-by-line restatement (in exact big-integer arithmetic, oracle/bls12_381.py) of
    QAP::from_r1cs / compute_quotient_polynomial   /root/reference/crates/groth16-qap/src/lib.rs:95-187,225-271
    CRS::generate_from_qap                         /root/reference/crates/groth16-setup/src/lib.rs:141-268
    Witness::validate, Prover::prove               /root/reference/crates/groth16-core/src/lib.rs:112-131,139-272
    Verifier::verify                               /root/reference/crates/groth16-core/src/lib.rs:308-355
including the reference's quirks that bit-exact parity has to reproduce: every
"F -> Fr" conversion keeps only the low 64-bit limb, h_g1[i] = [s^i/delta]_1 without Z(s), and
`qap.degree()` equals the domain size.  ark-poly's radix-2 domain is restated from its published
algorithm (generator 7, two-adicity 32; third-party crate ark-poly 0.4.2, not vendored).

PARITY UNPINNED by the reference: its tests use thread_rng and pin no proof bytes.
"""
from __future__ import annotations

from bls12_381 import G1, G2, G1_GEN, G2_GEN, R, M64, proof_bytes
import pairing

TWO_ADICITY = 32
TWO_ADIC_ROOT = pow(7, (R - 1) >> TWO_ADICITY, R)


def t64(x):
    """`Fr::from(x.into_bigint().as_ref()[0])`: keep the low 64-bit limb."""
    return (x % R) & M64


def inv(x):
    return pow(x, -1, R)


# ---------------------------------------------------------------- ark-poly radix-2 domain
class Domain:
    def __init__(self, num_coeffs):
        size = 1
        while size < num_coeffs:
            size *= 2
        self.size = size
        log = size.bit_length() - 1
        assert log <= TWO_ADICITY
        g = TWO_ADIC_ROOT
        for _ in range(log, TWO_ADICITY):
            g = g * g % R
        self.group_gen = g

    def ifft(self, evals):
        n = self.size
        evals = list(evals) + [0] * (n - len(evals))
        gi, ni = inv(self.group_gen), inv(n)
        out = []
        for i in range(n):
            w = pow(gi, i, R)
            acc, x = 0, 1
            for j in range(n):
                acc += evals[j] * x
                x = x * w % R
            out.append(acc % R * ni % R)
        return out


def trim(p):
    p = [c % R for c in p]
    while p and p[-1] == 0:
        p.pop()
    return p


def poly_eval(p, x):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % R
    return acc


def poly_add(a, b):
    n = max(len(a), len(b))
    return trim([(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)])


def poly_mul(a, b):
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % R
    return trim(out)


def degree(p):
    return len(p) - 1 if p else 0          # ark: degree of the zero polynomial is 0


# ---------------------------------------------------------------- R1CS -> QAP (qap/src/lib.rs:95-187)
class QAP:
    def __init__(self, constraints, num_variables):
        """constraints: list of (A, B, C) dicts {variable index: coefficient}."""
        self.constraints = constraints
        self.num_constraints = len(constraints)
        self.num_variables = num_variables
        self.domain = Domain(max(1, self.num_constraints))     # next_power_of_two(0) == 1
        self.n = self.domain.size

        def polys(k):
            out = []
            for v in range(num_variables):
                evals = [constraints[i][k].get(v, 0) % R for i in range(self.num_constraints)]
                out.append(trim(self.domain.ifft(evals)))
            return out
        self.a_polys, self.b_polys, self.c_polys = polys(0), polys(1), polys(2)
        self.vanishing = [R - 1] + [0] * (self.n - 1) + [1]

    def degree(self):
        d = max(degree(p) for p in self.a_polys + self.b_polys + self.c_polys)
        return max(d, degree(self.vanishing))

    def evaluate_at(self, point, assignment):
        a = sum(w * poly_eval(p, point) for w, p in zip(assignment, self.a_polys)) % R
        b = sum(w * poly_eval(p, point) for w, p in zip(assignment, self.b_polys)) % R
        c = sum(w * poly_eval(p, point) for w, p in zip(assignment, self.c_polys)) % R
        return a, b, c

    def domain_evals(self, assignment):
        """(a_i, b_i, c_i) = (<A-row i, w>, <B-row i, w>, <C-row i, w>) for the n domain points (zero rows
        beyond the constraints): the evaluations of A(x) = sum_k w_k A_k(x) etc. on the domain, since
        A_k is the interpolant of column k (from_r1cs, qap/src/lib.rs:143-170)."""
        out = []
        for k in range(3):
            ev = [sum(coef * assignment[v] for v, coef in row[k].items()) % R for row in self.constraints]
            out.append(ev + [0] * (self.n - len(ev)))
        return out

    def quotient(self, assignment):
        """compute_quotient_polynomial (qap/src/lib.rs:225-271) incl. ark's divide_by_vanishing_poly."""
        def comb(polys):
            acc = []
            for w, p in zip(assignment, polys):
                if w % R:
                    acc = poly_add(acc, [c * w % R for c in p])
            return acc
        num = poly_add(poly_mul(comb(self.a_polys), comb(self.b_polys)), [(-c) % R for c in comb(self.c_polys)])
        n = self.n
        if len(num) < n:
            q, rem = [], num
        else:
            q = list(num[n:])
            for i in range(1, len(num) // n):
                for k, c in enumerate(num[n * (i + 1):]):
                    q[k] = (q[k] + c) % R
            rem = list(num[:n])
            for k, c in enumerate(q):
                if k < len(rem):
                    rem[k] = (rem[k] + c) % R
            q, rem = trim(q), trim(rem)
        if rem:
            raise ValueError("QAPError::PolynomialDivisionFailed")
        return q


# ---------------------------------------------------------------- setup (setup/src/lib.rs:141-268)
def setup(qap: QAP, params: dict, num_public: int):
    for k in ("alpha", "beta", "gamma", "delta"):
        if params[k] % R == 0:
            raise ValueError("SetupError::InvalidParams")
    if num_public >= qap.num_variables:
        raise ValueError("SetupError::InvalidParams")
    alpha, beta, gamma, delta, s = (t64(params[k]) for k in ("alpha", "beta", "gamma", "delta", "s"))
    g1 = lambda k: G1.mul(G1_GEN, k)
    g2 = lambda k: G2.mul(G2_GEN, k)
    a_vals = [poly_eval(p, s) for p in qap.a_polys]
    b_vals = [poly_eval(p, s) for p in qap.b_polys]
    c_vals = [poly_eval(p, s) for p in qap.c_polys]
    # exponents (kept for tests) and points
    exps = dict(
        a=[t64(v) for v in a_vals], b=[t64(v) for v in b_vals],
        ic=[t64((beta * a_vals[i] + alpha * b_vals[i] + c_vals[i]) * inv(delta)) for i in range(num_public + 1, qap.num_variables)],
        vk_ic=[t64((beta * a_vals[i] + alpha * b_vals[i] + c_vals[i]) * inv(gamma)) for i in range(0, num_public + 1)],
        h=[t64(pow(s, i, R) * inv(delta)) for i in range(qap.degree())],
    )
    pk = dict(
        alpha_g1=g1(params["alpha"]), beta_g1=g1(params["beta"]), beta_g2=g2(params["beta"]),
        delta_g1=g1(params["delta"]), delta_g2=g2(params["delta"]),
        a_g1=[g1(k) for k in exps["a"]], b_g1=[g1(k) for k in exps["b"]], b_g2=[g2(k) for k in exps["b"]],
        ic_g1=[g1(k) for k in exps["ic"]], h_g1=[g1(k) for k in exps["h"]],
        num_public=num_public, qap=qap, exps=exps,
    )
    vk = dict(alpha_g1=pk["alpha_g1"], beta_g2=pk["beta_g2"], gamma_g2=g2(params["gamma"]), delta_g2=pk["delta_g2"],
              ic_g1=[g1(k) for k in exps["vk_ic"]], num_public=num_public)
    return pk, vk


# ---------------------------------------------------------------- prove (core/src/lib.rs:139-272)
def prover_inputs(pk, assignment):
    """The host-side part of Prover::prove that feeds the MSMs: validate, truncate, quotient."""
    qap = pk["qap"]
    if len(assignment) != qap.num_variables:
        raise ValueError("GrothError::InvalidWitness(length)")
    a, b, c = qap.evaluate_at(qap.domain.group_gen, assignment)
    if a * b % R != c:
        raise ValueError("GrothError::InvalidWitness(constraints)")
    assignment_fr = [t64(w) for w in assignment]
    h_coeffs = [t64(c) for c in qap.quotient(assignment)]
    return assignment_fr, h_coeffs


def prove(pk, assignment, r, s):
    """Returns the proof (A in G1, B in G2, C in G1) for fixed randomness (r, s)."""
    w, h = prover_inputs(pk, assignment)
    msm1 = lambda terms: G1.msm_naive([p for _, p in terms], [k for k, _ in terms])
    msm2 = lambda terms: G2.msm_naive([p for _, p in terms], [k for k, _ in terms])
    a_terms = [(1, pk["alpha_g1"])] + [(wi, pk["a_g1"][i]) for i, wi in enumerate(w) if wi and i < len(pk["a_g1"])] + [(r, pk["delta_g1"])]
    pi_a = msm1(a_terms)
    b_terms = [(1, pk["beta_g2"])] + [(wi, pk["b_g2"][i]) for i, wi in enumerate(w) if wi and i < len(pk["b_g2"])] + [(s, pk["delta_g2"])]
    pi_b = msm2(b_terms)
    h_terms = [(c, p) for c, p in zip(h, pk["h_g1"]) if c]
    h_s = msm1(h_terms) if h_terms else None
    c_terms = []
    npub = pk["num_public"]
    for i in range(npub + 1, len(w)):
        if w[i] and (i - npub - 1) < len(pk["ic_g1"]):
            c_terms.append((w[i], pk["ic_g1"][i - npub - 1]))
    if h_s is not None:
        c_terms.append((1, h_s))
    if pi_a is not None:
        c_terms.append((s, pi_a))
    b1_terms = [(1, pk["beta_g1"])] + [(wi, pk["b_g1"][i]) for i, wi in enumerate(w) if wi and i < len(pk["b_g1"])]
    pi_b1 = msm1(b1_terms)
    if pi_b1 is not None:
        c_terms.append((r, pi_b1))
    pi_c = msm1(c_terms) if c_terms else None
    return pi_a, pi_b, pi_c


# ---------------------------------------------------------------- verify (core/src/lib.rs:308-355)
def verify(vk, proof, public_inputs):
    if len(public_inputs) != vk["num_public"]:
        raise ValueError("GrothError::InvalidWitness(public inputs)")
    pub = [t64(x) for x in public_inputs]
    terms = [(1, vk["ic_g1"][0])] + [(x, vk["ic_g1"][i + 1]) for i, x in enumerate(pub) if x]
    ic = G1.msm_naive([p for _, p in terms], [k for k, _ in terms])
    a, b, c = proof
    return pairing.pairing_product_is_one([
        (a, b), (G1.neg(vk["alpha_g1"]), vk["beta_g2"]), (G1.neg(ic), vk["gamma_g2"]), (G1.neg(c), vk["delta_g2"])])


# ---------------------------------------------------------------- BASELINE config 1 circuits
def circuit_mul():
    """x * y = z with w = [1, z=12, x=3, y=4]... exactly the reference's test circuit
    (crates/groth16-core/src/lib.rs:446-471): variables [1, x, y, z], one public input."""
    constraints = [({1: 1}, {2: 1}, {3: 1})]
    return constraints, 4, [1, 3, 4, 12], 1


def circuit_cubic():
    """x^3 + x + 5 = 35 (BASELINE config 1; defined in SURVEY.md 8d): vars [1, out, x, sym1, y, sym2]."""
    constraints = [({2: 1}, {2: 1}, {3: 1}), ({3: 1}, {2: 1}, {4: 1}), ({4: 1, 2: 1}, {0: 1}, {5: 1}), ({5: 1, 0: 5}, {0: 1}, {1: 1})]
    return constraints, 6, [1, 35, 3, 9, 27, 30], 1


P_VERIFY = dict(alpha=11, beta=13, gamma=1, delta=1, s=17)
P_RAND = dict(
    alpha=0x1f2e3d4c5b6a79880123456789abcdef0fedcba98765432100112233445566778 % R,
    beta=0x2a3b4c5d6e7f80910a1b2c3d4e5f60718293a4b5c6d7e8f9012345670abcdef1 % R,
    gamma=0x3141592653589793238462643383279502884197169399375105820974944592 % R,
    delta=0x2718281828459045235360287471352662497757247093699959574966967627 % R,
    s=0x1618033988749894848204586834365638117720309179805762862135448622 % R,
)
FIXED_R = 0x0123456789abcdef0fedcba9876543211032547698badcfe1357924680acebdf % R
FIXED_S = 0x0fedcba987654321123456789abcdef00f1e2d3c4b5a69788796a5b4c3d2e1f0 % R


def proof_to_bytes(proof, compressed=True):
    return proof_bytes(proof[0], proof[1], proof[2], compressed)
