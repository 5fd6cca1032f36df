"""Parity of the GPU quotient polynomial (g16_quotient_h) with the reference-semantics model."""
import numpy as np

import bls12_381 as bls
import groth16_ref as ref


def fr_arr(vals):
    return np.array([bls.fr_to_mont(v) for v in vals], dtype=np.uint64).reshape(-1, 4)


def fr_list(arr):
    return [bls.fr_from_mont(list(x)) for x in arr]


def check_toy_circuits(ctx):
    """H from the domain evaluations == QAP::compute_quotient_polynomial of the model (un-truncated)."""
    for circuit in (ref.circuit_mul, ref.circuit_cubic):
        constraints, nvars, w, npub = circuit()
        qap = ref.QAP(constraints, nvars)
        a, b, c = qap.domain_evals(w)
        h = fr_list(ctx.quotient_h(fr_arr(a), fr_arr(b), fr_arr(c)))
        exp = qap.quotient(w)
        assert ref.trim(h) == exp, circuit.__name__
        assert len(h) == qap.n


def check_random_polynomials(ctx, log_sizes=(0, 1, 2, 3, 5, 8)):
    """Random A, B of degree < n; C := low(A*B) + high(A*B) makes A*B - C = high(A*B) * (x^n - 1) exactly."""
    import groth16_cuda
    rng = bls.SplitMix64(0x517)
    for log_n in log_sizes:
        n = 1 << log_n
        dom = ref.Domain(n)
        A = [bls.random_fr(rng) for _ in range(n)]
        B = [bls.random_fr(rng) for _ in range(n)]
        AB = [0] * (2 * n - 1) if n > 0 else []
        for i, x in enumerate(A):
            for j, y in enumerate(B):
                AB[i + j] = (AB[i + j] + x * y) % bls.R
        H = AB[n:] + [0] * (n - len(AB[n:]))
        C = [(AB[i] + H[i]) % bls.R for i in range(n)]
        pts = [pow(dom.group_gen, i, bls.R) for i in range(n)]
        ev = lambda p: [ref.poly_eval(p, x) for x in pts]
        h = fr_list(ctx.quotient_h(fr_arr(ev(A)), fr_arr(ev(B)), fr_arr(ev(C))))
        assert h == H, log_n
        if n >= 2:   # an unsatisfied constraint must be reported like the reference's PolynomialDivisionFailed
            bad = ev(C)
            bad[1] = (bad[1] + 1) % bls.R
            try:
                ctx.quotient_h(fr_arr(ev(A)), fr_arr(ev(B)), fr_arr(bad))
            except groth16_cuda.MSMError as e:
                assert "division failed" in str(e)
            else:
                raise AssertionError("non-vanishing A*B - C must be rejected")
