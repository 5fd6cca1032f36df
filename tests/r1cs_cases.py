"""Parity of the sparse R1CS path (g16_r1cs_*, g16_setup_crs, g16_prove_r1cs) with the reference-semantics
model (oracle/groth16_ref.py: dense QAP::from_r1cs, CRS::generate_from_qap, Prover::prove) and, for sizes
the dense model cannot reach, with a direct big-integer evaluation of the same sums."""
import numpy as np

import bls12_381 as bls
import groth16_ref as ref
import prove_cases

R = bls.R


def fr_arr(vals):
    return np.array([bls.fr_to_mont(v % R) for v in vals], dtype=np.uint64).reshape(-1, 4)


def fr_list(arr):
    return [bls.fr_from_mont(list(x)) for x in arr]


def to_csr(constraints, k):
    """list of (A, B, C) dicts {variable: coefficient} -> CSR arrays of matrix k."""
    row_ptr, col, val = [0], [], []
    for row in constraints:
        for v, coef in sorted(row[k].items()):
            col.append(v)
            val.append(coef % R)
        row_ptr.append(len(col))
    return (np.array(row_ptr, dtype=np.uint32), np.array(col, dtype=np.uint32),
            fr_arr(val) if val else np.zeros((0, 4), dtype=np.uint64))


def upload(ctx, constraints, nvars):
    return ctx.r1cs_upload(len(constraints), nvars, to_csr(constraints, 0), to_csr(constraints, 1), to_csr(constraints, 2))


def chain_circuit(seed, m, n_public=1, width=3, bits=255):
    """Satisfiable synthetic R1CS: variables [1, public..., seeds..., products...]; constraint i multiplies two random
    sparse combinations of earlier variables and defines a new variable as the product."""
    rng = bls.SplitMix64(seed)
    w = [1] + [bls.random_fr(rng, bits) for _ in range(n_public + 2)]
    constraints = []
    for _ in range(m):
        def comb():
            d = {}
            for _ in range(1 + rng.next() % width):
                d[rng.next() % len(w)] = bls.random_fr(rng, bits)
            return d
        a, b = comb(), comb()
        va = sum(c * w[v] for v, c in a.items()) % R
        vb = sum(c * w[v] for v, c in b.items()) % R
        constraints.append((a, b, {len(w): 1}))
        w.append(va * vb % R)
    return constraints, len(w), w, n_public


def lagrange_at(n, s):
    dom = ref.Domain(n)
    zn = (pow(s, n, R) - 1) * pow(n, -1, R) % R
    out = []
    for i in range(n):
        wi = pow(dom.group_gen, i, R)
        d = (s - wi) % R
        out.append(1 if d == 0 else zn * wi % R * pow(d, -1, R) % R)
    return out


def sparse_eval_at(constraints, nvars, s):
    n = ref.Domain(max(1, len(constraints))).size
    lag = lagrange_at(n, s)
    out = [[0] * nvars for _ in range(3)]
    for i, row in enumerate(constraints):
        for k in range(3):
            for v, coef in row[k].items():
                if v < nvars:
                    out[k][v] = (out[k][v] + coef * lag[i]) % R
    return out


def sparse_domain_evals(constraints, w):
    n = ref.Domain(max(1, len(constraints))).size
    out = [[0] * n for _ in range(3)]
    for i, row in enumerate(constraints):
        for k in range(3):
            out[k][i] = sum(coef * w[v] for v, coef in row[k].items() if v < len(w)) % R
    return out


CIRCUITS = [ref.circuit_mul, ref.circuit_cubic, lambda: chain_circuit(0xc1, 5), lambda: chain_circuit(0xc2, 11, n_public=2)]


def check_evals_small(ctx):
    """domain_evals and eval_at against the dense model (QAP::from_r1cs + Horner evaluation)."""
    for mk in CIRCUITS:
        constraints, nvars, w, npub = mk()
        qap = ref.QAP(constraints, nvars)
        dev = upload(ctx, constraints, nvars)
        assert dev.domain_size == qap.n
        got = [fr_list(x) for x in ctx.r1cs_domain_evals(dev, fr_arr(w))]
        assert got == [list(x) for x in qap.domain_evals(w)]
        dom = ref.Domain(qap.n)
        for s in (ref.P_RAND["s"], 17, 0, 1, pow(dom.group_gen, 3 % qap.n, R)):   # incl. points ON the domain
            ga, gb, gc = (fr_list(x) for x in ctx.r1cs_eval_at(dev, fr_arr([s])[0]))
            assert ga == [ref.poly_eval(p, s) for p in qap.a_polys], s
            assert gb == [ref.poly_eval(p, s) for p in qap.b_polys], s
            assert gc == [ref.poly_eval(p, s) for p in qap.c_polys], s
        dev.free()


def check_duplicate_entries(ctx):
    """A (row, variable) pair given twice in the CSR input: the LAST coefficient wins, as in the assignment loop of
    QAP::from_r1cs (`a_evals[row][var] = coeff`, qap/src/lib.rs:121-138) -- not the sum.  Columns >= num_variables
    are ignored."""
    constraints, nvars, w, _ = chain_circuit(0xd0, 6)
    clean = [to_csr(constraints, k) for k in range(3)]
    dirty = []
    for k in range(3):
        row_ptr, col, val = [0], [], []
        for i, row in enumerate(constraints):
            items = sorted(row[k].items())
            v0, c0 = items[0]
            col += [v0, nvars + 5]                       # a stale coefficient first, and an out-of-range column
            val += [(c0 + 12345 + i) % R, 777]
            for v, coef in items[1:]:
                col.append(v); val.append(coef % R)
            col.append(v0); val.append(c0 % R)           # the value that must survive
            row_ptr.append(len(col))
        dirty.append((np.array(row_ptr, dtype=np.uint32), np.array(col, dtype=np.uint32), fr_arr(val)))
    d_clean = ctx.r1cs_upload(len(constraints), nvars, *clean)
    d_dirty = ctx.r1cs_upload(len(constraints), nvars, *dirty)
    for a, b in zip(ctx.r1cs_domain_evals(d_clean, fr_arr(w)), ctx.r1cs_domain_evals(d_dirty, fr_arr(w))):
        assert (a == b).all()
    s = fr_arr([ref.P_RAND["s"]])[0]
    for a, b in zip(ctx.r1cs_eval_at(d_clean, s), ctx.r1cs_eval_at(d_dirty, s)):
        assert (a == b).all()
    d_clean.free(); d_dirty.free()


def check_evals_long_lines(ctx, m=700):
    """> 256 entries in one line (the constant column, one wide row): block-summed lines; ignored out-of-range columns."""
    constraints, nvars, w, npub = chain_circuit(0xc3, m, bits=64)
    for i in range(0, m, 2):
        constraints[i][0][0] = (i + 7)                       # the constant appears in every second A row
    constraints[5] = ({v: v + 1 for v in range(min(nvars, 400))}, constraints[5][1], constraints[5][2])   # one wide row
    constraints[6][1][nvars + 3] = 99                        # variable index beyond num_variables: dropped
    dev = upload(ctx, constraints, nvars)
    got = [fr_list(x) for x in ctx.r1cs_domain_evals(dev, fr_arr(w))]
    assert got == sparse_domain_evals(constraints, w)
    s = ref.P_RAND["s"]
    got = [fr_list(x) for x in ctx.r1cs_eval_at(dev, fr_arr([s])[0])]
    assert got == sparse_eval_at(constraints, nvars, s)
    dev.free()


def crs_points(crs):
    g1s = lambda xy, inf: [bls.g1_from_mont(list(p), int(f)) for p, f in zip(xy, inf)]
    g2s = lambda xy, inf: [bls.g2_from_mont(list(p), int(f)) for p, f in zip(xy, inf)]
    out = {}
    for k in ("alpha_g1", "beta_g1", "delta_g1"):
        out[k] = bls.g1_from_mont(list(crs[k]), int(not crs[k].any()))
    for k in ("beta_g2", "gamma_g2", "delta_g2"):
        out[k] = bls.g2_from_mont(list(crs[k]), int(not crs[k].any()))
    for k in ("a_g1", "b_g1", "ic_g1", "vk_ic_g1", "h_g1"):
        out[k] = g1s(crs[k], crs[k + "_inf"])
    out["b_g2"] = g2s(crs["b_g2"], crs["b_g2_inf"])
    return out


def params_arr(params):
    return {k: fr_arr([v])[0] for k, v in params.items()}


def check_setup_and_prove(ctx, circuits=None, golden=None):
    """g16_setup_crs == CRS::generate_from_qap of the model (every point of both keys), and g16_prove_r1cs on the
    device-resident key == Prover::prove of the model (proof bytes), for fixed SetupParams and fixed (r, s)."""
    out = {}
    for ci, mk in enumerate(circuits or CIRCUITS):
        constraints, nvars, w, npub = mk()
        qap = ref.QAP(constraints, nvars)
        dev = upload(ctx, constraints, nvars)
        for pname, params in (("Pverify", ref.P_VERIFY), ("Prand", ref.P_RAND)):
            pk, vk = ref.setup(qap, params, npub)
            crs, dpk = ctx.setup_crs(dev, params_arr(params), npub, want_host=True, want_device_pk=True)
            got = crs_points(crs)
            for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2", "a_g1", "b_g1", "b_g2", "ic_g1", "h_g1"):
                assert got[k] == pk[k], (ci, pname, k)
            assert got["gamma_g2"] == vk["gamma_g2"] and got["vk_ic_g1"] == vk["ic_g1"], (ci, pname)
            expect = ref.prove(pk, w, ref.FIXED_R, ref.FIXED_S)
            (a, ai), (b, bi), (c, ci_) = ctx.prove_r1cs(dpk, dev, fr_arr(w), fr_arr([ref.FIXED_R])[0], fr_arr([ref.FIXED_S])[0])
            proof = (bls.g1_from_mont(list(a), ai), bls.g2_from_mont(list(b), bi), bls.g1_from_mont(list(c), ci_))
            assert ref.proof_to_bytes(proof) == ref.proof_to_bytes(expect), (ci, pname)
            # the host-array key through g16_pk_upload + g16_prove gives the same proof
            w_t, h_t = ref.prover_inputs(pk, w)
            hpk = ctx.pk_upload({**{k: crs[k] for k in crs if k not in ("gamma_g2", "vk_ic_g1", "vk_ic_g1_inf")}})
            (a2, ai2), (b2, bi2), (c2, ci2) = ctx.prove(hpk, fr_arr(w_t), fr_arr(h_t) if h_t else None,
                                                        fr_arr([ref.FIXED_R])[0], fr_arr([ref.FIXED_S])[0])
            assert (a2 == a).all() and (b2 == b).all() and (c2 == c).all() and (ai2, bi2, ci2) == (ai, bi, ci_)
            hpk.free(); dpk.free()
            out[f"circuit{ci}_{pname}"] = ref.proof_to_bytes(proof).hex()
            if golden is not None:
                assert golden[f"circuit{ci}_{pname}"] == out[f"circuit{ci}_{pname}"]
        dev.free()
    return out


def check_errors(ctx):
    import groth16_cuda
    constraints, nvars, w, npub = chain_circuit(0xc4, 6)
    dev = upload(ctx, constraints, nvars)
    P = params_arr(ref.P_RAND)

    def fails(fn, text):
        try:
            fn()
        except groth16_cuda.MSMError as e:
            assert text in str(e), str(e)
        else:
            raise AssertionError("expected an error containing " + text)
    zero = dict(P); zero["gamma"] = fr_arr([0])[0]
    fails(lambda: ctx.setup_crs(dev, zero, npub), "must be non-zero")                       # SetupParams::validate
    fails(lambda: ctx.setup_crs(dev, P, nvars), "less than total variables")                # setup/src/lib.rs:148-152
    trunc0 = dict(P); trunc0["delta"] = fr_arr([1 << 64])[0]                                # low limb zero: the reference unwraps None
    fails(lambda: ctx.setup_crs(dev, trunc0, npub), "truncation")
    _, dpk = ctx.setup_crs(dev, P, npub, want_host=False, want_device_pk=True)
    r, s = fr_arr([ref.FIXED_R])[0], fr_arr([ref.FIXED_S])[0]
    fails(lambda: ctx.prove_r1cs(dpk, dev, fr_arr(w[:-1]), r, s), "does not match QAP variables")
    bad1 = list(w); bad1[nvars - 6 + 1] += 1     # the variable defined by constraint 1: Witness::validate looks at row 1 only
    fails(lambda: ctx.prove_r1cs(dpk, dev, fr_arr(bad1), r, s), "does not satisfy QAP constraints")
    bad4 = list(w); bad4[nvars - 1] += 1         # last constraint violated: validate passes, the division fails
    fails(lambda: ctx.prove_r1cs(dpk, dev, fr_arr(bad4), r, s), "division failed")
    ctx.prove_r1cs(dpk, dev, fr_arr(w), r, s)    # and the context is still usable afterwards
    dpk.free(); dev.free()


def check_large_properties(ctx, oracle, log_m=12):
    """Sizes the dense model cannot reach: the polynomial identity A*B - C = H*Z at a random point ties the row
    products, the quotient and the column products (Lagrange basis) together in exact arithmetic; the proof from
    g16_prove_r1cs must equal the five MSMs of the C oracle on the exported key."""
    m = 1 << log_m
    constraints, nvars, w, npub = chain_circuit(0xc5 + log_m, m, width=2, bits=64)
    dev = upload(ctx, constraints, nvars)
    a, b, c = ctx.r1cs_domain_evals(dev, fr_arr(w))
    h = fr_list(ctx.quotient_h(a, b, c))
    x = 0x1234567 * 0x89abcdef % R
    va, vb, vc = (fr_list(v) for v in ctx.r1cs_eval_at(dev, fr_arr([x])[0]))
    A = sum(wi * v for wi, v in zip(w, va)) % R
    B = sum(wi * v for wi, v in zip(w, vb)) % R
    C = sum(wi * v for wi, v in zip(w, vc)) % R
    H = ref.poly_eval(h, x)
    assert (A * B - C) % R == H * (pow(x, dev.domain_size, R) - 1) % R
    dev.free()
