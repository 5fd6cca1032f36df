// C ABI of the sparse R1CS path: setup (CRS generation) and prove for real circuits (include/g16_cuda.h,
// "sparse R1CS").  Also compiled with -DG16_EMU by tests/emu.
#include "api_common.cuh"
#include "r1cs_engine.cuh"

struct g16_r1cs { std::unique_ptr<R1cs> r; };

namespace {

// Scratch of one setup call.  Everything in it derives from the toxic waste (alpha .. s): the destructor -- reached
// on success and on every error path alike -- zeroes the buffers before it returns them to the allocator.
struct SetupWork {
    DevBuf params, blk, lag, vals, ab, ic, h, singles_sc, flags;
    stream_t st = nullptr;
    explicit SetupWork(stream_t s) : st(s) {}
    SetupWork(const SetupWork &) = delete;
    SetupWork &operator=(const SetupWork &) = delete;
    ~SetupWork() {
        DevBuf *all[] = {&params, &blk, &lag, &vals, &ab, &ic, &h, &singles_sc, &flags};
#ifndef G16_EMU
        for (DevBuf *b : all)
            if (b->p) cudaMemsetAsync(b->p, 0, b->cap, st);   // destructors must not throw: unchecked on purpose
        cudaStreamSynchronize(st);
#else
        for (DevBuf *b : all)
            if (b->p) memset(b->p, 0, b->cap);
#endif
        for (DevBuf *b : all) b->release();
    }
};

// device array of (prefix + n) packed points owned by a Bases object
template <class F>
std::unique_ptr<Bases> make_device_bases(Context *c, size_t n_total) {
    std::unique_ptr<Bases> b(new Bases);
    b->ctx = c; b->group = GroupOf<F>::id; b->n = n_total;
    BasesShard sh;
    sh.dev = 0; sh.cuda_dev = c->devs[0].id; sh.begin = 0; sh.n = n_total; sh.owned = true;
    sh.pts = (uint32_t *)dev_alloc(std::max<size_t>(n_total, 1) * 2 * FieldWords<F>::N * 4);
    b->shards.push_back(sh);
    return b;
}

template <class F>
void export_points(Device &dv, const uint32_t *d_pts, size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    constexpr size_t W = 2 * FieldWords<F>::N;
    if (!n) return;
    if (out_xy) copy_d2h(out_xy, d_pts, n * W * 4, dv.stream);
    if (out_inf) {
        uint8_t *d_fl = dv.ws.fb_flags.as<uint8_t>(n);
        k_export_flags<F>(dv.stream, n, d_pts, d_fl);
        copy_d2h(out_inf, d_fl, n, dv.stream);
        stream_sync(dv.stream);   // fb_flags is reused by the next export
    }
}

}  // namespace

extern "C" {

int g16_r1cs_upload(g16_ctx *ctx, size_t num_constraints, size_t num_variables, const g16_csr *a, const g16_csr *b,
                    const g16_csr *c, g16_r1cs **out) {
    if (!ctx || !out) return G16_ERR_INVALID;
    *out = nullptr;
    return guarded(ctx, [&] {
        single_device(ctx);
        require(a && b && c, "NULL matrix");
        CsrView v[3] = {{a->row_ptr, a->col, a->val}, {b->row_ptr, b->col, b->val}, {c->row_ptr, c->col, c->val}};
        for (auto &m : v) require(num_constraints == 0 || (m.row_ptr && (m.row_ptr[num_constraints] == 0 || (m.col && m.val))), "NULL CSR array");
        std::unique_ptr<g16_r1cs> h(new g16_r1cs);
        h->r = r1cs_upload(&ctx->c, num_constraints, num_variables, v);
        *out = h.release();
    });
}
void g16_r1cs_free(g16_r1cs *r) { delete r; }
size_t g16_r1cs_domain_size(const g16_r1cs *r) { return r ? (size_t)1 << r->r->log_n : 0; }

int g16_r1cs_domain_evals(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t *assignment, size_t num_vars, uint64_t *a_evals,
                          uint64_t *b_evals, uint64_t *c_evals) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(r1cs && r1cs->r->ctx == &ctx->c, "bad r1cs handle");
        require(assignment && a_evals && b_evals && c_evals, "NULL argument");
        const R1cs &r = *r1cs->r;
        if (num_vars != r.nv)
            throw Error{G16_ERR_LENGTH, "Assignment length " + std::to_string(num_vars) + " does not match QAP variables " + std::to_string(r.nv)};
        size_t n = (size_t)1 << r.log_n;
        uint32_t *d_w = dv.ws.prove_w.as<uint32_t>(r.nv * 8 + 8);
        copy_h2d(d_w, assignment, r.nv * 32, dv.stream);
        uint32_t *abc = dv.ws.ntt_abc.as<uint32_t>(3 * n * 8 + 8);
        r1cs_domain_evals_device(dv, r, d_w, abc);
        copy_d2h(a_evals, abc, n * 32, dv.stream);
        copy_d2h(b_evals, abc + n * 8, n * 32, dv.stream);
        copy_d2h(c_evals, abc + 2 * n * 8, n * 32, dv.stream);
        stream_sync(dv.stream);
    });
}

int g16_r1cs_eval_at(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t s[4], uint64_t *a_vals, uint64_t *b_vals,
                     uint64_t *c_vals) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(r1cs && r1cs->r->ctx == &ctx->c, "bad r1cs handle");
        require(s && a_vals && b_vals && c_vals, "NULL argument");
        const R1cs &r = *r1cs->r;
        size_t n = (size_t)1 << r.log_n;
        SetupWork w(dv.stream);
        uint64_t params[20] = {0};
        memcpy(params + 16, s, 32);
        uint32_t *d_params = w.params.as<uint32_t>(40);
        copy_h2d(d_params, params, 160, dv.stream);
        uint32_t *blk = w.blk.as<uint32_t>(k_setup_scalar_words());
        const uint32_t *consts = ntt_prepare(dv, r.log_n);
        k_setup_scalars(dv.stream, d_params, consts, r.log_n, 0, blk);
        uint32_t *vals = w.vals.as<uint32_t>(3 * r.nv * 8);
        r1cs_eval_at_device(dv, r, blk, w.lag.as<uint32_t>(n * 8), vals);
        copy_d2h(a_vals, vals, r.nv * 32, dv.stream);
        copy_d2h(b_vals, vals + r.nv * 8, r.nv * 32, dv.stream);
        copy_d2h(c_vals, vals + 2 * r.nv * 8, r.nv * 32, dv.stream);
        stream_sync(dv.stream);
    });
}

// CRS::generate_from_qap (crates/groth16-setup/src/lib.rs:141-268) from the sparse constraint system.
int g16_setup_crs(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t alpha[4], const uint64_t beta[4], const uint64_t gamma[4],
                  const uint64_t delta[4], const uint64_t s[4], size_t num_public, g16_crs_host *out, g16_pk **pk_out) {
    if (!ctx) return G16_ERR_INVALID;
    if (pk_out) *pk_out = nullptr;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(r1cs && r1cs->r->ctx == &ctx->c, "bad r1cs handle");
        require(alpha && beta && gamma && delta && s, "NULL argument");
        const R1cs &r = *r1cs->r;
        // SetupParams::validate (setup/src/lib.rs:127-136) and the num_public check (:148-152)
        auto is_zero = [](const uint64_t *x) { return (x[0] | x[1] | x[2] | x[3]) == 0; };
        if (is_zero(alpha) || is_zero(beta) || is_zero(gamma) || is_zero(delta))
            throw Error{G16_ERR_INVALID, "Invalid setup parameters: Setup parameters must be non-zero"};
        if (num_public >= r.nv)
            throw Error{G16_ERR_INVALID, "Invalid setup parameters: Number of public inputs must be less than total variables"};
        size_t n = (size_t)1 << r.log_n, nv = r.nv;
        stream_t st = dv.stream;
        SetupWork w(st);
        uint64_t params[20];
        memcpy(params, alpha, 32); memcpy(params + 4, beta, 32); memcpy(params + 8, gamma, 32); memcpy(params + 12, delta, 32);
        memcpy(params + 16, s, 32);
        uint32_t *d_params = w.params.as<uint32_t>(40);
        copy_h2d(d_params, params, 160, st);
        uint32_t *blk = w.blk.as<uint32_t>(k_setup_scalar_words());
        const uint32_t *consts = ntt_prepare(dv, r.log_n);
        k_setup_scalars(st, d_params, consts, r.log_n, 1, blk);
        uint32_t flag_words[8];
        copy_d2h(flag_words, blk + 8 * 8, 32, st);
        stream_sync(st);
        // the reference unwraps inverse(gamma_t) / inverse(delta_t) of the TRUNCATED values and would panic
        if (flag_words[0] & 3u) throw Error{G16_ERR_INVALID, "Invalid setup parameters: gamma or delta is zero after 64-bit truncation"};
        // exponents
        uint32_t *vals = w.vals.as<uint32_t>(3 * nv * 8);
        r1cs_eval_at_device(dv, r, blk, w.lag.as<uint32_t>(n * 8), vals);
        uint32_t *ab = w.ab.as<uint32_t>(2 * nv * 8), *ic = w.ic.as<uint32_t>(nv * 8), *hx = w.h.as<uint32_t>(n * 8);
        k_crs_exponents(st, nv, vals, blk, (uint32_t)num_public, ab, ic);
        k_crs_h_exponents(st, n, blk, hx);
        // single points use the FULL-width parameters (setup/src/lib.rs:166-171): G1 [alpha, beta, delta], G2 [beta, gamma, delta]
        uint64_t sg[24];
        memcpy(sg, alpha, 32); memcpy(sg + 4, beta, 32); memcpy(sg + 8, delta, 32);
        memcpy(sg + 12, beta, 32); memcpy(sg + 16, gamma, 32); memcpy(sg + 20, delta, 32);
        uint32_t *d_sg = w.singles_sc.as<uint32_t>(48);
        copy_h2d(d_sg, sg, 192, st);
        size_t n_ic = nv - num_public - 1;
        // proving key arrays with their ad-hoc single points in front (layout of g16_pk, see g16_prove)
        auto A = make_device_bases<Fq>(&ctx->c, nv + 2);
        auto B2 = make_device_bases<Fq2>(&ctx->c, nv + 2);
        auto B1 = make_device_bases<Fq>(&ctx->c, nv + 1);
        auto IC = make_device_bases<Fq>(&ctx->c, nv);          // vk part first, then the pk part
        auto H = make_device_bases<Fq>(&ctx->c, n);
        uint32_t *pa = A->shards[0].pts, *pb2 = B2->shards[0].pts, *pb1 = B1->shards[0].pts, *pic = IC->shards[0].pts, *ph = H->shards[0].pts;
        uint32_t *s1 = dv.ws.staging.as<uint32_t>(3 * 24 + 3 * 48);   // alpha_g1, beta_g1, delta_g1 | beta_g2, gamma_g2, delta_g2
        uint32_t *s2 = s1 + 3 * 24;
        fixed_base_device<Fq>(dv, G1_GENERATOR, d_sg, 3, s1);
        fixed_base_device<Fq2>(dv, G2_GENERATOR, d_sg + 24, 3, s2);
        fixed_base_device<Fq>(dv, G1_GENERATOR, ab, nv, pa + 2 * 24);
        fixed_base_device<Fq>(dv, G1_GENERATOR, ab + nv * 8, nv, pb1 + 1 * 24);
        fixed_base_device<Fq2>(dv, G2_GENERATOR, ab + nv * 8, nv, pb2 + 2 * 48);
        fixed_base_device<Fq>(dv, G1_GENERATOR, ic, nv, pic);
        fixed_base_device<Fq>(dv, G1_GENERATOR, hx, n, ph);
        copy_d2d(pa, s1, 96, st); copy_d2d(pa + 24, s1 + 48, 96, st);             // [alpha_g1, delta_g1]
        copy_d2d(pb1, s1 + 24, 96, st);                                           // [beta_g1]
        copy_d2d(pb2, s2, 192, st); copy_d2d(pb2 + 48, s2 + 96, 192, st);         // [beta_g2, delta_g2]
        if (out) {
            if (out->alpha_g1) copy_d2h(out->alpha_g1, s1, 96, st);
            if (out->beta_g1) copy_d2h(out->beta_g1, s1 + 24, 96, st);
            if (out->delta_g1) copy_d2h(out->delta_g1, s1 + 48, 96, st);
            if (out->beta_g2) copy_d2h(out->beta_g2, s2, 192, st);
            if (out->gamma_g2) copy_d2h(out->gamma_g2, s2 + 48, 192, st);
            if (out->delta_g2) copy_d2h(out->delta_g2, s2 + 96, 192, st);
            export_points<Fq>(dv, pa + 2 * 24, nv, out->a_g1, out->a_g1_inf);
            export_points<Fq>(dv, pb1 + 24, nv, out->b_g1, out->b_g1_inf);
            export_points<Fq2>(dv, pb2 + 2 * 48, nv, out->b_g2, out->b_g2_inf);
            export_points<Fq>(dv, pic, num_public + 1, out->vk_ic_g1, out->vk_ic_g1_inf);
            export_points<Fq>(dv, pic + (num_public + 1) * 24, n_ic, out->ic_g1, out->ic_g1_inf);
            export_points<Fq>(dv, ph, n, out->h_g1, out->h_g1_inf);
        }
        stream_sync(st);
        if (pk_out) {
            std::unique_ptr<g16_pk> pk(new g16_pk);
            pk->ctx = &ctx->c;
            pk->a = std::move(A); pk->b2 = std::move(B2); pk->b1 = std::move(B1); pk->h = std::move(H);
            // the proving key keeps only the private part of the IC query
            pk->ic = make_device_bases<Fq>(&ctx->c, n_ic);
            copy_d2d(pk->ic->shards[0].pts, pic + (num_public + 1) * 24, n_ic * 96, st);
            stream_sync(st);
            pk->a_len = nv; pk->b1_len = nv; pk->b2_len = nv; pk->ic_len = n_ic; pk->h_len = n; pk->num_public = num_public;
            *pk_out = pk.release();
        }
    });
}

// Prover::prove (crates/groth16-core/src/lib.rs:139-272) from the un-truncated witness: validate, domain
// evaluations, quotient polynomial, 64-bit truncations and the MSM schedule, all on the device.
int g16_prove_r1cs(g16_ctx *ctx, const g16_pk *pk, const g16_r1cs *r1cs, const uint64_t *assignment, size_t num_vars,
                   const uint64_t r_[4], const uint64_t s_[4], uint64_t a_xy[12], uint8_t *a_inf, uint64_t b_xy[24],
                   uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(pk && pk->ctx == &ctx->c, "bad pk handle");
        require(r1cs && r1cs->r->ctx == &ctx->c, "bad r1cs handle");
        require(assignment && r_ && s_ && a_xy && b_xy && c_xy, "NULL argument");
        const R1cs &r = *r1cs->r;
        if (num_vars != r.nv)
            throw Error{G16_ERR_LENGTH, "Invalid witness: Assignment length " + std::to_string(num_vars) + " does not match QAP variables " + std::to_string(r.nv)};
        require(num_vars > pk->num_public, "assignment shorter than the public inputs");
        size_t n = (size_t)1 << r.log_n;
        stream_t st = dv.stream;
        Workspace &ws = dv.ws;
        uint32_t *d_full = ws.ntt_out.as<uint32_t>((n + 2 * r.nv) * 8 + 16);   // H coefficients | full witness | truncated witness
        uint32_t *d_h = d_full, *d_wfull = d_full + n * 8, *d_w = d_wfull + r.nv * 8;
        uint32_t *abc = ws.ntt_abc.as<uint32_t>(3 * n * 8 + 8);
        uint32_t *flags = abc + 3 * n * 8;     // [0] rows with A*B != C, [1] Witness::validate failed
        copy_h2d(d_wfull, assignment, r.nv * 32, st);
        dev_memset(flags, 0, 8, st);
        r1cs_domain_evals_device(dv, r, d_wfull, abc);
        k_validate_row(st, abc, (uint32_t)n, flags);
        quotient_device(dv, r.log_n, abc, flags, d_h);
        k_truncate64(st, n, d_h, d_h);
        k_truncate64(st, r.nv, d_wfull, d_w);
        uint32_t fl[2] = {0, 0};
        ProveDeviceInputs in;
        in.d_w = d_w; in.d_h = d_h; in.d_flags = flags; in.flags_out = fl;
        prove_single_device(&ctx->c, pk, nullptr, num_vars, nullptr, n, r_, s_, a_xy, a_inf, b_xy, b_inf, c_xy, c_inf, &in);
        if (fl[1]) throw Error{G16_ERR_INVALID, "Invalid witness: Witness does not satisfy QAP constraints"};
        if (fl[0]) throw Error{G16_ERR_INVALID, "Polynomial division failed: non-zero remainder"};
    });
}

}  // extern "C"
