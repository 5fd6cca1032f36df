// C ABI of libg16cuda.so (declared in include/g16_cuda.h).  Also compiled as plain C++ with
// -DG16_EMU by tests/emu (see rt.cuh) -- that build is test infrastructure only.
#include "../../include/g16_cuda.h"
#include "api_common.cuh"

extern "C" {

int g16_device_count(void) {
#ifndef G16_EMU
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
#else
    return 1;
#endif
}

int g16_ctx_create(const int *devices, int ndev, g16_ctx **out) {
    if (!out) return G16_ERR_INVALID;
    *out = nullptr;
    return guarded(nullptr, [&] {
#ifndef G16_EMU
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error{G16_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0")};
#else
        int count = 64;
#endif
        std::unique_ptr<g16_ctx> ctx(new g16_ctx);
        std::vector<int> ids;
        if (!devices || ndev <= 0) {
            int cur = 0;
#ifndef G16_EMU
            G16_CUDA_CHECK(cudaGetDevice(&cur));
#endif
            ids.push_back(cur);
        } else {
            ids.assign(devices, devices + ndev);
        }
        for (int id : ids) {
            if (id < 0 || id >= count) throw Error{G16_ERR_NO_DEVICE, "device index out of range"};
            Device d;
            d.id = id;
#ifndef G16_EMU
            G16_CUDA_CHECK(cudaSetDevice(id));
            int sms = 0;
            G16_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, id));
            d.sm_count = (uint32_t)std::max(1, sms);
            G16_CUDA_CHECK(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
            d.own_stream = true;
#endif
            ctx->c.devs.push_back(std::move(d));
        }
        *out = ctx.release();
    });
}

void g16_ctx_destroy(g16_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->c.devs) {
#ifndef G16_EMU
        cudaSetDevice(d.id);
        if (d.stream) cudaStreamSynchronize(d.stream);
#endif
        d.ws.release();
        d.timer.destroy();
        for (auto &l : d.extra) {
#ifndef G16_EMU
            if (l->stream) cudaStreamSynchronize(l->stream);
#endif
            l->ws.release();
            l->timer.destroy();
#ifndef G16_EMU
            if (l->own_stream && l->stream) cudaStreamDestroy(l->stream);
#endif
        }
#ifndef G16_EMU
        if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
#endif
    }
#ifndef G16_EMU
    if (ctx->c.prove_epoch) cudaEventDestroy((cudaEvent_t)ctx->c.prove_epoch);
#endif
    delete ctx;
}

const char *g16_last_error(const g16_ctx *ctx) { return ctx ? ctx->c.err.c_str() : create_error().c_str(); }

int g16_ctx_set_stream(g16_ctx *ctx, void *cuda_stream) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &d = single_device(ctx);
#ifndef G16_EMU
        if (d.own_stream && d.stream) { cudaStreamSynchronize(d.stream); cudaStreamDestroy(d.stream); }
        d.stream = (cudaStream_t)cuda_stream;
#else
        d.stream = cuda_stream;
#endif
        d.own_stream = false;
    });
}

int g16_ctx_synchronize(g16_ctx *ctx) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        for (auto &d : ctx->c.devs) {
            set_device(d.id);
            stream_sync(d.stream);
            for (auto &l : d.extra) stream_sync(l->stream);
        }
    });
}

int g16_ctx_set_window_bits(g16_ctx *ctx, unsigned c) {
    if (!ctx || c > 24 || c == 1) return G16_ERR_INVALID;
    ctx->c.c_override = c;
    return G16_OK;
}

int g16_ctx_set_h2d_pipeline_min(g16_ctx *ctx, size_t min_scalars) {
    if (!ctx) return G16_ERR_INVALID;
    ctx->c.h2d_pipe_min = min_scalars ? min_scalars : H2D_PIPE_MIN;
    return G16_OK;
}

// ---- bases ------------------------------------------------------------------------------
extern "C++" {
template <class F>
static int bases_upload_impl(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, g16_bases **out) {
    if (!ctx || !out) return G16_ERR_INVALID;
    *out = nullptr;
    return guarded(ctx, [&] {
        require(xy || n == 0, "xy is NULL");
        std::unique_ptr<g16_bases> h(new g16_bases);
        h->b = bases_upload<F>(&ctx->c, xy, inf, n);
        *out = h.release();
    });
}
}  // extern "C++"
int g16_g1_bases_upload(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, g16_bases **out) {
    return bases_upload_impl<Fq>(ctx, xy, inf, n, out);
}
int g16_g2_bases_upload(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, g16_bases **out) {
    return bases_upload_impl<Fq2>(ctx, xy, inf, n, out);
}

extern "C++" {
template <class F>
static int bases_from_device_impl(g16_ctx *ctx, const void *dev_xy, size_t n, g16_bases **out) {
    if (!ctx || !out) return G16_ERR_INVALID;
    *out = nullptr;
    return guarded(ctx, [&] {
        single_device(ctx);
        require(dev_xy || n == 0, "dev_xy is NULL");
        std::unique_ptr<g16_bases> h(new g16_bases);
        h->b.reset(new Bases);
        h->b->ctx = &ctx->c; h->b->group = GroupOf<F>::id; h->b->n = n;
        BasesShard sh;
        sh.dev = 0; sh.cuda_dev = ctx->c.devs[0].id; sh.pts = (uint32_t *)dev_xy; sh.begin = 0; sh.n = n; sh.owned = false;
        h->b->shards.push_back(sh);
        *out = h.release();
    });
}
}  // extern "C++"
int g16_g1_bases_from_device(g16_ctx *ctx, const void *dev_xy, size_t n, g16_bases **out) {
    return bases_from_device_impl<Fq>(ctx, dev_xy, n, out);
}
int g16_g2_bases_from_device(g16_ctx *ctx, const void *dev_xy, size_t n, g16_bases **out) {
    return bases_from_device_impl<Fq2>(ctx, dev_xy, n, out);
}
int g16_bases_precompute(g16_ctx *ctx, g16_bases *bases, unsigned window_bits, size_t budget_bytes, unsigned *used_bits) {
    if (!ctx || !bases) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(bases->b->ctx == &ctx->c, "bases belong to another context");
        if (budget_bytes == 0) budget_bytes = (size_t)48 << 30;
        unsigned c = bases->b->group == GROUP_G1 ? bases_precompute<Fq>(&ctx->c, bases->b.get(), window_bits, budget_bytes)
                                                 : bases_precompute<Fq2>(&ctx->c, bases->b.get(), window_bits, budget_bytes);
        if (used_bits) *used_bits = c;
    });
}
void g16_bases_free(g16_bases *bases) { delete bases; }
size_t g16_bases_len(const g16_bases *bases) { return bases ? bases->b->n : 0; }

// ---- MSM --------------------------------------------------------------------------------
extern "C++" {
template <class F>
static int msm_impl(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, uint64_t *out_xy,
                    uint8_t *out_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(bases && out_xy, "NULL argument");
        require(scalars || n == 0, "scalars is NULL");
        require(bases->b->ctx == &ctx->c, "bases belong to another context");
        msm_host<F>(&ctx->c, bases->b.get(), scalars, n, out_xy, out_inf);
    });
}
}  // extern "C++"
int g16_g1_msm(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, uint64_t out_xy[12],
               uint8_t *out_inf) {
    return msm_impl<Fq>(ctx, bases, scalars, n, out_xy, out_inf);
}
int g16_g2_msm(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, uint64_t out_xy[24],
               uint8_t *out_inf) {
    return msm_impl<Fq2>(ctx, bases, scalars, n, out_xy, out_inf);
}

extern "C++" {
template <class F>
static int msm_oneshot_impl(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, const uint64_t *scalars, size_t n,
                            uint64_t *out_xy, uint8_t *out_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(out_xy, "out_xy is NULL");
        require((xy && scalars) || n == 0, "NULL input");
        std::unique_ptr<Bases> b = bases_upload<F>(&ctx->c, xy, inf, n);
        msm_host<F>(&ctx->c, b.get(), scalars, n, out_xy, out_inf);
    });
}
}  // extern "C++"
int g16_g1_msm_oneshot(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, const uint64_t *scalars, size_t n,
                       uint64_t out_xy[12], uint8_t *out_inf) {
    return msm_oneshot_impl<Fq>(ctx, xy, inf, scalars, n, out_xy, out_inf);
}
int g16_g2_msm_oneshot(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, const uint64_t *scalars, size_t n,
                       uint64_t out_xy[24], uint8_t *out_inf) {
    return msm_oneshot_impl<Fq2>(ctx, xy, inf, scalars, n, out_xy, out_inf);
}

extern "C++" {
template <class F>
static int msm_device_impl(g16_ctx *ctx, const g16_bases *bases, const void *dev_scalars, size_t n, void *dev_out_affine,
                           void *dev_out_partial) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(bases && bases->b->ctx == &ctx->c && bases->b->group == GroupOf<F>::id, "bad bases handle");
        require(bases->b->shards.size() == 1, "bases are sharded");
        if (n > bases->b->n) throw Error{G16_ERR_LENGTH, "more scalars than bases"};
        require(dev_scalars || n == 0, "dev_scalars is NULL");
        msm_run<F>(dv, bases->b->shards[0], (const uint32_t *)dev_scalars, n, true, ctx->c.c_override,
                   (uint32_t *)dev_out_partial, (uint32_t *)dev_out_affine);
    });
}
}  // extern "C++"
int g16_g1_msm_device(g16_ctx *ctx, const g16_bases *bases, const void *dev_scalars, size_t n, void *dev_out_affine,
                      void *dev_out_partial) {
    return msm_device_impl<Fq>(ctx, bases, dev_scalars, n, dev_out_affine, dev_out_partial);
}
int g16_g2_msm_device(g16_ctx *ctx, const g16_bases *bases, const void *dev_scalars, size_t n, void *dev_out_affine,
                      void *dev_out_partial) {
    return msm_device_impl<Fq2>(ctx, bases, dev_scalars, n, dev_out_affine, dev_out_partial);
}

extern "C++" {
template <class F>
static int msm_h2d_impl(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, void *dev_out_affine,
                        void *dev_out_partial) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        single_device(ctx);
        require(bases && bases->b->ctx == &ctx->c, "bad bases handle");
        require(scalars || n == 0, "scalars is NULL");
        require(dev_out_affine || dev_out_partial, "no output");
        msm_launch<F>(&ctx->c, bases->b.get(), scalars, n, 0, (uint32_t *)dev_out_partial, (uint32_t *)dev_out_affine);
    });
}
}  // extern "C++"
int g16_g1_msm_async(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, void *dev_out_affine,
                     void *dev_out_partial) {
    return msm_h2d_impl<Fq>(ctx, bases, scalars, n, dev_out_affine, dev_out_partial);
}
int g16_g2_msm_async(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, void *dev_out_affine,
                     void *dev_out_partial) {
    return msm_h2d_impl<Fq2>(ctx, bases, scalars, n, dev_out_affine, dev_out_partial);
}

extern "C++" {
template <class F>
static int combine_impl(g16_ctx *ctx, const void *dev_partials, size_t k, void *dev_out_affine) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(dev_out_affine && (dev_partials || k == 0), "NULL argument");
        k_partial_combine<F>(dv.stream, (const uint32_t *)dev_partials, (uint32_t)k, nullptr, (uint32_t *)dev_out_affine);
    });
}
}  // extern "C++"
int g16_g1_combine_partials_device(g16_ctx *ctx, const void *dev_partials, size_t k, void *dev_out_affine) {
    return combine_impl<Fq>(ctx, dev_partials, k, dev_out_affine);
}
int g16_g2_combine_partials_device(g16_ctx *ctx, const void *dev_partials, size_t k, void *dev_out_affine) {
    return combine_impl<Fq2>(ctx, dev_partials, k, dev_out_affine);
}

// ---- fixed base ---------------------------------------------------------------------------
extern "C++" {
template <class F>
static int fixed_base_impl(g16_ctx *ctx, const uint64_t *base_xy, const uint64_t *scalars, size_t n, uint64_t *out_xy,
                           uint8_t *out_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(base_xy && (n == 0 || (scalars && out_xy)), "NULL argument");
        fixed_base_host<F>(&ctx->c, base_xy, scalars, n, out_xy, out_inf);
    });
}
}  // extern "C++"
int g16_g1_fixed_base_mul(g16_ctx *ctx, const uint64_t base_xy[12], const uint64_t *scalars, size_t n,
                          uint64_t *out_xy, uint8_t *out_inf) {
    return fixed_base_impl<Fq>(ctx, base_xy, scalars, n, out_xy, out_inf);
}
int g16_g2_fixed_base_mul(g16_ctx *ctx, const uint64_t base_xy[24], const uint64_t *scalars, size_t n,
                          uint64_t *out_xy, uint8_t *out_inf) {
    return fixed_base_impl<Fq2>(ctx, base_xy, scalars, n, out_xy, out_inf);
}
extern "C++" {
template <class F>
static int fixed_base_device_impl(g16_ctx *ctx, const uint64_t *base_xy, const void *dev_scalars, size_t n,
                                  void *dev_out_xy) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(base_xy && (n == 0 || (dev_scalars && dev_out_xy)), "NULL argument");
        fixed_base_device<F>(dv, base_xy, (const uint32_t *)dev_scalars, n, (uint32_t *)dev_out_xy);
    });
}
}  // extern "C++"
int g16_g1_fixed_base_mul_device(g16_ctx *ctx, const uint64_t base_xy[12], const void *dev_scalars, size_t n,
                                 void *dev_out_xy) {
    return fixed_base_device_impl<Fq>(ctx, base_xy, dev_scalars, n, dev_out_xy);
}
int g16_g2_fixed_base_mul_device(g16_ctx *ctx, const uint64_t base_xy[24], const void *dev_scalars, size_t n,
                                 void *dev_out_xy) {
    return fixed_base_device_impl<Fq2>(ctx, base_xy, dev_scalars, n, dev_out_xy);
}

// ---- proving key + prove schedule --------------------------------------------------------------
// Concatenate `k` single points (host) in front of a host array and upload.
extern "C++" {
template <class F>
static std::unique_ptr<Bases> upload_with_prefix(Context *c, std::initializer_list<const uint64_t *> singles,
                                                 const uint64_t *xy, const uint8_t *inf, size_t n) {
    constexpr size_t PW = FieldWords<F>::N;  // u64 words per point
    size_t k = singles.size();
    std::vector<uint64_t> buf((k + n) * PW);
    std::vector<uint8_t> flags(k + n, 0);
    size_t i = 0;
    for (const uint64_t *p : singles) {
        memcpy(buf.data() + i * PW, p, PW * 8);
        // a single point at infinity is passed as all-zero coordinates
        bool z = true;
        for (size_t j = 0; j < PW; ++j) z = z && p[j] == 0;
        flags[i] = z;
        ++i;
    }
    if (n) memcpy(buf.data() + k * PW, xy, n * PW * 8);
    if (inf) memcpy(flags.data() + k, inf, n);
    return bases_upload<F>(c, buf.data(), flags.data(), k + n);
}
}  // extern "C++"

int g16_pk_upload(g16_ctx *ctx, const g16_pk_host *pk, g16_pk **out) {
    if (!ctx || !pk || !out) return G16_ERR_INVALID;
    *out = nullptr;
    return guarded(ctx, [&] {
        require(pk->alpha_g1 && pk->beta_g1 && pk->delta_g1 && pk->beta_g2 && pk->delta_g2, "NULL single point");
        std::unique_ptr<g16_pk> h(new g16_pk);
        h->ctx = &ctx->c;
        h->a = upload_with_prefix<Fq>(&ctx->c, {pk->alpha_g1, pk->delta_g1}, pk->a_g1, pk->a_g1_inf, pk->a_len);
        h->b2 = upload_with_prefix<Fq2>(&ctx->c, {pk->beta_g2, pk->delta_g2}, pk->b_g2, pk->b_g2_inf, pk->b2_len);
        h->b1 = upload_with_prefix<Fq>(&ctx->c, {pk->beta_g1}, pk->b_g1, pk->b_g1_inf, pk->b1_len);
        h->ic = upload_with_prefix<Fq>(&ctx->c, {}, pk->ic_g1, pk->ic_g1_inf, pk->ic_len);
        h->h = upload_with_prefix<Fq>(&ctx->c, {}, pk->h_g1, pk->h_g1_inf, pk->h_len);
        h->a_len = pk->a_len; h->b1_len = pk->b1_len; h->b2_len = pk->b2_len; h->ic_len = pk->ic_len; h->h_len = pk->h_len;
        h->num_public = pk->num_public;
        *out = h.release();
    });
}
int g16_pk_precompute_bits(g16_ctx *ctx, g16_pk *pk, unsigned scalar_bits) {
    if (!ctx || !pk) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(pk->ctx == &ctx->c, "bad pk handle");
        require(scalar_bits <= 256, "scalar_bits must be at most 256 (0 = full width)");
        size_t budget = (size_t)24 << 30;
        for (Bases *b : {pk->a.get(), pk->b1.get(), pk->ic.get(), pk->h.get()})
            if (b->n >= 256) bases_precompute<Fq>(&ctx->c, b, 0, budget, scalar_bits);
        if (pk->b2->n >= 256) bases_precompute<Fq2>(&ctx->c, pk->b2.get(), 0, budget, scalar_bits);
    });
}
int g16_pk_precompute(g16_ctx *ctx, g16_pk *pk) { return g16_pk_precompute_bits(ctx, pk, 0); }
void g16_pk_free(g16_pk *pk) { delete pk; }


extern "C++" {
// Single-device fast path of the prove schedule: the assignment is copied to the device once, every MSM gets its
// (prefix ++ assignment) scalar vector by a device-to-device copy, the five big MSMs run on five lanes (stream +
// workspace each), and pi_C = private part + H + s pi_A + r pi_B' is finished on the device: the two multiplications
// are double-and-add chains side by side in one warp (ScalarMulAffine, 2 ms; the 3-point MSM through the whole
// pipeline that did this before cost 4 ms of serial window folding), then one four-term fold.  The host waits once.
//
// The prove is bound by the SUM of the work of its kernels, not by their latencies (profiles/r02_run{9..13}_prove_timeline_*.txt):
// blocks of a younger grid are not dispatched while an older grid of the same priority has pending blocks, so sort stages
// and tails of one lane queue behind the accumulation grids of the others, but the GPU is busy throughout.  Measured
// against this schedule and not adopted: sort stages of all lanes first + accumulations chained by events + tails on
// high-priority streams with blocks small enough for the registers one retiring accumulate block frees (same 52.5 ms at
// full width, 30.8 instead of 27.3 ms with the reference's 64-bit scalars: the G2 bucket reduction is 5 ms of real
// multiplier work at low efficiency wherever it runs, and small tail blocks cost an extra tree level); the same for
// pi_A and pi_B' only (54.3 / 30.4 ms); independent lanes with the chains on the producing lanes (56.9 / 30.9 ms).
void prove_single_device(Context *c, const g16_pk *pk, const uint64_t *w, size_t num_vars, const uint64_t *h,
                         size_t num_h, const uint64_t *r, const uint64_t *s, uint64_t *a_xy, uint8_t *a_inf,
                         uint64_t *b_xy, uint8_t *b_inf, uint64_t *c_xy, uint8_t *c_inf, const ProveDeviceInputs *dev) {
    constexpr size_t PW1 = 48, AW1 = 25, PW2 = 96, AW2 = 49;
    Device &d0 = c->devs[0];
    set_device(d0.id);
    Device &LA = lane_of(d0, 0), &LB = lane_of(d0, 1), &LH = lane_of(d0, 2), &LB1 = lane_of(d0, 3), &LC = lane_of(d0, 4);
    unsigned co = c->c_override;
#ifndef G16_EMU
    if (d0.timer.enabled) {   // test hook: per-lane stage timeline of this prove (g16_ctx_prove_timeline)
        for (Device *l : {&LB, &LH, &LB1, &LC}) { l->timer.enabled = true; l->timer.valid = false; }
        if (!c->prove_epoch) { cudaEvent_t e; G16_CUDA_CHECK(cudaEventCreate(&e)); c->prove_epoch = e; }
        G16_CUDA_CHECK(cudaEventRecord((cudaEvent_t)c->prove_epoch, LA.stream));
    }
#endif
    // small host staging block (kept alive until the final synchronisation): prefixes and the two multipliers
    std::vector<uint64_t> hs(4 * 16);
    auto put = [&](size_t slot, const uint64_t *x) { memcpy(hs.data() + 4 * slot, x, 32); };
    put(0, FR_ONE_MONT); put(1, r);          // pi_A prefix   [1, r]
    put(2, FR_ONE_MONT); put(3, s);          // pi_B prefix   [1, s]
    put(4, FR_ONE_MONT);                     // pi_B' prefix  [1]
    put(6, s); put(7, r);                    // multipliers of pi_A and pi_B' inside pi_C

    // assignment: one H2D, shared by four MSMs
    // (or already on the device, produced on lane 0's stream by the R1CS path)
    const uint32_t *d_w = dev ? dev->d_w : nullptr;
    if (!d_w) {
        uint32_t *buf = LA.ws.prove_w.as<uint32_t>(num_vars * 8 + 8);
        copy_h2d(buf, w, num_vars * 32, LA.stream);
        d_w = buf;
    }
    constexpr size_t AB1 = (2 * AW1 + 3) / 4 * 4;     // two affine records, padded: what follows is accessed 16 bytes at a time
    uint32_t *d_misc = LA.ws.prove_misc.as<uint32_t>(16 * 8 + AB1 + 4 * PW1 + AW1 + 64);
    uint32_t *d_small = d_misc;                       // 16 scalars
    uint32_t *d_ab1 = d_misc + 16 * 8;                // pi_A and pi_B' affine records, back to back (inputs of the chains)
    uint32_t *d_cparts = d_ab1 + AB1;                 // 4 projective terms of pi_C: private part, H, s pi_A, r pi_B'
    uint32_t *d_c_aff = d_cparts + 4 * PW1;           // pi_C affine
    copy_h2d(d_small, hs.data(), 16 * 32, LA.stream);
    for (Device *l : {&LB, &LB1, &LC}) stream_wait(l->stream, LA.stream);
    if (dev) stream_wait(LH.stream, LA.stream);

    auto prefixed = [&](Device &L, size_t slot, size_t k, size_t n) {
        uint32_t *d = L.ws.scalars.as<uint32_t>((k + n) * 8 + 8);
        copy_d2d(d, d_small + slot * 8, k * 32, L.stream);
        copy_d2d(d + k * 8, d_w, n * 32, L.stream);
        return d;
    };
    // pi_A (lane 0), pi_B in G2 (lane 1), [H(s)]_1 (lane 2), pi_B' (lane 3), private part of pi_C (lane 4)
    size_t na = std::min(num_vars, pk->a_len), nb2 = std::min(num_vars, pk->b2_len), nb1 = std::min(num_vars, pk->b1_len);
    uint32_t *t_priv = d_cparts, *t_h = d_cparts + PW1, *t_mul = d_cparts + 2 * PW1;
    uint32_t *oa = LA.ws.out.as<uint32_t>(PW1 + AW1) + PW1;
    msm_run<Fq>(LA, pk->a->shards[0], prefixed(LA, 0, 2, na), na + 2, true, co, nullptr, oa);
    uint32_t *ob = LB.ws.out.as<uint32_t>(PW2 + AW2) + PW2;
    msm_run<Fq2>(LB, pk->b2->shards[0], prefixed(LB, 2, 2, nb2), nb2 + 2, true, co, nullptr, ob);
    size_t nh = (h || (dev && dev->d_h)) ? std::min(num_h, pk->h_len) : 0;
    if (dev && dev->d_h) {
        msm_run<Fq>(LH, pk->h->shards[0], dev->d_h, nh, true, co, t_h, nullptr);
    } else {
        uint32_t *d_h = LH.ws.scalars.as<uint32_t>(nh * 8 + 8);
        copy_h2d(d_h, h, nh * 32, LH.stream);
        msm_run<Fq>(LH, pk->h->shards[0], d_h, nh, true, co, t_h, nullptr);     // nh == 0 -> identity
    }
    uint32_t *ob1 = LB1.ws.out.as<uint32_t>(PW1 + AW1) + PW1;
    msm_run<Fq>(LB1, pk->b1->shards[0], prefixed(LB1, 4, 1, nb1), nb1 + 1, true, co, nullptr, ob1);
    size_t first_priv = pk->num_public + 1;
    size_t nic = num_vars > first_priv ? std::min(num_vars - first_priv, pk->ic_len) : 0;
    msm_run<Fq>(LC, pk->ic->shards[0], d_w + first_priv * 8, nic, true, co, t_priv, nullptr);
    // the rest of pi_C on lane 4 once pi_A, H and pi_B' exist (device-side dependency, no host wait)
    for (Device *l : {&LA, &LH, &LB1}) stream_wait(LC.stream, l->stream);
    copy_d2d(d_ab1, oa, AW1 * 4, LC.stream);
    copy_d2d(d_ab1 + AW1, ob1, AW1 * 4, LC.stream);
    k_scalar_mul_affine<Fq>(LC.stream, 2, d_small + 6 * 8, d_ab1, (uint32_t)AW1, t_mul);
    k_partial_combine<Fq>(LC.stream, d_cparts, 4, nullptr, d_c_aff);

    uint32_t ra[AW1], rb[AW2], rc[AW1];
    copy_d2h(ra, oa, AW1 * 4, LA.stream);
    copy_d2h(rb, ob, AW2 * 4, LB.stream);
    copy_d2h(rc, d_c_aff, AW1 * 4, LC.stream);
    if (dev && dev->d_flags && dev->flags_out) copy_d2h(dev->flags_out, dev->d_flags, 8, LA.stream);
    for (Device *l : {&LA, &LB, &LH, &LB1, &LC}) stream_sync(l->stream);
    memcpy(a_xy, ra, 96); memcpy(b_xy, rb, 192); memcpy(c_xy, rc, 96);
    if (a_inf) *a_inf = (uint8_t)ra[24];
    if (b_inf) *b_inf = (uint8_t)rb[48];
    if (c_inf) *c_inf = (uint8_t)rc[24];
}

// The same schedule on a multi-device context: every resident array is sharded by index range, so device k runs the
// five MSMs over ITS slices on five lanes; its slice of the assignment (and of the H coefficients) is copied once and
// shared by the four MSMs that read it -- no host-side concatenation, the (1, r) / (1, s) prefixes are spliced in on
// the device that owns the head of an array.  The 5 x G partial sums travel to device 0, which folds them, adds the
// ad-hoc terms of pi_C and returns the proof; the host waits once.
void prove_multi_device(Context *c, const g16_pk *pk, const uint64_t *w, size_t num_vars, const uint64_t *h, size_t num_h,
                        const uint64_t *r, const uint64_t *s, uint64_t *a_xy, uint8_t *a_inf, uint64_t *b_xy, uint8_t *b_inf,
                        uint64_t *c_xy, uint8_t *c_inf) {
    constexpr size_t PW1 = 48, AW1 = 25, PW2 = 96, AW2 = 49;
    const size_t G = c->devs.size();
    const unsigned co = c->c_override;
    std::vector<uint64_t> hs(4 * 16);
    auto put = [&](size_t slot, const uint64_t *x) { memcpy(hs.data() + 4 * slot, x, 32); };
    put(0, FR_ONE_MONT); put(1, r);              // pi_A prefix   [1, r]
    put(2, FR_ONE_MONT); put(3, s);              // pi_B prefix   [1, s]
    put(4, FR_ONE_MONT);                         // pi_B' prefix  [1]
    put(6, s); put(7, r);                        // multipliers of pi_A and pi_B' inside pi_C

    const size_t first_priv = pk->num_public + 1;
    struct Job { const Bases *bases; size_t prefix, slot, n_src, src_off; int lane; bool g2, from_h; };
    const Job jobs[5] = {
        {pk->a.get(), 2, 0, std::min(num_vars, pk->a_len), 0, 0, false, false},
        {pk->b2.get(), 2, 2, std::min(num_vars, pk->b2_len), 0, 1, true, false},
        {pk->h.get(), 0, 0, h ? std::min(num_h, pk->h_len) : 0, 0, 2, false, true},
        {pk->b1.get(), 1, 4, std::min(num_vars, pk->b1_len), 0, 3, false, false},
        {pk->ic.get(), 0, 0, num_vars > first_priv ? std::min(num_vars - first_priv, pk->ic_len) : 0, first_priv, 4, false, false},
    };
    // device 0, lane 4 collects: partial sums [job][device] (G2 slots for all, simpler indexing)
    Device &d0 = c->devs[0];
    set_device(d0.id);
    Device &LC0 = lane_of(d0, 4);
    uint32_t *parts = LC0.ws.partials.as<uint32_t>(5 * (G + 1) * PW2 + 4 * AW2 + 3 * 24 + 64);
    auto part_of = [&](int job, size_t k) { return parts + ((size_t)job * (G + 1) + k) * PW2; };

    for (size_t k = 0; k < G; ++k) {
        Device &dv = c->devs[k];
        set_device(dv.id);
        Device *L[5];
        for (int i = 0; i < 5; ++i) L[i] = &lane_of(dv, i);
        // slice of every job on this device, and the range of the assignment they read
        size_t lo[5], hi[5], wlo = ~(size_t)0, whi = 0;
        for (int j = 0; j < 5; ++j) {
            const Job &jb = jobs[j];
            size_t n_total = jb.prefix + jb.n_src;
            const BasesShard &sh = jb.bases->shards[k];
            lo[j] = std::min(sh.begin, n_total); hi[j] = std::min(sh.begin + sh.n, n_total);
            if (!jb.from_h && hi[j] > std::max(lo[j], jb.prefix)) {
                wlo = std::min(wlo, std::max(lo[j], jb.prefix) - jb.prefix + jb.src_off);
                whi = std::max(whi, hi[j] - jb.prefix + jb.src_off);
            }
        }
        if (whi <= wlo) { wlo = 0; whi = 0; }
        uint32_t *d_w = L[0]->ws.prove_w.as<uint32_t>((whi - wlo) * 8 + 8);
        copy_h2d(d_w, w + wlo * 4, (whi - wlo) * 32, L[0]->stream);
        uint32_t *d_small = L[0]->ws.prove_misc.as<uint32_t>(16 * 8 + 64);
        copy_h2d(d_small, hs.data(), 16 * 32, L[0]->stream);
        for (int i = 1; i < 5; ++i) stream_wait(L[i]->stream, L[0]->stream);
        for (int j = 0; j < 5; ++j) {
            const Job &jb = jobs[j];
            Device &ln = *L[jb.lane];
            const BasesShard &sh = jb.bases->shards[k];
            size_t cnt = hi[j] - lo[j];
            const size_t PW = jb.g2 ? PW2 : PW1;
            uint32_t *d_out = ln.ws.out.as<uint32_t>(PW2 + AW2);
            const uint32_t *d_sc = nullptr;
            if (cnt) {
                if (jb.from_h) {
                    uint32_t *d_h = ln.ws.scalars.as<uint32_t>(cnt * 8 + 8);
                    copy_h2d(d_h, h + lo[j] * 4, cnt * 32, ln.stream);
                    d_sc = d_h;
                } else if (lo[j] >= jb.prefix) {
                    d_sc = d_w + (lo[j] - jb.prefix + jb.src_off - wlo) * 8;   // a plain slice of the assignment
                } else {
                    // head of the array: prefix scalars, then the assignment from its first element
                    uint32_t *d = ln.ws.scalars.as<uint32_t>(cnt * 8 + 8);
                    size_t np = std::min(jb.prefix, hi[j]) - lo[j];
                    copy_d2d(d, d_small + (jb.slot + lo[j]) * 8, np * 32, ln.stream);
                    copy_d2d(d + np * 8, d_w + (jb.src_off - wlo) * 8, (cnt - np) * 32, ln.stream);
                    d_sc = d;
                }
            }
            if (jb.g2) msm_run<Fq2>(ln, sh, d_sc, cnt, true, co, d_out, nullptr, lo[j] - std::min(lo[j], sh.begin));
            else msm_run<Fq>(ln, sh, d_sc, cnt, true, co, d_out, nullptr, lo[j] - std::min(lo[j], sh.begin));
            copy_peer(part_of(j, k), d_out, PW * 4, ln.stream);
        }
    }
    // device 0: wait for every lane of every device, fold, finish pi_C
    for (size_t k = 0; k < G; ++k)
        for (int i = 0; i < 5; ++i) {
            Device &ln = lane_of(c->devs[k], i);
            if (&ln != &LC0) stream_wait_xdev(LC0.stream, d0.id, ln.stream, c->devs[k].id);
        }
    set_device(d0.id);
    uint32_t *aff = parts + 5 * (G + 1) * PW2;               // A, B (G2), H, B' affine results, AW2 words apart
    uint32_t *oa = aff, *ob = aff + AW2, *ob1 = aff + 3 * AW2;
    auto fold_g1 = [&](int job, uint32_t *out_aff) {
        // G1 partials sit PW2 words apart: compact them in place before the fold (slot 0 stays)
        for (size_t k = 1; k < G; ++k) copy_d2d(part_of(job, 0) + k * PW1, part_of(job, k), PW1 * 4, LC0.stream);
        k_partial_combine<Fq>(LC0.stream, part_of(job, 0), (uint32_t)G, nullptr, out_aff);
    };
    fold_g1(0, oa);
    fold_g1(3, ob1);
    // s * pi_A and r * pi_B' (one launch, two chains side by side in one warp), then the G2 fold and the H fold
    uint32_t *terms = part_of(4, 0);                          // [private part x G | H | s pi_A | r pi_B']
    const uint32_t *d_small0 = (const uint32_t *)lane_of(d0, 0).ws.prove_misc.p;
    for (size_t k = 1; k < G; ++k) copy_d2d(terms + k * PW1, part_of(4, k), PW1 * 4, LC0.stream);
    k_scalar_mul_affine<Fq>(LC0.stream, 2, d_small0 + 6 * 8, oa, (uint32_t)(3 * AW2), terms + (G + 1) * PW1);
    k_partial_combine<Fq2>(LC0.stream, part_of(1, 0), (uint32_t)G, nullptr, ob);
    for (size_t k = 1; k < G; ++k) copy_d2d(part_of(2, 0) + k * PW1, part_of(2, k), PW1 * 4, LC0.stream);
    k_partial_combine<Fq>(LC0.stream, part_of(2, 0), (uint32_t)G, terms + G * PW1, nullptr);
    uint32_t *d_c_aff = aff + 4 * AW2;
    k_partial_combine<Fq>(LC0.stream, terms, (uint32_t)G + 3, nullptr, d_c_aff);

    uint32_t ra[AW1], rb[AW2], rc[AW1];
    copy_d2h(ra, oa, AW1 * 4, LC0.stream);
    copy_d2h(rb, ob, AW2 * 4, LC0.stream);
    copy_d2h(rc, d_c_aff, AW1 * 4, LC0.stream);
    stream_sync(LC0.stream);
    memcpy(a_xy, ra, 96); memcpy(b_xy, rb, 192); memcpy(c_xy, rc, 96);
    if (a_inf) *a_inf = (uint8_t)ra[24];
    if (b_inf) *b_inf = (uint8_t)rb[48];
    if (c_inf) *c_inf = (uint8_t)rc[24];
}
}  // extern "C++"

// The MSM schedule of Prover::prove (crates/groth16-core/src/lib.rs:164-271).  The reference
// builds fresh (scalar, point) lists with zero scalars filtered out; here the CRS arrays stay
// resident, the unfiltered scalar vectors are sent (a zero scalar contributes no digit) and the
// ad-hoc terms (alpha, r*delta, ...) ride along as extra entries -- the sums are the same group
// elements.
int g16_prove(g16_ctx *ctx, const g16_pk *pk, const uint64_t *assignment_fr, size_t num_vars, const uint64_t *h_coeffs,
              size_t num_h, const uint64_t r[4], const uint64_t s[4], uint64_t a_xy[12], uint8_t *a_inf,
              uint64_t b_xy[24], uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(pk && pk->ctx == &ctx->c, "bad pk handle");
        require(assignment_fr && r && s && a_xy && b_xy && c_xy, "NULL argument");
        require(num_vars > pk->num_public, "assignment shorter than the public inputs");
        Context *c = &ctx->c;
        if (c->devs.size() == 1) {
            prove_single_device(c, pk, assignment_fr, num_vars, h_coeffs, num_h, r, s, a_xy, a_inf, b_xy, b_inf, c_xy, c_inf, nullptr);
            return;
        }
        prove_multi_device(c, pk, assignment_fr, num_vars, h_coeffs, num_h, r, s, a_xy, a_inf, b_xy, b_inf, c_xy, c_inf);
    });
}

// ---- quotient polynomial -----------------------------------------------------------------------
int g16_quotient_h(g16_ctx *ctx, const uint64_t *a_evals, const uint64_t *b_evals, const uint64_t *c_evals, size_t n,
                   uint64_t *h_coeffs) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(a_evals && b_evals && c_evals && h_coeffs, "NULL argument");
        require(n >= 1 && (n & (n - 1)) == 0 && n <= ((size_t)1 << 28), "domain size must be a power of two <= 2^28");
        uint32_t log_n = 0;
        while (((size_t)1 << log_n) < n) ++log_n;
        Device &dv = ctx->c.devs[0];
        set_device(dv.id);
        if (!quotient_host(dv, a_evals, b_evals, c_evals, log_n, h_coeffs))
            throw Error{G16_ERR_INVALID, "Polynomial division failed: non-zero remainder"};
    });
}

int g16_quotient_h_device(g16_ctx *ctx, void *dev_abc, size_t n, void *dev_h, void *dev_bad_rows) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(dev_abc && dev_h && dev_bad_rows, "NULL argument");
        require(n >= 1 && (n & (n - 1)) == 0 && n <= ((size_t)1 << 28), "domain size must be a power of two <= 2^28");
        uint32_t log_n = 0;
        while (((size_t)1 << log_n) < n) ++log_n;
        dev_memset(dev_bad_rows, 0, 4, dv.stream);
        quotient_device(dv, log_n, (uint32_t *)dev_abc, (uint32_t *)dev_bad_rows, (uint32_t *)dev_h);
    });
}

// ---- test hooks --------------------------------------------------------------------------------
unsigned long long g16_launch_count(void) { return launch_count(); }

int g16_ctx_set_item_max(g16_ctx *ctx, unsigned item_max) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(item_max == 0 || (item_max >= 4 && item_max <= k_item_max()), "item_max must be 0 or in [4, 256]");
        for (auto &d : ctx->c.devs) {
            d.item_max_override = item_max;
            for (auto &l : d.extra) l->item_max_override = item_max;
        }
    });
}
int g16_ctx_enable_stage_timing(g16_ctx *ctx, int on) {
    if (!ctx) return G16_ERR_INVALID;
    for (auto &d : ctx->c.devs) {
        d.timer.enabled = on != 0; d.timer.valid = false;
        for (auto &l : d.extra) { l->timer.enabled = false; l->timer.valid = false; }   // lanes: switched on per prove
    }
    return G16_OK;
}
int g16_ctx_last_stage_ms(g16_ctx *ctx, float ms[6], unsigned plan[3]) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        stream_sync(dv.stream);
        require(dv.timer.read(ms), "no timed MSM has run (g16_ctx_enable_stage_timing)");
        if (plan) { plan[0] = dv.last_plan.c; plan[1] = dv.last_plan.nwin; plan[2] = dv.last_plan.nb; }
    });
}

int g16_ctx_prove_timeline(g16_ctx *ctx, float t[35]) {
    if (!ctx || !t) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &d0 = single_device(ctx);
        for (int lane = 0; lane < 5; ++lane) {
            Device &l = lane_of(d0, lane);
            stream_sync(l.stream);
            require(l.timer.read_since(ctx->c.prove_epoch, t + 7 * lane), "no timed prove has run (g16_ctx_enable_stage_timing)");
        }
    });
}
int g16_debug_fq_op(g16_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(a && out, "NULL argument");
        uint32_t *d = dv.ws.staging.as<uint32_t>(n * 36 + 36);
        copy_h2d(d, a, n * 48, dv.stream);
        if (b) copy_h2d(d + n * 12, b, n * 48, dv.stream);
        k_debug_fq_op(dv.stream, n, op, d, b ? d + n * 12 : nullptr, d + n * 24);
        copy_d2h(out, d + n * 24, n * 48, dv.stream);
        stream_sync(dv.stream);
    });
}
int g16_debug_fr_from_mont(g16_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(a && out, "NULL argument");
        uint32_t *d = dv.ws.staging.as<uint32_t>(n * 16 + 16);
        copy_h2d(d, a, n * 32, dv.stream);
        k_debug_fr_from_mont(dv.stream, n, d, d + n * 8);
        copy_d2h(out, d + n * 8, n * 32, dv.stream);
        stream_sync(dv.stream);
    });
}
extern "C++" {
template <class F>
static int debug_add_impl(g16_ctx *ctx, const uint64_t *p, const uint8_t *p_inf, const uint64_t *q, const uint8_t *q_inf,
                          uint64_t *out_xy, uint8_t *out_inf, size_t n) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        Device &dv = single_device(ctx);
        require(p && q && out_xy, "NULL argument");
        constexpr size_t W = 2 * FieldWords<F>::N;
        uint32_t *dp = import_points<F>(dv, p, p_inf, n);
        uint32_t *dq = nullptr;
        try {
            dq = import_points<F>(dv, q, q_inf, n);
            uint32_t *d_out = dv.ws.fb_out.as<uint32_t>(n * W);
            k_debug_add<F>(dv.stream, n, dp, dq, d_out);
            copy_d2h(out_xy, d_out, n * W * 4, dv.stream);
            if (out_inf) {
                uint8_t *d_fl = dv.ws.fb_flags.as<uint8_t>(n);
                k_export_flags<F>(dv.stream, n, d_out, d_fl);
                copy_d2h(out_inf, d_fl, n, dv.stream);
            }
            stream_sync(dv.stream);
        } catch (...) { dev_free(dp); dev_free(dq); throw; }
        dev_free(dp); dev_free(dq);
    });
}
}  // extern "C++"
int g16_debug_g1_add(g16_ctx *ctx, const uint64_t *p, const uint8_t *p_inf, const uint64_t *q, const uint8_t *q_inf,
                     uint64_t *out_xy, uint8_t *out_inf, size_t n) {
    return debug_add_impl<Fq>(ctx, p, p_inf, q, q_inf, out_xy, out_inf, n);
}
int g16_debug_g2_add(g16_ctx *ctx, const uint64_t *p, const uint8_t *p_inf, const uint64_t *q, const uint8_t *q_inf,
                     uint64_t *out_xy, uint8_t *out_inf, size_t n) {
    return debug_add_impl<Fq2>(ctx, p, p_inf, q, q_inf, out_xy, out_inf, n);
}

}  // extern "C"
