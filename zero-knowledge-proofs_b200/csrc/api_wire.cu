// C ABI of the wire format (include/g16_cuda.h, "wire format"): ark CanonicalSerialize / CanonicalDeserialize of
// G1Affine / G2Affine vectors and of `Proof` (crates/groth16-core/src/lib.rs:27-36).  Also compiled with
// -DG16_EMU by tests/emu.
#include "api_common.cuh"

namespace {

template <class F>
void serialize_points(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
    Device &dv = single_device(ctx);
    require((xy && out) || n == 0, "NULL argument");
    if (!n) return;
    constexpr size_t PW = 2 * FieldWords<F>::N;           // words per packed point
    const size_t out_words = n * (compressed ? PW / 2 : PW);
    uint32_t *pts = import_points<F>(dv, xy, inf, n);
    try {
        uint32_t *d_out = dv.ws.out.as<uint32_t>(out_words);
        k_point_encode<F>(dv.stream, n, pts, compressed != 0, d_out);
        copy_d2h(out, d_out, out_words * 4, dv.stream);
        stream_sync(dv.stream);
    } catch (...) { dev_free(pts); throw; }
    dev_free(pts);
}

// returns the number of rejected elements; first_bad = index of the first one
template <class F>
size_t deserialize_points(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                          uint8_t *out_inf, uint8_t *status, size_t &first_bad, uint8_t &first_code) {
    Device &dv = single_device(ctx);
    require((bytes && out_xy) || n == 0, "NULL argument");
    first_bad = 0; first_code = 0;
    if (!n) return 0;
    constexpr size_t PW = 2 * FieldWords<F>::N;
    const size_t in_words = n * (compressed ? PW / 2 : PW);
    uint32_t *d_in = dv.ws.staging.as<uint32_t>(in_words + n * PW + (n + 3) / 4 + 4);
    uint32_t *d_pts = d_in + in_words;
    uint8_t *d_status = (uint8_t *)(d_pts + n * PW);
    copy_h2d(d_in, bytes, in_words * 4, dv.stream);
    k_point_decode<F>(dv.stream, n, d_in, compressed != 0, validate != 0, d_pts, d_status);
    copy_d2h(out_xy, d_pts, n * PW * 4, dv.stream);
    std::vector<uint8_t> st(n);
    copy_d2h(st.data(), d_status, n, dv.stream);
    if (out_inf) {
        uint8_t *d_fl = dv.ws.fb_flags.as<uint8_t>(n);
        k_export_flags<F>(dv.stream, n, d_pts, d_fl);
        copy_d2h(out_inf, d_fl, n, dv.stream);
    }
    stream_sync(dv.stream);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        if (st[i]) {
            if (!bad) { first_bad = i; first_code = st[i]; }
            ++bad;
        }
    }
    if (status) memcpy(status, st.data(), n);
    return bad;
}

template <class F>
int deserialize_entry(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                      uint8_t *out_inf, uint8_t *status) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        size_t first = 0;
        uint8_t code = 0;
        size_t bad = deserialize_points<F>(ctx, bytes, n, compressed, validate, out_xy, out_inf, status, first, code);
        if (bad) {
            throw Error{G16_ERR_INVALID, std::string(code == WIRE_STATUS_UNEXPECTED_FLAGS ? "UnexpectedFlags" : "InvalidData") +
                                             " at element " + std::to_string(first) + " (" + std::to_string(bad) + " rejected)"};
        }
    });
}

}  // namespace

extern "C" {

int g16_g1_serialize(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] { serialize_points<Fq>(ctx, xy, inf, n, compressed, out); });
}
int g16_g2_serialize(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] { serialize_points<Fq2>(ctx, xy, inf, n, compressed, out); });
}
int g16_g1_deserialize(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                       uint8_t *out_inf, uint8_t *status) {
    return deserialize_entry<Fq>(ctx, bytes, n, compressed, validate, out_xy, out_inf, status);
}
int g16_g2_deserialize(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                       uint8_t *out_inf, uint8_t *status) {
    return deserialize_entry<Fq2>(ctx, bytes, n, compressed, validate, out_xy, out_inf, status);
}

int g16_proof_serialize(g16_ctx *ctx, const uint64_t a_xy[12], uint8_t a_inf, const uint64_t b_xy[24], uint8_t b_inf,
                        const uint64_t c_xy[12], uint8_t c_inf, int compressed, uint8_t *out) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(a_xy && b_xy && c_xy && out, "NULL argument");
        const size_t g1 = compressed ? 48 : 96, g2 = 2 * g1;
        uint64_t ac[24];
        uint8_t ac_inf[2] = {a_inf, c_inf};
        memcpy(ac, a_xy, 96); memcpy(ac + 12, c_xy, 96);
        uint8_t tmp[192];
        serialize_points<Fq>(ctx, ac, ac_inf, 2, compressed, tmp);
        memcpy(out, tmp, g1);
        memcpy(out + g1 + g2, tmp + g1, g1);
        serialize_points<Fq2>(ctx, b_xy, &b_inf, 1, compressed, out + g1);
    });
}
int g16_proof_deserialize(g16_ctx *ctx, const uint8_t *bytes, int compressed, int validate, uint64_t a_xy[12], uint8_t *a_inf,
                          uint64_t b_xy[24], uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf) {
    if (!ctx) return G16_ERR_INVALID;
    return guarded(ctx, [&] {
        require(bytes && a_xy && b_xy && c_xy, "NULL argument");
        const size_t g1 = compressed ? 48 : 96, g2 = 2 * g1;
        uint8_t tmp[192], inf2[2], infb = 0;
        uint64_t ac[24];
        memcpy(tmp, bytes, g1);
        memcpy(tmp + g1, bytes + g1 + g2, g1);
        size_t first = 0;
        uint8_t code = 0;
        const char *field = nullptr;
        if (deserialize_points<Fq>(ctx, tmp, 2, compressed, validate, ac, inf2, nullptr, first, code)) field = first ? "c" : "a";
        uint8_t code_b = 0;
        size_t fb = 0;
        // ark reads a, b, c in order and stops at the first error
        if (deserialize_points<Fq2>(ctx, bytes + g1, 1, compressed, validate, b_xy, &infb, nullptr, fb, code_b) &&
            (!field || field[0] == 'c')) { field = "b"; code = code_b; }
        if (field)
            throw Error{G16_ERR_INVALID, std::string(code == WIRE_STATUS_UNEXPECTED_FLAGS ? "UnexpectedFlags" : "InvalidData") +
                                             " in proof." + field};
        memcpy(a_xy, ac, 96); memcpy(c_xy, ac + 12, 96);
        if (a_inf) *a_inf = inf2[0];
        if (c_inf) *c_inf = inf2[1];
        if (b_inf) *b_inf = infb;
    });
}

}  // extern "C"
