// Definitions shared by the translation units that implement the C ABI (api.cu, api_r1cs.cu).
#pragma once
#include "../../include/g16_cuda.h"
#include "engine.cuh"

using namespace g16;

struct g16_ctx { Context c; };
struct g16_bases { std::unique_ptr<Bases> b; };

struct g16_pk {
    Context *ctx = nullptr;
    // resident arrays with the ad-hoc single points prepended (see g16_prove)
    std::unique_ptr<Bases> a;    // [alpha_g1, delta_g1, a_g1...]
    std::unique_ptr<Bases> b2;   // [beta_g2, delta_g2, b_g2...]
    std::unique_ptr<Bases> b1;   // [beta_g1, b_g1...]
    std::unique_ptr<Bases> ic;   // ic_g1
    std::unique_ptr<Bases> h;    // h_g1
    size_t a_len = 0, b1_len = 0, b2_len = 0, ic_len = 0, h_len = 0, num_public = 0;
};

inline std::string &create_error() { static thread_local std::string e; return e; }
static const uint64_t FR_ONE_MONT[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                                        0x1824b159acc5056fULL};


template <class Fn>
inline int guarded(g16_ctx *ctx, Fn &&fn) {
    try {
        fn();
        return G16_OK;
    } catch (const Error &e) {
        if (ctx) ctx->c.err = e.msg; else create_error() = e.msg;
        return e.code;
    } catch (const std::bad_alloc &) {
        if (ctx) ctx->c.err = "host allocation failed"; else create_error() = "host allocation failed";
        return G16_ERR_OOM;
    } catch (...) {
        if (ctx) ctx->c.err = "unknown error"; else create_error() = "unknown error";
        return G16_ERR_INVALID;
    }
}

inline void require(bool ok, const char *what) {
    if (!ok) throw Error{G16_ERR_INVALID, what};
}

inline Device &single_device(g16_ctx *ctx) {
    require(ctx->c.devs.size() == 1, "this entry point needs a single-device context");
    set_device(ctx->c.devs[0].id);
    return ctx->c.devs[0];
}


// inputs of the prove schedule that already live on the device (written on lane 0's stream): the truncated
// assignment, the truncated H coefficients and two flag words read back together with the proof
struct ProveDeviceInputs {
    const uint32_t *d_w = nullptr, *d_h = nullptr, *d_flags = nullptr;
    uint32_t *flags_out = nullptr;
};
void prove_single_device(Context *c, const g16_pk *pk, const uint64_t *w, size_t num_vars, const uint64_t *h, size_t num_h,
                         const uint64_t *r, const uint64_t *s, uint64_t *a_xy, uint8_t *a_inf, uint64_t *b_xy, uint8_t *b_inf,
                         uint64_t *c_xy, uint8_t *c_inf, const ProveDeviceInputs *dev);
