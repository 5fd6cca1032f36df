// Host side of the sparse R1CS path: device-resident constraint matrices, the domain
// evaluations that feed the quotient polynomial, the QAP evaluations at s and the CRS exponents that feed
// the fixed-base kernels.  Allocation, copies and launches only -- the arithmetic is in r1cs_kernels.cuh.
#pragma once
#include "engine.cuh"

namespace g16 {

// ark-bls12-381 generators, Montgomery limbs (`G1Projective::generator()` / `G2Projective::generator()`,
// /root/reference/crates/groth16-setup/src/lib.rs:162-163); pinned by tests/test_oracle_kat.py
static const uint64_t G1_GENERATOR[12] = {
    0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL, 0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL,
    0x120177419e0bfb75ULL, 0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL, 0x51ac582950405194ULL,
    0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL};
static const uint64_t G2_GENERATOR[24] = {
    0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL, 0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL,
    0x058191924350bcd7ULL, 0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL, 0x1b1ab6cc8541b367ULL,
    0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL, 0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL,
    0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL, 0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL,
    0x86adac6a3be4eba0ULL, 0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL};

// one orientation of the stacked matrices (A | B | C) on the device
struct StackedCsr {
    uint32_t *ptr = nullptr, *idx = nullptr, *val = nullptr, *long_lines = nullptr;
    size_t lines = 0, nnz = 0, n_long = 0;
    void release() {
        dev_free(ptr); dev_free(idx); dev_free(val); dev_free(long_lines);
        ptr = idx = val = long_lines = nullptr;
    }
};

struct R1cs {
    Context *ctx = nullptr;        // identity check only; never dereferenced after upload
    int dev = 0;
    int cuda_dev = 0;              // CUDA ordinal of the device that holds the matrices
    size_t m = 0, nv = 0;          // constraints, variables
    uint32_t log_n = 0;            // domain = next_power_of_two(m)  (`QAP::from_r1cs`, qap/src/lib.rs:100)
    StackedCsr rows, cols;         // by constraint (prove) / by variable (setup)
    ~R1cs() {
        set_device_nothrow(cuda_dev);
        rows.release(); cols.release();
    }
};

struct HostCsr {
    std::vector<uint32_t> ptr, idx;
    std::vector<uint64_t> val;   // 4 u64 per entry
};

inline void upload_stacked(Device &dv, const HostCsr &h, StackedCsr &d) {
    d.lines = h.ptr.size() - 1;
    d.nnz = h.idx.size();
    std::vector<uint32_t> longs;
    uint32_t thr = k_spmv_long_threshold();
    for (size_t t = 0; t < d.lines; ++t)
        if (h.ptr[t + 1] - h.ptr[t] > thr) longs.push_back((uint32_t)t);
    d.n_long = longs.size();
    d.ptr = (uint32_t *)dev_alloc(h.ptr.size() * 4);
    d.idx = (uint32_t *)dev_alloc(d.nnz * 4);
    d.val = (uint32_t *)dev_alloc(d.nnz * 32);
    d.long_lines = (uint32_t *)dev_alloc(longs.size() * 4);
    copy_h2d(d.ptr, h.ptr.data(), h.ptr.size() * 4, dv.stream);
    copy_h2d(d.idx, h.idx.data(), d.nnz * 4, dv.stream);
    copy_h2d(d.val, h.val.data(), d.nnz * 32, dv.stream);
    copy_h2d(d.long_lines, longs.data(), longs.size() * 4, dv.stream);
    stream_sync(dv.stream);
}

// mats[k] = {row_ptr (m + 1), col (nnz_k), val (nnz_k x 4 u64)} for k = A, B, C.  Entries whose variable index
// is >= num_variables are dropped like the reference does (qap/src/lib.rs:121-138).  When a (row, variable) pair
// appears more than once in a matrix the LAST value wins, as in the reference's `a_evals[row][var] = coeff`
// assignment loop (qap/src/lib.rs:121-138) -- the SpMV kernels would otherwise sum the duplicates.
struct CsrView { const uint32_t *row_ptr, *col; const uint64_t *val; };
inline std::unique_ptr<R1cs> r1cs_upload(Context *ctx, size_t m, size_t nv, const CsrView mats[3]) {
    if (nv == 0) throw Error{G16_ERR_INVALID, "R1CS needs at least the constant variable"};
    if (3 * m + 1 >= 0xffffffffull || 3 * nv + 1 >= 0xffffffffull) throw Error{G16_ERR_INVALID, "R1CS too large"};
    std::unique_ptr<R1cs> r(new R1cs);
    r->ctx = ctx; r->dev = 0; r->m = m; r->nv = nv;
    size_t n = 1;
    while (n < m) { n <<= 1; ++r->log_n; }
    if (r->log_n > 28) throw Error{G16_ERR_INVALID, "domain larger than 2^28"};
    HostCsr by_row, by_col;
    by_row.ptr.assign(3 * m + 1, 0);
    by_col.ptr.assign(3 * nv + 1, 0);
    // pass 1: count the kept entries per row line and per column line.  seen[var] = 1 + line of the last row that
    // mentioned var: a repeated (row, variable) pair is counted once
    std::vector<uint64_t> seen(nv, 0);
    uint64_t kept = 0;
    for (int k = 0; k < 3; ++k) {
        const CsrView &v = mats[k];
        if (m && (!v.row_ptr || v.row_ptr[0] != 0)) throw Error{G16_ERR_INVALID, "row_ptr must start at 0"};
        for (size_t i = 0; i < m; ++i) {
            if (v.row_ptr[i + 1] < v.row_ptr[i]) throw Error{G16_ERR_INVALID, "row_ptr must be non-decreasing"};
            const uint64_t line = (uint64_t)k * m + i + 1;
            for (uint32_t e = v.row_ptr[i]; e < v.row_ptr[i + 1]; ++e) {
                uint32_t c = v.col[e];
                if (c >= nv || seen[c] == line) continue;
                seen[c] = line;
                ++by_row.ptr[k * m + i + 1];
                ++by_col.ptr[k * nv + c + 1];
                ++kept;
            }
        }
    }
    if (kept >= 0xffffffffull) throw Error{G16_ERR_INVALID, "more than 2^32 non-zero coefficients"};
    for (size_t t = 0; t < 3 * m; ++t) by_row.ptr[t + 1] += by_row.ptr[t];
    for (size_t t = 0; t < 3 * nv; ++t) by_col.ptr[t + 1] += by_col.ptr[t];
    by_row.idx.resize(kept); by_row.val.resize(kept * 4);
    by_col.idx.resize(kept); by_col.val.resize(kept * 4);
    std::vector<uint32_t> cur(by_col.ptr.begin(), by_col.ptr.end() - 1);
    // pass 2: fill; slot_row / slot_col remember where the pair of the current line went, so that a later duplicate
    // overwrites the value in place
    std::fill(seen.begin(), seen.end(), 0);
    std::vector<uint32_t> slot_row(nv, 0), slot_col(nv, 0);
    size_t w = 0;
    for (int k = 0; k < 3; ++k) {
        const CsrView &v = mats[k];
        for (size_t i = 0; i < m; ++i) {
            const uint64_t line = (uint64_t)k * m + i + 1;
            for (uint32_t e = v.row_ptr[i]; e < v.row_ptr[i + 1]; ++e) {
                uint32_t c = v.col[e];
                if (c >= nv) continue;
                if (seen[c] == line) {
                    memcpy(&by_row.val[4 * (size_t)slot_row[c]], v.val + 4 * (size_t)e, 32);
                    memcpy(&by_col.val[4 * (size_t)slot_col[c]], v.val + 4 * (size_t)e, 32);
                    continue;
                }
                seen[c] = line;
                by_row.idx[w] = c;
                memcpy(&by_row.val[4 * w], v.val + 4 * (size_t)e, 32);
                slot_row[c] = (uint32_t)w;
                ++w;
                uint32_t p = cur[k * nv + c]++;
                by_col.idx[p] = (uint32_t)i;
                memcpy(&by_col.val[4 * (size_t)p], v.val + 4 * (size_t)e, 32);
                slot_col[c] = p;
            }
        }
    }
    Device &dv = ctx->devs[0];
    set_device(dv.id);
    r->cuda_dev = dv.id;
    upload_stacked(dv, by_row, r->rows);
    upload_stacked(dv, by_col, r->cols);
    return r;
}

// abc (device, 3 * n Fr) = evaluations of A, B, C on the domain for the assignment d_w (device, nv Fr)
inline void r1cs_domain_evals_device(Device &dv, const R1cs &r, const uint32_t *d_w, uint32_t *abc) {
    size_t n = (size_t)1 << r.log_n;
    dev_memset(abc, 0, 3 * n * 32, dv.stream);   // rows beyond the constraints evaluate to zero
    if (r.m)
        k_spmv(dv.stream, 3 * r.m, r.rows.ptr, r.rows.idx, r.rows.val, d_w, (uint32_t)r.m, (uint32_t)n, r.rows.long_lines,
               r.rows.n_long, abc);
}

// vals (device, 3 * nv Fr) = A_j(s_t), B_j(s_t), C_j(s_t) for the truncated s of the setup block `blk`
inline void r1cs_eval_at_device(Device &dv, const R1cs &r, const uint32_t *blk, uint32_t *lag, uint32_t *vals) {
    size_t n = (size_t)1 << r.log_n;
    const uint32_t *consts = ntt_prepare(dv, r.log_n);
    k_lagrange_at(dv.stream, n, consts, blk, lag);
    k_spmv(dv.stream, 3 * r.nv, r.cols.ptr, r.cols.idx, r.cols.val, lag, (uint32_t)r.nv, (uint32_t)r.nv, r.cols.long_lines,
           r.cols.n_long, vals);
}

}  // namespace g16
