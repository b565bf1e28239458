// SRS handling: `.ptau` loader (replaces readBinFile + readPTauHeader + fd.readToBuffer, reference
// src/grandsum/mset_eq_kzg_prover.js:15-16,83-85 and src/ptau_utils.js:3-24), a device generator of
// synthetic SRS [tau^i]_1 (the Hermez file of .github/workflows/tests.yml:15-19 cannot be fetched
// offline) and the matching `.ptau` writer.  Container layout:
//   "ptau" | version u32 | nSections u32 | { id u32 | size u64 | payload }*
//   section 1: n8 u32 | q (n8 B LE) | power u32 | ceremonyPower u32
//   section 2: tauG1, 64 B affine Montgomery-LE per point;  section 3: tauG2, 128 B per point
#include <stdio.h>
#include <string.h>

#include <map>

#include "common.cuh"

namespace kzg {

struct PtauSection {
    uint64_t offset, size;
};

static int ptau_scan(kzg_ctx* ctx, FILE* f, const char* path, std::map<uint32_t, std::vector<PtauSection>>& sections) {
    uint8_t hdr[12];
    if (fread(hdr, 1, 12, f) != 12) return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": Invalid File format");
    if (memcmp(hdr, "ptau", 4) != 0) return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": Invalid File format");
    uint32_t version, nsec;
    memcpy(&version, hdr + 4, 4);
    memcpy(&nsec, hdr + 8, 4);
    if (version > 1) return set_err(ctx, KZG_ERR_FORMAT, "Version not supported");
    uint64_t pos = 12;
    for (uint32_t i = 0; i < nsec; i++) {
        uint8_t sh[12];
        if (fseeko(f, (off_t)pos, SEEK_SET) != 0 || fread(sh, 1, 12, f) != 12)
            return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": truncated section table");
        uint32_t id;
        uint64_t size;
        memcpy(&id, sh, 4);
        memcpy(&size, sh + 4, 8);
        sections[id].push_back({pos + 12, size});
        pos += 12 + size;
    }
    return KZG_OK;
}

static const uint8_t BN254_Q_LE[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71, 0x68, 0x91, 0x6a, 0x81, 0x97,
                                       0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

static int ptau_header(kzg_ctx* ctx, FILE* f, const char* path, std::map<uint32_t, std::vector<PtauSection>>& sections,
                       uint32_t* power, uint32_t* ceremony_power) {
    if (!sections.count(1)) return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": File has no  header");
    if (sections[1].size() > 1) return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": File has more than one header");
    const PtauSection s = sections[1][0];
    uint8_t buf[44];
    if (fseeko(f, (off_t)s.offset, SEEK_SET) != 0 || fread(buf, 1, 4, f) != 4)
        return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": truncated header");
    uint32_t n8;
    memcpy(&n8, buf, 4);
    if (n8 != 32) return set_err(ctx, KZG_ERR_FORMAT, "Curve not supported");  // getCurveFromQ: only BN254 here
    if (fread(buf + 4, 1, 40, f) != 40) return set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": truncated header");
    if (memcmp(buf + 4, BN254_Q_LE, 32) != 0) return set_err(ctx, KZG_ERR_FORMAT, "Curve not supported");
    memcpy(power, buf + 36, 4);
    memcpy(ceremony_power, buf + 40, 4);
    if (s.size != 44) return set_err(ctx, KZG_ERR_FORMAT, "Invalid PTau header size");
    return KZG_OK;
}

// ---- device SRS generator -------------------------------------------------------------------------
// table[w][d-1] = d * 2^(8w) * G1 (XYZZ), w < 32, d in 1..255
__global__ void srs_table_kernel(G1XYZZ* __restrict__ table) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32) return;
    G1Affine g;
    g.x = fp_one<FqP>();
    g.y = fp_dbl(fp_one<FqP>());
    G1XYZZ base = xyzz_from_affine(g);
    for (uint32_t i = 0; i < 8 * w; i++) base = xyzz_dbl(base);
    G1XYZZ acc = base;
    for (uint32_t d = 1; d <= 255; d++) {
        G1XYZZ* slot = table + (size_t)w * 255 + (d - 1);
        fp_store(&slot->x, acc.x);
        fp_store(&slot->y, acc.y);
        fp_store(&slot->zz, acc.zz);
        fp_store(&slot->zzz, acc.zzz);
        xyzz_add(acc, base);
    }
}

__global__ void __launch_bounds__(128) srs_points_kernel(const G1XYZZ* __restrict__ table, Fr tau, uint64_t first, uint64_t n,
                                                         G1Affine* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = fp_from_mont(fp_pow_u64(tau, first + i));  // tau^(first+i) as a plain integer
    G1XYZZ acc = xyzz_inf();
    for (uint32_t w = 0; w < 32; w++) {
        uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xffu;
        if (d) {
            const G1XYZZ* slot = table + (size_t)w * 255 + (d - 1);
            G1XYZZ t;
            t.x = fp_load<FqP>(&slot->x);
            t.y = fp_load<FqP>(&slot->y);
            t.zz = fp_load<FqP>(&slot->zz);
            t.zzz = fp_load<FqP>(&slot->zzz);
            xyzz_add(acc, t);
        }
    }
    G1Affine a = xyzz_to_affine(acc);
    fp_store(&out[i].x, a.x);
    fp_store(&out[i].y, a.y);
}

// ---------------------------------------------------------------------------------------------
// Lagrange-basis SRS:  [L_i(tau)]_1 = (1/n) sum_j w^(-ij) [tau^j]_1  -- the inverse DFT of the
// monomial points, carried out in the GROUP: radix-2 DIF stages over extended-Jacobian points, one thread per
// butterfly, the twiddle applied as a 254-bit double-and-add scalar multiplication (~4 k Montgomery products per
// butterfly; n log n / 2 of them: 0.6 s at n = 2^20 -- a one-off per SRS, like the window table).  With it a polynomial
// given by its EVALUATIONS on H is committed directly: commit(lagrange_srs, evals) == commit(srs, iNTT(evals)).
// ---------------------------------------------------------------------------------------------
__device__ G1XYZZ g1_scalar_mul(const G1XYZZ& p, const Fr& k_std) {
    // k_std: standard-form integer < r.  MSB-first double-and-add; real loops around noinline-sized bodies
    G1XYZZ acc = xyzz_inf();
    int top = 7;
    while (top >= 0 && k_std.l[top] == 0) top--;
    if (top < 0 || xyzz_is_inf(p)) return acc;
#pragma unroll 1
    for (int limb = top; limb >= 0; limb--) {
        const uint32_t w = k_std.l[limb];
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            acc = xyzz_dbl(acc);
            if ((w >> bit) & 1u) xyzz_add(acc, p);
        }
    }
    return acc;
}
__global__ void __launch_bounds__(128) g1_lift_kernel(const G1Affine* __restrict__ in, uint64_t n, G1XYZZ* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    p.x = fp_load<FqP>(&in[i].x);
    p.y = fp_load<FqP>(&in[i].y);
    const G1XYZZ v = xyzz_from_affine(p);
    fp_store(&out[i].x, v.x);
    fp_store(&out[i].y, v.y);
    fp_store(&out[i].zz, v.zz);
    fp_store(&out[i].zzz, v.zzz);
}
__device__ __forceinline__ G1XYZZ load_pt(const G1XYZZ* p) {
    G1XYZZ v;
    v.x = fp_load<FqP>(&p->x);
    v.y = fp_load<FqP>(&p->y);
    v.zz = fp_load<FqP>(&p->zz);
    v.zzz = fp_load<FqP>(&p->zzz);
    return v;
}
__device__ __forceinline__ void store_pt(G1XYZZ* p, const G1XYZZ& v) {
    fp_store(&p->x, v.x);
    fp_store(&p->y, v.y);
    fp_store(&p->zz, v.zz);
    fp_store(&p->zzz, v.zzz);
}
// one DIF stage of half-size h with the INVERSE twiddles: (a, b) -> (a + b, (a - b) w_{2h}^-j)
__global__ void __launch_bounds__(128) g1_dif_stage_kernel(G1XYZZ* __restrict__ pts, uint64_t n, uint64_t h, uint32_t log_2h,
                                                           const Fr* __restrict__ tw_lo, const Fr* __restrict__ tw_hi) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n / 2) return;
    const uint64_t j = b & (h - 1);
    const uint64_t i0 = ((b - j) << 1) + j, i1 = i0 + h;
    G1XYZZ a = load_pt(pts + i0);
    const G1XYZZ d = load_pt(pts + i1);
    G1XYZZ s = a;
    xyzz_add(s, d);
    G1XYZZ nd = d;
    nd.y = fp_neg(nd.y);
    xyzz_add(a, nd);  // a - b
    if (j != 0) {
        // w_{2h}^-j = Winv^(j 2^26 / 2h): composite table lookup, then out of Montgomery form
        const uint32_t e = (uint32_t)(j << (NTT_MAX_LOG - log_2h));
        Fr w = fp_load<FrP>(tw_hi + (e >> TW_BITS));
        const uint32_t l = e & (TW_SIZE - 1);
        if (l) w = fp_mul(w, fp_load<FrP>(tw_lo + l));
        a = g1_scalar_mul(a, fp_from_mont(w));
    }
    store_pt(pts + i0, s);
    store_pt(pts + i1, a);
}
// bit-reversed position -> natural position, times n^-1, to canonical affine
__global__ void __launch_bounds__(128) g1_unscramble_kernel(const G1XYZZ* __restrict__ pts, uint64_t n, uint32_t log_n, Fr n_inv_std,
                                                            G1Affine* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t src = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    const G1XYZZ v = g1_scalar_mul(load_pt(pts + src), n_inv_std);
    const G1Affine r = xyzz_to_affine(v);
    fp_store(&out[i].x, r.x);
    fp_store(&out[i].y, r.y);
}

}  // namespace kzg

using namespace kzg;

extern "C" {

int kzg_ptau_read_header(kzg_ctx* ctx, const char* path, uint32_t* power, uint32_t* ceremony_power) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !path || !power || !ceremony_power) return KZG_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(ctx, KZG_ERR_IO, std::string(path) + ": cannot open");
    std::map<uint32_t, std::vector<PtauSection>> sections;
    int r = ptau_scan(ctx, f, path, sections);
    if (r == KZG_OK) r = ptau_header(ctx, f, path, sections, power, ceremony_power);
    fclose(f);
    return r;
}

int kzg_ptau_read_tau_g2(kzg_ctx* ctx, const char* path, uint8_t out[128]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !path || !out) return KZG_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(ctx, KZG_ERR_IO, std::string(path) + ": cannot open");
    std::map<uint32_t, std::vector<PtauSection>> sections;
    int r = ptau_scan(ctx, f, path, sections);
    if (r == KZG_OK) {
        if (!sections.count(3) || sections[3][0].size < 256) {
            r = set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": no tauG2 section");
        } else if (fseeko(f, (off_t)(sections[3][0].offset + 128), SEEK_SET) != 0 || fread(out, 1, 128, f) != 128) {
            r = set_err(ctx, KZG_ERR_IO, std::string(path) + ": short read");
        }
    }
    fclose(f);
    return r;
}

int kzg_srs_from_host(kzg_ctx* ctx, const uint8_t* affine, uint64_t n_points, kzg_srs** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !out || (!affine && n_points)) return KZG_ERR_ARG;
    kzg_srs* s = new kzg_srs();
    s->n = n_points;
    if (n_points) {
        cudaError_t e = cudaMalloc((void**)&s->d, sizeof(G1Affine) * n_points);
        if (e != cudaSuccess) {
            delete s;
            return set_err(ctx, KZG_ERR_NOMEM, std::string("SRS allocation failed: ") + cudaGetErrorString(e));
        }
        e = cudaMemcpyAsync(s->d, affine, sizeof(G1Affine) * n_points, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaFree(s->d);
            delete s;
            return set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    *out = s;
    return KZG_OK;
}

// the points [first, first + n_points) of section 2 (clamped to the section): the shard of one device
int kzg_srs_load_ptau_range(kzg_ctx* ctx, const char* path, uint64_t first, uint64_t n_points, kzg_srs** out,
                            uint32_t* power_out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !path || !out) return KZG_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(ctx, KZG_ERR_IO, std::string(path) + ": cannot open");
    std::map<uint32_t, std::vector<PtauSection>> sections;
    uint32_t power = 0, cpower = 0;
    int r = ptau_scan(ctx, f, path, sections);
    if (r == KZG_OK) r = ptau_header(ctx, f, path, sections, &power, &cpower);
    if (r == KZG_OK && !sections.count(2)) r = set_err(ctx, KZG_ERR_FORMAT, std::string(path) + ": no tauG1 section");
    if (r != KZG_OK) {
        fclose(f);
        return r;
    }
    const PtauSection s2 = sections[2][0];
    const uint64_t avail = s2.size / 64;
    const uint64_t lo = first < avail ? first : avail;
    const uint64_t n = n_points < avail - lo ? n_points : avail - lo;  // the reference over-reads by one point at n = 2^power
    std::vector<uint8_t> host((size_t)n * 64);
    if (n && (fseeko(f, (off_t)(s2.offset + 64 * lo), SEEK_SET) != 0 || fread(host.data(), 1, host.size(), f) != host.size())) {
        fclose(f);
        return set_err(ctx, KZG_ERR_IO, std::string(path) + ": short read in tauG1");
    }
    fclose(f);
    kzg_srs* s = nullptr;
    KZG_TRY(kzg_srs_from_host(ctx, host.data(), n, &s));
    s->power = power;
    if (power_out) *power_out = power;
    *out = s;
    return KZG_OK;
}

int kzg_srs_load_ptau(kzg_ctx* ctx, const char* path, uint64_t n_points, kzg_srs** out, uint32_t* power_out) {
    return kzg_srs_load_ptau_range(ctx, path, 0, n_points, out, power_out);
}

// [L_i(tau)]_1, i < 2^n_bits, from the first 2^n_bits monomial points of `srs`
int kzg_srs_lagrange(kzg_ctx* ctx, kzg_srs* srs, uint32_t n_bits, kzg_srs** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !out || n_bits > 24) return KZG_ERR_ARG;
    const uint64_t n = 1ull << n_bits;
    if (srs->n < n) return set_err(ctx, KZG_ERR_ARG, "lagrange SRS: not enough monomial points");
    kzg_srs* s = new kzg_srs();
    s->n = n;
    s->power = srs->power;
    G1XYZZ* work = nullptr;
    cudaError_t e = cudaMalloc((void**)&s->d, sizeof(G1Affine) * n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&work, sizeof(G1XYZZ) * n);
    if (e != cudaSuccess) {
        cudaFree(s->d);
        delete s;
        return set_err(ctx, KZG_ERR_NOMEM, std::string("lagrange SRS allocation failed: ") + cudaGetErrorString(e));
    }
    const uint32_t blocks = (uint32_t)((n + 127) / 128);
    KZG_LAUNCH(ctx, g1_lift_kernel, blocks, 128, 0, srs->d, n, work);
    for (uint32_t st = 0; st < n_bits; st++) {
        const uint64_t h = n >> (st + 1);
        KZG_LAUNCH(ctx, g1_dif_stage_kernel, (uint32_t)((n / 2 + 127) / 128), 128, 0, work, n, h, n_bits - st, ctx->tw_lo[1],
                   ctx->tw_hi[1]);
    }
    Fr nn = fp_zero<FrP>();
    nn.l[0] = (uint32_t)n;
    nn.l[1] = (uint32_t)(n >> 32);
    const Fr n_inv_std = fp_from_mont(fp_inv(fp_to_mont(nn)));
    KZG_LAUNCH(ctx, g1_unscramble_kernel, blocks, 128, 0, work, n, n_bits, n_inv_std, s->d);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(work);
    if (e != cudaSuccess) {
        cudaFree(s->d);
        delete s;
        return set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return KZG_OK;
}

int kzg_srs_generate(kzg_ctx* ctx, const uint8_t tau_std[32], uint64_t n_points, kzg_srs** out) {
    kzg::DeviceGuard _dg(ctx);
    return kzg_srs_generate_range(ctx, tau_std, 0, n_points, out);
}

int kzg_srs_generate_range(kzg_ctx* ctx, const uint8_t tau_std[32], uint64_t first, uint64_t n_points, kzg_srs** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !tau_std || !out) return KZG_ERR_ARG;
    kzg_srs* s = new kzg_srs();
    s->n = n_points;
    uint32_t power = 0;
    while ((2ull << power) < first + n_points) power++;  // a power-p file holds ~2^(p+1) points
    s->power = power;
    if (n_points == 0) {
        *out = s;
        return KZG_OK;
    }
    cudaError_t e = cudaMalloc((void**)&s->d, sizeof(G1Affine) * n_points);
    if (e != cudaSuccess) {
        delete s;
        return set_err(ctx, KZG_ERR_NOMEM, std::string("SRS allocation failed: ") + cudaGetErrorString(e));
    }
    G1XYZZ* table = nullptr;
    e = cudaMallocAsync((void**)&table, sizeof(G1XYZZ) * 32 * 255, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(s->d);
        delete s;
        return set_err(ctx, KZG_ERR_NOMEM, cudaGetErrorString(e));
    }
    Fr tau = fp_to_mont(fr_from_bytes(tau_std));
    KZG_LAUNCH(ctx, srs_table_kernel, 1, 32, 0, table);
    KZG_LAUNCH(ctx, srs_points_kernel, (uint32_t)((n_points + 127) / 128), 128, 0, table, tau, first, n_points, s->d);
    e = cudaGetLastError();
    cudaFreeAsync(table, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(s->d);
        delete s;
        return set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return KZG_OK;
}

int kzg_srs_download(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, uint64_t count, uint8_t* out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || (!out && count)) return KZG_ERR_ARG;
    if (first + count > srs->n) return set_err(ctx, KZG_ERR_ARG, "SRS download out of bounds");
    if (!count) return KZG_OK;
    KZG_CUDA(ctx, cudaMemcpyAsync(out, srs->d + first, sizeof(G1Affine) * count, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KZG_OK;
}

int kzg_srs_write_ptau(kzg_ctx* ctx, kzg_srs* srs, uint32_t power, const uint8_t g2_one[128], const uint8_t g2_tau[128],
                       const char* path) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !g2_one || !g2_tau || !path) return KZG_ERR_ARG;
    std::vector<uint8_t> pts((size_t)srs->n * 64);
    KZG_TRY(kzg_srs_download(ctx, srs, 0, srs->n, pts.data()));
    FILE* f = fopen(path, "wb");
    if (!f) return set_err(ctx, KZG_ERR_IO, std::string(path) + ": cannot create");
    auto w32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
    auto w64 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
    fwrite("ptau", 1, 4, f);
    w32(1);
    w32(3);
    w32(1);
    w64(44);
    w32(32);
    fwrite(BN254_Q_LE, 1, 32, f);
    w32(power);
    w32(power);
    w32(2);
    w64(pts.size());
    fwrite(pts.data(), 1, pts.size(), f);
    w32(3);
    w64(256);
    fwrite(g2_one, 1, 128, f);
    fwrite(g2_tau, 1, 128, f);
    bool ok = ferror(f) == 0;
    ok &= fclose(f) == 0;
    if (!ok) return set_err(ctx, KZG_ERR_IO, std::string(path) + ": write failed");
    return KZG_OK;
}

uint64_t kzg_srs_len(kzg_srs* srs) { return srs ? srs->n : 0; }
void* kzg_srs_device_ptr(kzg_srs* srs) { return srs ? (void*)srs->d : nullptr; }

int kzg_srs_free(kzg_ctx* ctx, kzg_srs* srs) {
    kzg::DeviceGuard _dg(ctx);
    if (!srs) return KZG_OK;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(srs->d);
    cudaFree(srs->table);
    delete srs;
    return KZG_OK;
}

}  // extern "C"
