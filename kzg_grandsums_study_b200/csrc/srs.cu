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
