// Host-side helpers of the C ABI: the Keccak-256 sponge of the Fiat-Shamir transcript and the byte
// conversions around it.  By mandate the transcript stays on the host (reference
// src/Keccak256Transcript.js:7-52); these are the O(1)-sized pieces a host binding needs so that it
// does not have to carry its own big-integer code.
#include <string.h>

#include "common.cuh"

namespace kzg {

static inline uint64_t rol64(uint64_t x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }

// Keccak-f[1600], state indexed a[x + 5 y]
static void keccak_f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    static const unsigned ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], ROT[x + 5 * y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[rnd];
    }
}

// 32-byte big-endian integer -> 8 little-endian limbs
static void be32_to_limbs(const uint8_t in[32], uint32_t l[8]) {
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = in + 28 - 4 * i;
        l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
static void limbs_to_be32(const uint32_t l[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) {
        uint8_t* p = out + 28 - 4 * i;
        p[0] = (uint8_t)(l[i] >> 24);
        p[1] = (uint8_t)(l[i] >> 16);
        p[2] = (uint8_t)(l[i] >> 8);
        p[3] = (uint8_t)l[i];
    }
}

}  // namespace kzg

using namespace kzg;

extern "C" {

// js-sha3 `keccak256`: rate 136, pad10*1 with the ORIGINAL domain byte 0x01 (not SHA3's 0x06)
void kzg_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t a[25];
    memset(a, 0, sizeof(a));
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w;
            memcpy(&w, data + 8 * i, 8);
            a[i] ^= w;
        }
        keccak_f1600(a);
        data += rate;
        len -= rate;
    }
    uint8_t blk[136];
    memset(blk, 0, sizeof(blk));
    if (len) memcpy(blk, data, len);
    blk[len] ^= 0x01;
    blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) {
        uint64_t w;
        memcpy(&w, blk + 8 * i, 8);
        a[i] ^= w;
    }
    keccak_f1600(a);
    memcpy(out, a, 32);
}

// G1.toRprUncompressed (Keccak256Transcript.js:42): x || y, 32 B big-endian standard form each.
// Infinity: zeros with 0x40 in byte 0.
void kzg_g1_to_rpr_uncompressed(const uint8_t in[64], uint8_t out[64]) {
    G1Affine p;
    memcpy(p.x.l, in, 32);
    memcpy(p.y.l, in + 32, 32);
    if (g1_affine_is_inf(p)) {
        memset(out, 0, 64);
        out[0] |= 0x40;
        return;
    }
    Fq x = fp_from_mont(p.x), y = fp_from_mont(p.y);
    limbs_to_be32(x.l, out);
    limbs_to_be32(y.l, out + 32);
}

// Fr.toRprBE (Keccak256Transcript.js:45)
void kzg_fr_to_rpr_be(const uint8_t in[32], uint8_t out[32]) {
    Fr a = fp_from_mont(fr_from_bytes(in));
    limbs_to_be32(a.l, out);
}

// Fr.e(Scalar.fromRprBE(hash)) (Keccak256Transcript.js:50-51): big-endian integer mod r -> Montgomery.
// 2^256 < 6r, so at most five conditional subtractions reduce the raw hash.
void kzg_fr_from_hash_be(const uint8_t in[32], uint8_t out[32]) {
    uint32_t l[8];
    be32_to_limbs(in, l);
    while (fp_geq_mod<FrP>(l)) {
        int64_t bw = 0;
        for (int i = 0; i < 8; i++) {
            bw += (int64_t)l[i] - (int64_t)FrP::mod(i);
            l[i] = (uint32_t)bw;
            bw >>= 32;
        }
    }
    Fr a;
    for (int i = 0; i < 8; i++) a.l[i] = l[i];
    fr_to_bytes(fp_to_mont(a), out);
}

}  // extern "C"
