// CPU check of the __host__ __device__ quant-block decoding used by the CUDA GEMV
// (rwkv.cpp_b200/csrc/kernels/quant_decode.cuh) against a plain scalar restatement of the ggml block
// layouts (ggml-common.h:161-221, ggml-quants.c:255-347). Built and run by tests/test_host_logic.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../rwkv.cpp_b200/csrc/kernels/quant_decode.cuh"

using namespace rwkv;

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

// scalar dequant of block `b` of a row in native layout -> 32 ints (stored values), d, m
static void scalar_decode(int type, const uint8_t * row, int b, int q[32], float & d, float & m) {
    const int bb = dtype_block_bytes(type);
    const uint8_t * p = row + (size_t) b * bb;
    uint16_t h; memcpy(&h, p, 2); d = fp16_to_fp32(h); m = 0.f;
    uint32_t qh = 0; const uint8_t * qs;
    switch (type) {
        case DT_Q4_0: qs = p + 2; break;
        case DT_Q4_1: memcpy(&h, p + 2, 2); m = fp16_to_fp32(h); qs = p + 4; break;
        case DT_Q5_0: memcpy(&qh, p + 2, 4); qs = p + 6; break;
        case DT_Q5_1: memcpy(&h, p + 2, 2); m = fp16_to_fp32(h); memcpy(&qh, p + 4, 4); qs = p + 8; break;
        default: qs = p + 2; break;
    }
    if (type == DT_Q8_0) { for (int j = 0; j < 32; j++) q[j] = (int8_t) qs[j]; return; }
    const bool five = type == DT_Q5_0 || type == DT_Q5_1;
    for (int j = 0; j < 16; j++) {
        q[j] = qs[j] & 0x0F; q[j + 16] = qs[j] >> 4;
        if (five) { q[j] |= ((qh >> j) & 1) << 4; q[j + 16] |= ((qh >> (j + 16)) & 1) << 4; }
    }
}

template <int TYPE> static int check_type(const char * name) {
    using TR = QTraits<TYPE>;
    const int nblk = 14;   // 7 pairs
    const int bb = dtype_block_bytes(TYPE);
    std::vector<uint8_t> row((size_t) nblk * bb + 64);
    int bad = 0;
    for (int trial = 0; trial < 200; trial++) {
        for (auto & x : row) x = (uint8_t) rnd();
        // keep the fp16 scale fields finite
        for (int b = 0; b < nblk; b++) {
            uint16_t h = fp32_to_fp16(((int) (rnd() % 2000) - 1000) / 4096.0f);
            memcpy(&row[(size_t) b * bb], &h, 2);
            if (TR::HAS_MIN) { h = fp32_to_fp16(((int) (rnd() % 2000) - 1000) / 512.0f); memcpy(&row[(size_t) b * bb + 2], &h, 2); }
        }
        int8_t act[nblk * 32];
        for (auto & a : act) a = (int8_t) ((int) (rnd() % 255) - 127);
        const int nunits = nblk / TR::UNIT_BLOCKS;
        for (int u = 0; u < nunits; u++) {
            // what the loader leaves in HBM: the file bytes with every Q5 block's fifth-bit word in device order (qh5_to_device)
            std::vector<uint8_t> dev(row.begin() + (size_t) u * TR::UNIT_WORDS * 4, row.begin() + (size_t) (u + 1) * TR::UNIT_WORDS * 4);
            if (TYPE == DT_Q5_0 || TYPE == DT_Q5_1) {
                for (int b = 0; b < TR::UNIT_BLOCKS; b++) {
                    uint8_t * qp = dev.data() + (size_t) b * bb + (TYPE == DT_Q5_0 ? 2 : 4);
                    uint32_t qh; memcpy(&qh, qp, 4);
                    if (qh5_to_file(qh5_to_device(qh)) != qh) { printf("%s: qh5 round trip\n", name); bad++; }
                    qh = qh5_to_device(qh); memcpy(qp, &qh, 4);
                }
            }
            uint32_t w[TR::UNIT_WORDS];
            memcpy(w, dev.data(), sizeof(w));
            for (int b = 0; b < TR::UNIT_BLOCKS; b++) {
                const int blk = u * TR::UNIT_BLOCKS + b;
                BlockQ bq; decode_block<TYPE>(w, b, bq);
                int q[32]; float d, m; scalar_decode(TYPE, row.data(), blk, q, d, m);
                for (int j = 0; j < 32; j++) {
                    int got = (int) (int8_t) ((uint32_t) bq.q[j / 4] >> (8 * (j % 4)));
                    if (got != q[j]) { if (bad < 5) printf("%s: trial %d blk %d elem %d: got %d want %d\n", name, trial, blk, j, got, q[j]); bad++; }
                }
                if (bq.d != d || bq.m != m) { if (bad < 5) printf("%s: scale mismatch blk %d\n", name, blk); bad++; }
                // block_dot against the scalar formula
                int a8[8]; memcpy(a8, act + blk * 32, 32);
                int isum = 0, asum = 0;
                for (int j = 0; j < 32; j++) { isum += (q[j] - TR::OFFSET) * act[blk * 32 + j]; asum += act[blk * 32 + j]; }
                ActScale as; as.d = 0.0123f; as.s = TR::HAS_MIN ? 0.5f : (float) asum;
                float want = fmaf(d * as.d, (float) (TR::HAS_MIN ? isum : isum), 1.0f);
                if (TR::HAS_MIN) want = fmaf(m, as.s, want);
                float got = block_dot<TYPE>(bq, a8, as, 1.0f);
                if (got != want) { if (bad < 5) printf("%s: dot mismatch blk %d: %g vs %g\n", name, blk, got, want); bad++; }
            }
        }
    }
    printf("%s: %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    int bad = 0;
    bad += check_type<DT_Q4_0>("Q4_0");
    bad += check_type<DT_Q4_1>("Q4_1");
    bad += check_type<DT_Q5_0>("Q5_0");
    bad += check_type<DT_Q5_1>("Q5_1");
    bad += check_type<DT_Q8_0>("Q8_0");
    for (uint32_t n = 0; n < 16; n++) {
        uint32_t want = 0;
        for (int b = 0; b < 4; b++) want |= ((n >> b) & 1u) << (8 * b + 4);
        if (spread_bit5(n) != want) { printf("spread_bit5(%u) wrong\n", n); bad++; }
    }
    return bad ? 1 : 0;
}
