// Persistent single-token kernel: see decode_persistent.h for the design. Device code below reuses the ring, the
// activation staging and the consumers of gemv_tma_device.cuh unchanged; the stages folded in between (LayerNorm + mix,
// v6 lerp, WKV5/6 step) restate glue.cu / wkv.cu for T = 1 with the same per-element operations and reduction trees.
#include "decode_persistent.h"
#include "gemv_tma_device.cuh"
#include "decode_steps.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace rwkv {
namespace dp {

using namespace tma;
using namespace steps;

constexpr int LERP_MAXF4 = 2;              // float4 per lane per mix kept in registers across the barrier: mix <= 64
constexpr size_t DYN_SMEM_BUDGET = 110 * 1024;   // + ~2 KB static: two CTAs per SM
constexpr int PHASE_MARK_BASE = 1024, PHASE_MARK_MAX = 700;     // layout of the optional trace buffer (4096 u64)
constexpr size_t RED_BYTES = (size_t) 2 * MAX_TILE_ROWS * CONSUMER_WARPS * sizeof(float);

// ---- the program as the kernel reads it: one fully resolved record per (CTA, phase), CTA-major, so that a CTA's next record is
// ONE contiguous block it can pull into shared memory with cp.async a whole phase ahead (the descriptor chain "which problem is
// mine -> its GemvProblem -> the stage parameters" cost ~2 us of dependent L2 round trips per phase when read on demand).
struct LnLocal {
    const float * x, * ln_w, * ln_b, * state_in, * coef;
    float * state_out, * out_xx, * out_sx;
    int formula, C;
};
struct LerpLocal {
    const float * w2, * z, * xx, * sx;
    const float * maa[5];
    float * out[5];
    int C, mix, c0, n;        // this CTA's channels [c0, c0 + n), n <= 32
};
struct alignas(16) CtaPhase {
    int op, active, local, my_tiles, head, pad_[3];
    GemvProblem P;
    union { LnLocal ln; WkvStep wkv; LerpLocal lerp; } u;
};
static_assert(sizeof(CtaPhase) % 16 == 0 && sizeof(CtaPhase) <= 512, "CtaPhase must be a whole number of 16-byte cp.async granules");
constexpr int REC_GRANULES = (int) (sizeof(CtaPhase) / 16);

struct Args {
    const CtaPhase * records;        // [grid][n_phases]
    int n_phases;
    unsigned long long * bar;
    unsigned long long bar_base;
    uint32_t stage_bytes, tmp_offset, region_bytes;
    unsigned long long * trace;
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    return g;
}
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void * p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// `bytes` of immutable data starting at p (128-byte aligned) -> L2, spread over the 256 consumer threads
__device__ __forceinline__ void prefetch_l2_span(const void * p, int bytes) {
    for (int off = (int) threadIdx.x * 128; off < bytes; off += CONSUMER_THREADS * 128) prefetch_l2(reinterpret_cast<const uint8_t *>(p) + off);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__host__ __device__ __forceinline__ int tiles_of(const GemvProblem & P, int local) {
    const int n_tiles = (P.M + P.tile_rows - 1) / P.tile_rows;
    return (local >= 0 && local < P.n_cta && local < n_tiles) ? (n_tiles - local + P.n_cta - 1) / P.n_cta : 0;
}

// ---- LayerNorm + token shift + mixing for T = 1, by the 256 consumer threads, bit-identical to ln_mix_kernel<PER> (glue.cu):
// that kernel runs 1024 threads, thread v summing channels v, v + 1024, ... in double, then a warp xor-tree, 32 slots and a second
// xor-tree over the slots. Consumer thread t plays the four virtual threads v = t + 256 q: virtual warp (t >> 5) + 8 q, same lane.
// The per-channel parameters (LayerNorm weight and bias, the previous token's LN(x), the mixing vector) never change during the
// launch: the first eight channels' worth is requested before the x loads and the reductions (LnRegs), the rest before the first
// output is stored, so none of these loads is serialised behind a store it might alias.
struct LnRegs { float w[8], b[8], pv[8], cf[8]; };
__device__ __forceinline__ void ln_prefetch(const LnLocal & L, LnRegs & r, int m0) {
    const int t = threadIdx.x;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int c = t + 256 * (m0 + m);
        const bool live = c < L.C;
        r.w[m] = live ? L.ln_w[c] : 0.f;
        r.b[m] = live ? L.ln_b[c] : 0.f;
        r.pv[m] = live ? L.state_in[c] : 0.f;
        r.cf[m] = live ? L.coef[c] : 0.f;
    }
}
__device__ __forceinline__ void ln_emit(const LnLocal & L, const LnRegs & r, const float * xa, float scale_a, int m0, float * tmp, bool writer) {
    const int t = threadIdx.x;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int c = t + 256 * (m0 + m);
        if (c < L.C) {
            const float a = __fadd_rn(__fmul_rn(__fmul_rn(xa[m0 + m], scale_a), r.w[m]), r.b[m]);     // LN(x)
            const float b = r.pv[m];                                                                  // LN(x) of the previous token
            const float mc = r.cf[m];
            tmp[c] = (L.formula == 0) ? __fadd_rn(__fmul_rn(a, mc), __fsub_rn(b, __fmul_rn(b, mc)))
                                      : __fadd_rn(__fmul_rn(__fsub_rn(b, a), mc), a);
            if (writer) {
                L.state_out[c] = a;
                if (L.out_sx) L.out_sx[c] = __fsub_rn(b, a);
                if (L.out_xx) L.out_xx[c] = a;
            }
        }
    }
}
__device__ __forceinline__ void ln_mix_stage(const LnLocal & L, float * tmp, double (* slots)[32], bool writer) {
    const int C = L.C;
    LnRegs first;
    ln_prefetch(L, first, 0);            // in flight during the x loads and the two reductions
    float xa[LN_MAXCH];                   // channel t + 256 m: centred x
    float scale_a;
    ln_center_scale_256(L.x, C, xa, scale_a, slots);
    if (C > 2048) {       // CTA-uniform
        LnRegs second;
        ln_prefetch(L, second, 8);
        ln_emit(L, first, xa, scale_a, 0, tmp, writer);
        ln_emit(L, second, xa, scale_a, 8, tmp, writer);
    } else {
        ln_emit(L, first, xa, scale_a, 0, tmp, writer);
    }
}

// ---- v6 data-dependent lerp for T = 1 (v6_lerp_decode_kernel, glue.cu): 8 lanes per channel, the same eight partial sums and
// xor-tree. The W2 rows (immutable) are pulled into registers between arriving at the grid barrier and waiting on it.
struct LerpRegs { float4 w[5][LERP_MAXF4]; float maa[5]; int c; bool live; };
__device__ __forceinline__ void lerp_prefetch(const LerpLocal & p, LerpRegs & r) {
    const int C = p.C, mix = p.mix;
    const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
    r.c = p.c0 + grp;
    r.live = grp < p.n && r.c < C;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const float4 * wrow = reinterpret_cast<const float4 *>(p.w2 + ((size_t) j * C + (r.live ? r.c : 0)) * mix);
#pragma unroll
        for (int q = 0; q < LERP_MAXF4; q++) {
            const int i4 = sub + 8 * q;
            r.w[j][q] = (r.live && i4 < mix / 4) ? __ldg(wrow + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        r.maa[j] = r.live ? p.maa[j][r.c] : 0.f;
    }
}
__device__ __forceinline__ void lerp_run(const LerpLocal & p, const LerpRegs & r, float * zs) {
    const int mix = p.mix, sub = threadIdx.x & 7;
    if (p.n > 0) for (int i = threadIdx.x; i < 5 * mix; i += CONSUMER_THREADS) zs[i] = p.z[i];
    const float sx = r.live ? p.sx[r.c] : 0.f, xx = r.live ? p.xx[r.c] : 0.f;
    consumer_barrier();
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float acc = 0.f;
        const float * zj = zs + j * mix;
#pragma unroll
        for (int q = 0; q < LERP_MAXF4; q++) {
            const int i4 = sub + 8 * q;
            if (i4 < mix / 4) {
                const float4 w = r.w[j][q], z = reinterpret_cast<const float4 *>(zj)[i4];
                acc = __fmaf_rn(w.x, z.x, acc); acc = __fmaf_rn(w.y, z.y, acc);
                acc = __fmaf_rn(w.z, z.z, acc); acc = __fmaf_rn(w.w, z.w, acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (r.live && sub == 0) p.out[j][r.c] = __fadd_rn(__fmul_rn(__fadd_rn(acc, r.maa[j]), sx), xx);
    }
}

// ---- the GEMV of a phase on this CTA's tiles: stage the (single) activation column, then the unchanged consumers
template <bool STAGE_V2>
__device__ void run_gemv(Shared & sh, uint8_t * ring, uint32_t stage_bytes, uint8_t * act, float * red, int it0, int local, int my_tiles, unsigned long long * marks) {
    const GemvProblem & P = sh.P;
    if (marks && threadIdx.x == 0) marks[0] = global_timer();      // stage inputs ready (LayerNorm / nothing done)
    stage_column<8, STAGE_V2>(P, 0, act, sh.red_d);
    consumer_barrier();
    if (marks && threadIdx.x == 0) marks[1] = global_timer();      // activation column staged
    const size_t colb = act_bytes_per_column(P.type, P.K);
#define RWKV_DP_REGS(T_) consume_quant_regs<T_, true>(sh, ring, stage_bytes, act, 0, red, it0, my_tiles, local, P.n_cta)
#define RWKV_DP_SMEM(T_) consume_smem<T_, 1, true>(sh, ring, stage_bytes, act, colb, 0, 1, red, it0, my_tiles, local, P.n_cta)
    switch (P.type) {
        case DT_Q4_0: RWKV_DP_REGS(DT_Q4_0); break;
        case DT_Q4_1: RWKV_DP_REGS(DT_Q4_1); break;
        case DT_Q5_0: RWKV_DP_REGS(DT_Q5_0); break;
        case DT_Q5_1: RWKV_DP_REGS(DT_Q5_1); break;
        case DT_Q8_0: RWKV_DP_REGS(DT_Q8_0); break;
        case DT_F16: RWKV_DP_SMEM(DT_F16); break;
        default: RWKV_DP_SMEM(DT_F32); break;
    }
#undef RWKV_DP_REGS
#undef RWKV_DP_SMEM
    if (marks && threadIdx.x == 0) marks[2] = global_timer();      // this CTA's tiles consumed
}

// STAGE_V2 = the experimental bundle selected by RWKV_B200_STAGE_V2=1 (written after the last GPU run of round 1, not yet executed):
// per-block activation staging (gemv_tma_device.cuh), red.release arrival without separate fences, L2 prefetch of the LayerNorm
// parameters and the WKV state while waiting in the barrier. The <false> instantiation is the validated kernel, SASS-identical.
template <bool STAGE_V2>
__global__ void __launch_bounds__(THREADS, 2) decode_persistent_kernel(const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ Shared sh;
    __shared__ __align__(16) CtaPhase rec[2];
    __shared__ double slots[2][32];

    if (threadIdx.x == 0) {
        sh.trace = nullptr;
        for (int s = 0; s < NSTAGES; s++) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], CONSUMER_WARPS); }
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t stage_bytes = a.stage_bytes;
    uint8_t * ring = smem;
    uint8_t * act = smem + (size_t) NSTAGES * stage_bytes;
    float * tmp = reinterpret_cast<float *>(act + a.tmp_offset);
    const int n = a.n_phases;
    const CtaPhase * mine = a.records + (size_t) blockIdx.x * (size_t) n;

    if (threadIdx.x >= CONSUMER_THREADS) {
        // ===== producer: one thread walks this CTA's records and streams its weight tiles of every phase through the ring, as
        // far ahead of the consumers as the ring allows (across phase boundaries: weights are immutable) =====
        if (threadIdx.x == CONSUMER_THREADS) {
            const uint64_t policy = policy_evict_first();
            int it = 0;
            for (int ph = 0; ph < n; ph++) {
                const CtaPhase & R = mine[ph];
                const int my_tiles = R.my_tiles;
                if (my_tiles == 0) continue;
                const int local = R.local, tile_rows = R.P.tile_rows, n_cta = R.P.n_cta, M = R.P.M;
                const uint8_t * Wb = reinterpret_cast<const uint8_t *>(R.P.W);
                const size_t pitch = (size_t) R.P.pitch;
                for (int i = 0; i < my_tiles; i++, it++) {
                    const int row0 = (local + i * n_cta) * tile_rows;
                    const int rows = min(tile_rows, M - row0);
                    const int s = it % NSTAGES;
                    if (it >= NSTAGES) mbar_wait_t<true>(&sh.empty[s], (uint32_t) (((it / NSTAGES) - 1) & 1));
                    const uint32_t bytes = (uint32_t) ((size_t) rows * pitch);
                    mbar_expect_tx(&sh.full[s], bytes);
                    bulk_copy_g2s(ring + (size_t) s * stage_bytes, Wb + (size_t) row0 * pitch, bytes, &sh.full[s], policy);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    float * red = reinterpret_cast<float *>(act + a.region_bytes);
    const int tid = threadIdx.x;
    if (tid < REC_GRANULES) cp_async16(reinterpret_cast<uint8_t *>(&rec[0]) + tid * 16, reinterpret_cast<const uint8_t *>(mine) + tid * 16);
    cp_async_commit();
    int it = 0;
    for (int ph = 0; ph < n; ph++) {
        // (1) this phase's record has landed; the previous phase is complete on this CTA: arrive at the grid barrier
        cp_async_wait_all();
        consumer_barrier();
        if (ph > 0 && tid == 0) {
            if constexpr (STAGE_V2) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(a.bar) : "memory");     // 1.65 vs 1.95 us per barrier (microbench)
            else { __threadfence(); atomicAdd(a.bar, 1ull); }
        }
        const CtaPhase & R = rec[ph & 1];
        const int op = R.op, active = R.active, local = R.local, my_tiles = R.my_tiles;
        // (2) work that needs nothing from the previous phase: the next record, immutable operands of this phase's stage
        if (ph + 1 < n && tid < REC_GRANULES)
            cp_async16(reinterpret_cast<uint8_t *>(&rec[(ph + 1) & 1]) + tid * 16, reinterpret_cast<const uint8_t *>(mine + ph + 1) + tid * 16);
        cp_async_commit();
        LerpRegs lr;
        if (op == DOP_LERP) lerp_prefetch(R.u.lerp, lr);
        if constexpr (STAGE_V2) {
            // immutable operands of this phase's stage, requested into L2 while the CTA waits for the others: by the time a layer
            // comes round again its small vectors have been flushed by the 165 MB of weights streamed in between
            if (op == DOP_LNMIX_GEMV && active) {
                const LnLocal & L = R.u.ln;
                prefetch_l2_span(L.ln_w, L.C * 4); prefetch_l2_span(L.ln_b, L.C * 4);
                prefetch_l2_span(L.coef, L.C * 4); prefetch_l2_span(L.state_in, L.C * 4);
            } else if (op == DOP_GEMV_WKV && R.head >= 0) {
                const WkvStep & w = R.u.wkv;
                const size_t hb = (size_t) R.head * w.S;
                prefetch_l2_span(w.state_in + hb * w.S, w.S * w.S * 4);
                if (tid < 2) prefetch_l2((tid == 0 ? w.lnx_w : w.lnx_b) + hb);
            }
        }
        // (3) wait for every CTA
        if (tid == 0) {
            if (active) {
                sh.P = R.P;
                if (op == DOP_LNMIX_GEMV) { sh.P.x = tmp; sh.P.ldx = 0; }
            }
            if (ph > 0) {
                const unsigned long long target = a.bar_base + (unsigned long long) ph * gridDim.x;
                const long long t0 = clock64();
                unsigned spins = 0;
                while (ld_acquire_u64(a.bar) < target) {
                    if ((++spins & 0x3FFu) == 0 && clock64() - t0 > GUARD_CYCLES) __trap();
                }
                if constexpr (!STAGE_V2) __threadfence();      // the acquire load above already orders everything after it
            }
            if (a.trace && blockIdx.x == 0) a.trace[ph] = global_timer();
        }
        consumer_barrier();
        // (4) the phase. CTA 0 leaves three intra-phase time marks per phase behind the boundary array when tracing is armed.
        unsigned long long * marks = (a.trace && blockIdx.x == 0 && ph < PHASE_MARK_MAX) ? a.trace + PHASE_MARK_BASE + 4 * ph : nullptr;
        switch (op) {
            case DOP_LNMIX_GEMV:
                if (active) {
                    ln_mix_stage(R.u.ln, tmp, slots, blockIdx.x == 0);
                    consumer_barrier();
                    run_gemv<STAGE_V2>(sh, ring, stage_bytes, act, red, it, local, my_tiles, marks);
                }
                break;
            case DOP_LERP:
                lerp_run(R.u.lerp, lr, tmp);
                break;
            case DOP_GEMV_WKV: {
                if (active) run_gemv<STAGE_V2>(sh, ring, stage_bytes, act, red, it, local, my_tiles, marks);
                const int h = R.head;
                if (h >= 0) {
                    consumer_barrier();     // the head's decay values just stored by this CTA are visible to all its threads
                    switch (R.u.wkv.S) {
                        case 8: wkv6_step<8>(R.u.wkv, h, tmp); break;
                        case 16: wkv6_step<16>(R.u.wkv, h, tmp); break;
                        case 32: wkv6_step<32>(R.u.wkv, h, tmp); break;
                        default: wkv6_step<64>(R.u.wkv, h, tmp); break;
                    }
                }
                break;
            }
            default:
                if (active) run_gemv<STAGE_V2>(sh, ring, stage_bytes, act, red, it, local, my_tiles, marks);
                break;
        }
        it += my_tiles;
    }
    if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[n] = global_timer();
}

}  // namespace dp

// ---- host side ----------------------------------------------------------------------------------------------------
bool gemv_tma_plan(GemvBatch & batch, int total_ctas, long long stage_bytes, size_t * max_col_bytes);   // gemv_tma.cu

static bool stage_v2_enabled() {
    static const bool on = [] { const char * e = getenv("RWKV_B200_STAGE_V2"); return e && atoi(e) != 0; }();
    return on;
}

void decode_program_free(DecodeProgram & program) {
    if (program.records) cudaFree(program.records);
    program = DecodeProgram();
}

// Replays the tile walk of every CTA exactly as the kernel's producer and consumers compute it and checks that each tile of
// each matrix is taken exactly once, fits a ring stage, and that a WKV phase maps one tile to one head. A program that fails
// this is never launched (the kernel has no way to recover from a tile-count mismatch between producer and consumers).
static bool program_selfcheck(const std::vector<DecodePhase> & phases, const DecodeProgram & program) {
    std::vector<int> taken;
    for (size_t ph = 0; ph < phases.size(); ph++) {
        const DecodePhase & D = phases[ph];
        const GemvBatch & B = D.batch;
        if (B.n < 0 || B.n > GEMV_MAX_PROBLEMS) return false;
        for (int i = 0; i < B.n; i++) {
            const GemvProblem & P = B.p[i];
            if (P.tile_rows <= 0 || P.n_cta <= 0 || P.first_cta < 0 || P.first_cta + P.n_cta > program.grid) return false;
            if (i > 0 && P.first_cta != B.p[i - 1].first_cta + B.p[i - 1].n_cta) return false;
            if (i == 0 && P.first_cta != 0) return false;
            if ((long long) P.tile_rows * P.pitch > (long long) program.stage_bytes) return false;
            if (P.tile_rows > tma::MAX_TILE_ROWS || P.tile_rows % (tma::CONSUMER_WARPS / P.wk) != 0) return false;
            if (tma::act_bytes_per_column(P.type, P.K) > program.region_bytes) return false;
            const int n_tiles = (P.M + P.tile_rows - 1) / P.tile_rows;
            taken.assign((size_t) n_tiles, 0);
            for (int cta = 0; cta < program.grid; cta++) {
                int pi = 0;                                   // the kernel's problem lookup
                for (int j = 1; j < B.n; j++) if (cta >= B.p[j].first_cta) pi = j;
                if (pi != i) continue;
                const int local = cta - P.first_cta;
                const int my = (local < P.n_cta && local < n_tiles) ? (n_tiles - local + P.n_cta - 1) / P.n_cta : 0;
                for (int t = 0; t < my; t++) {
                    const int tile = local + t * P.n_cta;
                    if (tile < 0 || tile >= n_tiles) return false;
                    taken[(size_t) tile]++;
                }
            }
            for (int t = 0; t < n_tiles; t++) if (taken[(size_t) t] != 1) return false;
        }
        if (D.op == DOP_GEMV_WKV && B.n == 1) {
            const GemvProblem & P = B.p[0];
            if (P.tile_rows != D.wkv.S || P.n_cta != D.wkv.H || P.M != D.wkv.H * D.wkv.S) return false;
        }
        if (D.op == DOP_LNMIX_GEMV && (size_t) program.tmp_offset + (size_t) D.ln.C * 4 > program.region_bytes) return false;
        if (D.op == DOP_LERP && (size_t) program.tmp_offset + (size_t) 5 * D.lerp.mix * 4 > program.region_bytes) return false;
    }
    return (size_t) tma::NSTAGES * program.stage_bytes + program.region_bytes + dp::RED_BYTES == program.smem_bytes && program.smem_bytes <= dp::DYN_SMEM_BUDGET;
}

static std::vector<dp::CtaPhase> flatten_program(const std::vector<DecodePhase> & phases, const DecodeProgram & program);
static bool records_selfcheck(const std::vector<dp::CtaPhase> & rec, const std::vector<DecodePhase> & phases, const DecodeProgram & program);

// Host-only part of decode_program_build: shapes -> shared-memory layout -> tiles and CTA shares.
bool decode_program_plan(std::vector<DecodePhase> & phases, int num_sms, DecodeProgram & program) {
    using namespace dp;
    program = DecodeProgram();
    const int total_ctas = 2 * num_sms;
    if (phases.empty() || num_sms <= 0) return false;
    // pass 1: shapes -> shared-memory layout
    size_t max_col = 0, tmp_off = 0, tmp_need = 0;
    for (DecodePhase & D : phases) {
        size_t col = 0;
        if (D.batch.n > 0) {
            D.batch.T = 1;
            if (!gemv_tma_plan(D.batch, total_ctas, tma::NOMINAL_STAGE_BYTES, &col)) return false;
            if (col > max_col) max_col = col;
        }
        if (D.op == DOP_LNMIX_GEMV) {
            if (D.batch.n == 0 || D.ln.C > 256 * LN_MAXCH) return false;
            if (col > tmp_off) tmp_off = col;
            if ((size_t) D.ln.C * 4 > tmp_need) tmp_need = (size_t) D.ln.C * 4;
        } else if (D.op == DOP_LERP) {
            const V6LerpParams & p = D.lerp;
            const int cpc = (p.C + total_ctas - 1) / total_ctas;
            if (D.batch.n != 0 || (p.mix & 3) || p.mix / 4 > 8 * LERP_MAXF4 || cpc > 32 || p.T != 1) return false;
            if ((size_t) 5 * p.mix * 4 > tmp_need) tmp_need = (size_t) 5 * p.mix * 4;
        } else if (D.op == DOP_GEMV_WKV) {
            const Wkv6Params & p = D.wkv;
            if (!(p.S == 8 || p.S == 16 || p.S == 32 || p.S == 64) || p.H > total_ctas || p.T != 1 || D.batch.n > 1) return false;
            if (col > tmp_off) tmp_off = col;
            if ((size_t) p.S * 4 > tmp_need) tmp_need = (size_t) p.S * 4;
        } else if (D.op != DOP_GEMV || D.batch.n == 0) {
            return false;
        }
    }
    tmp_off = (tmp_off + 15) & ~(size_t) 15;
    size_t region = tmp_off + tmp_need;
    if (max_col > region) region = max_col;
    region = (region + 127) & ~(size_t) 127;
    if (region + RED_BYTES + 3 * 16 * 1024 > DYN_SMEM_BUDGET) return false;
    const long long stage_bytes = (long long) ((DYN_SMEM_BUDGET - region - RED_BYTES) / tma::NSTAGES / 1024 * 1024);
    // pass 2: tiles and CTA shares for the real stage size
    for (DecodePhase & D : phases) {
        if (D.batch.n == 0) continue;
        if (!gemv_tma_plan(D.batch, total_ctas, stage_bytes, nullptr)) return false;
        D.batch.stage_bytes = stage_bytes;
        D.batch.trace = nullptr;
        if (D.op == DOP_GEMV_WKV) {     // one tile = one head, CTA h owns head h
            GemvProblem & p = D.batch.p[0];
            const int S = D.wkv.S, wr = tma::CONSUMER_WARPS / p.wk;
            if (p.M != D.wkv.H * S || S % wr != 0 || S > tma::MAX_TILE_ROWS || (long long) S * p.pitch > stage_bytes) return false;
            p.tile_rows = S; p.n_cta = D.wkv.H; p.first_cta = 0;
        }
    }
    program.grid = total_ctas;
    program.stage_bytes = (uint32_t) stage_bytes;
    program.tmp_offset = (uint32_t) tmp_off;
    program.region_bytes = (uint32_t) region;
    program.smem_bytes = (size_t) tma::NSTAGES * (size_t) stage_bytes + region + RED_BYTES;
    program.n_phases = (int) phases.size();
    if (!program_selfcheck(phases, program)) {
        fprintf(stderr, "rwkv_b200: the persistent decode program failed its self-check; using the per-launch path\n");
        return false;
    }
    return true;
}

bool decode_program_plan_check_records(std::vector<DecodePhase> & phases, int num_sms, DecodeProgram & program) {
    if (!decode_program_plan(phases, num_sms, program)) return false;
    return records_selfcheck(flatten_program(phases, program), phases, program);
}

// The planned phase list resolved per CTA (see CtaPhase): what each CTA does in each phase, with the same problem lookup and
// tile arithmetic the self-check replayed.
static std::vector<dp::CtaPhase> flatten_program(const std::vector<DecodePhase> & phases, const DecodeProgram & program) {
    using namespace dp;
    const int grid = program.grid, n = (int) phases.size();
    std::vector<CtaPhase> rec((size_t) grid * (size_t) n);
    memset(rec.data(), 0, rec.size() * sizeof(CtaPhase));
    int prev_active = 0;      // CTAs [0, prev_active) had tiles in the previous phase
    for (int ph = 0; ph < n; ph++) {
        const DecodePhase & D = phases[(size_t) ph];
        const GemvBatch & B = D.batch;
        int used = 0;
        for (int i = 0; i < B.n; i++) used = B.p[i].first_cta + B.p[i].n_cta;
        // lerp channels go to the CTAs that sat idle in the previous phase when those alone can take them (<= 32 channels each)
        int lerp_first = 0, lerp_ctas = grid, lerp_cpc = 0;
        if (D.op == DOP_LERP) {
            if (prev_active < grid && (long long) (grid - prev_active) * 32 >= D.lerp.C) { lerp_first = prev_active; lerp_ctas = grid - prev_active; }
            lerp_cpc = (D.lerp.C + lerp_ctas - 1) / lerp_ctas;
        }
        for (int cta = 0; cta < grid; cta++) {
            CtaPhase & R = rec[(size_t) cta * (size_t) n + (size_t) ph];
            R.op = D.op; R.head = -1; R.local = cta;
            if (B.n > 0) {
                int pi = 0;
                for (int j = 1; j < B.n; j++) if (cta >= B.p[j].first_cta) pi = j;
                const GemvProblem & P = B.p[pi];
                R.local = cta - P.first_cta;
                R.my_tiles = tiles_of(P, R.local);
                R.active = R.my_tiles > 0;
                if (R.active) R.P = P; else R.local = cta;
                if (D.op == DOP_LNMIX_GEMV && R.active) {
                    LnLocal & l = R.u.ln;
                    l.x = D.ln.x; l.ln_w = D.ln.ln_w; l.ln_b = D.ln.ln_b; l.state_in = D.ln.state_in; l.coef = D.ln.coef[pi];
                    l.state_out = D.ln.state_out; l.out_xx = D.ln.out_xx; l.out_sx = D.ln.out_sx; l.formula = D.ln.formula; l.C = D.ln.C;
                }
            }
            if (D.op == DOP_GEMV_WKV) {
                const Wkv6Params & w = D.wkv;
                R.head = (B.n > 0) ? (R.active ? R.local : -1) : (cta < w.H ? cta : -1);
                WkvStep & k = R.u.wkv;
                k.r = w.r; k.k = w.k; k.v = w.v; k.td = w.td; k.tf = w.tf; k.state_in = w.state_in; k.lnx_w = w.lnx_w; k.lnx_b = w.lnx_b; k.g = w.g;
                k.state_out = w.state_out; k.y = w.y; k.eps = w.eps; k.td_per_token = w.td_per_token; k.per_head_scalars = w.per_head_scalars;
                k.H = w.H; k.S = w.S;
            } else if (D.op == DOP_LERP) {
                const V6LerpParams & v = D.lerp;
                LerpLocal & l = R.u.lerp;
                l.w2 = v.w2; l.z = v.z; l.xx = v.xx; l.sx = v.sx; l.C = v.C; l.mix = v.mix;
                for (int j = 0; j < 5; j++) { l.maa[j] = v.maa[j]; l.out[j] = v.out[j]; }
                const int idx = cta - lerp_first;
                l.c0 = (idx >= 0) ? idx * lerp_cpc : v.C;
                l.n = (idx >= 0 && l.c0 < v.C) ? (v.C - l.c0 < lerp_cpc ? v.C - l.c0 : lerp_cpc) : 0;
                if (l.n == 0) l.c0 = 0;
            }
        }
        prev_active = used;
    }
    return rec;
}

// Host check of the flattened records: every tile of every phase exactly once (again, now from what the kernel will actually
// read), every lerp channel exactly once, at most 32 lerp channels per CTA.
static bool records_selfcheck(const std::vector<dp::CtaPhase> & rec, const std::vector<DecodePhase> & phases, const DecodeProgram & program) {
    using namespace dp;
    const int grid = program.grid, n = (int) phases.size();
    std::vector<int> seen;
    for (int ph = 0; ph < n; ph++) {
        const DecodePhase & D = phases[(size_t) ph];
        for (int i = 0; i < D.batch.n; i++) {
            const GemvProblem & P = D.batch.p[i];
            const int n_tiles = (P.M + P.tile_rows - 1) / P.tile_rows;
            seen.assign((size_t) n_tiles, 0);
            for (int cta = 0; cta < grid; cta++) {
                const CtaPhase & R = rec[(size_t) cta * n + ph];
                if (!R.active || R.P.W != P.W) continue;
                for (int t = 0; t < R.my_tiles; t++) {
                    const int tile = R.local + t * R.P.n_cta;
                    if (tile < 0 || tile >= n_tiles) return false;
                    seen[(size_t) tile]++;
                }
            }
            for (int t = 0; t < n_tiles; t++) if (seen[(size_t) t] != 1) return false;
        }
        if (D.op == DOP_LERP) {
            seen.assign((size_t) D.lerp.C, 0);
            for (int cta = 0; cta < grid; cta++) {
                const LerpLocal & l = rec[(size_t) cta * n + ph].u.lerp;
                if (l.n < 0 || l.n > 32 || l.c0 < 0 || l.c0 + l.n > D.lerp.C) return false;
                for (int c = l.c0; c < l.c0 + l.n; c++) seen[(size_t) c]++;
            }
            for (int c = 0; c < D.lerp.C; c++) if (seen[(size_t) c] != 1) return false;
        }
        if (D.op == DOP_GEMV_WKV) {
            seen.assign((size_t) D.wkv.H, 0);
            for (int cta = 0; cta < grid; cta++) { const int h = rec[(size_t) cta * n + ph].head; if (h >= D.wkv.H) return false; if (h >= 0) seen[(size_t) h]++; }
            for (int h = 0; h < D.wkv.H; h++) if (seen[(size_t) h] != 1) return false;
        }
    }
    return true;
}

bool decode_program_build(std::vector<DecodePhase> & phases, const DeviceInfo & dev, DecodeProgram & program) {
    using namespace dp;
    if (!decode_program_plan(phases, dev.num_sms, program)) { program = DecodeProgram(); return false; }
    static bool attr_set_dev[64] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    cur_dev = (cur_dev < 0 || cur_dev >= 64) ? 0 : cur_dev;
    if (!attr_set_dev[cur_dev]) {
        if (cudaFuncSetAttribute(decode_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DYN_SMEM_BUDGET) != cudaSuccess ||
            cudaFuncSetAttribute(decode_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DYN_SMEM_BUDGET) != cudaSuccess) { cudaGetLastError(); return false; }
        attr_set_dev[cur_dev] = true;
    }
    int per_sm = 0, per_sm_v2 = 0, coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cur_dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_v2, decode_persistent_kernel<true>, tma::THREADS, program.smem_bytes) != cudaSuccess) per_sm_v2 = 0;
    if (stage_v2_enabled() && per_sm_v2 < 2) { cudaGetLastError(); program = DecodeProgram(); return false; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel<false>, tma::THREADS, program.smem_bytes) != cudaSuccess || per_sm < 2 || !coop) {
        cudaGetLastError();
        fprintf(stderr, "rwkv_b200: the persistent decode kernel cannot keep 2 CTAs per SM resident (%d, cooperative launch %d); using the per-launch path\n", per_sm, coop);
        program = DecodeProgram();
        return false;
    }
    const std::vector<CtaPhase> rec = flatten_program(phases, program);
    if (!records_selfcheck(rec, phases, program)) {
        fprintf(stderr, "rwkv_b200: the persistent decode program's per-CTA records failed their self-check; using the per-launch path\n");
        program = DecodeProgram();
        return false;
    }
    if (cudaMalloc(&program.records, rec.size() * sizeof(CtaPhase)) != cudaSuccess) { cudaGetLastError(); program = DecodeProgram(); return false; }
    if (cudaMemcpy(program.records, rec.data(), rec.size() * sizeof(CtaPhase), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        decode_program_free(program);
        return false;
    }
    program.record_bytes = rec.size() * sizeof(CtaPhase);
    program.supported = true;
    return true;
}

cudaError_t decode_program_launch(const DecodeProgram & program, unsigned long long * barrier_counter, unsigned long long barrier_base,
                                  unsigned long long * trace, cudaStream_t stream) {
    if (!program.supported) return cudaErrorNotSupported;
    dp::Args a;
    a.records = reinterpret_cast<const dp::CtaPhase *>(program.records); a.n_phases = program.n_phases;
    a.bar = barrier_counter; a.bar_base = barrier_base;
    a.stage_bytes = program.stage_bytes; a.tmp_offset = program.tmp_offset; a.region_bytes = program.region_bytes;
    a.trace = trace;
    void * args[] = {&a};
    g_kernel_launches++;
    const void * kernel = stage_v2_enabled() ? reinterpret_cast<const void *>(dp::decode_persistent_kernel<true>) : reinterpret_cast<const void *>(dp::decode_persistent_kernel<false>);
    return cudaLaunchCooperativeKernel(kernel, dim3((unsigned) program.grid), dim3(tma::THREADS), args, program.smem_bytes, stream);
}

}  // namespace rwkv
