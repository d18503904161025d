// Single-token evaluation as ONE persistent kernel ("layer program").
//
// A decode step of the per-launch path is ~10 dependent kernels per layer; each pays ~4.5 us of launch dependency
// (profiles/r1_trace_decode_v6.csv), which at 7B costs more than streaming the layer's weights. Here the whole token is one
// cooperative launch of 2 CTAs per SM: every CTA walks the same list of phases; inside a phase it runs the unchanged
// streaming GEMV consumers of gemv_tma_device.cuh on its share of the phase's matrices, phases are separated by a
// grid-wide barrier (one atomic per CTA), and the TMA producer thread of every CTA keeps streaming the NEXT phase's
// weight tiles into its ring while the consumers sit in the barrier (weights are immutable, so the producer never waits
// for activations). The small stages between GEMVs are folded in:
//   DOP_LNMIX_GEMV  LayerNorm + token shift + mixing (rwkv_carry_x and the lerps, rwkv_graph.inc:56-82, 94-97, 310-311)
//                   recomputed by every CTA of the phase straight into its activation staging buffer, then the GEMV
//   DOP_LERP        v6 data-dependent lerp (rwkv_graph.inc:323-346), channels spread over all CTAs
//   DOP_GEMV_WKV    optional decay GEMV (one tile = one head) followed by the head's WKV5/6 step + head norm + ln_x + gate
//   DOP_GEMV        plain batch of fused dequantize-GEMVs
// Arithmetic per output is the per-launch path's, operation for operation (same lane mapping and summation trees),
// so a token evaluated here is bit-identical to one evaluated by the kernels of gemv_tma.cu / glue.cu / wkv.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "gemv.h"
#include "ops.h"

namespace rwkv {

enum DecodeOp : int { DOP_GEMV = 0, DOP_LNMIX_GEMV = 1, DOP_LERP = 2, DOP_GEMV_WKV = 3 };

struct DecodeLnMix {
    const float * x, * ln_w, * ln_b;
    const float * state_in;                   // [C] LN(x) of the previous token
    float * state_out;                        // [C] <- LN(x), written by CTA 0
    const float * coef[GEMV_MAX_PROBLEMS];    // mixing vector that feeds problem i of the phase's batch
    float * out_xx, * out_sx;                 // optional [C] side outputs (v6: LN(x) and prev - LN(x)), written by CTA 0
    int formula, C;                           // formulas of LnMixParams
};

struct DecodePhase {
    int op;
    GemvBatch batch;        // batch.n == 0: no GEMV in this phase
    DecodeLnMix ln;         // DOP_LNMIX_GEMV
    V6LerpParams lerp;      // DOP_LERP
    Wkv6Params wkv;         // DOP_GEMV_WKV (T == 1)
};

struct DecodeProgram {      // device-resident, immutable once built
    void * records = nullptr;           // device: one resolved record per (CTA, phase), CTA-major (decode_persistent.cu: CtaPhase)
    size_t record_bytes = 0;
    int n_phases = 0;
    int grid = 0;
    uint32_t stage_bytes = 0, tmp_offset = 0, region_bytes = 0;
    size_t smem_bytes = 0;
    bool supported = false;
};

// Plans every batch (warps per row, tile heights, CTA shares), lays out the shared memory and uploads the phase list.
// Returns false (program.supported == false) when some shape does not fit this kernel; the caller then keeps using the
// per-launch path. `phases` holds unplanned batches (W, type, K, M, x, y, epilogue operands).
bool decode_program_build(std::vector<DecodePhase> & phases, const DeviceInfo & dev, DecodeProgram & program);
// The host-only part of the above (no CUDA calls): planning, layout and the tile-walk self-check.
bool decode_program_plan(std::vector<DecodePhase> & phases, int num_sms, DecodeProgram & program);
// decode_program_plan + the per-CTA flattening and its self-check (still host only).
bool decode_program_plan_check_records(std::vector<DecodePhase> & phases, int num_sms, DecodeProgram & program);
void decode_program_free(DecodeProgram & program);

// Enqueues one token. `barrier_counter` is a device u64 that only this kernel touches; `barrier_base` its value before the
// launch (the host adds decode_program_barrier_arrivals(program) per launch). `trace` (optional, device, n_phases + 1 u64)
// receives %globaltimer of CTA 0 at every phase boundary.
cudaError_t decode_program_launch(const DecodeProgram & program, unsigned long long * barrier_counter, unsigned long long barrier_base,
                                  unsigned long long * trace, cudaStream_t stream);
inline unsigned long long decode_program_barrier_arrivals(const DecodeProgram & p) {
    return p.n_phases > 0 ? (unsigned long long) (p.n_phases - 1) * (unsigned long long) p.grid : 0ull;
}

}  // namespace rwkv
