// On-device sampling of the next token from the logits the head just produced, so that only a token id crosses PCIe
// (SURVEY.md 8f-4). Follows the reference's python/sampling.py:10-52 step for step: softmax -> (logit bias) -> temperature 0 =
// argmax -> top-p cutoff -> power by 1/temperature -> renormalise -> inverse-CDF draw, where the draw uses a uniform number `u`
// supplied by the caller exactly as numpy's RandomState.choice uses its one random_sample(): cdf = cumsum(p); cdf /= cdf[-1];
// index = searchsorted(cdf, u, side="right").
//
// One CTA of 1024 threads, thread t owns the PER consecutive vocabulary entries [t*PER, (t+1)*PER) in registers (n_vocab <= 65536).
// The top-p cutoff -- the value of the sorted probabilities at the first index whose running sum exceeds top_p (:42-45) -- is found
// without sorting: it is the largest value v for which sum(p[p >= v]) > top_p, located by bisection on the float bit pattern
// (31 masked block sums over registers).
#include "ops.h"
#include "gemv.h"

namespace rwkv {
namespace {

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_WARPS = SAMPLE_THREADS / 32;

__device__ __forceinline__ double warp_sum_dd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ double block_sum_dd(double v, double * slots) {      // every thread gets the total
    v = warp_sum_dd(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = v;
    __syncthreads();
    return warp_sum_dd(slots[threadIdx.x & 31]);
}
__device__ float block_max_f(float v, float * slots) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = slots[threadIdx.x & 31];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}

template <int PER>
__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const SampleParams p) {
    __shared__ double dslots[SAMPLE_WARPS];
    __shared__ float fslots[SAMPLE_WARPS];
    __shared__ int first_flagged;
    pdl_prologue();
    const int tid = threadIdx.x, V = p.n_vocab, base = tid * PER;
    const float * src = p.logits;
    if (p.n_bias > 0) {     // logits + bias through the scratch copy (a dict has no duplicate ids, :30-33)
        for (int i = tid; i < V; i += SAMPLE_THREADS) p.scratch[i] = p.logits[i];
        __syncthreads();
        for (int i = tid; i < p.n_bias; i += SAMPLE_THREADS) if (p.bias_ids[i] < (uint32_t) V) p.scratch[p.bias_ids[i]] += p.bias_values[i];
        __syncthreads();
        src = p.scratch;
    }
    float v[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        v[j] = (base + j < V) ? src[base + j] : -INFINITY;
        mx = fmaxf(mx, v[j]);
    }
    mx = block_max_f(mx, fslots);
    if (p.temperature == 0.0f) {    // :39-40 argmax, first index among equals
        int best = V;
#pragma unroll
        for (int j = 0; j < PER; j++) if (base + j < V && v[j] == mx && base + j < best) best = base + j;
        if (tid == 0) first_flagged = V;
        __syncthreads();
        if (best < V) atomicMin(&first_flagged, best);
        __syncthreads();
        if (tid == 0) { *p.token_out = (uint32_t) first_flagged; if (p.prob_out) *p.prob_out = 1.0f; }
        return;
    }
    // softmax (:5-8)
    double s = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { v[j] = (base + j < V) ? expf(v[j] - mx) : 0.f; s += (double) v[j]; }
    const float total = (float) block_sum_dd(s, dslots);
#pragma unroll
    for (int j = 0; j < PER; j++) v[j] = __fdiv_rn(v[j], total);
    // top-p (:42-45)
    const float top_p = (p.top_p == 0.0f) ? 1.0f : p.top_p;
    if (top_p < 1.0f) {
        auto mass_at_least = [&](float thr) {
            double m = 0;
#pragma unroll
            for (int j = 0; j < PER; j++) m += (v[j] >= thr) ? (double) v[j] : 0.0;
            return block_sum_dd(m, dslots);
        };
        float cutoff;
        if (!(mass_at_least(0.0f) > (double) top_p)) {
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < PER; j++) m2 = fmaxf(m2, v[j]);
            cutoff = block_max_f(m2, fslots);       // np.argmax of an all-False mask is 0: the largest probability
        } else {
            uint32_t lo = 0u, hi = 0x7F800000u;     // f(lo) > top_p >= f(hi)
            while (hi - lo > 1u) {
                const uint32_t mid = lo + (hi - lo) / 2u;
                if (mass_at_least(__uint_as_float(mid)) > (double) top_p) lo = mid; else hi = mid;
            }
            cutoff = __uint_as_float(lo);
        }
#pragma unroll
        for (int j = 0; j < PER; j++) if (v[j] < cutoff) v[j] = 0.f;
    }
    if (p.temperature != 1.0f) {    // :47-48
        const float inv_t = 1.0f / p.temperature;
#pragma unroll
        for (int j = 0; j < PER; j++) v[j] = (v[j] > 0.f) ? powf(v[j], inv_t) : 0.f;
    }
    double s2 = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) s2 += (double) v[j];
    const float total2 = (float) block_sum_dd(s2, dslots);
    // :50-52  p / sum(p), then RandomState.choice's inverse CDF in double
    double mine = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { v[j] = __fdiv_rn(v[j], total2); mine += (double) v[j]; }
    // exclusive prefix of the per-thread masses: warp scan, then a scan of the 32 warp totals
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += up;
    }
    __syncthreads();
    if ((tid & 31) == 31) dslots[tid >> 5] = incl;
    if (tid == 0) first_flagged = SAMPLE_THREADS;
    __syncthreads();
    double warp_base = 0, grand = 0;
    for (int w = 0; w < SAMPLE_WARPS; w++) { const double t = dslots[w]; if (w < (tid >> 5)) warp_base += t; grand += t; }
    const double before = warp_base + incl - mine;
    if ((before + mine) / grand > p.u) atomicMin(&first_flagged, tid);
    __syncthreads();
    const int owner = first_flagged;
    if (owner == SAMPLE_THREADS) {                 // u >= cdf[-1] (only for u == 1.0): the last token with any mass
        if (tid == 0) { *p.token_out = (uint32_t) (V - 1); if (p.prob_out) *p.prob_out = 0.f; }
        return;
    }
    if (tid == owner) {
        double c = before;
        int pick = min(base + PER - 1, V - 1);
        float pp = 0.f;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            c += (double) v[j];
            if (c / grand > p.u) { pick = base + j; pp = v[j]; break; }
        }
        *p.token_out = (uint32_t) min(pick, V - 1);
        if (p.prob_out) *p.prob_out = pp;
    }
}

}  // namespace

cudaError_t launch_sample(const SampleParams & p, cudaStream_t s) {
    if (p.n_vocab <= 0 || p.n_vocab > SAMPLE_THREADS * 64) return cudaErrorInvalidValue;
    g_kernel_launches++;
    const int per = (p.n_vocab + SAMPLE_THREADS - 1) / SAMPLE_THREADS;
    if (per <= 1) return launch_pdl(sample_kernel<1>, dim3(1), dim3(SAMPLE_THREADS), 0, s, p);
    if (per <= 4) return launch_pdl(sample_kernel<4>, dim3(1), dim3(SAMPLE_THREADS), 0, s, p);
    if (per <= 16) return launch_pdl(sample_kernel<16>, dim3(1), dim3(SAMPLE_THREADS), 0, s, p);
    if (per <= 50) return launch_pdl(sample_kernel<50>, dim3(1), dim3(SAMPLE_THREADS), 0, s, p);
    return launch_pdl(sample_kernel<64>, dim3(1), dim3(SAMPLE_THREADS), 0, s, p);
}

}  // namespace rwkv
