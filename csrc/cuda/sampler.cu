// Temperature / top-p sampling on the device: the logits never leave the GPU.
//
// Semantics = the reference's Sampler::sample (src/tokenizer.cpp:426-512; host twin: csrc/host/text.cpp Sampler::sample):
//   logits / T -> softmax -> coin = xorshift* F32 -> multinomial CDF walk (topp outside (0,1)) or nucleus sampling: drop
//   probabilities below (1 - topp) / (n - 1), order the rest by (probability desc, index asc), cut at the first prefix whose sum
//   exceeds topp, draw r = coin * prefix sum and walk the order until the running sum exceeds r.
// No sort is materialised: both "first prefix whose sum exceeds X" questions are answered by a radix descent over the float keys
// (4 levels x 8 bits, a 256-bin histogram of probability mass per level) followed by a rank selection among equal keys. All
// prefix arithmetic is 2^-40 fixed point in 64-bit integers, so histogram atomics are order independent: the sampler is bit
// reproducible run to run and across ranks — under tensor parallelism every rank samples the same token from the gathered
// logits with its own copy of the generator state, and no token broadcast is needed.
//
// Tensor parallelism (K4, reference: gather of vocab slices to the root, src/llm.cpp:587-599): logitsGatherKernel pushes this
// rank's vocabulary slice into every rank's gather buffer (one multimem.st per 16 bytes through the NVSwitch multicast mapping, or
// unicast peer stores) and bumps an arrival counter on every rank; the sampler waits for nRanks arrivals. No NCCL on the step path.
#include "../common/dl_expf.h"
#include "kernels.h"

namespace dl {
namespace {

constexpr int kSampThreads = 1024;
constexpr float kFix = 1099511627776.0f;   // 2^40

__device__ __forceinline__ unsigned long long toFix(float p) { return (unsigned long long)(p * kFix); }

struct SampleArgs {
    const float *logits;       // [n] (full vocabulary on this rank: local buffer or the gather buffer)
    float *probs;              // [n] scratch: exp((l - max) / T)
    uint32_t n;
    float temperature, topp;
    unsigned long long *rng;   // xorshift* state
    int *tokenOut, *pos, *history;
    uint32_t historyCap;
    // tensor parallelism: wait until every rank's slice has arrived in `logits`
    const unsigned int *gatherFlag;
    unsigned int *gatherEpoch;     // device counter of completed gathers (the kernel advances it)
    uint32_t nRanks;
    int *debugOut;             // optional [4]: {last index in order?, ...} unused by the product path
};

__device__ float blockMax(float v, float *red) {
    v = warpMax(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[lane];          // 32 warps
    t = warpMax(t);
    return t;
}

// Deterministic block sum: fixed tree over warps.
__device__ float blockSum(float v, float *red) {
    v = warpSum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[lane];
    t = warpSum(t);
    return t;
}

// Exclusive prefix over the 1024 threads (thread order), for 64-bit integers; returns the exclusive prefix, *total gets the sum.
__device__ unsigned long long blockExclusiveScan(unsigned long long v, unsigned long long *warpTotals, unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) warpTotals[warp] = inc;
    __syncthreads();
    unsigned long long wt = warpTotals[lane];
    unsigned long long winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    const unsigned long long warpExcl = __shfl_sync(0xffffffffu, winc - wt, warp);
    if (total) *total = __shfl_sync(0xffffffffu, winc, 31);
    return warpExcl + inc - v;
}

// Finds, in the order (key desc, index asc) over the candidates (prob >= cutoff), the first element whose inclusive prefix sum
// (fixed point) exceeds `target`. Outputs its index and that prefix sum. If the total candidate mass does not exceed the target,
// returns the LAST element of the order and the total (the reference's loop falls off its end the same way).
struct Found { int index; unsigned long long prefix; };

__device__ Found radixFind(const SampleArgs &a, float inv, float cutoff, unsigned long long target, unsigned long long *hist /*[256]*/,
                           unsigned long long *scratch /*[64]*/, int *sInt /*[8]*/) {
    const uint32_t n = a.n;
    uint32_t prefix = 0;                 // decided high bits of the key
    unsigned long long base = 0;         // mass of all candidates whose key is above the current prefix range
    bool exhausted = false;              // the whole candidate mass is <= target
    for (int level = 0; level < 4; level++) {
        const int shift = 24 - 8 * level;
        for (int b = threadIdx.x; b < 256; b += kSampThreads) hist[b] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += kSampThreads) {
            const float p = a.probs[i] * inv;
            if (p < cutoff) continue;
            const uint32_t key = __float_as_uint(p);
            if (level > 0 && (key >> (shift + 8)) != prefix) continue;
            atomicAdd(&hist[(key >> shift) & 255u], toFix(p));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long run = base;
            int chosen = -1, lowestNonEmpty = -1;
            for (int b = 255; b >= 0; b--) {
                if (hist[b] == 0) continue;
                lowestNonEmpty = b;
                if (run + hist[b] > target) { chosen = b; break; }
                run += hist[b];
            }
            if (chosen < 0) {            // never exceeds: continue towards the smallest key so that the last element is found
                chosen = lowestNonEmpty;
                run -= (lowestNonEmpty >= 0 ? hist[lowestNonEmpty] : 0);
                sInt[1] = 1;
            } else {
                sInt[1] = 0;
            }
            sInt[0] = chosen;
            scratch[0] = run;
        }
        __syncthreads();
        if (sInt[0] < 0) return Found{(int)n - 1, base};     // no candidate at all (cannot happen for n >= 2: max prob >= 1/n > cutoff)
        exhausted = exhausted || sInt[1] != 0;
        prefix = (prefix << 8) | (uint32_t)sInt[0];
        base = scratch[0];
        __syncthreads();
    }
    // all elements with key == prefix share one probability p*: pick the j-th in index order
    const float pStar = __uint_as_float(prefix);
    const unsigned long long fStar = toFix(pStar);
    // per-thread contiguous index chunks keep the index order
    const uint32_t chunk = (n + kSampThreads - 1) / kSampThreads;
    const uint32_t lo = threadIdx.x * chunk, hi = min(lo + chunk, n);
    unsigned long long cnt = 0;
    for (uint32_t i = lo; i < hi; i++) cnt += (__float_as_uint(a.probs[i] * inv) == prefix) ? 1ull : 0ull;
    unsigned long long totalTies = 0;
    const unsigned long long before = blockExclusiveScan(cnt, scratch + 8, &totalTies);
    // number of ties needed: smallest j >= 1 with base + j * f* > target (all of them when the mass is exhausted)
    unsigned long long need;
    if (exhausted || base + totalTies * fStar <= target) need = totalTies;
    else need = (target - base) / fStar + 1;       // base <= target here
    if (need < 1) need = 1;
    if (need > totalTies) need = totalTies;
    __syncthreads();
    if (before < need && need <= before + cnt) {
        unsigned long long seen = before;
        for (uint32_t i = lo; i < hi; i++) {
            if (__float_as_uint(a.probs[i] * inv) == prefix && ++seen == need) { sInt[2] = (int)i; break; }
        }
    }
    __syncthreads();
    return Found{sInt[2], base + need * fStar};
}

__global__ void __launch_bounds__(kSampThreads, 1) sampleKernel(SampleArgs a) {
    __shared__ float red[32];
    __shared__ unsigned long long hist[256];
    __shared__ unsigned long long scratch[64];
    __shared__ int sInt[8];
    __shared__ float sCoin;
    const uint32_t n = a.n;
    if (a.gatherFlag && a.nRanks > 1) {
        if (threadIdx.x == 0) {
            const unsigned int want = (*a.gatherEpoch + 1u) * a.nRanks;
            unsigned int v;
            uint32_t spins = 0;
            do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.gatherFlag) : "memory"); } while ((int)(v - want) < 0 && ++spins < (1u << 28));
            *a.gatherEpoch += 1u;
        }
        __syncthreads();
    }
    // softmax statistics (IEEE division / shared exp / integer normaliser: the host sampler computes the very same numbers)
    float m = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += kSampThreads) m = fmaxf(m, __fdiv_rn(a.logits[i], a.temperature));
    m = blockMax(m, red);
    // per-thread sums over contiguous chunks (index order), combined by a fixed tree
    const uint32_t chunk = (n + kSampThreads - 1) / kSampThreads;
    const uint32_t lo = threadIdx.x * chunk, hi = min(lo + chunk, n);
    unsigned long long sFix = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const float e = expNeg(__fsub_rn(__fdiv_rn(a.logits[i], a.temperature), m));
        a.probs[i] = e;
        sFix += toFix(e);
    }
    unsigned long long sTotal = 0;
    (void)blockExclusiveScan(sFix, scratch + 8, &sTotal);
    const float inv = __fdiv_rn(1.0f, __fdiv_rn((float)sTotal, kFix));
    if (threadIdx.x == 0) {
        unsigned long long st = *a.rng;          // xorshift* (reference src/tokenizer.cpp:25-36)
        st ^= st >> 12; st ^= st << 25; st ^= st >> 27;
        *a.rng = st;
        const uint32_t u = (uint32_t)((st * 0x2545F4914F6CDD1Dull) >> 32);
        sCoin = (float)(u >> 8) * 5.9604644775390625e-8f;   // / 2^24, exact
        sInt[2] = (int)n - 1;
    }
    __syncthreads();
    const float coin = sCoin;
    int token;
    if (a.topp <= 0.f || a.topp >= 1.f) {
        // multinomial walk in index order
        unsigned long long mine = 0;
        for (uint32_t i = lo; i < hi; i++) mine += toFix(a.probs[i] * inv);
        unsigned long long total = 0;
        const unsigned long long before = blockExclusiveScan(mine, scratch + 8, &total);
        const unsigned long long c = toFix(coin);
        if (before <= c && c < before + mine) {
            unsigned long long cdf = before;
            for (uint32_t i = lo; i < hi; i++) {
                cdf += toFix(a.probs[i] * inv);
                if (c < cdf) { sInt[2] = (int)i; break; }
            }
        }
        __syncthreads();
        token = sInt[2];
    } else {
        const float cutoff = __fdiv_rn(1.0f - a.topp, (float)(n - 1));
        const Found cut = radixFind(a, inv, cutoff, toFix(a.topp), hist, scratch, sInt);
        // r = coin * cumulative (the reference multiplies the float prefix sum)
        const float cumulative = __fdiv_rn((float)cut.prefix, kFix);
        const unsigned long long r = toFix(coin * cumulative);
        __syncthreads();
        const Found pick = radixFind(a, inv, cutoff, r, hist, scratch, sInt);
        token = pick.index;
    }
    if (threadIdx.x == 0) {
        a.tokenOut[0] = token;
        if (a.pos) {
            const int p = a.pos[0] + 1;
            a.pos[0] = p;
            if (a.history && (uint32_t)p < a.historyCap) a.history[p] = token;
        }
    }
}

// Copies this rank's vocabulary slice into the gather buffer of every rank and signals its arrival.
__global__ void __launch_bounds__(256) logitsGatherKernel(const float *local, uint32_t v0, uint32_t rank, uint32_t nRanks, float *gatherMc,
                                                           float *const *gatherUc, unsigned int *flagMc, unsigned int *const *flagUc,
                                                           unsigned int *blockCounter) {
    const uint32_t nVec = v0 / 4;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nVec; i += gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4 *>(local)[i];
        const size_t off = (size_t)rank * v0 + (size_t)i * 4;
        if (gatherMc) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gatherMc + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
            for (uint32_t p = 0; p < nRanks; p++) *reinterpret_cast<float4 *>(gatherUc[p] + off) = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(blockCounter, 1u);
        if (prev == gridDim.x - 1) {
            *blockCounter = 0;
            __threadfence_system();
            if (flagMc) asm volatile("multimem.red.release.sys.global.add.u32 [%0], 1;" ::"l"(flagMc) : "memory");
            else for (uint32_t p = 0; p < nRanks; p++) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flagUc[p]) : "memory");
        }
    }
}

}  // namespace

int launchSample(const float *logits, float *probs, uint32_t n, float temperature, float topp, unsigned long long *rng, int *tokenOut, int *pos,
                 int *history, uint32_t historyCap, const unsigned int *gatherFlag, unsigned int *gatherEpoch, uint32_t nRanks,
                 cudaStream_t stream) {
    if (n < 2 || temperature <= 0.f) return -1;
    SampleArgs a{};
    a.logits = logits; a.probs = probs; a.n = n; a.temperature = temperature; a.topp = topp; a.rng = rng; a.tokenOut = tokenOut; a.pos = pos;
    a.history = history; a.historyCap = historyCap; a.gatherFlag = gatherFlag; a.gatherEpoch = gatherEpoch; a.nRanks = nRanks;
    sampleKernel<<<1, kSampThreads, 0, stream>>>(a);
    DL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launchLogitsGather(const float *local, uint32_t v0, uint32_t rank, uint32_t nRanks, float *gatherMc, float *const *gatherUcDev,
                       unsigned int *flagMc, unsigned int *const *flagUcDev, unsigned int *blockCounter, cudaStream_t stream) {
    if (v0 % 4) return -1;
    logitsGatherKernel<<<32, 256, 0, stream>>>(local, v0, rank, nRanks, gatherMc, gatherUcDev, flagMc, flagUcDev, blockCounter);
    DL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace dl

// Stand-alone entry for tests: samples one token from `logits` (device f32 [n]); rngState is a device u64.
DL_EXPORT int dl_sample_logits(const float *logits, float *probsScratch, uint32_t n, float temperature, float topp, unsigned long long *rngState,
                               int *tokenOut, cudaStream_t stream) {
    return dl::launchSample(logits, probsScratch, n, temperature, topp, rngState, tokenOut, nullptr, nullptr, 0, nullptr, nullptr, 1, stream);
}
