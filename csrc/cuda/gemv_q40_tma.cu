// TMA-fed decode GEMV: y[t][d] = W_q40[d][n] · q80(x[t][n]), 1..8 tokens, fused prologue/epilogue.
//
// Same contract as gemv_q40.cu (which stays as the fallback for shapes the bulk-copy alignment rules reject), but
// the weight stream is moved by the TMA unit instead of by per-thread loads:
//   * warp 16 (one elected lane) is the producer: it cuts the CTA's contiguous row tile into stages of ~16-32 KB
//     and issues `cp.async.bulk` (global -> shared, mbarrier complete_tx) for nibbles and scales; a ring of up to
//     ~190 KB per SM is in flight, none of it costing registers;
//   * the ring is filled *before* griddepcontrol.wait, i.e. while the previous kernel of the layer is still
//     running (PDL) — weights are constants, only the activations depend on the predecessor;
//   * warps 0-15 consume: per step a warp takes 4 rows x 32 quant blocks from shared memory, the activation block
//     of each lane is read once and reused for the 4 rows (dp4a), 4 partial sums are reduced with a 6-shuffle
//     transposing butterfly, per-(row,segment) partials are summed in fixed order (deterministic);
//   * full/empty mbarrier pairs recycle stages when the tile is larger than the ring (w13, logits).
#include "kernels.h"
#include "tma_common.cuh"

namespace dl {

enum { PRO_RMSNORM = 0, PRO_PLAIN = 1 };
enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_ARGMAX = 3, EPI_MOE_DOWN = 4 };

struct TmaGemvGeom {
    uint32_t stageRows, nStages, stageBytes, maxTileRows;
};

// NB == 1 (decode) is compiled for two resident CTAs per SM: while kernel N computes, the CTAs of kernel N+1 are already
// on the SM (PDL) with their ring filling, so consecutive kernels of a layer overlap load and compute.
template <int PRO, int EPI, int NB>
__global__ void __launch_bounds__(kTmaThreads, (NB == 1 ? 2 : 1)) gemvQ40TmaKernel(GemvArgs a, TmaGemvGeom geo) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nblk = a.n / 32;
    const uint32_t nseg = (nblk + 31) / 32;
    const uint32_t rowQsBytes = nblk * 16, rowScBytes = nblk * 2;

    // ---- tile (pair aligned, like the fallback kernel); MoE launches carry one tile grid per routing slot ----
    const bool moe = a.moeCtasPerSlot != 0;
    const uint32_t nTiles = moe ? a.moeCtasPerSlot : gridDim.x;
    const uint32_t tileIdx = moe ? blockIdx.x % a.moeCtasPerSlot : blockIdx.x;
    const uint32_t slot = moe ? blockIdx.x / a.moeCtasPerSlot : 0;
    const uint32_t nPairs = a.d / 2;
    const uint32_t pairBegin = (uint32_t)(((uint64_t)tileIdx * nPairs) / nTiles);
    const uint32_t pairEnd = (uint32_t)(((uint64_t)(tileIdx + 1) * nPairs) / nTiles);
    const uint32_t rowBase = pairBegin * 2;
    const uint32_t tileRows = (pairEnd - pairBegin) * 2;
    const uint32_t SR = geo.stageRows;
    const uint32_t nFills = (tileRows + SR - 1) / SR;

    // ---- shared memory carve-up: [ring stages][planes][dx][dx8][partial][red][barriers] ----
    uint8_t *ring = smem;
    uint4 *planeA = reinterpret_cast<uint4 *>(ring + (size_t)geo.nStages * geo.stageBytes);
    uint4 *planeB = planeA + (size_t)NB * nblk;
    float *dxs = reinterpret_cast<float *>(planeB + (size_t)NB * nblk);
    float *dx8 = dxs + (size_t)NB * nblk;
    float *partial = dx8 + (size_t)NB * nblk;                               // [maxTileRows(+2)][nseg][NB]
    float *red = partial + (size_t)(geo.maxTileRows + 2) * nseg * NB;        // [32]: 16 reduction slots + 16 arg-max indices
    uint64_t *fullBar = reinterpret_cast<uint64_t *>(red + 32);             // [nStages]
    uint64_t *emptyBar = fullBar + kMaxStages;                              // [nStages]

    traceStamp(a.trace, 0);
    pdlLaunchDependents();
    if (tid == 0) {
        for (uint32_t s = 0; s < geo.nStages; s++) {
            mbarInit(&fullBar[s], 1);
            mbarInit(&emptyBar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // =============================== producer ===============================
        if (lane == 0) {
            const uint8_t *qsBase = reinterpret_cast<const uint8_t *>(a.qs);
            const uint8_t *scBase = reinterpret_cast<const uint8_t *>(a.scales);
            if (moe) {
                pdlWait();   // the expert choice is produced by the router kernel
                const uint32_t e = (uint32_t)a.expertIdx[slot] - a.moeFirstExpert;
                if (e >= a.moeNumLocal) return;   // expert lives on another rank (expert parallelism): nothing to stream
                qsBase += (uint64_t)e * a.expertQsStride * 4;
                scBase += (uint64_t)e * a.expertScaleStride * 2;
            }
            const uint64_t policy = policyEvictFirst();
            uint32_t st = 0, par = 0;
            bool wrapped = false;
            for (uint32_t f = 0; f < nFills; f++) {
                if (wrapped) mbarWait(&emptyBar[st], par ^ 1u);
                const uint32_t r0 = f * SR;
                const uint32_t rows = min(SR, tileRows - r0);
                const uint32_t bq = rows * rowQsBytes, bs = rows * rowScBytes;
                uint8_t *dst = ring + (size_t)st * geo.stageBytes;
                mbarExpectTx(&fullBar[st], bq + bs);
                tmaBulkLoad(dst, qsBase + (uint64_t)(rowBase + r0) * rowQsBytes, bq, &fullBar[st], policy);
                tmaBulkLoad(dst + (size_t)SR * rowQsBytes, scBase + (uint64_t)(rowBase + r0) * rowScBytes, bs, &fullBar[st], policy);
                if (++st == geo.nStages) { st = 0; par ^= 1u; wrapped = true; }
            }
        }
        return;
    }

    // =============================== consumers ===============================
    // norm weights are constants: pull them towards L2 while we wait for the predecessor
    if (PRO == PRO_RMSNORM) {
        for (uint32_t i = tid * 32; i < a.n; i += kConsumerThreads * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.normW + i));
    }
    pdlWait();
    traceStamp(a.trace, 1);
    const bool active = !moe || ((uint32_t)a.expertIdx[slot] - a.moeFirstExpert) < a.moeNumLocal;
    const float *inBase = a.in + (size_t)slot * a.inSlotStride;

    // ---- prologue: (rmsnorm) + q80 quantisation into the dp4a plane layout ----
    if (active) {
        const uint32_t nVec = a.n / 4;
#pragma unroll 1
        for (int t = 0; t < NB; t++) {
            const float4 *x4 = reinterpret_cast<const float4 *>(inBase + (size_t)t * a.inStride);
            float inv = 1.f;
            if (PRO == PRO_RMSNORM) {
                float ss = 0.f;
                for (uint32_t i = tid; i < nVec; i += kConsumerThreads) {
                    const float4 v = x4[i];
                    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                }
                ss = consumerSum(ss, red);
                inv = rsqrtf(ss / (float)a.n + a.eps);
            }
            uint8_t *pa = reinterpret_cast<uint8_t *>(planeA + (size_t)t * nblk);
            uint8_t *pb = reinterpret_cast<uint8_t *>(planeB + (size_t)t * nblk);
            for (uint32_t base = 0; base < nVec; base += kConsumerThreads) {
                const uint32_t i = base + tid;
                const bool act = i < nVec;
                float4 v = act ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                if (PRO == PRO_RMSNORM && act) {
                    const float4 w = reinterpret_cast<const float4 *>(a.normW)[i];
                    v.x = w.x * (v.x * inv); v.y = w.y * (v.y * inv); v.z = w.z * (v.z * inv); v.w = w.w * (v.w * inv);
                }
                float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                const float d = amax / 127.f;
                const float id = d != 0.f ? 1.f / d : 0.f;
                const int q0 = __float2int_rn(v.x * id), q1 = __float2int_rn(v.y * id);
                const int q2 = __float2int_rn(v.z * id), q3 = __float2int_rn(v.w * id);
                int qsum = q0 + q1 + q2 + q3;
                qsum += __shfl_xor_sync(0xffffffffu, qsum, 1);
                qsum += __shfl_xor_sync(0xffffffffu, qsum, 2);
                qsum += __shfl_xor_sync(0xffffffffu, qsum, 4);
                if (act) {
                    const uint32_t b = i >> 3, sub = i & 7, k = sub >> 1, odd = sub & 1;
                    uint8_t *wa = pa + (size_t)b * 16 + k * 4 + odd;
                    uint8_t *wb = pb + (size_t)b * 16 + k * 4 + odd;
                    wa[0] = (uint8_t)(int8_t)q0; wa[2] = (uint8_t)(int8_t)q1;
                    wb[0] = (uint8_t)(int8_t)q2; wb[2] = (uint8_t)(int8_t)q3;
                    if (sub == 0) {
                        const float dq = __half2float(__float2half_rn(d));
                        dxs[(size_t)t * nblk + b] = dq;
                        dx8[(size_t)t * nblk + b] = dq * 8.f * (float)qsum;
                    }
                }
            }
        }
    }
    consumerBarrier();
    traceStamp(a.trace, 2);

    // ---- main loop over ring stages ----
    // stage index / parity / step rotation are carried incrementally: no integer division inside the fill loop
    const uint32_t stepsPerFullStage = (SR / kRowsPerStep) * nseg;
    const uint32_t rotInc = stepsPerFullStage % kConsumerWarps;
    const uint32_t gInc = kConsumerWarps / nseg, segInc = kConsumerWarps - gInc * nseg;
    const uint32_t recipNseg = 65536u / nseg + 1u;   // (s * recip) >> 16 == s / nseg for s < 16, nseg <= 16
    uint32_t st = 0, par = 0, rot = 0;
    for (uint32_t f = 0; active && f < nFills; f++) {
        const uint32_t r0 = f * SR;
        const uint32_t rows = min(SR, tileRows - r0);
        const uint32_t nGroups = (rows + kRowsPerStep - 1) / kRowsPerStep;
        const uint32_t nSteps = nGroups * nseg;
        const uint8_t *stage = ring + (size_t)st * geo.stageBytes;
        const uint4 *sq = reinterpret_cast<const uint4 *>(stage);
        const uint16_t *ss = reinterpret_cast<const uint16_t *>(stage + (size_t)SR * rowQsBytes);
        // steps are dealt round-robin over the 16 warps *across* fills (a stage may hold fewer than 16 steps)
        const uint32_t firstStep = ((uint32_t)warp - rot) & (kConsumerWarps - 1);
        rot = (rot + rotInc) & (kConsumerWarps - 1);
        // every warp waits (even one without steps in this fill): it keeps all warps within one ring revolution, so no
        // warp can arrive twice on the same empty-barrier phase
        mbarWait(&fullBar[st], par);
        uint32_t g = (firstStep * recipNseg) >> 16, seg = firstStep - g * nseg;
        for (uint32_t s = firstStep; s < nSteps; s += kConsumerWarps) {
            const uint32_t blk = seg * 32 + lane;
            const uint32_t rl = g * kRowsPerStep;                  // first row of the group inside the stage
            float acc[kRowsPerStep][NB];
#pragma unroll
            for (int r = 0; r < kRowsPerStep; r++)
#pragma unroll
                for (int t = 0; t < NB; t++) acc[r][t] = 0.f;
            if (blk < nblk) {
                uint4 A[NB], B[NB];
                float dxv[NB], dx8v[NB];
#pragma unroll
                for (int t = 0; t < NB; t++) {
                    A[t] = planeA[t * nblk + blk];
                    B[t] = planeB[t * nblk + blk];
                    dxv[t] = dxs[t * nblk + blk];
                    dx8v[t] = dx8[t * nblk + blk];
                }
                // Rows past the end of a short last stage read stale ring bytes (always inside the stage buffer because
                // stageRows % 4 == 0); their sums are simply not stored below, so no branch is needed here.
                const uint4 *qp = sq + rl * nblk + blk;
                const uint16_t *sp = ss + rl * nblk + blk;
#pragma unroll
                for (int r = 0; r < kRowsPerStep; r++) {
                    const uint4 q = qp[r * nblk];
                    const float dw = __half2float(__ushort_as_half(sp[r * nblk]));
                    // low nibbles: q & 0x0f0f0f0f; high nibbles are used in place (q & 0xf0f0f0f0 = 16 * nibble), the factor
                    // 16 is removed with one shift after the dot product (exact: the partial sum is a multiple of 16)
                    const uint32_t ml = 0x0f0f0f0fu, mh = 0xf0f0f0f0u;
                    const uint32_t l0 = q.x & ml, h0 = q.x & mh, l1 = q.y & ml, h1 = q.y & mh;
                    const uint32_t l2 = q.z & ml, h2 = q.z & mh, l3 = q.w & ml, h3 = q.w & mh;
#pragma unroll
                    for (int t = 0; t < NB; t++) {
                        int lo = dp4a(l0, A[t].x, 0);
                        int hi = dp4a(h0, B[t].x, 0);
                        lo = dp4a(l1, A[t].y, lo);
                        hi = dp4a(h1, B[t].y, hi);
                        lo = dp4a(l2, A[t].z, lo);
                        hi = dp4a(h2, B[t].z, hi);
                        lo = dp4a(l3, A[t].w, lo);
                        hi = dp4a(h3, B[t].w, hi);
                        const int dot = lo + (hi >> 4);
                        acc[r][t] = dw * (dxv[t] * (float)dot - dx8v[t]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < NB; t++) {
                const float v = reduce4(acc[0][t], acc[1][t], acc[2][t], acc[3][t], lane);
                if ((lane & 7) == 0) {
                    const uint32_t r = lane >> 3;
                    if (rl + r < rows) partial[((r0 + rl + r) * nseg + seg) * NB + t] = v;
                }
            }
            g += gInc;
            seg += segInc;
            if (seg >= nseg) { seg -= nseg; g++; }
        }
        __syncwarp();
        if (lane == 0) mbarArrive(&emptyBar[st]);
        if (++st == geo.nStages) { st = 0; par ^= 1u; }
    }
    consumerBarrier();

    // ---- epilogue ----
    auto rowSum = [&](uint32_t r, uint32_t t) {
        float v = 0.f;
        for (uint32_t sg = 0; sg < nseg; sg++) v += partial[((size_t)r * nseg + sg) * NB + t];
        return v;
    };
    if (EPI == EPI_SWIGLU) {
        const uint32_t tilePairs = tileRows / 2;
        float *outBase = a.out + (size_t)slot * a.outSlotStride;
        for (uint32_t i = tid; i < tilePairs * NB; i += kConsumerThreads) {
            const uint32_t p = i / NB, t = i - p * NB;
            outBase[(size_t)t * a.outStride + pairBegin + p] = active ? gateAct(rowSum(2 * p, t), a.act) * rowSum(2 * p + 1, t) : 0.f;
        }
    } else if (EPI == EPI_RESIDUAL || EPI == EPI_MOE_DOWN) {
        __shared__ bool sLastSlot;
        if (EPI == EPI_MOE_DOWN) {
            // every routing slot leaves its weighted product in the scratch; the last slot CTA of a row tile sums the
            // slots in fixed order (deterministic) and carries on with the residual add / all-reduce
            const float wgt = active ? a.expertWeight[slot] : 0.f;
            for (uint32_t i = tid; i < tileRows; i += kConsumerThreads)
                a.moeScratch[(size_t)slot * a.d + rowBase + i] = active ? rowSum(i, 0) * wgt : 0.f;
            __threadfence();
            consumerBarrier();
            if (tid == 0) {
                const unsigned int prev = atomicAdd(&a.moeCounters[tileIdx], 1u);
                sLastSlot = prev == a.kActive - 1;
                if (sLastSlot) a.moeCounters[tileIdx] = 0;
            }
            consumerBarrier();
            if (!sLastSlot) { traceStamp(a.trace, 3); return; }
            __threadfence();
        }
        auto value = [&](uint32_t r, uint32_t t) {
            if (EPI == EPI_MOE_DOWN) {
                float v = 0.f;
                for (uint32_t j = 0; j < a.kActive; j++) v += __ldcg(a.moeScratch + (size_t)j * a.d + rowBase + r);
                return v;
            }
            return rowSum(r, t);
        };
        if (a.ar.nRanks > 1) {
            // ---- fused one-shot all-reduce over NVLink peer memory + residual add -------------------------------
            // Every rank computed the partial product of the same row tile in the CTA with the same index. LL protocol:
            // every 8-byte word carries (partial value, valid flag); the partial rows are stored into slot[myRank] of
            // *every* rank, receivers poll the words of all sources, sum them in rank order (bit-identical result on
            // every rank) and clear them for the all-reduce after next (the slots are double buffered by parity).
            const ArArgs &ar = a.ar;
            const size_t slotBase = (size_t)(ar.parity * ar.nRanks + ar.rank) * ar.slotStride;
            for (uint32_t i = tid; i < tileRows * NB; i += kConsumerThreads) {
                const uint32_t r = i / NB, t = i - r * NB;
                const float v = value(r, t);
                const size_t off = slotBase + (size_t)t * ar.dim + rowBase + r;
#pragma unroll 1
                for (uint32_t p = 0; p < ar.nRanks; p++) stLL(ar.slots[(ar.rank + p) % ar.nRanks] + off, __float_as_uint(v), 1u);
            }
            uint64_t *mine = ar.slots[ar.rank];
            for (uint32_t i = tid; i < tileRows * NB; i += kConsumerThreads) {
                const uint32_t r = i / NB, t = i - r * NB;
                float sum = 0.f;
                for (uint32_t sr = 0; sr < ar.nRanks; sr++) {
                    uint64_t *w = mine + (size_t)(ar.parity * ar.nRanks + sr) * ar.slotStride + (size_t)t * ar.dim + rowBase + r;
                    uint2 v = ldLL(w);
                    while (v.y == 0u) v = ldLL(w);
                    sum += __uint_as_float(v.x);
                    stLL(w, 0u, 0u);
                }
                a.out[(size_t)t * a.outStride + rowBase + r] += sum;
            }
        } else {
            for (uint32_t i = tid; i < tileRows * NB; i += kConsumerThreads) {
                const uint32_t r = i / NB, t = i - r * NB;
                a.out[(size_t)t * a.outStride + rowBase + r] += value(r, t);
            }
        }
    } else {
        float best = -INFINITY;
        int bestIdx = 0x7fffffff;
        for (uint32_t i = tid; i < tileRows * NB; i += kConsumerThreads) {
            const uint32_t r = i / NB, t = i - r * NB;
            const float v = rowSum(r, t);
            a.out[(size_t)t * a.outStride + rowBase + r] = v;
            if (EPI == EPI_ARGMAX && v > best && a.rowOffsetGlobal + rowBase + r < a.vocabLimit) { best = v; bestIdx = (int)(a.rowOffsetGlobal + rowBase + r); }   // rows ascend per thread
        }
        if (EPI == EPI_ARGMAX) {
            // greedy sampling fused into the logits kernel: CTA-level arg-max, then the last CTA to finish reduces
            // the per-CTA candidates, publishes the next token and advances the position (all on the device)
            auto better = [](float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); };
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
            }
            float *sv = red;                                   // 16 floats
            int *si = reinterpret_cast<int *>(red + 16);
            consumerBarrier();
            if (lane == 0) { sv[warp] = best; si[warp] = bestIdx; }
            consumerBarrier();
            __shared__ bool lastCta;
            if (warp == 0) {
                best = lane < kConsumerWarps ? sv[lane] : -INFINITY;
                bestIdx = lane < kConsumerWarps ? si[lane] : 0x7fffffff;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                    if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
                }
                if (lane == 0) {
                    a.argVal[blockIdx.x] = best;
                    a.argIdx[blockIdx.x] = bestIdx;
                    __threadfence();
                    const unsigned int prev = atomicAdd(a.argCounter, 1u);
                    lastCta = prev == gridDim.x - 1;
                    if (lastCta) *a.argCounter = 0;
                }
            }
            consumerBarrier();
            if (lastCta && warp == 0) {
                __threadfence();
                best = -INFINITY;
                bestIdx = 0x7fffffff;
                for (uint32_t i = lane; i < gridDim.x; i += 32) {
                    const float v = __ldcg(a.argVal + i);
                    const int ix = __ldcg(a.argIdx + i);
                    if (better(v, ix, best, bestIdx)) { best = v; bestIdx = ix; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                    if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
                }
                if (a.ar.nRanks > 1) {
                    // vocabulary-parallel logits: exchange the per-rank winners over peer memory; every rank then
                    // picks the same global token, so no token broadcast is needed afterwards
                    const ArArgs &ar = a.ar;
                    if (lane < ar.nRanks) stLL(ar.cand[lane] + ar.rank, __float_as_uint(best), (uint32_t)bestIdx + 1u);
                    if (lane < ar.nRanks) {
                        uint64_t *w = ar.cand[ar.rank] + lane;
                        uint2 v = ldLL(w);
                        while (v.y == 0u) v = ldLL(w);
                        best = __uint_as_float(v.x);
                        bestIdx = (int)(v.y - 1u);
                        stLL(w, 0u, 0u);
                    } else {
                        best = -INFINITY;
                        bestIdx = 0x7fffffff;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                        if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
                    }
                }
                if (lane == 0) {
                    a.tokenOut[0] = bestIdx;
                    const int p = a.posInOut[0] + 1;
                    a.posInOut[0] = p;
                    if (a.history && (uint32_t)p < a.historyCap) a.history[p] = bestIdx;
                }
            }
        }
    }
    traceStamp(a.trace, 3);
}

template <int PRO, int EPI, int NB>
static int launchTma(const GemvArgs &a, const TmaGemvGeom &geo, int grid, size_t smemBytes, cudaStream_t stream, bool pdl) {
    auto kernel = gemvQ40TmaKernel<PRO, EPI, NB>;
    static size_t configured = 0;
    if (smemBytes > configured) {
        DL_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        configured = smemBytes;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTmaThreads);
    cfg.dynamicSmemBytes = smemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, a, geo));
    return 0;
}

// Returns 1 when the shape cannot use the TMA path (caller falls back), <0 on errors, 0 on success.
int gemvQ40Tma(int pro, int epi, int nb, GemvArgs a, int numSms, cudaStream_t stream, bool pdl) {
    if (a.d % 2 || a.n % 128) return 1;   // bulk copies need 16-byte aligned row starts for the fp16 scale rows
    a.act = gHiddenAct;
    if (a.vocabLimit == 0) a.vocabLimit = 0xffffffffu;
    const uint32_t nblk = a.n / 32, nseg = (nblk + 31) / 32;
    const uint32_t nPairs = a.d / 2;
    int grid = (int)(nPairs < (uint32_t)numSms ? nPairs : (uint32_t)numSms);
    uint32_t tilesPerMatrix = (uint32_t)grid;
    if (a.moeCtasPerSlot) {
        if (nb != 1) return -4;
        tilesPerMatrix = a.moeCtasPerSlot;
        grid = (int)(a.moeCtasPerSlot * a.kActive);
    }
    TmaGemvGeom geo{};
    geo.maxTileRows = 2 * ((nPairs + tilesPerMatrix - 1) / tilesPerMatrix);
    const uint32_t rowBytes = nblk * 18;
    // a stage should hold >= 16 steps (one per consumer warp) so the per-stage barrier traffic is amortised
    uint32_t sr = (64 + nseg - 1) / nseg;
    sr = (sr + 3) / 4 * 4;
    while (sr > 4 && sr * rowBytes > 36 * 1024) sr -= 4;
    if (sr * rowBytes > 48 * 1024) return 1;   // rows too wide for a 4-row stage: per-thread-load kernel handles it
    if (sr > 64) sr = 64;
    geo.stageRows = sr;
    geo.stageBytes = (sr * rowBytes + 127) / 128 * 128;
    const size_t fixedBytes = (size_t)nb * nblk * (16 + 16 + 4 + 4) + (size_t)(geo.maxTileRows + 2) * nseg * nb * 4 + 32 * 4 +
                              2 * kMaxStages * 8 + 256;
    const size_t budget = (nb == 1 ? 112 : 226) * 1024;   // nb == 1: leave room for the next kernel's CTA on the same SM
    if (fixedBytes + geo.stageBytes > budget) return 1;
    uint32_t stages = (uint32_t)((budget - fixedBytes) / geo.stageBytes);
    const uint32_t need = (geo.maxTileRows + sr - 1) / sr;
    if (stages > need) stages = need;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 1) return 1;
    geo.nStages = stages;
    const size_t smemBytes = fixedBytes + (size_t)stages * geo.stageBytes;
#define DL_TMA_CASE(P, E, N) \
    if (pro == P && epi == E && nb == N) return launchTma<P, E, N>(a, geo, grid, smemBytes, stream, pdl);
#define DL_TMA_NB(P, E) DL_TMA_CASE(P, E, 1) DL_TMA_CASE(P, E, 2) DL_TMA_CASE(P, E, 4) DL_TMA_CASE(P, E, 8)
    DL_TMA_NB(PRO_RMSNORM, EPI_STORE)
    DL_TMA_NB(PRO_PLAIN, EPI_RESIDUAL)
    DL_TMA_NB(PRO_RMSNORM, EPI_SWIGLU)
    DL_TMA_NB(PRO_PLAIN, EPI_STORE)
    DL_TMA_CASE(PRO_PLAIN, EPI_SWIGLU, 1)
    DL_TMA_CASE(PRO_RMSNORM, EPI_ARGMAX, 1)
    DL_TMA_CASE(PRO_PLAIN, EPI_MOE_DOWN, 1)
#undef DL_TMA_NB
#undef DL_TMA_CASE
    return -3;
}

}  // namespace dl
