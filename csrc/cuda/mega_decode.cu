// Persistent single-token decode kernel: one launch per generated token for dense models (Llama / Qwen3).
//
// The multi-kernel decode path pays ~3 µs of grid-completion latency at each of its 161 kernel boundaries and restarts the
// weight stream at every launch. Here one CTA per SM stays resident for the whole token:
//   * warp 16 (producer) walks the list of all weight matrices of the token and streams this CTA's row tile of each one
//     through a single shared-memory ring with cp.async.bulk + mbarriers. It never synchronises with the phases — only
//     ring space limits it — so the HBM stream runs continuously across phase boundaries and layers.
//   * warps 0-15 (consumers) execute the phases embedding -> per layer [QKV GEMV | attention | WO GEMV + residual |
//     W1|W3 GEMV + SwiGLU | W2 GEMV + residual] -> logits GEMV + arg-max, separated by software grid barriers (one atomic
//     arrive + acquire spin, ~1 µs). Activation reads bypass L1 (ld.global.cg): L1 is not invalidated inside a launch.
//   * tensor-parallel runs use the same LL-word all-reduce in the WO / W2 epilogues and the cross-rank arg-max.
// The kernel takes over the roles of NnExecutor's step loop + barriers (reference src/nn/nn-executor.cpp:137-175) on the GPU.
#include "kernels.h"
#include "tma_common.cuh"

namespace dl {

enum { MP_QKV = 0, MP_WO = 1, MP_W13 = 2, MP_W2 = 3, MP_LOGITS = 4 };   // index into MegaArgs::ph

static uint32_t megaStageRows(uint32_t n, uint32_t stageBytes) {
    const uint32_t rowBytes = (n / 32) * 18;
    uint32_t sr = stageBytes / rowBytes;
    sr = sr / 4 * 4;
    return sr > 64 ? 64 : sr;
}

struct MegaSmem {
    uint8_t *ring;
    uint4 *planeA, *planeB;
    float *dxs, *dx8, *partial, *red;
    uint64_t *fullBar, *emptyBar;
};

__device__ __forceinline__ void gridBarrier(unsigned int *ctr, unsigned int &target, int tid) {
    consumerBarrier();   // orders every consumer thread's global writes before thread 0's release below
    if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        target += gridDim.x;
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    consumerBarrier();
}

__device__ __forceinline__ float4 ldcg4(const float4 *p) { return __ldcg(p); }

// Ring position shared (by construction) by the producer and the consumers: stage index + mbarrier phase parity, advanced
// once per fill. Kept incrementally — the hot loop contains no integer division (every `%`/`/` by a runtime value costs a
// ~20-instruction I2F/MUFU.RCP/F2I sequence per warp per fill, which used to dominate the per-fill overhead).
struct RingPos {
    uint32_t st, par;
    __device__ __forceinline__ void advance(uint32_t nStages) {
        if (++st == nStages) { st = 0; par ^= 1u; }
    }
};

// This CTA's pair-aligned row tile of a phase: the first `pairsRem` CTAs own one pair more (host-computed quotient/remainder).
__device__ __forceinline__ void megaTile(const MegaPhase &P, uint32_t &pairBegin, uint32_t &tileRows) {
    const uint32_t b = blockIdx.x;
    pairBegin = b * P.pairsQ + min(b, P.pairsRem);
    tileRows = 2 * (P.pairsQ + (b < P.pairsRem ? 1u : 0u));
}

// One GEMV phase on the consumer warps.
template <int PRO, int EPI>
__device__ void megaGemv(const MegaArgs &m, const MegaSmem &sm, const MegaPhase &P, const float *in, const float *normW, float *out,
                         uint32_t arParity, RingPos &ring, int tid, uint32_t &slot) {
    auto stamp = [&]() { if (m.trace && tid == 0 && blockIdx.x < m.traceCtas) m.trace[(size_t)blockIdx.x * m.traceStride + slot] = globalTimerNs(); slot++; };
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t n = P.n;
    const uint32_t nblk = P.nblk, nseg = P.nseg;
    const uint32_t rowQsBytes = nblk * 16;
    uint32_t pairBegin, tileRows;
    megaTile(P, pairBegin, tileRows);
    const uint32_t rowBase = pairBegin * 2;
    const uint32_t SR = P.stageRows;
    uint4 *planeA = sm.planeA, *planeB = sm.planeB;
    float *dxs = sm.dxs, *dx8 = sm.dx8, *partial = sm.partial, *red = sm.red;

    // ---- prologue: (rmsnorm) + q80 quantisation of the activation vector (one pass: the vector stays in registers) ----
    {
        const uint32_t nVec = n / 4;
        const float4 *x4 = reinterpret_cast<const float4 *>(in);
        // plain phases: 8 float4 x 512 threads = 16384 elements; RMS-norm phases (n = dim <= 8192) hold the norm weights in the
        // other half of that register budget, loaded together with x so only one L2 round trip sits on the critical path
        constexpr int kMaxVec = PRO == PRO_RMSNORM_ ? 4 : 8;
        float4 xv[kMaxVec];
        float4 wv[PRO == PRO_RMSNORM_ ? kMaxVec : 1];
        if (PRO == PRO_RMSNORM_) {
#pragma unroll
            for (int k = 0; k < kMaxVec; k++) {
                const uint32_t i = k * kConsumerThreads + tid;
                wv[k] = i < nVec ? __ldg(reinterpret_cast<const float4 *>(normW) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxVec; k++) {
            const uint32_t i = k * kConsumerThreads + tid;
            xv[k] = i < nVec ? ldcg4(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
        }
        float inv = 1.f;
        if (PRO == PRO_RMSNORM_) {
            ss = consumerSum(ss, red);
            inv = rsqrtf(ss / (float)n + m.eps);
        }
        uint8_t *pa = reinterpret_cast<uint8_t *>(planeA);
        uint8_t *pb = reinterpret_cast<uint8_t *>(planeB);
#pragma unroll
        for (int k = 0; k < kMaxVec; k++) {
            const uint32_t i = k * kConsumerThreads + tid;
            if (k * kConsumerThreads >= nVec) break;     // block-uniform
            const bool act = i < nVec;
            float4 v = xv[k];
            if (PRO == PRO_RMSNORM_ && act) {
                const float4 w = wv[PRO == PRO_RMSNORM_ ? k : 0];
                v.x = w.x * (v.x * inv); v.y = w.y * (v.y * inv); v.z = w.z * (v.z * inv); v.w = w.w * (v.w * inv);
            }
            float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float dq = amax / 127.f;
            const float id = dq != 0.f ? 1.f / dq : 0.f;
            const int q0 = __float2int_rn(v.x * id), q1 = __float2int_rn(v.y * id);
            const int q2 = __float2int_rn(v.z * id), q3 = __float2int_rn(v.w * id);
            int qsum = q0 + q1 + q2 + q3;
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 1);
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 2);
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 4);
            if (act) {
                const uint32_t b = i >> 3, sub = i & 7, kk = sub >> 1, odd = sub & 1;
                uint8_t *wa = pa + (size_t)b * 16 + kk * 4 + odd;
                uint8_t *wb = pb + (size_t)b * 16 + kk * 4 + odd;
                wa[0] = (uint8_t)(int8_t)q0; wa[2] = (uint8_t)(int8_t)q1;
                wb[0] = (uint8_t)(int8_t)q2; wb[2] = (uint8_t)(int8_t)q3;
                if (sub == 0) {
                    const float dr = __half2float(__float2half_rn(dq));
                    dxs[b] = dr;
                    dx8[b] = dr * 8.f * (float)qsum;
                }
            }
        }
    }
    consumerBarrier();
    stamp();   // prologue done

    // ---- main loop over this phase's fills ----
    const uint32_t gInc = P.gInc, segInc = P.segInc;
    uint32_t rot = 0;   // (fill index * steps per full stage) mod 16: rotates the step -> warp assignment from fill to fill
    for (uint32_t r0 = 0; r0 < tileRows; r0 += SR) {
        const uint32_t st = ring.st;
        const uint32_t rows = min(SR, tileRows - r0);
        const uint32_t nSteps = ((rows + kRowsPerStep - 1) / kRowsPerStep) * nseg;
        const uint8_t *stage = sm.ring + st * m.stageBytes;
        const uint4 *sq = reinterpret_cast<const uint4 *>(stage);
        const uint16_t *ss = reinterpret_cast<const uint16_t *>(stage + SR * rowQsBytes);
        const uint32_t firstStep = ((uint32_t)warp - rot) & (kConsumerWarps - 1);
        rot = (rot + P.rotInc) & (kConsumerWarps - 1);
        uint32_t g = (firstStep * P.recipNseg) >> 16, seg = firstStep - g * nseg;   // firstStep / nseg, firstStep % nseg
        mbarWait(&sm.fullBar[st], ring.par);
        // two steps are processed together (independent dependency chains -> the LDS/IDP/SHFL latencies overlap)
        auto stepDot = [&](uint32_t g_, uint32_t seg_, float (&acc)[kRowsPerStep]) {
            const uint32_t blk = seg_ * 32 + lane;
            const uint32_t rl = g_ * kRowsPerStep;
#pragma unroll
            for (int r = 0; r < kRowsPerStep; r++) acc[r] = 0.f;
            if (blk < nblk) {
                const uint4 A = planeA[blk], B = planeB[blk];
                const float dxv = dxs[blk], dx8v = dx8[blk];
                const uint4 *qp = sq + rl * nblk + blk;
                const uint16_t *sp = ss + rl * nblk + blk;
#pragma unroll
                for (int r = 0; r < kRowsPerStep; r++) {
                    const uint4 q = qp[r * nblk];
                    const float dw = __half2float(__ushort_as_half(sp[r * nblk]));
                    const uint32_t ml = 0x0f0f0f0fu, mh = 0xf0f0f0f0u;
                    int lo = dp4a(q.x & ml, A.x, 0), hi = dp4a(q.x & mh, B.x, 0);
                    lo = dp4a(q.y & ml, A.y, lo); hi = dp4a(q.y & mh, B.y, hi);
                    lo = dp4a(q.z & ml, A.z, lo); hi = dp4a(q.z & mh, B.z, hi);
                    lo = dp4a(q.w & ml, A.w, lo); hi = dp4a(q.w & mh, B.w, hi);
                    acc[r] = dw * (dxv * (float)(lo + (hi >> 4)) - dx8v);
                }
            }
        };
        auto stepStore = [&](uint32_t g_, uint32_t seg_, float v) {
            const uint32_t rl = g_ * kRowsPerStep;
            if ((lane & 7) == 0) {
                const uint32_t r = lane >> 3;
                if (rl + r < rows) partial[(r0 + rl + r) * nseg + seg_] = v;
            }
        };
        auto advance = [&](uint32_t &g_, uint32_t &seg_) {
            g_ += gInc;
            seg_ += segInc;
            if (seg_ >= nseg) { seg_ -= nseg; g_++; }
        };
        uint32_t s = firstStep;
        for (; s + kConsumerWarps < nSteps; s += 2 * kConsumerWarps) {
            uint32_t g2 = g, seg2 = seg;
            advance(g2, seg2);
            float a0[kRowsPerStep], a1[kRowsPerStep];
            stepDot(g, seg, a0);
            stepDot(g2, seg2, a1);
            const float v0 = reduce4(a0[0], a0[1], a0[2], a0[3], lane);
            const float v1 = reduce4(a1[0], a1[1], a1[2], a1[3], lane);
            stepStore(g, seg, v0);
            stepStore(g2, seg2, v1);
            g = g2; seg = seg2;
            advance(g, seg);
        }
        if (s < nSteps) {
            float a0[kRowsPerStep];
            stepDot(g, seg, a0);
            stepStore(g, seg, reduce4(a0[0], a0[1], a0[2], a0[3], lane));
        }
        __syncwarp();
        if (lane == 0) mbarArrive(&sm.emptyBar[st]);
        ring.advance(m.nStages);
    }
    consumerBarrier();
    stamp();   // main loop done

    // ---- epilogue ----
    auto rowSum = [&](uint32_t r) {
        float v = 0.f;
        for (uint32_t sg = 0; sg < nseg; sg++) v += partial[r * nseg + sg];
        return v;
    };
    if (EPI == EPI_SWIGLU_) {
        for (uint32_t p = tid; p < tileRows / 2; p += kConsumerThreads) out[pairBegin + p] = gateAct(rowSum(2 * p), m.act) * rowSum(2 * p + 1);
    } else if (EPI == EPI_RESIDUAL_) {
        if (m.ar.nRanks > 1) {
            const ArArgs &ar = m.ar;
            const size_t slotBase = (size_t)(arParity * ar.nRanks + ar.rank) * ar.slotStride;
            for (uint32_t r = tid; r < tileRows; r += kConsumerThreads) {
                const float v = rowSum(r);
#pragma unroll 1
                for (uint32_t p = 0; p < ar.nRanks; p++) stLL(ar.slots[(ar.rank + p) % ar.nRanks] + slotBase + rowBase + r, __float_as_uint(v), 1u);
            }
            uint64_t *mine = ar.slots[ar.rank];
            for (uint32_t r = tid; r < tileRows; r += kConsumerThreads) {
                float sum = 0.f;
                for (uint32_t sr = 0; sr < ar.nRanks; sr++) {
                    uint64_t *w = mine + (size_t)(arParity * ar.nRanks + sr) * ar.slotStride + rowBase + r;
                    uint2 v = ldLL(w);
                    while (v.y == 0u) v = ldLL(w);
                    sum += __uint_as_float(v.x);
                    stLL(w, 0u, 0u);
                }
                out[rowBase + r] = __ldcg(out + rowBase + r) + sum;
            }
        } else {
            for (uint32_t r = tid; r < tileRows; r += kConsumerThreads) out[rowBase + r] = __ldcg(out + rowBase + r) + rowSum(r);
        }
    } else {
        float best = -INFINITY;
        int bestIdx = 0x7fffffff;
        for (uint32_t r = tid; r < tileRows; r += kConsumerThreads) {
            const float v = rowSum(r);
            out[rowBase + r] = v;
            if (EPI == EPI_ARGMAX_ && v > best && m.rowOffsetGlobal + rowBase + r < m.vocabLimit) { best = v; bestIdx = (int)(m.rowOffsetGlobal + rowBase + r); }
        }
        if (EPI == EPI_ARGMAX_) {
            auto better = [](float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); };
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
            }
            float *sv = red;
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
                    m.argVal[blockIdx.x] = best;
                    m.argIdx[blockIdx.x] = bestIdx;
                    __threadfence();
                    const unsigned int prev = atomicAdd(m.argCounter, 1u);
                    lastCta = prev == gridDim.x - 1;
                    if (lastCta) *m.argCounter = 0;
                }
            }
            consumerBarrier();
            if (lastCta && warp == 0) {
                __threadfence();
                best = -INFINITY;
                bestIdx = 0x7fffffff;
                for (uint32_t i = lane; i < gridDim.x; i += 32) {
                    const float v = __ldcg(m.argVal + i);
                    const int ix = __ldcg(m.argIdx + i);
                    if (better(v, ix, best, bestIdx)) { best = v; bestIdx = ix; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                    if (better(ov, oi, best, bestIdx)) { best = ov; bestIdx = oi; }
                }
                if (m.ar.nRanks > 1) {
                    const ArArgs &ar = m.ar;
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
                if (lane == 0 && m.greedyAdvance) {
                    m.tokens[0] = bestIdx;
                    const int p = m.pos[0] + 1;
                    m.pos[0] = p;
                    if (m.history && (uint32_t)p < m.seqLen) m.history[p] = bestIdx;
                }
            }
        }
    }
}

// Attention phase: work items (head, split) are dealt round-robin to the CTAs; the 16 consumer warps of a CTA share the
// positions of one item. Same math as attnFusedKernel (decode_ops.cu).
template <int HD>
__device__ void megaAttention(const MegaArgs &m, const MegaSmem &sm, const MegaLayer &L, int p, int tid) {
    constexpr int DPL = HD / 32;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t nPos = (uint32_t)p + 1;
    uint32_t eff = (nPos + 255) / 256;
    if (eff > m.nSplits) eff = m.nSplits;
    if (eff < 1) eff = 1;
    const uint32_t kvMul = m.nHeads / m.nKvHeads;
    const uint32_t qDim = m.nHeads * HD, kvDim = m.nKvHeads * HD;
    float *sAcc = sm.partial;                 // [16][HD]
    float *sM = sm.partial + 16 * HD;         // [16]
    float *sL = sM + 16;                      // [16]
    __shared__ bool sLast;
    const float2 *ropeRow = reinterpret_cast<const float2 *>(m.rope) + (size_t)p * (HD / 2) + lane * (DPL / 2);
    auto normRope = [&](float *v, const float *nw) {
        if (nw) {
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < DPL; i++) ss += v[i] * v[i];
            ss = warpSum(ss);
            const float inv = rsqrtf(ss / (float)HD + m.eps);
#pragma unroll
            for (int i = 0; i < DPL; i++) v[i] = nw[lane * DPL + i] * (v[i] * inv);
        }
#pragma unroll
        for (int k = 0; k < DPL / 2; k++) {
            const float2 cs = ropeRow[k];
            const float x0 = v[2 * k] * cs.x - v[2 * k + 1] * cs.y;
            const float x1 = v[2 * k] * cs.y + v[2 * k + 1] * cs.x;
            v[2 * k] = x0; v[2 * k + 1] = x1;
        }
    };
    for (uint32_t item = blockIdx.x; item < m.nHeads * eff; item += gridDim.x) {
        const uint32_t h = item / eff, split = item - h * eff;
        const uint32_t kvh = h / kvMul;
        const uint32_t chunk = (nPos + eff - 1) / eff;
        const uint32_t begin = split * chunk;
        const uint32_t end = min(begin + chunk, nPos);
        const bool ownsNew = end == nPos;
        const uint32_t cachedEnd = ownsNew ? end - 1 : end;
        float q[DPL];
#pragma unroll
        for (int i = 0; i < DPL; i++) q[i] = __ldcg(m.qkv + (size_t)h * HD + lane * DPL + i);
        normRope(q, L.qNorm);
        const float scale = rsqrtf((float)HD);
#pragma unroll
        for (int i = 0; i < DPL; i++) q[i] *= scale;
        float mx = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
        for (int i = 0; i < DPL; i++) acc[i] = 0.f;
        __nv_bfloat16 *kHead = L.kCache + (size_t)kvh * m.seqLen * HD;
        __nv_bfloat16 *vHead = L.vCache + (size_t)kvh * m.seqLen * HD;
        if (ownsNew && warp == 0) {
            float kn[DPL], vn[DPL];
#pragma unroll
            for (int i = 0; i < DPL; i++) {
                kn[i] = __ldcg(m.qkv + qDim + (size_t)kvh * HD + lane * DPL + i);
                vn[i] = __ldcg(m.qkv + qDim + kvDim + (size_t)kvh * HD + lane * DPL + i);
            }
            normRope(kn, L.kNorm);
            __nv_bfloat162 kb[DPL / 2], vb[DPL / 2];
#pragma unroll
            for (int k = 0; k < DPL / 2; k++) {
                kb[k] = __floats2bfloat162_rn(kn[2 * k], kn[2 * k + 1]);
                vb[k] = __floats2bfloat162_rn(vn[2 * k], vn[2 * k + 1]);
                const float2 kf = __bfloat1622float2(kb[k]), vf = __bfloat1622float2(vb[k]);
                kn[2 * k] = kf.x; kn[2 * k + 1] = kf.y;
                vn[2 * k] = vf.x; vn[2 * k + 1] = vf.y;
            }
            if (h % kvMul == 0) {
                __nv_bfloat162 *kd = reinterpret_cast<__nv_bfloat162 *>(kHead + (size_t)p * HD + lane * DPL);
                __nv_bfloat162 *vd = reinterpret_cast<__nv_bfloat162 *>(vHead + (size_t)p * HD + lane * DPL);
#pragma unroll
                for (int k = 0; k < DPL / 2; k++) { kd[k] = kb[k]; vd[k] = vb[k]; }
            }
            float dsum = 0.f;
#pragma unroll
            for (int i = 0; i < DPL; i++) dsum += q[i] * kn[i];
            dsum = warpSum(dsum);
            mx = dsum; l = 1.f;
#pragma unroll
            for (int i = 0; i < DPL; i++) acc[i] = vn[i];
        }
        const __nv_bfloat16 *kBase = kHead + lane * DPL;
        const __nv_bfloat16 *vBase = vHead + lane * DPL;
        constexpr int UN = 4;
        for (uint32_t s0 = begin + warp * UN; s0 < cachedEnd; s0 += kConsumerWarps * UN) {
            float kf[UN][DPL], vf[UN][DPL];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const uint32_t s = s0 + u;
                if (s < cachedEnd) {
                    if constexpr (DPL == 4) {
                        const uint2 kr = *reinterpret_cast<const uint2 *>(kBase + (size_t)s * HD);
                        const uint2 vr = *reinterpret_cast<const uint2 *>(vBase + (size_t)s * HD);
                        const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&kr.x));
                        const float2 k1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&kr.y));
                        const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&vr.x));
                        const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&vr.y));
                        kf[u][0] = k0.x; kf[u][1] = k0.y; kf[u][2] = k1.x; kf[u][3] = k1.y;
                        vf[u][0] = v0.x; vf[u][1] = v0.y; vf[u][2] = v1.x; vf[u][3] = v1.y;
                    } else {
                        const uint32_t kr = *reinterpret_cast<const uint32_t *>(kBase + (size_t)s * HD);
                        const uint32_t vr = *reinterpret_cast<const uint32_t *>(vBase + (size_t)s * HD);
                        const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&kr));
                        const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&vr));
                        kf[u][0] = k0.x; kf[u][1] = k0.y;
                        vf[u][0] = v0.x; vf[u][1] = v0.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < DPL; i++) { kf[u][i] = 0.f; vf[u][i] = 0.f; }
                }
            }
            float sc[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                float dsum = 0.f;
#pragma unroll
                for (int i = 0; i < DPL; i++) dsum += q[i] * kf[u][i];
                sc[u] = warpSum(dsum);
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                if (s0 + u < cachedEnd) {
                    const float mNew = fmaxf(mx, sc[u]);
                    const float corr = __expf(mx - mNew);
                    const float pr = __expf(sc[u] - mNew);
                    l = l * corr + pr;
#pragma unroll
                    for (int i = 0; i < DPL; i++) acc[i] = acc[i] * corr + pr * vf[u][i];
                    mx = mNew;
                }
            }
        }
        consumerBarrier();   // sAcc may still be read by the previous item
#pragma unroll
        for (int i = 0; i < DPL; i++) sAcc[warp * HD + lane * DPL + i] = acc[i];
        if (lane == 0) { sM[warp] = mx; sL[warp] = l; }
        consumerBarrier();
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) M = fmaxf(M, sM[w]);
        float num = 0.f, Lsum = 0.f;
        if (tid < HD) {
#pragma unroll
            for (int w = 0; w < kConsumerWarps; w++) {
                const float wgt = (sM[w] == -INFINITY) ? 0.f : __expf(sM[w] - M);
                num += sAcc[w * HD + tid] * wgt;
                Lsum += sL[w] * wgt;
            }
        }
        float *outRow = m.z + (size_t)h * HD;
        if (eff == 1) {
            if (tid < HD) outRow[tid] = num / Lsum;
            continue;
        }
        float *pOut = m.attnPartial + ((size_t)h * m.nSplits + split) * (HD + 2);
        if (tid < HD) pOut[tid] = num;
        if (tid == 0) { pOut[HD] = M; pOut[HD + 1] = Lsum; }
        __threadfence();
        consumerBarrier();
        if (tid == 0) {
            const unsigned int prev = atomicAdd(&m.attnCounters[h], 1u);
            sLast = (prev == eff - 1);
            if (sLast) m.attnCounters[h] = 0;
        }
        consumerBarrier();
        if (sLast) {
            __threadfence();
            const float *pIn = m.attnPartial + (size_t)h * m.nSplits * (HD + 2);
            float gM = -INFINITY;
            for (uint32_t s = 0; s < eff; s++) gM = fmaxf(gM, __ldcg(pIn + (size_t)s * (HD + 2) + HD));
            if (tid < HD) {
                float n2 = 0.f, den = 0.f;
                for (uint32_t s = 0; s < eff; s++) {
                    const float w = __expf(__ldcg(pIn + (size_t)s * (HD + 2) + HD) - gM);
                    n2 += w * __ldcg(pIn + (size_t)s * (HD + 2) + tid);
                    den += w * __ldcg(pIn + (size_t)s * (HD + 2) + HD + 1);
                }
                outRow[tid] = n2 / den;
            }
        }
    }
}

template <int HD>
__global__ void __launch_bounds__(kTmaThreads, 1) megaDecodeKernel(const __grid_constant__ MegaArgs m) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    MegaSmem sm;
    sm.ring = smem;
    sm.planeA = reinterpret_cast<uint4 *>(smem + (size_t)m.nStages * m.stageBytes);
    sm.planeB = sm.planeA + m.planeBlocks;
    sm.dxs = reinterpret_cast<float *>(sm.planeB + m.planeBlocks);
    sm.dx8 = sm.dxs + m.planeBlocks;
    sm.partial = sm.dx8 + m.planeBlocks;
    sm.red = sm.partial + m.partialFloats;
    sm.fullBar = reinterpret_cast<uint64_t *>(sm.red + 32);
    sm.emptyBar = sm.fullBar + kMaxStages;

    if (tid == 0) {
        for (uint32_t s = 0; s < m.nStages; s++) {
            mbarInit(&sm.fullBar[s], 1);
            mbarInit(&sm.emptyBar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t qDim = m.nHeads * m.headDim, kvDim = m.nKvHeads * m.headDim, qkvDim = qDim + 2 * kvDim;

    if (warp == kConsumerWarps) {
        // =============================== producer: every weight matrix of the token, back to back ===============================
        if (lane == 0) {
            const uint64_t policy = policyEvictFirst();
            RingPos pr{0u, 0u};
            bool wrapped = false;   // every stage has been filled once: from now on wait for the consumers to free it
            auto stream = [&](const uint8_t *qs, const uint8_t *sc, const MegaPhase &P) {
                const uint32_t rowQsBytes = P.nblk * 16, rowScBytes = P.nblk * 2;
                uint32_t pairBegin, tileRows;
                megaTile(P, pairBegin, tileRows);
                const uint32_t rowBase = pairBegin * 2, SR = P.stageRows;
                for (uint32_t r0 = 0; r0 < tileRows; r0 += SR) {
                    const uint32_t st = pr.st;
                    if (wrapped) mbarWait(&sm.emptyBar[st], pr.par ^ 1u);
                    const uint32_t rows = min(SR, tileRows - r0);
                    const uint32_t bq = rows * rowQsBytes, bs = rows * rowScBytes;
                    uint8_t *dst = sm.ring + st * m.stageBytes;
                    mbarExpectTx(&sm.fullBar[st], bq + bs);
                    tmaBulkLoad(dst, qs + (uint64_t)(rowBase + r0) * rowQsBytes, bq, &sm.fullBar[st], policy);
                    tmaBulkLoad(dst + SR * rowQsBytes, sc + (uint64_t)(rowBase + r0) * rowScBytes, bs, &sm.fullBar[st], policy);
                    pr.advance(m.nStages);
                    if (pr.st == 0) wrapped = true;
                }
            };
            for (uint32_t l = 0; l < m.nLayers; l++) {
                const MegaLayer &L = m.layers[l];
                stream(L.qkvQs, L.qkvSc, m.ph[MP_QKV]);
                stream(L.woQs, L.woSc, m.ph[MP_WO]);
                stream(L.w13Qs, L.w13Sc, m.ph[MP_W13]);
                stream(L.w2Qs, L.w2Sc, m.ph[MP_W2]);
            }
            stream(m.wclsQs, m.wclsSc, m.ph[MP_LOGITS]);
        }
        return;
    }

    // =============================== consumers ===============================
    unsigned int barTarget = 0;
    RingPos ring{0u, 0u};
    uint32_t slot = 0;
    auto stamp = [&]() { if (m.trace && tid == 0 && blockIdx.x < m.traceCtas) m.trace[(size_t)blockIdx.x * m.traceStride + slot] = globalTimerNs(); slot++; };
    stamp();
    // embedding: CTA c copies its slice of the row
    {
        int tok = m.tokens[0];
        if (tok < 0 || (uint32_t)tok >= m.vocabFull) tok = 0;
        const uint32_t per = (m.dim + gridDim.x - 1) / gridDim.x;
        const uint32_t b = blockIdx.x * per, e = min(b + per, m.dim);
        for (uint32_t i = b + tid; i < e; i += kConsumerThreads) m.x[i] = m.embedding[(size_t)tok * m.dim + i];
    }
    int p = m.pos[0];
    if (p < 0) p = 0;
    if ((uint32_t)p >= m.seqLen) p = m.seqLen - 1;
    gridBarrier(m.gridCounter, barTarget, tid);
    auto prefetchVec = [&](const float *p) {   // norm weights are constants: pull them towards L2 ahead of their phase
        for (uint32_t i = (blockIdx.x * kConsumerThreads + tid) * 32; i < m.dim; i += gridDim.x * kConsumerThreads * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i));
    };
    for (uint32_t l = 0; l < m.nLayers; l++) {
        const MegaLayer &L = m.layers[l];
        prefetchVec(L.norm1);
        prefetchVec(l + 1 < m.nLayers ? m.layers[l + 1].norm0 : m.finalNorm);
        stamp();
        megaGemv<PRO_RMSNORM_, EPI_STORE_>(m, sm, m.ph[MP_QKV], m.x, L.norm0, m.qkv, 0, ring, tid, slot);
        stamp();
        gridBarrier(m.gridCounter, barTarget, tid);
        stamp();
        megaAttention<HD>(m, sm, L, p, tid);
        stamp();
        gridBarrier(m.gridCounter, barTarget, tid);
        stamp();
        megaGemv<PRO_PLAIN_, EPI_RESIDUAL_>(m, sm, m.ph[MP_WO], m.z, nullptr, m.x, 0, ring, tid, slot);
        stamp();
        gridBarrier(m.gridCounter, barTarget, tid);
        stamp();
        megaGemv<PRO_RMSNORM_, EPI_SWIGLU_>(m, sm, m.ph[MP_W13], m.x, L.norm1, m.h, 0, ring, tid, slot);
        stamp();
        gridBarrier(m.gridCounter, barTarget, tid);
        stamp();
        megaGemv<PRO_PLAIN_, EPI_RESIDUAL_>(m, sm, m.ph[MP_W2], m.h, nullptr, m.x, 1, ring, tid, slot);
        stamp();
        gridBarrier(m.gridCounter, barTarget, tid);
    }
    stamp();
    megaGemv<PRO_RMSNORM_, EPI_ARGMAX_>(m, sm, m.ph[MP_LOGITS], m.x, m.finalNorm, m.logits, 0, ring, tid, slot);
    stamp();
}

// Host: geometry + launch. Returns 1 if the model shape cannot use the persistent kernel.
int launchMegaDecode(MegaArgs m, int numSms, cudaStream_t stream) {
    if (m.headDim != 64 && m.headDim != 128) return 1;
    const uint32_t qDim = m.nHeads * m.headDim;
    const uint32_t ns[4] = {m.dim, qDim, m.ffDim, m.dim};
    uint32_t maxN = 0;
    for (uint32_t n : ns) {
        if (n % 128) return 1;
        if (n > maxN) maxN = n;
    }
    if (maxN > 8 * kConsumerThreads * 4 || m.dim > 4 * kConsumerThreads * 4) return 1;   // activation (+ norm) vectors live in registers during the prologue
    const uint32_t grid = (uint32_t)numSms;
    m.act = gHiddenAct;
    if (m.vocabLimit == 0) m.vocabLimit = 0xffffffffu;
    // partial buffer: rows of the largest tile x segments; also hosts the attention scratch (16 x HD + 32 floats)
    const uint32_t ds[5] = {qDim + 2 * m.nKvHeads * m.headDim, m.dim, 2 * m.ffDim, m.dim, m.vocab};
    const uint32_t dn[5] = {m.dim, qDim, m.dim, m.ffDim, m.dim};
    uint32_t partial = 16 * m.headDim + 64;
    for (int i = 0; i < 5; i++) {
        if (ds[i] % 2) return 1;
        const uint32_t tile = 2 * ((ds[i] / 2 + grid - 1) / grid) + 2;
        const uint32_t nseg = (dn[i] / 32 + 31) / 32;
        if (tile * nseg > partial) partial = tile * nseg;
    }
    m.partialFloats = (partial + 3) / 4 * 4;
    m.planeBlocks = maxN / 32;
    const size_t fixedBytes = (size_t)m.planeBlocks * (16 + 16 + 4 + 4) + (size_t)m.partialFloats * 4 + 32 * 4 + 2 * kMaxStages * 8 + 256;
    const size_t budget = 226 * 1024;
    // stage must hold >= 4 rows of the widest matrix
    const uint32_t maxRowBytes = (maxN / 32) * 18;
    uint32_t stageBytes = 36 * 1024;
    if (maxRowBytes * 4 > stageBytes) stageBytes = (maxRowBytes * 4 + 127) / 128 * 128;
    if (fixedBytes + 2 * (size_t)stageBytes > budget) return 1;
    uint32_t stages = (uint32_t)((budget - fixedBytes) / stageBytes);
    if (stages > (uint32_t)kMaxStages) stages = kMaxStages;
    m.stageBytes = stageBytes;
    m.nStages = stages;
    // per-phase geometry, so the device code contains no integer division
    for (int i = 0; i < 5; i++) {
        MegaPhase &P = m.ph[i];
        P.d = ds[i]; P.n = dn[i];
        P.nblk = dn[i] / 32;
        P.nseg = (P.nblk + 31) / 32;
        P.stageRows = megaStageRows(dn[i], stageBytes);
        P.pairsQ = (ds[i] / 2) / grid;
        P.pairsRem = (ds[i] / 2) % grid;
        P.recipNseg = 65536u / P.nseg + 1u;                    // (s * recip) >> 16 == s / nseg for s < 16
        P.gInc = (uint32_t)kConsumerWarps / P.nseg;
        P.segInc = (uint32_t)kConsumerWarps - P.gInc * P.nseg;
        P.rotInc = ((P.stageRows / kRowsPerStep) * P.nseg) % (uint32_t)kConsumerWarps;
        for (uint32_t sft = 0; sft < (uint32_t)kConsumerWarps; sft++)
            if (((sft * P.recipNseg) >> 16) != sft / P.nseg) return 1;
    }
    const size_t smemBytes = fixedBytes + (size_t)stages * stageBytes;
    static size_t configured[2] = {0, 0};
    const int v = m.headDim == 128 ? 1 : 0;
    if (smemBytes > configured[v]) {
        if (v) DL_CUDA_CHECK(cudaFuncSetAttribute(megaDecodeKernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        else DL_CUDA_CHECK(cudaFuncSetAttribute(megaDecodeKernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        configured[v] = smemBytes;
    }
    DL_CUDA_CHECK(cudaMemsetAsync(m.gridCounter, 0, sizeof(unsigned int), stream));
    if (v) megaDecodeKernel<128><<<grid, kTmaThreads, smemBytes, stream>>>(m);
    else megaDecodeKernel<64><<<grid, kTmaThreads, smemBytes, stream>>>(m);
    DL_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace dl
