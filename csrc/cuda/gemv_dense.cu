// Matrix-vector kernels for `.m` files whose matrices are *not* q40: f32 weights (the reference's F32_F32_F32 matmul
// with `--buffer-float-type f32`, src/nn/nn-cpu-ops.cpp:1138-1160, vulkan/matmul-forward-f32-f32-f32.comp) and f16
// weights (which the reference can store but not run). Activations stay f32 — no q80 round trip on this path.
//
// Same fusion contract as the q40 kernels (gemv_q40.cu): rmsnorm prologue, store / residual / SwiGLU epilogues,
// 1..8 tokens per launch, one persistent CTA per SM owning a contiguous pair-aligned row tile. The weight stream is
// 4-8x larger than q40, so the kernel is a pure streaming loop: every lane keeps kUnroll 16-byte loads in flight
// (64 KB per SM), activations are staged once per CTA in shared memory, row sums are reduced in a fixed order.
#include "kernels.h"

namespace dl {

namespace {

constexpr int kDenseThreads = 512;
constexpr int kDenseWarps = kDenseThreads / 32;
constexpr int kDenseUnroll = 8;

template <typename WT> struct Chunk;   // one 16-byte load of a weight row
template <> struct Chunk<float> {
    static constexpr int kElems = 4;
    __device__ static __forceinline__ void unpack(const uint4 &q, float (&w)[4]) {
        w[0] = __uint_as_float(q.x); w[1] = __uint_as_float(q.y); w[2] = __uint_as_float(q.z); w[3] = __uint_as_float(q.w);
    }
};
template <> struct Chunk<__half> {
    static constexpr int kElems = 8;
    __device__ static __forceinline__ void unpack(const uint4 &q, float (&w)[8]) {
        const uint32_t v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v[i]));
            w[2 * i] = f.x; w[2 * i + 1] = f.y;
        }
    }
};

__device__ __forceinline__ float blockSumDense(float v, float *red) {
    v = warpSum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < kDenseWarps) ? red[lane] : 0.f;
    t = warpSum(t);
    __syncthreads();
    return t;
}

template <typename WT, int PRO, int EPI, int NB>
__global__ void __launch_bounds__(kDenseThreads, 1) gemvDenseKernel(GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int E = Chunk<WT>::kElems;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint32_t nPairs = a.d / 2;
    const uint32_t pairBegin = (uint32_t)(((uint64_t)blockIdx.x * nPairs) / gridDim.x);
    const uint32_t pairEnd = (uint32_t)(((uint64_t)(blockIdx.x + 1) * nPairs) / gridDim.x);
    const uint32_t rowBase = pairBegin * 2;
    const uint32_t tileRows = (pairEnd - pairBegin) * 2;

    float *xs = reinterpret_cast<float *>(smem);                 // [NB][n] (normalised) activations
    float *rowOut = xs + (size_t)NB * a.n;                       // [maxTileRows][NB]
    float *red = rowOut + (size_t)a.maxTileRows * NB;            // [16]

    pdlLaunchDependents();
    pdlWait();

    // ---- prologue: (rmsnorm) activations -> shared memory ----
    {
        const uint32_t nVec = a.n / 4;
#pragma unroll 1
        for (int t = 0; t < NB; t++) {
            const float4 *x4 = reinterpret_cast<const float4 *>(a.in + (size_t)t * a.inStride);
            float inv = 1.f;
            if (PRO == PRO_RMSNORM_) {
                float ss = 0.f;
                for (uint32_t i = tid; i < nVec; i += kDenseThreads) {
                    const float4 v = x4[i];
                    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                }
                ss = blockSumDense(ss, red);
                inv = rsqrtf(ss / (float)a.n + a.eps);
            }
            float4 *dst = reinterpret_cast<float4 *>(xs + (size_t)t * a.n);
            for (uint32_t i = tid; i < nVec; i += kDenseThreads) {
                float4 v = x4[i];
                if (PRO == PRO_RMSNORM_) {
                    const float4 w = reinterpret_cast<const float4 *>(a.normW)[i];
                    v.x = w.x * (v.x * inv); v.y = w.y * (v.y * inv); v.z = w.z * (v.z * inv); v.w = w.w * (v.w * inv);
                }
                dst[i] = v;
            }
        }
    }
    __syncthreads();

    // ---- main loop: one warp per row, kDenseUnroll x 16 B in flight per lane ----
    const uint32_t nChunks = a.n / E;
    const WT *wBase = reinterpret_cast<const WT *>(a.qs);
    for (uint32_t r = warp; r < tileRows; r += kDenseWarps) {
        const uint4 *row = reinterpret_cast<const uint4 *>(wBase + (size_t)(rowBase + r) * a.n);
        float acc[NB];
#pragma unroll
        for (int t = 0; t < NB; t++) acc[t] = 0.f;
        for (uint32_t c0 = lane; c0 < nChunks; c0 += 32 * kDenseUnroll) {
            uint4 q[kDenseUnroll];
#pragma unroll
            for (int u = 0; u < kDenseUnroll; u++) {
                const uint32_t c = c0 + u * 32;
                q[u] = c < nChunks ? ldgStream16(row + c) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kDenseUnroll; u++) {
                const uint32_t c = c0 + u * 32;
                if (c < nChunks) {
                    float w[E];
                    Chunk<WT>::unpack(q[u], w);
#pragma unroll
                    for (int t = 0; t < NB; t++) {
                        const float4 *xv = reinterpret_cast<const float4 *>(xs + (size_t)t * a.n + (size_t)c * E);
#pragma unroll
                        for (int k = 0; k < E / 4; k++) {
                            const float4 x = xv[k];
                            acc[t] = fmaf(w[4 * k], x.x, acc[t]);
                            acc[t] = fmaf(w[4 * k + 1], x.y, acc[t]);
                            acc[t] = fmaf(w[4 * k + 2], x.z, acc[t]);
                            acc[t] = fmaf(w[4 * k + 3], x.w, acc[t]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NB; t++) {
            const float v = warpSum(acc[t]);
            if (lane == 0) rowOut[(size_t)r * NB + t] = v;
        }
    }
    __syncthreads();

    // ---- epilogue ----
    if (EPI == EPI_SWIGLU_) {
        const uint32_t tilePairs = tileRows / 2;
        for (uint32_t i = tid; i < tilePairs * NB; i += kDenseThreads) {
            const uint32_t p = i / NB, t = i - p * NB;
            const float g = rowOut[(size_t)(2 * p) * NB + t], up = rowOut[(size_t)(2 * p + 1) * NB + t];
            a.out[(size_t)t * a.outStride + pairBegin + p] = gateAct(g, a.act) * up;
        }
    } else {
        for (uint32_t i = tid; i < tileRows * NB; i += kDenseThreads) {
            const uint32_t r = i / NB, t = i - r * NB;
            float *o = a.out + (size_t)t * a.outStride + rowBase + r;
            const float v = rowOut[(size_t)r * NB + t];
            if (EPI == EPI_RESIDUAL_) *o += v;
            else *o = v;
        }
    }
}

template <typename WT, int PRO, int EPI, int NB>
int launchDense(const GemvArgs &a, int grid, size_t smemBytes, cudaStream_t stream, bool pdl) {
    auto kernel = gemvDenseKernel<WT, PRO, EPI, NB>;
    static size_t configured = 0;   // per instantiation
    if (smemBytes > configured) {
        DL_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
        configured = smemBytes;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kDenseThreads);
    cfg.dynamicSmemBytes = smemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, a));
    return 0;
}

template <typename WT>
int dispatchDense(int pro, int epi, int nb, const GemvArgs &a, int grid, size_t smemBytes, cudaStream_t stream, bool pdl) {
#define DL_DENSE_CASE(P, E, N) \
    if (pro == P && epi == E && nb == N) return launchDense<WT, P, E, N>(a, grid, smemBytes, stream, pdl);
#define DL_DENSE_NB(P, E) DL_DENSE_CASE(P, E, 1) DL_DENSE_CASE(P, E, 2) DL_DENSE_CASE(P, E, 4) DL_DENSE_CASE(P, E, 8)
    DL_DENSE_NB(PRO_RMSNORM_, EPI_STORE_)
    DL_DENSE_NB(PRO_PLAIN_, EPI_RESIDUAL_)
    DL_DENSE_NB(PRO_RMSNORM_, EPI_SWIGLU_)
    DL_DENSE_NB(PRO_PLAIN_, EPI_STORE_)
#undef DL_DENSE_NB
#undef DL_DENSE_CASE
    return -3;
}

}  // namespace

size_t gemvDenseSmemBytes(uint32_t n, uint32_t maxTileRows, int nb) {
    return (size_t)nb * n * 4 + (size_t)maxTileRows * nb * 4 + 16 * 4 + 16;
}

// wtype: 1 = f32, 2 = f16. `a.qs` points at the row-major [d][n] matrix, `a.scales` is unused.
int gemvDense(int wtype, int pro, int epi, int nb, GemvArgs a, int numSms, cudaStream_t stream, bool pdl) {
    if (a.d % 2 || a.n % 8) return -1;
    a.act = gHiddenAct;
    if (a.expertIdx || a.moeCtasPerSlot || a.ar.nRanks > 1) return -40;   // MoE routing / in-kernel all-reduce: q40 kernels only
    const uint32_t nPairs = a.d / 2;
    const int grid = (int)(nPairs < (uint32_t)numSms ? nPairs : (uint32_t)numSms);
    a.maxTileRows = 2 * ((nPairs + grid - 1) / grid);
    const size_t smemBytes = gemvDenseSmemBytes(a.n, a.maxTileRows, nb);
    if (smemBytes > 227 * 1024) return -2;
    if (wtype == 1) return dispatchDense<float>(pro, epi, nb, a, grid, smemBytes, stream, pdl);
    if (wtype == 2) return dispatchDense<__half>(pro, epi, nb, a, grid, smemBytes, stream, pdl);
    return -4;
}

}  // namespace dl

// Standalone entry point (tests / microbenchmarks).
DL_EXPORT int dl_gemv_dense(int wtype, int pro, int epi, int nb, const void *w, uint32_t d, uint32_t n, const float *in,
                            uint32_t inStride, const float *normW, float eps, float *out, uint32_t outStride, int numSms,
                            cudaStream_t stream, int pdl) {
    dl::GemvArgs a{};
    a.qs = (const uint32_t *)w;
    a.d = d; a.n = n;
    a.in = in; a.normW = normW; a.eps = eps;
    a.out = out; a.inStride = inStride; a.outStride = outStride;
    return dl::gemvDense(wtype, pro, epi, nb, a, numSms, stream, pdl != 0);
}
