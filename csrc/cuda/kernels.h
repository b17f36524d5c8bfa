// Internal C++ interface between the kernel translation units and the engine.
#pragma once
#include "common.cuh"

namespace dl {

enum { PRO_RMSNORM_ = 0, PRO_PLAIN_ = 1 };
enum { EPI_STORE_ = 0, EPI_RESIDUAL_ = 1, EPI_SWIGLU_ = 2 };

struct GemvArgs {
    const uint32_t *qs;
    const __half *scales;
    uint32_t d, n;
    const float *in;
    const float *normW;
    float eps;
    float *out;
    uint32_t inStride, outStride;
    uint32_t maxTileRows;
    const int *expertIdx;
    uint32_t slot, kActive;
    uint64_t expertQsStride, expertScaleStride;
    const float *expertWeight;
};
int gemvQ40(int pro, int epi, int nb, GemvArgs a, int numSms, cudaStream_t stream, bool pdl);

struct RopeKvArgs {
    float *qkv;
    uint32_t qkvStride;
    const int *pos;
    const float *rope;
    const float *qNorm;
    const float *kNorm;
    float eps;
    uint32_t nHeads, nKvHeads, headDim, seqLen;
    __nv_bfloat16 *kCache;
    __nv_bfloat16 *vCache;
};
int launchRopeKv(const RopeKvArgs &a, int nb, cudaStream_t stream, bool pdl);

struct AttnArgs {
    const float *qkv;
    uint32_t qkvStride;
    const int *pos;
    const __nv_bfloat16 *kCache, *vCache;
    uint32_t nHeads, nKvHeads, headDim, seqLen, nSplits;
    float *partial;
    unsigned int *counters;
    float *out;
    uint32_t outStride;
};
int launchAttnDecode(const AttnArgs &a, int nb, cudaStream_t stream, bool pdl);

int launchEmbedding(const float *table, const int *tokens, float *x, uint32_t dim, uint32_t xStride, uint32_t vocab, int nb,
                    cudaStream_t stream);
int launchArgmaxAdvance(const float *logits, uint32_t vocab, int *tokenOut, int *pos, int *history, uint32_t historyCap,
                        cudaStream_t stream, bool pdl);

}  // namespace dl
