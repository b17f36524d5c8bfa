// `.m` model file: header, tensor directory, tensor-parallel slicing rules, RoPE tables.
//
// Parity targets (behaviour, not code): reference src/llm.cpp:36-116 (header), :614-669 (tensor walk),
// src/nn/nn-core.cpp:211-322 (slicers/splitters), :326-383 (rope cache), converter/writer.py:109-145.
//
// Design difference: the reference interleaves "walk the file" with "send to worker sockets". Here the
// file is described once as a flat *tensor directory* (name, layer, expert, dtype, shape, byte range,
// partitioning rule); every rank then pulls exactly the byte ranges it owns (or a peer-memory scatter
// does), and the GPU repack kernels consume (offset, pitch) pairs. No per-tensor callbacks.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "quants.hpp"

namespace dl {

enum ArchType : int32_t { ARCH_LLAMA = 0xABCD00, ARCH_QWEN3 = 0xABCD01, ARCH_QWEN3_MOE = 0xABCD02 };
enum RopeType : int32_t { ROPE_LLAMA = 0, ROPE_FALCON = 1, ROPE_LLAMA3_1 = 2 };
enum HiddenAct : int32_t { ACT_GELU = 0, ACT_SILU = 1 };

enum HeaderKey : int32_t {
    K_VERSION = 0, K_ARCH_TYPE = 1, K_DIM = 2, K_HIDDEN_DIM = 3, K_N_LAYERS = 4, K_N_HEADS = 5,
    K_N_KV_HEADS = 6, K_N_EXPERTS = 7, K_N_ACTIVE_EXPERTS = 8, K_VOCAB_SIZE = 9, K_SEQ_LEN = 10,
    K_HIDDEN_ACT = 11, K_ROPE_THETA = 12, K_WEIGHT_FLOAT_TYPE = 13, K_ROPE_SCALING_FACTOR = 14,
    K_ROPE_SCALING_LOW_FREQ_FACTOR = 15, K_ROPE_SCALING_HIGH_FREQ_FACTOR = 16,
    K_ROPE_SCALING_ORIG_MAX_SEQ_LEN = 17, K_ROPE_TYPE = 18, K_HEAD_DIM = 19, K_NORM_EPSILON = 20,
    K_MOE_HIDDEN_DIM = 21,
};

constexpr int32_t kModelMagic = 0xA00ABCD;

struct ModelHeader {
    uint64_t headerSize = 0;   // bytes, counted from file start (magic + size + kv pairs)
    uint64_t fileSize = 0;
    int32_t version = 0;
    ArchType arch = ARCH_LLAMA;
    uint32_t dim = 0, hiddenDim = 0, moeHiddenDim = 0;
    uint32_t nLayers = 0, nHeads = 0, nKvHeads = 0, headDim = 0;
    uint32_t nExperts = 0, nActiveExperts = 0;
    uint32_t vocabSize = 0, seqLen = 0, origSeqLen = 0;
    uint32_t qDim = 0, kvDim = 0;
    HiddenAct hiddenAct = ACT_SILU;
    RopeType ropeType = ROPE_LLAMA;
    float ropeTheta = 10000.f;
    float ropeScalingFactor = 1.f, ropeScalingLowFreqFactor = 0.f, ropeScalingHighFreqFactor = 0.f;
    uint32_t ropeScalingOrigMaxSeqLen = 0;
    float normEpsilon = 1e-5f;
    FloatType weightType = F_UNK;

    uint32_t ffDim() const { return arch == ARCH_QWEN3_MOE ? moeHiddenDim : hiddenDim; }
    bool qkNorm() const { return arch == ARCH_QWEN3 || arch == ARCH_QWEN3_MOE; }
};

// Parses the header from the first bytes of a file (data must cover at least 8 bytes + the kv area).
ModelHeader parseModelHeader(const uint8_t *data, size_t nBytes, uint64_t fileSize, uint32_t maxSeqLen);
ModelHeader loadModelHeader(const std::string &path, uint32_t maxSeqLen);
// Serialises (key,value) pairs in the order given; returns the header bytes (magic, size, pairs).
std::vector<uint8_t> buildModelHeader(const std::vector<std::pair<int32_t, int32_t>> &kv);
std::string describeModelHeader(const ModelHeader &h);   // the "💡 ..." lines of the CLI

const char *archName(ArchType a);
const char *ropeName(RopeType r);

// How a tensor is partitioned over tensor-parallel ranks.
enum Partition : int32_t {
    PART_ROOT = 0,       // only rank 0 needs it (reference: loadRoot — token embedding)
    PART_REPLICATE = 1,  // every rank holds a full copy (norm weights, MoE router)
    PART_ROWS = 2,       // split output dim d: rank r owns rows [r*d/N, (r+1)*d/N)
    PART_COLS = 3,       // split input dim n: rank r owns columns [r*n/N, (r+1)*n/N) of every row
};

struct TensorEntry {
    std::string name;     // e.g. "block_matmul_q"
    uint32_t layer = 0;
    uint32_t expert = 0;
    FloatType type = F_32;
    uint64_t d = 1;       // rows (output dim); 1 for vectors
    uint64_t n = 0;       // columns (input dim) — quant blocks run along n
    uint64_t offset = 0;  // absolute byte offset in the file
    uint64_t nBytes = 0;
    Partition part = PART_REPLICATE;
};

// Fixed tensor order of the format. Throws if the sizes do not add up to fileSize.
std::vector<TensorEntry> buildTensorDirectory(const ModelHeader &h, bool checkFileSize = true);

struct SliceRange {     // byte geometry of one rank's share of a tensor
    uint64_t firstRow = 0, nRows = 0;        // rows owned
    uint64_t rowBytes = 0;                   // pitch of a full row in the file
    uint64_t colByteOffset = 0, colBytes = 0;  // bytes owned inside each row
    uint64_t firstCol = 0, nCols = 0;        // element geometry
    uint64_t totalBytes() const { return nRows * colBytes; }
};
SliceRange sliceTensor(const TensorEntry &t, uint32_t rank, uint32_t nRanks);
// Copies the rank's slice into `out` (tightly packed, row-major). Returns bytes written.
uint64_t extractSlice(const TensorEntry &t, const uint8_t *fileBase, uint32_t rank, uint32_t nRanks, uint8_t *out);

// RoPE angle table, f32 [seqLen][headDim/2][2] = (cos, sin) of pos * freq_j, freq_j = theta^(-2j/headDim),
// with the Llama-3.1 frequency rescaling when ropeType == ROPE_LLAMA3_1 and factor != 1.
// The same table serves the interleaved (Llama) and half-split (NeoX/Falcon) conventions: both rotate
// pair j by angle pos*freq_j, they only disagree on which two lanes form pair j.
void buildRopeTable(const ModelHeader &h, uint32_t seqLen, float *out);
float ropeFrequency(const ModelHeader &h, uint32_t pairIndex);

uint64_t requiredDeviceBytes(const ModelHeader &h, uint32_t nRanks, uint32_t kvBytesPerElem);

}  // namespace dl
