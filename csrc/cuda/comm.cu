// Peer-memory plumbing for the in-kernel collectives: a per-rank symmetric arena (cudaMalloc) whose IPC handle is
// exchanged through torch.distributed, so every rank holds raw device pointers into every peer's arena (NVLink P2P).
//
// Reference component replaced: NnNetwork (TCP full mesh, src/nn/nn-network.cpp:295-539) and the sync steps built on
// it (syncWithRoot / syncNodeSlices, :541-632). There is no host in the data path here: kernels store into peer arenas
// and spin on flags that peers set (see the all-reduce epilogue in gemv_q40_tma.cu and gemm kernels).
#include <cstring>

#include "common.cuh"

DL_EXPORT int dl_comm_alloc(size_t bytes, void **ptr) {
    DL_CUDA_CHECK(cudaMalloc(ptr, bytes));
    DL_CUDA_CHECK(cudaMemset(*ptr, 0, bytes));
    DL_CUDA_CHECK(cudaDeviceSynchronize());
    return 0;
}

DL_EXPORT int dl_comm_free(void *ptr) {
    DL_CUDA_CHECK(cudaFree(ptr));
    return 0;
}

DL_EXPORT int dl_comm_ipc_handle(void *ptr, void *out64) {
    cudaIpcMemHandle_t h;
    DL_CUDA_CHECK(cudaIpcGetMemHandle(&h, ptr));
    static_assert(sizeof(h) == 64, "unexpected IPC handle size");
    std::memcpy(out64, &h, 64);
    return 0;
}

DL_EXPORT int dl_comm_ipc_open(const void *handle64, void **ptr) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    DL_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

DL_EXPORT int dl_comm_ipc_close(void *ptr) {
    DL_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return 0;
}

DL_EXPORT int dl_comm_memset(void *ptr, int value, size_t bytes, cudaStream_t stream) {
    DL_CUDA_CHECK(cudaMemsetAsync(ptr, value, bytes, stream));
    return 0;
}
