// Symmetric peer-memory arena on the CUDA virtual-memory-management API, with an NVLS multicast mapping.
//
// Every rank (one process per GPU) creates one physical allocation, exports it as a POSIX file descriptor and passes the
// descriptor to its peers over abstract unix datagram sockets (SCM_RIGHTS) — no torch / NCCL involved, so the same code
// bootstraps the Python front end and the native multi-GPU launcher. Each rank then holds
//   * uc[r]  : a unicast mapping of rank r's arena (plain ld/st go over NVLink to that GPU), and
//   * mc     : one multicast mapping bound to all arenas: `multimem.st` / `multimem.red` on mc + off are replicated by the
//              NVSwitch into every rank's arena at offset off, `multimem.ld_reduce` returns the switch-side sum.
// Reference component replaced: the TCP mesh bootstrap + per-socket send/recv of NnNetwork (src/nn/nn-network.cpp:295-539).
#include <cuda.h>
#include <errno.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kMaxR = 8;

// ---- driver entry points, resolved at run time (the library carries no link-time dependency on libcuda) ----
struct Drv {
    CUresult (*getGran)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*exportH)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*importH)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*mcCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*mcAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*mcBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*mcGetGran)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*devGetAttr)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*errStr)(CUresult, const char **) = nullptr;
    bool ok = false;
};

template <typename F>
bool resolve(F &fn, const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

Drv &drv() {
    static Drv d;
    static bool tried = false;
    if (!tried) {
        tried = true;
        bool ok = true;
        ok &= resolve(d.getGran, "cuMemGetAllocationGranularity");
        ok &= resolve(d.memCreate, "cuMemCreate");
        ok &= resolve(d.memRelease, "cuMemRelease");
        ok &= resolve(d.exportH, "cuMemExportToShareableHandle");
        ok &= resolve(d.importH, "cuMemImportFromShareableHandle");
        ok &= resolve(d.reserve, "cuMemAddressReserve");
        ok &= resolve(d.addrFree, "cuMemAddressFree");
        ok &= resolve(d.map, "cuMemMap");
        ok &= resolve(d.unmap, "cuMemUnmap");
        ok &= resolve(d.setAccess, "cuMemSetAccess");
        ok &= resolve(d.devGetAttr, "cuDeviceGetAttribute");
        ok &= resolve(d.errStr, "cuGetErrorString");
        // multicast entry points are optional (older drivers): their absence only disables the NVLS mapping
        resolve(d.mcCreate, "cuMulticastCreate");
        resolve(d.mcAddDevice, "cuMulticastAddDevice");
        resolve(d.mcBindMem, "cuMulticastBindMem");
        resolve(d.mcGetGran, "cuMulticastGetGranularity");
        d.ok = ok;
    }
    return d;
}

#define DL_DRV_CHECK(expr)                                                                                  \
    do {                                                                                                    \
        CUresult _r = (expr);                                                                               \
        if (_r != CUDA_SUCCESS) {                                                                           \
            const char *_s = nullptr;                                                                       \
            if (drv().errStr) drv().errStr(_r, &_s);                                                        \
            std::fprintf(stderr, "CUDA driver error %d (%s) at %s:%d\n", (int)_r, _s ? _s : "?", __FILE__, __LINE__); \
            return -(1000 + (int)_r);                                                                       \
        }                                                                                                   \
    } while (0)

enum MsgKind : uint32_t { MSG_ARENA_FD = 1, MSG_MC_FD = 2, MSG_TOKEN = 3 };
struct Msg {
    uint32_t kind, src, phase, pad;
};

struct Vmm {
    uint32_t rank = 0, nRanks = 1;
    int dev = 0;
    size_t bytes = 0;            // mapped size (granularity-rounded)
    std::string tag;
    int sock = -1;
    CUmemGenericAllocationHandle local = 0, peers[kMaxR] = {}, mcHandle = 0;
    CUdeviceptr uc[kMaxR] = {}, mc = 0;
    bool wantMc = false, haveMc = false;
    int localFd = -1;
    // inbox
    int arenaFd[kMaxR];
    int mcFd = -1;
    bool mcSeen = false;         // rank 0's multicast message arrived (its phase field says whether creation succeeded)
    uint32_t mcRootOk = 0;
    uint32_t tokens[16] = {};
    Vmm() { for (int &f : arenaFd) f = -1; }
};

void sockAddr(const Vmm &v, uint32_t r, sockaddr_un &a, socklen_t &len) {
    std::memset(&a, 0, sizeof(a));
    a.sun_family = AF_UNIX;
    const std::string name = "dllama-b200-" + v.tag + "-" + std::to_string(r);
    a.sun_path[0] = '\0';   // abstract namespace: nothing to unlink, vanishes with the process
    std::memcpy(a.sun_path + 1, name.data(), name.size());
    len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

int sendMsg(Vmm &v, uint32_t dst, const Msg &m, int fd) {
    sockaddr_un a;
    socklen_t alen;
    sockAddr(v, dst, a, alen);
    msghdr h{};
    iovec io{(void *)&m, sizeof(m)};
    h.msg_name = &a; h.msg_namelen = alen; h.msg_iov = &io; h.msg_iovlen = 1;
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    if (fd >= 0) {
        std::memset(ctrl, 0, sizeof(ctrl));
        h.msg_control = ctrl; h.msg_controllen = sizeof(ctrl);
        cmsghdr *c = CMSG_FIRSTHDR(&h);
        c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
        std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    }
    // the peer may not have bound its socket yet: retry for up to 120 s
    for (int attempt = 0; attempt < 12000; attempt++) {
        if (sendmsg(v.sock, &h, 0) == (ssize_t)sizeof(m)) return 0;
        if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN && errno != ENOBUFS) break;
        timespec ts{0, 10 * 1000 * 1000};
        nanosleep(&ts, nullptr);
    }
    std::fprintf(stderr, "dl_vmm: sendmsg to rank %u failed: %s\n", dst, std::strerror(errno));
    return -1;
}

// Receives one message (blocking up to timeoutMs) and files it into the inbox.
int pump(Vmm &v, int timeoutMs) {
    pollfd p{v.sock, POLLIN, 0};
    const int pr = poll(&p, 1, timeoutMs);
    if (pr <= 0) return -1;
    Msg m{};
    msghdr h{};
    iovec io{&m, sizeof(m)};
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    h.msg_iov = &io; h.msg_iovlen = 1; h.msg_control = ctrl; h.msg_controllen = sizeof(ctrl);
    if (recvmsg(v.sock, &h, 0) != (ssize_t)sizeof(m)) return -1;
    int fd = -1;
    for (cmsghdr *c = CMSG_FIRSTHDR(&h); c; c = CMSG_NXTHDR(&h, c))
        if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
    if (m.kind == MSG_ARENA_FD && m.src < (uint32_t)kMaxR) v.arenaFd[m.src] = fd;
    else if (m.kind == MSG_MC_FD) { v.mcFd = fd; v.mcSeen = true; v.mcRootOk = m.phase; }
    else if (m.kind == MSG_TOKEN && m.phase < 16) v.tokens[m.phase]++;
    return 0;
}

// All-to-all token exchange over the bootstrap sockets: returns when every peer has reached `phase`.
int barrier(Vmm &v, uint32_t phase) {
    Msg m{MSG_TOKEN, v.rank, phase, 0};
    for (uint32_t r = 0; r < v.nRanks; r++)
        if (r != v.rank && sendMsg(v, r, m, -1) != 0) return -1;
    while (v.tokens[phase] < v.nRanks - 1)
        if (pump(v, 120000) != 0) { std::fprintf(stderr, "dl_vmm: rank %u timed out in barrier %u\n", v.rank, phase); return -1; }
    return 0;
}

int mapHandle(Vmm &v, CUmemGenericAllocationHandle h, size_t gran, CUdeviceptr *out) {
    Drv &d = drv();
    DL_DRV_CHECK(d.reserve(out, v.bytes, gran, 0, 0));
    DL_DRV_CHECK(d.map(*out, v.bytes, 0, h, 0));
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = v.dev; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DL_DRV_CHECK(d.setAccess(*out, v.bytes, &acc, 1));
    return 0;
}

}  // namespace

// Reports whether the current device supports the VMM path (bit 0) and multicast objects (bit 1).
DL_EXPORT int dl_vmm_supported(int *flags) {
    *flags = 0;
    int dev = 0;
    DL_CUDA_CHECK(cudaGetDevice(&dev));
    DL_CUDA_CHECK(cudaFree(nullptr));
    Drv &d = drv();
    if (!d.ok) return 0;
    int vmm = 0, fdOk = 0, mcast = 0;
    d.devGetAttr(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
    d.devGetAttr(&fdOk, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
    d.devGetAttr(&mcast, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    if (vmm && fdOk) *flags |= 1;
    if (vmm && fdOk && mcast && d.mcCreate && d.mcAddDevice && d.mcBindMem && d.mcGetGran) *flags |= 2;
    return 0;
}

// Step 1 (no communication): allocate this rank's arena and open the bootstrap socket. `tag` must be identical on all ranks of
// the job and unique per job (e.g. master port + a counter).
DL_EXPORT void *dl_vmm_create(uint32_t rank, uint32_t nRanks, size_t bytes, const char *tag, int wantMulticast) {
    if (nRanks < 1 || nRanks > (uint32_t)kMaxR || rank >= nRanks) return nullptr;
    int flags = 0;
    if (dl_vmm_supported(&flags) != 0 || !(flags & 1)) return nullptr;
    Drv &d = drv();
    Vmm *v = new Vmm();
    v->rank = rank; v->nRanks = nRanks; v->tag = tag ? tag : "job";
    v->wantMc = wantMulticast && (flags & 2) && nRanks > 1;
    cudaGetDevice(&v->dev);
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = v->dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (d.getGran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) { delete v; return nullptr; }
    if (v->wantMc) {
        CUmulticastObjectProp mp{};
        mp.numDevices = nRanks; mp.size = bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (d.mcGetGran(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
    }
    v->bytes = (bytes + gran - 1) / gran * gran;
    if (d.memCreate(&v->local, v->bytes, &prop, 0) != CUDA_SUCCESS) { delete v; return nullptr; }
    if (d.exportH(&v->localFd, v->local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) { d.memRelease(v->local); delete v; return nullptr; }
    v->sock = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
    sockaddr_un a;
    socklen_t alen;
    sockAddr(*v, rank, a, alen);
    if (v->sock < 0 || bind(v->sock, (sockaddr *)&a, alen) != 0) {
        std::fprintf(stderr, "dl_vmm: bind failed: %s\n", std::strerror(errno));
        d.memRelease(v->local); delete v; return nullptr;
    }
    return v;
}

// Step 2 (collective): exchange descriptors, map every peer, set up the multicast object. Returns 0 on success.
DL_EXPORT int dl_vmm_connect(void *h) {
    Vmm &v = *(Vmm *)h;
    Drv &d = drv();
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = v.dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    DL_DRV_CHECK(d.getGran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    // own mapping
    if (mapHandle(v, v.local, gran, &v.uc[v.rank]) != 0) return -2;
    DL_CUDA_CHECK(cudaMemset((void *)v.uc[v.rank], 0, v.bytes));
    DL_CUDA_CHECK(cudaDeviceSynchronize());
    // arena descriptors to every peer
    Msg m{MSG_ARENA_FD, v.rank, 0, 0};
    for (uint32_t r = 0; r < v.nRanks; r++)
        if (r != v.rank && sendMsg(v, r, m, v.localFd) != 0) return -3;
    for (uint32_t r = 0; r < v.nRanks; r++) {
        if (r == v.rank) continue;
        while (v.arenaFd[r] < 0)
            if (pump(v, 120000) != 0) { std::fprintf(stderr, "dl_vmm: rank %u: no arena descriptor from rank %u\n", v.rank, r); return -4; }
        DL_DRV_CHECK(d.importH(&v.peers[r], (void *)(intptr_t)v.arenaFd[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
        close(v.arenaFd[r]);
        if (mapHandle(v, v.peers[r], gran, &v.uc[r]) != 0) return -5;
    }
    if (barrier(v, 0) != 0) return -6;   // every arena is zeroed and mapped everywhere
    if (v.wantMc) {
        int mcOk = 1;
        if (v.rank == 0) {
            CUmulticastObjectProp mp{};
            mp.numDevices = v.nRanks; mp.size = v.bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            int fd = -1;
            if (d.mcCreate(&v.mcHandle, &mp) != CUDA_SUCCESS || d.exportH(&fd, v.mcHandle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
                std::fprintf(stderr, "dl_vmm: multicast object creation failed; continuing with unicast peer stores\n");
                mcOk = 0;
            }
            Msg mm{MSG_MC_FD, 0, (uint32_t)mcOk, 0};
            for (uint32_t r = 1; r < v.nRanks; r++)
                if (sendMsg(v, r, mm, mcOk ? fd : -1) != 0) return -7;
            if (fd >= 0) close(fd);
        } else {
            // rank 0 always sends MSG_MC_FD; its phase field is 0 when the object could not be created there
            while (!v.mcSeen)
                if (pump(v, 120000) != 0) return -8;
            if (!v.mcRootOk || v.mcFd < 0) mcOk = 0;
            if (mcOk && d.importH(&v.mcHandle, (void *)(intptr_t)v.mcFd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) mcOk = 0;
            if (v.mcFd >= 0) close(v.mcFd);
        }
        if (mcOk && d.mcAddDevice(v.mcHandle, v.dev) != CUDA_SUCCESS) mcOk = 0;
        if (barrier(v, 1) != 0) return -9;           // all devices added before any memory is bound
        if (mcOk && d.mcBindMem(v.mcHandle, 0, v.local, 0, v.bytes, 0) != CUDA_SUCCESS) mcOk = 0;
        if (mcOk && mapHandle(v, v.mcHandle, gran, &v.mc) != 0) mcOk = 0;
        v.haveMc = mcOk != 0;
        if (barrier(v, 2) != 0) return -10;          // all bindings in place before the first multimem instruction
    }
    return 0;
}

DL_EXPORT void *dl_vmm_ptr(void *h, uint32_t rank) { return (void *)((Vmm *)h)->uc[rank]; }
DL_EXPORT void *dl_vmm_mc_ptr(void *h) { return ((Vmm *)h)->haveMc ? (void *)((Vmm *)h)->mc : nullptr; }
DL_EXPORT size_t dl_vmm_bytes(void *h) { return ((Vmm *)h)->bytes; }

// Host-side barrier over the bootstrap sockets (phases 3..14 are free for callers; used by the native launcher).
DL_EXPORT int dl_vmm_barrier(void *h, uint32_t phase) {
    if (phase < 3 || phase > 14) return -1;
    return barrier(*(Vmm *)h, phase);
}

DL_EXPORT void dl_vmm_destroy(void *h) {
    Vmm *v = (Vmm *)h;
    if (!v) return;
    Drv &d = drv();
    cudaDeviceSynchronize();
    if (v->mc) { d.unmap(v->mc, v->bytes); d.addrFree(v->mc, v->bytes); }
    for (uint32_t r = 0; r < v->nRanks; r++) {
        if (v->uc[r]) { d.unmap(v->uc[r], v->bytes); d.addrFree(v->uc[r], v->bytes); }
        if (r != v->rank && v->peers[r]) d.memRelease(v->peers[r]);
    }
    if (v->mcHandle) d.memRelease(v->mcHandle);
    if (v->local) d.memRelease(v->local);
    if (v->localFd >= 0) close(v->localFd);
    if (v->sock >= 0) close(v->sock);
    delete v;
}

// ---- device-side self test of the mappings (tools/probe_multicast.py, tests/test_gpu_tp.py) ----
namespace {
__global__ void mcSelfTestKernel(float *ucLocal, float *mc, float *const *peersUc, uint32_t rank, uint32_t nRanks, uint32_t n, int phase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (phase == 0) {
        // unicast peer stores: slot [rank][i] of every peer
        for (uint32_t p = 0; p < nRanks; p++) peersUc[p][(size_t)(1 + rank) * n + i] = (float)(rank * 1000 + i);
        // multicast reduce-add into region 0 (replicated in every arena)
        if (mc) asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc + i), "f"((float)(rank + 1)) : "memory");
    } else if (mc) {
        // switch-side reduction of region [1 + nRanks + 1] (each rank wrote its own copy before)
        float v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc + (size_t)(2 + nRanks) * n + i) : "memory");
        ucLocal[(size_t)(3 + nRanks) * n + i] = v;
    }
}
}  // namespace

DL_EXPORT int dl_vmm_selftest_kernel(void *h, uint32_t n, int phase, cudaStream_t stream) {
    Vmm &v = *(Vmm *)h;
    float **tab = nullptr;
    DL_CUDA_CHECK(cudaMalloc(&tab, sizeof(float *) * kMaxR));
    float *host[kMaxR] = {};
    for (uint32_t r = 0; r < v.nRanks; r++) host[r] = (float *)v.uc[r];
    DL_CUDA_CHECK(cudaMemcpyAsync(tab, host, sizeof(host), cudaMemcpyHostToDevice, stream));
    mcSelfTestKernel<<<(n + 255) / 256, 256, 0, stream>>>((float *)v.uc[v.rank], v.haveMc ? (float *)v.mc : nullptr, tab, v.rank, v.nRanks, n, phase);
    DL_CUDA_CHECK(cudaGetLastError());
    DL_CUDA_CHECK(cudaStreamSynchronize(stream));
    DL_CUDA_CHECK(cudaFree(tab));
    return 0;
}
