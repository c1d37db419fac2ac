// Exchange buffers in peer (NVLink) memory: plain cudaMalloc blocks shared between the one-process-per-GPU ranks of
// a node through CUDA IPC handles.  Used by the fused reduce + cross-GPU sum of the streaming pass
// (pla_stream_pass_peer_f64); the handles travel between the processes over torch.distributed (parallel.PeerComm).
#include <cstring>
#include "common.cuh"
#include "../../include/parla_b200.h"

using namespace pla;

extern "C" int pla_peer_alloc(size_t bytes, void** dev_ptr) {
    PLA_CHECK_ARG(bytes > 0, 1, "zero bytes");
    PLA_CHECK_ARG(dev_ptr != nullptr, 2, "dev_ptr is null");
    void* p = nullptr;
    PLA_CUDA(cudaMalloc(&p, bytes));            // not from torch's caching allocator: IPC needs a whole allocation
    PLA_CUDA(cudaMemset(p, 0, bytes));          // flag 0 is never a valid epoch
    PLA_CUDA(cudaDeviceSynchronize());
    *dev_ptr = p;
    return 0;
}

extern "C" int pla_peer_free(void* dev_ptr) {
    if (dev_ptr != nullptr) PLA_CUDA(cudaFree(dev_ptr));
    return 0;
}

extern "C" int pla_peer_export(const void* dev_ptr, unsigned char* handle64) {
    PLA_CHECK_ARG(dev_ptr != nullptr, 1, "dev_ptr is null");
    PLA_CHECK_ARG(handle64 != nullptr, 2, "handle is null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    PLA_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int pla_peer_import(const unsigned char* handle64, void** peer_ptr) {
    PLA_CHECK_ARG(handle64 != nullptr, 1, "handle is null");
    PLA_CHECK_ARG(peer_ptr != nullptr, 2, "peer_ptr is null");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    PLA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = p;
    return 0;
}

extern "C" int pla_peer_close(void* peer_ptr) {
    if (peer_ptr != nullptr) PLA_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}
