// Peer-visible device memory for the Z-sharded pass.
//
// The exchange steps of the sharded pass are written by the producing kernels straight into the
// consumer GPU's memory over NVLink / NVSwitch (plain stores + a release flag); the consumer kernel
// spins on the flag in its own HBM.  That needs buffers every rank of the box can map: they are
// cudaMalloc'ed here (legacy CUDA IPC cannot export sub-allocations of a caching allocator, nor
// expandable segments) and exchanged as 64-byte IPC handles by the caller (skoots_b200/sharded.py,
// over torch.distributed).  These five calls are the only ones in the library that allocate or
// synchronise; they are set-up / tear-down, never on the per-pass path.
#include <string.h>

#include "skb_common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == SKB_PEER_HANDLE_BYTES, "IPC handle size");

#define SKB_CUDA_TRY(call, what)                                                  \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            skb_set_error("%s: CUDA error %s", what, cudaGetErrorString(e__));    \
            (void)cudaGetLastError();                                             \
            return SKB_E_CUDA;                                                    \
        }                                                                         \
    } while (0)

extern "C" int skb_peer_alloc(size_t bytes, void** ptr) {
    SKB_REQUIRE(ptr && bytes > 0, "skb_peer_alloc: bad argument");
    void* p = nullptr;
    SKB_CUDA_TRY(cudaMalloc(&p, bytes), "skb_peer_alloc");
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        skb_set_error("skb_peer_alloc: CUDA error %s", cudaGetErrorString(e));
        return SKB_E_CUDA;
    }
    *ptr = p;
    return SKB_OK;
}

extern "C" int skb_peer_free(void* ptr) {
    if (!ptr) return SKB_OK;
    SKB_CUDA_TRY(cudaFree(ptr), "skb_peer_free");
    return SKB_OK;
}

extern "C" int skb_peer_export(void* ptr, uint8_t handle[SKB_PEER_HANDLE_BYTES]) {
    SKB_REQUIRE(ptr && handle, "skb_peer_export: NULL pointer");
    cudaIpcMemHandle_t h;
    SKB_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr), "skb_peer_export");
    memcpy(handle, &h, sizeof(h));
    return SKB_OK;
}

extern "C" int skb_peer_open(const uint8_t handle[SKB_PEER_HANDLE_BYTES], void** ptr) {
    SKB_REQUIRE(ptr && handle, "skb_peer_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    // the flag also enables peer access from the current device to the exporting one
    SKB_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "skb_peer_open");
    *ptr = p;
    return SKB_OK;
}

extern "C" int skb_peer_close(void* ptr) {
    if (!ptr) return SKB_OK;
    SKB_CUDA_TRY(cudaIpcCloseMemHandle(ptr), "skb_peer_close");
    return SKB_OK;
}
