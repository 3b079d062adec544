// Peer-memory plumbing of the multi-GPU fused update: one process per GPU, every rank cudaMalloc's its
// exchange buffer, publishes the CUDA IPC handle (the host exchanges the 64 opaque bytes through
// torch.distributed) and maps the peers' buffers.  The data path itself is inside ppo_tc_kernel
// (mlp_tc.cu): plain 8-byte stores / loads on these pointers over NVLink.
#include <cuda_runtime.h>

#include <cstring>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == TS_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int ts_peer_alloc(int64_t bytes, void** ptr_out, uint8_t* handle_out) {
    TS_REQUIRE(bytes > 0 && ptr_out && handle_out, "ts_peer_alloc: bad arguments");
    void* p = nullptr;
    TS_CUDA(cudaMalloc(&p, (size_t)bytes));
    TS_CUDA(cudaMemset(p, 0, (size_t)bytes));
    TS_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        tsb::set_error("ts_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return 1;
    }
    std::memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return 0;
}

extern "C" int ts_peer_open(const uint8_t* handle, void** ptr_out) {
    TS_REQUIRE(handle && ptr_out, "ts_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    TS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return 0;
}

extern "C" int ts_peer_close(void* ptr) {
    if (ptr) TS_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int ts_peer_free(void* ptr) {
    if (ptr) TS_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int64_t ts_ppo_peer_buffer_bytes(const ts_actor_critic_desc* desc, int32_t world) {
    if (!desc || world < 1 || world > tsb::kMaxPeers) return 0;
    const int64_t width = desc->n_params + TS_PPO_GRAD_EXTRA;
    return (int64_t)tsb::kPeerHeaderBytes + (int64_t)world * 2 * width * 8;
}
