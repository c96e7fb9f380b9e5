// common.cuh — error plumbing, launch counter and host-pointer staging shared by the C-ABI files.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/padeops_b200.h"

namespace pdo {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define PDO_CUDA(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return ::pdo::fail(                                                                               \
                _e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? PDO_E_NODEVICE : PDO_E_CUDA,   \
                "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));                         \
    } while (0)

inline bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Grow-only device scratch for calls that arrive with HOST pointers (the drop-in path).
struct StagePool {
    void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[4] = {0, 0, 0, 0};
    int get(int slot, size_t bytes, void** out) {
        if (cap[slot] < bytes) {
            if (buf[slot]) cudaFree(buf[slot]);
            buf[slot] = nullptr;
            cap[slot] = 0;
            PDO_CUDA(cudaMalloc(&buf[slot], bytes));
            cap[slot] = bytes;
        }
        *out = buf[slot];
        return 0;
    }
};
StagePool& stage_pool();

// Runs body(dev_in, dev_out) with device views of (in, out); stages through the pool when they are host memory.
template <class Body>
int with_device_views(const void* in, size_t in_bytes, void* out, size_t out_bytes, cudaStream_t st, Body body) {
    const bool in_dev = is_device_ptr(in), out_dev = is_device_ptr(out);
    const void* din = in;
    void* dout = out;
    if (!in_dev) {
        void* b = nullptr;
        if (int rc = stage_pool().get(0, in_bytes, &b)) return rc;
        PDO_CUDA(cudaMemcpyAsync(b, in, in_bytes, cudaMemcpyHostToDevice, st));
        din = b;
    }
    if (!out_dev) {
        void* b = nullptr;
        if (int rc = stage_pool().get(1, out_bytes, &b)) return rc;
        dout = b;
    }
    if (int rc = body(din, dout)) return rc;
    if (!out_dev) {
        PDO_CUDA(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, st));
        PDO_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

}  // namespace pdo
