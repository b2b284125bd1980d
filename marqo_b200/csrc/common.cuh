// Host-side helpers shared by every translation unit of libmarqo_b200.so:
// status codes, thread-local error text, CUDA error checks, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/marqo_b200.h"

namespace mb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define MB_CUDA(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            ::mb::fail(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define MB_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) ::mb::fail(B200_ERR_INVALID_ARG, __VA_ARGS__); \
    } while (0)

// Wraps a C-ABI body: exceptions -> status code + thread-local message.
template <class F>
int guarded(F&& f) {
    try {
        f();
        return B200_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return B200_ERR_INTERNAL;
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        MB_CUDA(cudaGetDevice(&prev));
        if (prev != dev) MB_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// 2D row-major tensor map: inner dim `cols` (contiguous), outer dim `rows`, row pitch in bytes.
// box = {box_cols, box_rows}; swizzle 128B requires box_cols * elem_size == 128.
CUtensorMap make_tmap_2d(const void* base, CUtensorMapDataType dtype, uint32_t elem_bytes, uint64_t cols,
                         uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols, uint32_t box_rows,
                         CUtensorMapSwizzle swizzle);

int sm_count(int device);

inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mb
