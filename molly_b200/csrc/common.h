// Host-side helpers shared by the translation units: error reporting, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace molly {

// thread-local last error string, surfaced through molly_last_error()
void set_last_error(const std::string& msg);
const char* get_last_error();

#define MOLLY_CHECK(cond, code, ...)                              \
    do {                                                          \
        if (!(cond)) {                                            \
            char _buf[512];                                       \
            snprintf(_buf, sizeof(_buf), __VA_ARGS__);            \
            ::molly::set_last_error(_buf);                        \
            return (code);                                        \
        }                                                         \
    } while (0)

#define MOLLY_CUDA(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            char _buf[512];                                                                           \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                     __FILE__, __LINE__);                                                             \
            ::molly::set_last_error(_buf);                                                            \
            return MOLLY_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

enum { MOLLY_OK = 0, MOLLY_ERR_INVALID = 1, MOLLY_ERR_CUDA = 2, MOLLY_ERR_UNSUPPORTED = 3, MOLLY_ERR_WORKSPACE = 4 };

// 2-D row-major tensor map: `rows` x `cols` elements of `elem_bytes`, leading dimension `ld` elements,
// box = box_rows x box_cols, swizzle span = box_cols * elem_bytes (must be 32/64/128 B) or none.
// Out-of-bounds box elements are zero-filled on load.
int make_tma_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                uint32_t box_cols, uint32_t elem_bytes, bool swizzle = true);

int device_sm_count();

void prof_start();
int prof_stop(int* launches, double* ms, double* work);

}  // namespace molly
