// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/wvd.h"

namespace wvd {

// thread-local error string, returned by wvd_last_error()
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);

#define WVD_CHECK_CUDA(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return ::wvd::set_error(WVD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                          \
    } while (0)

#define WVD_REQUIRE(cond, ...)                                            \
    do {                                                                  \
        if (!(cond)) return ::wvd::set_error(WVD_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// 2-D bf16 row-major tensor map with 128-byte swizzle.
//   rows x cols elements, row stride ld (elements); box = box_rows x 64 columns (64 bf16 = 128 B).
// Out-of-bounds elements are filled with zeros on load and clipped on store.
// Maps are cached by (ptr, rows, cols, ld, box_rows): encoding costs ~1 us, the cache makes it ~50 ns.
int get_tensor_map_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols = 64);

int sm_count();

// One-time per-DEVICE setup (cudaFuncSetAttribute is per device; a process may drive several GPUs): returns true the
// first time it is called with this flag word on the current device.  Thread-safe.
bool first_use_on_current_device(unsigned long long* flag_word);

}  // namespace wvd
