// C-ABI plumbing: error reporting, library info, TMA tensor-map creation (driver entry point resolved at run time,
// so libwvd.so has no link-time dependency on libcuda and loads on a CPU-only box for the symbol checks).
#include <stdarg.h>
#include <string.h>

#include "host_utils.h"

namespace wvd {

std::string& last_error_ref() {
    thread_local std::string err;
    return err;
}

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows, box_cols;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
        h ^= (k.cols * 31 + k.ld * 131 + k.box_rows * 17 + k.box_cols + (h << 6) + (h >> 2));
        return static_cast<size_t>(h);
    }
};

int get_tensor_map_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows, box_cols};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            memcpy(out, &it->second, sizeof(CUtensorMap));
            return WVD_OK;
        }
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(WVD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
        return set_error(WVD_ERR_INVALID, "tensor map: base pointer and row pitch must be 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(WVD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 8192) cache.clear();
        cache[key] = m;
    }
    memcpy(out, &m, sizeof(CUtensorMap));
    return WVD_OK;
}

bool first_use_on_current_device(unsigned long long* flag_word) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;      // unknown: configure again (idempotent)
    const unsigned long long bit = 1ull << dev;
    return (__atomic_fetch_or(flag_word, bit, __ATOMIC_ACQ_REL) & bit) == 0;
}

int gemm_read_diag(unsigned long long* out);
int attn_read_diag(unsigned long long* out);
int attn_pair_read_diag(unsigned long long* out);
int attn_cg2_read_diag(unsigned long long* out);

}  // namespace wvd

extern "C" __attribute__((visibility("default"))) const char* wvd_last_error(void) { return wvd::last_error_ref().c_str(); }
extern "C" __attribute__((visibility("default"))) int wvd_version(void) { return 200; /* 0.2.0 */ }
#ifndef WVD_SOURCE_HASH
#define WVD_SOURCE_HASH "unknown"
#endif
extern "C" __attribute__((visibility("default"))) const char* wvd_build_info(void) {
    return "wvd 0.2.0 sm_100a src=" WVD_SOURCE_HASH;
}
extern "C" __attribute__((visibility("default"))) int wvd_sm_arch(void) { return 100; }

extern "C" __attribute__((visibility("default"))) int wvd_debug_flags(unsigned long long out[8]) {
    using namespace wvd;
    // a kernel whose watchdog fired has trapped: the synchronise below then reports the sticky launch failure
    WVD_CHECK_CUDA(cudaDeviceSynchronize());
    unsigned long long a[8] = {0}, b[8] = {0}, c[8] = {0};
    if (gemm_read_diag(a) != 0 || attn_read_diag(b) != 0 || attn_pair_read_diag(c) != 0)
        return set_error(WVD_ERR_CUDA, "reading diagnostics failed");
    if (c[0] != 0) { b[0] += c[0]; b[1] = c[1]; b[2] = c[2]; b[3] = c[3]; }      // all attention kernels report as "attn"
    unsigned long long d[8] = {0};
    if (attn_cg2_read_diag(d) != 0) return set_error(WVD_ERR_CUDA, "reading diagnostics failed");
    if (d[0] != 0) { b[0] += d[0]; b[1] = d[1]; b[2] = d[2]; b[3] = d[3]; }
    for (int i = 0; i < 8; ++i) out[i] = a[i];
    out[0] = a[0] + b[0];
    if (b[0] != 0) { out[1] = b[1]; out[2] = b[2]; out[3] = b[3]; }
    out[4] = a[0];
    out[5] = b[0];
    return WVD_OK;
}
