#include "common.h"

#include <cstring>
#include <unordered_map>

#include <mutex>

namespace k5 {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // The driver entry point avoids a link-time dependency on libcuda.so.
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

// Stream-ordered 32-bit memory operations (cuStreamWriteValue32 / cuStreamWaitValue32): executed by the stream's
// front end, so they make progress while persistent kernels hold every SM - which a one-thread signalling kernel
// launched on a second stream does not (the temporal shard's overlapped all-gather depends on it).
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamValueFn get_stream_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<StreamValueFn>(p);
    return nullptr;
}
int stream_write_u32(cudaStream_t st, uint32_t* addr, uint32_t value) {
    static StreamValueFn fn = get_stream_fn("cuStreamWriteValue32");
    K5_REQUIRE(fn != nullptr, "cuStreamWriteValue32 entry point not available");
    const CUresult r = fn(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuStreamWriteValue32 failed with CUresult " + std::to_string(static_cast<int>(r)));
        return K5_ERR_CUDA;
    }
    return K5_OK;
}
int stream_wait_geq_u32(cudaStream_t st, const uint32_t* addr, uint32_t value) {
    static StreamValueFn fn = get_stream_fn("cuStreamWaitValue32");
    K5_REQUIRE(fn != nullptr, "cuStreamWaitValue32 entry point not available");
    const CUresult r = fn(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(const_cast<uint32_t*>(addr)), value,
                          CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuStreamWaitValue32 failed with CUresult " + std::to_string(static_cast<int>(r)));
        return K5_ERR_CUDA;
    }
    return K5_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return K5_ERR_CUDA;
    }
    K5_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    K5_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
    K5_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
    // A forward encodes ~1000 maps and all of them recur every forward (same engine buffers, same shapes): the
    // descriptor is a pure function of (device, base, shape, pitch, box), so it is cached instead of re-encoded.
    struct Key {
        uint64_t v[6];
        bool operator==(const Key& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            uint64_t h = 0x9e3779b97f4a7c15ull;
            for (uint64_t x : k.v) h = (h ^ x) * 0xff51afd7ed558ccdull + (h >> 29);
            return static_cast<size_t>(h);
        }
    };
    static std::mutex m;
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    const Key key = {{reinterpret_cast<uint64_t>(base), rows, cols, ld, box_rows, static_cast<uint64_t>(current_device())}};
    {
        std::lock_guard<std::mutex> lk(m);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return K5_OK;
        }
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return K5_ERR_CUDA;
    }
    std::lock_guard<std::mutex> lk(m);
    if (cache.size() >= 8192) cache.clear();        // callers with ever-changing pointers (tests) must not grow it
    cache.emplace(key, *out);
    return K5_OK;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < K5_MAX_DEVICES ? dev : K5_MAX_DEVICES - 1;
}

int sm_count() {
    static PerDevice<int> pd;
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(pd.m);
    if (!pd.set[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        pd.v[dev] = n;
        pd.set[dev] = true;
    }
    return pd.v[dev];
}

}  // namespace k5
