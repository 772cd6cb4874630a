// Host-side helpers shared by all translation units of libk5: error reporting, TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/k5.h"

namespace k5 {

typedef __nv_bfloat16 bf16;

// Error codes returned through the C ABI come from include/k5.h (K5_OK, K5_ERR_*).

void set_last_error(const std::string& msg);
const char* get_last_error();

#define K5_CHECK_CUDA(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::k5::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return K5_ERR_CUDA;                                                               \
        }                                                                                           \
    } while (0)

#define K5_REQUIRE(cond, msg)                                                                       \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            ::k5::set_last_error(std::string(msg) + " [" #cond "]");                                \
            return K5_ERR_INVALID;                                                            \
        }                                                                                           \
    } while (0)

#define K5_TRY(expr)                                                                                \
    do {                                                                                            \
        int _rc = (expr);                                                                           \
        if (_rc != 0) return _rc;                                                                   \
    } while (0)

// Row-major 2-D bf16 tensor [rows, cols] with a row pitch of ld elements -> TMA map with a
// [box_rows x 64]-element box and 128-byte swizzle (64 bf16 = 128 B, one swizzle atom wide).
// Out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// Per-device state.  cudaFuncSetAttribute, the cluster occupancy and the SM count belong to a DEVICE, and one process
// may drive several (the pipeline's device_map puts the DiT and the VAE on different GPUs, kandinsky/utils.py:23-34),
// so every "configure once" in the launchers is keyed by the current device and guarded by a mutex.
constexpr int K5_MAX_DEVICES = 64;
int current_device();               // cudaGetDevice, clamped to [0, K5_MAX_DEVICES)
template <typename T>
struct PerDevice {
    std::mutex m;
    bool set[K5_MAX_DEVICES] = {};
    T v[K5_MAX_DEVICES] = {};
};
int sm_count();                     // of the current device
// cuStreamWriteValue32 / cuStreamWaitValue32 (>=, wrap-safe 32-bit compare) on `st`; addr may be peer-mapped memory
int stream_write_u32(cudaStream_t st, uint32_t* addr, uint32_t value);
int stream_wait_geq_u32(cudaStream_t st, const uint32_t* addr, uint32_t value);

}  // namespace k5
