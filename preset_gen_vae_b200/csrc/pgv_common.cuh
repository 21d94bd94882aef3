// Host-side plumbing shared by all translation units: handle, error reporting, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pgv.h"

struct pgv_handle {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    // cuTensorMapEncodeTiled, resolved through the runtime so that libpgv.so does not link libcuda directly
    CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    // cuTensorMapEncodeIm2col (activation operand of the channels-last convolutions); NULL if the driver does not export it
    CUresult (*encode_im2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                              const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    int driver_version;
};

namespace pgv {

// Hardtanh / clamp that lets NaN through like torch's (fminf / fmaxf would replace it by a bound and hide a diverged model from the
// NaN guard of train.py:245)
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

int set_error(int code, const char* fmt, ...);

#define PGV_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return pgv::set_error(-1, __VA_ARGS__);      \
    } while (0)

#define PGV_CUDA(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return pgv::set_error(static_cast<int>(e__), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                  __FILE__, __LINE__);                                                      \
    } while (0)

#define PGV_LAUNCH_CHECK() PGV_CUDA(cudaGetLastError())

// fp32 tensor map, 128-byte swizzle, zero fill out of bounds.  dims[0] is the contiguous dimension; strides_bytes has
// rank-1 entries (dimension 0 is implicitly 4 bytes).  box[0] must be 32 (128 bytes).
int make_tmap_f32(const pgv_handle* h, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

extern int g_use_pdl;        // pgv_debug_set_pdl: programmatic dependent launch of the PDL-aware kernels (default on)

// Launch of a PDL-aware kernel (one that starts with griddepcontrol.launch_dependents and executes griddepcontrol.wait before its
// first dependent access): lets it overlap its launch and prologue with the tail of its predecessor in the stream.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pgv
