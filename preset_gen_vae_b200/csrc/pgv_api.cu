// Library-level entry points of include/pgv.h: version, errors, handle, TMA descriptor helper.
#include <string.h>

#include "pgv_common.cuh"

namespace pgv {

static thread_local char g_last_error[512] = "";
int g_use_pdl = 0;      // measured on B200: with the attribute the captured step was 0.14 ms SLOWER (early-resident CTAs); kept as a switch

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int make_tmap_f32(const pgv_handle* h, CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box) {
    if (!h || !h->encode_tiled) return set_error(-2, "pgv handle has no cuTensorMapEncodeTiled");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(-1, "TMA base address must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            if (gstr[i - 1] % 16 != 0) return set_error(-1, "TMA stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i - 1]);
        }
    }
    CUresult r = h->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                                 gdim, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(static_cast<int>(r), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

}  // namespace pgv

extern "C" {

int pgv_version(void) { return PGV_VERSION; }

const char* pgv_last_error(void) { return pgv::g_last_error; }

int pgv_init(pgv_handle** out, int device) {
    if (!out) return pgv::set_error(-1, "pgv_init: out is NULL");
    *out = nullptr;
    int count = 0;
    PGV_CUDA(cudaGetDeviceCount(&count));
    PGV_CHECK_ARG(device >= 0 && device < count, "pgv_init: device %d out of range (%d visible)", device, count);
    cudaDeviceProp prop;
    PGV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return pgv::set_error(-3, "pgv_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                              prop.major, prop.minor);
    pgv_handle* h = new pgv_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
        delete h;
        return pgv::set_error(-4, "pgv_init: cannot resolve cuTensorMapEncodeTiled (%s)", cudaGetErrorString(e));
    }
    h->encode_tiled = reinterpret_cast<decltype(h->encode_tiled)>(fn);
    fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
    h->encode_im2col = (e == cudaSuccess && q == cudaDriverEntryPointSuccess) ? reinterpret_cast<decltype(h->encode_im2col)>(fn) : nullptr;
    h->driver_version = 0;
    cudaDriverGetVersion(&h->driver_version);
    *out = h;
    return 0;
}

void pgv_destroy(pgv_handle* h) { delete h; }

int pgv_sm_count(const pgv_handle* h) { return h ? h->sm_count : 0; }

int pgv_debug_set_pdl(int on) { pgv::g_use_pdl = on ? 1 : 0; return 0; }

}  // extern "C"
