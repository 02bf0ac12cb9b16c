// Library-wide state: version, thread-local error string, cached device properties.
#include "gn_common.cuh"
#include <stdarg.h>
#include <mutex>

static thread_local char g_err[512] = "";

void gn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

GN_API const char* gn_last_error(void) { return g_err; }

GN_API int gn_version(void) { return 100; }  // 0.1.0

int gn_num_sms() {
    static int sms = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) { sms = 148; return; }
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    });
    return sms;
}

GN_API int gn_device_sm_count(void) { return gn_num_sms(); }

static int g_pdl = -1;
int gn_pdl_enabled() {
    if (g_pdl < 0) g_pdl = gn_env_flag("GN_NO_PDL") ? 0 : 1;
    return g_pdl;
}
// Switch programmatic dependent launch on / off for every later launch of this process (returns the previous setting).  The host
// side switches it OFF in data-parallel runs: with kernels that trigger their dependents early, the 2-GPU step (NCCL all-reduce
// between backward and optimizer) hung on the B200 pool, while every single-process configuration ran clean.
GN_API int gn_set_pdl(int on) {
    const int prev = gn_pdl_enabled();
    g_pdl = (on && !gn_env_flag("GN_NO_PDL")) ? 1 : 0;
    return prev;
}

// ------------------------------------------------------------------------------------------------
#include "gn_tma.cuh"
typedef CUresult (*gn_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static gn_encode_fn get_encode() {
    static gn_encode_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (gn_encode_fn)p;
    });
    return fn;
}

int gn_tmap_encode(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* gaddr, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
    gn_encode_fn fn = get_encode();
    GN_REQUIRE(fn != nullptr, GN_EDRIVER, "cuTensorMapEncodeTiled is not available from the driver");
    GN_REQUIRE(((uintptr_t)gaddr & 15) == 0, GN_EALIGN, "TMA: global address %p is not 16-byte aligned", gaddr);
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        gs[i] = strides_bytes[i];
        GN_REQUIRE((gs[i] & 15) == 0, GN_EALIGN, "TMA: stride %llu of dim %d is not a multiple of 16 bytes", (unsigned long long)gs[i], i + 1);
    }
    CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(gaddr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);   // 256B promotion over-fetches column slices of the concat buffers
    GN_REQUIRE(r == CUDA_SUCCESS, GN_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
    return GN_OK;
}
