// Process-wide plumbing of libwiw_b200.so: last-error string, device probe, TMA descriptor encoder.
#include "wiw_common.cuh"
#include "../../include/wiw.h"

#include <mutex>

namespace wiw {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }

int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : 0;
}

// per-device cache: one process may drive engines on several GPUs
int num_sms() {
    static int cached[64] = {0};
    const int dev = current_device() & 63;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

bool first_call_on_device(unsigned long long* mask) {
    const unsigned long long bit = 1ull << (current_device() & 63);
    if (*mask & bit) return false;
    *mask |= bit;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
        return 1;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[512];
        snprintf(buf, sizeof(buf),
                 "cuTensorMapEncodeTiled failed (%d): rank=%d base=%p dims=[%llu,%llu,%llu,%llu,%llu] "
                 "strides=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                 (int)r, rank, base, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                 (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                 (unsigned long long)(rank > 4 ? gdim[4] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
                 (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0),
                 (unsigned long long)(rank > 4 ? gstr[3] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
                 rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
        set_last_error(buf);
        return 1;
    }
    return 0;
}

}  // namespace wiw

extern "C" const char* wiw_last_error(void) { return wiw::g_last_error.c_str(); }

extern "C" int wiw_abi_version(void) { return 1; }

extern "C" int wiw_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
