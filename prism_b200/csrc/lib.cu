// lib.cu -- library-level entry points of libprism_b200.so
#include "common.cuh"

std::atomic<long long> g_pb_launches{0};

int pb_sm_count()
{
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    }
    cached = n;
    return n;
}

extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

long long pb_launch_count(void) { return g_pb_launches.load(std::memory_order_relaxed); }

const char *pb_error_string(int code)
{
    switch (code) {
        case PB_OK: return "ok";
        case PB_E_ARG: return "invalid argument";
        case PB_E_CAPACITY: return "capacity must be a power of two in [2, 2^30] with size <= capacity";
        case PB_E_UNSUPPORTED: return "unsupported option";
        case PB_E_POOL: return "aux observation pool exhausted";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

}  // extern "C"
