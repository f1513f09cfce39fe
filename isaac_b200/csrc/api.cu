// Library-level glue of libhectorb200.so: error strings, launch counter, device probe.
#include <stdarg.h>

#include <atomic>

#include "hb_common.cuh"

namespace hb {

thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int g_use_pdl = -1;
int g_gemm_pdl = 1;
int g_gae_serial_min_envs = 8192;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace hb

extern "C" {

const char *hb_last_error(void) { return hb::g_error; }
int hb_abi_version(void) { return HB_ABI_VERSION; }
int64_t hb_launch_count(void) { return hb::g_launches.load(std::memory_order_relaxed); }
void hb_launch_count_reset(void) { hb::g_launches.store(0, std::memory_order_relaxed); }

// struct sizes, so that a foreign-language binding can verify its mirror of the header
int hb_sizeof_env_params(void) { return (int)sizeof(hb_env_params); }
int hb_sizeof_env_buffers(void) { return (int)sizeof(hb_env_buffers); }
int hb_sizeof_env_noise(void) { return (int)sizeof(hb_env_noise); }

// The kernels are sm_100a-only (no other cubin or PTX is embedded): refuse anything else loudly.
int hb_check_device(void) {
    int dev = 0;
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDevice(&dev));
    HB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        hb::set_error("libhectorb200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
        return HB_ERR_UNSUPPORTED;
    }
    return HB_OK;
}

}  // extern "C"
