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
int g_gemm_snake = 1;
int g_coop_launch = 0;
int g_gae_serial_min_envs = 8192;
int g_gae_threads = 0;

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
int hb_sizeof_gemm_desc(void) { return (int)sizeof(hb_gemm_desc); }
int hb_sizeof_adam_params(void) { return (int)sizeof(hb_adam_params); }
int hb_sizeof_optim_state(void) { return (int)sizeof(hb_optim_state); }
int hb_sizeof_dp_comm(void) { return (int)sizeof(hb_dp_comm); }

// Observation tensors are [N, width] views of rows at a 128-byte pitch: a host (or dense-buffer) consumer copies the valid
// columns of all rows with one 2-D DMA transfer instead of the whole pitched block.
int hb_copy_rows(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes, int64_t width_bytes, int64_t rows,
                 void *stream) {
    HB_REQUIRE(dst && src && rows > 0 && width_bytes > 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
               "hb_copy_rows: bad arguments");
    HB_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)rows,
                              cudaMemcpyDefault, (cudaStream_t)stream));
    return HB_OK;
}

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

// ---- CUDA graphs of this library's launches ------------------------------------------------------------------
// A host that replays a fixed sequence of hb_* calls (one env step) can capture it here instead of through its
// framework's graph object: no allocator / generator bookkeeping around capture and replay (torch.cuda.graph
// synchronises, collects garbage and empties its cache on every capture - the step graphs are captured lazily, one per
// pair of observation buffers).  The capture must contain hb_* calls only (nothing that allocates); `stream` must not
// be the legacy default stream.
int hb_graph_begin(void *stream) {
    HB_REQUIRE(stream != nullptr, "hb_graph_begin: capture needs a non-default stream");
    HB_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return HB_OK;
}

int hb_graph_end(void *stream, void **graph_exec) {
    HB_REQUIRE(stream != nullptr && graph_exec != nullptr, "hb_graph_end: null argument");
    cudaGraph_t graph = nullptr;
    HB_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        hb::set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
        return HB_ERR_CUDA;
    }
    *graph_exec = exec;
    return HB_OK;
}

int hb_graph_launch(void *graph_exec, void *stream) {
    HB_REQUIRE(graph_exec != nullptr, "hb_graph_launch: null graph");
    HB_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return HB_OK;
}

int hb_graph_destroy(void *graph_exec) {
    if (graph_exec) HB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return HB_OK;
}

}  // extern "C"
