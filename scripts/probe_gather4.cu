// Probe for TMA tile::gather4 on sm_100a (minibatch rows gathered inside the GEMM's operand path).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe_gather4 scripts/probe_gather4.cu -lcuda
// 1. which tensor-map box a gather4 wants ({W,1} or {W,4}), where its four rows land in shared memory, and how the
//    128-byte swizzle modes apply to them;
// 2. bandwidth of 128-row x 128-byte operand stages filled by one tile load against 32 gather4 issued by a warp,
//    with identity and with random row indices (M x 616 floats, larger than L2).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

static inline int ck_(cudaError_t e, const char *file, int line) {
    if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line); exit(1); }
    return 0;
}
#define CK(x) ck_((x), __FILE__, __LINE__)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t ph) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *b, uint32_t ph, long long limit) {
    const long long t0 = clock64();
    while (!mbar_try(b, ph))
        if (clock64() - t0 > limit) return false;
    return true;
}
__device__ __forceinline__ void gather4(void *smem, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tile2d(void *smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// ---- 1. layout probe: one gather4 of rows {5, 17, 2, 40}, columns 32..63, into a zeroed 4 KB buffer ----
__global__ void layout_probe(const __grid_constant__ CUtensorMap map, uint32_t expect, float *out, int *done) {
    __shared__ __align__(1024) float buf[1024];
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = -1.0f;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar, expect);
        gather4(buf, &map, 32, 5, 17, 2, 40, &bar);
        *done = mbar_wait_bounded(&bar, 0, 2000000) ? 1 : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}

// ---- 2. bandwidth: stages of 128 rows x 32 floats ----
constexpr int NST = 8, STAGE_FLOATS = 128 * 32;
template <int MODE>      // 0: tile load (box 32 x 128), 1: gather4 x 32 lanes
__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, const int *__restrict__ idx, int m_tiles,
                                                    int k_blocks, float *sink) {
    extern __shared__ __align__(1024) unsigned char raw[];
    float *stages = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[NST], empty[NST];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t it = 0;
    float acc = 0.f;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        int4 rows = make_int4(0, 0, 0, 0);
        if (MODE == 1 && warp == 0) rows = __ldg(reinterpret_cast<const int4 *>(idx + mt * 128) + lane);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
            const uint32_t s = it % NST, ph = (it / NST) & 1;
            if (warp == 0) {
                while (!mbar_try(&empty[s], ph ^ 1)) {}
                if (lane == 0) mbar_expect_tx(&full[s], STAGE_FLOATS * 4);
                __syncwarp();
                if (MODE == 0) {
                    if (lane == 0) tile2d(stages + s * STAGE_FLOATS, &map, kb * 32, mt * 128, &full[s]);
                } else {
                    gather4(stages + s * STAGE_FLOATS + lane * 128, &map, kb * 32, rows.x, rows.y, rows.z, rows.w, &full[s]);
                }
            } else {
                while (!mbar_try(&full[s], ph)) {}
                acc += stages[s * STAGE_FLOATS + lane * 33];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUresult encode(EncodeFn enc, CUtensorMap *map, float *base, long rows, int cols, int ld, int bw, int bh, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    // ---------------- layout ----------------
    const int R = 64, C = 64;
    std::vector<float> h(R * C);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) h[r * C + c] = r * 100.0f + c;         // value = 100 row + col
    float *d, *out;
    int *done;
    CK(cudaMalloc(&d, R * C * 4)), CK(cudaMalloc(&out, 4096)), CK(cudaMalloc(&done, 4));
    CK(cudaMemcpy(d, h.data(), R * C * 4, cudaMemcpyHostToDevice));
    struct Variant { const char *name; int bh; CUtensorMapSwizzle sw; uint32_t expect; };
    const Variant variants[] = {
        {"box{32,1} swizzle none", 1, CU_TENSOR_MAP_SWIZZLE_NONE, 512},
        {"box{32,1} swizzle 128B", 1, CU_TENSOR_MAP_SWIZZLE_128B, 512},
        {"box{32,1} swizzle 128B atom 32B", 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 512},
        {"box{32,1} swizzle none, expect 128", 1, CU_TENSOR_MAP_SWIZZLE_NONE, 128},
        {"box{32,4} swizzle none", 4, CU_TENSOR_MAP_SWIZZLE_NONE, 512},          // may fault: runs last
    };
    auto run_variant = [&](const Variant &v) {
        CUtensorMap map;
        CUresult r = encode(enc, &map, d, R, C, C, 32, v.bh, v.sw);
        printf("== %s: encode rc=%d\n", v.name, (int)r);
        if (r != CUDA_SUCCESS) return 0;
        CK(cudaMemset(done, 0xff, 4));
        layout_probe<<<1, 128>>>(map, v.expect, out, done);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<float> o(1024);
        int hd;
        CK(cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost)), CK(cudaMemcpy(&hd, done, 4, cudaMemcpyDeviceToHost));
        int written = 0;
        for (float x : o) written += x >= 0.0f;
        printf("   barrier completed=%d floats written=%d\n", hd, written);
        for (int row = 0; row < 8; ++row) {          // 128-byte lines of the buffer: first float of each 16-byte chunk
            printf("   line %d:", row);
            for (int ch = 0; ch < 8; ++ch) printf(" %7.0f", o[row * 32 + ch * 4]);
            printf("\n");
        }
        return 0;
    };
    for (int i = 0; i < 4; ++i)
        if (run_variant(variants[i])) return 2;
    // ---------------- bandwidth ----------------
    const long M = 98304;
    const int K = 615, LD = 616, KB = (K + 31) / 32;
    float *big, *sink;
    int *idx;
    CK(cudaMalloc(&big, M * LD * 4)), CK(cudaMalloc(&sink, 4)), CK(cudaMalloc(&idx, M * 4));
    CK(cudaMemset(big, 0, M * LD * 4));
    std::vector<int> hi(M);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)), CK(cudaEventCreate(&e1));
    CUtensorMap map_tile, map_g;
    if (encode(enc, &map_tile, big, M, K, LD, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS) { printf("tile map failed\n"); return 3; }
    if (encode(enc, &map_g, big, M, K, LD, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS) { printf("gather map failed\n"); return 3; }
    const size_t smem = NST * STAGE_FLOATS * 4 + 1024;
    CK(cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double bytes = (double)M * KB * 128;      // M rows x KB blocks x 128 bytes
    for (int mode = 0; mode < 3; ++mode) {
        for (long i = 0; i < M; ++i) hi[i] = (int)i;
        if (mode == 2) {
            srand(1);
            for (long i = M - 1; i > 0; --i) { long j = rand() % (i + 1); int t = hi[i]; hi[i] = hi[j]; hi[j] = t; }
        }
        CK(cudaMemcpy(idx, hi.data(), M * 4, cudaMemcpyHostToDevice));
        for (int grid : {148, 296}) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaEventRecord(e0));
                if (mode == 0) stream_kernel<0><<<grid, 64, smem>>>(map_tile, idx, (int)(M / 128), KB, sink);
                else stream_kernel<1><<<grid, 64, smem>>>(map_g, idx, (int)(M / 128), KB, sink);
                CK(cudaEventRecord(e1));
                cudaError_t e = cudaEventSynchronize(e1);
                if (e != cudaSuccess) { printf("stream kernel error: %s\n", cudaGetErrorString(e)); return 4; }
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                best = ms < best ? ms : best;
            }
            printf("stream %-26s grid %3d: %.3f ms  %.0f GB/s\n", mode == 0 ? "tile load 128x32" : (mode == 1 ? "gather4 x32, identity rows" : "gather4 x32, random rows"),
                   grid, best, bytes / best * 1e-6);
        }
    }
    fflush(stdout);
    return run_variant(variants[4]);
}
