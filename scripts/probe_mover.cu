// Probe: how fast can ONE CTA per SM (a few "mover" warps) shift its tile's frame-stack rows?  (The fused post-physics launch
// gives every 32-env tile mover warps; at 4096 envs there is one tile per SM.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe_mover scripts/probe_mover.cu
// Variants: cp.async into a private shared-memory ring (depth x rows per group), plain register loads (unroll), mover warps.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_stream4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_stream4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int ROT>
__device__ __forceinline__ float4 rot4(const float4 &v, const float4 &w) {
    if (ROT == 1) return make_float4(v.y, v.z, v.w, w.x);
    if (ROT == 2) return make_float4(v.z, v.w, w.x, w.y);
    if (ROT == 3) return make_float4(v.w, w.x, w.y, w.z);
    return v;
}

// ---- variant 1: cp.async, private 32-byte slots, DEPTH groups of R rows in flight ----
template <int ROW, int LD, int FRAME, int MT, int DEPTH, bool CA>
__device__ __forceinline__ void mover_async(const float *prev, float *next, int nv, uint32_t mtid, float4 *ring, int R) {
    constexpr uint32_t KEEP = ROW - FRAME, NV = (KEEP + 3) / 4, FVEC = FRAME >> 2, ROT = FRAME & 3, JT = (NV + MT - 1) / MT, SLOT = MT * 2;
    const uint32_t cap = DEPTH * (uint32_t)R * JT, ring_s = smem_u32(ring);
    const float *src = prev + 4u * (mtid + FVEC);
    float *dst = next + 4u * mtid;
    const int ngroups = (nv + R - 1) / R;
    uint32_t s_i = 0, s_c = 0;
    int r_i = 0, r_c = 0;
    auto issue = [&]() {
        for (int g = 0; g < R; ++g, ++r_i) {
            if (r_i < nv) {
#pragma unroll
                for (uint32_t q = 0; q < JT; ++q) {
                    if (q * MT + mtid < NV) {
                        const uint32_t d = ring_s + s_i * (SLOT * 16u);
                        if (CA) {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 4u * q * MT) : "memory");
                            if (ROT) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16u), "l"(src + 4u * q * MT + 4) : "memory");
                        } else {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + 4u * q * MT) : "memory");
                            if (ROT) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u), "l"(src + 4u * q * MT + 4) : "memory");
                        }
                    }
                    s_i = (s_i + 1 == cap) ? 0 : s_i + 1;
                }
                src += LD;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto consume = [&]() {
        for (int g = 0; g < R; ++g, ++r_c) {
            if (r_c < nv) {
#pragma unroll
                for (uint32_t q = 0; q < JT; ++q) {
                    const uint32_t cv = q * MT + mtid;
                    if (cv < NV) {
                        const float4 v = ring[s_c * SLOT], w = ring[s_c * SLOT + 1];
                        const float4 o = rot4<ROT>(v, w);
                        if (4u * cv + 4u <= KEEP) st_stream4(reinterpret_cast<float4 *>(dst + 4u * q * MT), o);
                        else { const float ov[4] = {o.x, o.y, o.z, o.w};
                               for (uint32_t k = 0; k < 4u; ++k) if (4u * cv + k < KEEP) dst[4u * q * MT + k] = ov[k]; }
                    }
                    s_c = (s_c + 1 == cap) ? 0 : s_c + 1;
                }
                dst += LD;
            }
        }
    };
#pragma unroll 1
    for (int gi = 0; gi < DEPTH; ++gi) issue();
#pragma unroll 1
    for (int gi = 0; gi < ngroups; ++gi) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        consume();
        issue();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int MT, int DEPTH, bool CA>
__global__ void __launch_bounds__(MT) k_async(const float *pa, float *na, const float *pb, float *nb, int n, int R) {
    extern __shared__ __align__(128) float sm[];
    const int env0 = blockIdx.x * 32, nv = min(32, n - env0);
    float4 *ring = reinterpret_cast<float4 *>(sm) + threadIdx.x * 2u;
    mover_async<1050, 1056, 70, MT, DEPTH, CA>(pb + (size_t)env0 * 1056, nb + (size_t)env0 * 1056, nv, threadIdx.x, ring, R);
    mover_async<615, 640, 41, MT, DEPTH, CA>(pa + (size_t)env0 * 640, na + (size_t)env0 * 640, nv, threadIdx.x, ring, R);
}

// ---- variant 2: plain loads, U destination vectors per thread in flight (two aligned loads each, the second hits L1) ----
template <int ROW, int LD, int FRAME, int MT, int U>
__device__ __forceinline__ void mover_regs(const float *prev, float *next, int nv, uint32_t mtid) {
    constexpr uint32_t KEEP = ROW - FRAME, NV = (KEEP + 3) / 4, FVEC = FRAME >> 2, ROT = FRAME & 3;
    const uint32_t total = (uint32_t)nv * NV;                 // destination vectors of the tile
    for (uint32_t base = 0; base < total; base += MT * U) {
        float4 v[U], w[U];
        uint32_t r[U], cv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j = base + u * MT + mtid;
            r[u] = j / NV, cv[u] = j - r[u] * NV;
            if (j < total) {
                const float4 *s = reinterpret_cast<const float4 *>(prev + (size_t)r[u] * LD) + cv[u] + FVEC;
                v[u] = __ldg(s);
                if (ROT) w[u] = __ldg(s + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j = base + u * MT + mtid;
            if (j < total) {
                const float4 o = rot4<ROT>(v[u], w[u]);
                float *d = next + (size_t)r[u] * LD + 4u * cv[u];
                if (4u * cv[u] + 4u <= KEEP) st_stream4(reinterpret_cast<float4 *>(d), o);
                else { const float ov[4] = {o.x, o.y, o.z, o.w};
                       for (uint32_t k = 0; k < 4u; ++k) if (4u * cv[u] + k < KEEP) d[k] = ov[k]; }
            }
        }
    }
}
template <int MT, int U>
__global__ void __launch_bounds__(MT) k_regs(const float *pa, float *na, const float *pb, float *nb, int n) {
    extern __shared__ __align__(128) float sm[];
    const int env0 = blockIdx.x * 32, nv = min(32, n - env0);
    mover_regs<1050, 1056, 70, MT, U>(pb + (size_t)env0 * 1056, nb + (size_t)env0 * 1056, nv, threadIdx.x);
    mover_regs<615, 640, 41, MT, U>(pa + (size_t)env0 * 640, na + (size_t)env0 * 640, nv, threadIdx.x);
}


// ---- elimination: stores only / loads only, one CTA per tile ----
template <int MT>
__global__ void __launch_bounds__(MT) k_store_only(float *na, float *nb, int n) {
    extern __shared__ __align__(128) float sm[];
    const int env0 = blockIdx.x * 32, nv = min(32, n - env0);
    const float4 o = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (int r = 0; r < nv; ++r)
        for (uint32_t cv = threadIdx.x; cv < 245; cv += MT) st_stream4(reinterpret_cast<float4 *>(nb + (size_t)(env0 + r) * 1056) + cv, o);
    for (int r = 0; r < nv; ++r)
        for (uint32_t cv = threadIdx.x; cv < 143; cv += MT) st_stream4(reinterpret_cast<float4 *>(na + (size_t)(env0 + r) * 640) + cv, o);
}
template <int MT, int U>
__global__ void __launch_bounds__(MT) k_load_only(const float *pa, const float *pb, float *sink, int n) {
    extern __shared__ __align__(128) float sm[];
    const int env0 = blockIdx.x * 32, nv = min(32, n - env0);
    float acc = 0.f;
    const uint32_t total_b = nv * 246u, total_a = nv * 145u;
    for (uint32_t base = 0; base < total_b; base += MT * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j = base + u * MT + threadIdx.x, r = j / 246u, cv = j - r * 246u;
            v[u] = j < total_b ? ld_stream4(reinterpret_cast<const float4 *>(pb + (size_t)(env0 + r) * 1056) + cv + 17) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].w;
    }
    for (uint32_t base = 0; base < total_a; base += MT * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t j = base + u * MT + threadIdx.x, r = j / 145u, cv = j - r * 145u;
            v[u] = j < total_a ? ld_stream4(reinterpret_cast<const float4 *>(pa + (size_t)(env0 + r) * 640) + cv + 10) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].w;
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
}
// the same bytes with the production mapping: flat blocks of 256 threads x 4 vectors over the whole buffer
__global__ void __launch_bounds__(256) k_flat(const float *pa, float *na, const float *pb, float *nb, uint32_t va, uint32_t vb) {
    const uint32_t i0 = (blockIdx.x * 256u + threadIdx.x);
    float4 v[4];
    const uint32_t stride = gridDim.x * 256u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = i0 + u * stride;
        v[u] = i < vb ? ld_stream4(reinterpret_cast<const float4 *>(pb) + i) : (i - vb < va ? ld_stream4(reinterpret_cast<const float4 *>(pa) + (i - vb)) : make_float4(0, 0, 0, 0));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const uint32_t i = i0 + u * stride;
        if (i < vb) st_stream4(reinterpret_cast<float4 *>(nb) + i, v[u]);
        else if (i - vb < va) st_stream4(reinterpret_cast<float4 *>(na) + (i - vb), v[u]);
    }
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4096, SETS = 6;
    float *pa[SETS], *na[SETS], *pb[SETS], *nb[SETS];
    for (int s = 0; s < SETS; ++s) {
        CK(cudaMalloc(&pa[s], (size_t)n * 640 * 4)); CK(cudaMalloc(&na[s], (size_t)n * 640 * 4));
        CK(cudaMalloc(&pb[s], (size_t)n * 1056 * 4)); CK(cudaMalloc(&nb[s], (size_t)n * 1056 * 4));
        CK(cudaMemset(pa[s], 1, (size_t)n * 640 * 4)); CK(cudaMemset(pb[s], 1, (size_t)n * 1056 * 4));
    }
    const int tiles = (n + 31) / 32;
    const double bytes = 2.0 * n * (574 + 980) * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, auto launch) {
        for (int s = 0; s < SETS; ++s) launch(s);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 4 * SETS;
        for (int i = 0; i < reps; ++i) launch(i % SETS);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-58s %8.2f us per launch  %7.1f GB/s\n", name, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
        CK(cudaGetLastError());
    };
#define ASYNC(MT, DEPTH, CA, R, PAD, label)                                                                            \
    { auto k = k_async<MT, DEPTH, CA>; size_t smem = (size_t)DEPTH * R * 2 * MT * 32 + PAD;                            \
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                            \
      timeit(label, [&](int s) { k<<<tiles, MT, smem>>>(pa[s], na[s], pb[s], nb[s], n, R); }); }
#define REGS(MT, U, PAD, label)                                                                                        \
    { auto k = k_regs<MT, U>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));    \
      timeit(label, [&](int s) { k<<<tiles, MT, (size_t)PAD>>>(pa[s], na[s], pb[s], nb[s], n); }); }
    printf("n = %d envs, %d tiles (one CTA each), %.1f MB moved per launch (read + write)\n", n, tiles, bytes / 1e6);
    const int BIG = 120 * 1024;      // dynamic shared memory that forces one CTA per SM
    ASYNC(128, 8, true, 2, 0, "cp.async.ca 4 warps depth 8 x 2 rows (64 KB ring)");
    ASYNC(128, 8, false, 2, 0, "cp.async.cg 4 warps depth 8 x 2 rows");
    ASYNC(128, 8, true, 1, BIG, "cp.async.ca 4 warps depth 8 x 1 row, 1 CTA/SM");
    ASYNC(128, 4, true, 1, BIG, "cp.async.ca 4 warps depth 4 x 1 row, 1 CTA/SM");
    ASYNC(128, 2, true, 1, BIG, "cp.async.ca 4 warps depth 2 x 1 row, 1 CTA/SM");
    ASYNC(256, 8, true, 1, 0, "cp.async.ca 8 warps depth 8 x 1 row (128 KB ring)");
    ASYNC(256, 4, true, 1, BIG - 64 * 1024, "cp.async.ca 8 warps depth 4 x 1 row, 1 CTA/SM");
    ASYNC(512, 4, true, 1, 0, "cp.async.ca 16 warps depth 4 x 1 row (128 KB ring)");
    ASYNC(512, 2, true, 1, BIG - 64 * 1024, "cp.async.ca 16 warps depth 2 x 1 row, 1 CTA/SM");
    REGS(128, 4, BIG, "register loads 4 warps unroll 4, 1 CTA/SM");
    REGS(128, 8, BIG, "register loads 4 warps unroll 8, 1 CTA/SM");
    REGS(128, 16, BIG, "register loads 4 warps unroll 16, 1 CTA/SM");
    REGS(256, 8, BIG, "register loads 8 warps unroll 8, 1 CTA/SM");
    REGS(512, 8, BIG, "register loads 16 warps unroll 8, 1 CTA/SM");
    REGS(1024, 4, BIG, "register loads 32 warps unroll 4, 1 CTA/SM");
    REGS(256, 4, 0, "register loads 8 warps unroll 4, no smem (occupancy free)");
    { auto k = k_store_only<128>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      timeit("STORES ONLY 4 warps, 1 CTA/SM (half the bytes)", [&](int s) { k<<<tiles, 128, BIG>>>(na[s], nb[s], n); }); }
    { auto k = k_store_only<512>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      timeit("STORES ONLY 16 warps, 1 CTA/SM (half the bytes)", [&](int s) { k<<<tiles, 512, BIG>>>(na[s], nb[s], n); }); }
    { auto k = k_load_only<128, 8>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      timeit("LOADS ONLY 4 warps unroll 8, 1 CTA/SM (half the bytes)", [&](int s) { k<<<tiles, 128, BIG>>>(pa[s], pb[s], na[0], n); }); }
    { auto k = k_load_only<512, 8>; CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      timeit("LOADS ONLY 16 warps unroll 8, 1 CTA/SM (half the bytes)", [&](int s) { k<<<tiles, 512, BIG>>>(pa[s], pb[s], na[0], n); }); }
    { const uint32_t va = (uint32_t)n * 160, vb = (uint32_t)n * 264; const uint32_t blocks = (va + vb + 1023) / 1024;
      timeit("FLAT copy of the whole pitched buffers, 256 x 4 vectors", [&](int s) { k_flat<<<blocks, 256>>>(pa[s], na[s], pb[s], nb[s], va, vb); }); }
    return 0;
}
