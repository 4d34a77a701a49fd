// stream_probe.cu — standalone memory-pipeline probe for K1's access pattern (NOT part of the library).
// Copies B x C planes of HW fp32 (reads C plane-streams, writes C plane-streams, as K1 with grad does)
// with different kernel structures, to find what each structure can sustain on B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <array>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int C = 7;

// ---- 1. flat copy: grid-stride float4, U loads in flight per thread
template <int U>
__global__ void __launch_bounds__(256) flat_copy(const float4* __restrict__ in, float4* __restrict__ out, long long n4) {
    const long long stride = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(out + i + u * stride, v[u]);
    }
    for (; i < n4; i += stride) __stcs(out + i, __ldcs(in + i));
}

// ---- 2. plane walk, direct: thread owns 4 px, walks C planes (K1 direct structure), persistent
template <int PF>  // PF = 1: prefetch next item's planes before storing the current (2x regs)
__global__ void __launch_bounds__(256) planes_direct(const float* __restrict__ in, float* __restrict__ out, long long hw, int B) {
    const long long ipi = hw / 4, n_items = ipi * B;
    const long long stride = (long long)gridDim.x * 256;
    long long item = (long long)blockIdx.x * 256 + threadIdx.x;
    if (PF == 0) {
        for (; item < n_items; item += stride) {
            const long long b = item / ipi, g = item - b * ipi;
            const long long off = b * C * hw + g * 4;
            float4 v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __ldcs(reinterpret_cast<const float4*>(in + off + c * hw));
#pragma unroll
            for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(out + off + c * hw), v[c]);
        }
    } else {
        float4 cur[C], nxt[C];
        long long off = 0;
        if (item < n_items) {
            const long long b = item / ipi, g = item - b * ipi;
            off = b * C * hw + g * 4;
#pragma unroll
            for (int c = 0; c < C; ++c) cur[c] = __ldcs(reinterpret_cast<const float4*>(in + off + c * hw));
        }
        for (; item < n_items; item += stride) {
            const long long ni = item + stride;
            long long noff = 0;
            if (ni < n_items) {
                const long long b = ni / ipi, g = ni - b * ipi;
                noff = b * C * hw + g * 4;
#pragma unroll
                for (int c = 0; c < C; ++c) nxt[c] = __ldcs(reinterpret_cast<const float4*>(in + noff + c * hw));
            }
#pragma unroll
            for (int c = 0; c < C; ++c) __stcs(reinterpret_cast<float4*>(out + off + c * hw), cur[c]);
#pragma unroll
            for (int c = 0; c < C; ++c) cur[c] = nxt[c];
            off = noff;
        }
    }
}

// ---- 3. plane walk, bulk-copy pipeline (K1 TMA structure): producer lane + idle consumers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// mode 0: load + store (stage recycled after the store has read it); mode 1: load only
// touch: consumers read+write the stage (LDS.128 / STS.128) to emulate K1's shared-memory traffic
__device__ __forceinline__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// flags: 1 = also bulk-load P label bytes per stage, 2 = store a u8 "argmax" per pixel (STG), 4 = u16 private
// histogram RMW per pixel in 25 KB of shared memory, 8 = one mbarrier arrive per warp instead of per thread
__global__ void __launch_bounds__(288) planes_bulk(const float* __restrict__ in, float* __restrict__ out, long long hw, int B,
                                                  int P, int S, int mode, int touch, int lag, int flags = 0,
                                                  const unsigned char* __restrict__ lab = nullptr, unsigned char* __restrict__ amax = nullptr,
                                                  float* __restrict__ sink = nullptr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[32];
    const int tid = threadIdx.x;
    const uint32_t bar0 = smem_u32(bars), st0 = smem_u32(smem);
    const int hist_bytes = (flags & 4) ? 49 * 256 * 2 : 0;
    const uint32_t st0h = st0 + hist_bytes;
    unsigned short* hist = reinterpret_cast<unsigned short*>(smem);
    if (flags & 4) for (int i = tid; i < 49 * 256; i += 288) hist[i] = 0;
    const int stage_bytes = C * P * 4 + ((flags & 1) ? P : 0);
    const int done_count = (flags & 8) ? 8 : 256;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (16 + s), done_count); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long cpi = hw / P, n_chunks = cpi * B;
    const long long mine = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto off_of = [&](long long i) { const long long q = blockIdx.x + i * gridDim.x; const long long b = q / cpi, k = q - b * cpi; return b * C * hw + k * P; };
    if (tid >= 256) {
        if (tid == 256) {
            auto load = [&](long long i) {
                const int s = (int)(i % S); const long long off = off_of(i);
                mbar_expect_tx(bar0 + 8 * s, stage_bytes);
                for (int c = 0; c < C; ++c) bulk_g2s(st0h + s * stage_bytes + c * P * 4, in + off + c * hw, P * 4, bar0 + 8 * s);
                if (flags & 1) { const long long q = blockIdx.x + i * gridDim.x; bulk_g2s(st0h + s * stage_bytes + C * P * 4, lab + q * P, P, bar0 + 8 * s); }
            };
            const long long pre = mine < S ? mine : S;
            for (long long i = 0; i < pre; ++i) load(i);
            for (long long i = 0; i < mine; ++i) {
                const int s = (int)(i % S);
                mbar_wait(bar0 + 8 * (16 + s), (uint32_t)((i / S) & 1));
                if (mode == 0) {
                    const long long off = off_of(i);
                    for (int c = 0; c < C; ++c) bulk_s2g(out + off + c * hw, st0h + s * stage_bytes + c * P * 4, P * 4);
                    bulk_commit();
                    // stage (i - lag) is free once its store has left smem
                    if (i >= lag && i - lag + S < mine) {
                        if (lag == 1) bulk_wait_read<1>(); else if (lag == 2) bulk_wait_read<2>(); else bulk_wait_read<0>();
                        load(i - lag + S);
                    }
                } else if (i + S < mine) load(i + S);
            }
            if (mode == 0) bulk_wait_all();
        }
    } else {
        for (long long i = 0; i < mine; ++i) {
            const int s = (int)(i % S);
            mbar_wait(bar0 + 8 * s, (uint32_t)((i / S) & 1));
            unsigned char* stg = smem + hist_bytes + s * stage_bytes;
            if (touch == 1) {
                float4* st = reinterpret_cast<float4*>(stg);
                for (int j = tid; j < C * P / 4; j += 256) { float4 v = st[j]; v.x += 1.f; st[j] = v; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            } else if (touch == 2) {
                // K1-like: thread owns VP = P/256 consecutive pixels x C planes
                const int VP = P / 256;
                float acc = 0.f;
                for (int k0 = 0; k0 < VP; k0 += 2) {   // two pixels at a time (LDS.64)
                    float x[2][C];
#pragma unroll
                    for (int c = 0; c < C; ++c) { const float2 v = *reinterpret_cast<const float2*>(stg + (c * P + tid * VP + k0) * 4); x[0][c] = v.x; x[1][c] = v.y; }
                    unsigned int am = 0;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float m = x[k][0];
#pragma unroll
                        for (int c = 1; c < C; ++c) m = fmaxf(m, x[k][c]);
                        int arg = C - 1;
#pragma unroll
                        for (int c = C - 2; c >= 0; --c) arg = (x[k][c] == m) ? c : arg;
                        float sum = 0.f;
#pragma unroll
                        for (int c = 0; c < C; ++c) { x[k][c] = ex2f_((x[k][c] - m) * 1.4426950408889634f); sum += x[k][c]; }
                        acc += lg2f_(sum) + m;
                        const float r = rcpf_(sum) * 1e-7f;
#pragma unroll
                        for (int c = 0; c < C; ++c) x[k][c] *= r;
                        am |= (unsigned)arg << (8 * k);
                        if (flags & 4) { unsigned short* h = hist + (arg * 7 + (arg ^ 1) % 7) * 256 + tid; *h = (unsigned short)(*h + 1); }
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) *reinterpret_cast<float2*>(stg + (c * P + tid * VP + k0) * 4) = make_float2(x[0][c], x[1][c]);
                    if (flags & 2) { const long long q = blockIdx.x + i * gridDim.x; *reinterpret_cast<unsigned short*>(amax + q * P + tid * VP + k0) = (unsigned short)am; }
                }
                if (acc == 123.456f) sink[0] = acc;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (flags & 8) { __syncwarp(); if ((tid & 31) == 0) mbar_arrive(bar0 + 8 * (16 + s)); continue; }
            mbar_arrive(bar0 + 8 * (16 + s));
        }
    }
}

int main(int argc, char** argv) {
    const int B = 16; const long long hw = 1024 * 1024;
    const long long n = (long long)B * C * hw;
    const int NS = 3;
    float *in[NS], *out[NS];
    for (int i = 0; i < NS; ++i) { CK(cudaMalloc(&in[i], n * 4)); CK(cudaMalloc(&out[i], n * 4)); CK(cudaMemset(in[i], 1, n * 4)); CK(cudaMemset(out[i], 0, n * 4)); }
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto bench = [&](const char* name, double bytes, auto launch) {
        for (int i = 0; i < 5; ++i) launch(i % NS);
        CK(cudaDeviceSynchronize());
        const int iters = 50;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) launch(i % NS);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
        printf("%-58s %8.1f us  %7.1f GB/s\n", name, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9);
    };
    const double rw = 2.0 * n * 4, ro = 1.0 * n * 4;
    char name[128];
    for (int per_sm : {4, 8, 16}) {
        snprintf(name, sizeof name, "flat_copy U=4 grid=%dxSMs", per_sm);
        bench(name, rw, [&](int k) { flat_copy<4><<<sms * per_sm, 256>>>((const float4*)in[k], (float4*)out[k], n / 4); });
        snprintf(name, sizeof name, "flat_copy U=8 grid=%dxSMs", per_sm);
        bench(name, rw, [&](int k) { flat_copy<8><<<sms * per_sm, 256>>>((const float4*)in[k], (float4*)out[k], n / 4); });
    }
    bench("cudaMemcpyAsync D2D", rw, [&](int k) { CK(cudaMemcpyAsync(out[k], in[k], n * 4, cudaMemcpyDeviceToDevice)); });
    for (int per_sm : {2, 3, 4, 6, 8}) {
        snprintf(name, sizeof name, "planes_direct PF=0 grid=%dxSMs", per_sm);
        bench(name, rw, [&](int k) { planes_direct<0><<<sms * per_sm, 256>>>(in[k], out[k], hw, B); });
        snprintf(name, sizeof name, "planes_direct PF=1 grid=%dxSMs", per_sm);
        bench(name, rw, [&](int k) { planes_direct<1><<<sms * per_sm, 256>>>(in[k], out[k], hw, B); });
    }
    CK(cudaFuncSetAttribute(planes_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    unsigned char *lab, *amax; float* sink;
    CK(cudaMalloc(&lab, B * hw)); CK(cudaMalloc(&amax, B * hw)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(lab, 1, B * hw));
    struct Cfg { int per_sm, P, S, touch, lag, flags; };
    std::vector<Cfg> cfgs;
    for (int rep = 0; rep < 2; ++rep)
        for (auto geo : std::vector<std::array<int, 3>>{{2, 512, 4}, {2, 512, 5}, {2, 1024, 3}, {3, 512, 3}})
            for (int touch : {1, 2})
                for (int flags : {0, 1, 2, 3, 4, 7, 8, 15}) {
                    if (touch == 1 && (flags & ~9)) continue;
                    cfgs.push_back({geo[0], geo[1], geo[2], touch, 1, flags});
                }
    for (auto c : cfgs) {
        const int smem = c.S * (C * c.P * 4 + ((c.flags & 1) ? c.P : 0)) + ((c.flags & 4) ? 49 * 512 : 0);
        snprintf(name, sizeof name, "bulk ctas/SM=%d P=%d S=%d touch=%d flags=%2d (%d KB)", c.per_sm, c.P, c.S, c.touch, c.flags, smem / 1024);
        bench(name, rw, [&](int k) { planes_bulk<<<sms * c.per_sm, 288, smem>>>(in[k], out[k], hw, B, c.P, c.S, 0, c.touch, c.lag, c.flags, lab, amax, sink); });
    }
    return 0;
}
