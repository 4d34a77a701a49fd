// rw_probe.cu — read-only, write-only and mixed streaming rates on B200 (context for the K5 tiler roofline).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) fill(float4* __restrict__ out, long long n4, float v) {
    const long long stride = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const float4 val = make_float4(v, v, v, v);
    for (; i + (U - 1) * stride < n4; i += U * stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(out + i + u * stride, val);
    }
    for (; i < n4; i += stride) __stcs(out + i, val);
}
template <int U>
__global__ void __launch_bounds__(256) rsum(const float4* __restrict__ in, long long n4, float* sink) {
    const long long stride = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    float a = 0.f;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) a += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (a == 123.456f) *sink = a;
}
// read 1 byte, write 4 bytes per element (u8 -> f32 cast), R:W = 1:4 like the tiler
template <int U>
__global__ void __launch_bounds__(256) cast_u8_f32(const unsigned int* __restrict__ in, float4* __restrict__ out, long long n4) {
    const long long stride = (long long)gridDim.x * 256;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        unsigned int w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = __ldcs(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(out + i + u * stride, make_float4((float)(w[u] & 0xff), (float)((w[u] >> 8) & 0xff), (float)((w[u] >> 16) & 0xff), (float)(w[u] >> 24)));
    }
}
int main() {
    const long long n = 1ll << 30;  // 1 Gi floats = 4 GiB out; u8 in = 1 GiB
    float* out; unsigned int* in; float* sink;
    CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, n * 4));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto bench = [&](const char* name, double bytes, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        const int iters = 10;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms / iters * 1e3, bytes / (ms / iters * 1e-3) / 1e9);
    };
    char name[96];
    const long long n4 = n / 4;
    bench("cudaMemsetAsync 4 GiB", n * 4.0, [&] { CK(cudaMemsetAsync(out, 0, n * 4)); });
    for (int per_sm : {2, 4, 8, 16, 32}) {
        snprintf(name, sizeof name, "fill U=4 grid=%dxSMs (write only)", per_sm);
        bench(name, n * 4.0, [&] { fill<4><<<sms * per_sm, 256>>>((float4*)out, n4, 1.f); });
        snprintf(name, sizeof name, "rsum U=8 grid=%dxSMs (read only)", per_sm);
        bench(name, n * 4.0, [&] { rsum<8><<<sms * per_sm, 256>>>((const float4*)in, n4, sink); });
        snprintf(name, sizeof name, "cast u8->f32 U=4 grid=%dxSMs (R:W 1:4)", per_sm);
        bench(name, n * 5.0, [&] { cast_u8_f32<4><<<sms * per_sm, 256>>>(in, (float4*)out, n4); });
        snprintf(name, sizeof name, "cast u8->f32 U=8 grid=%dxSMs (R:W 1:4)", per_sm);
        bench(name, n * 5.0, [&] { cast_u8_f32<8><<<sms * per_sm, 256>>>(in, (float4*)out, n4); });
    }
    return 0;
}
