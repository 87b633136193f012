// Latency microbenchmarks for the FPS critical path on sm_100a: redux, vote, flo, shfl, LDS, barrier.
#include <cstdio>
#include <cuda_runtime.h>
#define REP 256
__global__ void k_lat(unsigned* out, long long* cyc, int mode) {
    __shared__ unsigned s[1024];
    unsigned v = threadIdx.x * 2654435761u + out[0];
    s[threadIdx.x] = v & 31;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {  // redux max chain
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = __reduce_max_sync(0xffffffffu, v) + (threadIdx.x & 1);
    } else if (mode == 1) {  // ballot chain
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = __ballot_sync(0xffffffffu, v & 1) + threadIdx.x;
    } else if (mode == 2) {  // ffs chain
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = __ffs(v | 0x80000000u) + v;
    } else if (mode == 3) {  // shfl chain
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    } else if (mode == 4) {  // LDS dependent chain
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = s[v & 1023];
    } else if (mode == 5) {  // barrier chain
        for (int i = 0; i < REP; ++i) { __syncthreads(); }
    } else if (mode == 6) {  // STS -> barrier -> LDS round trip
        for (int i = 0; i < REP; ++i) { s[(threadIdx.x + 32) & 1023] = v; __syncthreads(); v = s[threadIdx.x] + 1; }
    } else if (mode == 7) {  // FMNMX/FSETP/FSEL dependent chain
        float f = __uint_as_float(v & 0x3fffffff), g = 1.0f;
#pragma unroll 16
        for (int i = 0; i < REP; ++i) { f = fminf(f * 1.0001f, g); if (f > g) g = f; }
        v = __float_as_uint(f + g);
    } else if (mode == 8) {  // redux+ballot+ffs+LDS.128 combo (v2 final stage)
        float4* s4 = reinterpret_cast<float4*>(s);
#pragma unroll 4
        for (int i = 0; i < REP; ++i) {
            unsigned m = __reduce_max_sync(0xffffffffu, v);
            unsigned b = __ballot_sync(0xffffffffu, v == m);
            int w = __ffs(b) - 1;
            float4 c = s4[w & 31];
            v = __float_as_uint(c.x) + threadIdx.x + i;
        }
    } else if (mode == 9) {  // popc
#pragma unroll 16
        for (int i = 0; i < REP; ++i) v = __popc(v) + v;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[1 + threadIdx.x] = v;
}
int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64 * 8);
    cudaMemset(out, 0, 4096 * 4);
    const char* names[] = {"redux.max", "ballot", "ffs", "shfl", "LDS dep", "bar.sync", "STS+bar+LDS", "fmnmx/fsetp/fsel x2", "redux+ballot+ffs+LDS128", "popc"};
    for (int threads : {32, 256, 512, 1024}) {
        for (int mode = 0; mode < 10; ++mode) {
            k_lat<<<1, threads>>>(out, cyc, mode);
            k_lat<<<1, threads>>>(out, cyc, mode);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads %4d  %-26s %7.1f cycles/op\n", threads, names[mode], (double)c / REP);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
