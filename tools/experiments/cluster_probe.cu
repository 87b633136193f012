// Which thread-block cluster sizes (including non-powers of two) can B200 keep co-resident, and where do the
// CTAs of such clusters land?  nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void __launch_bounds__(1024, 1) probe_kernel(int* smid, int* rank, int spin) {
    extern __shared__ float s[];
    unsigned sm, r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    if (threadIdx.x == 0) {
        smid[blockIdx.x] = (int)sm;
        rank[blockIdx.x] = (int)r;
        s[0] = 1.0f;
        long long t0 = clock64();
        while (clock64() - t0 < spin) {}
    }
}

int main() {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int threads[] = {256, 512, 768, 1024};
    const int smem_kb[] = {48, 96, 116, 200, 226};
    printf("cluster threads smem_KB max_active_clusters ctas\n");
    for (int C = 2; C <= 16; ++C)
        for (int t : threads)
            for (int kb : smem_kb) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(C * 148, 1, 1);
                cfg.blockDim = dim3(t, 1, 1);
                cfg.dynamicSmemBytes = (size_t)kb * 1024;
                cudaLaunchAttribute a[1];
                a[0].id = cudaLaunchAttributeClusterDimension;
                a[0].val.clusterDim.x = C;
                a[0].val.clusterDim.y = 1;
                a[0].val.clusterDim.z = 1;
                cfg.attrs = a;
                cfg.numAttrs = 1;
                int num = -1;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&num, probe_kernel, &cfg);
                if (e != cudaSuccess) {
                    printf("%2d %4d %3d  error %s\n", C, t, kb, cudaGetErrorString(e));
                    cudaGetLastError();
                } else
                    printf("%2d %4d %3d %4d %4d\n", C, t, kb, num, num * C);
            }
    // actual placement: 8 clusters of C CTAs at 226 KB each (1 CTA per SM), CTAs held for ~1 ms
    int *d_sm, *d_rank;
    cudaMalloc(&d_sm, 4096 * 4);
    cudaMalloc(&d_rank, 4096 * 4);
    for (int C : {16, 15, 14, 13, 12, 10}) {
        const int nclu = 8, grid = C * nclu;
        cudaMemset(d_sm, 0xff, 4096 * 4);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(512, 1, 1);
        cfg.dynamicSmemBytes = 226 * 1024;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = C;
        a[0].val.clusterDim.y = 1;
        a[0].val.clusterDim.z = 1;
        cfg.attrs = a;
        cfg.numAttrs = 1;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d_sm, d_rank, 2000000);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<int> sm(grid), rk(grid);
        cudaMemcpy(sm.data(), d_sm, grid * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(rk.data(), d_rank, grid * 4, cudaMemcpyDeviceToHost);
        printf("launch C=%d x %d clusters: %s / %s, %.3f ms (one wave ~1.0 ms at 1.9 GHz)\n", C, nclu, cudaGetErrorString(e), cudaGetErrorString(e2), ms);
        for (int c = 0; c < nclu; ++c) {
            printf("  cluster %d SMs:", c);
            for (int r = 0; r < C; ++r) printf(" %d", sm[c * C + r]);
            printf("  ranks ok=%d\n", rk[c * C] == 0 && rk[c * C + C - 1] == C - 1);
        }
    }
    return 0;
}
