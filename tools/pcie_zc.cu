// Host <-> device paths for a COLUMN CHUNK of the reference's row-major (R, k) host block (pcb_apply_host):
//   (a) cudaMemcpy2DAsync with narrow rows, one stream vs the rows split over several streams (several copy engines?)
//   (b) a kernel reading pinned host memory directly (zero copy) and writing the planar device columns, and the reverse
//   (c) (b) in both directions at once on two streams
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/pcie_zc tools/pcie_zc.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 cplx;
// host block H[R][k] (row-major) columns [j0, j0+w) -> device planar D[j][R]
__global__ void k_gather(const cplx* __restrict__ H, cplx* __restrict__ D, long long R, int k, int j0, int w) {
    const long long n = R * w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / w; const int j = (int)(e % w);
        D[(long long)j * R + r] = H[r * k + j0 + j];
    }
}
__global__ void k_scatter(cplx* __restrict__ H, const cplx* __restrict__ D, long long R, int k, int j0, int w) {
    const long long n = R * w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / w; const int j = (int)(e % w);
        H[r * k + j0 + j] = D[(long long)j * R + r];
    }
}
int main() {
    const long long R = 5184000; const int k = 16; const size_t el = 16;
    cplx *h, *d, *d2;
    cudaHostAlloc(&h, R * k * el, cudaHostAllocMapped);
    cudaMalloc(&d, R * k * el); cudaMalloc(&d2, R * k * el);
    cudaMemset(d2, 0, R * k * el);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaStream_t st[8]; for (int i = 0; i < 8; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    float ms; int ne = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); ne = p.asyncEngineCount;
    printf("asyncEngineCount %d\n", ne);
    auto join = [&](int other) { cudaEvent_t t; cudaEventCreateWithFlags(&t, cudaEventDisableTiming); cudaEventRecord(t, st[other]); cudaStreamWaitEvent(st[0], t, 0); cudaEventDestroy(t); };
    for (int rep = 0; rep < 2; ++rep) {
        // both directions at once, 8 columns each way (the pipeline's steady state), by different engines
        for (int variant = 0; variant < 6; ++variant) {
            cudaDeviceSynchronize(); cudaEventRecord(e0, st[0]); cudaStreamWaitEvent(st[1], e0, 0);
            const char* name = "";
            switch (variant) {
                case 0: name = "DMA contiguous H2D + DMA contiguous D2H (halves)";
                    cudaMemcpyAsync(d, h, R * 8 * el, cudaMemcpyHostToDevice, st[0]);
                    cudaMemcpyAsync(h + R * 8, d2, R * 8 * el, cudaMemcpyDeviceToHost, st[1]); break;
                case 1: name = "DMA 2-D H2D + DMA 2-D D2H";
                    cudaMemcpy2DAsync(d, 8 * el, h, k * el, 8 * el, R, cudaMemcpyHostToDevice, st[0]);
                    cudaMemcpy2DAsync(h + 8, k * el, d2, 8 * el, 8 * el, R, cudaMemcpyDeviceToHost, st[1]); break;
                case 2: name = "DMA 2-D H2D + zero-copy scatter (296 blocks)";
                    cudaMemcpy2DAsync(d, 8 * el, h, k * el, 8 * el, R, cudaMemcpyHostToDevice, st[0]);
                    k_scatter<<<296, 256, 0, st[1]>>>(h, d2, R, k, 8, 8); break;
                case 3: name = "zero-copy gather (296 blocks) + DMA 2-D D2H";
                    k_gather<<<296, 256, 0, st[0]>>>(h, d, R, k, 0, 8);
                    cudaMemcpy2DAsync(h + 8, k * el, d2, 8 * el, 8 * el, R, cudaMemcpyDeviceToHost, st[1]); break;
                case 4: name = "zero-copy gather + zero-copy scatter (296 blocks each)";
                    k_gather<<<296, 256, 0, st[0]>>>(h, d, R, k, 0, 8);
                    k_scatter<<<296, 256, 0, st[1]>>>(h, d2, R, k, 8, 8); break;
                case 5: name = "zero-copy gather (592 blocks) + zero-copy scatter (148 blocks)";
                    k_gather<<<592, 256, 0, st[0]>>>(h, d, R, k, 0, 8);
                    k_scatter<<<148, 256, 0, st[1]>>>(h, d2, R, k, 8, 8); break;
            }
            join(1);
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("%-62s %.1f GB/s aggregate (%.1f ms)\n", name, 2.0 * R * 8 * el / ms / 1e6, ms);
        }
        // the same with 4 columns each way
        for (int variant = 0; variant < 3; ++variant) {
            cudaDeviceSynchronize(); cudaEventRecord(e0, st[0]); cudaStreamWaitEvent(st[1], e0, 0);
            const char* name = "";
            switch (variant) {
                case 0: name = "4 cols: DMA 2-D H2D + DMA 2-D D2H";
                    cudaMemcpy2DAsync(d, 4 * el, h, k * el, 4 * el, R, cudaMemcpyHostToDevice, st[0]);
                    cudaMemcpy2DAsync(h + 8, k * el, d2, 4 * el, 4 * el, R, cudaMemcpyDeviceToHost, st[1]); break;
                case 1: name = "4 cols: DMA 2-D H2D + zero-copy scatter (296 blocks)";
                    cudaMemcpy2DAsync(d, 4 * el, h, k * el, 4 * el, R, cudaMemcpyHostToDevice, st[0]);
                    k_scatter<<<296, 256, 0, st[1]>>>(h, d2, R, k, 8, 4); break;
                case 2: name = "4 cols: zero-copy gather + zero-copy scatter (296 blocks each)";
                    k_gather<<<296, 256, 0, st[0]>>>(h, d, R, k, 0, 4);
                    k_scatter<<<296, 256, 0, st[1]>>>(h, d2, R, k, 8, 4); break;
            }
            join(1);
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("%-62s %.1f GB/s aggregate (%.1f ms)\n", name, 2.0 * R * 4 * el / ms / 1e6, ms);
        }
        // single direction references
        for (int w : {8, 4}) {
            cudaDeviceSynchronize(); cudaEventRecord(e0, st[0]);
            k_gather<<<296, 256, 0, st[0]>>>(h, d, R, k, 0, w);
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("zero-copy gather  %d cols, 296 blocks: %.1f GB/s\n", w, R * w * el / ms / 1e6);
            cudaEventRecord(e0, st[0]);
            k_scatter<<<296, 256, 0, st[0]>>>(h, d2, R, k, 0, w);
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("zero-copy scatter %d cols, 296 blocks: %.1f GB/s\n", w, R * w * el / ms / 1e6);
            cudaEventRecord(e0, st[0]);
            cudaMemcpy2DAsync(h + 8, k * el, d2, w * el, w * el, R, cudaMemcpyDeviceToHost, st[0]);
            cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("DMA 2-D D2H %d cols: %.1f GB/s\n", w, R * w * el / ms / 1e6);
        }
    }
    return 0;
}
