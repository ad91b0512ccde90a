// Does a strided (2-D) host<->device copy of a column chunk of the reference's row-major (R, k) block run at PCIe speed?
//   nvcc -O3 -o tools/_bin/pcie_2d tools/pcie_2d.cu
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t R = 5184000, k = 16, el = 16;
    char *h, *d;
    cudaMallocHost(&h, R * k * el);
    cudaMalloc(&d, R * k * el);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); cudaMemcpyAsync(d, h, R * k * el, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("contiguous H2D %.1f GB/s\n", R * k * el / ms / 1e6);
        for (size_t cols : {8, 4, 2}) {
            cudaEventRecord(e0);
            cudaMemcpy2DAsync(d, cols * el, h, k * el, cols * el, R, cudaMemcpyHostToDevice);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("2-D H2D %zu of 16 columns (%zu B rows): %.1f GB/s\n", cols, cols * el, R * cols * el / ms / 1e6);
            cudaEventRecord(e0);
            cudaMemcpy2DAsync(h, k * el, d, cols * el, cols * el, R, cudaMemcpyDeviceToHost);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("2-D D2H %zu of 16 columns: %.1f GB/s\n", cols, R * cols * el / ms / 1e6);
        }
        // full duplex: H2D of one half while D2H of the other
        cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
        cudaEventRecord(e0);
        cudaMemcpyAsync(d, h, R * k * el / 2, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(h + R * k * el / 2, d + R * k * el / 2, R * k * el / 2, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("full duplex contiguous halves: %.1f GB/s aggregate\n", R * k * el / ms / 1e6);
    }
    return 0;
}
