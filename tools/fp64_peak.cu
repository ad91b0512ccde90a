// Micro-benchmark: sustained FP64 throughput of one B200 -- DFMA (vector pipe) and DMMA (mma.sync.m8n8k4.f64) --
// and the shared-memory LDS.128 rate, to size the Gram/update/FFT kernels against real ceilings.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// larger FP64 MMA shapes (PTX: sm_90+): m16n8k8 = 1024 FMA per warp instruction, A 4 regs, B 2 regs, C 4 regs per lane.
// On sm_100a ptxas lowers it to a sequence of DMMA.8x8x4 (cuobjdump -sass shows no other DMMA shape), so it brings no
// operand-traffic advantage over issuing m8n8k4 directly; kept here as the evidence.
__global__ void k_dmma16(double* out, int iters) {
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-6, b1 = b0 + 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds(double* out, int iters) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, 1.0);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double2 v = sm[(idx + i * 256) & 2047];
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + 1) & 2047;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {256, 512, 1024}) {
        const int blocks = sms * (2048 / threads), iters = 20000;
        k_dfma<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  threads/CTA %4d: %.2f TFLOP/s\n", threads, 2.0 * 16 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
        k_dmma<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  threads/CTA %4d: %.2f TFLOP/s\n", threads, 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
        k_dmma16<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma16<<<blocks, threads>>>(out, iters / 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA m16n8k8 threads/CTA %4d: %.2f TFLOP/s\n", threads, 2.0 * 1024 * 4 * (iters / 4) * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12);
    }
    {
        const int threads = 512, blocks = sms * 4, iters = 20000;
        k_lds<<<blocks, threads, 32768>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_lds<<<blocks, threads, 32768>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("LDS.128: %.1f B/clk/SM at %d MHz (%.2f TB/s)\n", 16.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / sms / (p.clockRate * 1e3),
               p.clockRate / 1000, 16.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    printf("SMs %d clock %d MHz\n", sms, p.clockRate / 1000);
    return 0;
}
