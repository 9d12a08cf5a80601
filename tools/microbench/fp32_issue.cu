// FP32 issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2 (f32x2), 3 distinct
// register operands, 8 independent chains per thread.  Prints warp-instructions per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp32_issue tools/microbench/fp32_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, float a0, float b0, int iters) {
    float2 x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = make_float2(a0 + i, a0 - i);
        y[i] = make_float2(b0 * (i + 1), b0 / (i + 1));
        z[i] = make_float2(threadIdx.x * 1e-3f, i);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {  // scalar FFMA, 3 distinct registers
                z[i].x = fmaf(x[i].x, y[i].x, z[i].x);
                z[i].y = fmaf(x[i].y, y[i].y, z[i].y);
            } else if (MODE == 1) {  // packed FFMA2
                z[i] = __ffma2_rn(x[i], y[i], z[i]);
            } else if (MODE == 2) {  // scalar FADD, 2 registers
                z[i].x = z[i].x + x[i].x;
                z[i].y = z[i].y + x[i].y;
            } else if (MODE == 3) {  // packed FADD2
                z[i] = __fadd2_rn(z[i], x[i]);
            } else if (MODE == 4) {  // scalar FFMA with a shared multiplier (operand reuse)
                z[i].x = fmaf(x[0].x, y[i].x, z[i].x);
                z[i].y = fmaf(x[0].x, y[i].y, z[i].y);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += z[i].x + z[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int instrPerIter, int flopsPerInstr) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int threads = 256, blocks = sms * 8, iters = 4096;
    float *out;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    k<MODE><<<blocks, threads>>>(out, 1.0f, 0.999f, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0f, 0.999f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double warpInstr = (double)blocks * threads / 32 * iters * instrPerIter;
    double perSec = warpInstr / (ms * 1e-3);
    printf("%-34s %8.3f ms  %.2f warp-instr/clk/SM at %d MHz nominal  (%.1f TFLOP/s)\n", name, ms,
           perSec / sms / (khz * 1e3), khz / 1000, perSec * 32 * flopsPerInstr / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("scalar FFMA (3 regs)", 16, 2);
    run<1>("packed FFMA2 (3 reg pairs)", 8, 4);
    run<2>("scalar FADD (2 regs)", 16, 1);
    run<3>("packed FADD2", 8, 2);
    run<4>("scalar FFMA (shared multiplier)", 16, 2);
    return 0;
}
