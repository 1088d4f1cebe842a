// Throughput of the integer max-plus instructions on sm_100a (tuning aid; the
// numbers feed DESIGN.md's issue roofline).  Build+run: nvcc -arch=sm_100a ...
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define NACC 8
template <int OP>
__global__ void k(int *out, int a0, int b0, int c0) {
    int acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    int b = b0, c = c0;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) acc[i] = __viaddmax_s32(acc[i], b, c);              // VIADDMNMX
            if (OP == 1) acc[i] = __vimax3_s32(acc[i], b, c + i);            // VIMNMX3
            if (OP == 2) acc[i] = max(acc[i], c + it);                       // VIMNMX (+ uniform add)
            if (OP == 3) acc[i] = acc[i] + b;                                // VIADD / IADD3
            if (OP == 4) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a0), "r"(b)); // IMAD
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(c));    // PRMT
            if (OP == 6) acc[i] = __viaddmax_s16x2(acc[i], b, c);            // packed 16-bit
            if (OP == 7) acc[i] = __vimax3_s16x2(acc[i], b, c + i);
            if (OP == 8) acc[i] = __vimax_s32_relu(acc[i] , c + i);
            if (OP == 10) acc[i] = __vimax_s16x2_relu(acc[i], c + i);        // VIMNMX.S16x2.RELU
            if (OP == 11) acc[i] = __vmaxs2(acc[i], c + i);             // VIMNMX.S16x2
            if (OP == 12) acc[i] = __viaddmax_u16x2(acc[i], b, c);           // VIADDMNMX.U16x2
            if (OP == 13) acc[i] = __vimax3_s16x2_relu(acc[i], b, c + i);    // VIMNMX3.S16x2.RELU
            if (OP == 9) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a0), "r"(b)); acc[i] = max(acc[i], c); } // IMAD+VIMNMX
        }
        b ^= it;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, int ninstr) {
    int *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(int));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warps = 4; warps <= 32; warps *= 2) {   // warps per SM
        k<OP><<<148, warps * 32>>>(out, 1, 3, 5);
        cudaEventRecord(a);
        k<OP><<<148, warps * 32>>>(out, 1, 3, 5);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        int dev; cudaGetDevice(&dev);
        int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        double cycles = ms * 1e-3 * khz * 1e3;
        double instr_per_smsp = (double)ITER * NACC * ninstr * warps / 4.0;
        printf("%-22s warps/SM=%2d  cycles/warp-instr/SMSP=%.2f (at nominal %d MHz)\n", name, warps,
               cycles / instr_per_smsp, khz / 1000);
    }
    cudaFree(out);
}
int main() {
    run<0>("VIADDMNMX", 1); run<1>("VIMNMX3", 1); run<2>("VIMNMX", 1); run<3>("IADD", 1);
    run<4>("IMAD", 1); run<5>("PRMT", 1); run<6>("VIADDMNMX.16x2", 1); run<7>("VIMNMX3.16x2", 1);
    run<8>("VIMNMX.RELU", 1); run<9>("IMAD+VIMNMX pair", 2);
    run<10>("VIMNMX.S16x2.RELU", 1); run<11>("VIMNMX.S16x2", 1); run<12>("VIADDMNMX.U16x2", 1);
    run<13>("VIMNMX3.S16x2.RELU", 1);
    return 0;
}
