// Pipe-bound rate of the affine cell's instruction mix with NO dependency chain
// between rows (tuning aid): PRMT + 3 VIADDMNMX + VIMNMX.RELU + IMAD + 0.5 VIMNMX3.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
#define ROWS 16
__device__ __forceinline__ int prmt_sx(unsigned lo, unsigned hi, unsigned sel) {
    int d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel)); return d;
}
__device__ __forceinline__ int imad(int m, int one, int open) {
    int d; asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(m), "r"(one), "r"(open)); return d;
}
template <int MODE>
__global__ void k(int *out, int ext, int open, int one, unsigned x0, unsigned x1) {
    int M[ROWS], D[ROWS], I[ROWS]; unsigned sel[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { M[r] = threadIdx.x + r; D[r] = r; I[r] = -r; sel[r] = (r & 3) * 0x1111u | 0x8880u; }
    int cm = 0;
    for (int it = 0; it < ITER; ++it) {
        unsigned X0 = x0 + it, X1 = x1 ^ it;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            int sc = prmt_sx(X0, X1, sel[r]);
            D[r] = __viaddmax_s32(D[r], ext, M[r]);
            int X = __viaddmax_s32(M[(r + ROWS - 1) % ROWS], sc, D[r]);
            I[r] = __viaddmax_s32(I[r], ext, M[r]);           // independent per row (no vertical chain)
            int Mv = __vimax_s32_relu(X, I[r]);
            if (MODE == 0) M[r] = imad(Mv, one, open); else M[r] = Mv + open;
            if (r & 1) cm = __vimax3_s32(cm, M[r], M[r - 1]);
            if (MODE >= 2) {}
        }
    }
    int s = cm;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) s += M[r] + D[r] + I[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed: two lattices per register (s16x2); G add as viaddmax (MODE 2) or IMAD on biased halves (MODE 3)
template <int MODE>
__global__ void kp(int *out, int ext, int open, int one, unsigned x0, unsigned x1) {
    unsigned M[ROWS], D[ROWS], I[ROWS]; unsigned sel[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { M[r] = threadIdx.x + r; D[r] = r; I[r] = r * 3; sel[r] = ((r & 3) | ((r & 3) | 8) << 4 | (4 + (r & 3)) << 8 | ((4 + (r & 3)) | 8) << 12); }
    unsigned cm = 0;
    const unsigned ext2 = (unsigned)ext, open2 = (unsigned)open, neg = 0x80008000u;
    for (int it = 0; it < ITER; ++it) {
        unsigned X0 = x0 + it, X1 = x1 ^ it;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            unsigned sc; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(X0), "r"(X1), "r"(sel[r]));
            D[r] = __viaddmax_s16x2(D[r], ext2, M[r]);
            unsigned X = __viaddmax_s16x2(M[(r + ROWS - 1) % ROWS], sc, D[r]);
            I[r] = __viaddmax_s16x2(I[r], ext2, M[r]);
            unsigned Mv = __vimax_s16x2_relu(X, I[r]);
            if (MODE == 2) M[r] = __viaddmax_s16x2(Mv, open2, neg);
            else M[r] = (unsigned)imad((int)Mv, one, (int)open2);
            if (r & 1) cm = __vimax3_s16x2(cm, M[r], M[r - 1]);
        }
    }
    unsigned s = cm;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) s += M[r] + D[r] + I[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void runp(const char *name) {
    int *out; cudaMalloc(&out, 148 * 1024 * sizeof(int));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warps = 4; warps <= 16; warps *= 2) {
        kp<MODE><<<148, warps * 32>>>(out, 0xfffcfffc, 0xfff4fff4, 1, 0x05fcfcfc, 0xfcfc05fc);
        cudaEventRecord(a);
        kp<MODE><<<148, warps * 32>>>(out, 0xfffcfffc, 0xfff4fff4, 1, 0x05fcfcfc, 0xfcfc05fc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        double cycles = ms * 1e-3 * khz * 1e3;
        double rows_per_smsp = (double)ITER * ROWS * warps / 4.0;
        printf("%-10s warps/SM=%2d cycles per packed warp-row per SMSP = %.2f  -> %.0f GCUPS chip-wide (2 cells/row)\n", name, warps,
               cycles / rows_per_smsp, 2 * 148 * 4 * 32 / (cycles / rows_per_smsp) * khz * 1e3 / 1e9);
    }
}
template <int MODE> void run(const char *name) {
    int *out; cudaMalloc(&out, 148 * 1024 * sizeof(int));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int warps = 4; warps <= 16; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, -4, -12, 1, 0x05fcfcfc, 0xfcfc05fc);
        cudaEventRecord(a);
        k<MODE><<<148, warps * 32>>>(out, -4, -12, 1, 0x05fcfcfc, 0xfcfc05fc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        double cycles = ms * 1e-3 * khz * 1e3;
        double rows_per_smsp = (double)ITER * ROWS * warps / 4.0;
        printf("%-10s warps/SM=%2d cycles per warp-row per SMSP = %.2f  -> %.0f GCUPS chip-wide\n", name, warps,
               cycles / rows_per_smsp, 148 * 4 * 32 / (cycles / rows_per_smsp) * khz * 1e3 / 1e9);
    }
}
int main() { run<0>("imad"); run<1>("iadd"); runp<2>("p16 dpxadd"); runp<3>("p16 imad"); return 0; }
