#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void bench(unsigned *out, unsigned a0, unsigned b0, int iters)
{
    unsigned x0 = a0 + threadIdx.x, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, x4 = x0 * 11, x5 = x0 * 13, x6 = x0 * 17, x7 = x0 * 19;
    unsigned y = b0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {  // VABSDIFF4.ACC
                x0 = __vsadu4(x0, y) + x1; x1 = __vsadu4(x1, y) + x2; x2 = __vsadu4(x2, y) + x3; x3 = __vsadu4(x3, y) + x4;
                x4 = __vsadu4(x4, y) + x5; x5 = __vsadu4(x5, y) + x6; x6 = __vsadu4(x6, y) + x7; x7 = __vsadu4(x7, y) + x0;
            } else if (MODE == 1) {  // LOP3 baseline
                x0 = (x0 ^ y) & x1; x1 = (x1 ^ y) & x2; x2 = (x2 ^ y) & x3; x3 = (x3 ^ y) & x4;
                x4 = (x4 ^ y) & x5; x5 = (x5 ^ y) & x6; x6 = (x6 ^ y) & x7; x7 = (x7 ^ y) & x0;
            } else {  // POPC
                x0 = __popc(x0 ^ y) + x1; x1 = __popc(x1 ^ y) + x2; x2 = __popc(x2 ^ y) + x3; x3 = __popc(x3 ^ y) + x4;
                x4 = __popc(x4 ^ y) + x5; x5 = __popc(x5 ^ y) + x6; x6 = __popc(x6 ^ y) + x7; x7 = __popc(x7 ^ y) + x0;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <int MODE> float run(unsigned *d, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<148 * 4, 256>>>(d, 1, 2, 10);
    cudaEventRecord(e0); bench<MODE><<<148 * 4, 256>>>(d, 1, 0x01020304, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 2000;
    const double ops = 148.0 * 4 * 256 / 32 * iters * 64;  // warp-level ops of the measured kind (8 x 8 per iteration)
    float a = run<0>(d, iters), b = run<1>(d, iters), c = run<2>(d, iters);
    printf("VABSDIFF4(+add): %.3f ms -> %.2f warp-ops/clk/SM\n", a, ops / (a * 1e-3 * 1.965e9) / 148);
    printf("LOP3 x1 (xor-and): %.3f ms -> %.2f warp-ops/clk/SM\n", b, ops / (b * 1e-3 * 1.965e9) / 148);
    printf("POPC(+xor+add): %.3f ms -> %.2f warp-ops/clk/SM\n", c, ops / (c * 1e-3 * 1.965e9) / 148);
    return 0;
}
