#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: 8 VABSDIFF4 ; 1: 8 LOP3 ; 2: 8 VABSDIFF4 + 8 LOP3 ; 3: 8 IMAD ; 4: 8 VABSDIFF4 + 8 IMAD ; 5: 8 LOP3 + 8 IMAD
template <int MODE>
__global__ void bench(unsigned *out, unsigned a0, unsigned b0, int iters)
{
    unsigned x[8], z[8];
    for (int i = 0; i < 8; i++) { x[i] = a0 * (2 * i + 3) + threadIdx.x; z[i] = a0 * (2 * i + 5) + threadIdx.x * 3; }
    unsigned y = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0 || MODE == 2 || MODE == 4) x[i] = __vsadu4(x[i], y) + x[(i + 1) & 7];
                if (MODE == 1 || MODE == 2 || MODE == 5) z[i] = (z[i] ^ y) & z[(i + 3) & 7];
                if (MODE == 3) x[i] = x[i] * y + x[(i + 1) & 7];
                if (MODE == 4 || MODE == 5) { if (MODE == 4) z[i] = z[i] * y + z[(i + 3) & 7]; else x[i] = x[i] * y + x[(i + 1) & 7]; }
            }
        }
    }
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s += x[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(unsigned *d, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<148 * 4, 256>>>(d, 1, 2, 10);
    cudaEventRecord(e0); bench<MODE><<<148 * 4, 256>>>(d, 7, 0x01020305, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 2000;
    const char *names[6] = {"8 VABSDIFF4", "8 LOP3", "8 VABSDIFF4 + 8 LOP3", "8 IMAD", "8 VABSDIFF4 + 8 IMAD", "8 LOP3 + 8 IMAD"};
    float ms[6] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters)};
    const double groups = 148.0 * 4 * 256 / 32 * iters * 4;  // warp-level groups of 8 (or 8+8) per SM-set
    for (int i = 0; i < 6; i++) printf("%-24s %.3f ms  -> %.2f clk per group of 8(+8) per SMSP\n", names[i], ms[i], ms[i] * 1e-3 * 1.965e9 / (groups / (148 * 4)));
    return 0;
}
