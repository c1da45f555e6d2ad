// dependent-chain latency micro-benchmarks (1 warp, 1 block)
#include <cstdio>
#include <cuda_runtime.h>
#include <cmath>
template <int OP>
__global__ void k(double* out, int iters, double a, double b, double c) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) x = fma(x, b, c);
        if (OP == 1) x = x + c;
        if (OP == 2) x = x * b;
        if (OP == 3) x = c / x + 1.0;           // division + add
        if (OP == 4) x = log(x) + 2.0;          // log + add
        if (OP == 5) x = exp(x * 1e-3) ;        // mul + exp
        if (OP == 6) x = pow(x, 1.0000001) + 1e-9;
        if (OP == 7) x = sqrt(x) + 1.0;
        if (OP == 8) { double s, co; sincos(x, &s, &co); x = s + co + 1.0; }
        if (OP == 9) x = atan2(x, 2.0) + 1.0;
        if (OP == 10) x = asin(x * 1e-3) + 1.0;
        if (OP == 11) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 12) x = (x < b) ? x + c : x - c;  // DSETP + select + add
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / iters; }
}
int main() {
    double* d; cudaMalloc(&d, 64);
    const char* names[] = {"DFMA","DADD","DMUL","div+add","log+add","mul+exp","pow+add","sqrt+add","sincos+2add","atan2+add","asin+mul+add","shfl64","dsetp+sel+add"};
    double h[2];
    #define RUN(OP) k<OP><<<1,32>>>(d, 4096, 1.5, 1.0000001, 1e-9); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-14s %8.1f cycles/iter (x=%g)\n", names[OP], h[1], h[0]);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    // warm second pass
    RUN(0) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
