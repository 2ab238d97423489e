// Microbenchmark: issue rate of FFMA (3-register), FFMA2 (packed f32x2), FSEL/FSETP and mixes on one SM quadrant.
// nvcc -gencode arch=compute_100a,code=sm_100a -o fp32_rates fp32_rates.cu && ./fp32_rates
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
template <int MODE> __global__ void k(float *out, float a, float b, long long *cyc) {
    float x[8]; u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a + i + threadIdx.x; float2 t = make_float2(x[i], x[i] + 1.f); p[i] = *reinterpret_cast<u64 *>(&t); }
    float2 bb = make_float2(b, b * 1.0001f), aa = make_float2(a, a * 0.999f);
    u64 pb = *reinterpret_cast<u64 *>(&bb), pa = *reinterpret_cast<u64 *>(&aa);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                                                   // FFMA
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));   // FFMA2
            if (MODE == 2) { x[i] = fmaf(x[i], a, b); x[i] = x[i] > 0.5f ? x[i] : a; }                  // FFMA + FSETP/FSEL
            if (MODE == 3) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); x[i] = x[i] > b ? x[i] : x[(i + 1) & 7]; }
            if (MODE == 4) x[i] = x[i] > b ? x[i] : x[(i + 1) & 7] + 0.f;                              // selects only
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2 *>(&p[i]); s += x[i] + t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, int ops_per_iter) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, 1.0001f, 0.5f, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SM %2d  cycles/iter %.1f  warp-instr/clk/SMSP %.3f\n", name, warps, (double) h / ITERS, (double) ops_per_iter * (warps / 4.0) / ((double) h / ITERS));
    }
}
int main() {
    run<0>("FFMA x8", 8);
    run<1>("FFMA2 x8 (16 fma)", 8);
    run<2>("FFMA+FSETP+FSEL x8", 24);
    run<3>("FFMA2+FSETP+FSEL x8", 24);
    run<4>("FSETP+FSEL(+FADD) x8", 24);
    return 0;
}
