// Microbenchmark: issue throughput of the instructions K1/K2 are built from, on
// one SM-full of warps (sm_100a).  Prints warp-instructions / clk / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 2048
#define NACC 8

typedef unsigned long long u64;

template <int OP>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cycles, float seed) {
    float a[NACC], b = seed * 1.0001f, c = seed * 0.5f;
    u64 p[NACC];
    uint32_t h[NACC];
    uint32_t hb = __float_as_uint(seed) | 0x3f803f80u;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        a[k] = seed + k + threadIdx.x;
        float2 t = make_float2(a[k], a[k] + 1.f);
        p[k] = *reinterpret_cast<u64*>(&t);
        h[k] = 0x3f803f80u + k + threadIdx.x;
    }
    float2 bb = make_float2(b, c);
    u64 pb = *reinterpret_cast<u64*>(&bb);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
            if (OP == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[k]) : "l"(pb));
            if (OP == 6) asm volatile("set.le.bf16x2.bf16x2 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            if (OP == 7) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            if (OP == 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(h[k]) : "r"(hb), "r"(h[(k + 1) % NACC]));
            if (OP == 10) {   // FADD + IADD alternating (dual pipe)
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            }
            if (OP == 11) {   // FMUL + FADD alternating
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[(k + 4) % NACC]) : "f"(c));
            }
            if (OP == 12) {   // FSETP + predicated add
                asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.s32 %0, %0, 1;}" : "+r"(h[k]) : "f"(a[k]), "f"(b));
            }
            if (OP == 13) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(h[k]) : "r"(hb));
            if (OP == 14) asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            if (OP == 15) {   // FADD2 + HSET2 mix
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(pb));
                asm volatile("set.le.bf16x2.bf16x2 %0, %0, %1;" : "+r"(h[k]) : "r"(hb));
            }
            if (OP == 16) {   // set.le.f32 -> mask form
                asm volatile("set.le.u32.f32 %0, %1, %2;" : "=r"(h[k]) : "f"(a[k]), "f"(b));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        float2 t = *reinterpret_cast<float2*>(&p[k]);
        s += a[k] + t.x + t.y + __uint_as_float(h[k]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    bench<OP><<<148, threads>>>(out, cyc, 1.5f);
    bench<OP><<<148, threads>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double winst = (double)ITERS * NACC * per_iter * (threads / 32);
    printf("%-28s threads=%4d  warp-inst/clk/SM = %.3f   (err=%s)\n", name, threads, winst / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {1024, 512}) {
        run<0>("FADD", 1, threads);
        run<1>("FMUL", 1, threads);
        run<2>("FFMA", 1, threads);
        run<3>("FADD2 (f32x2)", 1, threads);
        run<4>("FMUL2 (f32x2)", 1, threads);
        run<5>("FFMA2 (f32x2)", 1, threads);
        run<6>("HSET2.BF16 le", 1, threads);
        run<7>("HADD2.BF16", 1, threads);
        run<8>("IADD", 1, threads);
        run<9>("LOP3", 1, threads);
        run<10>("FADD+IADD pair", 2, threads);
        run<11>("FMUL+FADD pair", 2, threads);
        run<12>("FSETP+@P IADD pair", 2, threads);
        run<13>("PRMT", 1, threads);
        run<14>("VHMNMX bf16x2 min", 1, threads);
        run<15>("FADD2+HSET2 pair", 2, threads);
        run<16>("FSET.le mask", 1, threads);
    }
    return 0;
}
