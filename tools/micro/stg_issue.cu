// Microbenchmark: cost of warp-level global store instructions on one SM as a function of the access width
// and of the number of active lanes (run under gpurun: nvcc -arch=sm_100a stg_issue.cu -o stg_issue && ./stg_issue).
// Every warp stores into its own 8 KB window (L2-resident, write-through L1), 24 warps per SM, 148 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(768, 1) k(uint8_t* buf, int reps, unsigned long long* cyc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* base = buf + (static_cast<size_t>(blockIdx.x) * 24 + warp) * 8192;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const int slot = r & 15;  // 16 x 512 B slots
        uint8_t* p = base + slot * 512;
        if (MODE == 0) p[lane] = static_cast<uint8_t>(r);                                            // 32 lanes x 1 B
        if (MODE == 1) reinterpret_cast<uint16_t*>(p)[lane] = static_cast<uint16_t>(r);              // 32 x 2 B
        if (MODE == 2) reinterpret_cast<uint32_t*>(p)[lane] = r;                                     // 32 x 4 B
        if (MODE == 3) reinterpret_cast<uint4*>(p)[lane] = make_uint4(r, r, r, r);                   // 32 x 16 B
        if (MODE == 4 && lane < 8) reinterpret_cast<uint4*>(p)[lane] = make_uint4(r, r, r, r);       // 8 x 16 B
        if (MODE == 5 && lane < 8) reinterpret_cast<uint32_t*>(p)[lane] = r;                         // 8 x 4 B
        if (MODE == 6 && lane < 16) reinterpret_cast<uint2*>(p)[lane] = make_uint2(r, r);            // 16 x 8 B
        if (MODE == 7) reinterpret_cast<uint2*>(p)[lane] = make_uint2(r, r);                         // 32 x 8 B
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, uint8_t* buf, unsigned long long* cyc, int bytes_per_instr) {
    const int reps = 20000;
    k<MODE><<<148, 768>>>(buf, 100, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 768>>>(buf, reps, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double per_instr = avg / (static_cast<double>(reps) * 24);  // SM cycles per warp-level store instruction
    printf("%-14s %8.3f ms  %6.2f cycles per warp store (per SM)  %7.1f GB/s chip\n", name, ms, per_instr,
           148.0 * 24 * reps * bytes_per_instr / (ms * 1e-3) / 1e9);
}

int main() {
    uint8_t* buf;
    unsigned long long* cyc;
    cudaMalloc(&buf, static_cast<size_t>(148) * 24 * 8192);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("32 x u8", buf, cyc, 32);
    run<1>("32 x u16", buf, cyc, 64);
    run<2>("32 x u32", buf, cyc, 128);
    run<7>("32 x u64", buf, cyc, 256);
    run<3>("32 x u128", buf, cyc, 512);
    run<4>("8 x u128", buf, cyc, 128);
    run<5>("8 x u32", buf, cyc, 32);
    run<6>("16 x u64", buf, cyc, 128);
    return 0;
}
