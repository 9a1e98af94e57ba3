// Instruction-supply microbenchmark: 28 desynchronised warps per SM each loop over a straight-line
// body of N FFMAs (4 independent accumulators, so a lone warp is issue- not latency-bound).
// Prints achieved warp-instructions/cycle/SM versus body size: the I-cache capacity knee.
#include <cstdio>
#include <cuda_runtime.h>
#define F1 a0 = fmaf(a0, x, y); a1 = fmaf(a1, x, y); a2 = fmaf(a2, x, y); a3 = fmaf(a3, x, y);
#define F4 F1 F1 F1 F1
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define F256 F64 F64 F64 F64
#define F1K F256 F256 F256 F256
#define F4K F1K F1K F1K F1K
template <int KIND> __global__ void k(float* out, int iters, int desync, long long* cyc) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, x = 1.0001f, y = 0.5f;
  if (desync) { long long t0 = clock64(); long long wait = ((threadIdx.x >> 5) * 7919LL) % 20011; while (clock64() - t0 < wait) {} }
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (KIND == 0) { F64 }            // 256 instr
    if (KIND == 1) { F256 }           // 1k
    if (KIND == 2) { F256 F256 }      // 2k
    if (KIND == 3) { F1K }            // 4k
    if (KIND == 4) { F1K F1K }        // 8k
    if (KIND == 5) { F4K }            // 16k
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const int sizes[6] = {256, 1024, 2048, 4096, 8192, 16384};
  for (int warps : {4, 8, 16, 28}) for (int desync : {0, 1}) {
    printf("warps/SM %2d desync %d:", warps, desync);
    for (int kind = 0; kind < 6; kind++) {
      int iters = (1 << 22) / sizes[kind];
      auto run = [&](int it) {
        switch (kind) { case 0: k<0><<<148, warps * 32>>>(out, it, desync, cyc); break; case 1: k<1><<<148, warps * 32>>>(out, it, desync, cyc); break;
          case 2: k<2><<<148, warps * 32>>>(out, it, desync, cyc); break; case 3: k<3><<<148, warps * 32>>>(out, it, desync, cyc); break;
          case 4: k<4><<<148, warps * 32>>>(out, it, desync, cyc); break; default: k<5><<<148, warps * 32>>>(out, it, desync, cyc); break; } };
      run(2); cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); run(iters); cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = (double)iters * sizes[kind] * warps;            // per SM
      double cycles = ms * 1e-3 * 1.965e9;
      printf("  %5d:%5.2f", sizes[kind], instr / cycles);
    }
    printf("   (warp-instr/cycle/SM)\n");
  }
  return 0;
}
