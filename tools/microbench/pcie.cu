// Host-link micro-benchmark behind the design of gmr_retarget_batch_host (DESIGN.md "End to end").
// What can the link of THIS box do for the solver's access pattern?
//   (1) bulk DMA: cudaMemcpyAsync H2D / D2H, alone and both directions at once, 2-D (strided) copies by frame segment
//   (2) zero-copy: the solve kernel's own pattern - one warp per clip walks its frames in order and pulls one frame's
//       keypoints (168 B of positions + 224 B of quaternions, two arrays) per step with cp.async, D frames ahead -
//       straight from mapped pinned host memory; and writes one frame of qpos (144 B) per step to mapped pinned memory
//   (3) both at once (SM reads competing with a DMA stream)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie pcie.cu ;  run: ./pcie [clips] [frames]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NH = 14, NQ = 36;

// one warp per clip, frames in order.  mode 0: 4-byte pos + 16-byte quat cp.async per body (round-1 stage_frame);
// mode 1: 8-byte chunks of the contiguous pos block + 16-byte chunks of the quat block (coalesced);
// mode 2: TMA bulk copies (cp.async.bulk, mbarrier completion): the 16-byte aligned body of the pos block (160 B) + an
//         8-byte cp.async for its head or tail (frames are 168 B apart: odd frames start 8 bytes off), the quat block (224 B)
template <int DEPTH, int MODE>
__global__ void __launch_bounds__(512, 1)
read_kernel(const float* __restrict__ pos, const float* __restrict__ quat, int C, int T, int spin, float* __restrict__ sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* ring = reinterpret_cast<float*>(smem) + (size_t)warp * DEPTH * 112;      // per frame: 42 + 56 floats, padded
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(blockDim.x >> 5) * DEPTH * 112 * 4) + warp * DEPTH;
  if (MODE == 2) {
    if (lane == 0) for (int i = 0; i < DEPTH; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  uint32_t phase = 0;                                                             // bit i: parity of ring slot i
  float acc = 0.f;
  for (int c = blockIdx.x * wpc + warp; c < C; c += gridDim.x * wpc) {
    const float* p = pos + (size_t)c * T * NH * 3;
    const float* q = quat + (size_t)c * T * NH * 4;
    auto stage = [&](int t) {
      float* dst = ring + (t % DEPTH) * 112;
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
      const float* pf = p + (size_t)t * NH * 3;
      const float* qf = q + (size_t)t * NH * 4;
      if (MODE == 0) {
        if (lane < NH) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 12 * lane), "l"(pf + 3 * lane) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 12 * lane + 4), "l"(pf + 3 * lane + 1) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 12 * lane + 8), "l"(pf + 3 * lane + 2) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 176 + 16 * lane), "l"(qf + 4 * lane) : "memory");
        }
      } else if (MODE == 2) {
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(pf) & 15u);   // 0 or 8; data lands at dst + off
        if (lane == 0) {
          const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + (t % DEPTH));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(160u + 224u) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(d + (off ? 16u : 0u)), "l"(reinterpret_cast<const char*>(pf) + (off ? 8 : 0)), "r"(160u), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(d + 192u), "l"(qf), "r"(224u), "r"(bar) : "memory");
        }
        if (lane == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + (off ? 8u : 160u)), "l"(reinterpret_cast<const char*>(pf) + (off ? 0 : 160)) : "memory");
      } else {
        if (lane < 21) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 8 * lane), "l"(pf + 2 * lane) : "memory");
        if (lane < NH) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 176 + 16 * lane), "l"(qf + 4 * lane) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < DEPTH - 1 && t < T; t++) stage(t);
    for (int t = 0; t < T; t++) {
      if (t + DEPTH - 1 < T) stage(t + DEPTH - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
      if (MODE == 2) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + (t % DEPTH)), par = (phase >> (t % DEPTH)) & 1u;
        uint32_t done = 0;
        while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
        phase ^= 1u << (t % DEPTH);
      }
      __syncwarp();
      const float* src = ring + (t % DEPTH) * 112;
      acc += src[lane] + src[44 + lane];
      for (int s = 0; s < spin; s++) acc = fmaf(acc, 1.0000001f, 1e-9f);           // stands in for the solve
      __syncwarp();
    }
  }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
}

__global__ void __launch_bounds__(512, 1)
write_kernel(float* __restrict__ out, int C, int T, int spin) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float acc = (float)lane;
  for (int c = blockIdx.x * wpc + warp; c < C; c += gridDim.x * wpc) {
    float* o = out + (size_t)c * T * NQ;
    for (int t = 0; t < T; t++) {
      for (int s = 0; s < spin; s++) acc = fmaf(acc, 1.0000001f, 1e-9f);
      for (int i = lane; i < NQ; i += 32) o[(size_t)t * NQ + i] = acc + i;
    }
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 4096, T = argc > 2 ? atoi(argv[2]) : 300;
  const size_t b_pos = (size_t)C * T * NH * 3 * 4, b_quat = (size_t)C * T * NH * 4 * 4, b_out = (size_t)C * T * NQ * 4;
  float *h_pos, *h_quat, *h_out, *d_pos, *d_quat, *d_out, *d_sink;
  CK(cudaHostAlloc(&h_pos, b_pos, cudaHostAllocMapped)); CK(cudaHostAlloc(&h_quat, b_quat, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_out, b_out, cudaHostAllocMapped));
  for (size_t i = 0; i < b_pos / 4; i++) h_pos[i] = (float)(i & 1023);
  for (size_t i = 0; i < b_quat / 4; i++) h_quat[i] = (float)(i & 511);
  CK(cudaMalloc(&d_pos, b_pos)); CK(cudaMalloc(&d_quat, b_quat)); CK(cudaMalloc(&d_out, b_out)); CK(cudaMalloc(&d_sink, 4096));
  cudaStream_t s0, s1; CK(cudaStreamCreate(&s0)); CK(cudaStreamCreate(&s1));
  cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
  int sms = 148; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("{\"clips\": %d, \"frames\": %d, \"h2d_MB\": %.1f, \"d2h_MB\": %.1f, \"sms\": %d}\n", C, T, (b_pos + b_quat) / 1e6, b_out / 1e6, sms);

  for (int rep = 0; rep < 3; rep++) {      // bulk DMA
    CK(cudaEventRecord(e0, s0));
    CK(cudaMemcpyAsync(d_pos, h_pos, b_pos, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(d_quat, h_quat, b_quat, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(e1, s0)); CK(cudaStreamSynchronize(s0));
    const float h2d = time_ms(e0, e1);
    CK(cudaEventRecord(e0, s0)); CK(cudaMemcpyAsync(h_out, d_out, b_out, cudaMemcpyDeviceToHost, s0)); CK(cudaEventRecord(e1, s0)); CK(cudaStreamSynchronize(s0));
    const float d2h = time_ms(e0, e1);
    CK(cudaEventRecord(e0, s0)); CK(cudaEventRecord(f0, s1));
    CK(cudaMemcpyAsync(d_pos, h_pos, b_pos, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(d_quat, h_quat, b_quat, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpyAsync(h_out, d_out, b_out, cudaMemcpyDeviceToHost, s1));
    CK(cudaEventRecord(e1, s0)); CK(cudaEventRecord(f1, s1)); CK(cudaDeviceSynchronize());
    printf("{\"test\": \"dma\", \"rep\": %d, \"h2d_ms\": %.2f, \"h2d_GBs\": %.1f, \"d2h_ms\": %.2f, \"d2h_GBs\": %.1f, \"duplex_h2d_ms\": %.2f, \"duplex_d2h_ms\": %.2f}\n",
           rep, h2d, (b_pos + b_quat) / h2d / 1e6, d2h, b_out / d2h / 1e6, time_ms(e0, e1), time_ms(f0, f1));
  }
  for (int S : {4, 16, 64}) {               // strided copies: frames [0, S) of every clip
    CK(cudaEventRecord(e0, s0));
    CK(cudaMemcpy2DAsync(d_pos, (size_t)S * NH * 12, h_pos, (size_t)T * NH * 12, (size_t)S * NH * 12, C, cudaMemcpyHostToDevice, s0));
    CK(cudaMemcpy2DAsync(d_quat, (size_t)S * NH * 16, h_quat, (size_t)T * NH * 16, (size_t)S * NH * 16, C, cudaMemcpyHostToDevice, s0));
    CK(cudaEventRecord(e1, s0)); CK(cudaStreamSynchronize(s0));
    const double mb = (double)C * S * NH * 28 / 1e6; const float ms = time_ms(e0, e1);
    printf("{\"test\": \"dma_2d\", \"segment_frames\": %d, \"MB\": %.1f, \"ms\": %.3f, \"GBs\": %.2f}\n", S, mb, ms, mb / ms);
  }
  {                                          // many small contiguous copies (one clip each)
    const int n = 512;
    CK(cudaEventRecord(e0, s0));
    for (int c = 0; c < n; c++) {
      CK(cudaMemcpyAsync(d_pos + (size_t)c * T * NH * 3, h_pos + (size_t)c * T * NH * 3, (size_t)T * NH * 12, cudaMemcpyHostToDevice, s0));
      CK(cudaMemcpyAsync(d_quat + (size_t)c * T * NH * 4, h_quat + (size_t)c * T * NH * 4, (size_t)T * NH * 16, cudaMemcpyHostToDevice, s0));
    }
    CK(cudaEventRecord(e1, s0)); CK(cudaStreamSynchronize(s0));
    const double mb = (double)n * T * NH * 28 / 1e6; const float ms = time_ms(e0, e1);
    printf("{\"test\": \"dma_per_clip\", \"copies\": %d, \"MB\": %.1f, \"ms\": %.3f, \"GBs\": %.2f}\n", 2 * n, mb, ms, mb / ms);
  }

  float *m_pos, *m_quat, *m_out;
  CK(cudaHostGetDevicePointer(&m_pos, h_pos, 0)); CK(cudaHostGetDevicePointer(&m_quat, h_quat, 0)); CK(cudaHostGetDevicePointer(&m_out, h_out, 0));
  auto run_read = [&](const char* name, int depth, int mode, const float* p, const float* q, int spin, bool with_dma) {
    const int wpc = 16; const size_t smem = (size_t)wpc * depth * 112 * 4;
    for (int rep = 0; rep < 2; rep++) {
      if (with_dma) { CK(cudaEventRecord(f0, s1)); CK(cudaMemcpyAsync(d_pos, h_pos, b_pos, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(d_quat, h_quat, b_quat, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(f1, s1)); }
      CK(cudaEventRecord(e0, s0));
      if (mode == 2 && depth == 2) read_kernel<2, 2><<<sms, wpc * 32, smem + wpc * depth * 8, s0>>>(p, q, C, T, spin, d_sink);
      else if (mode == 2) read_kernel<4, 2><<<sms, wpc * 32, smem + wpc * depth * 8, s0>>>(p, q, C, T, spin, d_sink);
      else if (depth == 2 && mode == 0) read_kernel<2, 0><<<sms, wpc * 32, smem, s0>>>(p, q, C, T, spin, d_sink);
      else if (depth == 2) read_kernel<2, 1><<<sms, wpc * 32, smem, s0>>>(p, q, C, T, spin, d_sink);
      else if (depth == 4) read_kernel<4, 1><<<sms, wpc * 32, smem, s0>>>(p, q, C, T, spin, d_sink);
      else { printf("unsupported depth\n"); return; }
      CK(cudaGetLastError());
      CK(cudaEventRecord(e1, s0)); CK(cudaDeviceSynchronize());
      const float ms = time_ms(e0, e1);
      printf("{\"test\": \"%s\", \"depth\": %d, \"mode\": %d, \"spin\": %d, \"rep\": %d, \"ms\": %.2f, \"GBs\": %.2f", name, depth, mode, spin, rep, ms, (b_pos + b_quat) / ms / 1e6);
      if (with_dma) printf(", \"concurrent_dma_ms\": %.2f", time_ms(f0, f1));
      printf("}\n");
    }
  };
  run_read("read_hbm", 2, 1, d_pos, d_quat, 0, false);
  run_read("read_zero_copy", 2, 0, m_pos, m_quat, 0, false);
  run_read("read_zero_copy", 2, 1, m_pos, m_quat, 0, false);
  run_read("read_zero_copy", 2, 2, m_pos, m_quat, 0, false);
  run_read("read_hbm", 2, 2, d_pos, d_quat, 0, false);
  run_read("read_zero_copy", 4, 2, m_pos, m_quat, 0, false);
  run_read("read_zero_copy", 4, 1, m_pos, m_quat, 0, false);
  // with a stand-in for the solve (~80 ms of ALU time per batch when data is free): does the link keep up?
  run_read("read_hbm_paced", 2, 1, d_pos, d_quat, 6000, false);
  run_read("read_zero_copy_paced", 2, 1, m_pos, m_quat, 6000, false);
  run_read("read_zero_copy_paced", 4, 1, m_pos, m_quat, 6000, false);
  run_read("read_zero_copy_paced", 2, 2, m_pos, m_quat, 6000, false);
  run_read("read_zero_copy_with_dma", 4, 1, m_pos, m_quat, 0, true);

  for (int rep = 0; rep < 2; rep++) {
    for (int k = 0; k < 2; k++) {
      float* dst = k ? m_out : d_out;
      CK(cudaEventRecord(e0, s0)); write_kernel<<<sms, 512, 0, s0>>>(dst, C, T, 0); CK(cudaGetLastError()); CK(cudaEventRecord(e1, s0)); CK(cudaDeviceSynchronize());
      const float ms = time_ms(e0, e1);
      printf("{\"test\": \"%s\", \"rep\": %d, \"ms\": %.2f, \"GBs\": %.2f}\n", k ? "write_zero_copy" : "write_hbm", rep, ms, b_out / ms / 1e6);
    }
  }
  // zero-copy writes + zero-copy reads at once (the full streaming pattern), paced
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0, s0));
    read_kernel<4, 1><<<sms / 2, 512, 16 * 4 * 112 * 4, s0>>>(m_pos, m_quat, C, T, 0, d_sink);
    write_kernel<<<sms / 2, 512, 0, s1>>>(m_out, C, T, 0);
    CK(cudaEventRecord(e1, s0)); CK(cudaDeviceSynchronize());
    printf("{\"test\": \"zero_copy_read+write\", \"rep\": %d, \"read_ms\": %.2f}\n", rep, time_ms(e0, e1));
  }
  return 0;
}
