// libgmr_b200.so — sm_100a kernels and the C ABI of include/gmr_b200.h.
//
// One persistent CTA per SM, one warp per clip (gmr_solver.cuh).  The per-robot constant
// block is staged once per CTA into shared memory (one TMA bulk copy, cp.async.bulk, issued
// by an elected thread and awaited on an mbarrier); per-warp solver state follows it in the
// same dynamic allocation.  Work is sharded by clip with no inter-warp communication, so the
// grid is sized to the SM count and warps stride over clips.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "gmr_solver.cuh"

#define GMR_FLAG_INTERNAL_CONVOY 0x80000000u   // set by launch(), never by callers


namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int set_err(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_err(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  (void)cudaGetLastError();      // do not leave a sticky "last error" behind for the caller's next launch
  return GMR_ECUDA;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_err(_e, #call); } while (0)

template <typename R> constexpr int consts_bytes() { return (int)((sizeof(GmrConsts<R>) + 15) / 16 * 16); }

// ---- TMA bulk copy global -> shared, completion on an mbarrier -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_consts_tma(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t bar_a = smem_u32(bar), dst_a = smem_u32(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_a), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
  }
  // everyone waits for phase 0
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
  }
}

template <typename R, typename IO, int MAXWARPS>
__global__ void __launch_bounds__(MAXWARPS * 32, 1)
gmr_retarget_kernel(const __grid_constant__ GmrDims dm, const __grid_constant__ GmrScal<R> ks,
                    const GmrConsts<R>* __restrict__ gconsts, const float* __restrict__ pos, const float* __restrict__ quat,
                    const float* __restrict__ ratio, int C, int T, const IO* __restrict__ qinit, IO* __restrict__ qout,
                    int32_t* __restrict__ iters, IO* __restrict__ err, IO* __restrict__ tg, uint32_t flags,
                    int* __restrict__ queue) {
  extern __shared__ __align__(128) unsigned char gmr_dyn_smem[];
  unsigned char* const smem = gmr_dyn_smem;
  constexpr int CB = consts_bytes<R>();
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);                 // 16 bytes reserved for the mbarrier
  GmrConsts<R>* mc = reinterpret_cast<GmrConsts<R>*>(smem + 16);
  stage_consts_tma(mc, gconsts, CB, bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int wel = dm.warp_elems;
  WarpSolver<R> ws(*mc, dm, ks, (uint32_t)(16 + CB) + (uint32_t)warp * (uint32_t)wel * (uint32_t)sizeof(R), lane);
  ws.convoy = (flags & GMR_FLAG_INTERNAL_CONVOY) != 0;
  ws.cta_active = reinterpret_cast<int*>(smem + 8);                  // second half of the mbarrier's 16-byte slot
  if (ws.convoy) {
    if (threadIdx.x == 0) *ws.cta_active = wpc;
    __syncthreads();
  }
  const int nq = dm.nq, nhum = dm.nhum;
  // Clips come from a global queue (one atomic per clip): clips differ several-fold in the number of
  // IK steps they need, so a warp that finishes early takes the next clip instead of idling behind a
  // static assignment.  The first gridDim * wpc clips are handed out without touching the queue.
  const int nw = gridDim.x * wpc;
  for (int c = warp * gridDim.x + blockIdx.x; c < C;) {
    const size_t f0 = (size_t)c * T;
    ws.template run_clip<IO>(pos + f0 * nhum * 3, quat + f0 * nhum * 4, ratio ? R(ratio[c]) : R(1), T,
                             qinit ? qinit + (size_t)c * nq : nullptr, qout + f0 * nq,
                             iters ? iters + 2 * f0 : nullptr, err ? err + 2 * f0 : nullptr,
                             tg ? tg + f0 * nhum * 7 : nullptr, flags);
    int nxt = 0;
    if (lane == 0) nxt = nw + atomicAdd(queue, 1);
    c = __shfl_sync(0xffffffffu, nxt, 0);
  }
  ws.convoy_retire();
}

constexpr int QUEUE_RING = 256;
constexpr int MAXW_F32 = 28;   // 28 warps * 32 lanes * 72 registers = one SM's register file
constexpr int MAXW_F64 = 16;   // 16 warps * 32 lanes * 128 registers

}  // namespace

struct GmrModel {
  int device = 0;
  int num_sms = 0;
  int max_smem = 0;
  GmrConsts<float>* d_f32 = nullptr;
  GmrConsts<double>* d_f64 = nullptr;
  GmrConsts<float> h_f32;
  GmrDims dims32{}, dims64{};
  GmrScal<float> ks32{};
  GmrScal<double> ks64{};
  int wel32 = 0, wel64 = 0;      // per-warp shared-memory elements
  int* d_queue = nullptr;        // ring of clip-queue counters, one per launch in flight
  std::atomic<uint32_t> queue_next{0};
  // lazily created resources of the host-buffer entry
  std::mutex host_mu;
  cudaStream_t hs[2] = {nullptr, nullptr};
  void* hbuf[2] = {nullptr, nullptr};
  size_t hbuf_bytes[2] = {0, 0};
};

namespace {

struct DeviceGuard {
  int prev = -1; bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    else (void)cudaGetLastError();
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename R> int max_warps();
template <> int max_warps<float>() { return MAXW_F32; }
template <> int max_warps<double>() { return MAXW_F64; }

template <typename R> size_t smem_bytes(const GmrModel* m, int wpc) {
  const int wel = sizeof(R) == 4 ? m->wel32 : m->wel64;
  return 16 + (size_t)consts_bytes<R>() + (size_t)wpc * wel * sizeof(R);
}
template <typename R> const GmrScal<R>& scal_of(const GmrModel* m);
template <> const GmrScal<float>& scal_of<float>(const GmrModel* m) { return m->ks32; }
template <> const GmrScal<double>& scal_of<double>(const GmrModel* m) { return m->ks64; }

// warps per CTA: as many as fit (shared memory, register file), but no more than the clips need
// CTAs (= independent convoys) per SM: two convoys overlap one's stragglers with the other's work
inline int ctas_per_sm() {
  static const int n = [] { const char* e = getenv("GMR_CTAS_PER_SM"); int v = e ? atoi(e) : 1; return v < 1 ? 1 : (v > 7 ? 7 : v); }();
  return n;
}
template <typename R> int pick_wpc(const GmrModel* m, int C) {
  int wpc = max_warps<R>() / ctas_per_sm();
  if (const char* cap = getenv("GMR_WPC_CAP")) { int c = atoi(cap); if (c >= 1 && c < wpc) wpc = c; }   // tuning knob
  while (wpc > 1 && smem_bytes<R>(m, wpc) > (size_t)m->max_smem) wpc--;
  int need = (C + m->num_sms * ctas_per_sm() - 1) / (m->num_sms * ctas_per_sm());
  if (need < 1) need = 1;
  if (need < wpc) wpc = need;
  return wpc;
}

template <typename R, typename IO, int MAXWARPS>
int launch(GmrModel* m, const GmrConsts<R>* dc, const float* pos, const float* quat, const float* ratio, int C, int T,
           const IO* qinit, IO* qout, int32_t* iters, IO* err, IO* tg, uint32_t flags, cudaStream_t st) {
  if (C == 0 || T == 0) return GMR_OK;
  auto kern = gmr_retarget_kernel<R, IO, MAXWARPS>;
  const int wpc = pick_wpc<R>(m, C);
  const size_t smem = smem_bytes<R>(m, wpc);
  if (smem > (size_t)m->max_smem) return set_err(GMR_ELIMIT, "model does not fit in shared memory");
  static thread_local const void* configured = nullptr;   // opt in to the large dynamic allocation once per kernel
  if (configured != (const void*)kern) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m->max_smem));
    configured = (const void*)kern;
  }
  int grid = (C + wpc - 1) / wpc;
  if (grid > m->num_sms * ctas_per_sm()) grid = m->num_sms * ctas_per_sm();
  // convoy mode (see gmr_solver.cuh) paid off while the hot path was 3x the instruction cache; with the
  // branch-sparse factorisation free-running warps are as fast on balanced batches and 8-10 % faster on
  // mixed ones (a slow clip no longer drags its CTA through a rendezvous per factorisation).  GMR_CONVOY=1 re-enables.
  static const int cv_env = getenv("GMR_CONVOY") ? atoi(getenv("GMR_CONVOY")) : 0;
  const bool convoy = cv_env != 0 && wpc >= 2;
  flags = (flags & 0xffffu) | (convoy ? GMR_FLAG_INTERNAL_CONVOY : 0u);
  int* queue = m->d_queue + (m->queue_next.fetch_add(1) % QUEUE_RING);
  CK(cudaMemsetAsync(queue, 0, sizeof(int), st));
  kern<<<grid, wpc * 32, smem, st>>>(sizeof(R) == 4 ? m->dims32 : m->dims64, scal_of<R>(m), dc, pos, quat, ratio, C, T, qinit, qout, iters, err, tg, flags, queue);
  g_launches.fetch_add(1);
  CK(cudaGetLastError());
  return GMR_OK;
}

}  // namespace

extern "C" {

int gmr_model_create(const GmrModelDesc* desc, int device, GmrModel** out) {
  if (!out) return set_err(GMR_EINVAL, "out is null");
  *out = nullptr;
  GmrModel* m = new (std::nothrow) GmrModel();
  if (!m) return set_err(GMR_ENOMEM, "host allocation failed");
  const char* why = nullptr;
  auto* hd = new (std::nothrow) GmrConsts<double>();
  if (!hd) { delete m; return set_err(GMR_ENOMEM, "host allocation failed"); }
  int rc = gmr_fill_consts<float>(desc, &m->h_f32, &why);
  if (rc == GMR_OK) rc = gmr_fill_consts<double>(desc, hd, &why);
  if (rc != GMR_OK) { delete hd; delete m; return set_err(rc, why ? why : "invalid model"); }
  m->dims32 = gmr_dims_of(m->h_f32); gmr_dims_layout<float>(m->dims32);
  m->dims64 = gmr_dims_of(m->h_f32); gmr_dims_layout<double>(m->dims64);
  m->ks32 = gmr_scal_of(m->h_f32);
  m->ks64 = gmr_scal_of(*hd);
  m->wel32 = m->dims32.warp_elems;
  m->wel64 = m->dims64.warp_elems;
  m->device = device;
  DeviceGuard g(device);
  cudaError_t e = g.ok ? cudaSuccess : cudaErrorInvalidDevice;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_f32, consts_bytes<float>());
  if (e == cudaSuccess) e = cudaMalloc(&m->d_f64, consts_bytes<double>());
  if (e == cudaSuccess) e = cudaMalloc(&m->d_queue, QUEUE_RING * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(m->d_f32, 0, consts_bytes<float>());
  if (e == cudaSuccess) e = cudaMemset(m->d_f64, 0, consts_bytes<double>());
  if (e == cudaSuccess) e = cudaMemcpy(m->d_f32, &m->h_f32, sizeof(GmrConsts<float>), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_f64, hd, sizeof(GmrConsts<double>), cudaMemcpyHostToDevice);
  delete hd;
  if (e != cudaSuccess) {
    int code = cuda_err(e, "gmr_model_create");
    if (m->d_f32) cudaFree(m->d_f32);
    if (m->d_f64) cudaFree(m->d_f64);
    if (m->d_queue) cudaFree(m->d_queue);
    delete m;
    return code;
  }
  if (smem_bytes<float>(m, 1) > (size_t)m->max_smem) {
    cudaFree(m->d_f32); cudaFree(m->d_f64); cudaFree(m->d_queue); delete m;
    return set_err(GMR_ELIMIT, "model does not fit in shared memory");
  }
  *out = m;
  return GMR_OK;
}

int gmr_model_destroy(GmrModel* m) {
  if (!m) return GMR_OK;
  DeviceGuard g(m->device);
  for (int i = 0; i < 2; i++) {
    if (m->hs[i]) { cudaStreamSynchronize(m->hs[i]); cudaStreamDestroy(m->hs[i]); }
    if (m->hbuf[i]) cudaFree(m->hbuf[i]);
  }
  cudaFree(m->d_f32);
  cudaFree(m->d_f64);
  cudaFree(m->d_queue);
  delete m;
  return GMR_OK;
}

int gmr_retarget_batch(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                       const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out, float* targets_out,
                       uint32_t flags, void* cuda_stream) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if ((C > 0 && T > 0) && (!pos || !quat || !qpos_out)) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  if ((reinterpret_cast<uintptr_t>(quat) & 15u) != 0) return set_err(GMR_EINVAL, "quat must be 16-byte aligned");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  if (flags & GMR_FLAG_COMPUTE_F64)
    return launch<double, float, MAXW_F64>(m, m->d_f64, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                           targets_out, flags, (cudaStream_t)cuda_stream);
  return launch<float, float, MAXW_F32>(m, m->d_f32, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                        targets_out, flags, (cudaStream_t)cuda_stream);
}

int gmr_retarget_batch_f64(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                           const double* qpos_init, double* qpos_out, int32_t* iters_out, double* err_out,
                           double* targets_out, uint32_t flags, void* cuda_stream) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if ((C > 0 && T > 0) && (!pos || !quat || !qpos_out)) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  if ((reinterpret_cast<uintptr_t>(quat) & 15u) != 0) return set_err(GMR_EINVAL, "quat must be 16-byte aligned");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  return launch<double, double, MAXW_F64>(m, m->d_f64, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                          targets_out, flags, (cudaStream_t)cuda_stream);
}

// Host-buffer entry: clips are cut into chunks; chunk i+1's host->device copy and chunk i-1's
// device->host copy overlap chunk i's solve on two streams with private staging buffers.
int gmr_retarget_batch_host(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                            const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out, uint32_t flags) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if (C == 0 || T == 0) return GMR_OK;
  if (!pos || !quat || !qpos_out) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  std::lock_guard<std::mutex> lk(m->host_mu);
  const int nq = m->h_f32.nq, nh = m->h_f32.nhum;
  // per-clip device bytes (each array padded to 16 bytes per chunk below)
  const size_t b_pos = (size_t)T * nh * 3 * 4, b_quat = (size_t)T * nh * 4 * 4, b_q = (size_t)T * nq * 4,
               b_it = iters_out ? (size_t)T * 2 * 4 : 0, b_err = err_out ? (size_t)T * 2 * 4 : 0,
               b_ratio = ratio ? 4 : 0, b_init = qpos_init ? (size_t)nq * 4 : 0;
  const size_t per_clip = b_pos + b_quat + b_q + b_it + b_err + b_ratio + b_init;
  // chunks of whole "waves" (one clip per resident warp) so that every chunk fills the GPU
  const int wave = m->num_sms * ((flags & GMR_FLAG_COMPUTE_F64) ? pick_wpc<double>(m, C) : pick_wpc<float>(m, C));
  int chunk = C;
  if (C >= 4 * wave) chunk = ((C / 4 + wave - 1) / wave) * wave;            // >= 4 chunks when there is enough work
  const size_t budget = (size_t)8 << 30;
  while ((size_t)chunk * per_clip > budget && chunk > wave) chunk -= wave;
  for (int i = 0; i < 2; i++) if (!m->hs[i]) CK(cudaStreamCreateWithFlags(&m->hs[i], cudaStreamNonBlocking));
  int rc = GMR_OK;
  int k = 0;
  for (int c0 = 0; c0 < C && rc == GMR_OK; c0 += chunk, k ^= 1) {
    const int n = (C - c0 < chunk) ? C - c0 : chunk;
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_pos = 0, o_quat = o_pos + pad(b_pos * n), o_q = o_quat + pad(b_quat * n), o_it = o_q + pad(b_q * n),
                 o_err = o_it + pad(b_it * n), o_ratio = o_err + pad(b_err * n), o_init = o_ratio + pad(b_ratio * n),
                 total = o_init + pad(b_init * n);
    cudaStream_t st = m->hs[k];
    CK(cudaStreamSynchronize(st));                           // staging buffer k is free again
    if (m->hbuf_bytes[k] < total) {
      if (m->hbuf[k]) CK(cudaFree(m->hbuf[k]));
      m->hbuf[k] = nullptr; m->hbuf_bytes[k] = 0;
      CK(cudaMalloc(&m->hbuf[k], total));
      m->hbuf_bytes[k] = total;
    }
    char* d = (char*)m->hbuf[k];
    CK(cudaMemcpyAsync(d + o_pos, pos + (size_t)c0 * T * nh * 3, b_pos * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_quat, quat + (size_t)c0 * T * nh * 4, b_quat * n, cudaMemcpyHostToDevice, st));
    if (ratio) CK(cudaMemcpyAsync(d + o_ratio, ratio + c0, b_ratio * n, cudaMemcpyHostToDevice, st));
    if (qpos_init) CK(cudaMemcpyAsync(d + o_init, qpos_init + (size_t)c0 * nq, b_init * n, cudaMemcpyHostToDevice, st));
    if (flags & GMR_FLAG_COMPUTE_F64)
      rc = launch<double, float, MAXW_F64>(m, m->d_f64, (const float*)(d + o_pos), (const float*)(d + o_quat),
                                           ratio ? (const float*)(d + o_ratio) : nullptr, n, T,
                                           qpos_init ? (const float*)(d + o_init) : nullptr, (float*)(d + o_q),
                                           iters_out ? (int32_t*)(d + o_it) : nullptr, err_out ? (float*)(d + o_err) : nullptr,
                                           nullptr, flags, st);
    else
      rc = launch<float, float, MAXW_F32>(m, m->d_f32, (const float*)(d + o_pos), (const float*)(d + o_quat),
                                          ratio ? (const float*)(d + o_ratio) : nullptr, n, T,
                                          qpos_init ? (const float*)(d + o_init) : nullptr, (float*)(d + o_q),
                                          iters_out ? (int32_t*)(d + o_it) : nullptr, err_out ? (float*)(d + o_err) : nullptr,
                                          nullptr, flags, st);
    if (rc != GMR_OK) break;
    CK(cudaMemcpyAsync(qpos_out + (size_t)c0 * T * nq, d + o_q, b_q * n, cudaMemcpyDeviceToHost, st));
    if (iters_out) CK(cudaMemcpyAsync(iters_out + (size_t)c0 * T * 2, d + o_it, b_it * n, cudaMemcpyDeviceToHost, st));
    if (err_out) CK(cudaMemcpyAsync(err_out + (size_t)c0 * T * 2, d + o_err, b_err * n, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaStreamSynchronize(m->hs[i]);
    if (e != cudaSuccess && rc == GMR_OK) rc = cuda_err(e, "cudaStreamSynchronize");
  }
  return rc;
}

int64_t gmr_launch_count(void) { return g_launches.load(); }

const char* gmr_last_error(void) { return g_err.c_str(); }

int gmr_kernel_info(GmrModel* m, int32_t precision_bits, int32_t* threads_per_cta, int32_t* clips_per_cta,
                    int32_t* smem_bytes_out, int32_t* regs_per_thread, int32_t* ctas_per_sm) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  DeviceGuard g(m->device);
  cudaFuncAttributes fa{};
  int wpc; size_t smem;
  if (precision_bits == 64) {
    CK(cudaFuncGetAttributes(&fa, gmr_retarget_kernel<double, double, MAXW_F64>));
    wpc = pick_wpc<double>(m, 1 << 30); smem = smem_bytes<double>(m, wpc);
  } else {
    CK(cudaFuncGetAttributes(&fa, gmr_retarget_kernel<float, float, MAXW_F32>));
    wpc = pick_wpc<float>(m, 1 << 30); smem = smem_bytes<float>(m, wpc);
  }
  if (threads_per_cta) *threads_per_cta = wpc * 32;
  if (clips_per_cta) *clips_per_cta = wpc;
  if (smem_bytes_out) *smem_bytes_out = (int32_t)smem;
  if (regs_per_thread) *regs_per_thread = fa.numRegs;
  if (ctas_per_sm) *ctas_per_sm = 1;
  return GMR_OK;
}

}  // extern "C"
