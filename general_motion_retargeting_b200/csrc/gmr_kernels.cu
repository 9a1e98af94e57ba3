// libgmr_b200.so — sm_100a kernels and the C ABI of include/gmr_b200.h.
//
// One persistent CTA per SM, one warp per clip (gmr_solver.cuh).  The per-robot constant
// block is staged once per CTA into shared memory (one TMA bulk copy, cp.async.bulk, issued
// by an elected thread and awaited on an mbarrier); per-warp solver state follows it in the
// same dynamic allocation.  Work is sharded by clip with no inter-warp communication: the grid is
// sized to the SM count and warps take clips from a global queue; launch() orders the queue (hard
// clips first) or, for batches of 1-4 waves, runs the two-phase schedule (frame 0 -> classify ->
// the rest with slow clips on their own SMs), see DESIGN.md "Clip scheduling".  Also here: the
// motion-array epilogue kernel, the human-frame producers (gmr_producers.cuh) and the live stream.
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "gmr_solver.cuh"
#include "gmr_producers.cuh"

#define GMR_FLAG_INTERNAL_CONVOY 0x80000000u   // set by launch(), never by callers
#define GMR_FLAG_INTERNAL_NOSTEAL 0x40000000u  // experiment knob (GMR_NO_STEAL=1): dense SMs never take slow clips


namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
std::atomic<long long*> g_trace{nullptr};      // gmr_debug_trace

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

// Scheduling hint for batches larger than the GPU's warp slots: clips whose first target asks for a large root
// rotation away from the starting configuration are the ones that settle on joint limits and need several times
// the work of a median clip (on the benchmark clips every slow clip is above 2.5 rad).  They go to
// the front of the clip queue so that the longest chains start at t = 0 instead of in the last wave.  Pure
// ordering: a clip's result does not depend on when or where it runs.
//   order[0 .. n_hard) hard clips, order[C .. C + n_easy) easy clips, counts = {n_hard, n_easy}
template <typename R, typename IO>
__global__ void __launch_bounds__(256)
gmr_order_kernel(const GmrConsts<R>* __restrict__ mc, const float* __restrict__ quat, const IO* __restrict__ qinit,
                 int C, int T, int* __restrict__ counts, int* __restrict__ order) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int nhum = mc->nhum, hr = mc->hroot, nq = mc->nq;
  const float* q = quat + ((size_t)c * T * nhum + hr) * 4;
  const R* o = mc->hroff + 4 * hr;
  // target root orientation = human root * table-1 rotation offset (motion_retarget.py:241-246), unnormalised
  const float a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
  const float b0 = (float)o[0], b1 = (float)o[1], b2 = (float)o[2], b3 = (float)o[3];
  const float t0 = a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, t1 = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
              t2 = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, t3 = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0;
  float r0, r1, r2, r3;
  if (qinit) { const IO* qi = qinit + (size_t)c * nq + 3; r0 = (float)qi[0]; r1 = (float)qi[1]; r2 = (float)qi[2]; r3 = (float)qi[3]; }
  else { r0 = (float)mc->qpos0[3]; r1 = (float)mc->qpos0[4]; r2 = (float)mc->qpos0[5]; r3 = (float)mc->qpos0[6]; }
  const float dot = fabsf(t0 * r0 + t1 * r1 + t2 * r2 + t3 * r3) *
                    rsqrtf(fmaxf((t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3) * (r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3), 1e-30f));
  const bool hard = dot < 0.4078f;                 // rotation angle 2 acos(|dot|) > 2.3 rad
  if (hard) order[atomicAdd(&counts[0], 1)] = c;
  else order[C + atomicAdd(&counts[1], 1)] = c;
}

// ---- clip scheduler of the scheduled launch (gmr_retarget_kernel, part_w > 0) ---------------------------------------
// Two FIFO rings of clip ids in global memory: ring entries are 64-bit {ticket + 1, clip}; the slow ring occupies
// ring[0 .. cap), the normal ring ring[cap .. 2 cap).  gmr_classify_kernel fills the first n_slow0 / n_norm0 tickets, then the
// warps push and pop.  Every operation is a handful of independent atomics (no compare-and-swap loop, no ordered publish):
//   push: ticket = res++ ; ring[ticket % cap] = {ticket + 1, clip} ; avail++
//   pop : if (avail-- <= 0) { avail++ ; empty } else { ticket = head++ ; wait until ring[ticket % cap].ticket == ticket + 1 }
// (`avail` counts published entries, so a claimed ticket's entry is at most a few instructions away.)  A clip is in at most one
// ring at a time, so a ring never holds more than C entries; with cap = 2 C + 2 an entry could only be overwritten before its
// claimant reads it if C + 1 further tickets were claimed in between.
enum { SQ_AVAIL_N = 0, SQ_NSLOW0 = 1, SQ_NNORM0 = 2, SQ_AVAIL_S = 3, SQ_RES_N = 4, SQ_HEAD_N = 5, SQ_RES_S = 6, SQ_HEAD_S = 7,
       SQ_DONE = 8, SQ_IDLE_STEAL = 9, SQ_LIVE_SLOW = 10 /* slow clips alive, relative to n_slow0 */,
       SQ_FR_NORM = 12 /* frames solved so far by clips taken as normal */, SQ_INTS = 16 };
__device__ __forceinline__ int ldv(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
// Is a slow clip on the batch's critical path?  It is when it lags the normal class, measured in frames solved (every clip of
// a batch has about the same number of frames): t frames of its own against fn / nn frames per normal clip - such a clip will
// still be running when the bulk is done, so every microsecond it loses to a neighbour is lost for the batch, and its SM stays
// exclusive.  A slow clip that is ahead of the bulk (many waves of normal clips per warp slot) can share its SM: the bulk, not
// the clip, ends the batch.  "Critical" until the normal class has reported progress; never when there is no normal class to
// speak of (under an eighth of the clips: then the slow clips ARE the bulk and throughput is all that counts).
__device__ __forceinline__ bool no_normal_class(const int* q) {
  const long long nn = q[SQ_NNORM0] - ldv(q + SQ_LIVE_SLOW), ns = q[SQ_NSLOW0] + ldv(q + SQ_LIVE_SLOW);
  return 8 * nn < ns;
}
// `margin` (frames): the clip stays critical until it is that far AHEAD of the bulk - the decision is only revisited at
// segment boundaries, and a clip that falls behind between two of them ends the batch after the bulk.
__device__ __forceinline__ bool slow_clip_critical(const int* q, int t_done, int margin) {
  const long long fn = ldv(q + SQ_FR_NORM), nn = q[SQ_NNORM0] - ldv(q + SQ_LIVE_SLOW);
  if (no_normal_class(q)) return false;
  if (fn <= 0) return true;
  return (long long)(t_done - margin) * nn < fn;
}
// Sparse part of the grid: part_w warps on each of the first b_slow CTAs.  When the suspects do not fit into part_pct % of
// the SMs at part_w per SM, there is no sparse part at all (b_slow = 0): so many slow clips are not a tail to protect but a
// large share of the work, and every SM serves both rings at full occupancy, slow clips first.
__device__ __forceinline__ void sched_geometry(int n_s0, int part_w, int part_pct, int grid, int wpc, int* pw_out, int* b_slow_out) {
  const int pw = part_w < wpc ? part_w : wpc;
  const int max_blocks = grid * part_pct / 100, need = (n_s0 + pw - 1) / pw;
  *pw_out = pw;
  *b_slow_out = need <= max_blocks ? need : 0;
}
__device__ __forceinline__ int sq_pop(int* q, const long long* ring, int cap, int avail_i, int head_i, int n0) {
  if (n0 + ldv(q + avail_i) <= 0) return -1;                          // cheap look before touching the counter
  if (n0 + atomicSub(q + avail_i, 1) <= 0) { atomicAdd(q + avail_i, 1); return -1; }
  const int t = atomicAdd(q + head_i, 1);
  const volatile long long* e = ring + t % cap;
  long long v;
  while ((int)((v = *e) >> 32) != t + 1) { }
  __threadfence();                                                    // the clip's state was written before the entry
  return (int)(v & 0xffffffffll);
}
__device__ __forceinline__ void sq_push(int* q, long long* ring, int cap, int avail_i, int res_i, int n0, int c) {
  const int t = n0 + atomicAdd(q + res_i, 1);
  __threadfence();
  *reinterpret_cast<volatile long long*>(ring + t % cap) = ((long long)(t + 1) << 32) | (long long)(unsigned)c;
  __threadfence();
  atomicAdd(q + avail_i, 1);
}

// After frame 0 (solved for every clip by a first launch) a clip that will be slow is recognisable: its task error stays
// above 1 where a converged clip is at ~0.07-0.1 (from frame 1 on ANY threshold between 0.3 and 1.9 splits the benchmark
// clips identically; after frame 0 a quarter of the suspects still converge within a few frames and are handed back by the
// scheduler).  Fills the two rings of the scheduled launch.
__global__ void __launch_bounds__(256)
gmr_classify_kernel(const double* __restrict__ state, int stride, int nq, int C, int cap, double thresh,
                    int* __restrict__ counts, long long* __restrict__ ring) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (state[(size_t)c * stride + nq + 1] > thresh) { const int t = atomicAdd(&counts[0], 1); ring[t] = ((long long)(t + 1) << 32) | c; }
  else { const int t = atomicAdd(&counts[1], 1); ring[cap + t] = ((long long)(t + 1) << 32) | c; }
}

template <typename R, typename IO, int MAXWARPS>
__global__ void __launch_bounds__(MAXWARPS * 32, 1)
gmr_retarget_kernel(const __grid_constant__ GmrDims dm, const __grid_constant__ GmrScal<R> ks,
                    const GmrConsts<R>* __restrict__ gconsts, const __grid_constant__ GmrIO<IO> io, uint32_t kflags,
                    int* __restrict__ queue, const int* __restrict__ order, long long* __restrict__ rings, int part_w, int part_pct_margin,
                    int seg_frames, double slow_err) {
  const int part_pct = part_pct_margin & 0xff, crit_margin = part_pct_margin >> 8;     // (percentage | margin in frames << 8)
  extern __shared__ __align__(128) unsigned char gmr_dyn_smem[];
  unsigned char* const smem = gmr_dyn_smem;
  constexpr int CB = consts_bytes<R>();
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);                 // 16 bytes reserved for the mbarrier
  GmrConsts<R>* mc = reinterpret_cast<GmrConsts<R>*>(smem + 16);
  stage_consts_tma(mc, gconsts, CB, bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int wel = dm.warp_elems;
  WarpSolver<R> ws(*mc, dm, ks, (uint32_t)(16 + CB) + (uint32_t)warp * (uint32_t)wel * (uint32_t)sizeof(R), lane);
  ws.stream_init();
  ws.convoy = (kflags & GMR_FLAG_INTERNAL_CONVOY) != 0;
  const int C = io.C;
  ws.cta_active = reinterpret_cast<int*>(smem + 8);                  // second half of the mbarrier's 16-byte slot
  if (ws.convoy) {
    if (threadIdx.x == 0) *ws.cta_active = wpc;
    __syncthreads();
  }
  // ONE loop and one inlined copy of the solver (two call sites would double the kernel's code, which already exceeds
  // the instruction cache), serving both kinds of launch:
  //
  // Plain launch (part_w == 0): clips come from a global queue (one atomic per clip) - clips differ several-fold in the
  // number of IK steps they need, so a warp that finishes early takes the next clip instead of idling behind a static
  // assignment.  The first gridDim * wpc clips are handed out without touching the queue; with `order`
  // (gmr_order_kernel) the hard clips come first.
  //
  // Scheduled launch (part_w > 0; every clip already has a state: frame 0 was solved by a first launch and
  // gmr_classify_kernel split the clips by the task error it left).  Clips advance in SEGMENTS of seg_frames frames; between
  // segments a clip lives in its state record and in one of two FIFO rings:
  //   * slow clips (error still above slow_err after the last frame: stuck in a local minimum with joints on their limits,
  //     2-3x the IK steps of a normal clip and ~1.7 factorisations per step - they are the batch's critical path) run on
  //     "sparse" SMs, part_w warps per SM (2 per scheduler: ~13 % slower than a lone warp instead of ~60 % on a full SM),
  //     and keep their warp from segment to segment;
  //   * normal clips go round robin through the other ("dense", fully occupied) SMs: after every segment a clip goes to
  //     the back of the ring while others are waiting, so all of them finish within one segment of each other instead of
  //     in whole-clip waves;
  //   * a clip changes class when its error crosses slow_err (false suspects leave the sparse SMs after one segment);
  //   * when the normal ring runs dry, warps 0-3 of the dense SMs (one per scheduler) take slow clips, and slow clips on
  //     sparse SMs hand themselves over at their next segment boundary: the tail of the batch runs at lone-warp speed.
  // Pure scheduling: a clip's frames are solved in order from its own state, whichever warps run its segments.
  // Nothing of the scheduler stays in registers across a segment (the solver needs all of them): the warp's role, what it
  // is running and the clip id live in two spare words next to the warp's mbarrier and are re-read between segments.
  enum { W_SLOW_BLOCK = 1, W_STEALER = 2, W_TOOK_SLOW = 4, W_COUNTED_IDLE = 8, W_HAVE = 16, W_SPARE = 32, W_CRITICAL = 64 };
  volatile int* const wword = reinterpret_cast<volatile int*>(reinterpret_cast<unsigned char*>(ws.s_bar()) + 8);   // [0] role bits, [1] clip
  int* const n_res = reinterpret_cast<int*>(smem + 12);               // slow clips resident on this (sparse) SM
  int* const n_crit = reinterpret_cast<int*>(smem + 4);               // ... of which on the critical path (the constants' mbarrier is dead by now)
  if (part_w > 0) {
    if (threadIdx.x == 0) { *n_res = 0; *n_crit = 0; }
    __syncthreads();
    int pw, b_slow;
    sched_geometry(queue[SQ_NSLOW0], part_w, part_pct, (int)gridDim.x, wpc, &pw, &b_slow);
    const bool slow_block = (int)blockIdx.x < b_slow;
    if (lane == 0) wword[0] = slow_block ? (W_SLOW_BLOCK | (warp >= pw ? W_SPARE : 0))
                                         : ((warp < 4 && !(kflags & GMR_FLAG_INTERNAL_NOSTEAL)) ? W_STEALER : 0);
    __syncwarp();
  }
  int p = warp * (int)gridDim.x + (int)blockIdx.x;                    // plain launch: position in the clip order
  for (;;) {
    int c = -1;
    if (part_w == 0) {
      if (p >= C) break;
      const int n_hard = order ? queue[1] : 0;                        // queue = {next position, n_hard, n_easy, -}
      c = order ? (p < n_hard ? order[p] : order[C + p - n_hard]) : p;
    } else {
      int nap = 1000;
      if (lane == 0) {
        int w = wword[0];
        if (w & W_HAVE) c = wword[1];                                  // the clip keeps its warp for another segment
        else {
          const int cap = 2 * C + 2;
          const int n_s0 = queue[SQ_NSLOW0], n_n0 = queue[SQ_NNORM0];
          w &= ~W_TOOK_SLOW;
          if (w & W_SLOW_BLOCK) {
            // A sparse SM takes slow clips up to its fair share of the ones alive (on its first part_w warps) and nothing else
            // while it hosts any: a normal clip beside them would cost every slow clip on the SM more than it gains.  With no
            // slow clip left or waiting, the SM turns dense (all warps, normal ring).
            const bool waiting = n_s0 + ldv(queue + SQ_AVAIL_S) > 0;
            const int here = *reinterpret_cast<volatile int*>(n_res);
            if (!(w & W_SPARE) && waiting) {
              int pw, b_slow;
              sched_geometry(n_s0, part_w, part_pct, (int)gridDim.x, wpc, &pw, &b_slow);
              const int alive = n_s0 + ldv(queue + SQ_LIVE_SLOW);
              int target = (alive + b_slow - 1) / b_slow;
              if (target < 1) target = 1;
              if (here < target) {
                c = sq_pop(queue, rings, cap, SQ_AVAIL_S, SQ_HEAD_S, n_s0);
                if (c >= 0) { w |= W_TOOK_SLOW | W_CRITICAL; atomicAdd(n_res, 1); atomicAdd(n_crit, 1); }
              }
            }
            if (c < 0 && ((here == 0 && !waiting) || *reinterpret_cast<volatile int*>(n_crit) == 0))
              c = sq_pop(queue, rings + cap, cap, SQ_AVAIL_N, SQ_HEAD_N, n_n0);
            if (c < 0 && (w & W_SPARE)) nap = 20000;
          } else {
            int pw_, b_slow_;
            sched_geometry(n_s0, part_w, part_pct, (int)gridDim.x, wpc, &pw_, &b_slow_);
            if (b_slow_ == 0 && n_s0 > 0) {                            // no sparse part: slow clips first, on any warp
              c = sq_pop(queue, rings, cap, SQ_AVAIL_S, SQ_HEAD_S, n_s0); if (c >= 0) w |= W_TOOK_SLOW;
            }
            if (c < 0) c = sq_pop(queue, rings + cap, cap, SQ_AVAIL_N, SQ_HEAD_N, n_n0);
            // out of normal clips: one warp per scheduler takes a slow clip (lone-warp speed); every warp does when slow
            // clips are queueing for want of sparse capacity (more than one per sparse SM waiting) or are the bulk themselves
            if (c < 0 && ((w & W_STEALER) || n_s0 + ldv(queue + SQ_AVAIL_S) > (int)gridDim.x / 2 || no_normal_class(queue))) {
              c = sq_pop(queue, rings, cap, SQ_AVAIL_S, SQ_HEAD_S, n_s0); if (c >= 0) w |= W_TOOK_SLOW;
            }
          }
          if (c < 0) {
            if (ldv(queue + SQ_DONE) >= C) c = -2;
            else if ((w & W_STEALER) && !(w & W_COUNTED_IDLE)) { atomicAdd(queue + SQ_IDLE_STEAL, 1); w |= W_COUNTED_IDLE; }
          } else {
            if (w & W_COUNTED_IDLE) { atomicSub(queue + SQ_IDLE_STEAL, 1); w &= ~W_COUNTED_IDLE; }
            w |= W_HAVE;
          }
          wword[0] = w; wword[1] = c;
        }
      }
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c == -2) break;
      if (c < 0) { __nanosleep(__shfl_sync(0xffffffffu, nap, 0)); continue; }
      __threadfence();                                                 // the clip's state may have been written by another warp
    }
    const int nt = ws.template run_clip<IO>(io, c, part_w > 0 ? seg_frames : 0);
    if (part_w == 0) {
      int nxt = 0;
      if (lane == 0) nxt = gridDim.x * wpc + atomicAdd(queue, 1);
      p = __shfl_sync(0xffffffffu, nxt, 0);
      continue;
    }
    // scheduled launch: what happens to the clip after this segment
    __threadfence();                                                   // its state record, before anything is published
    __syncwarp();
    int act = 0;                                                       // 1: back off after handing a slow clip over
    if (lane == 0) {
      const int cc = wword[1];
      int w = wword[0];
      const bool took_slow = (w & W_TOOK_SLOW) != 0, resident = took_slow && (w & W_SLOW_BLOCK);
      if (!took_slow) atomicAdd(queue + SQ_FR_NORM, seg_frames);
      if (resident) {                                                  // refresh this clip's "critical" mark
        const bool crit = nt >= 0 && slow_clip_critical(queue, nt, crit_margin);
        if (crit != ((w & W_CRITICAL) != 0)) { atomicAdd(n_crit, crit ? 1 : -1); w ^= W_CRITICAL; }
      }
      int T_c = io.T;
      if (io.ex.lengths) { T_c = io.ex.lengths[cc]; T_c = T_c < 0 ? 0 : (T_c > io.T ? io.T : T_c); }
      if (nt < 0 || nt >= T_c) {                                       // complete (or stopped by a fatal status)
        atomicAdd(queue + SQ_DONE, 1);
        if (took_slow) atomicSub(queue + SQ_LIVE_SLOW, 1);
        if (resident) { atomicSub(n_res, 1); if (w & W_CRITICAL) atomicSub(n_crit, 1); }
        w &= ~(W_HAVE | W_CRITICAL);
      } else {
        const bool now_slow = (double)ws.s_piv()[1] > slow_err;
        const int n_s0 = queue[SQ_NSLOW0];
        bool keep;
        if (now_slow && took_slow) {
          // a slow clip keeps its warp, unless it sits on a sparse SM that holds more than its fair share of the slow clips
          // alive, or lone-speed slots are idle elsewhere
          keep = true;
          if (resident) {
            int pw, b_slow;
            sched_geometry(n_s0, part_w, part_pct, (int)gridDim.x, wpc, &pw, &b_slow);
            const int alive = n_s0 + ldv(queue + SQ_LIVE_SLOW);
            int target = (alive + b_slow - 1) / b_slow;
            if (target < 1) target = 1;
            keep = !((ldv(queue + SQ_IDLE_STEAL) > 0 && !no_normal_class(queue)) || *reinterpret_cast<volatile int*>(n_res) > target);
          }
        } else if (!now_slow && !took_slow) {
          // a normal clip keeps its warp while nobody is waiting in the normal ring - except on a sparse SM that has slow work again
          keep = queue[SQ_NNORM0] + ldv(queue + SQ_AVAIL_N) <= 0 &&
                 !((w & W_SLOW_BLOCK) && (*reinterpret_cast<volatile int*>(n_crit) > 0 || n_s0 + ldv(queue + SQ_AVAIL_S) > 0));
        } else keep = false;                                           // the clip changes class
        if (!keep) {
          const int cap = 2 * C + 2;
          if (resident) { atomicSub(n_res, 1); if (w & W_CRITICAL) atomicSub(n_crit, 1); }
          w &= ~W_CRITICAL;
          if (now_slow != took_slow) atomicAdd(queue + SQ_LIVE_SLOW, now_slow ? 1 : -1);
          if (now_slow) sq_push(queue, rings, cap, SQ_AVAIL_S, SQ_RES_S, n_s0, cc);
          else sq_push(queue, rings + cap, cap, SQ_AVAIL_N, SQ_RES_N, queue[SQ_NNORM0], cc);
          w &= ~W_HAVE;
          act = now_slow && took_slow;
        }
      }
      wword[0] = w;
    }
    if (__shfl_sync(0xffffffffu, act, 0)) __nanosleep(4000);           // let the slot that should take it win the pop
  }
  ws.convoy_retire();
}

// Motion-file epilogue (gmr_finalize_motion): one thread per (clip, frame, qpos element).  HBM-bound elementwise
// pass: reads qpos once (4 nq bytes per frame, coalesced), writes the three arrays of the pkl once.
__global__ void __launch_bounds__(256)
gmr_finalize_kernel(const float* __restrict__ qpos, const float* __restrict__ lowest, const int32_t* __restrict__ lengths,
                    int C, int T, int nq, int height_adjust, int origin_offset,
                    float* __restrict__ root_pos, float* __restrict__ root_rot, float* __restrict__ dof_pos) {
  const size_t n = (size_t)C * T * nq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t f = i / nq;                 // frame index c * T + t
    const int e = (int)(i - f * nq);
    const int c = (int)(f / T), t = (int)(f - (size_t)c * T);
    const bool live = !lengths || t < lengths[c];
    float v = live ? qpos[i] : 0.0f;
    if (e < 3) {
      if (live) {
        if (e == 2) { if (height_adjust && lowest) v -= lowest[c]; }
        else if (origin_offset) v -= qpos[(size_t)c * T * nq + e];          // first frame of the clip
      }
      root_pos[f * 3 + e] = v;
    } else if (e < 7) {
      root_rot[f * 4 + (e == 3 ? 3 : e - 4)] = v;                           // wxyz -> xyzw
    } else {
      dof_pos[f * (nq - 7) + (e - 7)] = v;
    }
  }
}

// Mixed-robot batches (BASELINE.json configs[4]): ONE launch serves up to GMR_MAX_MULTI robot-uniform buckets.  The
// CTAs are divided among the buckets in proportion to their work; a CTA stages its bucket's constant block and
// copies the bucket's dims / scalars / batch pointers from the kernel parameters into shared memory (they cannot
// be compile-time constant-bank operands here, which costs a few per cent per solve), then serves its bucket's
// clip queue.  Running the buckets as separate launches instead makes their tails add up (one CTA per SM).
#define GMR_MAX_MULTI 8
template <typename R, typename IO> struct GmrMultiArgs {
  int32_t n, _pad[3];
  int32_t cta_end[GMR_MAX_MULTI];                  // CTAs [cta_end[r-1], cta_end[r]) serve bucket r
  GmrDims dm[GMR_MAX_MULTI];
  GmrScal<R> ks[GMR_MAX_MULTI];
  const GmrConsts<R>* gc[GMR_MAX_MULTI];
  GmrIO<IO> io[GMR_MAX_MULTI];
  int* queue[GMR_MAX_MULTI];                       // {next position, n_hard, n_easy, -}
  const int* order[GMR_MAX_MULTI];                 // hard-first permutation or null
};
struct MultiLocal { GmrDims dm; int32_t r, cta_begin, cta_count, _pad; };

template <typename R, typename IO, int MAXWARPS>
__global__ void __launch_bounds__(MAXWARPS * 32, 1)
gmr_retarget_multi_kernel(const __grid_constant__ GmrMultiArgs<R, IO> mu) {
  extern __shared__ __align__(128) unsigned char gmr_dyn_smem[];
  unsigned char* const smem = gmr_dyn_smem;
  constexpr int CB = consts_bytes<R>();
  constexpr int LB = (int)((sizeof(MultiLocal) + sizeof(GmrScal<R>) + sizeof(GmrIO<IO>) + 15) / 16 * 16);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  GmrConsts<R>* mc = reinterpret_cast<GmrConsts<R>*>(smem + 16);
  MultiLocal* loc = reinterpret_cast<MultiLocal*>(smem + 16 + CB);
  GmrScal<R>* ks = reinterpret_cast<GmrScal<R>*>(smem + 16 + CB + sizeof(MultiLocal));
  GmrIO<IO>* io = reinterpret_cast<GmrIO<IO>*>(smem + 16 + CB + sizeof(MultiLocal) + sizeof(GmrScal<R>));
  int r = 0;
  while (r + 1 < mu.n && (int)blockIdx.x >= mu.cta_end[r]) r++;
  if (threadIdx.x == 0) {
    loc->dm = mu.dm[r]; loc->r = r; loc->cta_begin = r ? mu.cta_end[r - 1] : 0; loc->cta_count = mu.cta_end[r] - loc->cta_begin;
    *ks = mu.ks[r]; *io = mu.io[r];
  }
  stage_consts_tma(mc, mu.gc[r], CB, bar);                             // contains the __syncthreads that publishes the copies above
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const GmrDims& dm = loc->dm;
  WarpSolver<R> ws(*mc, dm, *ks, (uint32_t)(16 + CB + LB) + (uint32_t)warp * (uint32_t)dm.warp_elems * (uint32_t)sizeof(R), lane);
  ws.stream_init();
  const int C = io->C, lb = (int)blockIdx.x - loc->cta_begin, lg = loc->cta_count;
  int* queue = mu.queue[r];
  const int* order = mu.order[r];
  const int nw = lg * wpc;
  const int n_hard = order ? queue[1] : 0;
  for (int p = warp * lg + lb; p < C;) {
    const int c = order ? (p < n_hard ? order[p] : order[C + p - n_hard]) : p;
    ws.template run_clip<IO>(*io, c);
    int nxt = 0;
    if (lane == 0) nxt = nw + atomicAdd(queue, 1);
    p = __shfl_sync(0xffffffffu, nxt, 0);
  }
}

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
  // Stream-ordered scratch (clip-queue counters, clip permutations, the two-phase hand-over state) comes from the
  // model's OWN memory pool with an unlimited release threshold: the device's default pool gives freed memory back to
  // the OS at every synchronisation, and a caller that synchronises after each call (e.g. to read the status words)
  // then pays a fresh OS allocation per call — measured: 80 ms calls turning into 0.5-1 s ones.
  cudaMemPool_t pool = nullptr;
  // lazily created resources of the host-buffer entry
  std::mutex host_mu;
  cudaStream_t hs[2] = {nullptr, nullptr};
  cudaStream_t ms[8] = {};       // side streams of gmr_retarget_multi (created on first use, guarded by host_mu)
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

// stream-ordered scratch memory from the model's pool, released in stream order when the scope ends: every launch
// in flight owns its counters and lists, whatever the number of streams or the depth of the queue
struct StreamScratch {
  cudaMemPool_t pool; cudaStream_t st; char* p = nullptr;
  StreamScratch(cudaMemPool_t pl, cudaStream_t s) : pool(pl), st(s) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocFromPoolAsync(reinterpret_cast<void**>(&p), bytes, pool, st); }
  ~StreamScratch() { if (p) { cudaFreeAsync(p, st); } }
  StreamScratch(const StreamScratch&) = delete;
  StreamScratch& operator=(const StreamScratch&) = delete;
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
template <typename R> int pick_wpc(const GmrModel* m, int C, int sms = 0) {
  if (sms <= 0) sms = m->num_sms;
  int wpc = max_warps<R>() / ctas_per_sm();
  if (const char* cap = getenv("GMR_WPC_CAP")) { int c = atoi(cap); if (c >= 1 && c < wpc) wpc = c; }   // tuning knob
  while (wpc > 1 && smem_bytes<R>(m, wpc) > (size_t)m->max_smem) wpc--;
  int need = (C + sms * ctas_per_sm() - 1) / (sms * ctas_per_sm());
  if (need < 1) need = 1;
  if (need < wpc && !getenv("GMR_WPC_FULL")) wpc = need;                 // GMR_WPC_FULL: experiment knob, always the full CTA
  return wpc;
}

template <typename R, typename IO, int MAXWARPS>
int launch(GmrModel* m, const GmrConsts<R>* dc, const float* pos, const float* quat, const float* ratio, int C, int T,
           const IO* qinit, IO* qout, int32_t* iters, IO* err, IO* tg, uint32_t flags, cudaStream_t st,
           const GmrBatchExtra* extra = nullptr, int* own_queue = nullptr, int sm_share = 0) {
  // sm_share > 0: this launch may use only that many SMs (CTAs); other launches run beside it (gmr_retarget_multi)
  if (C == 0 || T == 0) return GMR_OK;
  const int sms = sm_share > 0 && sm_share < m->num_sms ? sm_share : m->num_sms;
  GmrBatchExtra ex{};
  if (extra) ex = *extra;
  auto kern = gmr_retarget_kernel<R, IO, MAXWARPS>;
  const int wpc = pick_wpc<R>(m, C, sms);
  const size_t smem = smem_bytes<R>(m, wpc);
  if (smem > (size_t)m->max_smem) return set_err(GMR_ELIMIT, "model does not fit in shared memory");
  // opt in to the large dynamic allocation: once per (kernel instantiation, device) — the attribute is per device
  static std::atomic<uint64_t> configured{0};
  const uint64_t dev_bit = 1ull << (m->device & 63);
  if (!(configured.load(std::memory_order_acquire) & dev_bit)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m->max_smem));
    configured.fetch_or(dev_bit, std::memory_order_release);
  }
  int grid = (C + wpc - 1) / wpc;
  if (grid > sms * ctas_per_sm()) grid = sms * ctas_per_sm();
  // convoy mode (see gmr_solver.cuh) paid off while the hot path was 3x the instruction cache; with the
  // branch-sparse factorisation free-running warps are as fast on balanced batches and 8-10 % faster on
  // mixed ones (a slow clip no longer drags its CTA through a rendezvous per factorisation).  GMR_CONVOY=1 re-enables.
  static const int cv_env = getenv("GMR_CONVOY") ? atoi(getenv("GMR_CONVOY")) : 0;
  const bool convoy = cv_env != 0 && wpc >= 2;
  static const int nosteal_env = getenv("GMR_NO_STEAL") ? atoi(getenv("GMR_NO_STEAL")) : 0;
  static const int force_sched_env = getenv("GMR_FORCE_SCHED") ? atoi(getenv("GMR_FORCE_SCHED")) : 0;
  flags = (flags & 0xffffu) | (convoy ? GMR_FLAG_INTERNAL_CONVOY : 0u) | (nosteal_env ? GMR_FLAG_INTERNAL_NOSTEAL : 0u);
  GmrIO<IO> io{};
  io.pos = pos; io.quat = quat; io.ratio = ratio; io.qinit = qinit; io.qout = qout; io.iters = iters; io.err = err; io.tg = tg;
  io.C = C; io.T = T; io.flags = flags & 0xffffu; io.ex = ex; io.trace = g_trace.load();
  const GmrDims& dims = sizeof(R) == 4 ? m->dims32 : m->dims64;
  const int slots = grid * wpc;
  static const int lpt_env = getenv("GMR_LPT") ? atoi(getenv("GMR_LPT")) : 1;
  // warps per slow SM (0 = off) and share of the SMs the slow clips may take: measured optima on the benchmark mix
  // (tools/prof/mix_case.py, ~10 % of the clips classified slow).  float32: 4 warps / 55 % (64 ms vs 70 ms at 40 %,
  // worse beyond 65 %).  float64 has 16 instead of 28 warps per SM to give away, so sparse slow SMs cost it more
  // capacity: 6 warps / 45 % (82 ms vs 86 ms at 4 / 55 % and 88-90 ms at 8-10 warps / 30-35 %).
  static const int part_env = getenv("GMR_PARTITION") ? atoi(getenv("GMR_PARTITION")) : 8;

  // ---- more clips than the sparse part could hold: frame 0 for every clip, classify, then the scheduled launch ----
  // (see gmr_retarget_kernel).  pw = warps per sparse SM: 2 per scheduler keeps a slow clip within ~13 % of the speed of a
  // lone warp (tools/prof/slow_curve.py: 21.5 / 24.4 / 27.9 / 34.3 us per IK step at 1-4 / 8 / 10 / 16 warps per SM) at
  // twice the capacity of 1 per scheduler; the number of sparse SMs follows from the number of suspects.  Segment
  // length: 25 frames = 2-3 ms per visit, thousands of IK steps per ring operation.
  static const int seg_env = getenv("GMR_SEGMENT") ? atoi(getenv("GMR_SEGMENT")) : 25;
  static const double slow_err_env = getenv("GMR_SLOW_ERR") ? atof(getenv("GMR_SLOW_ERR")) : 1.0;
  const bool two_phase = part_env > 0 && seg_env > 0 && !own_queue && !convoy && !(flags & GMR_FLAG_NO_SOLVE) && T >= 16 &&
                         (C > 2 * sms * part_env || force_sched_env);
  if (two_phase) {
    const int stride = gmr_state_stride(dims.nq), cap = 2 * C + 2;
    const size_t b_state = ((size_t)C * stride * sizeof(double) + 255) & ~(size_t)255, b_ring = (size_t)2 * cap * sizeof(long long);
    StreamScratch scratch(m->pool, st);                      // freed in stream order on every exit path
    CK(scratch.alloc(256 + b_state + b_ring));               // [ launch A's clip queue, the scheduler's counters | state | rings ]
    int* q1 = reinterpret_cast<int*>(scratch.p);
    int* q2 = q1 + 16;
    double* state = reinterpret_cast<double*>(scratch.p + 256);
    long long* rings = reinterpret_cast<long long*>(scratch.p + 256 + b_state);
    CK(cudaMemsetAsync(q1, 0, 256, st));
    // launch A: frame 0 of every clip (balanced work: every clip starts from the same configuration)
    GmrIO<IO> ioa = io;
    ioa.t_begin = 0; ioa.t_end = 1; ioa.state = state;
    kern<<<grid, wpc * 32, smem, st>>>(dims, scal_of<R>(m), dc, ioa, flags, q1, (const int*)nullptr, (long long*)nullptr, 0, 0, 0, 0.0);
    CK(cudaGetLastError());
    gmr_classify_kernel<<<(C + 255) / 256, 256, 0, st>>>(state, stride, dims.nq, C, cap, slow_err_env, q2 + SQ_NSLOW0, rings);
    CK(cudaGetLastError());
    // launch B: frames 1 .. T-1, scheduled
    GmrIO<IO> iob = io;
    iob.t_begin = 1; iob.t_end = 0; iob.state = state;
    static const int pct_env = getenv("GMR_PARTITION_PCT") ? atoi(getenv("GMR_PARTITION_PCT")) : 50;
    // a slow clip stays critical until it is one segment ahead of the bulk (the decision is revisited once per segment)
    static const int margin_env = getenv("GMR_CRIT_MARGIN") ? atoi(getenv("GMR_CRIT_MARGIN")) : -1;
    const int margin = margin_env >= 0 ? (margin_env > 0xffff ? 0xffff : margin_env) : seg_env;
    kern<<<grid, wpc * 32, smem, st>>>(dims, scal_of<R>(m), dc, iob, flags, q2, (const int*)nullptr, rings, part_env, pct_env | (margin << 8), seg_env, slow_err_env);
    CK(cudaGetLastError());
    g_launches.fetch_add(3);
    return GMR_OK;
  }

  // more clips than warp slots: hard clips first (stream-ordered scratch for the clip queue and the permutation)
  const bool lpt = lpt_env && C > slots;
  StreamScratch scratch(m->pool, st);
  if (!own_queue || lpt) CK(scratch.alloc(256 + (lpt ? (size_t)2 * C * sizeof(int) : 0)));
  int* queue = own_queue ? own_queue : reinterpret_cast<int*>(scratch.p);
  CK(cudaMemsetAsync(queue, 0, 4 * sizeof(int), st));
  int* order = nullptr;
  if (lpt) {
    order = reinterpret_cast<int*>(scratch.p + 256);
    gmr_order_kernel<R, IO><<<(C + 255) / 256, 256, 0, st>>>(dc, quat, qinit, C, T, queue + 1, order);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
  }
  kern<<<grid, wpc * 32, smem, st>>>(dims, scal_of<R>(m), dc, io, flags, queue, order, (long long*)nullptr, 0, 0, 0, 0.0);
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
  if (e == cudaSuccess) {
    cudaMemPoolProps pp{};
    pp.allocType = cudaMemAllocationTypePinned; pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice; pp.location.id = device;
    e = cudaMemPoolCreate(&m->pool, &pp);
    if (e == cudaSuccess) { uint64_t keep = UINT64_MAX; e = cudaMemPoolSetAttribute(m->pool, cudaMemPoolAttrReleaseThreshold, &keep); }
  }
  if (e == cudaSuccess) e = cudaMemset(m->d_f32, 0, consts_bytes<float>());
  if (e == cudaSuccess) e = cudaMemset(m->d_f64, 0, consts_bytes<double>());
  if (e == cudaSuccess) e = cudaMemcpy(m->d_f32, &m->h_f32, sizeof(GmrConsts<float>), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_f64, hd, sizeof(GmrConsts<double>), cudaMemcpyHostToDevice);
  delete hd;
  if (e != cudaSuccess) {
    int code = cuda_err(e, "gmr_model_create");
    if (m->d_f32) cudaFree(m->d_f32);
    if (m->d_f64) cudaFree(m->d_f64);
    if (m->pool) cudaMemPoolDestroy(m->pool);
    delete m;
    return code;
  }
  if (smem_bytes<float>(m, 1) > (size_t)m->max_smem) {
    cudaFree(m->d_f32); cudaFree(m->d_f64); cudaMemPoolDestroy(m->pool); delete m;
    return set_err(GMR_ELIMIT, "model does not fit in shared memory");
  }
  *out = m;
  return GMR_OK;
}

int gmr_model_destroy(GmrModel* m) {
  if (!m) return GMR_OK;
  DeviceGuard g(m->device);
  for (auto& st : m->ms) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (int i = 0; i < 2; i++) {
    if (m->hs[i]) { cudaStreamSynchronize(m->hs[i]); cudaStreamDestroy(m->hs[i]); }
  }
  cudaFree(m->d_f32);
  cudaFree(m->d_f64);
  if (m->pool) cudaMemPoolDestroy(m->pool);      // outstanding stream-ordered frees complete first (CUDA defers the release)
  delete m;
  return GMR_OK;
}

static int check_batch_args(GmrModel* m, const void* pos, const void* quat, const void* qpos_out, int32_t C, int32_t T) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if ((C > 0 && T > 0) && (!pos || !quat || !qpos_out)) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  if ((reinterpret_cast<uintptr_t>(quat) & 15u) != 0) return set_err(GMR_EINVAL, "quat must be 16-byte aligned");
  return GMR_OK;
}

int gmr_retarget_batch_ex(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                          const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out, float* targets_out,
                          const GmrBatchExtra* extra, uint32_t flags, void* cuda_stream) {
  if (int rc = check_batch_args(m, pos, quat, qpos_out, C, T)) return rc;
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  if (flags & GMR_FLAG_COMPUTE_F64)
    return launch<double, float, MAXW_F64>(m, m->d_f64, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                           targets_out, flags, (cudaStream_t)cuda_stream, extra);
  return launch<float, float, MAXW_F32>(m, m->d_f32, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                        targets_out, flags, (cudaStream_t)cuda_stream, extra);
}

int gmr_retarget_batch(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                       const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out, float* targets_out,
                       uint32_t flags, void* cuda_stream) {
  return gmr_retarget_batch_ex(m, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, targets_out, nullptr, flags, cuda_stream);
}

int gmr_retarget_batch_f64_ex(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                              const double* qpos_init, double* qpos_out, int32_t* iters_out, double* err_out,
                              double* targets_out, const GmrBatchExtra* extra, uint32_t flags, void* cuda_stream) {
  if (int rc = check_batch_args(m, pos, quat, qpos_out, C, T)) return rc;
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  return launch<double, double, MAXW_F64>(m, m->d_f64, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out,
                                          targets_out, flags, (cudaStream_t)cuda_stream, extra);
}

int gmr_retarget_batch_f64(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                           const double* qpos_init, double* qpos_out, int32_t* iters_out, double* err_out,
                           double* targets_out, uint32_t flags, void* cuda_stream) {
  return gmr_retarget_batch_f64_ex(m, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, targets_out, nullptr, flags, cuda_stream);
}

int gmr_finalize_motion(GmrModel* m, const float* qpos, const float* lowest_z, const int32_t* lengths, int32_t C, int32_t T,
                        int32_t height_adjust, int32_t origin_offset, float* root_pos_out, float* root_rot_xyzw_out,
                        float* dof_pos_out, void* cuda_stream) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if (C == 0 || T == 0) return GMR_OK;
  if (!qpos || !root_pos_out || !root_rot_xyzw_out || (m->h_f32.nh > 0 && !dof_pos_out))
    return set_err(GMR_EINVAL, "qpos and the three output arrays are required");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  const size_t n = (size_t)C * T * m->h_f32.nq;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)m->num_sms * 16;             // grid-stride: a multiple of the SM count
  if (blocks > cap) blocks = cap;
  gmr_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(
      qpos, lowest_z, lengths, C, T, m->h_f32.nq, height_adjust, origin_offset, root_pos_out, root_rot_xyzw_out, dof_pos_out);
  g_launches.fetch_add(1);
  CK(cudaGetLastError());
  return GMR_OK;
}

}  // extern "C"

namespace {
template <typename R, int MAXWARPS>
int launch_multi(const GmrBatchDesc* b, int n, uint32_t flags, cudaStream_t st) {
  GmrModel* m0 = b[0].model;
  auto kern = gmr_retarget_multi_kernel<R, float, MAXWARPS>;
  constexpr int LB = (int)((sizeof(MultiLocal) + sizeof(GmrScal<R>) + sizeof(GmrIO<float>) + 15) / 16 * 16);
  // warps per CTA: what the largest robot's state allows; CTAs: the device's SMs, split by work (clips x frames)
  int wel = 0;
  double work_total = 0;
  for (int i = 0; i < n; i++) {
    const int w = sizeof(R) == 4 ? b[i].model->wel32 : b[i].model->wel64;
    if (w > wel) wel = w;
    work_total += (double)b[i].C * b[i].T;
  }
  int wpc = max_warps<R>();
  while (wpc > 1 && 16 + (size_t)consts_bytes<R>() + LB + (size_t)wpc * wel * sizeof(R) > (size_t)m0->max_smem) wpc--;
  const size_t smem = 16 + (size_t)consts_bytes<R>() + LB + (size_t)wpc * wel * sizeof(R);
  if (smem > (size_t)m0->max_smem) return set_err(GMR_ELIMIT, "models do not fit in shared memory");
  static std::atomic<uint64_t> configured{0};
  const uint64_t dev_bit = 1ull << (m0->device & 63);
  if (!(configured.load(std::memory_order_acquire) & dev_bit)) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m0->max_smem));
    configured.fetch_or(dev_bit, std::memory_order_release);
  }
  const int grid = m0->num_sms;
  if (grid < n) return set_err(GMR_ELIMIT, "more buckets than SMs");
  GmrMultiArgs<R, float> mu{};
  mu.n = n;
  StreamScratch scratch(m0->pool, st);
  size_t order_ints = 0;
  for (int i = 0; i < n; i++) order_ints += (size_t)2 * b[i].C;
  CK(scratch.alloc(256 + order_ints * sizeof(int)));       // [ one clip queue per bucket | lists ]
  CK(cudaMemsetAsync(scratch.p, 0, 256, st));
  int* order_base = reinterpret_cast<int*>(scratch.p + 256);
  int assigned = 0;
  double acc = 0;
  for (int i = 0; i < n; i++) {
    GmrModel* m = b[i].model;
    acc += (double)b[i].C * b[i].T;
    int end = i + 1 == n ? grid : (int)(grid * (acc / work_total) + 0.5);
    if (end < assigned + 1) end = assigned + 1;                       // at least one CTA per bucket
    if (end > grid - (n - 1 - i)) end = grid - (n - 1 - i);
    mu.cta_end[i] = end;
    assigned = end;
    mu.dm[i] = sizeof(R) == 4 ? m->dims32 : m->dims64;
    mu.ks[i] = scal_of<R>(m);
    mu.gc[i] = sizeof(R) == 4 ? reinterpret_cast<const GmrConsts<R>*>(m->d_f32) : reinterpret_cast<const GmrConsts<R>*>(m->d_f64);
    GmrIO<float>& io = mu.io[i];
    io.pos = b[i].pos; io.quat = b[i].quat; io.ratio = b[i].ratio; io.qinit = b[i].qpos_init; io.qout = b[i].qpos_out;
    io.iters = b[i].iters_out; io.err = b[i].err_out; io.tg = nullptr; io.C = b[i].C; io.T = b[i].T; io.flags = flags & 0xffffu;
    int* q = reinterpret_cast<int*>(scratch.p) + 4 * i;
    mu.queue[i] = q;
    // hard clips first within each bucket (same pre-pass as the single-robot path)
    gmr_order_kernel<R, float><<<(b[i].C + 255) / 256, 256, 0, st>>>(mu.gc[i], b[i].quat, b[i].qpos_init, b[i].C, b[i].T, q + 1, order_base);
    CK(cudaGetLastError());
    mu.order[i] = order_base;
    order_base += (size_t)2 * b[i].C;
    g_launches.fetch_add(1);
  }
  kern<<<grid, wpc * 32, smem, st>>>(mu);
  g_launches.fetch_add(1);
  CK(cudaGetLastError());
  return GMR_OK;
}
}  // namespace

extern "C" {

int gmr_retarget_multi(const GmrBatchDesc* batches, int32_t n, uint32_t flags, void* cuda_stream) {
  if (n < 0 || (n > 0 && !batches)) return set_err(GMR_EINVAL, "bad bucket list");
  if (n == 0) return GMR_OK;
  if (n > GMR_MAX_MULTI) return set_err(GMR_ELIMIT, "more than 8 buckets in one launch");
  if (flags & GMR_FLAG_NO_SOLVE) return set_err(GMR_EINVAL, "GMR_FLAG_NO_SOLVE is not supported by gmr_retarget_multi");
  for (int i = 0; i < n; i++) {
    const GmrBatchDesc& d = batches[i];
    if (!d.model) return set_err(GMR_EINVAL, "model is null");
    if (d.model->device != batches[0].model->device) return set_err(GMR_EINVAL, "all buckets must live on one device");
    if (d.C <= 0 || d.T <= 0) return set_err(GMR_EINVAL, "empty bucket");
    if (!d.pos || !d.quat || !d.qpos_out) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
    if ((reinterpret_cast<uintptr_t>(d.quat) & 15u) != 0) return set_err(GMR_EINVAL, "quat must be 16-byte aligned");
  }
  GmrModel* m0 = batches[0].model;
  DeviceGuard g(m0->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the models' device");
  static const int one_kernel = getenv("GMR_MULTI_KERNEL") ? atoi(getenv("GMR_MULTI_KERNEL")) : 0;
  if (one_kernel || m0->num_sms < 2 * n) {
    if (flags & GMR_FLAG_COMPUTE_F64) return launch_multi<double, MAXW_F64>(batches, n, flags, (cudaStream_t)cuda_stream);
    return launch_multi<float, MAXW_F32>(batches, n, flags, (cudaStream_t)cuda_stream);
  }
  // Default: every bucket is an ordinary single-robot batch (two-phase / scheduled launch included) confined to its share
  // of the SMs, all buckets side by side on the first model's side streams, joined back into the caller's stream.  The
  // shares follow the work: clips x frames x DoFs (the per-frame cost grows with the size of the system).
  cudaStream_t st = (cudaStream_t)cuda_stream;
  double work[GMR_MAX_MULTI], total = 0;
  for (int i = 0; i < n; i++) { work[i] = (double)batches[i].C * batches[i].T * (6 + batches[i].model->h_f32.nh); total += work[i]; }
  int share[GMR_MAX_MULTI], used = 0;
  for (int i = 0; i < n; i++) { share[i] = (int)(m0->num_sms * work[i] / total); if (share[i] < 1) share[i] = 1; used += share[i]; }
  for (int i = 0; used < m0->num_sms; i = (i + 1) % n) { share[i]++; used++; }       // hand out the remainder
  for (int i = 0; used > m0->num_sms; i = (i + 1) % n) if (share[i] > 1) { share[i]--; used--; }
  std::lock_guard<std::mutex> lk(m0->host_mu);
  cudaEvent_t fork = nullptr, join[GMR_MAX_MULTI] = {};
  int rc = GMR_OK;
  auto cleanup = [&]() { if (fork) cudaEventDestroy(fork); for (auto& e : join) if (e) cudaEventDestroy(e); };
  #define CKM(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return cuda_err(_e, #call); } } while (0)
  CKM(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  CKM(cudaEventRecord(fork, st));
  for (int i = 0; i < n && rc == GMR_OK; i++) {
    if (!m0->ms[i]) CKM(cudaStreamCreateWithFlags(&m0->ms[i], cudaStreamNonBlocking));
    cudaStream_t si = m0->ms[i];
    CKM(cudaStreamWaitEvent(si, fork, 0));
    const GmrBatchDesc& d = batches[i];
    if (flags & GMR_FLAG_COMPUTE_F64)
      rc = launch<double, float, MAXW_F64>(d.model, d.model->d_f64, d.pos, d.quat, d.ratio, d.C, d.T, d.qpos_init, d.qpos_out,
                                           d.iters_out, d.err_out, nullptr, flags, si, nullptr, nullptr, share[i]);
    else
      rc = launch<float, float, MAXW_F32>(d.model, d.model->d_f32, d.pos, d.quat, d.ratio, d.C, d.T, d.qpos_init, d.qpos_out,
                                          d.iters_out, d.err_out, nullptr, flags, si, nullptr, nullptr, share[i]);
    if (rc != GMR_OK) break;
    CKM(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
    CKM(cudaEventRecord(join[i], si));
    CKM(cudaStreamWaitEvent(st, join[i], 0));
  }
  #undef CKM
  cleanup();                                                       // destruction is deferred until the events have completed
  return rc;
}

// ---- single live stream ---------------------------------------------------------------------------------
}  // extern "C"

struct GmrStream {
  GmrModel* m = nullptr;
  cudaStream_t st = nullptr;
  char* h_in = nullptr;    // pinned: pos [nh,3] f32 | quat [nh,4] f32 (16-byte aligned) | ratio f32
  char* h_out = nullptr;   // pinned: qpos [nq] f64 | err [2] f64 | targets [nh,7] f64 | iters [2] i32
  char* d_buf = nullptr;   // device: the same two blocks, then the state: qpos [nq] f64 | working sets [4] u32
  size_t in_bytes = 0, out_bytes = 0, o_quat = 0, o_ratio = 0, o_err = 0, o_tg = 0, o_it = 0, o_state = 0, o_warm = 0;
  cudaGraphExec_t graph[4] = {nullptr, nullptr, nullptr, nullptr};      // by (flags & 3)
  bool graph_failed = false;
  std::mutex mu;
};

namespace {
inline size_t up16(size_t b) { return (b + 15) & ~(size_t)15; }

// copy in -> solve (one clip, one frame, float64) -> keep the new configuration -> copy out, on s->st
int stream_enqueue(GmrStream* s, uint32_t flags) {
  GmrModel* m = s->m;
  char* d = s->d_buf;
  CK(cudaMemcpyAsync(d, s->h_in, s->in_bytes, cudaMemcpyHostToDevice, s->st));
  char* dout = d + up16(s->in_bytes);
  GmrBatchExtra ex{};
  ex.warm_state = (flags & GMR_FLAG_NO_SOLVE) ? nullptr : reinterpret_cast<uint32_t*>(d + s->o_warm);
  int rc = launch<double, double, MAXW_F64>(m, m->d_f64, (const float*)d, (const float*)(d + s->o_quat), (const float*)(d + s->o_ratio), 1, 1,
                                            (const double*)(d + s->o_state), (double*)dout, (int32_t*)(dout + s->o_it),
                                            (double*)(dout + s->o_err), (double*)(dout + s->o_tg), flags & 3u, s->st, &ex,
                                            reinterpret_cast<int*>(d + s->o_warm + 16));     // {position, n_hard, n_easy}: 12 bytes
  if (rc != GMR_OK) return rc;
  if (!(flags & GMR_FLAG_NO_SOLVE))
    CK(cudaMemcpyAsync(d + s->o_state, dout, (size_t)m->h_f32.nq * 8, cudaMemcpyDeviceToDevice, s->st));
  CK(cudaMemcpyAsync(s->h_out, dout, s->out_bytes, cudaMemcpyDeviceToHost, s->st));
  return GMR_OK;
}
}  // namespace

extern "C" {

int gmr_stream_create(GmrModel* m, double height_ratio, GmrStream** out) {
  if (!out) return set_err(GMR_EINVAL, "out is null");
  *out = nullptr;
  if (!m) return set_err(GMR_EINVAL, "model is null");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  GmrStream* s = new (std::nothrow) GmrStream();
  if (!s) return set_err(GMR_ENOMEM, "host allocation failed");
  s->m = m;
  const size_t nh = m->h_f32.nhum, nq = m->h_f32.nq;
  s->o_quat = up16(nh * 3 * 4); s->o_ratio = s->o_quat + nh * 4 * 4; s->in_bytes = s->o_ratio + 16;
  s->o_err = up16(nq * 8); s->o_tg = s->o_err + 16; s->o_it = s->o_tg + up16(nh * 7 * 8); s->out_bytes = s->o_it + 16;
  s->o_state = up16(s->in_bytes) + up16(s->out_bytes); s->o_warm = s->o_state + up16(nq * 8);
  cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMallocHost(&s->h_in, s->in_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&s->h_out, s->out_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_buf, s->o_warm + 32);          // + working sets + this stream's own clip-queue counter
  if (e == cudaSuccess) e = cudaMemset(s->d_buf, 0, s->o_warm + 32);
  if (e != cudaSuccess) { int code = cuda_err(e, "gmr_stream_create"); gmr_stream_destroy(s); return code; }
  *reinterpret_cast<float*>(s->h_in + s->o_ratio) = (float)height_ratio;
  int rc = gmr_stream_reset(s, nullptr);
  if (rc != GMR_OK) { gmr_stream_destroy(s); return rc; }
  *out = s;
  return GMR_OK;
}

int gmr_stream_destroy(GmrStream* s) {
  if (!s) return GMR_OK;
  DeviceGuard g(s->m->device);
  if (s->st) cudaStreamSynchronize(s->st);
  for (auto& ge : s->graph) if (ge) cudaGraphExecDestroy(ge);
  if (s->d_buf) cudaFree(s->d_buf);
  if (s->h_in) cudaFreeHost(s->h_in);
  if (s->h_out) cudaFreeHost(s->h_out);
  if (s->st) cudaStreamDestroy(s->st);
  (void)cudaGetLastError();
  delete s;
  return GMR_OK;
}

int gmr_stream_reset(GmrStream* s, const double* qpos) {
  if (!s) return set_err(GMR_EINVAL, "stream is null");
  DeviceGuard g(s->m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  std::lock_guard<std::mutex> lk(s->mu);
  const int nq = s->m->h_f32.nq;
  std::vector<double> q(nq);
  for (int i = 0; i < nq; i++) q[i] = qpos ? qpos[i] : (double)s->m->h_f32.qpos0[i];
  if (!qpos) {                       // qpos0 in full precision comes from the float64 constant block on the device
    CK(cudaMemcpyAsync(s->d_buf + s->o_state, reinterpret_cast<const char*>(s->m->d_f64) + offsetof(GmrConsts<double>, qpos0),
                       (size_t)nq * 8, cudaMemcpyDeviceToDevice, s->st));
  } else {
    CK(cudaMemcpyAsync(s->d_buf + s->o_state, q.data(), (size_t)nq * 8, cudaMemcpyHostToDevice, s->st));
  }
  CK(cudaMemsetAsync(s->d_buf + s->o_warm, 0, 16, s->st));
  CK(cudaStreamSynchronize(s->st));
  return GMR_OK;
}

int gmr_stream_retarget(GmrStream* s, const float* pos, const float* quat, uint32_t flags, double* qpos_out,
                        int32_t* iters_out, double* err_out, double* targets_out) {
  if (!s) return set_err(GMR_EINVAL, "stream is null");
  if (!pos || !quat || !qpos_out) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  GmrModel* m = s->m;
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  std::lock_guard<std::mutex> lk(s->mu);
  const size_t nh = m->h_f32.nhum, nq = m->h_f32.nq;
  memcpy(s->h_in, pos, nh * 3 * 4);
  memcpy(s->h_in + s->o_quat, quat, nh * 4 * 4);
  const uint32_t f = flags & 3u;
  if (!s->graph[f] && !s->graph_failed) {
    // first call with these flags: run once eagerly (also opts the kernel into its shared-memory size), then
    // capture the same sequence for every later frame
    int rc = stream_enqueue(s, f);
    if (rc != GMR_OK) return rc;
    CK(cudaStreamSynchronize(s->st));
    cudaGraph_t gr = nullptr;
    bool ok = cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      ok = stream_enqueue(s, f) == GMR_OK;                  // recorded, not executed
      cudaError_t e = cudaStreamEndCapture(s->st, &gr);
      ok = ok && e == cudaSuccess && gr;
    }
    if (ok) ok = cudaGraphInstantiate(&s->graph[f], gr, 0) == cudaSuccess;
    if (gr) cudaGraphDestroy(gr);
    if (!ok) { s->graph[f] = nullptr; s->graph_failed = true; (void)cudaGetLastError(); }
  } else if (s->graph[f]) {
    CK(cudaGraphLaunch(s->graph[f], s->st));
    CK(cudaStreamSynchronize(s->st));
  } else {
    int rc = stream_enqueue(s, f);
    if (rc != GMR_OK) return rc;
    CK(cudaStreamSynchronize(s->st));
  }
  memcpy(qpos_out, s->h_out, nq * 8);
  if (err_out) memcpy(err_out, s->h_out + s->o_err, 16);
  if (targets_out) memcpy(targets_out, s->h_out + s->o_tg, nh * 7 * 8);
  if (iters_out) memcpy(iters_out, s->h_out + s->o_it, 8);
  return GMR_OK;
}

// Host-buffer entry.  Page-locked (pinned) host arrays are used IN PLACE: the solve kernel pulls each frame's keypoints
// over the host link with TMA bulk copies one frame ahead of the solve (it needs ~6 GB/s of a link that serves > 30 GB/s
// of such requests, tools/microbench/pcie.cu) and writes each frame's qpos straight into the caller's array, so the
// transfers ride inside the solve instead of in front of and behind it, and the whole batch is ONE schedule (two-phase
// partition included) — cutting clips into chunks to overlap copies would make every chunk pay its own tail.
// Pageable arrays cannot be mapped: they are staged through stream-ordered device scratch (cudaMemcpyAsync before /
// after the solve), in chunks of whole waves only when the staging would exceed a memory budget.
int gmr_retarget_batch_host(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                            const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out, uint32_t flags) {
  return gmr_retarget_batch_host_ex(m, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, nullptr, flags);
}

}  // extern "C"

namespace {
// device-visible alias of a host array, or null when the array is pageable
void* mapped_alias(const void* h) {
  if (!h) return nullptr;
  static const bool off = getenv("GMR_HOST_STAGED") && atoi(getenv("GMR_HOST_STAGED")) != 0;     // A/B knob: always stage
  if (off) return nullptr;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  if (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice) return a.devicePointer;
  return nullptr;
}

struct HostArr {
  const void* h_in = nullptr;   // input array on the host (or null)
  void* h_out = nullptr;        // output array on the host (or null)
  size_t per_clip = 0;          // bytes per clip
  char* alias = nullptr;        // device-visible alias of the whole array (pinned memory), else staged
  char* dev = nullptr;          // device pointer for the current chunk
  bool present() const { return h_in || h_out; }
  const char* host() const { return h_in ? (const char*)h_in : (const char*)h_out; }
};

int host_run(GmrModel* m, HostArr (&arr)[8], int C, int T, uint32_t flags) {
  enum { A_POS, A_QUAT, A_RATIO, A_INIT, A_Q, A_IT, A_ERR, A_ST };
  size_t staged_per_clip = 0;
  for (auto& a : arr) if (a.present()) { a.alias = (char*)mapped_alias(a.host()); if (!a.alias) staged_per_clip += a.per_clip; }
  const int wave = m->num_sms * ((flags & GMR_FLAG_COMPUTE_F64) ? pick_wpc<double>(m, C) : pick_wpc<float>(m, C));
  const size_t budget = (size_t)8 << 30;
  int chunk = C;
  while ((size_t)chunk * staged_per_clip > budget && chunk > wave) chunk -= wave;
  for (int i = 0; i < 2; i++) if (!m->hs[i]) CK(cudaStreamCreateWithFlags(&m->hs[i], cudaStreamNonBlocking));
  if (arr[A_ST].present() && arr[A_ST].alias) memset(arr[A_ST].h_out, 0, (size_t)C * 4);   // mapped status words: bits are OR-ed in
  int k = 0;
  for (int c0 = 0; c0 < C; c0 += chunk, k ^= 1) {
    const int n = (C - c0 < chunk) ? C - c0 : chunk;
    cudaStream_t st = m->hs[k];
    StreamScratch scratch(m->pool, st);
    size_t total = 0;
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    for (auto& a : arr) if (a.present() && !a.alias) total += pad(a.per_clip * n);
    if (total) CK(scratch.alloc(total));
    size_t o = 0;
    for (auto& a : arr) {
      if (!a.present()) { a.dev = nullptr; continue; }
      if (a.alias) { a.dev = a.alias + a.per_clip * c0; continue; }
      a.dev = scratch.p + o; o += pad(a.per_clip * n);
      if (a.h_in) CK(cudaMemcpyAsync(a.dev, (const char*)a.h_in + a.per_clip * c0, a.per_clip * n, cudaMemcpyHostToDevice, st));
    }
    GmrBatchExtra ex{};
    if (arr[A_ST].present()) {
      ex.status = (int32_t*)arr[A_ST].dev;
      if (!arr[A_ST].alias) CK(cudaMemsetAsync(arr[A_ST].dev, 0, (size_t)n * 4, st));
    }
    int rc;
    if (flags & GMR_FLAG_COMPUTE_F64)
      rc = launch<double, float, MAXW_F64>(m, m->d_f64, (const float*)arr[A_POS].dev, (const float*)arr[A_QUAT].dev, (const float*)arr[A_RATIO].dev,
                                           n, T, (const float*)arr[A_INIT].dev, (float*)arr[A_Q].dev, (int32_t*)arr[A_IT].dev,
                                           (float*)arr[A_ERR].dev, nullptr, flags, st, &ex);
    else
      rc = launch<float, float, MAXW_F32>(m, m->d_f32, (const float*)arr[A_POS].dev, (const float*)arr[A_QUAT].dev, (const float*)arr[A_RATIO].dev,
                                          n, T, (const float*)arr[A_INIT].dev, (float*)arr[A_Q].dev, (int32_t*)arr[A_IT].dev,
                                          (float*)arr[A_ERR].dev, nullptr, flags, st, &ex);
    if (rc != GMR_OK) return rc;
    for (auto& a : arr)
      if (a.h_out && !a.alias) CK(cudaMemcpyAsync((char*)a.h_out + a.per_clip * c0, a.dev, a.per_clip * n, cudaMemcpyDeviceToHost, st));
  }
  return GMR_OK;
}
}  // namespace

extern "C" {

int gmr_retarget_batch_host_ex(GmrModel* m, const float* pos, const float* quat, const float* ratio, int32_t C, int32_t T,
                               const float* qpos_init, float* qpos_out, int32_t* iters_out, float* err_out,
                               int32_t* status_out, uint32_t flags) {
  if (!m) return set_err(GMR_EINVAL, "model is null");
  if (C < 0 || T < 0) return set_err(GMR_EINVAL, "negative batch size");
  if (C == 0 || T == 0) return GMR_OK;
  if (!pos || !quat || !qpos_out) return set_err(GMR_EINVAL, "pos, quat and qpos_out are required");
  if ((reinterpret_cast<uintptr_t>(quat) & 15u) != 0) return set_err(GMR_EINVAL, "quat must be 16-byte aligned");
  DeviceGuard g(m->device);
  if (!g.ok) return set_err(GMR_ECUDA, "cannot select the model's device");
  std::lock_guard<std::mutex> lk(m->host_mu);              // the two host-entry streams are the model's
  const size_t nq = m->h_f32.nq, nh = m->h_f32.nhum;
  HostArr arr[8];
  arr[0].h_in = pos;        arr[0].per_clip = (size_t)T * nh * 12;
  arr[1].h_in = quat;       arr[1].per_clip = (size_t)T * nh * 16;
  arr[2].h_in = ratio;      arr[2].per_clip = 4;
  arr[3].h_in = qpos_init;  arr[3].per_clip = nq * 4;
  arr[4].h_out = qpos_out;  arr[4].per_clip = (size_t)T * nq * 4;
  arr[5].h_out = iters_out; arr[5].per_clip = (size_t)T * 8;
  arr[6].h_out = err_out;   arr[6].per_clip = (size_t)T * 8;
  arr[7].h_out = status_out; arr[7].per_clip = 4;
  int rc = host_run(m, arr, C, T, flags);
  // whatever happened, nothing of this call may still be in flight when the caller gets its arrays back
  for (int i = 0; i < 2; i++) {
    if (!m->hs[i]) continue;
    cudaError_t e = cudaStreamSynchronize(m->hs[i]);
    if (e != cudaSuccess && rc == GMR_OK) rc = cuda_err(e, "cudaStreamSynchronize");
  }
  return rc;
}

// ---- human-frame producers -------------------------------------------------------------------------------
static int producer_grid(int frames) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = (frames + 7) / 8, cap = sms * 8;          // 8 frames per block, a multiple of the SM count
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

int gmr_produce_bvh_frames(const float* lrot, const float* lpos, const int32_t* parents, int32_t F, int32_t J,
                           const int32_t* pos_joint, const int32_t* rot_joint, int32_t nh,
                           float* pos_out, float* quat_out, void* cuda_stream) {
  if (F < 0) return set_err(GMR_EINVAL, "negative frame count");
  if (F == 0) return GMR_OK;
  if (!lrot || !lpos || !parents || !pos_joint || !rot_joint || !pos_out || !quat_out) return set_err(GMR_EINVAL, "null argument");
  if ((reinterpret_cast<uintptr_t>(lrot) & 15u) || (reinterpret_cast<uintptr_t>(quat_out) & 15u)) return set_err(GMR_EINVAL, "quaternion arrays must be 16-byte aligned");
  // positions and orientations may come from different joints (FootMod): both sets of chains are needed
  std::vector<int32_t> both(2 * (size_t)(nh > 0 ? nh : 0));
  for (int b = 0; b < nh; b++) { both[b] = rot_joint[b]; both[nh + b] = pos_joint[b]; }
  GmrChain ch{};
  const char* why = "";
  if (nh < 1 || nh > 16) return set_err(GMR_ELIMIT, "1..16 output bodies");
  int rc = gmr_prod::build_chain(parents, J, both.data(), 2 * nh, &ch, &why);
  if (rc != 0) return set_err(rc == -4 ? GMR_ELIMIT : GMR_EINVAL, why);
  for (int b = 0; b < nh; b++) ch.pos_lane[b] = ch.rot_lane[nh + b];
  ch.nh = nh;
  gmr_prod::gmr_bvh_kernel<<<producer_grid(F), 256, 0, (cudaStream_t)cuda_stream>>>(ch, lrot, lpos, F, J, pos_out, quat_out);
  g_launches.fetch_add(1);
  CK(cudaGetLastError());
  return GMR_OK;
}

int gmr_produce_smplx_frames(const float* global_orient, const float* full_pose, const float* joints, const int32_t* parents,
                             int32_t F, int32_t NJ, int32_t NJ_joints, int32_t F_out, const int32_t* body_joint, int32_t nh,
                             float* pos_out, float* quat_out, void* cuda_stream) {
  if (F < 0 || F_out < 0) return set_err(GMR_EINVAL, "negative frame count");
  if (F == 0 || F_out == 0) return GMR_OK;
  if (!global_orient || !full_pose || !joints || !parents || !body_joint || !pos_out || !quat_out) return set_err(GMR_EINVAL, "null argument");
  if (reinterpret_cast<uintptr_t>(quat_out) & 15u) return set_err(GMR_EINVAL, "quat_out must be 16-byte aligned");
  if (F_out > F) return set_err(GMR_EINVAL, "F_out must not exceed F (the reference only resamples downwards)");
  GmrChain ch{};
  const char* why = "";
  int rc = gmr_prod::build_chain(parents, NJ, body_joint, nh, &ch, &why);
  if (rc != 0) return set_err(rc == -4 ? GMR_ELIMIT : GMR_EINVAL, why);
  for (int b = 0; b < nh; b++) {
    if (body_joint[b] >= NJ_joints) return set_err(GMR_EINVAL, "joint index beyond the joints array");
    ch.pos_lane[b] = (uint8_t)body_joint[b];
  }
  gmr_prod::gmr_smplx_kernel<<<producer_grid(F_out), 256, 0, (cudaStream_t)cuda_stream>>>(
      ch, global_orient, full_pose, joints, F, NJ, NJ_joints, F_out, F_out != F ? 1 : 0, pos_out, quat_out);
  g_launches.fetch_add(1);
  CK(cudaGetLastError());
  return GMR_OK;
}

int64_t gmr_launch_count(void) { return g_launches.load(); }

// Profiling aid, not part of the documented ABI: while a device buffer of [2C,4] int64 is registered, every clip of a
// single-robot launch records {start ns, end ns, SM | warp << 16, solves | factorisations << 32} (tools/prof/timeline.py).
void gmr_debug_trace(long long* device_buffer) { g_trace.store(device_buffer); }

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
