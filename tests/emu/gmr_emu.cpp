// Lane-serial HOST build of the CUDA kernel body (general_motion_retargeting_b200/csrc/
// gmr_solver.cuh compiled with GMR_EMULATE): a development aid that lets the warp-level
// algorithm be debugged against the oracle on a machine without a GPU.  It is test
// infrastructure: nothing in the package loads it and it is never a fallback for the GPU path.
#define GMR_EMULATE 1
#include "../../general_motion_retargeting_b200/csrc/gmr_solver.cuh"

#include <atomic>
#include <thread>
#include <vector>

template <typename R>
static int run(const GmrModelDesc* d, const float* pos, const float* quat, const float* ratio, int C, int T,
               const double* qpos_init, double* qpos_out, int32_t* iters_out, double* err_out, double* tg_out,
               uint32_t flags, int nthreads, int64_t* refactor_out, const GmrBatchExtra* ex) {
  auto* mc = new GmrConsts<R>();
  const char* why = nullptr;
  int rc = gmr_fill_consts<R>(d, mc, &why);
  if (rc != GMR_OK) { delete mc; return rc; }
  GmrDims dims = gmr_dims_of(*mc);
  gmr_dims_layout<R>(dims);
  const GmrScal<R> ks = gmr_scal_of(*mc);
  const int wel = dims.warp_elems;
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads > C) nthreads = C > 0 ? C : 1;
  GmrIO<double> io{};
  io.pos = pos; io.quat = quat; io.ratio = ratio; io.qinit = qpos_init; io.qout = qpos_out; io.iters = iters_out; io.err = err_out;
  io.tg = tg_out; io.C = C; io.T = T; io.flags = flags;
  if (ex) io.ex = *ex;
  std::atomic<int> next{0};
  std::atomic<int64_t> refac{0};
  auto work = [&]() {
    std::vector<R> sm((size_t)wel + 16, R(0));
    for (;;) {
      int c = next.fetch_add(1);
      if (c >= C) break;
      auto* ws = new WarpSolver<R>(*mc, dims, ks, sm.data());
      ws->template run_clip<double>(io, c);
      refac += ws->stat_refactor;
      delete ws;
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (refactor_out) *refactor_out = refac.load();
  delete mc;
  return GMR_OK;
}

extern "C" int gmr_emu_retarget_batch(const GmrModelDesc* d, const float* pos, const float* quat, const float* ratio,
                                      int32_t C, int32_t T, const double* qpos_init, double* qpos_out,
                                      int32_t* iters_out, double* err_out, double* tg_out, uint32_t flags,
                                      int32_t nthreads, int32_t precision_bits, int64_t* refactor_out) {
  if (precision_bits == 32)
    return run<float>(d, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, tg_out, flags, nthreads, refactor_out, nullptr);
  return run<double>(d, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, tg_out, flags, nthreads, refactor_out, nullptr);
}

// same with the GmrBatchExtra members (host pointers here): ragged lengths, local body positions, lowest z
extern "C" int gmr_emu_retarget_batch_ex(const GmrModelDesc* d, const float* pos, const float* quat, const float* ratio,
                                         int32_t C, int32_t T, const double* qpos_init, double* qpos_out,
                                         int32_t* iters_out, double* err_out, double* tg_out, const GmrBatchExtra* ex,
                                         uint32_t flags, int32_t nthreads, int32_t precision_bits) {
  if (precision_bits == 32)
    return run<float>(d, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, tg_out, flags, nthreads, nullptr, ex);
  return run<double>(d, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, tg_out, flags, nthreads, nullptr, ex);
}

// order in which the emulator visits the lanes of a block (a permutation of 0..31; not thread safe: set it between runs)
// Per-warp shared-memory layout of precision `bits` for a model: out[0..15] = { GS_VAR, o_u, o_xq, o_xp, o_sd, o_tg, o_in,
// warp_elems, PX, SD, TG, MT, row stride, ROWH, consts bytes, sizeof(R) } (tests/test_emulator.py checks alignment, bank-friendly
// strides, disjoint regions and the 16-warp budget the float64 kernel relies on for G1).
template <typename R> static int layout(const GmrModelDesc* d, int32_t* out) {
  auto* mc = new GmrConsts<R>();
  const char* why = nullptr;
  int rc = gmr_fill_consts<R>(d, mc, &why);
  if (rc != GMR_OK) { delete mc; return rc; }
  GmrDims dm = gmr_dims_of(*mc);
  gmr_dims_layout<R>(dm);
  const int32_t v[16] = {GS_VAR, dm.o_u, dm.o_xq, dm.o_xp, dm.o_sd, dm.o_tg, dm.o_in, dm.warp_elems, GmrLay<R>::PX, GmrLay<R>::SD,
                         GmrLay<R>::TG, GmrLay<R>::MT, dm.rs, GmrLay<R>::ROWH, (int32_t)((sizeof(GmrConsts<R>) + 15) / 16 * 16), (int32_t)sizeof(R)};
  for (int i = 0; i < 16; i++) out[i] = v[i];
  delete mc;
  return GMR_OK;
}
extern "C" int gmr_emu_layout(const GmrModelDesc* d, int32_t precision_bits, int32_t* out) {
  return precision_bits == 32 ? layout<float>(d, out) : layout<double>(d, out);
}

extern "C" int gmr_emu_set_lane_order(const int* order) {
  bool seen[32] = {};
  for (int i = 0; i < 32; i++) { if (order[i] < 0 || order[i] > 31 || seen[order[i]]) return -1; seen[order[i]] = true; }
  for (int i = 0; i < 32; i++) gmr_emu_lane_order[i] = order[i];
  return 0;
}

#ifdef GMR_STATS
extern "C" void gmr_emu_stats(long long* out, int reset) {
  for (int i = 0; i < 16; i++) { out[i] = gmr_stats[i].load(); if (reset) gmr_stats[i] = 0; }
}
#endif
