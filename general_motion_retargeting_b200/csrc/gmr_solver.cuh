// Warp-per-clip fused retargeting IK solver (the body of the sm_100a kernel).
//
// One warp owns one clip: frames are solved in order, warm-started, exactly like
// `for frame in frames: retargeter.retarget(frame)` in the reference
// (general_motion_retargeting/motion_retarget.py:139-185; callers
// scripts/smplx_to_robot_dataset.py:84-87).  This is NOT a port of the mink/MuJoCo/DAQP
// call sequence; the per-solve algebra is re-derived for a 32-lane warp:
//
//  * FK as a pointer-jumping scan over the tree of moving bodies (lane = hinge body; log2(depth) dependent
//    compositions, ping-pong between two pose buffers), positions kept relative to the floating root so
//    float32 keeps ~1e-7 m resolution anywhere in the world.
//  * The QP matrix H = damping*I + sum_t [(W J_t)^T (W J_t) + mu_t I] is never formed from
//    dense 6 x nv task Jacobians.  Each frame task is a 6x6 "spring inertia"
//    M_t = A_t^T A_t at the reference point (A_t = -W * Jlog_t * blkdiag(R_b^T) shifted to the
//    root); the M_t are summed leaf-to-root over the task tree, every hinge lane takes the
//    composite of its subtree (as in the composite-rigid-body algorithm) and gets its row as
//    H_ij = s_i^T Ic_i s_j over its ancestors j, with s_j the world-frame spatial axis of DoF j.
//    ~10x fewer flops than the dense J^T W^2 J and no shared-memory read-modify-write.
//  * H_ij is non-zero only between a hinge and its ancestors (branch-induced sparsity).  Row i
//    lives in REGISTERS of lane i, indexed by the DEPTH of the ancestor (6 floating-base slots, then
//    <= GMR_MAXD - 1 ancestors): every hinge of a root-to-leaf chain sees the same ancestor at the
//    same slot, so all register indices are static.  The factorisation is a branch-sparse L^T D L
//    that eliminates hinges by depth, leaves first (no fill-in): the hinges of one depth publish
//    their rows to shared memory, every lane absorbs its own pivot of that depth (branches side by side);
//    the floating base is the 6x6 Schur complement of everything, accumulated by 27 lanes in one
//    constant-stride loop over the published rows; back substitution travels by warp shuffle.
//  * Joint-limit box: exact primal active set (same optimum as the reference's DAQP solve),
//    working set kept per stage and warm-started from the previous solve; the common case
//    (unconstrained step feasible) costs one factorisation.
//
// The file is written as a sequence of "lane blocks" (GMR_LANES ... GMR_END).  On the GPU a
// block is straight-line code of one thread followed by __syncwarp(); defining GMR_EMULATE
// turns each block into a host loop over 32 lanes, which is how tests/emu debugs the
// warp-level logic on a machine without a GPU.  Rule: inside one block a lane never reads
// shared memory another lane writes in the same block.
// A warp issues IN ORDER: a block made of several `if (lane-dependent)` regions runs their load -> FMA
// chains one after the other, the same block written with selects overlaps them - the hot blocks are
// branch-free for that reason (ncu per source line: profiles/r2_by_phase_f64_*.txt).
#pragma once
#include <math.h>
#include <stdint.h>

#include "gmr_consts.h"

#ifdef GMR_EMULATE
#include <cmath>
#include <cstring>
#define GMR_FN inline
#define GMR_HD inline
// The emulator runs the lanes of a block one after the other, in the order gmr_emu_lane_order gives (identity by default).
// A lane block must not depend on that order - "inside one block a lane never reads shared memory another lane writes in
// the same block" - so tests/test_emulator.py runs the same clips under reversed and shuffled orders and demands
// bit-identical results: the CPU stand-in for a racecheck of the blocks' intra-warp hazards.
static int gmr_emu_lane_order[32] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
#define GMR_LANES for (int li_ = 0; li_ < 32; ++li_) { const int lane = gmr_emu_lane_order[li_]; LaneRegs<R>& L = lanes_[lane]; (void)L;
#define GMR_END }
#define GMR_UNROLL
#define GMR_NOUNROLL
#define GMR_SYNC()
// end the lane loop, broadcast per-lane register(s) of lane `src` to every lane, open a new lane loop
#define GMR_BCAST1(src, f0) } { const R bc0 = lanes_[src].f0; for (int li_ = 0; li_ < 32; ++li_) { const int lane = gmr_emu_lane_order[li_]; LaneRegs<R>& L = lanes_[lane]; (void)L;
#define GMR_BCAST2(src, f0, f1) } { const R bc0 = lanes_[src].f0, bc1 = lanes_[src].f1; for (int li_ = 0; li_ < 32; ++li_) { const int lane = gmr_emu_lane_order[li_]; LaneRegs<R>& L = lanes_[lane]; (void)L;
#define GMR_END_BCAST } }
#define GMR_END_BCAST_NOSYNC } }
#define GMR_END_NOSYNC }
#define GMR_DEPTH ((int)mc.hdepth[lane])
#define GMR_DESC (mc.desc_mask[lane])
#define GMR_SR (gmr_schur_r(lane))
#define GMR_SC (gmr_schur_c<R>(lane))
#define GMR_CTZ(x) __builtin_ctz(x)
#else
#define GMR_FN __device__ __forceinline__
#define GMR_HD __host__ __device__ inline
#define GMR_LANES { LaneRegs<R>& L = lanes_; const int lane = lane_; (void)L; (void)lane;
#define GMR_END } __syncwarp();
#define GMR_UNROLL _Pragma("unroll")
#define GMR_NOUNROLL _Pragma("unroll 1")
#define GMR_SYNC() __syncwarp()
// warp shuffles: no shared-memory round trip and no extra __syncwarp for a one-lane -> all-lanes hand-over
#define GMR_BCAST1(src, f0) const R bc0 = __shfl_sync(0xffffffffu, L.f0, src);
#define GMR_BCAST2(src, f0, f1) const R bc0 = __shfl_sync(0xffffffffu, L.f0, src), bc1 = __shfl_sync(0xffffffffu, L.f1, src);
#define GMR_END_BCAST } __syncwarp();
#define GMR_END_BCAST_NOSYNC }
#define GMR_END_NOSYNC }
#define GMR_DEPTH (dep_)
#define GMR_DESC (desc_)
#define GMR_SR (sr_)
#define GMR_SC (sc_)
#define GMR_CTZ(x) (__ffs(x) - 1)
#endif

// ---- shared-memory layout of one warp's state (units: elements of R) -----------------------
GMR_HD constexpr int gmr_pad4(int n) { return (n + 3) & ~3; }
// Published factor row of hinge k (written when k is eliminated, read by its ancestors and by the
// back substitution):  [ y_k, 1/d_k (, 0, 0 in float32) | H'_k,base(6) | H'_k,anc(depth 1) .. H'_k,anc(depth d_k - 1) | junk up to o_y ]
// (the header comes first so that every offset inside a row is a compile-time constant; Lrow(k) points at slot 0)
// Slots are indexed by the DEPTH of the ancestor, not by its hinge number: every hinge on the chain root .. k
// sees the same ancestor at the same slot, which is what lets lane i update row_i[s] -= a * row_k[s] with
// statically indexed registers.
GMR_HD constexpr int gmr_row_oy(int maxd) { return gmr_pad4(6 + (maxd > 0 ? maxd : 1)); }
// Record strides (elements of R).  Shared memory serves a 128-bit access of a quarter-warp in one wavefront only when the
// eight 16-byte units fall into different bank groups: records that consecutive lanes touch are therefore laid out at a
// stride that is ODD in 16-byte units (float64: 10-element poses = 5 units, 6-element spatial axes = 3, 30-element task
// blocks = 15, factor rows 18 = 9; float32: 12-element poses = 3 units, 28-element task blocks = 7, factor rows 20 = 5).
// ncu, round 2, balanced float64 batch: the shared-memory pipe was 73 % busy and 40 % of its wavefronts were bank
// conflicts of the old power-of-two strides (profiles/r2_smem_wavefronts_before.txt).
template <typename R> struct GmrLay;
template <> struct GmrLay<double> { enum { PX = 10, SD = 6, TG = 10, MT = 30, ROWH = 2 }; };
template <> struct GmrLay<float>  { enum { PX = 12, SD = 8, TG = 8,  MT = 28, ROWH = 4 }; };
// the row stride is a compile-time constant (all GMR_NS slots, whatever the robot's deepest chain): row addresses are
// k * constant + base, and the loop over all rows at the end of the factorisation unrolls with immediate offsets
template <typename R> GMR_HD constexpr int gmr_row_stride(int) { return GMR_NS + GmrLay<R>::ROWH; }

// fixed part
enum {
  GS_RED = 0,               // [32] reductions / hand-over scratch
  GS_XS = 32,               // [40] solution of the last linear solve (base 6, hinges)
  GS_ROOT = 72,             // [28] whole-tree composite: M(21) g(6) mu
  GS_LF = 100,              // [28] base 6x6 Schur complement (21, packed upper) + its right-hand side (6)
  GS_PIV = 128,             // [4]
  GS_RQ = 132,              // [4] normalised root quaternion of the last FK
  GS_BAR = 136,             // [4] mbarrier of the keypoint stream (8 bytes used)
  GS_BND = 140,             // [32] bound value of pinned hinges
  GS_LP = 172,              // [9][32] lane-private slots
  GS_Q = 172 + 9 * 32,      // [40] qpos
  GS_VAR = 172 + 9 * 32 + 40
};
enum { LP_F = 0, LP_DIAG = 6, LP_CI = 7, LP_X = 8 };

// staged raw keypoints of one frame (floats): [ positions nhum x 3, placed at the source's offset within its 16-byte line
// (<= 3 floats of slack in front, so that the TMA bulk copy of the aligned body lands 16-byte aligned) | quaternions nhum x 4 ]
GMR_HD constexpr int gmr_in_quat(int nhum) { return gmr_pad4(3 * nhum + 4); }
template <typename R> GMR_HD int gmr_in_elems(int nhum) { return gmr_pad4((int)(((gmr_in_quat(nhum) + 4 * nhum) * sizeof(float) + sizeof(R) - 1) / sizeof(R))); }
// variable part: U sd[SD nh] tg[TG nhum] in[staged floats], where the union U (first: its offset is a compile-time constant) holds
//   [ mt: task blocks, MT nt | body poses, PX nb: (quat[4], pos[3], pad ..) ]   while FK / task evaluation are live, and
//   [ published factor rows, nh * stride ]                                       from the factorisation to the next FK.
// (body poses are dead once the composites are built; integrate() takes the root quaternion from GS_RQ.)
// task blocks; during FK the same region is the second pose buffer of the scan (ping-pong), hence at least PX * nb
template <typename R> GMR_HD constexpr int gmr_mt_elems(int nt, int nb) { return gmr_pad4(GmrLay<R>::MT * nt > GmrLay<R>::PX * nb ? GmrLay<R>::MT * nt : GmrLay<R>::PX * nb); }
template <typename R> GMR_HD constexpr int gmr_sd_elems(int nh) { return gmr_pad4(GmrLay<R>::SD * (nh > 0 ? nh : 1)); }
template <typename R> GMR_HD int gmr_union_elems(int nb, int nh, int nt, int maxd) {
  const int u = gmr_mt_elems<R>(nt, nb) + gmr_pad4(GmrLay<R>::PX * nb);
  const int lr = (nh > 0 ? nh : 1) * gmr_row_stride<R>(maxd);
  return gmr_pad4(lr > u ? lr : u);
}
template <typename R> GMR_HD int gmr_warp_elems(int nb, int nh, int nhum, int nt, int maxd) {
  return GS_VAR + gmr_union_elems<R>(nb, nh, nt, maxd) + gmr_sd_elems<R>(nh) + gmr_pad4(GmrLay<R>::TG * nhum) + gmr_in_elems<R>(nhum);
}

// fills the layout fields of `d` for precision R (host side, before launch)
template <typename R> inline void gmr_dims_layout(GmrDims& d) {
  int o = GS_VAR;
  d.o_u = o;
  d.o_xq = o + gmr_mt_elems<R>(d.nt, d.nb);       // poses are one record of PX elements per body: orientation at +0, position at +4
  d.o_xp = d.o_xq + 4;
  o += gmr_union_elems<R>(d.nb, d.nh, d.nt, d.maxd);
  d.o_sd = o; o += gmr_sd_elems<R>(d.nh);
  d.o_tg = o; o += gmr_pad4(GmrLay<R>::TG * d.nhum);
  d.o_in = o; o += gmr_in_elems<R>(d.nhum);
  d.o_y = gmr_row_oy(d.maxd);
  d.rs = gmr_row_stride<R>(d.maxd);
  d.warp_elems = gmr_warp_elems<R>(d.nb, d.nh, d.nhum, d.nt, d.maxd);
}

// relative tolerance of the KKT multiplier sign test (the Lie threshold is a model parameter: GmrScal::lie)
#if defined(GMR_EMULATE) && defined(GMR_STATS)
// active-set statistics of the emulator build (tools/prof/active_set_stats.py): [0] solves, [1] SOLVE passes (factorisations),
// [2] passes that ended blocked (a bound added), [3] CHECK passes, [4] checks that released bounds, [5] bounds released,
// [6] sum of pinned joints at the end of a solve, [7] solves that ended with pins, [8] solves whose warm set was already optimal
#include <atomic>
static std::atomic<long long> gmr_stats[16];
#define GMR_STAT(i, n) gmr_stats[i] += (n)
#else
#define GMR_STAT(i, n)
#endif
template <typename R> struct GmrEps;
template <> struct GmrEps<float>  { static constexpr float  lam = 1e-5f;  };
template <> struct GmrEps<double> { static constexpr double lam = 1e-12; };

// per-lane registers that persist across lane blocks (what the factorisation keeps live;
// colder per-lane state sits in the lane-private shared-memory slots GS_LP)
template <typename R> struct LaneRegs {
  R row[GMR_NS];   // this hinge's row of H by slot: 6 floating-base columns, then its strict ancestors by depth
  R dg;            // its diagonal
  R rhs, dinv;
  R sacc;          // lanes 0..26: one entry of the base block's Schur complement (accumulated during the elimination)
  uint32_t piv;    // pivots of the depth being eliminated that this lane still has to absorb (its descendants)
};

// ---- tiny math helpers ----------------------------------------------------------------------
template <typename R> GMR_FN R g_sqrt(R x) { return sqrt(x); }
template <typename R> GMR_FN R g_atan2(R y, R x) { return atan2(y, x); }
// square root of a strictly positive, normal-range argument (callers branch away from ~0 first)
template <typename R> GMR_FN R g_sqrt_pos(R x) { return sqrt(x); }
template <typename R> GMR_FN R g_abs(R x) { return fabs(x); }
template <typename R> GMR_FN R g_rsqrt(R x) {
#ifdef GMR_EMULATE
  return R(1) / std::sqrt(x);
#else
  return rsqrt(x);
#endif
}
// sin/cos by quadrant reduction (x = k pi/2 + r, |r| <= pi/4) and minimax-free Taylor polynomials:
// hinge half-angles and IK rotation steps are a few radians at most, so a two-constant Cody-Waite
// reduction is exact enough (|x| < 1e4) and the whole thing is ~35 straight FMAs instead of the
// library routine with its large-argument path.  Error < 1 ulp-ish of the working precision.
template <typename R> GMR_FN void g_sincos(R x, R* s, R* c) {
  const R k = rint(x * R(0.63661977236758134308));
  R r = fma(-k, R(1.57079632679489655800e+00), x);
  r = fma(-k, R(6.12323399573676603587e-17), r);
  const R r2 = r * r;
  R sp, cp;
  if (sizeof(R) == 8) {
    sp = fma(r2, R(-7.6471637318198164759e-13), R(1.6059043836821614599e-10));
    sp = fma(r2, sp, R(-2.5052108385441718775e-08)); sp = fma(r2, sp, R(2.7557319223985890653e-06));
    sp = fma(r2, sp, R(-1.9841269841269841270e-04)); sp = fma(r2, sp, R(8.3333333333333333333e-03));
    sp = fma(r2, sp, R(-1.6666666666666666667e-01));
    cp = fma(r2, R(4.7794773323873852974e-14), R(-1.1470745597729724714e-11));
    cp = fma(r2, cp, R(2.0876756987868098979e-09)); cp = fma(r2, cp, R(-2.7557319223985890653e-07));
    cp = fma(r2, cp, R(2.4801587301587301587e-05)); cp = fma(r2, cp, R(-1.3888888888888888889e-03));
    cp = fma(r2, cp, R(4.1666666666666666667e-02)); cp = fma(r2, cp, R(-0.5));
  } else {
    sp = fma(r2, R(2.7557319223985890653e-06), R(-1.9841269841269841270e-04));
    sp = fma(r2, sp, R(8.3333333333333333333e-03)); sp = fma(r2, sp, R(-1.6666666666666666667e-01));
    cp = fma(r2, R(-2.7557319223985890653e-07), R(2.4801587301587301587e-05));
    cp = fma(r2, cp, R(-1.3888888888888888889e-03)); cp = fma(r2, cp, R(4.1666666666666666667e-02)); cp = fma(r2, cp, R(-0.5));
  }
  const R sr = fma(r * r2, sp, r), cr = fma(r2, cp, R(1));
  const int q = (int)k & 3;
  const R ss = (q & 1) ? cr : sr, cc = (q & 1) ? sr : cr;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
// reciprocal of a pivot (x >= damping > 0, far from the range limits): hardware seed + Newton steps, no slow path
template <typename R> GMR_FN R g_rcp_pos(R x) { return R(1) / x; }
#ifndef GMR_EMULATE
template <> __device__ __forceinline__ float g_rcp_pos<float>(float x) {
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}
template <> __device__ __forceinline__ double g_rcp_pos<double>(double x) {
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return fma(r, fma(-x, r, 1.0), r);
}
#endif
#ifndef GMR_EMULATE
template <> __device__ __forceinline__ float g_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ float g_rsqrt<float>(float x) { return rsqrtf(x); }
// float64 1/sqrt for positive, normal-range arguments (pivots >= damping, squared quaternion norms): hardware seed
// (~20 bits) + two Newton steps instead of the library routine with its special-case handling (~25 instructions,
// six of them back to back on the base block's dependent chain)
template <> __device__ __forceinline__ double g_rsqrt<double>(double x) {
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  r = fma(r, fma(-hx * r, r, 0.5), r);
  r = fma(r, fma(-hx * r, r, 0.5), r);
  return r;
}
template <> __device__ __forceinline__ double g_sqrt_pos<double>(double x) {
  const double r = g_rsqrt<double>(x);
  const double s = x * r;
  return fma(fma(-s, s, x), 0.5 * r, s);                 // one Heron correction
}
template <> __device__ __forceinline__ float g_sqrt_pos<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ float g_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> __device__ __forceinline__ float g_abs<float>(float x) { return fabsf(x); }
#endif

// global load that bypasses the (non-coherent) L1: clip state records and status words travel between SMs within one launch
#ifdef GMR_EMULATE
template <typename T> GMR_FN T g_ldcg(const T* p) { return *p; }
#else
template <typename T> GMR_FN T g_ldcg(const T* p) { return __ldcg(p); }
#endif

// 4 consecutive elements from 16-byte aligned shared memory (one LDS.128 / two LDS.128)
GMR_FN void g_ld4(const float* p, float* v) {
#ifdef GMR_EMULATE
  v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#else
  const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#endif
}
GMR_FN void g_ld4(const double* p, double* v) {
#ifdef GMR_EMULATE
  v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#else
  const double2 t = *reinterpret_cast<const double2*>(p), u = *reinterpret_cast<const double2*>(p + 2);
  v[0] = t.x; v[1] = t.y; v[2] = u.x; v[3] = u.y;
#endif
}

// 4 consecutive elements to 16-byte aligned shared memory
GMR_FN void g_st4(float* p, float a, float b, float c, float d) {
#ifdef GMR_EMULATE
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#else
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
#endif
}
GMR_FN void g_st4(double* p, double a, double b, double c, double d) {
#ifdef GMR_EMULATE
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#else
  *reinterpret_cast<double2*>(p) = make_double2(a, b); *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
#endif
}

// 2 consecutive elements (float64: one 128-bit access, 16-byte aligned; float32: one 64-bit access, 8-byte aligned)
GMR_FN void g_ld2(const float* p, float* v) {
#ifdef GMR_EMULATE
  v[0] = p[0]; v[1] = p[1];
#else
  const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
#endif
}
GMR_FN void g_ld2(const double* p, double* v) {
#ifdef GMR_EMULATE
  v[0] = p[0]; v[1] = p[1];
#else
  const double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y;
#endif
}
GMR_FN void g_st2(float* p, float a, float b) {
#ifdef GMR_EMULATE
  p[0] = a; p[1] = b;
#else
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
#endif
}
GMR_FN void g_st2(double* p, double a, double b) {
#ifdef GMR_EMULATE
  p[0] = a; p[1] = b;
#else
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
#endif
}

template <typename R> GMR_FN void q_mul(const R* a, const R* b, R* o) {
  R w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  R x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  R y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  R z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
// v' = R(q) v
template <typename R> GMR_FN void q_rot(const R* q, const R* v, R* o) {
  R tx = R(2) * (q[2] * v[2] - q[3] * v[1]), ty = R(2) * (q[3] * v[0] - q[1] * v[2]), tz = R(2) * (q[1] * v[1] - q[2] * v[0]);
  R ox = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty);
  R oy = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz);
  R oz = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
  o[0] = ox; o[1] = oy; o[2] = oz;
}
// v' = R(q)^T v
template <typename R> GMR_FN void q_rot_inv(const R* q, const R* v, R* o) {
  R tx = R(2) * (q[2] * v[2] - q[3] * v[1]), ty = R(2) * (q[3] * v[0] - q[1] * v[2]), tz = R(2) * (q[1] * v[1] - q[2] * v[0]);
  R ox = v[0] - q[0] * tx + (q[2] * tz - q[3] * ty);
  R oy = v[1] - q[0] * ty + (q[3] * tx - q[1] * tz);
  R oz = v[2] - q[0] * tz + (q[1] * ty - q[2] * tx);
  o[0] = ox; o[1] = oy; o[2] = oz;
}
template <typename R> GMR_FN void q_to_mat(const R* q, R* m) {
  R w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = R(2) * (x * y - w * z); m[2] = R(2) * (x * z + w * y);
  m[3] = R(2) * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = R(2) * (y * z - w * x);
  m[6] = R(2) * (x * z - w * y); m[7] = R(2) * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <typename R> GMR_FN void q_normalize(R* q) {      // mju_normalize4
  const R n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < R(1e-30)) { q[0] = R(1); q[1] = q[2] = q[3] = R(0); return; }
  const R inv = g_rsqrt(n2);
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
// C = A * B (3x3 row-major)
template <typename R> GMR_FN void m3_mul(const R* A, const R* B, R* C) {
  GMR_UNROLL
  for (int i = 0; i < 3; i++) {
    GMR_UNROLL
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  }
}
// C = A * B^T
template <typename R> GMR_FN void m3_mul_bt(const R* A, const R* B, R* C) {
  GMR_UNROLL
  for (int i = 0; i < 3; i++) {
    GMR_UNROLL
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
  }
}
// C = A^T * B
template <typename R> GMR_FN void m3_mul_at(const R* A, const R* B, R* C) {
  GMR_UNROLL
  for (int i = 0; i < 3; i++) {
    GMR_UNROLL
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
  }
}
template <typename R> GMR_FN void m3_skew(const R* v, R* S) {
  S[0] = R(0); S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = R(0); S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = R(0);
}

// lane e < 21 owns entry (r, c) of the packed upper triangle of the base block, lanes 21..26 its right-hand side
GMR_HD constexpr int gmr_schur_r(int e) { return e < 21 ? (e >= 6) + (e >= 11) + (e >= 15) + (e >= 18) + (e >= 20) : (e < 27 ? e - 21 : 0); }
template <typename R> GMR_HD constexpr int gmr_schur_c(int e) { return e < 21 ? gmr_schur_r(e) + e - (gmr_schur_r(e) * (13 - gmr_schur_r(e))) / 2 : (e < 27 ? -(int)GmrLay<R>::ROWH : 0); }
// index of (i,j) in the packed upper triangle of a symmetric 6x6 (21 entries)
GMR_HD constexpr int gmr_sym6(int i, int j) { return i <= j ? (i * (13 - i)) / 2 + (j - i) : (j * (13 - j)) / 2 + (i - j); }

// base pointers and sizes of one batch (see include/gmr_b200.h: gmr_retarget_batch); IO = buffer type of qpos/err/targets
template <typename IO> struct GmrIO {
  const float* pos; const float* quat; const float* ratio;
  const IO* qinit; IO* qout; int32_t* iters; IO* err; IO* tg;
  int32_t C, T; uint32_t flags, _pad;
  GmrBatchExtra ex;
  // a launch may cover only frames [t_begin, t_end) of every clip (t_end <= 0: to the end); `state`
  // ([C, nq + 6] doubles: qpos | running lowest z | last error | working sets as 4 x u32 | next frame | pad) carries a
  // clip from one launch - or one SEGMENT of the scheduled launch, see run_clip - to the next at full precision
  int32_t t_begin, t_end;
  double* state;
  // profiling aid (gmr_debug_trace): [2C,4] int64 {first start ns, last end ns, running ns, factorisations | segments << 32 | last SM << 48} per clip;
  // rows [0, C) are written by launches that start at frame 0, rows [C, 2C) by launches that continue a clip
  long long* trace;
};
GMR_HD constexpr int gmr_state_stride(int nq) { return nq + 6; }

// =============================================================================================
template <typename R>
struct WarpSolver {
  const GmrConsts<R>& mc;     // per-robot tables (shared memory on the GPU)
  const GmrDims& dm;          // sizes / masks (kernel parameter: constant bank)
  const GmrScal<R>& ks;       // solver knobs (kernel parameter)
#ifdef GMR_EMULATE
  R* sm;                      // this warp's state block
#else
  // this warp's state block as a 32-bit byte offset into the dynamic shared window: keeps every
  // access a plain LDS/STS [reg + imm] instead of 64-bit generic pointer arithmetic
  uint32_t sm_off;
  struct SmProxy {
    uint32_t off;
    __device__ __forceinline__ R* ptr() const { extern __shared__ __align__(128) unsigned char gmr_dyn_smem[]; return reinterpret_cast<R*>(gmr_dyn_smem + off); }
    __device__ __forceinline__ R* operator+(int i) const { return ptr() + i; }
    __device__ __forceinline__ R& operator[](int i) const { return ptr()[i]; }
  };
#endif
#ifdef GMR_EMULATE
  LaneRegs<R> lanes_[32];
#else
  LaneRegs<R> lanes_;
  int lane_;
  int dep_;                    // hinge depth of this lane (0: lane owns no hinge)
  uint32_t desc_;              // strict descendants of this lane's hinge
  int sr_, sc_;                // this lane's entry of the base block (row slot, column slot; the right-hand side is the row header)
#endif
  int stat_refactor;           // factorisations done (uniform)
  uint32_t stat_flags;         // GMR_STATUS_* events of the current clip (uniform)
  bool convoy;                 // CTA-wide rendezvous before every factorisation (see convoy_arrive)
  volatile int* cta_active;    // shared count of warps that still have clips (convoy mode)
  // working set carried from the previous solve of the same STAGE (uniform).  Four scalars, selected by
  // comparison: a dynamically indexed member array would push the whole solver object into local memory.
  uint32_t warm_lo0, warm_hi0, warm_lo1, warm_hi1;
  // keypoint stream (uniform): float offset of the staged positions inside their slot (= the source frame's offset within
  // its 16-byte line), whether a frame is in flight, and the phase of the warp's mbarrier
  uint32_t in_off, in_phase;
  bool in_flight;

#ifdef GMR_EMULATE
  GMR_FN WarpSolver(const GmrConsts<R>& m, const GmrDims& d, const GmrScal<R>& k, R* smem)
#else
  // smem_byte_off: offset of this warp's state block from the start of the dynamic shared window
  GMR_FN WarpSolver(const GmrConsts<R>& m, const GmrDims& d, const GmrScal<R>& k, uint32_t smem_byte_off, int lane)
#endif
                     : mc(m), dm(d), ks(k), stat_refactor(0), stat_flags(0), convoy(false), cta_active(nullptr), warm_lo0(0), warm_hi0(0), warm_lo1(0), warm_hi1(0),
                       in_off(0), in_phase(0), in_flight(false) {
#ifndef GMR_EMULATE
    lane_ = lane;
    dep_ = lane < d.nh ? (int)m.hdepth[lane] : 0;
    desc_ = lane < d.nh ? m.desc_mask[lane] : 0u;
    sr_ = gmr_schur_r(lane); sc_ = gmr_schur_c<R>(lane);
    sm_off = smem_byte_off;
#else
    sm = smem;
#endif
  }

#ifndef GMR_EMULATE
  #define sm (SmProxy{sm_off})
#endif
  GMR_FN R* s_red() const { return sm + GS_RED; }
  GMR_FN R* s_xs() const { return sm + GS_XS; }
  GMR_FN R* s_root() const { return sm + GS_ROOT; }
  GMR_FN R* s_lf() const { return sm + GS_LF; }
  GMR_FN R* s_piv() const { return sm + GS_PIV; }
  GMR_FN R* s_bnd() const { return sm + GS_BND; }
  GMR_FN R* s_rq() const { return sm + GS_RQ; }
  GMR_FN R& lp(int slot, int lane) const { return *(sm + (GS_LP + slot * 32 + lane)); }
  GMR_FN R* s_q() const { return sm + GS_Q; }
  GMR_FN R* s_sd() const { return sm + dm.o_sd; }
  GMR_FN R* s_tg() const { return sm + dm.o_tg; }
  GMR_FN float* s_in() const { return reinterpret_cast<float*>(sm + dm.o_in); }
  GMR_FN float* s_inq() const { return s_in() + gmr_in_quat(dm.nhum); }
  GMR_FN R* s_bar() const { return sm + GS_BAR; }
  enum { PX = GmrLay<R>::PX, SD = GmrLay<R>::SD, TG = GmrLay<R>::TG, MT = GmrLay<R>::MT };
  GMR_FN R* s_xp() const { return sm + dm.o_xp; }   // body poses: record b at PX * b, position here, orientation at s_xq
  GMR_FN R* s_xq() const { return sm + dm.o_xq; }
  GMR_FN R* s_U() const { return sm + GS_VAR; }
  GMR_FN R* s_at(int off) const { return sm + off; }
  GMR_FN R* s_mt() const { return s_U(); }      // task inertias, dead once the rows are built
  GMR_FN R* s_L() const { return s_U(); }       // packed factor rows
  enum { ROWH = GmrLay<R>::ROWH, ROWS = GMR_NS + GmrLay<R>::ROWH };
  // slot 0 of the published row of hinge k; its header (y_k, 1/d_k) sits at [-ROWH], [-ROWH + 1]
  GMR_FN R* Lrow(int k) const {
#ifdef GMR_EMULATE
    return sm + (GS_VAR + ROWH) + ROWS * k;
#else
    extern __shared__ __align__(128) unsigned char gmr_dyn_smem[];
    return reinterpret_cast<R*>(gmr_dyn_smem + (sm_off + (uint32_t)k * (uint32_t)(ROWS * sizeof(R)))) + (GS_VAR + ROWH);   // one IMAD + an immediate
#endif
  }
#ifndef GMR_EMULATE
  #undef sm
#endif

  // ------------------------------------------------------------------ configuration --------
  template <typename S> GMR_FN void set_qpos(const S* src) {
    GMR_LANES
      for (int i = lane; i < dm.nq; i += 32) s_q()[i] = R(src[i]);
    GMR_END
  }

  // forward kinematics of s_q -> s_xp (root-relative positions), s_xq (world orientations),
  // then the world-frame spatial axis of every hinge at the root origin -> s_sd[j] = (v, w).
  //
  // A pointer-jumping scan over the tree of MOVING bodies instead of a walk by tree level: lane j < nh owns hinge j's body,
  // lane nh the floating root; fixed bodies in between are folded into the records on the host (gmr_fill_consts).  Every lane
  // starts from its pose in its moving parent's frame, (P, cos(q/2) A + sin(q/2) B), and in step k composes the pose its
  // ancestor at distance 2^k has accumulated so far in front of its own: after ceil(log2(depth + 1)) steps (4 for G1's 10
  // hinges + root) every lane holds its world pose - 4 dependent compositions on 30 lanes instead of 11 on 2-6.  The
  // running pose stays in registers; each step reads the ancestors' poses from one buffer and publishes its own into the other.  Fixed bodies (hands, head, sensors: task frames and the epilogue need them) follow in one block.
  GMR_FN void fk() {
    R q[4] = {R(1), R(0), R(0), R(0)}, p[3] = {R(0), R(0), R(0)};
#ifdef GMR_EMULATE
    R eq[32][4], ep[32][3];
#define GMR_FK_LOAD for (int i_ = 0; i_ < 4; i_++) q[i_] = eq[lane][i_]; for (int i_ = 0; i_ < 3; i_++) p[i_] = ep[lane][i_];
#define GMR_FK_SAVE for (int i_ = 0; i_ < 4; i_++) eq[lane][i_] = q[i_]; for (int i_ = 0; i_ < 3; i_++) ep[lane][i_] = p[i_];
#else
#define GMR_FK_LOAD
#define GMR_FK_SAVE
#endif
    // The scan ping-pongs between the pose array and the task-block region (dead during FK): a step reads one and writes the
    // other, so it is ONE lane block (one __syncwarp) instead of a compose block and a publish block; the start buffer is
    // chosen so that the last step lands in the pose array.
    int nsteps = 0;
    for (int k = 0; k < 4; k++) if ((1 << k) <= dm.maxd) nsteps++;      // a chain of maxd hinges + the root needs 2^K >= maxd + 1
    int cur = (nsteps & 1) ? (int)GS_VAR : dm.o_xq;
    GMR_LANES
      if (lane < dm.nh) {
        R s, c; g_sincos(R(0.5) * s_q()[7 + lane], &s, &c);
        R A[4], B[4], P[4];
        g_ld4(mc.fk_A + 4 * lane, A); g_ld4(mc.fk_B + 4 * lane, B); g_ld4(mc.fk_pos + 4 * lane, P);
        q[0] = c * A[0] + s * B[0]; q[1] = c * A[1] + s * B[1]; q[2] = c * A[2] + s * B[2]; q[3] = c * A[3] + s * B[3];
        p[0] = P[0]; p[1] = P[1]; p[2] = P[2];
        R* o = s_at(cur) + PX * mc.hbody[lane];
        g_st4(o, q[0], q[1], q[2], q[3]);
        g_st4(o + 4, p[0], p[1], p[2], R(0));
      } else if (lane == dm.nh) {
        q[0] = s_q()[3]; q[1] = s_q()[4]; q[2] = s_q()[5]; q[3] = s_q()[6];
        q_normalize(q);
        p[0] = p[1] = p[2] = R(0);
        g_st4(s_at(cur), q[0], q[1], q[2], q[3]);
        g_st4(s_at(cur) + 4, R(0), R(0), R(0), R(0));
        g_st4(s_rq(), q[0], q[1], q[2], q[3]);
      }
      GMR_FK_SAVE
    GMR_END
    GMR_NOUNROLL                                                      // one copy of the step in the instruction stream
    for (int k = 0; k < nsteps; k++) {
      const int nxt = cur == dm.o_xq ? (int)GS_VAR : dm.o_xq;
      GMR_LANES
        GMR_FK_LOAD
        if (lane <= dm.nh) {
          int b = 0;                                                    // the root (lane nh) only copies itself across
          if (lane < dm.nh) {
            b = mc.hbody[lane];
            const int pb = (int)((mc.fk_idx[lane] >> (8 * k)) & 0xffu);
            if (pb != 0xff) {
              R pq[4], pp[4];
              g_ld4(s_at(cur) + PX * pb, pq); g_ld4(s_at(cur) + PX * pb + 4, pp);
              R off[3]; q_rot(pq, p, off);
              p[0] = pp[0] + off[0]; p[1] = pp[1] + off[1]; p[2] = pp[2] + off[2];
              R qn[4]; q_mul(pq, q, qn);
              // no renormalisation along the chain: the root quaternion is normalised exactly at the top of every FK and
              // each product of unit quaternions moves |q| by about one ulp
              q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; q[3] = qn[3];
            }
          }
          R* o = s_at(nxt) + PX * b;
          g_st4(o, q[0], q[1], q[2], q[3]);
          g_st4(o + 4, p[0], p[1], p[2], R(0));
        }
        GMR_FK_SAVE
      GMR_END
      cur = nxt;
    }
    GMR_LANES
      GMR_FK_LOAD
      if (lane < dm.nh) {
        R w[3]; q_rot(q, mc.axis + 3 * lane, w);
        R* o = s_sd() + SD * lane;
        // linear velocity of the reference point (root origin) under unit joint rate: w x (0 - d) = d x w
        g_st4(o, p[1] * w[2] - p[2] * w[1], p[2] * w[0] - p[0] * w[2], p[0] * w[1] - p[1] * w[0], w[0]);
        g_st2(o + 4, w[1], w[2]);
      }
      for (int e = dm.nh + lane; e < dm.nb - 1; e += 32) {               // fixed bodies: constant pose in a moving body's frame
        const uint32_t ix = mc.fk_idx[e];
        const int b = ix & 0xffu, ab = (ix >> 8) & 0xffu;
        R aq[4], ap[4], C[4], Cp[4];
        g_ld4(s_xq() + PX * ab, aq); g_ld4(s_xp() + PX * ab, ap);
        g_ld4(mc.fk_A + 4 * e, C); g_ld4(mc.fk_pos + 4 * e, Cp);
        R off[3]; q_rot(aq, Cp, off);
        R qf[4]; q_mul(aq, C, qf);
        g_st4(s_xp() + PX * b, ap[0] + off[0], ap[1] + off[1], ap[2] + off[2], R(0));
        g_st4(s_xq() + PX * b, qf[0], qf[1], qf[2], qf[3]);
      }
    GMR_END
#undef GMR_FK_LOAD
#undef GMR_FK_SAVE
  }

  // ------------------------------------------------------------------ targets (A1-A5) ------
  // Raw keypoints of one frame (staged by stage_frame: positions at s_in + in_off, quaternions at s_inq, float) ->
  // scaled + offset targets in s_tg[h] = (pos[3], pad, quat[4]).
  // Returns false (uniform) when a keypoint is not finite or a quaternion has no direction: the reference would raise
  // inside scipy / mink on such a frame (and its dataset script would skip the file); the caller stops the clip.
  GMR_FN bool update_targets(R ratio, bool to_ground) {
    bool ok = true;
    GMR_LANES
      if (lane == dm.hroot) { const float* in = s_in() + in_off + 3 * lane; s_red()[0] = R(in[0]); s_red()[1] = R(in[1]); s_red()[2] = R(in[2]); }
    GMR_END
    GMR_LANES
      bool fin = true;
      if (lane < dm.nhum) {
        const R rx = s_red()[0], ry = s_red()[1], rz = s_red()[2];
        const float* in = s_in() + in_off + 3 * lane;
        const float* iq = s_inq() + 4 * lane;
        const R sr = mc.hscale[dm.hroot] * ratio;
        R p[3];
        if (lane == dm.hroot) { p[0] = sr * rx; p[1] = sr * ry; p[2] = sr * rz; }
        else {
          const R s = mc.hscale[lane] * ratio;
          p[0] = (R(in[0]) - rx) * s + sr * rx; p[1] = (R(in[1]) - ry) * s + sr * ry; p[2] = (R(in[2]) - rz) * s + sr * rz;
        }
        R q[4] = {R(iq[0]), R(iq[1]), R(iq[2]), R(iq[3])};
        const R qn2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        // NaN fails every comparison; float32 inputs cannot overflow R here
        fin = qn2 > R(1e-30) && qn2 < R(INFINITY) && (g_abs(p[0]) + g_abs(p[1]) + g_abs(p[2])) < R(INFINITY);
        R n = R(1) / g_sqrt(qn2);
        q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
        R u[4]; q_mul(q, mc.hroff + 4 * lane, u);
        n = R(1) / g_sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
        u[0] *= n; u[1] *= n; u[2] *= n; u[3] *= n;
        R g[3]; q_rot(u, mc.hpoff + 3 * lane, g);
        R* o = s_tg() + TG * lane;
        g_st4(o, p[0] + g[0], p[1] + g[1], p[2] + g[2], R(0));
        g_st4(o + 4, u[0], u[1], u[2], u[3]);
      }
#ifdef GMR_EMULATE
      if (!fin) ok = false;
#else
      ok = __all_sync(0xffffffffu, fin);
#endif
    GMR_END
    if (to_ground) {      // offset_human_data_to_ground, motion_retarget.py:252-270
      R lowest = R(INFINITY);
      for (int h = 0; h < dm.nhum; h++) if ((dm.foot_mask >> h) & 1u) { R z = s_tg()[TG * h + 2]; if (z < lowest) lowest = z; }
      GMR_SYNC();
      GMR_LANES
        if (lane < dm.nhum) s_tg()[TG * lane + 2] = s_tg()[TG * lane + 2] - lowest + R(0.1);
      GMR_END
    }
    return ok;
  }

  // ------------------------------------------------------------------ tasks (A7, A9, A10) ---
  // Split in two so that the work a stage decision does not need is not done:
  //   task_error()      e = log(T_b^-1 T_t) = (rho, omega) per task -> s_red[t] = |e|^2 (unweighted, all tasks);
  //                     the lane keeps (omega, rho, theta^2, |v|^2, cos(theta/2), cV) and the body pose it used.
  //   task_build(stage) only when an IK step follows: P = Jinv(omega) R_b^T;  K' = -Jinv Q Jinv R_b^T - P [d_b]x
  //                     (the position rows' dependence on rotation, shifted to the root origin); then with the
  //                     stage's weights s_mt[t] = M_t packed upper (21) | g_t (6) | mu_t, where
  //   M = [[wp2 P^T P, wp2 P^T K'], [., wp2 K'^T K' + wr2 P^T P]],
  //   g = A'^T W e = -[wp2 P^T rho ; wp2 K'^T rho + wr2 P^T omega],  mu = lm (wp2 |rho|^2 + wr2 |omega|^2).
  // Per frame that is (solves + stages) error evaluations but only (solves) builds, and a stage switch
  // re-sums the same |e_t|^2 with the other stage's task mask instead of re-evaluating anything.
  GMR_FN void task_error() {
    GMR_LANES
      if (lane < dm.nt) {
        const int b = mc.tbody[lane], h = mc.thuman[lane];
        const R* tg = s_tg() + TG * h;
        R qb[4]; g_ld4(s_xq() + PX * b, qb);
        R d[4]; g_ld4(s_xp() + PX * b, d);
        R qi[4] = {qb[0], -qb[1], -qb[2], -qb[3]};
        R qt[4]; g_ld4(tg + 4, qt);
        R qe[4]; q_mul(qi, qt, qe);
        if (qe[0] < R(0)) { qe[0] = -qe[0]; qe[1] = -qe[1]; qe[2] = -qe[2]; qe[3] = -qe[3]; }
        // world offset target - body, via root-relative coordinates
        R dw[3] = {(tg[0] - s_q()[0]) - d[0], (tg[1] - s_q()[1]) - d[1], (tg[2] - s_q()[2]) - d[2]};
        R tb[3]; q_rot_inv(qb, dw, tb);
        // SO(3) log
        const R nsq = qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3];
        R fac;
        if (nsq < ks.lie) fac = R(2) / qe[0] - R(2) / R(3) * nsq / (qe[0] * qe[0] * qe[0]);
        else if (nsq < R(0.0025) * qe[0] * qe[0]) {
          // theta / 2 = atan(t), t = n / w < 0.05 (a task that has nearly converged - the common case): 2 atan(t) / n =
          // (2 / w) (1 - t^2/3 + t^4/5 - ...), truncated below 1e-22; the library atan2 is ~130 instructions
          const R iw = R(1) / qe[0], t2 = nsq * iw * iw;
          fac = R(2) * iw * (R(1) + t2 * (R(-1.0 / 3.0) + t2 * (R(1.0 / 5.0) + t2 * (R(-1.0 / 7.0) + t2 * (R(1.0 / 9.0) + t2 * (R(-1.0 / 11.0) + t2 * (R(1.0 / 13.0) + t2 * R(-1.0 / 15.0))))))));
        }
        else { const R n = g_sqrt_pos(nsq); fac = R(2) * g_atan2(n, qe[0]) / n; }
        R om[3] = {fac * qe[1], fac * qe[2], fac * qe[3]};
        const R th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
        // Jinv = I - S/2 + cV S^2, cV = (1 - (theta/2) cot(theta/2)) / theta^2 (also V^-1 of SE3.log)
        R cV;
        if (th2 < ks.lie) cV = R(1) / R(12);
        else if (th2 < R(0.0625)) cV = R(1) / R(12) + th2 * (R(1) / R(720) + th2 * (R(1) / R(30240) + th2 * R(1.0 / 1209600.0)));   // series: the closed form cancels badly for small theta
        else cV = (R(1) - R(0.5) * g_sqrt_pos(th2) * qe[0] / g_sqrt_pos(nsq)) / th2;
        // rho = Jinv tb = tb - (omega x tb) / 2 + cV omega x (omega x tb)
        R c1[3] = {om[1] * tb[2] - om[2] * tb[1], om[2] * tb[0] - om[0] * tb[2], om[0] * tb[1] - om[1] * tb[0]};
        R c2[3] = {om[1] * c1[2] - om[2] * c1[1], om[2] * c1[0] - om[0] * c1[2], om[0] * c1[1] - om[1] * c1[0]};
        R rho[3] = {tb[0] - R(0.5) * c1[0] + cV * c2[0], tb[1] - R(0.5) * c1[1] + cV * c2[1], tb[2] - R(0.5) * c1[2] + cV * c2[2]};
        s_red()[lane] = rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2] + th2;
        // hand-over to task_build(): parked in the task's own (still unused) M_t block rather than in registers
        R* o = s_mt() + MT * lane;
        g_st4(o, om[0], om[1], om[2], th2); g_st4(o + 4, rho[0], rho[1], rho[2], nsq);
        g_st4(o + 8, d[0], d[1], d[2], qe[0]); g_st4(o + 12, qb[0], qb[1], qb[2], qb[3]);
        o[16] = cV;
      }
    GMR_END
  }

  GMR_FN void task_build(int stage) {
    const uint32_t stage_mask = stage == 0 ? dm.in1_mask : dm.in2_mask;
    GMR_LANES
      if (lane < dm.nt) {
        R om[4], rho[4], d[4], qb[4], P[9], K[9];
        const R* in = s_mt() + MT * lane;
        g_ld4(in, om); g_ld4(in + 4, rho); g_ld4(in + 8, d); g_ld4(in + 12, qb);
        const R th2 = om[3], nsq = rho[3], cw = d[3], cV = in[16];
        {
          // Bq, Cq, Dq: Barfoot's Q coefficients
          R Bq, Cq, Dq;
          if (th2 < ks.lie) {
            Bq = R(1) / R(6); Cq = -R(1) / R(24); Dq = R(1) / R(120);
          } else if (th2 < R(0.0625)) {                    // series: the closed forms cancel badly for small theta
            Bq = R(1) / R(6) - th2 * (R(1) / R(120) - th2 * (R(1) / R(5040) - th2 * R(1.0 / 362880.0)));
            Cq = -R(1) / R(24) + th2 * (R(1) / R(720) - th2 * (R(1) / R(40320) - th2 * R(1.0 / 3628800.0)));
            Dq = R(1) / R(120) - th2 * (R(1) / R(2520) - th2 * (R(1) / R(120960) - th2 * R(1.0 / 9979200.0)));
          } else {
            const R th = g_sqrt(th2);
            const R n = g_sqrt(nsq);                       // sin(theta/2); cw = cos(theta/2)
            const R st = R(2) * n * cw, ct = cw * cw - nsq;
            Bq = (th - st) / (th2 * th);
            Cq = (R(1) - th2 * R(0.5) - ct) / (th2 * th2);
            Dq = (R(2) * th - R(3) * st + th * ct) / (R(2) * th2 * th2 * th);
          }
          // With W = [omega]x, V = [rho]x, wr = omega . rho, c = rho x omega:
          //   W^2 = omega omega^T - theta^2 I,   V W = omega rho^T - wr I,   V W W = omega c^T - wr W,
          // so Barfoot's Q = V/2 + Bq (W V + V W + W V W) - Cq (V W W - W W V - 3 W V W) - Dq (W V W W + W W V W), with
          // W V W = -wr W and W V W W + W W V W = -2 wr W^2, is a symmetric part plus ONE skew matrix:
          //   Q = Bq (omega rho^T + rho omega^T) - 2 Dq wr omega omega^T + 2 wr (Dq theta^2 - Bq) I  +  [a]x,
          //   a = rho / 2 - (Bq + Cq) wr omega - Cq (c x omega)        (omega c^T - c omega^T = [c x omega]x)
          // ~65 multiply-adds instead of two 3x3 products and a nine-entry combination.
          R S[9], S2[9], Ji[9];
          m3_skew(om, S);
          GMR_UNROLL
          for (int i = 0; i < 3; i++) {
            GMR_UNROLL
            for (int j = 0; j < 3; j++) S2[3 * i + j] = om[i] * om[j] - (i == j ? th2 : R(0));
          }
          GMR_UNROLL
          for (int i = 0; i < 9; i++) Ji[i] = ((i & 3) == 0 ? R(1) : R(0)) - R(0.5) * S[i] + cV * S2[i];
          R Q[9];
          {
            const R wr = om[0] * rho[0] + om[1] * rho[1] + om[2] * rho[2];
            const R c[3] = {rho[1] * om[2] - rho[2] * om[1], rho[2] * om[0] - rho[0] * om[2], rho[0] * om[1] - rho[1] * om[0]};
            const R cw[3] = {c[1] * om[2] - c[2] * om[1], c[2] * om[0] - c[0] * om[2], c[0] * om[1] - c[1] * om[0]};
            const R k1 = (Bq + Cq) * wr, k2 = R(2) * Dq * wr, kd = R(2) * wr * (Dq * th2 - Bq);
            const R a[3] = {R(0.5) * rho[0] - k1 * om[0] - Cq * cw[0], R(0.5) * rho[1] - k1 * om[1] - Cq * cw[1], R(0.5) * rho[2] - k1 * om[2] - Cq * cw[2]};
            R A[9]; m3_skew(a, A);
            GMR_UNROLL
            for (int i = 0; i < 3; i++) {
              GMR_UNROLL
              for (int j = 0; j < 3; j++)
                Q[3 * i + j] = Bq * (om[i] * rho[j] + om[j] * rho[i]) - k2 * (om[i] * om[j]) + (i == j ? kd : R(0)) + A[3 * i + j];
            }
          }
          if (th2 < ks.lie) {
            // mink's SE3 Jacobian shortcut: for theta^2 < eps, jlog(T) is taken to be the 6x6 identity
            // (the rho-dependent block Q = [rho]x / 2 is dropped).  Reproduced for parity.
            GMR_UNROLL
            for (int i = 0; i < 9; i++) { Ji[i] = ((i & 3) == 0) ? R(1) : R(0); Q[i] = R(0); }
          }
          R Rb[9]; q_to_mat(qb, Rb);
          m3_mul_bt(Ji, Rb, P);                 // P = Jinv R_b^T
          R T1[9], T2[9];
          m3_mul(Ji, Q, T1); m3_mul(T1, P, T2); // Jinv Q Jinv R_b^T
          R Sd[9], PS[9]; m3_skew(d, Sd); m3_mul(P, Sd, PS);
          GMR_UNROLL
          for (int i = 0; i < 9; i++) K[i] = -T2[i] - PS[i];
        }
        const R r2 = rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2];
        R* o = s_mt() + MT * lane;
        const bool on = (stage_mask >> lane) & 1u;
        const R* wtab = stage == 0 ? mc.w1 : mc.w2;
        const R wp = on ? wtab[2 * lane] : R(0), wr = on ? wtab[2 * lane + 1] : R(0);
        const R wp2 = wp * wp, wr2 = wr * wr;
        R blk[28];                                            // assembled in registers (static indices), stored as 128-bit units
        {
          R PtP[9], PtK[9], KtK[9];
          m3_mul_at(P, P, PtP); m3_mul_at(P, K, PtK); m3_mul_at(K, K, KtK);
          GMR_UNROLL
          for (int i = 0; i < 3; i++) {
            GMR_UNROLL
            for (int j = 0; j < 3; j++) {
              if (j >= i) { blk[gmr_sym6(i, j)] = wp2 * PtP[3 * i + j]; blk[gmr_sym6(3 + i, 3 + j)] = wp2 * KtK[3 * i + j] + wr2 * PtP[3 * i + j]; }
              blk[gmr_sym6(i, 3 + j)] = wp2 * PtK[3 * i + j];
            }
          }
        }
        GMR_UNROLL
        for (int i = 0; i < 3; i++) {
          const R ptr = P[i] * rho[0] + P[3 + i] * rho[1] + P[6 + i] * rho[2];
          const R ktr = K[i] * rho[0] + K[3 + i] * rho[1] + K[6 + i] * rho[2];
          const R pto = P[i] * om[0] + P[3 + i] * om[1] + P[6 + i] * om[2];
          blk[21 + i] = -(wp2 * ptr);
          blk[24 + i] = -(wp2 * ktr + wr2 * pto);
        }
        blk[27] = ks.lm * (wp2 * r2 + wr2 * th2);
        GMR_UNROLL
        for (int c = 0; c < 7; c++) g_st4(o + 4 * c, blk[4 * c], blk[4 * c + 1], blk[4 * c + 2], blk[4 * c + 3]);
      }
    GMR_END
  }

  // unweighted error norm of the stage's tasks (error1()/error2(), motion_retarget.py:188-200):
  // lane t holds |e_t|^2 in s_red[t]; butterfly sum over the lanes of the stage's tasks
  GMR_FN R stage_error(int stage) const {
    const uint32_t m = stage == 0 ? dm.in1_mask : dm.in2_mask;
#ifdef GMR_EMULATE
    R part[32];
    for (int t = 0; t < 32; t++) part[t] = (t < dm.nt && ((m >> t) & 1u)) ? s_red()[t] : R(0);
    for (int o = 16; o > 0; o >>= 1) for (int t = 0; t < 32; t++) if ((t & o) == 0) { const R a = part[t], b = part[t ^ o]; part[t] = part[t ^ o] = a + b; }
    return g_sqrt(part[0]);
#else
    R v = (lane_ < dm.nt && ((m >> lane_) & 1u)) ? s_red()[lane_] : R(0);
    GMR_UNROLL
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return g_sqrt(v);
#endif
  }

  // ------------------------------------------------------------------ composites ------------
  // whole-tree composite for the floating base, per-hinge composite -> f_i = Ic_i s_i,
  // c_i = s_i . G_i, diag_i = s_i . f_i + damping + sum mu
  GMR_FN void composites() {
    // task-subtree composites in place (children before parents; lane e owns element e of every block, so
    // there is no cross-lane hazard), then the whole-tree composite = sum over the root tasks
    // lane = (pair q, 128-bit chunk c) of a round: 4 pairs x 7 chunks
    for (int r = 0; r < mc.ntround; r++) {
      GMR_LANES
        if (lane < 28) {
          const int q = lane / 7, c = lane - 7 * q;
          const uint32_t tp = mc.tround[4 * r + q];
          if (tp != 0xffffu) {                               // unused slot of the round
            R* dst = s_mt() + MT * (tp >> 8);
            const R* src = s_mt() + MT * (tp & 0xffu);
            if constexpr (sizeof(R) == 8) {
              // a lane's four elements as two 16-byte units that are 7 units apart: the seven lanes of a pair then
              // touch consecutive units in each access (4 c .. 4 c + 3 would put lanes c and c + 4 on the same banks)
              R a[2], b[2], a2[2], b2[2];
              g_ld2(src + 2 * c, a); g_ld2(dst + 2 * c, b); g_ld2(src + 14 + 2 * c, a2); g_ld2(dst + 14 + 2 * c, b2);
              g_st2(dst + 2 * c, a[0] + b[0], a[1] + b[1]); g_st2(dst + 14 + 2 * c, a2[0] + b2[0], a2[1] + b2[1]);
            } else {
              R a[4], b[4];
              g_ld4(src + 4 * c, a); g_ld4(dst + 4 * c, b);
              g_st4(dst + 4 * c, a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]);
            }
          }
        }
      GMR_END
    }
    GMR_LANES
      if (lane < 28) {
        R s = R(0);
        uint32_t rm = mc.troot_mask;
        while (rm) { const int t = GMR_CTZ(rm); rm &= rm - 1u; s += s_mt()[MT * t + lane]; }
        s_root()[lane] = s;
      }
    GMR_END
    GMR_LANES
      if (lane < dm.nh) {
        R acc[28];
        // every lane sums the task-subtree composites of its OWN top tasks (set bits of its mask; one for
        // most hinges of a humanoid); the trip count is the longest such list of the robot
        uint32_t tm = mc.top_mask[lane];
        if (tm) {                                            // the first (usually only) one is a plain load
          const R* m = s_mt() + MT * GMR_CTZ(tm);
          tm &= tm - 1u;
          GMR_UNROLL
          for (int c = 0; c < 7; c++) g_ld4(m + 4 * c, acc + 4 * c);
        } else {
          GMR_UNROLL
          for (int i = 0; i < 28; i++) acc[i] = R(0);
        }
        for (int it = 1; it < dm.tlmax; it++) {
          if (tm) {
            const int t = GMR_CTZ(tm);
            tm &= tm - 1u;
            const R* m = s_mt() + MT * t;
            GMR_UNROLL
            for (int c = 0; c < 7; c++) {
              R v[4]; g_ld4(m + 4 * c, v);
              acc[4 * c] += v[0]; acc[4 * c + 1] += v[1]; acc[4 * c + 2] += v[2]; acc[4 * c + 3] += v[3];
            }
          }
        }
        R sv[6]; g_ld4(s_sd() + SD * lane, sv); g_ld2(s_sd() + SD * lane + 4, sv + 4);
        R dg = R(0), ci = R(0);
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          R v = R(0);
          GMR_UNROLL
          for (int j = 0; j < 6; j++) v += acc[gmr_sym6(i, j)] * sv[j];
          lp(LP_F + i, lane) = v; dg += v * sv[i]; ci += acc[21 + i] * sv[i];
        }
        lp(LP_DIAG, lane) = dg + ks.damping + s_root()[27];
        lp(LP_CI, lane) = ci;
      }
    GMR_END
  }

  GMR_FN R bound_hi(int j) const { return ((dm.limited_mask >> j) & 1u) ? ks.gain * (mc.hi[j] - s_q()[7 + j]) : R(INFINITY); }
  GMR_FN R bound_lo(int j) const { return ((dm.limited_mask >> j) & 1u) ? -(ks.gain * (s_q()[7 + j] - mc.lo[j])) : -R(INFINITY); }

  // ------------------------------------------------------------------ rows of H -------------
  // H_ij = f_i . s_j is non-zero only when j is an ancestor of i (or i of j): branch-induced sparsity.
  // Lane i keeps its row by SLOT: row[0..5] = f_i (the base columns are unit twists at the root origin,
  // world-aligned), row[5 + d] = f_i . s_(ancestor of i at depth d) for d < depth_i, dg = diagonal,
  // rhs = -c_i.  Every lane defines ALL of its row registers here (lanes without a hinge get zeros): a
  // full redefinition is what tells the compiler that row[] is dead between two solves.
  GMR_FN void build_rows() {
    GMR_LANES
      {
        const bool act = lane < dm.nh;
        const int dep = GMR_DEPTH;
        R f[6];
        GMR_UNROLL
        for (int g = 0; g < 6; g++) { f[g] = act ? lp(LP_F + g, lane) : R(0); L.row[g] = f[g]; }
        // Branch-free across the lanes: a lane without an ancestor at depth d reads its own axis (the table holds its own
        // index there) and discards the product.  The depths then sit in two straight-line blocks - 1..6, and 7..9 for
        // robots with chains deeper than 7 hinges - whose load -> FMA chains overlap instead of running one after the
        // other behind branches (a warp issues in order; measured on a lone slow clip: 5.7 % of its time was here).
        static_assert(GMR_MAXD == 10, "the depth split below assumes ten slots");
        GMR_UNROLL
        for (int d = 1; d <= 6; d++) {
          const int j = mc.anc_of[lane * GMR_ANCS + d - 1];
          R s[6]; g_ld4(s_sd() + SD * j, s); g_ld2(s_sd() + SD * j + 4, s + 4);
          const R t = f[0] * s[0] + f[1] * s[1] + f[2] * s[2] + f[3] * s[3] + f[4] * s[4] + f[5] * s[5];
          L.row[5 + d] = d < dep ? t : R(0);
        }
        if (dm.maxd > 7) {                                       // uniform
          GMR_UNROLL
          for (int d = 7; d < GMR_MAXD; d++) {
            const int j = mc.anc_of[lane * GMR_ANCS + d - 1];
            R s[6]; g_ld4(s_sd() + SD * j, s); g_ld2(s_sd() + SD * j + 4, s + 4);
            const R t = f[0] * s[0] + f[1] * s[1] + f[2] * s[2] + f[3] * s[3] + f[4] * s[4] + f[5] * s[5];
            L.row[5 + d] = d < dep ? t : R(0);
          }
        } else {
          GMR_UNROLL
          for (int d = 7; d < GMR_MAXD; d++) L.row[5 + d] = R(0);
        }
        L.row[5 + GMR_MAXD] = R(0);                              // no hinge has a strict ancestor at depth GMR_MAXD
        L.dg = act ? lp(LP_DIAG, lane) : R(1);
        L.rhs = act ? -lp(LP_CI, lane) : R(0);
        L.dinv = R(1);
      }
    GMR_END
  }

  // stores this lane's row slots, scaled by `sc`, into its published row
  GMR_FN void publish_row(R* o, const LaneRegs<R>& L, R sc, int nchunk) const {
    GMR_UNROLL
    for (int c = 0; c < GMR_NS / 4; c++)
      if (c < nchunk) g_st4(o + 4 * c, L.row[4 * c] * sc, L.row[4 * c + 1] * sc, L.row[4 * c + 2] * sc, L.row[4 * c + 3] * sc);
  }

  // sum of term(p) over the set bits p of m, four terms at a time: their loads are issued together (a warp issues in order: a
  // plain loop pays one shared-memory round trip per bit)
  template <typename F> GMR_FN R sum_over_bits(uint32_t m, F term) const {
    R acc = R(0);
    while (m) {
      const int p0 = GMR_CTZ(m); m &= m - 1u;
      const bool h1 = m != 0u; const int p1 = h1 ? GMR_CTZ(m) : p0; m &= m - 1u;
      const bool h2 = m != 0u; const int p2 = h2 ? GMR_CTZ(m) : p0; m &= m - 1u;
      const bool h3 = m != 0u; const int p3 = h3 ? GMR_CTZ(m) : p0; m &= m - 1u;
      const R t0 = term(p0), t1 = term(p1), t2 = term(p2), t3 = term(p3);
      acc += t0; acc += h1 ? t1 : R(0); acc += h2 ? t2 : R(0); acc += h3 ? t3 : R(0);
    }
    return acc;
  }

  // Working set: a pinned hinge p keeps x_p = bound.  Its row becomes the identity with rhs = bound, its
  // column moves to the right-hand side of every row coupled to it (ancestors, descendants, base).
  GMR_FN void apply_pins(uint32_t pinned) {
    const int nchunk_all = dm.o_y >> 2;
    GMR_LANES
      if (lane < dm.nh && ((pinned >> lane) & 1u)) {
        R* o = Lrow(lane);
        publish_row(o, L, s_bnd()[lane], nchunk_all);               // (row of H) * bound
        g_st2(o - ROWH, R(0), R(0));                                  // header: "1 / d = 0" - no contribution to the base block
      }
    GMR_END
    // straight-line per lane (selects instead of branches: the nine depths' table look-ups, bound loads and FMAs overlap)
    GMR_LANES
      {
        const int dep = GMR_DEPTH;                                // 0 for a lane without a hinge: nothing below applies to it
        const bool mine = (pinned >> lane) & 1u;
        R r = L.rhs;
        GMR_UNROLL
        for (int d = 1; d < GMR_MAXD; d++) {                      // pinned ancestors: my own slots
          const int j = mc.anc_of[lane * GMR_ANCS + d - 1];
          const bool pj = d < dep && ((pinned >> j) & 1u);
          const R t = r - L.row[5 + d] * s_bnd()[j];              // s_bnd of an unpinned hinge is stale: computed, not selected
          r = pj ? t : r;
          L.row[5 + d] = (pj || mine) ? R(0) : L.row[5 + d];
        }
        // pinned descendants: their published rows
        r -= sum_over_bits(GMR_DESC & pinned, [&](int p) { return Lrow(p)[5 + dep]; });
        GMR_UNROLL
        for (int g = 0; g < 6; g++) L.row[g] = mine ? R(0) : L.row[g];
        L.row[5 + GMR_MAXD] = mine ? R(0) : L.row[5 + GMR_MAXD];
        L.dg = mine ? R(1) : L.dg;
        L.rhs = mine ? s_bnd()[lane] : r;
      }
      if (lane < 6)                                             // base right-hand side correction -> s_xs[0..5]
        s_xs()[lane] = sum_over_bits(pinned, [&](int p) { return Lrow(p)[lane]; });
    GMR_END
  }

  // one depth of the elimination: the hinges of depth lv publish their rows, their ancestors absorb them
  template <int NCH>
  GMR_FN void eliminate_depth(int lv, uint32_t pinned) {
    // a pinned hinge's row is the identity (apply_pins): eliminating it changes nothing, so it is skipped
    GMR_LANES
      if (GMR_DEPTH == lv && !((pinned >> lane) & 1u)) {
        const R di = g_rcp_pos(L.dg);
        L.dinv = di;
        R* o = Lrow(lane);
        publish_row(o, L, R(1), NCH);
        if constexpr (sizeof(R) == 8) g_st2(o - ROWH, L.rhs, di); else g_st4(o - ROWH, L.rhs, di, R(0), R(0));
      }
    GMR_END
    // float32: two pivots per pass, branch-free — every lane loads both published rows (broadcast reads) and absorbs
    // them with a multiplier that is zero unless the lane is an ancestor of that pivot; a lone pivot is paired with
    // itself at zero weight.  No divergent branch, the loads of the second pivot overlap the FMAs of the first
    // (lone warp -8 %).  float64 keeps the update behind a branch: its FMAs are the scarce resource (the FP64
    // pipe is half rate), and the paired form issues ~10 % more of them.
    uint32_t rem = mc.lvl_mask[lv] & ~pinned;
    if constexpr (sizeof(R) == 8) {
      // float64: every lane absorbs ITS OWN next pivot in a pass - pivots of one depth in different branches (left /
      // right leg, left / right arm) have different ancestors, so their row updates run side by side: for G1 16
      // passes instead of 29 (a lane still sees its pivots in ascending order: the results are unchanged bit for bit).
      // The base block's rank-1 updates of this depth follow in factor_solve (base_block_depth).
      bool more = false;
      GMR_LANES
        L.piv = GMR_DESC & rem;
#ifdef GMR_EMULATE
        if (L.piv) more = true;
#else
        more = __any_sync(0xffffffffu, L.piv != 0u);
#endif
      GMR_END_NOSYNC
      // (a table of pass counts per depth instead of the vote was tried: 2-3 % slower on 8192-clip batches)
      while (more) {
#ifdef GMR_EMULATE
        more = false;
#endif
        GMR_LANES
          if (L.piv) {
            const R* pk = Lrow(GMR_CTZ(L.piv));
            L.piv &= L.piv - 1u;
            R yd[2]; g_ld2(pk - ROWH, yd);                         // y_k, 1 / d_k
            const R hki = pk[5 + GMR_DEPTH];
            const R a = hki * yd[1];
            GMR_UNROLL
            for (int c = 0; c < NCH; c++) {
              R v[4]; g_ld4(pk + 4 * c, v);
              L.row[4 * c] -= a * v[0]; L.row[4 * c + 1] -= a * v[1]; L.row[4 * c + 2] -= a * v[2]; L.row[4 * c + 3] -= a * v[3];
            }
            L.dg -= a * hki;
            L.rhs -= a * yd[0];
          }
#ifdef GMR_EMULATE
          if (L.piv) more = true;
#else
          more = __any_sync(0xffffffffu, L.piv != 0u);
#endif
        GMR_END_NOSYNC
      }
    } else {
    while (rem) {
      const int k0 = GMR_CTZ(rem);
      rem &= rem - 1u;
      const bool two = rem != 0;
      const int k1 = two ? GMR_CTZ(rem) : k0;
      rem &= rem - 1u;                                               // (0 & anything) stays 0
      GMR_LANES
        const R* p0 = Lrow(k0);
        const R* p1 = Lrow(k1);
        const R y0 = p0[-ROWH], d0 = p0[-ROWH + 1], y1 = p1[-ROWH], d1 = two ? p1[-ROWH + 1] : R(0);
        const R h0 = p0[5 + GMR_DEPTH], h1 = p1[5 + GMR_DEPTH];
        const R a0 = ((GMR_DESC >> k0) & 1u) ? h0 * d0 : R(0), a1 = ((GMR_DESC >> k1) & 1u) ? h1 * d1 : R(0);
        GMR_UNROLL
        for (int c = 0; c < NCH; c++) {
          R v[4], w[4]; g_ld4(p0 + 4 * c, v); g_ld4(p1 + 4 * c, w);
          L.row[4 * c] -= a0 * v[0]; L.row[4 * c + 1] -= a0 * v[1]; L.row[4 * c + 2] -= a0 * v[2]; L.row[4 * c + 3] -= a0 * v[3];
          L.row[4 * c] -= a1 * w[0]; L.row[4 * c + 1] -= a1 * w[1]; L.row[4 * c + 2] -= a1 * w[2]; L.row[4 * c + 3] -= a1 * w[3];
        }
        L.dg -= a0 * h0; L.dg -= a1 * h1;
        L.rhs -= a0 * y0; L.rhs -= a1 * y1;
        // base block: S = A_bb - sum_k (1/d_k) h_k h_k^T,  b = b_b - sum_k (1/d_k) h_k y_k,  h_k = base slots of row k
        if (lane < 27) { L.sacc += p0[GMR_SR] * d0 * p0[GMR_SC]; L.sacc += p1[GMR_SR] * d1 * p1[GMR_SC]; }
      GMR_END_NOSYNC
    }
    }
  }

  // ------------------------------------------------------------------ factor + solve --------
  // Solves H x = rhs for the rows built above; result in s_xs[0..nv) (base first).
  //
  // Branch-sparse L^T D L factorisation (the elimination order of Featherstone's sparse joint-space
  // factorisation: leaves first, so no fill-in outside the ancestor pattern).  Hinges are eliminated by
  // DEPTH, deepest first; all hinges of one depth publish their rows at once (one __syncwarp per depth),
  // then every strict ancestor i of pivot k does  a = H_ki / d_k;  row_i[s] -= a * row_k[s]  over the
  // pivot's slots (the same slot is the same ancestor for both),  d_i -= a * H_ki,  rhs_i -= a * y_k
  // (the forward substitution rides along).  The 6 floating-base DoFs are the ancestors of everything:
  // their 6x6 Schur complement is accumulated afterwards by 27 lanes (21 entries + 6 right-hand sides) from
  // the published rows and solved densely; the back substitution walks the depths root -> leaves.
  // For G1: 10 depths instead of 29 sequential pivots, rows of <= 16 slots instead of 35.
  GMR_FN void factor_solve(uint32_t pinned) {
    const bool pinned_any = pinned != 0;
    // per-lane constants of this factorisation: my descendants, my entry (r, cc) of the base block
    // (lanes 0..20: packed upper triangle; lanes 21..26: the right-hand side, the row header's y)
    GMR_LANES
      L.sacc = R(0);
    GMR_END_NOSYNC
    {
      // slots [0, 5 + lv) = base + strict ancestors; the chunk count is a template argument so that shallow
      // depths do not issue (predicated-off) work for slots they do not have: depths 8.. use 4 chunks, 4..7 three, 1..3 two
      int lv = dm.maxd;
      for (; lv >= 8; lv--) eliminate_depth<GMR_NS / 4>(lv, pinned);
      for (; lv >= 4; lv--) eliminate_depth<3>(lv, pinned);
      for (; lv >= 1; lv--) eliminate_depth<2>(lv, pinned);
    }
    if constexpr (sizeof(R) == 8) {
      // base block: S = A_bb - sum_k (1/d_k) h_k h_k^T,  b = b_b - sum_k (1/d_k) h_k y_k,  h_k = base slots of the
      // published row k.  Every row is still in place after the last depth (pinned rows carry 1/d = 0), so one loop over
      // all rows does it: constant stride, immediate offsets, the loads of several rows in flight at once; two partial
      // sums halve the dependent chain.  Lanes >= 27 own "entry (0, 0)" and are never read.
      GMR_LANES
        const R* pr = Lrow(0) + GMR_SR;
        const R* pc = Lrow(0) + GMR_SC;
        const R* pd = Lrow(0) - ROWH + 1;
        R s0 = R(0), s1 = R(0);
        int k = 0;
        GMR_NOUNROLL
        for (; k + 4 <= dm.nh; k += 4) {
          const R a0 = pr[0] * pd[0], a1 = pr[ROWS] * pd[ROWS], a2 = pr[2 * ROWS] * pd[2 * ROWS], a3 = pr[3 * ROWS] * pd[3 * ROWS];
          s0 += a0 * pc[0]; s1 += a1 * pc[ROWS]; s0 += a2 * pc[2 * ROWS]; s1 += a3 * pc[3 * ROWS];
          pr += 4 * ROWS; pc += 4 * ROWS; pd += 4 * ROWS;
        }
        GMR_NOUNROLL
        for (; k < dm.nh; k++) { s0 += (pr[0] * pd[0]) * pc[0]; pr += ROWS; pc += ROWS; pd += ROWS; }
        L.sacc = s0 + s1;
      GMR_END_NOSYNC
    }
    GMR_LANES
      if (lane < 27) {
        const R* rt = s_root();
        R v;
        if (lane < 21) v = rt[lane] + ((GMR_SC == GMR_SR) ? ks.damping + rt[27] : R(0));
        else v = -rt[lane] - (pinned_any ? s_xs()[GMR_SR] : R(0));
        s_lf()[lane] = v - L.sacc;
      }
    GMR_END
    // dense 6x6 solve by lane 0: S = Lf Lf^T, x_b = Lf^-T Lf^-1 b
    GMR_LANES
      if (lane == 0) {
        const R* c = s_lf();
        R A[21];
        GMR_UNROLL
        for (int i = 0; i < 21; i++) A[i] = c[i];
        R y[6];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) y[i] = c[21 + i];
        R lf[21];                                              // row-wise lower, reciprocal diagonal
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          GMR_UNROLL
          for (int j = 0; j <= i; j++) {
            R s = A[gmr_sym6(j, i)];
            GMR_UNROLL
            for (int k = 0; k < j; k++) s -= lf[i * (i + 1) / 2 + k] * lf[j * (j + 1) / 2 + k];
            if (i == j) lf[i * (i + 1) / 2 + i] = g_rsqrt(s);
            else lf[i * (i + 1) / 2 + j] = s * lf[j * (j + 1) / 2 + j];
          }
          R s = y[i];
          GMR_UNROLL
          for (int k = 0; k < i; k++) s -= lf[i * (i + 1) / 2 + k] * y[k];
          y[i] = s * lf[i * (i + 1) / 2 + i];
        }
        R x[6];
        GMR_UNROLL
        for (int ii = 0; ii < 6; ii++) {
          const int i = 5 - ii;
          R s = y[i];
          GMR_UNROLL
          for (int k = i + 1; k < 6; k++) s -= lf[k * (k + 1) / 2 + i] * x[k];
          x[i] = s * lf[i * (i + 1) / 2 + i];
        }
        GMR_UNROLL
        for (int i = 0; i < 6; i++) s_xs()[i] = x[i];
      }
    GMR_END
    // back substitution, root -> leaves:  x_k = (y_k - sum_s row_k[s] x_(anc at slot s)) / d_k.  The ancestors' values travel by
    // warp shuffle (each lane names its ancestor at the level's depth; a lane that has none names itself and discards the
    // value): one multiply, one shuffle and one FMA per level instead of a store, a __syncwarp, a table look-up and a load.
    {
#ifdef GMR_EMULATE
      R xv[32], xmine[32];
      for (int i = 0; i < 32; i++) xmine[i] = R(0);
#else
      R xmine = R(0);
#endif
#ifndef GMR_EMULATE
      uint32_t aw[3];                                           // this lane's ancestors by depth, one byte each
#endif
      GMR_LANES
        R acc = L.rhs;
        GMR_UNROLL
        for (int c = 0; c < 6; c++) acc -= L.row[c] * s_xs()[c];
        L.rhs = acc;
#ifndef GMR_EMULATE
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(mc.anc_of + lane * GMR_ANCS);
        aw[0] = ap[0]; aw[1] = ap[1]; aw[2] = ap[2];
#endif
      GMR_END_NOSYNC
      GMR_UNROLL
      for (int lv = 1; lv <= GMR_MAXD; lv++) {
        if (lv <= dm.maxd) {
#ifdef GMR_EMULATE
          GMR_LANES
            xv[lane] = L.rhs * L.dinv;
            if (GMR_DEPTH == lv) xmine[lane] = xv[lane];
          GMR_END_NOSYNC
          GMR_LANES
            const R xa = xv[mc.anc_of[lane * GMR_ANCS + lv - 1]];
            if (GMR_DEPTH > lv) L.rhs -= L.row[5 + lv] * xa;
          GMR_END_NOSYNC
#else
          GMR_LANES
            const R x = L.rhs * L.dinv;                             // final on the lanes of depth lv
            xmine = GMR_DEPTH == lv ? x : xmine;
            const R xa = __shfl_sync(0xffffffffu, x, (int)((aw[(lv - 1) >> 2] >> (8 * ((lv - 1) & 3))) & 0xffu));
            L.rhs = GMR_DEPTH > lv ? L.rhs - L.row[5 + lv] * xa : L.rhs;
          GMR_END_NOSYNC
#endif
        }
      }
      GMR_LANES
#ifdef GMR_EMULATE
        if (lane < dm.nh) s_xs()[6 + lane] = xmine[lane];
#else
        if (lane < dm.nh) s_xs()[6 + lane] = xmine;
#endif
      GMR_END
    }
  }

  // ------------------------------------------------------------------ box QP (A10, A11) -----
  // min 1/2 x^T H x + c^T x, lo <= x_hinge <= hi: primal active set from a feasible start.
  // The step is left in s_xs[0..nv).  One code instance of build_rows/factor_solve: the loop
  // alternates between SOLVE passes and (only when bounds are pinned) a CHECK pass that
  // evaluates the KKT multipliers on the original rows.
  GMR_FN void solve_qp(int stage) {
    const R INF = R(INFINITY);
    const uint32_t w_lo = stage == 0 ? warm_lo0 : warm_lo1, w_hi = stage == 0 ? warm_hi0 : warm_hi1;
    GMR_LANES
      if (lane < dm.nh) {
        // feasible start: bounds that were active at the end of the previous solve of this stage stay in
        // the working set (their joints sit on, or creep towards, the limit), everything else at
        // clip(0, lo, hi).  Any feasible start gives the same (unique) optimum.  The two stages weigh the
        // tasks differently and settle on different working sets, so each keeps its own.
        const R bhi = bound_hi(lane), blo = bound_lo(lane);
        R x;
        if ((w_hi >> lane) & 1u) x = bhi;
        else if ((w_lo >> lane) & 1u) x = blo;
        else x = R(0) < blo ? blo : (R(0) > bhi ? bhi : R(0));
        lp(LP_X, lane) = x;
      }
    GMR_END
    uint32_t pin_lo = w_lo & dm.limited_mask, pin_hi = w_hi & dm.limited_mask;
    GMR_STAT(0, 1);
    int nchecks = 0;
    bool check = false;
    const int max_as = 8 * dm.nh + 16;
    int it = 0;
    for (; it < max_as; it++) {
      const uint32_t pinned = pin_lo | pin_hi;
      if (!check && pinned) {
        GMR_LANES
          if (lane < dm.nh) s_bnd()[lane] = ((pin_hi >> lane) & 1u) ? bound_hi(lane) : bound_lo(lane);
        GMR_END
      }
      if (!check) convoy_arrive();
      build_rows();
      if (!check) {
        if (pinned) apply_pins(pinned);
        factor_solve(pinned);
        stat_refactor++;
        GMR_STAT(1, 1);
        // ratio test towards the candidate
        bool blocked = false;
        GMR_LANES
          R al = INF;
          if (lane < dm.nh) {
            const R x = lp(LP_X, lane);
            if (!((pinned >> lane) & 1u)) {
              // step fraction to the first bound in the direction of the candidate; the division only happens for a
              // lane that is actually blocked (al < 1  <=>  the remaining distance is shorter than the step)
              const R p = s_xs()[6 + lane] - x;
              const R bhi = bound_hi(lane), blo = bound_lo(lane);
              if (p > R(0) && bhi < INF) { if (bhi - x < p) al = (bhi - x) / p; }
              else if (p < R(0) && blo > -INF) { if (blo - x > p) al = (blo - x) / p; }
            } else {
              s_xs()[6 + lane] = x;                        // pinned: stays on its bound
            }
            s_red()[lane] = al;
          }
#ifdef GMR_EMULATE
          if (al < R(1)) blocked = true;
#else
          blocked = __any_sync(0xffffffffu, al < R(1));
#endif
        GMR_END
        if (blocked) {
          R alpha = R(1); int blk = -1;
#ifdef GMR_EMULATE
          for (int j = 0; j < dm.nh; j++) { const R a = s_red()[j]; if (a < alpha) { alpha = a; blk = j; } }
#ifdef GMR_STATS
          { int nb_ = 0; for (int j = 0; j < dm.nh; j++) if (s_red()[j] < R(1)) nb_++; GMR_STAT(9, nb_); if (nb_ >= 2) GMR_STAT(10, 1); }
#endif
#else
          {   // smallest step and the lowest hinge that attains it: butterfly min + ballot
            const R mine = lane_ < dm.nh ? s_red()[lane_] : R(INFINITY);
            R mn = mine;
            GMR_UNROLL
            for (int o = 16; o > 0; o >>= 1) { const R v = __shfl_xor_sync(0xffffffffu, mn, o); if (v < mn) mn = v; }
            alpha = mn;
            blk = __ffs(__ballot_sync(0xffffffffu, mine == mn)) - 1;
          }
#endif
          GMR_SYNC();
          if (alpha < R(0)) alpha = R(0);
          GMR_LANES
            if (lane < dm.nh) {
              const R x = lp(LP_X, lane), p = s_xs()[6 + lane] - x;
              if (lane == blk) {
                const bool up = p > R(0);
                lp(LP_X, lane) = up ? bound_hi(lane) : bound_lo(lane);
                s_piv()[2] = up ? R(1) : R(-1);
              } else if (!((pinned >> lane) & 1u)) {
                lp(LP_X, lane) = x + alpha * p;
              }
            }
          GMR_END
          const bool up = s_piv()[2] > R(0);
          GMR_SYNC();
          if (up) pin_hi |= 1u << blk; else pin_lo |= 1u << blk;
          GMR_STAT(2, 1);
          continue;
        }
        GMR_LANES
          if (lane < dm.nh) lp(LP_X, lane) = s_xs()[6 + lane];
        GMR_END
        if (!pinned) break;
        check = true;
        GMR_STAT(3, 1);
        continue;
      }
      // CHECK pass: g = H x + c on the original rows (just rebuilt); x is in s_xs.  Every lane publishes
      // (its row) * x_lane, a pinned lane adds its own slots times the ancestors' x and collects its
      // column from the published rows of its descendants.
      GMR_LANES
        if (lane < dm.nh) publish_row(Lrow(lane), L, lp(LP_X, lane), dm.o_y >> 2);
      GMR_END
      GMR_LANES
        {
          // straight-line (a lane that is not pinned computes a value it discards): rows hold zeros in the slots a hinge
          // does not have, and the table names the hinge itself there
          const bool mine = (pinned >> lane) & 1u;
          const int dep = GMR_DEPTH;
          R g = lp(LP_CI, lane) + L.dg * lp(LP_X, lane);
          GMR_UNROLL
          for (int f = 0; f < 6; f++) g += L.row[f] * s_xs()[f];
          GMR_UNROLL
          for (int d = 1; d < GMR_MAXD; d++) g += L.row[5 + d] * s_xs()[6 + mc.anc_of[lane * GMR_ANCS + d - 1]];
          g += sum_over_bits(mine ? GMR_DESC : 0u, [&](int k) { return Lrow(k)[5 + dep]; });
          if (lane < dm.nh) s_red()[lane] = mine ? (((pin_lo >> lane) & 1u) ? g : -g) : INF;
        }
      GMR_END
      // release: the first two checks of a solve drop EVERY bound with a multiplier of the wrong sign (a
      // stage switch typically flips several at once), later ones only the worst (the textbook rule, which
      // cannot cycle).  The loop ends on a point that passes this KKT check, or - never observed, but not excluded for
      // the multi-release rule in floating point - on its iteration cap, which the clip's status then reports.
      R lmin = R(0), gmax = R(1); int worst = -1;
      uint32_t drop_all = 0;
#ifdef GMR_EMULATE
      for (int j = 0; j < dm.nh; j++) if ((pinned >> j) & 1u) {
        const R l = s_red()[j];
        if (g_abs(l) > gmax) gmax = g_abs(l);
        if (worst < 0 || l < lmin) { lmin = l; worst = j; }
      }
      for (int j = 0; j < dm.nh; j++) if (((pinned >> j) & 1u) && s_red()[j] < -GmrEps<R>::lam * gmax) drop_all |= 1u << j;
#else
      {   // butterfly max |lambda| and min lambda over the pinned lanes, ballots for the argmin and the drop set
        const bool mine_pinned = (pinned >> lane_) & 1u;
        const R l = mine_pinned ? s_red()[lane_] : R(INFINITY);
        R mx = mine_pinned ? g_abs(l) : R(0), mn = l;
        GMR_UNROLL
        for (int o = 16; o > 0; o >>= 1) {
          const R a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
          if (a > mx) mx = a;
          if (b < mn) mn = b;
        }
        if (mx > gmax) gmax = mx;
        lmin = mn;
        worst = __ffs(__ballot_sync(0xffffffffu, mine_pinned && l == mn)) - 1;
        drop_all = __ballot_sync(0xffffffffu, mine_pinned && l < -GmrEps<R>::lam * gmax);
      }
#endif
      if (worst < 0 || lmin >= -GmrEps<R>::lam * gmax) { GMR_SYNC(); break; }
      uint32_t drop = 1u << worst;
      if (nchecks < 2) drop |= drop_all;
      GMR_STAT(4, 1); GMR_STAT(5, __builtin_popcount(drop));
      nchecks++;
      GMR_SYNC();
      pin_lo &= ~drop; pin_hi &= ~drop;
      check = false;
    }
    if (it >= max_as) stat_flags |= GMR_STATUS_AS_CAP;            // feasible point, optimality not proven
    GMR_STAT(6, __builtin_popcount(pin_lo | pin_hi)); GMR_STAT(7, (pin_lo | pin_hi) ? 1 : 0);
    GMR_STAT(8, (pin_lo == (w_lo & dm.limited_mask) && pin_hi == (w_hi & dm.limited_mask)) ? 1 : 0);
    if (stage == 0) { warm_lo0 = pin_lo; warm_hi0 = pin_hi; } else { warm_lo1 = pin_lo; warm_hi1 = pin_hi; }
    // publish the final step (hinge part) for integration
    GMR_LANES
      if (lane < dm.nh) s_xs()[6 + lane] = lp(LP_X, lane);
    GMR_END
  }

  // ------------------------------------------------------------------ integration (A12) ------
  // v = dq/dt, mj_integratePos(qpos, v, dt): world translation, body-local rotation
  // (right-multiplied), hinge angles.  The base rotation was solved world-aligned; it is
  // rotated into the root frame first.
  GMR_FN void integrate() {
    GMR_LANES
      // v = dq / dt as a multiplication by the precomputed 1 / dt (a float64 division is ~20 instructions; the
      // reference's round trip through dt only contributes rounding)
      const R dt = ks.dt, idt = ks.inv_dt;
      if (lane < dm.nh) {
        const R v = s_xs()[6 + lane] * idt;
        s_q()[7 + lane] += dt * v;
      }
      if (lane == 0) {
        GMR_UNROLL
        for (int i = 0; i < 3; i++) { const R v = s_xs()[i] * idt; s_q()[i] += dt * v; }
        R q[4]; g_ld4(s_rq(), q);                              // normalised root quaternion (s_xq is overwritten by the factor rows)
        R ww[3] = {s_xs()[3], s_xs()[4], s_xs()[5]}, wl[3];
        q_rot_inv(q, ww, wl);
        R v[3] = {wl[0] * idt, wl[1] * idt, wl[2] * idt};
        R nrm = g_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        R qr[4] = {R(1), R(0), R(0), R(0)};
        if (nrm >= R(1e-15)) {
          const R ang = dt * nrm;
          if (ang != R(0)) {
            R s, c; g_sincos(R(0.5) * ang, &s, &c);
            const R k = s / nrm;
            qr[0] = c; qr[1] = v[0] * k; qr[2] = v[1] * k; qr[3] = v[2] * k;
          }
        }
        R qq[4] = {s_q()[3], s_q()[4], s_q()[5], s_q()[6]};
        q_normalize(qq);
        R o[4]; q_mul(qq, qr, o);
        s_q()[3] = o[0]; s_q()[4] = o[1]; s_q()[5] = o[2]; s_q()[6] = o[3];
      }
    GMR_END
  }

  // ------------------------------------------------------------------ keypoint stream -------
  // One frame's raw keypoints (positions nhum x 12 bytes, quaternions nhum x 16 bytes, two arrays) are staged into
  // shared memory by TMA bulk copies (cp.async.bulk -> UBLKCP) whose completion is counted on the warp's own
  // mbarrier: one elected lane issues the quaternion block and the 16-byte aligned body of the position block;
  // frames are 12 nhum bytes apart, so a position block may start 4/8/12 bytes off a 16-byte line: its <= 3 head
  // and <= 3 tail words travel as 4-byte cp.async (LDGSTS) and the block is placed at the same offset inside its
  // slot (in_off).  Bulk requests are what the host link serves best when the arrays live in mapped pinned host
  // memory (gmr_retarget_batch_host reads them in place: tools/microbench/pcie.cu measured 33 GB/s against 22 GB/s
  // for per-body cp.async; the solve needs 6), and one instruction instead of four per body for arrays in HBM.
  // frame_wait() must be called before update_targets() reads the frame; one frame is in flight at most.
  GMR_FN void stream_init() {
#ifndef GMR_EMULATE
    if (lane_ == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_bar())));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif
    in_phase = 0; in_flight = false; in_off = 0;
  }
  GMR_FN void stage_frame(const float* pos, const float* quat) {
    const int pbytes = 12 * dm.nhum;
#ifdef GMR_EMULATE
    in_off = 0;
    const int head = 0, mid = 0;
#else
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(pos) & 15u);
    in_off = a >> 2;
    int head = (int)((16u - a) & 15u);                        // bytes in front of the first 16-byte line boundary
    if (head > pbytes) head = pbytes;
    const int mid = (pbytes - head) & ~15;                    // the aligned body
#endif
    GMR_LANES
#ifdef GMR_EMULATE
      (void)head; (void)mid;
      if (lane < dm.nhum) {
        float* dp = s_in() + in_off + 3 * lane; float* dq = s_inq() + 4 * lane;
        const float* p = pos + 3 * lane; const float* q = quat + 4 * lane;
        dp[0] = p[0]; dp[1] = p[1]; dp[2] = p[2];
        dq[0] = q[0]; dq[1] = q[1]; dq[2] = q[2]; dq[3] = q[3];
      }
#else
      const uint32_t dpos = (uint32_t)__cvta_generic_to_shared(s_in() + in_off);
      if (lane == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar());
        const uint32_t qbytes = 16u * (uint32_t)dm.nhum;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(qbytes + (uint32_t)mid) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(s_inq())), "l"(quat), "r"(qbytes), "r"(bar) : "memory");
        if (mid > 0)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dpos + (uint32_t)head), "l"(reinterpret_cast<const char*>(pos) + head), "r"((uint32_t)mid), "r"(bar) : "memory");
      } else if (lane >= 8 && lane < 16) {
        // head words (lanes 8..10) and tail words (lanes 12..14)
        const int w = lane - 8;
        const int nhead = head >> 2, ntail = (pbytes - head - mid) >> 2;
        int idx = -1;
        if (w < 4) { if (w < nhead) idx = w; }
        else if (w - 4 < ntail) idx = ((head + mid) >> 2) + (w - 4);
        if (idx >= 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dpos + 4u * (uint32_t)idx), "l"(pos + idx) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    GMR_END
    in_flight = true;
  }
  GMR_FN void frame_wait() {
    if (!in_flight) return;
    in_flight = false;
#ifndef GMR_EMULATE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar());
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(bar), "r"(in_phase) : "memory");
    }
    in_phase ^= 1u;
    __syncwarp();
#endif
  }

  // ------------------------------------------------------------------ convoy -----------------
  // All warps of a CTA run the same ~16k-instruction solve on different clips.  Left alone they
  // drift apart, every warp streams its own copy of the code through the instruction caches and
  // the SM starves (ncu: stall "no instruction" ~10 of ~20 warp-cycles per issue; the hot path
  // is ~75 KB against a 32 KB instruction cache).  One CTA-wide rendezvous per factorisation
  // keeps the warps inside the same stretch of code, so a fetched line serves all of them.
  // Everything between two rendezvous is the same amount of work for every warp (one
  // factorisation + the bookkeeping up to the next one); extra active-set passes simply take
  // extra rounds.  Warps that ran out of clips keep arriving until the CTA is done.
  GMR_FN void convoy_arrive() {
#ifndef GMR_EMULATE
    if (convoy) __syncthreads();
#endif
  }
  GMR_FN void convoy_retire() {
#ifndef GMR_EMULATE
    if (!convoy) return;
    if (lane_ == 0) atomicSub((int*)cta_active, 1);
    for (;;) {
      __syncthreads();
      if (*cta_active <= 0) break;
    }
#endif
  }

  // ------------------------------------------------------------------ one clip (A13) --------
  // pos/quat point at this clip's first frame ([T,nhum,3] / [T,nhum,4]); outputs at this clip's
  // first frame too.  Null output pointers are skipped.  The per-frame control flow of
  // retarget() (motion_retarget.py:139-185) is flattened into one loop so that every phase
  // exists once in the instruction stream:
  //   per stage:  curr = err(); solve; next = err(); n = 0
  //               while curr - next > tol and n < max_iter: curr = next; solve; next = err(); n += 1
  // `io` holds the batch's base pointers (a kernel parameter: constant bank, no registers); everything
  // clip- or frame-specific is derived from the clip index c and the frame index where it is used, so that a
  // warp carries two integers through its clip instead of a dozen pointers.
  // Post-solve epilogue of the dataset scripts (scripts/smplx_to_robot_dataset.py:93-123), fused: the last FK
  // of a frame is the FK of the frame's answer, so
  //   ex.local_body_pos [C,T,nb,3]  FK with an identity root = R_root^T (x_b - x_root), and
  //   ex.lowest_z [C]               min over frames and bodies of the world z (the height adjustment of :118-123)
  // cost one pass over the body poses already in shared memory.
  // seg_frames > 0 (scheduled launch): run at most that many frames starting at the frame recorded in the clip's state,
  // save the state again and return the next frame to run (== the clip's length when it is complete, -1 when the clip
  // was stopped by a fatal status); the task error of the last frame solved (the scheduler's slow / normal signal)
  // is left in s_piv()[1].  seg_frames == 0: the launch's own frame range [t_begin, t_end).
  template <typename IO>
  GMR_FN int run_clip(const GmrIO<IO>& io, int c, int seg_frames = 0) {
    const uint32_t flags = io.flags;
    const size_t f0 = (size_t)c * io.T;
    int T = io.T;
    if (io.ex.lengths) { T = io.ex.lengths[c]; T = T < 0 ? 0 : (T > io.T ? io.T : T); }     // ragged batch: this clip's own length
    const float* const pos = io.pos + f0 * dm.nhum * 3;
    const float* const quat = io.quat + f0 * dm.nhum * 4;
    int tb = io.t_begin;
    if (seg_frames > 0) {
      tb = (int)g_ldcg(io.state + (size_t)c * gmr_state_stride(dm.nq) + dm.nq + 4);
      if (tb + seg_frames < T) T = tb + seg_frames;
    } else if (io.t_end > 0 && io.t_end < T) T = io.t_end;
    if (tb >= T) return tb;                                           // nothing of this clip in this launch's frame range
    const bool resume = tb > 0 && io.state;
    stat_flags = 0;
    stat_refactor = 0;
#ifndef GMR_EMULATE
    if (io.trace && lane_ == 0) {                                     // nothing of the trace stays live across the clip
      long long* o = io.trace + 4 * ((size_t)(io.t_begin > 0 ? io.C : 0) + c);
      long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if (seg_frames == 0 || g_ldcg(o) == 0) { o[0] = t0; o[2] = 0; o[3] = 0; }
      o[1] = t0;                                                      // start of this segment (overwritten by its end below)
    }
#endif
    int stat_frame = -1;                                              // frame of the first status event
    {
    const double* const st = io.state ? io.state + (size_t)c * gmr_state_stride(dm.nq) : nullptr;   // not kept live: re-derived at the end
    if (resume) { const double e = g_ldcg(st + dm.nq + 1); if (!(e == e)) return -1; }   // the clip was stopped earlier (error = NaN)
    if (resume) {
      GMR_LANES
        for (int i = lane; i < dm.nq; i += 32) s_q()[i] = R(g_ldcg(st + i));
      GMR_END
    }
    else if (io.qinit) set_qpos(io.qinit + (size_t)c * dm.nq);
    else set_qpos(mc.qpos0);
    warm_lo0 = warm_lo1 = warm_hi0 = warm_hi1 = 0;
    {
      const uint32_t* w = resume ? reinterpret_cast<const uint32_t*>(st + dm.nq + 2) : (io.ex.warm_state ? io.ex.warm_state + 4 * c : nullptr);
      if (w) { warm_lo0 = g_ldcg(w); warm_hi0 = g_ldcg(w + 1); warm_lo1 = g_ldcg(w + 2); warm_hi1 = g_ldcg(w + 3); }
    }
    GMR_LANES
      if (lane == 0) { s_piv()[0] = io.ratio ? R(io.ratio[c]) : R(1); s_piv()[3] = resume ? R(g_ldcg(st + dm.nq)) : R(INFINITY); }   // height ratio, running lowest z
    GMR_END
    }
    stage_frame(pos + (size_t)tb * dm.nhum * 3, quat + (size_t)tb * dm.nhum * 4);
    const int first_stage = dm.use1 ? 0 : 1;
    const bool any_stage = dm.use1 || dm.use2;
    int t = tb, stage = first_stage, nsolve = 0;
    int n0 = 0, n1 = 0;
    R e0 = R(0), e1 = R(0);
    R curr = R(0);
    bool frame_start = true, need_fk = true, need_err = true;
    for (;;) {
      if (frame_start) {
        frame_wait();
        if (!update_targets(s_piv()[0], (flags & GMR_FLAG_OFFSET_TO_GROUND) != 0)) {
          stat_flags |= GMR_STATUS_BAD_INPUT; stat_frame = t;
          break;                                                       // nothing in flight: the next frame is not staged yet
        }
        if (t + 1 < T) stage_frame(pos + (size_t)(t + 1) * dm.nhum * 3, quat + (size_t)(t + 1) * dm.nhum * 4);
        stage = first_stage; nsolve = 0; n0 = n1 = 0; e0 = e1 = R(0);
        frame_start = false; need_err = true;
      }
      if (need_fk) { fk(); need_fk = false; need_err = true; }
      bool solve = false, done = !any_stage;
      if (any_stage) {
        if (need_err) { task_error(); need_err = false; }     // a stage switch re-sums the same |e_t|^2 with the other mask
        const R e = stage_error(stage);
        if (flags & GMR_FLAG_NO_SOLVE) {                 // targets + errors only (update_targets / error1 / error2)
          if (stage == 0) e0 = e; else e1 = e;
          if (stage == 0 && dm.use2) { stage = 1; continue; }
          done = true;
        }
        else if (nsolve == 0) { curr = e; solve = true; }
        else if (curr - e > ks.tol && nsolve - 1 < dm.max_iter) { curr = e; solve = true; }
        else {
          if (stage == 0) { n0 = nsolve; e0 = e; } else { n1 = nsolve; e1 = e; }
          if (stage == 0 && dm.use2) { stage = 1; nsolve = 0; continue; }   // re-evaluate with stage-2 weights
          done = true;
        }
      }
      if (solve) {
        task_build(stage);
        composites();
        solve_qp(stage);
        integrate();
        need_fk = true;
        nsolve++;
        continue;
      }
      if (done) {
        const size_t f = (size_t)c * io.T + t;
        bool finite = true;
        GMR_LANES
          IO* qo = io.qout + f * dm.nq;
          bool fin = true;
          for (int i = lane; i < dm.nq; i += 32) { const R v = s_q()[i]; qo[i] = IO(v); fin = fin && (g_abs(v) < R(INFINITY)); }
#ifdef GMR_EMULATE
          if (!fin) finite = false;
#else
          finite = __all_sync(0xffffffffu, fin);
#endif
          if (lane == 0) {
            if (io.iters) { io.iters[2 * f] = n0; io.iters[2 * f + 1] = n1; }
            if (io.err) { io.err[2 * f] = IO(e0); io.err[2 * f + 1] = IO(e1); }
          }
          if (io.tg && lane < dm.nhum) {
            IO* o = io.tg + (f * dm.nhum + lane) * 7;
            const R* g = s_tg() + TG * lane;
            o[0] = IO(g[0]); o[1] = IO(g[1]); o[2] = IO(g[2]); o[3] = IO(g[4]); o[4] = IO(g[5]); o[5] = IO(g[6]); o[6] = IO(g[7]);
          }
        GMR_END
        if (io.ex.local_body_pos || io.ex.lowest_z) {
          GMR_LANES
            R zmin = R(INFINITY);
            R rq[4]; g_ld4(s_rq(), rq);
            for (int b = lane; b < dm.nb; b += 32) {
              R v[4]; g_ld4(s_xp() + PX * b, v);
              if (v[2] < zmin) zmin = v[2];
              if (io.ex.local_body_pos) {
                R o[3]; q_rot_inv(rq, v, o);
                float* dst = io.ex.local_body_pos + (f * dm.nb + b) * 3;
                dst[0] = (float)o[0]; dst[1] = (float)o[1]; dst[2] = (float)o[2];
              }
            }
            s_red()[lane] = zmin;
          GMR_END
#ifdef GMR_EMULATE
          R zm = s_red()[0];
          for (int j = 1; j < 32; j++) { const R z = s_red()[j]; if (z < zm) zm = z; }
#else
          R zm = s_red()[lane_];
          GMR_UNROLL
          for (int o = 16; o > 0; o >>= 1) { const R z = __shfl_xor_sync(0xffffffffu, zm, o); if (z < zm) zm = z; }
#endif
          zm += s_q()[2];
          const R low = s_piv()[3];
          GMR_SYNC();
          GMR_LANES
            if (lane == 0 && zm < low) s_piv()[3] = zm;
          GMR_END
        }
        if (stat_flags && stat_frame < 0) stat_frame = t;
        if (!finite) {                                                 // the reference's `assert dq is not None`
          stat_flags |= GMR_STATUS_NONFINITE;
          frame_wait();                                                // drain the prefetch of frame t + 1
          break;
        }
        t++;
        if (t == T) break;
        frame_start = true;
      }
    }
#ifndef GMR_EMULATE
    if (io.trace && lane_ == 0) {
      long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      const unsigned smid = blockIdx.x;                                // the CTA (one per SM): tells sparse from dense SMs
      long long* o = io.trace + 4 * ((size_t)(io.t_begin > 0 ? io.C : 0) + c);
      // segments of one clip accumulate into one row: running time, factorisations | segments << 32 | last SM << 48
      o[2] = g_ldcg(o + 2) + (t1 - g_ldcg(o + 1));
      o[1] = t1;
      o[3] = ((g_ldcg(o + 3) & 0xffffffffffffll) + (long long)stat_refactor + (1ll << 32)) | ((long long)smid << 48);
    }
#endif
    const bool stopped = (stat_flags & GMR_STATUS_FATAL) != 0;
    if (io.ex.status && stat_flags) {
      GMR_LANES
        if (lane == 0) {
          const int32_t old = g_ldcg(io.ex.status + c);
          io.ex.status[c] = old ? (old | (int32_t)stat_flags) : (int32_t)(stat_flags | ((uint32_t)stat_frame << 8));
        }
      GMR_END
    }
    if (io.state) {                                                   // hand-over to the launch that continues this clip
      double* const st = io.state + (size_t)c * gmr_state_stride(dm.nq);
      GMR_LANES
        for (int i = lane; i < dm.nq; i += 32) st[i] = (double)s_q()[i];
        if (lane == 0) {
          st[dm.nq] = (double)s_piv()[3];
          st[dm.nq + 1] = stopped ? (double)NAN : (double)(dm.use2 ? e1 : e0);
          uint32_t* w = reinterpret_cast<uint32_t*>(st + dm.nq + 2);
          w[0] = warm_lo0; w[1] = warm_hi0; w[2] = warm_lo1; w[3] = warm_hi1;
          st[dm.nq + 4] = (double)t;
        }
      GMR_END
    }
    if (io.ex.lowest_z || io.ex.warm_state) {
      GMR_LANES
        if (lane == 0) {
          if (io.ex.lowest_z) io.ex.lowest_z[c] = (float)s_piv()[3];
          if (io.ex.warm_state) { uint32_t* w = io.ex.warm_state + 4 * c; w[0] = warm_lo0; w[1] = warm_hi0; w[2] = warm_lo1; w[3] = warm_hi1; }
        }
      GMR_END
    }
    if (seg_frames > 0) {
      GMR_LANES
        if (lane == 0) s_piv()[1] = dm.use2 ? e1 : e0;
      GMR_END
    }
    return stopped ? -1 : t;                                          // next frame to run (>= the clip's length: complete)
  }
};
