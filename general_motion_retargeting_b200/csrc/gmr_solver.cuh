// Warp-per-clip fused retargeting IK solver (the body of the sm_100a kernel).
//
// One warp owns one clip: frames are solved in order, warm-started, exactly like
// `for frame in frames: retargeter.retarget(frame)` in the reference
// (general_motion_retargeting/motion_retarget.py:139-185; callers
// scripts/smplx_to_robot_dataset.py:84-87).  This is NOT a port of the mink/MuJoCo/DAQP
// call sequence; the per-solve algebra is re-derived for a 32-lane warp:
//
//  * FK by tree level (lanes = bodies of one depth), positions kept relative to the
//    floating root so float32 keeps ~1e-7 m resolution anywhere in the world.
//  * The QP matrix H = damping*I + sum_t [(W J_t)^T (W J_t) + mu_t I] is never formed from
//    dense 6 x nv task Jacobians.  Each frame task is a 6x6 "spring inertia"
//    M_t = A_t^T A_t at the reference point (A_t = -W * Jlog_t * blkdiag(R_b^T) shifted to the
//    root), every hinge lane sums the M_t of the tasks in its subtree (composite, as in the
//    composite-rigid-body algorithm) and gets its row as H_ij = s_i^T Ic_i s_j over its
//    ancestors j, with s_j the world-frame spatial axis of DoF j.  ~10x fewer flops than the
//    dense J^T W^2 J and no shared-memory read-modify-write.
//  * Row i of H lives in REGISTERS of lane i (static indexing, loops unrolled over the pivot
//    index); the Cholesky factorisation broadcasts the pivot row from shared memory
//    (one wavefront per 4 values) instead of shuffling or re-reading lane-private rows.
//    The 6 floating-base DoFs are factored redundantly by every lane (6x6), so no lane ever
//    needs a second pass when nv = 6 + nhinge > 32.
//  * Joint-limit box: exact primal active set (same optimum as the reference's DAQP solve);
//    the common case (unconstrained step feasible) costs one factorisation.
//
// The file is written as a sequence of "lane blocks" (GMR_LANES ... GMR_END).  On the GPU a
// block is straight-line code of one thread followed by __syncwarp(); defining GMR_EMULATE
// turns each block into a host loop over 32 lanes, which is how tests/emu debugs the
// warp-level logic on a machine without a GPU.  Rule: inside one block a lane never reads
// shared memory another lane writes in the same block.
#pragma once
#include <math.h>
#include <stdint.h>

#include "gmr_consts.h"

#ifdef GMR_EMULATE
#include <cmath>
#include <cstring>
#define GMR_FN inline
#define GMR_LANES for (int lane = 0; lane < 32; ++lane) { LaneRegs<R>& L = lanes_[lane]; (void)L;
#define GMR_END }
#define GMR_UNROLL
#else
#define GMR_FN __device__ __forceinline__
#define GMR_LANES { LaneRegs<R>& L = lanes_; const int lane = lane_; (void)L; (void)lane;
#define GMR_END } __syncwarp();
#define GMR_UNROLL _Pragma("unroll")
#endif

#define GMR_HD  // functions below are templates/inline; host+device qualifiers added per build
#if !defined(GMR_EMULATE)
#undef GMR_HD
#define GMR_HD __host__ __device__
#endif

// ---- shared-memory layout of one warp's state (units: elements of R) -----------------------
GMR_HD constexpr int gmr_pad4(int n) { return (n + 3) & ~3; }
// offset of pivot row k in the packed factor storage: row k holds 6 base columns + k hinge columns
GMR_HD constexpr int gmr_loff(int k) { int o = 0; for (int i = 0; i < k; i++) o += gmr_pad4(6 + i); return o; }

struct GmrWarpLayout {
  int q, xp, xq, sc, tg, sd, red, xs, root, lf, piv, bnd, tk, mt, lfac, total;
};
GMR_HD inline GmrWarpLayout gmr_warp_layout(int nb, int nh, int nhum, int nt) {
  GmrWarpLayout w{};
  int o = 0;
  w.q = o;    o += gmr_pad4(7 + nh);
  w.xp = o;   o += gmr_pad4(3 * nb);
  w.xq = o;   o += 4 * nb;
  w.sc = o;   o += gmr_pad4(2 * nh);
  w.tg = o;   o += 8 * nhum;
  w.sd = o;   o += 8 * (nh > 0 ? nh : 1);
  w.red = o;  o += 32;
  w.xs = o;   o += gmr_pad4(6 + nh);
  w.root = o; o += 28;
  w.lf = o;   o += 28;
  w.piv = o;  o += 4;
  w.bnd = o;  o += 32;
  // union: {task kinematics (24/task) + task inertias (28/task)}  vs  packed factor rows
  w.tk = o; w.mt = o + 24 * nt; w.lfac = o;
  int a = 52 * nt, b = gmr_loff(nh);
  o += gmr_pad4(a > b ? a : b);
  w.total = o;
  return w;
}

template <typename R> struct GmrEps;
template <> struct GmrEps<float>  { static constexpr float  lie = 1.1920929e-06f; static constexpr float  lam = 1e-5f;  };
template <> struct GmrEps<double> { static constexpr double lie = 2.220446049250313e-15; static constexpr double lam = 1e-12; };

// per-lane registers that persist across lane blocks
template <typename R> struct LaneRegs {
  R a[6];          // base (floating joint) columns of this hinge's row of H, then of L
  R row[GMR_NH];   // hinge columns j <= lane
  R rhs, dinv, tmp;
  R f[6];          // Ic_i * s_i
  R diag, ci;
  R x, xs, blo, bhi;
  float in_pos[3]; float in_quat[4];
};

// ---- tiny math helpers ----------------------------------------------------------------------
template <typename R> GMR_FN R g_sqrt(R x) { return sqrt(x); }
template <typename R> GMR_FN R g_atan2(R y, R x) { return atan2(y, x); }
template <typename R> GMR_FN R g_abs(R x) { return fabs(x); }
template <typename R> GMR_FN void g_sincos(R x, R* s, R* c) {
#ifdef GMR_EMULATE
  *s = std::sin(x); *c = std::cos(x);
#else
  sincos(x, s, c);
#endif
}
#ifndef GMR_EMULATE
template <> __device__ __forceinline__ void g_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ float g_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ float g_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> __device__ __forceinline__ float g_abs<float>(float x) { return fabsf(x); }
#endif

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
  R n = g_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < R(1e-15)) { q[0] = R(1); q[1] = q[2] = q[3] = R(0); return; }
  R inv = R(1) / n;
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

// index of (i,j), i<=j, in the packed upper triangle of a symmetric 6x6 (21 entries)
GMR_HD constexpr int gmr_sym6(int i, int j) { return i <= j ? (i * (13 - i)) / 2 + (j - i) : (j * (13 - j)) / 2 + (i - j); }

// =============================================================================================
template <typename R>
struct WarpSolver {
  const GmrConsts<R>& mc;
  R* sm;                      // this warp's shared-memory block
  GmrWarpLayout lay;
#ifdef GMR_EMULATE
  LaneRegs<R> lanes_[32];
#else
  LaneRegs<R> lanes_;
  int lane_;
#endif
  // per-solve statistics (uniform)
  int stat_refactor;
  uint32_t warm_lo, warm_hi;   // working set carried from the previous solve (uniform)

  GMR_FN WarpSolver(const GmrConsts<R>& m, R* smem
#ifndef GMR_EMULATE
                    , int lane
#endif
                    ) : mc(m), sm(smem), lay(gmr_warp_layout(m.nb, m.nh, m.nhum, m.nt)), stat_refactor(0), warm_lo(0), warm_hi(0) {
#ifndef GMR_EMULATE
    lane_ = lane;
#endif
  }

  GMR_FN R* s_q() const { return sm + lay.q; }
  GMR_FN R* s_xp() const { return sm + lay.xp; }
  GMR_FN R* s_xq() const { return sm + lay.xq; }
  GMR_FN R* s_sc() const { return sm + lay.sc; }
  GMR_FN R* s_tg() const { return sm + lay.tg; }
  GMR_FN R* s_sd() const { return sm + lay.sd; }
  GMR_FN R* s_red() const { return sm + lay.red; }
  GMR_FN R* s_xs() const { return sm + lay.xs; }
  GMR_FN R* s_root() const { return sm + lay.root; }
  GMR_FN R* s_lf() const { return sm + lay.lf; }
  GMR_FN R* s_piv() const { return sm + lay.piv; }
  GMR_FN R* s_bnd() const { return sm + lay.bnd; }
  GMR_FN R* s_tk() const { return sm + lay.tk; }
  GMR_FN R* s_mt() const { return sm + lay.mt; }
  GMR_FN R* s_L() const { return sm + lay.lfac; }

  // ------------------------------------------------------------------ configuration --------
  // qpos -> shared memory.  `src` has nq values (uniform pointer); all lanes cooperate.
  template <typename S> GMR_FN void set_qpos(const S* src) {
    GMR_LANES
      for (int i = lane; i < mc.nq; i += 32) s_q()[i] = R(src[i]);
    GMR_END
  }

  // forward kinematics of s_q -> s_xp (root-relative positions), s_xq (world orientations),
  // then the world-frame spatial axis of every hinge at the root origin -> s_sd[j] = (v, w).
  GMR_FN void fk() {
    GMR_LANES
      if (lane < mc.nh) {
        R s, c; g_sincos(R(0.5) * s_q()[7 + lane], &s, &c);
        s_sc()[2 * lane] = s; s_sc()[2 * lane + 1] = c;
      }
      if (lane == 0) {
        R q[4] = {s_q()[3], s_q()[4], s_q()[5], s_q()[6]};
        q_normalize(q);
        s_xq()[0] = q[0]; s_xq()[1] = q[1]; s_xq()[2] = q[2]; s_xq()[3] = q[3];
        s_xp()[0] = R(0); s_xp()[1] = R(0); s_xp()[2] = R(0);
      }
    GMR_END
    for (int l = 1; l < mc.nlevel; l++) {
      const int beg = mc.lvl_off[l], cnt = mc.lvl_off[l + 1] - beg;
      GMR_LANES
        for (int e = lane; e < cnt; e += 32) {
          const int b = mc.lvl_body[beg + e], p = mc.parent[b];
          const R* qp = s_xq() + 4 * p;
          R pq[4] = {qp[0], qp[1], qp[2], qp[3]};
          R off[3]; q_rot(pq, mc.bpos + 3 * b, off);
          s_xp()[3 * b] = s_xp()[3 * p] + off[0]; s_xp()[3 * b + 1] = s_xp()[3 * p + 1] + off[1]; s_xp()[3 * b + 2] = s_xp()[3 * p + 2] + off[2];
          R q[4]; q_mul(pq, mc.bquat + 4 * b, q);
          const int j = mc.bhinge[b];
          if (j >= 0) {
            const R s = s_sc()[2 * j], c = s_sc()[2 * j + 1];
            R ql[4] = {c, mc.axis[3 * j] * s, mc.axis[3 * j + 1] * s, mc.axis[3 * j + 2] * s};
            R t[4]; q_mul(q, ql, t); q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; q[3] = t[3];
          }
          // first-order renormalisation (|q| is 1 up to rounding): q *= 1.5 - 0.5 |q|^2
          const R k = R(1.5) - R(0.5) * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
          s_xq()[4 * b] = q[0] * k; s_xq()[4 * b + 1] = q[1] * k; s_xq()[4 * b + 2] = q[2] * k; s_xq()[4 * b + 3] = q[3] * k;
        }
      GMR_END
    }
    GMR_LANES
      if (lane < mc.nh) {
        const int b = mc.hbody[lane];
        const R* qb = s_xq() + 4 * b;
        R q[4] = {qb[0], qb[1], qb[2], qb[3]};
        R w[3]; q_rot(q, mc.axis + 3 * lane, w);
        const R* d = s_xp() + 3 * b;
        R* o = s_sd() + 8 * lane;
        // linear velocity of the reference point (root origin) under unit joint rate: w x (0 - d) = d x w
        o[0] = d[1] * w[2] - d[2] * w[1]; o[1] = d[2] * w[0] - d[0] * w[2]; o[2] = d[0] * w[1] - d[1] * w[0];
        o[3] = w[0]; o[4] = w[1]; o[5] = w[2]; o[6] = R(0); o[7] = R(0);
      }
    GMR_END
  }

  // ------------------------------------------------------------------ targets (A1-A5) ------
  // Raw keypoints of one frame (held per lane in L.in_pos/in_quat for lane < nhum) ->
  // scaled + offset targets in s_tg[h] = (pos[3], pad, quat[4]).
  GMR_FN void update_targets(R ratio, bool to_ground) {
    GMR_LANES
      if (lane == mc.hroot) { s_red()[0] = R(L.in_pos[0]); s_red()[1] = R(L.in_pos[1]); s_red()[2] = R(L.in_pos[2]); }
    GMR_END
    GMR_LANES
      if (lane < mc.nhum) {
        const R rx = s_red()[0], ry = s_red()[1], rz = s_red()[2];
        const R sr = mc.hscale[mc.hroot] * ratio;
        R p[3];
        if (lane == mc.hroot) { p[0] = sr * rx; p[1] = sr * ry; p[2] = sr * rz; }
        else {
          const R s = mc.hscale[lane] * ratio;
          p[0] = (R(L.in_pos[0]) - rx) * s + sr * rx; p[1] = (R(L.in_pos[1]) - ry) * s + sr * ry; p[2] = (R(L.in_pos[2]) - rz) * s + sr * rz;
        }
        R q[4] = {R(L.in_quat[0]), R(L.in_quat[1]), R(L.in_quat[2]), R(L.in_quat[3])};
        R n = R(1) / g_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
        R u[4]; q_mul(q, mc.hroff + 4 * lane, u);
        n = R(1) / g_sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
        u[0] *= n; u[1] *= n; u[2] *= n; u[3] *= n;
        R g[3]; q_rot(u, mc.hpoff + 3 * lane, g);
        R* o = s_tg() + 8 * lane;
        o[0] = p[0] + g[0]; o[1] = p[1] + g[1]; o[2] = p[2] + g[2]; o[3] = R(0);
        o[4] = u[0]; o[5] = u[1]; o[6] = u[2]; o[7] = u[3];
      }
    GMR_END
    if (to_ground) {      // offset_human_data_to_ground, motion_retarget.py:252-270
      R lowest = R(INFINITY);
      for (int h = 0; h < mc.nhum; h++) if ((mc.foot_mask >> h) & 1u) { R z = s_tg()[8 * h + 2]; if (z < lowest) lowest = z; }
#ifndef GMR_EMULATE
      __syncwarp();
#endif
      GMR_LANES
        if (lane < mc.nhum) s_tg()[8 * lane + 2] = s_tg()[8 * lane + 2] - lowest + R(0.1);
      GMR_END
    }
  }

  // ------------------------------------------------------------------ task kinematics (A7, A9)
  // Per task: error e = log(T_b^-1 T_t) = (rho, omega); P = Jinv(omega) R_b^T;
  // K' = -Jinv Q Jinv R_b^T - P [d_b]x  (the position rows' dependence on rotation, shifted to the
  // root origin).  s_tk[t] = rho(3) omega(3) P(9) K'(9); s_red[t] = |e|^2.
  GMR_FN void task_kinematics() {
    GMR_LANES
      if (lane < mc.nt) {
        const int b = mc.tbody[lane], h = mc.thuman[lane];
        const R* sq = s_xq() + 4 * b; const R* sd = s_xp() + 3 * b; const R* tg = s_tg() + 8 * h;
        R qb[4] = {sq[0], sq[1], sq[2], sq[3]};
        R d[3] = {sd[0], sd[1], sd[2]};
        R qi[4] = {qb[0], -qb[1], -qb[2], -qb[3]};
        R qt[4] = {tg[4], tg[5], tg[6], tg[7]};
        R qe[4]; q_mul(qi, qt, qe);
        if (qe[0] < R(0)) { qe[0] = -qe[0]; qe[1] = -qe[1]; qe[2] = -qe[2]; qe[3] = -qe[3]; }
        // world offset target - body, via root-relative coordinates
        R dw[3] = {(tg[0] - s_q()[0]) - d[0], (tg[1] - s_q()[1]) - d[1], (tg[2] - s_q()[2]) - d[2]};
        R tb[3]; q_rot_inv(qb, dw, tb);
        // SO(3) log
        const R nsq = qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3];
        R fac, cV, th2;
        R om[3];
        const bool small = nsq < GmrEps<R>::lie;
        if (small) {
          fac = R(2) / qe[0] - R(2) / R(3) * nsq / (qe[0] * qe[0] * qe[0]);
        } else {
          const R n = g_sqrt(nsq);
          fac = R(2) * g_atan2(n, qe[0]) / n;
        }
        om[0] = fac * qe[1]; om[1] = fac * qe[2]; om[2] = fac * qe[3];
        th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
        // Jinv = I - S/2 + cV S^2 with cV = (1 - (theta/2) cot(theta/2)) / theta^2 (also V^-1 of SE3.log)
        R Bq, Cq, Dq;   // Barfoot's Q coefficients
        if (th2 < GmrEps<R>::lie) {
          cV = R(1) / R(12);
          Bq = R(1) / R(6); Cq = -R(1) / R(24); Dq = R(1) / R(120);
        } else {
          const R th = g_sqrt(th2);
          const R n = g_sqrt(nsq);                       // sin(theta/2); qe[0] = cos(theta/2)
          cV = (R(1) - R(0.5) * th * qe[0] / n) / th2;
          if (th < R(0.25)) {                            // series: closed forms cancel badly for small theta
            Bq = R(1) / R(6) - th2 * (R(1) / R(120) - th2 * (R(1) / R(5040) - th2 / R(362880)));
            Cq = -R(1) / R(24) + th2 * (R(1) / R(720) - th2 * (R(1) / R(40320) - th2 / R(3628800)));
            Dq = R(1) / R(120) - th2 * (R(1) / R(2520) - th2 * (R(1) / R(120960) - th2 / R(9979200)));
          } else {
            const R st = R(2) * n * qe[0], ct = qe[0] * qe[0] - nsq;
            Bq = (th - st) / (th2 * th);
            Cq = (R(1) - th2 * R(0.5) - ct) / (th2 * th2);
            Dq = (R(2) * th - R(3) * st + th * ct) / (R(2) * th2 * th2 * th);
          }
        }
        R S[9], S2[9], Ji[9];
        m3_skew(om, S); m3_mul(S, S, S2);
        GMR_UNROLL
        for (int i = 0; i < 9; i++) Ji[i] = ((i & 3) == 0 ? R(1) : R(0)) - R(0.5) * S[i] + cV * S2[i];
        R rho[3] = {Ji[0] * tb[0] + Ji[1] * tb[1] + Ji[2] * tb[2], Ji[3] * tb[0] + Ji[4] * tb[1] + Ji[5] * tb[2], Ji[6] * tb[0] + Ji[7] * tb[1] + Ji[8] * tb[2]};
        // Q(rho, omega) with W V W = -(omega.rho) W,  W V W W + W W V W = -2 (omega.rho) W^2
        R V[9], VW[9], VWW[9], Q[9];
        m3_skew(rho, V); m3_mul(V, S, VW); m3_mul(VW, S, VWW);
        const R wr = om[0] * rho[0] + om[1] * rho[1] + om[2] * rho[2];
        GMR_UNROLL
        for (int i = 0; i < 3; i++) {
          GMR_UNROLL
          for (int j = 0; j < 3; j++) {
            const int ij = 3 * i + j, ji = 3 * j + i;
            const R wvw = -wr * S[ij];
            Q[ij] = R(0.5) * V[ij] + Bq * (VW[ji] + VW[ij] + wvw) - Cq * (VWW[ij] - VWW[ji] - R(3) * wvw) + Dq * (-R(2) * wr * S2[ij]);
          }
        }
        R Rb[9]; q_to_mat(qb, Rb);
        R P[9], T1[9], T2[9], K[9];
        m3_mul_bt(Ji, Rb, P);                 // P = Jinv R_b^T
        m3_mul(Ji, Q, T1); m3_mul(T1, P, T2); // Jinv Q Jinv R_b^T
        R Sd[9], PS[9]; m3_skew(d, Sd); m3_mul(P, Sd, PS);
        GMR_UNROLL
        for (int i = 0; i < 9; i++) K[i] = -T2[i] - PS[i];
        R* o = s_tk() + 24 * lane;
        o[0] = rho[0]; o[1] = rho[1]; o[2] = rho[2]; o[3] = om[0]; o[4] = om[1]; o[5] = om[2];
        GMR_UNROLL
        for (int i = 0; i < 9; i++) { o[6 + i] = P[i]; o[15 + i] = K[i]; }
        s_red()[lane] = rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2] + th2;
      }
    GMR_END
  }

  // unweighted error norm of the stage's tasks (error1()/error2(), motion_retarget.py:188-200)
  GMR_FN R stage_error(uint32_t stage_mask) const {
    R s = R(0);
    for (int t = 0; t < mc.nt; t++) if ((stage_mask >> t) & 1u) s += s_red()[t];
    return g_sqrt(s);
  }

  // ------------------------------------------------------------------ task inertias (A9, A10)
  // s_mt[t] = M_t packed upper (21) | g_t (6) | mu_t, for the stage's weights.
  GMR_FN void task_inertias(const R* wtab, uint32_t stage_mask) {
    GMR_LANES
      if (lane < mc.nt) {
        R* o = s_mt() + 28 * lane;
        const bool on = (stage_mask >> lane) & 1u;
        const R wp = on ? wtab[2 * lane] : R(0), wr = on ? wtab[2 * lane + 1] : R(0);
        const R wp2 = wp * wp, wr2 = wr * wr;
        const R* k = s_tk() + 24 * lane;
        R rho[3] = {k[0], k[1], k[2]}, om[3] = {k[3], k[4], k[5]};
        R P[9], K[9];
        GMR_UNROLL
        for (int i = 0; i < 9; i++) { P[i] = k[6 + i]; K[i] = k[15 + i]; }
        R PtP[9], PtK[9], KtK[9];
        m3_mul_at(P, P, PtP); m3_mul_at(P, K, PtK); m3_mul_at(K, K, KtK);
        // M = [[wp2 PtP, wp2 PtK],[., wp2 KtK + wr2 PtP]]
        GMR_UNROLL
        for (int i = 0; i < 3; i++) {
          GMR_UNROLL
          for (int j = 0; j < 3; j++) {
            if (j >= i) { o[gmr_sym6(i, j)] = wp2 * PtP[3 * i + j]; o[gmr_sym6(3 + i, 3 + j)] = wp2 * KtK[3 * i + j] + wr2 * PtP[3 * i + j]; }
            o[gmr_sym6(i, 3 + j)] = wp2 * PtK[3 * i + j];
          }
        }
        // g = A'^T W e = -[wp2 P^T rho ; wp2 K^T rho + wr2 P^T om]
        GMR_UNROLL
        for (int i = 0; i < 3; i++) {
          const R ptr = P[i] * rho[0] + P[3 + i] * rho[1] + P[6 + i] * rho[2];
          const R ktr = K[i] * rho[0] + K[3 + i] * rho[1] + K[6 + i] * rho[2];
          const R pto = P[i] * om[0] + P[3 + i] * om[1] + P[6 + i] * om[2];
          o[21 + i] = -(wp2 * ptr);
          o[24 + i] = -(wp2 * ktr + wr2 * pto);
        }
        o[27] = mc.lm * (wp2 * (rho[0] * rho[0] + rho[1] * rho[1] + rho[2] * rho[2]) + wr2 * (om[0] * om[0] + om[1] * om[1] + om[2] * om[2]));
      }
    GMR_END
    // whole-tree composite for the floating base: s_root[0..20] = sum M, [21..26] = sum g, [27] = sum mu
    GMR_LANES
      if (lane < 28) {
        R s = R(0);
        for (int t = 0; t < mc.nt; t++) s += s_mt()[28 * t + lane];
        s_root()[lane] = s;
      }
    GMR_END
    // per-hinge composite -> f_i = Ic_i s_i, c_i = s_i . G_i, diag
    GMR_LANES
      if (lane < mc.nh) {
        R acc[27];
        GMR_UNROLL
        for (int i = 0; i < 27; i++) acc[i] = R(0);
        const uint32_t tm = mc.task_mask[lane];
        for (int t = 0; t < mc.nt; t++) {
          if ((tm >> t) & 1u) {
            const R* m = s_mt() + 28 * t;
            GMR_UNROLL
            for (int i = 0; i < 27; i++) acc[i] += m[i];
          }
        }
        const R* s = s_sd() + 8 * lane;
        R sv[6] = {s[0], s[1], s[2], s[3], s[4], s[5]};
        R dg = R(0), ci = R(0);
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          R v = R(0);
          GMR_UNROLL
          for (int j = 0; j < 6; j++) v += acc[gmr_sym6(i, j)] * sv[j];
          L.f[i] = v; dg += v * sv[i]; ci += acc[21 + i] * sv[i];
        }
        L.diag = dg + mc.damping + s_root()[27];
        L.ci = ci;
      }
    GMR_END
  }

  // ------------------------------------------------------------------ rows of H -------------
  // lane i: row[j] = f_i . s_j for hinge ancestors j, row[i] = diag, a[] = f_i (base columns are
  // unit twists at the root origin, world-aligned), rhs = -c_i.  With pins: pinned lanes become
  // identity rows with rhs = bound, free lanes move the pinned columns to the right-hand side.
  GMR_FN void build_rows(uint32_t pinned) {
    GMR_LANES
      if (lane < mc.nh) {
        const uint32_t am = mc.anc_mask[lane];
        GMR_UNROLL
        for (int j = 0; j < GMR_NH; j++) {
          R v = R(0);
          if (j < mc.nh && ((am >> j) & 1u)) {
            const R* s = s_sd() + 8 * j;
            v = L.f[0] * s[0] + L.f[1] * s[1] + L.f[2] * s[2] + L.f[3] * s[3] + L.f[4] * s[4] + L.f[5] * s[5];
          }
          L.row[j] = (j == lane) ? L.diag : v;
        }
        GMR_UNROLL
        for (int g = 0; g < 6; g++) L.a[g] = L.f[g];
        L.rhs = -L.ci;
      }
    GMR_END
    if (pinned) {
      // pinned lanes publish (their column of H) * bound; s_bnd holds the bound values
      GMR_LANES
        if (lane < mc.nh && ((pinned >> lane) & 1u)) {
          const R bv = s_bnd()[lane];
          R* o = s_L() + gmr_loff_rt(lane);
          GMR_UNROLL
          for (int g = 0; g < 6; g++) o[g] = L.a[g] * bv;
          GMR_UNROLL
          for (int j = 0; j < GMR_NH; j++) if (j < lane) o[6 + j] = L.row[j] * bv;
        }
      GMR_END
      GMR_LANES
        if (lane < mc.nh) {
          if ((pinned >> lane) & 1u) {
            GMR_UNROLL
            for (int g = 0; g < 6; g++) L.a[g] = R(0);
            GMR_UNROLL
            for (int j = 0; j < GMR_NH; j++) L.row[j] = (j == lane) ? R(1) : R(0);
            L.rhs = s_bnd()[lane];
          } else {
            R r = L.rhs;
            GMR_UNROLL
            for (int j = 0; j < GMR_NH; j++) {
              if ((pinned >> j) & 1u) {
                if (j < lane) { r -= L.row[j] * s_bnd()[j]; L.row[j] = R(0); }
                else if (j > lane && j < mc.nh) r -= s_L()[gmr_loff_rt(j) + 6 + lane];
              }
            }
            L.rhs = r;
          }
        }
        // base right-hand side correction, gathered by lanes 0..5 into s_piv-free scratch s_xs[0..5]
        if (lane < 6) {
          R r = R(0);
          for (int j = 0; j < mc.nh; j++) if ((pinned >> j) & 1u) r += s_L()[gmr_loff_rt(j) + lane];
          s_xs()[lane] = r;
        }
      GMR_END
    }
  }

  GMR_FN static int gmr_loff_rt(int k) {      // runtime version of gmr_loff
    // sum_{i<k} pad4(6+i): rows come in groups of four equal lengths after the first two
    int o = 0;
    for (int i = 0; i < k; i++) o += (6 + i + 3) & ~3;
    return o;
  }

  // ------------------------------------------------------------------ factor + solve --------
  // Solves H x = rhs for the rows built above; result in s_xs[0..nv) (base first) and L.xs.
  GMR_FN void factor_solve(uint32_t pinned) {
    // 6x6 base block, factored redundantly by every lane (uniform data from s_root)
    GMR_LANES
      if (lane == 0) {
        const R* c = s_root();
        const R dd = mc.damping + c[27];
        R A[21];
        GMR_UNROLL
        for (int i = 0; i < 21; i++) A[i] = c[i];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) A[gmr_sym6(i, i)] += dd;
        R b[6];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) b[i] = -c[21 + i] - (pinned ? s_xs()[i] : R(0));
        // Cholesky A = Lf Lf^T, Lf stored row-wise lower: lf[i*(i+1)/2 + j]
        R lf[21];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          GMR_UNROLL
          for (int j = 0; j <= i; j++) {
            R s = A[gmr_sym6(j, i)];
            GMR_UNROLL
            for (int k = 0; k < j; k++) s -= lf[i * (i + 1) / 2 + k] * lf[j * (j + 1) / 2 + k];
            if (i == j) lf[i * (i + 1) / 2 + i] = g_sqrt(s);
            else lf[i * (i + 1) / 2 + j] = s / lf[j * (j + 1) / 2 + j];
          }
        }
        // yf = Lf^-1 b ; store Lf with reciprocal diagonal
        R y[6];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          R s = b[i];
          GMR_UNROLL
          for (int k = 0; k < i; k++) s -= lf[i * (i + 1) / 2 + k] * y[k];
          const R di = R(1) / lf[i * (i + 1) / 2 + i];
          lf[i * (i + 1) / 2 + i] = di;
          y[i] = s * di;
        }
        R* o = s_lf();
        GMR_UNROLL
        for (int i = 0; i < 21; i++) o[i] = lf[i];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) o[21 + i] = y[i];
      }
    GMR_END
    // every hinge row: a' = a Lf^-T (forward substitution on its own 6 values), rhs -= a'.yf
    GMR_LANES
      if (lane < mc.nh) {
        const R* lf = s_lf();
        R r = L.rhs;
        GMR_UNROLL
        for (int i = 0; i < 6; i++) {
          R s = L.a[i];
          GMR_UNROLL
          for (int k = 0; k < i; k++) s -= lf[i * (i + 1) / 2 + k] * L.a[k];
          s *= lf[i * (i + 1) / 2 + i];
          L.a[i] = s;
          r -= s * lf[21 + i];
        }
        L.rhs = r;
        R* o = s_L() + gmr_loff_rt(lane);
        GMR_UNROLL
        for (int g = 0; g < 6; g++) o[g] = L.a[g];
      }
    GMR_END
    // hinge block, column by column (pivot k unrolled so that row[] stays in registers)
    GMR_UNROLL
    for (int k = 0; k < GMR_NH; k++) {
      if (k < mc.nh) {
        GMR_LANES
          if (lane >= k && lane < mc.nh) {
            const R* lk = s_L() + gmr_loff(k);
            R s = L.row[k];
            GMR_UNROLL
            for (int g = 0; g < 6; g++) s -= L.a[g] * lk[g];
            GMR_UNROLL
            for (int m = 0; m < k; m++) s -= L.row[m] * lk[6 + m];
            if (lane == k) {
              const R d = g_sqrt(s), di = R(1) / d;
              L.row[k] = d; L.dinv = di;
              const R y = L.rhs * di;
              L.rhs = y;
              s_piv()[0] = di; s_piv()[1] = y;
            } else {
              L.tmp = s;
            }
          }
        GMR_END
        GMR_LANES
          if (lane > k && lane < mc.nh) {
            const R l = L.tmp * s_piv()[0];
            L.row[k] = l;
            s_L()[gmr_loff_rt(lane) + 6 + k] = l;
            L.rhs -= l * s_piv()[1];
          }
        GMR_END
      }
    }
    // back substitution: hinges from the last to the first, lanes 0..5 also accumulate A'^T x
    GMR_LANES
      L.tmp = (lane < 6) ? s_lf()[21 + lane] : R(0);      // zf accumulators (lanes 0..5)
    GMR_END
    GMR_UNROLL
    for (int kk = 0; kk < GMR_NH; kk++) {
      const int k = GMR_NH - 1 - kk;
      if (k < mc.nh) {
        GMR_LANES
          if (lane == k) { const R x = L.rhs * L.dinv; L.xs = x; s_xs()[6 + k] = x; }
        GMR_END
        GMR_LANES
          const R x = s_xs()[6 + k];
          const R* lk = s_L() + gmr_loff(k);
          if (lane < k) L.rhs -= lk[6 + lane] * x;
          if (lane < 6) L.tmp -= lk[lane] * x;
        GMR_END
      }
    }
    GMR_LANES
      if (lane < 6) s_xs()[lane] = L.tmp;   // zf = yf - A'^T x_h
    GMR_END
    GMR_LANES
      if (lane == 0) {                                     // x_f = Lf^-T zf
        const R* lf = s_lf();
        R z[6], x[6];
        GMR_UNROLL
        for (int i = 0; i < 6; i++) z[i] = s_xs()[i];
        GMR_UNROLL
        for (int ii = 0; ii < 6; ii++) {
          const int i = 5 - ii;
          R s = z[i];
          GMR_UNROLL
          for (int k = i + 1; k < 6; k++) s -= lf[k * (k + 1) / 2 + i] * x[k];
          x[i] = s * lf[i * (i + 1) / 2 + i];
        }
        GMR_UNROLL
        for (int i = 0; i < 6; i++) s_xs()[i] = x[i];
      }
    GMR_END
  }

  // ------------------------------------------------------------------ box QP (A10, A11) -----
  // min 1/2 x^T H x + c^T x, lo <= x_hinge <= hi: primal active set from clip(0, lo, hi).
  // Returns the number of factorisations; the step is left in s_xs[0..nv).
  GMR_FN int solve_qp() {
    const R INF = R(INFINITY);
    GMR_LANES
      if (lane < mc.nh) {
        if ((mc.limited_mask >> lane) & 1u) {
          const R qj = s_q()[7 + lane];
          L.bhi = mc.gain * (mc.hi[lane] - qj);
          L.blo = -(mc.gain * (qj - mc.lo[lane]));
        } else { L.bhi = INF; L.blo = -INF; }
        // feasible start: bounds that were active at the end of the previous solve stay in the
        // working set (their joints sit on, or creep towards, the limit), everything else at
        // clip(0, lo, hi).  Any feasible start gives the same (unique) optimum.
        if ((warm_hi >> lane) & 1u) L.x = L.bhi;
        else if ((warm_lo >> lane) & 1u) L.x = L.blo;
        else L.x = R(0) < L.blo ? L.blo : (R(0) > L.bhi ? L.bhi : R(0));
      }
    GMR_END
    uint32_t pin_lo = warm_lo & mc.limited_mask, pin_hi = warm_hi & mc.limited_mask;
    int nfac = 0;
    const int max_as = 4 * mc.nh + 8;
    for (int it = 0; it < max_as; it++) {
      const uint32_t pinned = pin_lo | pin_hi;
      if (pinned) {
        GMR_LANES
          if (lane < mc.nh) s_bnd()[lane] = ((pin_hi >> lane) & 1u) ? L.bhi : L.blo;
        GMR_END
      }
      build_rows(pinned);
      factor_solve(pinned);
      nfac++;
      // ratio test towards the candidate
      GMR_LANES
        if (lane < mc.nh) {
          R al = INF;
          if (!((pinned >> lane) & 1u)) {
            const R p = L.xs - L.x;
            if (p > R(0) && L.bhi < INF) al = (L.bhi - L.x) / p;
            else if (p < R(0) && L.blo > -INF) al = (L.blo - L.x) / p;
          } else {
            L.xs = L.x;                        // pinned: stays on its bound
          }
          s_red()[lane] = al;
        }
      GMR_END
      R alpha = R(1); int blk = -1;
      for (int j = 0; j < mc.nh; j++) { const R a = s_red()[j]; if (a < alpha) { alpha = a; blk = j; } }
#ifndef GMR_EMULATE
      __syncwarp();
#endif
      if (blk >= 0) {
        if (alpha < R(0)) alpha = R(0);
        int side = 0;
        GMR_LANES
          if (lane < mc.nh) {
            if (lane == blk) {
              const bool up = (L.xs - L.x) > R(0);
              L.x = up ? L.bhi : L.blo;
              s_piv()[2] = up ? R(1) : R(-1);
            } else if (!((pinned >> lane) & 1u)) {
              L.x += alpha * (L.xs - L.x);
            }
          }
        GMR_END
        side = s_piv()[2] > R(0) ? 1 : -1;
#ifndef GMR_EMULATE
        __syncwarp();
#endif
        if (side > 0) pin_hi |= 1u << blk; else pin_lo |= 1u << blk;
        continue;
      }
      GMR_LANES
        if (lane < mc.nh) L.x = L.xs;
      GMR_END
      if (!pinned) break;
      // KKT multipliers of the pinned bounds: g = H x + c on the ORIGINAL rows
      build_rows(0);
      GMR_LANES
        if (lane < mc.nh) {
          R* o = s_L() + gmr_loff_rt(lane);
          GMR_UNROLL
          for (int j = 0; j < GMR_NH; j++) if (j < lane) o[6 + j] = L.row[j] * L.x;
        }
      GMR_END
      GMR_LANES
        R lam = INF;
        if (lane < mc.nh) {
          if ((pinned >> lane) & 1u) {
            R g = L.ci + L.diag * L.x;
            GMR_UNROLL
            for (int f = 0; f < 6; f++) g += L.a[f] * s_xs()[f];
            GMR_UNROLL
            for (int j = 0; j < GMR_NH; j++) if (j < lane) g += L.row[j] * s_xs()[6 + j];
            for (int i = lane + 1; i < mc.nh; i++) g += s_L()[gmr_loff_rt(i) + 6 + lane];
            lam = ((pin_lo >> lane) & 1u) ? g : -g;
          }
          s_red()[lane] = lam;
        }
      GMR_END
      R lmin = R(0), gmax = R(1); int worst = -1;
      for (int j = 0; j < mc.nh; j++) if ((pinned >> j) & 1u) {
        const R l = s_red()[j];
        if (g_abs(l) > gmax) gmax = g_abs(l);
        if (worst < 0 || l < lmin) { lmin = l; worst = j; }
      }
#ifndef GMR_EMULATE
      __syncwarp();
#endif
      if (worst < 0 || lmin >= -GmrEps<R>::lam * gmax) break;
      pin_lo &= ~(1u << worst); pin_hi &= ~(1u << worst);
    }
    warm_lo = pin_lo; warm_hi = pin_hi;
    // publish the final step (hinge part) for integration
    GMR_LANES
      if (lane < mc.nh) s_xs()[6 + lane] = L.x;
    GMR_END
    return nfac;
  }

  // ------------------------------------------------------------------ integration (A12) ------
  // v = dq/dt, mj_integratePos(qpos, v, dt): world translation, body-local rotation
  // (right-multiplied), hinge angles.  The base rotation was solved world-aligned; it is
  // rotated into the root frame first.
  GMR_FN void integrate() {
    GMR_LANES
      const R dt = mc.dt;
      if (lane < mc.nh) {
        const R v = s_xs()[6 + lane] / dt;
        s_q()[7 + lane] += dt * v;
      }
      if (lane == 0) {
        GMR_UNROLL
        for (int i = 0; i < 3; i++) { const R v = s_xs()[i] / dt; s_q()[i] += dt * v; }
        R q[4] = {s_xq()[0], s_xq()[1], s_xq()[2], s_xq()[3]};     // normalised root quaternion
        R ww[3] = {s_xs()[3], s_xs()[4], s_xs()[5]}, wl[3];
        q_rot_inv(q, ww, wl);
        R v[3] = {wl[0] / dt, wl[1] / dt, wl[2] / dt};
        R nrm = g_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        R qr[4] = {R(1), R(0), R(0), R(0)};
        if (nrm >= R(1e-15)) {
          const R ang = dt * nrm;
          if (ang != R(0)) {
            R s, c; g_sincos(R(0.5) * ang, &s, &c);
            qr[0] = c; qr[1] = v[0] / nrm * s; qr[2] = v[1] / nrm * s; qr[3] = v[2] / nrm * s;
          }
        }
        R qq[4] = {s_q()[3], s_q()[4], s_q()[5], s_q()[6]};
        q_normalize(qq);
        R o[4]; q_mul(qq, qr, o);
        s_q()[3] = o[0]; s_q()[4] = o[1]; s_q()[5] = o[2]; s_q()[6] = o[3];
      }
    GMR_END
  }

  // ------------------------------------------------------------------ one stage (A13) --------
  // Assumes task_kinematics() is current.  Returns number of solves; *err = final stage error.
  GMR_FN int run_stage(int stage, R* err) {
    const uint32_t mask = stage == 0 ? mc.in1_mask : mc.in2_mask;
    const R* wtab = stage == 0 ? mc.w1 : mc.w2;
    R curr = stage_error(mask);
    int nsolve = 0, num_iter = 0;
    R next;
    for (;;) {
      task_inertias(wtab, mask);
      stat_refactor += solve_qp();
      integrate();
      fk();
      task_kinematics();
      next = stage_error(mask);
      nsolve++;
      if (nsolve == 1) { if (!(curr - next > mc.tol && num_iter < mc.max_iter)) break; }
      else { num_iter++; if (!(curr - next > mc.tol && num_iter < mc.max_iter)) break; }
      curr = next;
    }
    *err = next;
    return nsolve;
  }
  // ------------------------------------------------------------------ one clip ---------------
  GMR_FN void load_frame(const float* pos, const float* quat) {
    GMR_LANES
      if (lane < mc.nhum) {
        const float* p = pos + 3 * lane;
        L.in_pos[0] = p[0]; L.in_pos[1] = p[1]; L.in_pos[2] = p[2];
#ifdef GMR_EMULATE
        const float* q = quat + 4 * lane;
        L.in_quat[0] = q[0]; L.in_quat[1] = q[1]; L.in_quat[2] = q[2]; L.in_quat[3] = q[3];
#else
        const float4 q = __ldg(reinterpret_cast<const float4*>(quat) + lane);
        L.in_quat[0] = q.x; L.in_quat[1] = q.y; L.in_quat[2] = q.z; L.in_quat[3] = q.w;
#endif
      }
    GMR_END
  }

  // pos/quat point at this clip's first frame ([T,nhum,3] / [T,nhum,4]); outputs at this clip's
  // first frame too.  Null output pointers are skipped.
  template <typename IO>
  GMR_FN void run_clip(const float* pos, const float* quat, R ratio, int T, const IO* qinit, IO* qpos_out,
                       int32_t* iters_out, IO* err_out, IO* tg_out, uint32_t flags) {
    if (qinit) set_qpos(qinit); else set_qpos(mc.qpos0);
    warm_lo = warm_hi = 0;
    fk();
    if (T > 0) load_frame(pos, quat);
    for (int t = 0; t < T; t++) {
      update_targets(ratio, (flags & GMR_FLAG_OFFSET_TO_GROUND) != 0);
      if (t + 1 < T) load_frame(pos + (size_t)(t + 1) * mc.nhum * 3, quat + (size_t)(t + 1) * mc.nhum * 4);
      task_kinematics();
      int n1 = 0, n2 = 0; R e1 = R(0), e2 = R(0);
      if (mc.use1) n1 = run_stage(0, &e1);
      if (mc.use2) n2 = run_stage(1, &e2);
      GMR_LANES
        IO* qo = qpos_out + (size_t)t * mc.nq;
        for (int i = lane; i < mc.nq; i += 32) qo[i] = IO(s_q()[i]);
        if (lane == 0) {
          if (iters_out) { iters_out[2 * t] = n1; iters_out[2 * t + 1] = n2; }
          if (err_out) { err_out[2 * t] = IO(e1); err_out[2 * t + 1] = IO(e2); }
        }
        if (tg_out && lane < mc.nhum) {
          IO* o = tg_out + ((size_t)t * mc.nhum + lane) * 7;
          const R* g = s_tg() + 8 * lane;
          o[0] = IO(g[0]); o[1] = IO(g[1]); o[2] = IO(g[2]); o[3] = IO(g[4]); o[4] = IO(g[5]); o[5] = IO(g[6]); o[6] = IO(g[7]);
        }
      GMR_END
    }
  }
};
