// CPU ORACLE, native port (test infrastructure — NOT a product path).
//
// C++ port of oracle/gmr_oracle.py (which restates the reference's
// GeneralMotionRetargeting.retarget, reference
// general_motion_retargeting/motion_retarget.py:139-185, and the mink / MuJoCo / DAQP
// arithmetic it calls; see that file's header for the per-step citations and for the
// PARITY STATUS: **parity unpinned for A6-A12** — the third-party solvers are absent here).
// It exists because the NumPy restatement runs at ~20 frames/s: this port is the fair
// multi-core CPU baseline bench.py times next to the GPU and the checker
// for full-size parity runs.  Dense float64 arithmetic in the same order of operations as
// the NumPy oracle (dense 6 x nv task Jacobians, dense H, Cholesky, primal active set);
// nothing here is shared with the CUDA kernel.  `precision_bits = 32` re-runs every step in
// float32 to study rounding sensitivity of the iteration counts.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/gmr_b200.h"

namespace {

constexpr int MAXB = GMR_MAX_BODY, MAXH = GMR_MAX_HINGE, MAXV = 6 + GMR_MAX_HINGE,
              MAXT = GMR_MAX_TASK, MAXN = GMR_MAX_HUMAN;

// mink.lie.utils.get_epsilon(dtype): the threshold of every small-angle branch of the Lie code.  The value is not
// stated anywhere in /root/reference (mink is not vendored); upstream mink (jaxlie-derived lie/utils.py) is recalled
// as 1e-10 for float64 / 1e-5 for float32, round 1 assumed 10 * machine epsilon.  It is therefore a PARAMETER
// (GmrModelDesc.lie_eps; 0 = 1e-10) and the parity matrix runs with both values (tests/test_lie_eps.py).
static double g_lie_eps = 1e-10;
template <typename R> struct Eps { static R v() { return R(g_lie_eps); } };

template <typename R> struct V3 { R x, y, z; };
template <typename R> struct Q4 { R w, x, y, z; };

template <typename R> inline Q4<R> qmul(const Q4<R>& a, const Q4<R>& b) {            // mju_mulQuat
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
template <typename R> inline Q4<R> qconj(const Q4<R>& q) { return {q.w, -q.x, -q.y, -q.z}; }
template <typename R> inline Q4<R> qnormalize(const Q4<R>& q) {                       // mju_normalize4
  R n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (n < R(1e-15)) return {R(1), R(0), R(0), R(0)};
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
template <typename R> inline void q2mat(const Q4<R>& q, R m[9]) {                     // mju_quat2Mat
  R w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <typename R> inline V3<R> mv(const R m[9], const V3<R>& v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
}
template <typename R> inline V3<R> mtv(const R m[9], const V3<R>& v) {
  return {m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z};
}
template <typename R> inline V3<R> qrot(const Q4<R>& q, const V3<R>& v) { R m[9]; q2mat(q, m); return mv(m, v); }
template <typename R> inline V3<R> cross(const V3<R>& a, const V3<R>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename R> inline Q4<R> axis_angle(const V3<R>& a, R ang) {                // mju_axisAngle2Quat
  if (ang == R(0)) return {R(1), R(0), R(0), R(0)};
  R s = std::sin(R(0.5) * ang);
  return {std::cos(R(0.5) * ang), a.x * s, a.y * s, a.z * s};
}

// 3x3 helpers on row-major arrays
template <typename R> inline void skew(const R v[3], R S[9]) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
template <typename R> inline void mm3(const R A[9], const R B[9], R C[9]) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    R s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}
template <typename R> inline void tr3(const R A[9], R B[9]) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) B[3 * i + j] = A[3 * j + i]; }

template <typename R> void so3_log(const Q4<R>& q, R out[3]) {                        // mink SO3.log
  const R eps = Eps<R>::v();
  R w = q.w, nsq = q.x * q.x + q.y * q.y + q.z * q.z;
  bool taylor = nsq < eps;
  R ns = taylor ? R(1) : std::sqrt(nsq), ws = taylor ? w : R(1), f;
  R at = std::atan2(w < 0 ? -ns : ns, std::fabs(w));
  if (taylor) f = R(2) / ws - R(2) / R(3) * nsq / (ws * ws * ws);
  else if (std::fabs(w) < eps) f = (w > 0 ? R(1) : R(-1)) * R(M_PI) / ns;
  else f = R(2) * at / ns;
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

template <typename R> void se3_log(const Q4<R>& q, const V3<R>& t, R out[6]) {        // mink SE3.log
  const R eps = Eps<R>::v();
  R om[3]; so3_log(q, om);
  R th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  R S[9], S2[9]; skew(om, S); mm3(S, S, S2);
  R Vi[9];
  R coef;
  if (th2 < eps) coef = R(1) / R(12);
  else { R th = std::sqrt(th2), h = R(0.5) * th; coef = (R(1) - R(0.5) * th * std::cos(h) / std::sin(h)) / th2; }
  for (int i = 0; i < 9; i++) Vi[i] = (i % 4 == 0 ? R(1) : R(0)) - R(0.5) * S[i] + coef * S2[i];
  V3<R> r = mv(Vi, t);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

// mink's closed forms cancel catastrophically for tiny angles (see gmr_oracle.py STABLE_LIE);
// g_stable_lie switches to the well-conditioned, mathematically identical forms.
static bool g_stable_lie = false;

template <typename R> void so3_ljacinv(const R om[3], R J[9]) {                       // mink SO3.ljacinv
  const R eps = Eps<R>::v();
  R th = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]), A;
  if (th < eps) { R t2 = th * th; A = R(1) / R(12) * (R(1) + t2 / R(60) * (R(1) + t2 / R(42) * (R(1) + t2 / R(40)))); }
  else if (g_stable_lie) { R h = R(0.5) * th; A = (R(1) - h * std::cos(h) / std::sin(h)) / (th * th); }
  else A = (R(1) / (th * th)) * (R(1) - (th * std::sin(th) / (R(2) * (R(1) - std::cos(th)))));
  R S[9], S2[9]; skew(om, S); mm3(S, S, S2);
  for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? R(1) : R(0)) - R(0.5) * S[i] + A * S2[i];
}

template <typename R> void se3_ljacinv(const R xi[6], R J[36]) {                      // mink SE3.ljacinv + _getQ
  const R eps = Eps<R>::v();
  const R* rho = xi; const R* om = xi + 3;
  R th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  for (int i = 0; i < 36; i++) J[i] = (i % 7 == 0) ? R(1) : R(0);
  if (th2 < eps) return;
  R th = std::sqrt(th2), s = std::sin(th), c = std::cos(th);
  R B = (th - s) / (th2 * th), C = (R(1) - th2 / R(2) - c) / (th2 * th2),
    D = (R(2) * th - R(3) * s + th * c) / (R(2) * th2 * th2 * th);
  if (g_stable_lie && th2 < R(0.0625)) {
    B = R(1) / R(6) - th2 * (R(1) / R(120) - th2 * (R(1) / R(5040) - th2 / R(362880)));
    C = -R(1) / R(24) + th2 * (R(1) / R(720) - th2 * (R(1) / R(40320) - th2 / R(3628800)));
    D = R(1) / R(120) - th2 * (R(1) / R(2520) - th2 * (R(1) / R(120960) - th2 / R(9979200)));
  }
  R V[9], W[9], VW[9], WV[9], WVW[9], VWW[9], VWWt[9], WVWW[9], WWVW[9], Q[9];
  skew(rho, V); skew(om, W); mm3(V, W, VW); tr3(VW, WV); mm3(WV, W, WVW); mm3(VW, W, VWW); tr3(VWW, VWWt);
  mm3(WVW, W, WVWW); mm3(W, WVW, WWVW);
  for (int i = 0; i < 9; i++)
    Q[i] = R(0.5) * V[i] + B * (WV[i] + VW[i] + WVW[i]) - C * (VWW[i] - VWWt[i] - R(3) * WVW[i]) + D * (WVWW[i] + WWVW[i]);
  R Ji[9], T1[9], T2[9]; so3_ljacinv(om, Ji); mm3(Ji, Q, T1); mm3(T1, Ji, T2);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    J[6 * i + j] = Ji[3 * i + j]; J[6 * i + 3 + j] = -T2[3 * i + j]; J[6 * (i + 3) + j] = 0; J[6 * (i + 3) + 3 + j] = Ji[3 * i + j]; }
}

template <typename R> struct Model {
  int nb, nh, nv, nq, nhum, nt;
  int parent[MAXB], bhinge[MAXB];
  V3<R> bpos[MAXB]; Q4<R> bquat[MAXB];
  V3<R> axis[MAXH]; R lo[MAXH], hi[MAXH]; bool limited[MAXH];
  R qpos0[7 + MAXH];
  int hroot; R hscale[MAXN]; V3<R> hpoff[MAXN]; Q4<R> hroff[MAXN]; bool hfoot[MAXN];
  int tbody[MAXT], thuman[MAXT]; R w1[MAXT][2], w2[MAXT][2]; bool in1[MAXT], in2[MAXT];
  bool use1, use2;
  R damping, lm, gain, tol, dt; int max_iter;
};

template <typename R> void load_model(const GmrModelDesc* d, Model<R>& m) {
  m.nb = d->nbody; m.nh = d->nhinge; m.nv = 6 + m.nh; m.nq = 7 + m.nh; m.nhum = d->nhuman; m.nt = d->ntask;
  for (int b = 0; b < m.nb; b++) {
    m.parent[b] = d->body_parent[b]; m.bhinge[b] = d->body_hinge[b];
    m.bpos[b] = {R(d->body_pos[3 * b]), R(d->body_pos[3 * b + 1]), R(d->body_pos[3 * b + 2])};
    m.bquat[b] = {R(d->body_quat[4 * b]), R(d->body_quat[4 * b + 1]), R(d->body_quat[4 * b + 2]), R(d->body_quat[4 * b + 3])};
  }
  for (int j = 0; j < m.nh; j++) {
    m.axis[j] = {R(d->hinge_axis[3 * j]), R(d->hinge_axis[3 * j + 1]), R(d->hinge_axis[3 * j + 2])};
    m.lo[j] = R(d->hinge_lo[j]); m.hi[j] = R(d->hinge_hi[j]); m.limited[j] = d->hinge_limited[j] != 0;
  }
  for (int i = 0; i < m.nq; i++) m.qpos0[i] = R(d->qpos0[i]);
  m.hroot = d->human_root;
  for (int i = 0; i < m.nhum; i++) {
    m.hscale[i] = R(d->human_scale[i]);
    m.hpoff[i] = {R(d->human_pos_off[3 * i]), R(d->human_pos_off[3 * i + 1]), R(d->human_pos_off[3 * i + 2])};
    m.hroff[i] = {R(d->human_rot_off[4 * i]), R(d->human_rot_off[4 * i + 1]), R(d->human_rot_off[4 * i + 2]), R(d->human_rot_off[4 * i + 3])};
    m.hfoot[i] = d->human_foot[i] != 0;
  }
  for (int k = 0; k < m.nt; k++) {
    m.tbody[k] = d->task_body[k]; m.thuman[k] = d->task_human[k];
    m.w1[k][0] = R(d->task_w1[2 * k]); m.w1[k][1] = R(d->task_w1[2 * k + 1]);
    m.w2[k][0] = R(d->task_w2[2 * k]); m.w2[k][1] = R(d->task_w2[2 * k + 1]);
    m.in1[k] = d->task_in1[k] != 0; m.in2[k] = d->task_in2[k] != 0;
  }
  m.use1 = d->use_stage1 != 0; m.use2 = d->use_stage2 != 0;
  m.damping = R(d->damping); m.lm = R(d->lm_damping); m.gain = R(d->limit_gain); m.tol = R(d->tol); m.dt = R(d->timestep);
  m.max_iter = d->max_iter;
}

// One clip's solver state (mink.Configuration + the task targets).
template <typename R> struct Clip {
  const Model<R>& m;
  R qpos[7 + MAXH];
  V3<R> xpos[MAXB]; Q4<R> xquat[MAXB]; R xmat[MAXB][9]; V3<R> xaxis[MAXH];
  V3<R> tpos[MAXN]; Q4<R> tquat[MAXN];       // preprocessed targets per human body
  explicit Clip(const Model<R>& mm) : m(mm) {}

  void fk() {                                                                        // mj_kinematics
    for (int b = 0; b < m.nb; b++) {
      int p = m.parent[b];
      V3<R> xp; Q4<R> xq;
      if (p < 0) {
        xp = {qpos[0], qpos[1], qpos[2]};
        xq = qnormalize(Q4<R>{qpos[3], qpos[4], qpos[5], qpos[6]});
      } else {
        V3<R> o = mv(xmat[p], m.bpos[b]);
        xp = {xpos[p].x + o.x, xpos[p].y + o.y, xpos[p].z + o.z};
        xq = qmul(xquat[p], m.bquat[b]);
        int j = m.bhinge[b];
        if (j >= 0) {
          xaxis[j] = qrot(xq, m.axis[j]);
          xq = qmul(xq, axis_angle(m.axis[j], qpos[7 + j]));
        }
        xq = qnormalize(xq);
      }
      xpos[b] = xp; xquat[b] = xq; q2mat(xq, xmat[b]);
    }
  }

  void task_error(int k, R e[6]) const {                                              // FrameTask.compute_error
    int b = m.tbody[k], h = m.thuman[k];
    Q4<R> qi = qconj(xquat[b]);
    Q4<R> qbt = qmul(qi, tquat[h]);
    V3<R> a = qrot(qi, tpos[h]), c = qrot(qi, xpos[b]);
    V3<R> t = {a.x + (-c.x), a.y + (-c.y), a.z + (-c.z)};
    se3_log(qbt, t, e);
  }

  void frame_jacobian(int body, R J[6 * MAXV]) const {                                 // mj_jacBody + mink body-frame rotation
    const int nv = m.nv;
    R jp[3 * MAXV], jr[3 * MAXV];
    std::memset(jp, 0, sizeof(jp)); std::memset(jr, 0, sizeof(jr));
    V3<R> pt = xpos[body];
    for (int b = body; b >= 0; b = m.parent[b]) {
      int j = m.bhinge[b];
      if (j >= 0) {
        V3<R> ax = xaxis[j], d = {pt.x - xpos[b].x, pt.y - xpos[b].y, pt.z - xpos[b].z}, cp = cross(ax, d);
        jr[0 * nv + 6 + j] = ax.x; jr[1 * nv + 6 + j] = ax.y; jr[2 * nv + 6 + j] = ax.z;
        jp[0 * nv + 6 + j] = cp.x; jp[1 * nv + 6 + j] = cp.y; jp[2 * nv + 6 + j] = cp.z;
      }
      if (m.parent[b] < 0) {
        jp[0 * nv + 0] = 1; jp[1 * nv + 1] = 1; jp[2 * nv + 2] = 1;
        V3<R> d = {pt.x - xpos[b].x, pt.y - xpos[b].y, pt.z - xpos[b].z};
        for (int k = 0; k < 3; k++) {
          V3<R> col = {xmat[b][k], xmat[b][3 + k], xmat[b][6 + k]}, cp = cross(col, d);
          jr[0 * nv + 3 + k] = col.x; jr[1 * nv + 3 + k] = col.y; jr[2 * nv + 3 + k] = col.z;
          jp[0 * nv + 3 + k] = cp.x; jp[1 * nv + 3 + k] = cp.y; jp[2 * nv + 3 + k] = cp.z;
        }
      }
    }
    const R* Rb = xmat[body];
    for (int c = 0; c < nv; c++) {
      V3<R> a = mtv(Rb, V3<R>{jp[c], jp[nv + c], jp[2 * nv + c]});
      V3<R> w = mtv(Rb, V3<R>{jr[c], jr[nv + c], jr[2 * nv + c]});
      J[0 * nv + c] = a.x; J[1 * nv + c] = a.y; J[2 * nv + c] = a.z;
      J[3 * nv + c] = w.x; J[4 * nv + c] = w.y; J[5 * nv + c] = w.z;
    }
  }

  void task_jacobian(int k, R J[6 * MAXV]) const {                                     // FrameTask.compute_jacobian
    const int nv = m.nv;
    int b = m.tbody[k], h = m.thuman[k];
    R Jb[6 * MAXV]; frame_jacobian(b, Jb);
    Q4<R> ti = qconj(tquat[h]);
    Q4<R> qtb = qmul(ti, xquat[b]);
    V3<R> a = qrot(ti, xpos[b]), c = qrot(ti, tpos[h]);
    V3<R> t = {a.x + (-c.x), a.y + (-c.y), a.z + (-c.z)};
    R lg[6]; se3_log(qtb, t, lg);
    for (int i = 0; i < 6; i++) lg[i] = -lg[i];
    R JL[36]; se3_ljacinv(lg, JL);                   // jlog(T) = ljacinv(-log T)
    for (int r = 0; r < 6; r++) for (int cidx = 0; cidx < nv; cidx++) {
      R s = 0; for (int i = 0; i < 6; i++) s += JL[6 * r + i] * Jb[i * nv + cidx];
      J[r * nv + cidx] = -s;
    }
  }

  R stage_error(int stage) const {                                                     // error1 / error2
    R s = 0;
    for (int k = 0; k < m.nt; k++) {
      if (!(stage == 0 ? m.in1[k] : m.in2[k])) continue;
      R e[6]; task_error(k, e);
      for (int i = 0; i < 6; i++) s += e[i] * e[i];
    }
    return std::sqrt(s);
  }

  // exact box QP by primal active set (stands in for DAQP)
  bool solve_box_qp(const R* H, const R* c, const R* lo, const R* hi, R* x) const {
    const int n = m.nv;
    const R INF = std::numeric_limits<R>::infinity();
    int W[MAXV];
    for (int i = 0; i < n; i++) { x[i] = R(0) < lo[i] ? lo[i] : (R(0) > hi[i] ? hi[i] : R(0)); W[i] = 0; }
    R L[MAXV * MAXV], rhs[MAXV], xs[MAXV], g[MAXV]; int idx[MAXV];
    for (int it = 0; it < 10 * n + 10; it++) {
      int nf = 0; for (int i = 0; i < n; i++) if (W[i] == 0) idx[nf++] = i;
      for (int i = 0; i < n; i++) xs[i] = x[i];
      if (nf) {
        for (int a = 0; a < nf; a++) {
          R s = c[idx[a]];
          for (int j = 0; j < n; j++) if (W[j] != 0) s += H[idx[a] * n + j] * x[j];
          rhs[a] = -s;
        }
        for (int a = 0; a < nf; a++) for (int b = 0; b <= a; b++) {                   // Cholesky of H_FF
          R s = H[idx[a] * n + idx[b]];
          for (int k = 0; k < b; k++) s -= L[a * nf + k] * L[b * nf + k];
          if (a == b) { if (!(s > 0)) return false; L[a * nf + a] = std::sqrt(s); } else L[a * nf + b] = s / L[b * nf + b];
        }
        for (int a = 0; a < nf; a++) { R s = rhs[a]; for (int k = 0; k < a; k++) s -= L[a * nf + k] * rhs[k]; rhs[a] = s / L[a * nf + a]; }
        for (int a = nf - 1; a >= 0; a--) { R s = rhs[a]; for (int k = a + 1; k < nf; k++) s -= L[k * nf + a] * rhs[k]; rhs[a] = s / L[a * nf + a]; }
        for (int a = 0; a < nf; a++) xs[idx[a]] = rhs[a];
      }
      R alpha = 1; int blk = -1, side = 0;
      for (int a = 0; a < nf; a++) {
        int i = idx[a]; R p = xs[i] - x[i];
        if (p > 0 && hi[i] < INF) { R al = (hi[i] - x[i]) / p; if (al < alpha) { alpha = al; blk = i; side = 1; } }
        else if (p < 0 && lo[i] > -INF) { R al = (lo[i] - x[i]) / p; if (al < alpha) { alpha = al; blk = i; side = -1; } }
      }
      if (blk >= 0) {
        if (alpha < 0) alpha = 0;
        for (int i = 0; i < n; i++) x[i] += alpha * (xs[i] - x[i]);
        x[blk] = side > 0 ? hi[blk] : lo[blk]; W[blk] = side;
        continue;
      }
      for (int i = 0; i < n; i++) x[i] = xs[i];
      R gmax = 1; int worst = -1; R lmin = 0;
      for (int i = 0; i < n; i++) { R s = c[i]; for (int j = 0; j < n; j++) s += H[i * n + j] * x[j]; g[i] = s; if (std::fabs(s) > gmax) gmax = std::fabs(s); }
      for (int i = 0; i < n; i++) if (W[i] != 0) { R lam = W[i] < 0 ? g[i] : -g[i]; if (worst < 0 || lam < lmin) { lmin = lam; worst = i; } }
      if (worst < 0 || lmin >= -R(1e-12) * gmax) return true;
      W[worst] = 0;
    }
    return false;
  }

  bool solve_and_integrate(int stage) {                                               // mink.solve_ik + integrate_inplace
    const int nv = m.nv;
    R H[MAXV * MAXV], c[MAXV], J[6 * MAXV], WJ[6 * MAXV];
    for (int i = 0; i < nv * nv; i++) H[i] = 0;
    for (int i = 0; i < nv; i++) { H[i * nv + i] = m.damping; c[i] = 0; }
    for (int k = 0; k < m.nt; k++) {
      if (!(stage == 0 ? m.in1[k] : m.in2[k])) continue;
      const R* w = stage == 0 ? m.w1[k] : m.w2[k];
      task_jacobian(k, J);
      R e[6], We[6]; task_error(k, e);
      for (int r = 0; r < 6; r++) { R wr = w[r / 3]; We[r] = wr * (-e[r]); for (int i = 0; i < nv; i++) WJ[r * nv + i] = wr * J[r * nv + i]; }
      R mu = 0; for (int r = 0; r < 6; r++) mu += We[r] * We[r]; mu *= m.lm;
      for (int i = 0; i < nv; i++) {
        for (int j = 0; j < nv; j++) { R s = 0; for (int r = 0; r < 6; r++) s += WJ[r * nv + i] * WJ[r * nv + j]; H[i * nv + j] += s; }
        H[i * nv + i] += mu;
        R s = 0; for (int r = 0; r < 6; r++) s += We[r] * WJ[r * nv + i];
        c[i] += -s;
      }
    }
    const R INF = std::numeric_limits<R>::infinity();
    R lo[MAXV], hi[MAXV], dq[MAXV];
    for (int i = 0; i < 6; i++) { lo[i] = -INF; hi[i] = INF; }
    for (int j = 0; j < m.nh; j++) {
      if (m.limited[j]) { hi[6 + j] = m.gain * (m.hi[j] - qpos[7 + j]); lo[6 + j] = -(m.gain * (qpos[7 + j] - m.lo[j])); }
      else { lo[6 + j] = -INF; hi[6 + j] = INF; }
    }
    if (!solve_box_qp(H, c, lo, hi, dq)) return false;
    // v = dq / dt ; mj_integratePos(qpos, v, dt)
    R v[MAXV]; for (int i = 0; i < nv; i++) v[i] = dq[i] / m.dt;
    for (int i = 0; i < 3; i++) qpos[i] += m.dt * v[i];
    R nrm = std::sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
    V3<R> ax;
    if (nrm < R(1e-15)) { ax = {R(1), R(0), R(0)}; nrm = 0; } else ax = {v[3] / nrm, v[4] / nrm, v[5] / nrm};
    Q4<R> qr = axis_angle(ax, m.dt * nrm);
    Q4<R> q = qmul(qnormalize(Q4<R>{qpos[3], qpos[4], qpos[5], qpos[6]}), qr);
    qpos[3] = q.w; qpos[4] = q.x; qpos[5] = q.y; qpos[6] = q.z;
    for (int j = 0; j < m.nh; j++) qpos[7 + j] += m.dt * v[6 + j];
    fk();
    return true;
  }

  // scale_human_data + offset_human_data (+ offset_human_data_to_ground)
  void update_targets(const float* pos, const float* quat, R ratio, bool to_ground) {
    const int n = m.nhum;
    V3<R> root = {R(pos[3 * m.hroot]), R(pos[3 * m.hroot + 1]), R(pos[3 * m.hroot + 2])};
    R sr = m.hscale[m.hroot] * ratio;
    V3<R> sroot = {sr * root.x, sr * root.y, sr * root.z};
    for (int i = 0; i < n; i++) {
      V3<R> p;
      if (i == m.hroot) p = sroot;
      else {
        R s = m.hscale[i] * ratio;
        p = {(R(pos[3 * i]) - root.x) * s + sroot.x, (R(pos[3 * i + 1]) - root.y) * s + sroot.y, (R(pos[3 * i + 2]) - root.z) * s + sroot.z};
      }
      Q4<R> q = {R(quat[4 * i]), R(quat[4 * i + 1]), R(quat[4 * i + 2]), R(quat[4 * i + 3])};
      R nn = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
      q = {q.w / nn, q.x / nn, q.y / nn, q.z / nn};
      Q4<R> u = qmul(q, m.hroff[i]);
      nn = std::sqrt(u.w * u.w + u.x * u.x + u.y * u.y + u.z * u.z);
      u = {u.w / nn, u.x / nn, u.y / nn, u.z / nn};
      V3<R> g = qrot(u, m.hpoff[i]);
      tpos[i] = {p.x + g.x, p.y + g.y, p.z + g.z}; tquat[i] = u;
    }
    if (to_ground) {
      R lowest = std::numeric_limits<R>::infinity();
      for (int i = 0; i < n; i++) if (m.hfoot[i] && tpos[i].z < lowest) lowest = tpos[i].z;
      for (int i = 0; i < n; i++) tpos[i].z = tpos[i].z - lowest + R(0.1);
    }
  }

  bool run_stage(int stage, int& nsolve, R& err) {                                    // motion_retarget.py:143-182
    R curr = stage_error(stage);
    if (!solve_and_integrate(stage)) return false;
    R next = stage_error(stage);
    nsolve = 1; int it = 0;
    while (curr - next > m.tol && it < m.max_iter) {
      curr = next;
      if (!solve_and_integrate(stage)) return false;
      next = stage_error(stage);
      it++; nsolve++;
    }
    err = next;
    return true;
  }
};

template <typename R>
int run_batch(const GmrModelDesc* d, const float* pos, const float* quat, const float* ratio, int C, int T,
              const double* qpos_init, double* qpos_out, int32_t* iters_out, double* err_out, uint32_t flags, int nthreads) {
  auto* mp = new Model<R>();
  load_model(d, *mp);
  const Model<R>& m = *mp;
  std::atomic<int> next{0}, failed{0};
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > C) nthreads = C > 0 ? C : 1;
  auto work = [&]() {
    auto* clip = new Clip<R>(m);
    for (;;) {
      int c = next.fetch_add(1);
      if (c >= C) break;
      for (int i = 0; i < m.nq; i++) clip->qpos[i] = qpos_init ? R(qpos_init[(size_t)c * m.nq + i]) : m.qpos0[i];
      clip->fk();
      R rt = ratio ? R(ratio[c]) : R(1);
      for (int t = 0; t < T; t++) {
        size_t f = (size_t)c * T + t;
        clip->update_targets(pos + f * m.nhum * 3, quat + f * m.nhum * 4, rt, (flags & GMR_FLAG_OFFSET_TO_GROUND) != 0);
        int n1 = 0, n2 = 0; R e1 = 0, e2 = 0; bool ok = true;
        if (m.use1) ok = clip->run_stage(0, n1, e1);
        if (ok && m.use2) ok = clip->run_stage(1, n2, e2);
        if (!ok) failed.store(1);
        for (int i = 0; i < m.nq; i++) qpos_out[f * m.nq + i] = (double)clip->qpos[i];
        if (iters_out) { iters_out[2 * f] = n1; iters_out[2 * f + 1] = n2; }
        if (err_out) { err_out[2 * f] = (double)e1; err_out[2 * f + 1] = (double)e2; }
      }
    }
    delete clip;
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  delete mp;
  return failed.load() ? GMR_EINVAL : GMR_OK;
}

}  // namespace

extern "C" int gmr_oracle_retarget_batch(const GmrModelDesc* desc, const float* pos, const float* quat,
                                         const float* ratio, int32_t C, int32_t T, const double* qpos_init,
                                         double* qpos_out, int32_t* iters_out, double* err_out,
                                         uint32_t flags, int32_t nthreads, int32_t precision_bits) {
  if (!desc || !pos || !quat || !qpos_out || C < 0 || T < 0) return GMR_EINVAL;
  if (desc->nbody > MAXB || desc->nhinge > MAXH || desc->nhuman > MAXN || desc->ntask > MAXT) return GMR_ELIMIT;
  g_stable_lie = (flags & GMR_ORACLE_FLAG_STABLE_LIE) != 0;
  g_lie_eps = desc->lie_eps > 0 ? desc->lie_eps : 1e-10;
  if (precision_bits == 32)
    return run_batch<float>(desc, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, flags, nthreads);
  return run_batch<double>(desc, pos, quat, ratio, C, T, qpos_init, qpos_out, iters_out, err_out, flags, nthreads);
}
