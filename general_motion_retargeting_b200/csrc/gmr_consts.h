// Device-side constant block of one (source format, robot) pair: the flat tables the solve
// kernel stages into shared memory once per CTA.  Filled on the host by
// gmr_fill_consts() from a GmrModelDesc (include/gmr_b200.h).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/gmr_b200.h"

#define GMR_NH 32            // lanes of a warp = hinge DoFs (GMR_MAX_HINGE)
#define GMR_MAX_LEVEL 64
#define GMR_MAXD 10          // deepest chain of hinges root -> leaf the factorisation is unrolled for (G1: 10)
#define GMR_NS (6 + GMR_MAXD) // row slots of one hinge: 6 floating-base columns, then its ancestors by depth
#define GMR_ANCS 12          // bytes per row of the ancestor table (GMR_MAXD rounded up to whole words)

template <typename R>
struct GmrConsts {
  int32_t nb, nh, nhum, nt, nlevel, nq, nv, hroot;
  int32_t use1, use2, max_iter, _pad0;
  uint32_t in1_mask, in2_mask, limited_mask, foot_mask;
  R damping, lm, gain, tol, dt, lie, _pad1[2];
  int16_t parent[GMR_MAX_BODY];       // body -> parent body (-1 root)
  int16_t bhinge[GMR_MAX_BODY];       // body -> hinge index or -1
  int16_t lvl_body[GMR_MAX_BODY];     // bodies sorted by tree depth
  int16_t lvl_off[GMR_MAX_LEVEL + 2]; // lvl_body[lvl_off[l] .. lvl_off[l+1]) is depth l
  int16_t hbody[GMR_NH];              // hinge -> body
  int16_t tbody[GMR_MAX_TASK];        // task -> robot body
  int16_t thuman[GMR_MAX_TASK];       // task -> human body
  uint32_t anc_mask[GMR_NH];          // hinge i -> bitmask of hinges that are strict ancestors of i
  uint32_t desc_mask[GMR_NH];         // hinge i -> bitmask of hinges that are strict descendants of i
  uint32_t task_mask[GMR_NH];         // hinge i -> bitmask of tasks whose body lies in i's subtree
  uint32_t top_mask[GMR_NH];          // ... restricted to tasks whose parent task is outside i's subtree
  uint32_t troot_mask, _pad3[3];      // tasks without a parent task
  int8_t tpar[GMR_MAX_TASK];          // task -> nearest task on a strict ancestor body (task tree), -1 if none
  // leaf-to-root accumulation of the task tree in rounds of 4 independent (child, parent) pairs:
  // tround[4 r + q] = child | parent << 8; 0xffff = unused slot
  int32_t ntround, _pad4[3];
  uint16_t tround[GMR_MAX_TASK * 4];
  int32_t maxd, tlmax, _pad2[2];      // deepest hinge chain; most tasks in any hinge's subtree
  uint32_t lvl_mask[GMR_MAXD + 2];    // [d] -> bitmask of the hinges of depth d (1-based)
  uint8_t hdepth[GMR_NH];             // hinge i -> number of hinges on the chain root .. i (inclusive); 0 for unused lanes
  // [i][d-1] -> the hinge at depth d on i's chain (d <= hdepth[i]; own index from d = hdepth[i] on); rows of GMR_ANCS = 12 bytes,
  // 4-byte aligned: a lane fetches its whole row as three words
  alignas(4) uint8_t anc_of[GMR_NH * GMR_ANCS];
  // Forward-kinematics records.  FK is a pointer-jumping scan over the tree of MOVING bodies (hinge bodies + the root):
  // fixed bodies between two hinges are folded into constants on the host.
  //   entries [0, nh): hinge j.  Its body's pose in the frame of its nearest moving ancestor is
  //     ( P, cos(q/2) * A + sin(q/2) * B ),  A = (fixed transforms) * bquat,  B = A * (0, axis);
  //     fk_idx[j] = four bytes: the BODY index of the moving ancestor at distance 1, 2, 4, 8 (0xff: none)
  //   entries [nh, nh + nfix): a fixed body: fk_idx = body | body of its nearest moving ancestor << 8, (fk_pos, fk_A) = its
  //     constant pose in that ancestor's frame
  uint32_t fk_idx[GMR_MAX_BODY];
  alignas(16) R fk_pos[GMR_MAX_BODY * 4];   // 16-byte aligned: read with 128-bit loads
  alignas(16) R fk_A[GMR_MAX_BODY * 4];
  alignas(16) R fk_B[GMR_MAX_BODY * 4];
  R axis[GMR_NH * 3];
  R lo[GMR_NH], hi[GMR_NH];
  R qpos0[8 + GMR_NH];
  R hscale[GMR_MAX_HUMAN];
  R hpoff[GMR_MAX_HUMAN * 3];
  R hroff[GMR_MAX_HUMAN * 4];
  R w1[GMR_MAX_TASK * 2];
  R w2[GMR_MAX_TASK * 2];
};

// Scalars of the model that every thread needs all the time.  They travel as kernel
// parameters (constant bank), so they cost no registers and no shared-memory reads.
struct GmrDims {
  int32_t nb, nh, nhum, nt, nlevel, nq, hroot, use1, use2, max_iter;
  uint32_t in1_mask, in2_mask, limited_mask, foot_mask;
  int32_t maxd, tlmax;
  // per-warp shared-memory layout for the launching precision (element offsets, see gmr_solver.cuh)
  int32_t o_tg, o_in, o_xp, o_xq, o_u, warp_elems;
  int32_t o_y, rs;            // published factor row of a hinge: header (y, 1/d), then slots [0, o_y); stride rs elements
  int32_t o_sd, _pad_dims;    // spatial axes
};
// `lie`: the reference's small-angle threshold (GmrModelDesc.lie_eps), the same float64-derived value in both precisions:
// the branch decisions (Taylor vs closed form, the jlog = I shortcut) must follow the float64 reference, not the kernel's dtype
template <typename R> struct GmrScal { R damping, lm, gain, tol, dt, inv_dt, lie, _pad; };

template <typename R> inline GmrDims gmr_dims_of(const GmrConsts<R>& c) {
  GmrDims d;
  d.nb = c.nb; d.nh = c.nh; d.nhum = c.nhum; d.nt = c.nt; d.nlevel = c.nlevel; d.nq = c.nq; d.hroot = c.hroot;
  d.use1 = c.use1; d.use2 = c.use2; d.max_iter = c.max_iter;
  d.in1_mask = c.in1_mask; d.in2_mask = c.in2_mask; d.limited_mask = c.limited_mask; d.foot_mask = c.foot_mask;
  d.maxd = c.maxd; d.tlmax = c.tlmax;
  d.o_tg = d.o_in = d.o_xp = d.o_xq = d.o_u = d.warp_elems = d.o_y = d.rs = d.o_sd = d._pad_dims = 0;
  return d;
}
template <typename R> inline GmrScal<R> gmr_scal_of(const GmrConsts<R>& c) {
  GmrScal<R> s; s.damping = c.damping; s.lm = c.lm; s.gain = c.gain; s.tol = c.tol; s.dt = c.dt; s.inv_dt = R(1) / c.dt; s.lie = c.lie; s._pad = R(0); return s;
}

// Validates `d` and fills `c`.  Returns GMR_OK or a negative GMR_E* code; `why` (may be null)
// receives a static message.
template <typename R>
inline int gmr_fill_consts(const GmrModelDesc* d, GmrConsts<R>* c, const char** why) {
  auto fail = [&](int code, const char* msg) { if (why) *why = msg; return code; };
  if (!d) return fail(GMR_EINVAL, "null model description");
  if (d->nbody < 1 || d->nhinge < 0 || d->nhuman < 1 || d->ntask < 1) return fail(GMR_EINVAL, "empty model");
  if (d->nbody > GMR_MAX_BODY) return fail(GMR_ELIMIT, "too many bodies (GMR_MAX_BODY)");
  if (d->nhinge > GMR_MAX_HINGE) return fail(GMR_ELIMIT, "too many hinges (GMR_MAX_HINGE)");
  if (d->nhinge > GMR_NH - 1) return fail(GMR_ELIMIT, "more than 31 hinges: the FK scan keeps one lane for the floating root");
  if (d->nhuman > GMR_MAX_HUMAN) return fail(GMR_ELIMIT, "too many human bodies (GMR_MAX_HUMAN)");
  if (d->ntask > GMR_MAX_TASK) return fail(GMR_ELIMIT, "too many tasks (GMR_MAX_TASK)");
  if (!d->body_parent || !d->body_pos || !d->body_quat || !d->body_hinge || !d->qpos0 || !d->human_scale ||
      !d->human_pos_off || !d->human_rot_off || !d->human_foot || !d->task_body || !d->task_human || !d->task_w1 ||
      !d->task_w2 || !d->task_in1 || !d->task_in2 || (d->nhinge > 0 && (!d->hinge_axis || !d->hinge_lo || !d->hinge_hi || !d->hinge_limited)))
    return fail(GMR_EINVAL, "null array in model description");
  if (d->human_root < 0 || d->human_root >= d->nhuman) return fail(GMR_EINVAL, "human_root out of range");
  if (!(d->timestep > 0)) return fail(GMR_EINVAL, "timestep must be positive");

  GmrConsts<R>& m = *c;
  // zero everything (padding included) so the block is reproducible byte for byte
  { unsigned char* p = reinterpret_cast<unsigned char*>(c); for (size_t i = 0; i < sizeof(GmrConsts<R>); i++) p[i] = 0; }
  m.nb = d->nbody; m.nh = d->nhinge; m.nhum = d->nhuman; m.nt = d->ntask;
  m.nq = 7 + d->nhinge; m.nv = 6 + d->nhinge; m.hroot = d->human_root;
  m.use1 = d->use_stage1 != 0; m.use2 = d->use_stage2 != 0; m.max_iter = d->max_iter;
  m.damping = R(d->damping); m.lm = R(d->lm_damping); m.gain = R(d->limit_gain); m.tol = R(d->tol); m.dt = R(d->timestep);
  m.lie = R(d->lie_eps > 0 ? d->lie_eps : 1e-10);

  int depth[GMR_MAX_BODY];
  int maxdepth = 0;
  for (int b = 0; b < m.nb; b++) {
    int p = d->body_parent[b];
    if (b == 0) { if (p != -1) return fail(GMR_EINVAL, "body 0 must be the root (parent -1)"); depth[b] = 0; }
    else { if (p < 0 || p >= b) return fail(GMR_EINVAL, "bodies must be ordered parent-before-child"); depth[b] = depth[p] + 1; }
    if (depth[b] > maxdepth) maxdepth = depth[b];
    m.parent[b] = (int16_t)p;
    int j = d->body_hinge[b];
    if (j < -1 || j >= m.nh || (b == 0 && j != -1)) return fail(GMR_EINVAL, "body_hinge out of range");
    m.bhinge[b] = (int16_t)j;
  }
  if (maxdepth + 1 > GMR_MAX_LEVEL) return fail(GMR_ELIMIT, "kinematic tree too deep");
  m.nlevel = maxdepth + 1;
  { int n = 0;
    for (int l = 0; l <= maxdepth; l++) { m.lvl_off[l] = (int16_t)n; for (int b = 0; b < m.nb; b++) if (depth[b] == l) m.lvl_body[n++] = (int16_t)b; }
    m.lvl_off[maxdepth + 1] = (int16_t)n; }
  for (int j = 0; j < m.nh; j++) m.hbody[j] = -1;
  for (int b = 0; b < m.nb; b++) if (m.bhinge[b] >= 0) {
    if (m.hbody[m.bhinge[b]] != -1) return fail(GMR_EINVAL, "hinge owned by two bodies");
    m.hbody[m.bhinge[b]] = (int16_t)b;
  }
  int prev_body = 0;
  for (int j = 0; j < m.nh; j++) {
    if (m.hbody[j] < 0) return fail(GMR_EINVAL, "hinge without a body");
    if (m.hbody[j] <= prev_body && j > 0) return fail(GMR_EINVAL, "hinges must be numbered in body order");
    prev_body = m.hbody[j];
    for (int k = 0; k < 3; k++) m.axis[3 * j + k] = R(d->hinge_axis[3 * j + k]);
    m.lo[j] = R(d->hinge_lo[j]); m.hi[j] = R(d->hinge_hi[j]);
    if (d->hinge_limited[j]) { if (d->hinge_lo[j] > d->hinge_hi[j]) return fail(GMR_EINVAL, "hinge range lo > hi"); m.limited_mask |= 1u << j; }
    uint32_t anc = 0;
    for (int b = m.parent[m.hbody[j]]; b >= 0; b = m.parent[b]) if (m.bhinge[b] >= 0) anc |= 1u << m.bhinge[b];
    m.anc_mask[j] = anc;
  }
  // hinge depths, chains by depth, descendants, hinges grouped by depth (the elimination order of the
  // branch-sparse factorisation: deepest hinges first)
  m.maxd = 0;
  for (int j = 0; j < m.nh; j++) {
    int chain[GMR_MAX_BODY], n = 0;
    for (int b = m.hbody[j]; b >= 0; b = m.parent[b]) if (m.bhinge[b] >= 0) chain[n++] = m.bhinge[b];
    if (n > GMR_MAXD) return fail(GMR_ELIMIT, "hinge chain too deep (GMR_MAXD)");
    m.hdepth[j] = (uint8_t)n;
    if (n > m.maxd) m.maxd = n;
    for (int d = 1; d <= n; d++) m.anc_of[j * GMR_ANCS + d - 1] = (uint8_t)chain[n - d];
    for (int d = n + 1; d <= GMR_ANCS; d++) m.anc_of[j * GMR_ANCS + d - 1] = (uint8_t)j;
    for (int i = 0; i < m.nh; i++) if ((m.anc_mask[j] >> i) & 1u) m.desc_mask[i] |= 1u << j;
  }
  for (int j = 0; j < m.nh; j++) m.lvl_mask[m.hdepth[j]] |= 1u << j;
  { // FK scan tables (see fk_idx above): fold the fixed bodies above every body into a constant transform
    auto qmul = [](const double* a, const double* b, double* o) {
      const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
      o[0] = w; o[1] = x; o[2] = y; o[3] = z;
    };
    auto qrot = [](const double* q, const double* v, double* o) {
      const double tx = 2 * (q[2] * v[2] - q[3] * v[1]), ty = 2 * (q[3] * v[0] - q[1] * v[2]), tz = 2 * (q[1] * v[1] - q[2] * v[0]);
      const double ox = v[0] + q[0] * tx + (q[2] * tz - q[3] * ty), oy = v[1] + q[0] * ty + (q[3] * tx - q[1] * tz),
                   oz = v[2] + q[0] * tz + (q[1] * ty - q[2] * tx);
      o[0] = ox; o[1] = oy; o[2] = oz;
    };
    int anchor[GMR_MAX_BODY];            // nearest moving strict ancestor (a hinge body or the root), -1 for the root
    int nfix = 0;
    anchor[0] = -1;
    for (int b = 1; b < m.nb; b++) {
      double pa[3] = {d->body_pos[3 * b], d->body_pos[3 * b + 1], d->body_pos[3 * b + 2]};
      double qa[4] = {d->body_quat[4 * b], d->body_quat[4 * b + 1], d->body_quat[4 * b + 2], d->body_quat[4 * b + 3]};
      int cur = m.parent[b];
      while (cur > 0 && m.bhinge[cur] < 0) {            // a fixed body in between: compose its constant pose in front
        double t[3]; qrot(d->body_quat + 4 * cur, pa, t);
        for (int k = 0; k < 3; k++) pa[k] = d->body_pos[3 * cur + k] + t[k];
        double q2[4]; qmul(d->body_quat + 4 * cur, qa, q2);
        for (int k = 0; k < 4; k++) qa[k] = q2[k];
        cur = m.parent[cur];
      }
      anchor[b] = cur;
      const int j = m.bhinge[b];
      const int e = j >= 0 ? j : m.nh + nfix++;
      for (int k = 0; k < 3; k++) m.fk_pos[4 * e + k] = R(pa[k]);
      for (int k = 0; k < 4; k++) m.fk_A[4 * e + k] = R(qa[k]);
      if (j >= 0) {
        const double ax[4] = {0.0, d->hinge_axis[3 * j], d->hinge_axis[3 * j + 1], d->hinge_axis[3 * j + 2]};
        double B[4]; qmul(qa, ax, B);
        for (int k = 0; k < 4; k++) m.fk_B[4 * e + k] = R(B[k]);
      } else {
        m.fk_idx[e] = (uint32_t)b | ((uint32_t)cur << 8);
      }
    }
    for (int j = 0; j < m.nh; j++) {                    // moving ancestors at distance 1, 2, 4, 8
      uint32_t packed = 0;
      for (int k = 0; k < 4; k++) {
        int a = m.hbody[j];
        for (int s = 0; s < (1 << k) && a >= 0; s++) a = anchor[a];
        packed |= (uint32_t)(a >= 0 ? a : 0xff) << (8 * k);
      }
      m.fk_idx[j] = packed;
    }
  }
  for (int i = 0; i < m.nq; i++) m.qpos0[i] = R(d->qpos0[i]);
  for (int i = 0; i < m.nhum; i++) {
    m.hscale[i] = R(d->human_scale[i]);
    for (int k = 0; k < 3; k++) m.hpoff[3 * i + k] = R(d->human_pos_off[3 * i + k]);
    for (int k = 0; k < 4; k++) m.hroff[4 * i + k] = R(d->human_rot_off[4 * i + k]);
    if (d->human_foot[i]) m.foot_mask |= 1u << i;
  }
  for (int t = 0; t < m.nt; t++) {
    int b = d->task_body[t], h = d->task_human[t];
    if (b < 0 || b >= m.nb || h < 0 || h >= m.nhum) return fail(GMR_EINVAL, "task index out of range");
    m.tbody[t] = (int16_t)b; m.thuman[t] = (int16_t)h;
    m.w1[2 * t] = R(d->task_w1[2 * t]); m.w1[2 * t + 1] = R(d->task_w1[2 * t + 1]);
    m.w2[2 * t] = R(d->task_w2[2 * t]); m.w2[2 * t + 1] = R(d->task_w2[2 * t + 1]);
    if (d->task_in1[t]) m.in1_mask |= 1u << t;
    if (d->task_in2[t]) m.in2_mask |= 1u << t;
    for (int bb = b; bb >= 0; bb = m.parent[bb]) if (m.bhinge[bb] >= 0) m.task_mask[m.bhinge[bb]] |= 1u << t;
  }
  // task tree: the composite of a hinge is the sum of the task-subtree composites of its "top" tasks
  for (int t = 0; t < m.nt; t++) {
    int best = -1;
    for (int bb = m.tbody[t]; bb >= 0 && best < 0; bb = m.parent[bb]) {
      for (int u = 0; u < m.nt; u++) {
        if (u == t) continue;
        // a task on the same body counts as an ancestor only if it comes first (keeps the relation acyclic)
        if (m.tbody[u] == bb && (bb != m.tbody[t] || u < t)) { best = u; break; }
      }
    }
    m.tpar[t] = (int8_t)best;
    if (best < 0) m.troot_mask |= 1u << t;
  }
  { // list scheduling: a pair (t -> tpar[t]) is ready once every child of t has been added into t; a round takes
    // up to 4 ready pairs with distinct parents
    int pending[GMR_MAX_TASK];                      // children not yet accumulated into t
    bool done[GMR_MAX_TASK];
    for (int t = 0; t < m.nt; t++) { pending[t] = 0; done[t] = m.tpar[t] < 0; }
    for (int t = 0; t < m.nt; t++) if (m.tpar[t] >= 0) pending[m.tpar[t]]++;
    int left = 0;
    for (int t = 0; t < m.nt; t++) if (!done[t]) left++;
    m.ntround = 0;
    while (left > 0) {
      int q = 0, chosen[4];
      uint32_t parents_used = 0;
      for (int t = 0; t < m.nt && q < 4; t++)
        if (!done[t] && pending[t] == 0 && !((parents_used >> m.tpar[t]) & 1u)) { chosen[q++] = t; parents_used |= 1u << m.tpar[t]; }
      if (q == 0) return fail(GMR_EINVAL, "task tree is cyclic");
      for (int i = 0; i < 4; i++)
        m.tround[4 * m.ntround + i] = i < q ? (uint16_t)(chosen[i] | (m.tpar[chosen[i]] << 8)) : (uint16_t)0xffffu;
      for (int i = 0; i < q; i++) { done[chosen[i]] = true; pending[m.tpar[chosen[i]]]--; left--; }
      m.ntround++;
    }
  }
  m.tlmax = 0;
  for (int j = 0; j < m.nh; j++) {
    int n = 0;
    for (int t = 0; t < m.nt; t++)
      if (((m.task_mask[j] >> t) & 1u) && (m.tpar[t] < 0 || !((m.task_mask[j] >> m.tpar[t]) & 1u))) { m.top_mask[j] |= 1u << t; n++; }
    if (n > m.tlmax) m.tlmax = n;
  }
  return GMR_OK;
}
