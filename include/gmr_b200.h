/*
 * gmr_b200.h — C ABI of the B200-native batched retargeting IK (libgmr_b200.so) and of
 * its CPU float64 checker (oracle/liboracle.so, gmr_oracle_* symbols).
 *
 * The reference has no FFI layer: its boundary for this path is the Python class
 * `GeneralMotionRetargeting` (reference general_motion_retargeting/__init__.py:3,
 * general_motion_retargeting/motion_retarget.py:10-185).  The entry points below are what
 * a binding for that class needs; each cites the reference interface it replaces.
 * Plain pointers and sizes only; caller owns every buffer; every function returns 0 or a
 * negative GMR_E* code and never throws or aborts; gmr_last_error() gives the text.
 */
#ifndef GMR_B200_H
#define GMR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMR_OK            0
#define GMR_EINVAL       -1   /* bad argument / unsupported model               */
#define GMR_ECUDA        -2   /* CUDA runtime error (text in gmr_last_error())  */
#define GMR_ENOMEM       -3
#define GMR_ELIMIT       -4   /* model exceeds a compiled-in size limit          */

#define GMR_MAX_BODY     64
#define GMR_MAX_HINGE    32   /* one warp lane per hinge DoF                      */
#define GMR_MAX_HUMAN    32
#define GMR_MAX_TASK     32

/* per-clip status bits (GmrBatchExtra.status / status_out of the host entry).  The reference fails per clip with a
 * Python exception (mink TargetNotSet / `assert dq is not None` in motion_retarget.py:147-150, caught per file by
 * scripts/smplx_to_robot_dataset.py:62-76,96-100 which then skips the file); the batched kernel records the same
 * events per clip instead.  status = bits | (frame of the first event << 8). */
#define GMR_STATUS_BAD_INPUT   1u  /* non-finite keypoint or zero-norm quaternion: the clip stops at that frame, later
                                      frames of the clip are NOT written                                            */
#define GMR_STATUS_AS_CAP      2u  /* the active-set loop of some solve hit its iteration cap: the step taken is
                                      feasible but was not proven optimal; the clip continues                       */
#define GMR_STATUS_NONFINITE   4u  /* non-finite configuration after a step (the QP "has no solution"): clip stops   */
#define GMR_STATUS_FATAL       (GMR_STATUS_BAD_INPUT | GMR_STATUS_NONFINITE)

/* flags for gmr_retarget_batch */
#define GMR_FLAG_OFFSET_TO_GROUND  1u  /* retarget(..., offset_to_ground=True), motion_retarget.py:122-123,252-270 */
#define GMR_FLAG_COMPUTE_F64       4u  /* gmr_retarget_batch / _host: float32 buffers, float64 arithmetic in the kernel */
#define GMR_FLAG_NO_SOLVE          2u  /* update_targets()/error1()/error2() only: targets_out and err_out at the given
                                         configuration, qpos_out = qpos_init, no IK step (motion_retarget.py:117-136,188-200) */

/*
 * Flat description of one (source format, robot) pair: what the reference constructor
 * builds from the MJCF (`mj.MjModel.from_xml_path`, motion_retarget.py:27), the IK JSON
 * (:30-59) and `setup_retarget_configuration` (:74-114), plus the solver knobs the
 * reference passes to mink (`solver="daqp"`, `damping=0.5` :18-19, `lm_damping=1` :88,
 * `max_iter=10` :56, threshold 0.001 :153, mink ConfigurationLimit gain 0.95,
 * `model.opt.timestep` :146).  All arrays are host memory, copied by gmr_model_create.
 * Quaternions are wxyz.  Bodies are in MuJoCo order without the world body (body 0 = the
 * floating root); qpos = [x y z qw qx qy qz, hinge...].
 */
typedef struct GmrModelDesc {
  int32_t nbody, nhinge, nhuman, ntask;
  /* kinematic tree */
  const int32_t* body_parent;    /* [nbody]  -1 for the root                    */
  const double*  body_pos;       /* [nbody*3] offset in the parent frame        */
  const double*  body_quat;      /* [nbody*4]                                   */
  const int32_t* body_hinge;     /* [nbody]  hinge index owned by body, or -1   */
  const double*  hinge_axis;     /* [nhinge*3] body-local, unit                 */
  const double*  hinge_lo;       /* [nhinge]                                    */
  const double*  hinge_hi;       /* [nhinge]                                    */
  const uint8_t* hinge_limited;  /* [nhinge]                                    */
  const double*  qpos0;          /* [7+nhinge]                                  */
  /* target preprocessing (scale_human_data :209-232, offset_human_data :234-250) */
  int32_t        human_root;     /* index of human_root_name                    */
  const double*  human_scale;    /* [nhuman] human_scale_table, before the height ratio */
  const double*  human_pos_off;  /* [nhuman*3] table-1 pos_offset - ground_height*z (:91) */
  const double*  human_rot_off;  /* [nhuman*4] table-1 rot_offset, unit (:92)   */
  const uint8_t* human_foot;     /* [nhuman] name contains "Foot"/"foot" (:260) */
  /* frame tasks (:80-107): union of table-1 and table-2 entries with a non-zero weight */
  const int32_t* task_body;      /* [ntask] robot body index (frame_name)       */
  const int32_t* task_human;     /* [ntask] human body index                    */
  const double*  task_w1;        /* [ntask*2] position_cost, orientation_cost in stage 1 */
  const double*  task_w2;        /* [ntask*2] same, stage 2                     */
  const uint8_t* task_in1;       /* [ntask] member of tasks1                    */
  const uint8_t* task_in2;       /* [ntask] member of tasks2                    */
  int32_t        use_stage1, use_stage2;   /* use_ik_match_table1/2 (:51-52)    */
  /* solver knobs */
  double  damping;               /* Tikhonov damping passed to mink.solve_ik    */
  double  lm_damping;            /* FrameTask lm_damping                        */
  double  limit_gain;            /* ConfigurationLimit gain                     */
  double  tol;                   /* loop threshold on the error decrease        */
  double  timestep;              /* model.opt.timestep                          */
  int32_t max_iter;              /* conditional iterations per stage            */
  double  lie_eps;               /* mink.lie.utils.get_epsilon(float64): threshold of the small-angle branches of
                                    SO3.log / SE3.log / jlog (Taylor series, the jlog = I shortcut).  mink is not
                                    vendored in the reference; upstream is recalled as 1e-10, round 1 assumed
                                    2.2e-15 (10 eps).  0 selects 1e-10.                                          */
} GmrModelDesc;

typedef struct GmrModel GmrModel;  /* opaque; owns device copies of the tables; immutable after create */

/* ---- libgmr_b200.so (CUDA, sm_100a) ----------------------------------------------------- */

/* replaces GeneralMotionRetargeting.__init__ + setup_retarget_configuration
 * (motion_retarget.py:13-114): validate the description and stage it on `device`. */
int gmr_model_create(const GmrModelDesc* desc, int device, GmrModel** out);
int gmr_model_destroy(GmrModel* model);

/* replaces `for frame in frames: qpos = retargeter.retarget(frame)` over a batch of clips
 * (callers: scripts/smplx_to_robot_dataset.py:84-87, scripts/bvh_to_robot_dataset.py:96-103;
 * callee: motion_retarget.py:139-185).  All pointers are DEVICE pointers on the model's
 * device.  Clip c is solved as a fresh retargeter (qpos starts at qpos_init[c] if given,
 * else qpos0) whose frames t = 0..T-1 are solved in order, each warm-started from the last.
 *   pos   [C,T,nhuman,3] float32 metres, world Z-up      quat [C,T,nhuman,4] float32 wxyz
 *   ratio [C] float32 actual_human_height / human_height_assumption (NULL: 1.0)  (:36-43)
 *   qpos_init [C,nq] or NULL            qpos_out [C,T,nq] float32 (required)
 *   iters_out [C,T,2] int32 solves per stage, or NULL
 *   err_out   [C,T,2] float32 final error1()/error2() (:188-200), or NULL
 *   targets_out [C,T,nhuman,7] float32 = scaled_human_data (pos, quat wxyz) (:124), or NULL
 * Stream-ordered, asynchronous, no hidden synchronisation. */
int gmr_retarget_batch(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                       int32_t C, int32_t T, const float* qpos_init, float* qpos_out,
                       int32_t* iters_out, float* err_out, float* targets_out,
                       uint32_t flags, void* cuda_stream);

/* Optional extras of one batch (every member may be NULL): ragged clip lengths and the post-solve forward-
 * kinematics epilogue of the dataset scripts (scripts/smplx_to_robot_dataset.py:93-123,
 * scripts/bvh_to_robot_dataset.py:107-143), fused into the solve kernel. Device pointers. */
typedef struct GmrBatchExtra {
  const int32_t* lengths;     /* [C] frames of clip c (<= T); later frames are neither read nor written      */
  float* local_body_pos;      /* [C,T,nbody,3] KinematicsModel.forward_kinematics(0, identity, dof_pos)[0]   */
  float* lowest_z;            /* [C] min over frames and bodies of the world z (torch.min(body_pos[..., 2])) */
  uint32_t* warm_state;       /* [C,4] in/out: the solver's working sets (joints resting on a limit) carried from
                                 one call to the next when a clip is fed in pieces; zeros to start            */
  int32_t* status;            /* [C] out: GMR_STATUS_* bits | first event's frame << 8; 0 = clean.  The caller zeroes
                                 it (bits are OR-ed in, so one array can collect several calls)               */
} GmrBatchExtra;

/* gmr_retarget_batch with extras; flags may carry GMR_FLAG_COMPUTE_F64. */
int gmr_retarget_batch_ex(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                          int32_t C, int32_t T, const float* qpos_init, float* qpos_out,
                          int32_t* iters_out, float* err_out, float* targets_out,
                          const GmrBatchExtra* extra, uint32_t flags, void* cuda_stream);

/* Mixed-robot batches (BASELINE.json configs[4]: clips of several robots retargeted together).  A launch is
 * robot-uniform per CTA, so the caller groups the clips by robot ("buckets", one model handle each) and hands all
 * buckets to ONE call: the device's SMs are divided among the buckets by work and every bucket's slow clips overlap
 * with the other buckets' bulk (separate gmr_retarget_batch calls would serialise their tails).  float32 buffers,
 * flags: GMR_FLAG_COMPUTE_F64, GMR_FLAG_OFFSET_TO_GROUND.  At most 8 buckets, all on one device. */
typedef struct GmrBatchDesc {
  GmrModel* model;
  const float* pos; const float* quat; const float* ratio;   /* as gmr_retarget_batch */
  int32_t C, T;
  const float* qpos_init;                                     /* or NULL */
  float* qpos_out; int32_t* iters_out; float* err_out;        /* iters_out / err_out may be NULL */
} GmrBatchDesc;
int gmr_retarget_multi(const GmrBatchDesc* batches, int32_t n, uint32_t flags, void* cuda_stream);

/* The rest of the scripts' epilogue as one elementwise pass (replaces :97-131 / :112-143): splits qpos into
 * the arrays the motion pkl holds (consumers: data_loader.py:4-16, booster_gym/utils/motion_loader.py:42-100),
 *   root_pos [C,T,3], root_rot [C,T,4] xyzw (:103-104), dof_pos [C,T,nhinge]            (all float32, device)
 * with, per clip, root_pos.z -= lowest_z[c] if height_adjust (:118-123) and root_pos.xy -= root_pos.xy of the
 * clip's first frame if origin_offset (:125-128).  lengths/lowest_z may be NULL (no raggedness / no adjust). */
int gmr_finalize_motion(GmrModel* model, const float* qpos, const float* lowest_z, const int32_t* lengths,
                        int32_t C, int32_t T, int32_t height_adjust, int32_t origin_offset,
                        float* root_pos_out, float* root_rot_xyzw_out, float* dof_pos_out, void* cuda_stream);

/* ---- single live stream: `retargeter.retarget(frame)` once per incoming frame ------------------------------
 * (callers: scripts/smplx_to_robot.py:125, scripts/bvh_to_robot.py:118, scripts/optitrack_to_robot.py:40).
 * A stream owns what mink.Configuration owns in the reference (motion_retarget.py:75): the current qpos, which
 * warm-starts the next frame, plus pinned staging buffers, a CUDA stream and a captured CUDA graph
 * (copy in -> solve -> copy out), so one call costs one graph launch.  float64 arithmetic.  HOST pointers. */
typedef struct GmrStream GmrStream;
int gmr_stream_create(GmrModel* model, double height_ratio, GmrStream** out);
int gmr_stream_destroy(GmrStream* stream);
/* set the configuration (NULL: the model's qpos0), like constructing a fresh retargeter */
int gmr_stream_reset(GmrStream* stream, const double* qpos);
/* pos [nhuman,3], quat [nhuman,4] float32 -> qpos_out [nq] float64 (required); iters_out [2], err_out [2],
 * targets_out [nhuman,7] may be NULL.  flags: GMR_FLAG_OFFSET_TO_GROUND, GMR_FLAG_NO_SOLVE (update_targets /
 * error1 / error2: the configuration is left untouched).  Returns when qpos_out is valid. */
int gmr_stream_retarget(GmrStream* stream, const float* pos, const float* quat, uint32_t flags,
                        double* qpos_out, int32_t* iters_out, double* err_out, double* targets_out);

/* ---- human-frame producers: the loaders' arithmetic after file parsing, for a batch of frames ---------------
 * Output = the solver's input layout (pos [F,nh,3] metres Z-up, quat [F,nh,4] wxyz), device pointers; the small
 * index arrays (parents, *_joint) are HOST pointers.  Frames of many clips can be concatenated (frames are
 * independent) except across a resampling call.
 *
 * BVH / LAFAN1 (replaces utils/lafan1.py:17-35 after read_bvh: quat_fk of lafan_vendor/utils.py:88-103, the
 * Y-up -> Z-up rotation, cm -> m, and the LeftFootMod / RightFootMod bodies):
 *   lrot [F,J,4] local quaternions wxyz, lpos [F,J,3] local positions in cm (Anim.quats / Anim.pos),
 *   body b takes its position from joint pos_joint[b] and its orientation from joint rot_joint[b]. */
int gmr_produce_bvh_frames(const float* lrot, const float* lpos, const int32_t* parents, int32_t F, int32_t J,
                           const int32_t* pos_joint, const int32_t* rot_joint, int32_t nh,
                           float* pos_out, float* quat_out, void* cuda_stream);
/* SMPL-X (replaces utils/smpl.py:127-196 after the body model's forward pass): global_orient [F,3] and
 * full_pose [F,NJ,3] axis-angle, joints [F,NJ_joints,3]; F_out < F resamples like the reference
 * (target times linspace(0, F-1, F_out), SLERP of neighbouring rotations, linear positions), F_out == F copies. */
int gmr_produce_smplx_frames(const float* global_orient, const float* full_pose, const float* joints, const int32_t* parents,
                             int32_t F, int32_t NJ, int32_t NJ_joints, int32_t F_out, const int32_t* body_joint, int32_t nh,
                             float* pos_out, float* quat_out, void* cuda_stream);

/* float64 variant of the same kernel (identical semantics; qpos_init/qpos_out/err_out/
 * targets_out are double).  Exists so that parity can be checked without float32 rounding. */
int gmr_retarget_batch_f64(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                           int32_t C, int32_t T, const double* qpos_init, double* qpos_out,
                           int32_t* iters_out, double* err_out, double* targets_out,
                           uint32_t flags, void* cuda_stream);

/* float64-buffer variant with extras (GmrBatchExtra: ragged lengths, FK epilogue, working sets, per-clip status) */
int gmr_retarget_batch_f64_ex(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                              int32_t C, int32_t T, const double* qpos_init, double* qpos_out,
                              int32_t* iters_out, double* err_out, double* targets_out,
                              const GmrBatchExtra* extra, uint32_t flags, void* cuda_stream);

/* Same call with HOST buffers; returns when every output array is complete.  This is what a dataset script calls.
 * Page-locked (pinned) arrays are used in place: the kernel pulls each frame's keypoints over the host link one frame
 * ahead of the solve (TMA bulk copies from the mapped arrays) and writes qpos straight into qpos_out, so the transfer
 * rides inside the solve and the batch stays one schedule.  Pageable arrays are staged through device scratch with
 * cudaMemcpyAsync before / after the solve (serial, and in chunks of whole waves only beyond an 8 GB staging budget):
 * pin the large arrays (pos, quat, qpos_out) for end-to-end throughput.  Each array is treated on its own. */
int gmr_retarget_batch_host(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                            int32_t C, int32_t T, const float* qpos_init, float* qpos_out,
                            int32_t* iters_out, float* err_out, uint32_t flags);

/* ... with per-clip status words (GMR_STATUS_*; host pointer [C], written by the call, may be NULL) */
int gmr_retarget_batch_host_ex(GmrModel* model, const float* pos, const float* quat, const float* ratio,
                               int32_t C, int32_t T, const float* qpos_init, float* qpos_out,
                               int32_t* iters_out, float* err_out, int32_t* status_out, uint32_t flags);

/* number of kernel launches issued by this library since load (for bench bookkeeping) */
int64_t gmr_launch_count(void);
/* profiling aid: while a device buffer of [2C,4] int64 is registered (NULL to stop), every clip of a single-robot launch
 * records {start ns, end ns, SM | warp << 16, IK steps | factorisations << 32}; rows [0,C) by launches that start at
 * frame 0, rows [C,2C) by launches that continue a clip (tools/prof/timeline.py) */
void gmr_debug_trace(long long* device_buffer);
const char* gmr_last_error(void);
/* static properties of the solve kernel for (model, C): for reports */
int gmr_kernel_info(GmrModel* model, int32_t precision_bits, int32_t* threads_per_cta,
                    int32_t* clips_per_cta, int32_t* smem_bytes, int32_t* regs_per_thread,
                    int32_t* ctas_per_sm);

/* ---- oracle/liboracle.so (CPU checker; test infrastructure only) ----------------------- */

/* oracle-only flag: evaluate mink's SO3/SE3 Jacobian coefficients with cancellation-free forms */
#define GMR_ORACLE_FLAG_STABLE_LIE 0x10000u

/* Same semantics computed on the host in float64 (precision_bits = 64) or with every
 * arithmetic step in float32 (32, used to study rounding).  Host pointers.  nthreads <= 0
 * uses every hardware thread. */
int gmr_oracle_retarget_batch(const GmrModelDesc* desc, const float* pos, const float* quat,
                              const float* ratio, int32_t C, int32_t T, const double* qpos_init,
                              double* qpos_out, int32_t* iters_out, double* err_out,
                              uint32_t flags, int32_t nthreads, int32_t precision_bits);

#ifdef __cplusplus
}
#endif
#endif /* GMR_B200_H */
