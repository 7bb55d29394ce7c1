/* kontiki_b200 -- C ABI of the B200-native residual + Jacobian evaluation path of hovren/kontiki.
 *
 * This is the drop-in boundary.  In the reference every measurement is one
 * ceres::DynamicAutoDiffCostFunction whose operator is
 *     bool Residual::operator()(T const* const* params, T* residual)
 *         cpplib/include/kontiki/measurements/gyroscope_measurement.h:58-68
 *         cpplib/include/kontiki/measurements/accelerometer_measurement.h:60-70
 *         cpplib/include/kontiki/measurements/static_rscamera_measurement.h:112-123
 * registered by  *Measurement::AddToEstimator  (gyroscope_measurement.h:75-105, accelerometer_measurement.h:77-108,
 * static_rscamera_measurement.h:130-198) and called by Ceres as
 *     CostFunction::Evaluate(double const* const* parameters, double* residuals, double** jacobians)
 * once per measurement per evaluation from  TrajectoryEstimator::Solve  (cpplib/include/kontiki/trajectory_estimator.h:38-64).
 * The functions below replace that per-measurement operator by ONE batched evaluation per parameter point -- the
 * shape of ceres::EvaluationCallback::PrepareForEvaluation(evaluate_jacobians, new_evaluation_point) -- after which
 * each cost function only copies its slice out.  INTEGRATION.md shows the reference-side binding.
 *
 * Plain C: pointers and sizes only; no exceptions cross the boundary (std::range_error / std::runtime_error of the
 * reference become KTK_ERANGE / KTK_ERUNTIME, text in ktk_last_error()).
 *
 * Layouts (all fp64, quaternions stored x,y,z,w = Eigen coefficient order, as the reference's parameter blocks):
 *   SE3 knot           : 7 doubles [qx qy qz qw tx ty tz]            (uniform_se3_spline_trajectory.h:27,57)
 *   gyro / accel row   : r[3];  J[4][3][7] = d r / d knot(i0+k), k = 0..3, each block row-major 3 x 7 exactly like
 *                        Ceres' jacobians[k] of the 4 knot parameter blocks (locked IMU: the segment has exactly
 *                        those 4 knots, spline_base.h:371-403);  i0 = index of the first active knot
 *   static-RS row      : r[2];  J[114] = [ref window: 4 x (2 x 7)] [obs window: 4 x (2 x 7)] [d r / d rho (2)];
 *                        i0_ref, i0_obs = first active knot of the two spline evaluations.  A knot that is in both
 *                        windows receives the SUM of its two blocks (one parameter block in the reference);
 *                        ktk_expand_static_rs() produces the reference's structural per-block layout.
 * Split trajectory (UniformR3SplineTrajectory + UniformSO3SplineTrajectory, split_trajectory.h): knots are passed as
 *   [R3 knots: n_r3 x 3 | SO3 knots: n_so3 x 4 (x,y,z,w)], the parameter order of SplitEntity (split_trajectory.h:34-39);
 *   gyro row  : J[4 SO3 knots][3][4] (48);          the R3 blocks of the residual are structurally present but zero
 *   accel row : J[4 R3 knots][3][3] (36) | [4 SO3 knots][3][4] (48)
 *   static RS : J[ref R3 4x(2x3)] (24) | [ref SO3 4x(2x4)] (32) | [obs R3] (24) | [obs SO3] (32) | [d r/d rho] (2)
 *   Newton-RS row      : r[2];  J[58 + 14 W] = [ref window: 4 x (2 x 7)] [obs SPAN: W x (2 x 7)] [d r / d rho (2)].  The Newton iteration
 *                        on the row time (newton_rscamera_measurement.h:62-117) evaluates the spline at several times inside
 *                        [t0_obs, t0_obs + readout], so the row carries every knot of the residual's observation span
 *                        {t0_obs - 1e-3, t0_obs + readout + 1e-3} (:210-236): i0_obs = first knot of that span, W = knots of the
 *                        widest span in the group = (ktk_group_row_size - 58) / 14; blocks past a row's own span are zero.
 * Rows are returned in the caller's order of insertion, whatever order the device processes them in.
 */
#ifndef KONTIKI_B200_H_
#define KONTIKI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ktk_problem ktk_problem;

enum {
  KTK_OK = 0,
  KTK_ERANGE = -1,        /* std::range_error in the reference (time outside the trajectory / no segment found) */
  KTK_ERUNTIME = -2,      /* std::runtime_error / std::domain_error in the reference */
  KTK_EINVAL = -3,        /* bad argument at the boundary */
  KTK_ECUDA = -4,         /* CUDA runtime failure, or no usable sm_100 device */
  KTK_EUNSUPPORTED = -5   /* a reference feature outside the built path (see DESIGN.md "out of scope") */
};

enum {
  KTK_EVAL_RESIDUALS = 1u,
  KTK_EVAL_JACOBIANS = 2u,
  KTK_EVAL_ROBUST = 4u,   /* apply ceres::HuberLoss(huber_c) + Corrector to static-RS rows, as Ceres does after Evaluate
                             (static_rscamera_measurement.h:195-197) */
  KTK_EVAL_SENSOR_JACOBIANS = 8u,  /* also fill ktk_group_out.Js: the columns of the sensor's own parameter blocks
                             (sensors/sensors.h:135-165), needed when one of them is unlocked */
  KTK_EVAL_DEVICE_ORDER = 32u,   /* write row k of every output array in DEVICE order (rows sorted by first active knot) instead of the
                             caller's insertion order; ktk_get_row_order() gives the insertion index of device row k.  The values
                             are bit-identical to the default, permuted.  This is the fast layout for device-side consumers (the
                             Gauss-Newton products below take the same flag): a warp's 32 rows are one contiguous block and
                             leave the SM with one TMA bulk store instead of a scatter. */
  KTK_EVAL_LOCAL = 16u    /* knot blocks in LOCAL (tangent) coordinates, i.e. after the knots' ceres::LocalParameterization:
                             SE3 knots 6 columns [upsilon; omega] (LocalParameterizationSE3, uniform_se3_spline_trajectory.h:17-49),
                             SO3 knots 3 (EigenQuaternionParameterization, uniform_so3_spline_trajectory.h:21), R3 knots 3.
                             Rows shrink: IMU 4x3x6 = 72; static RS 4x2x6 + 4x2x6 + 2 = 98; split: gyro 36, accel 36 + 36,
                             static RS 24 + 24 + 24 + 24 + 2 = 98 (ktk_group_row_size_local).  The Gauss-Newton product
                             entry points below expect ambient rows. */
};

enum { KTK_GYROSCOPE = 0, KTK_ACCELEROMETER = 1, KTK_STATIC_RS = 2, KTK_NEWTON_RS = 3, KTK_POSITION = 4, KTK_ORIENTATION = 5, KTK_LIFTING_RS = 6 };
enum { KTK_CAMERA_PINHOLE = 0, KTK_CAMERA_ATAN = 1 };

/* sensors/sensors.h:91-109: relative pose + time offset; *_locked as the reference's lock flags (default locked). */
typedef struct {
  double q_ct[4];            /* x y z w */
  double p_ct[3];
  double time_offset;
  double max_time_offset;    /* sensors.h:107, default 0.1 */
  int32_t q_locked, p_locked, time_offset_locked;
} ktk_sensor;

/* sensors/camera.h:24-28 + sensors/pinhole_camera.h:25 (PinholeCamera) / sensors/atan_camera.h:20-22 (AtanCamera: the same
 * camera matrix plus the distortion centre wc and parameter gamma of the FOV/arctangent model, atan_camera.h:54-103). */
typedef struct {
  ktk_sensor base;
  int32_t rows, cols;
  double readout;
  double K[9];               /* row-major 3x3 */
  int32_t model;             /* KTK_CAMERA_PINHOLE / KTK_CAMERA_ATAN */
  int32_t reserved;
  double wc[2];              /* AtanCamera only */
  double gamma;              /* AtanCamera only, != 0 */
} ktk_camera;
typedef ktk_camera ktk_pinhole_camera;   /* model = KTK_CAMERA_PINHOLE */

/* Per measurement group output pointers (any may be NULL).  Host pointers for ktk_evaluate, device pointers for
 * ktk_evaluate_device.  Device J pointers must be 16-byte aligned (rows leave the SM as TMA bulk stores; cudaMalloc and
 * torch allocations are): KTK_EINVAL otherwise.  Sizes for n rows: r n*3 (IMU, position) / n*2 (camera) / n (orientation); J n*84 / n*114 / n*28
 * (ktk_group_row_size); i0 n; i0_b n (camera: obs). */
typedef struct {
  double* r;
  double* J;
  int32_t* i0;       /* SE3 (or the R3 part of a split trajectory): IMU i0;  camera i0_ref */
  int32_t* i0_b;     /*                                             camera i0_obs; unused for IMU */
  int32_t* i0_c;     /* SO3 part of a split trajectory: IMU i0;  camera i0_ref   (unused for SE3) */
  int32_t* i0_d;     /*                                 camera i0_obs */
  double* Js;        /* KTK_EVAL_SENSOR_JACOBIANS.  IMU rows: d r / d time_offset (3 per row); the relative pose of an IMU is not
                        applied by the reference (TODO.md:6) so those columns are zero, and d r / d bias = -weight I for a
                        ConstantBiasImu (constant_bias_imu.h:52-61).  Camera rows (static and NewtonRs; SE3 and split): 16 per row =
                        d r/d q_ct (2x4) | d r/d p_ct (2x3) | d r/d time_offset (2x1), each block row-major as Ceres'; LiftingRs rows:
                        24 per row = (3x4) | (3x3) | (3x1).  NewtonRs / LiftingRs: relative pose only, the time offset stays locked. */
} ktk_group_out;

const char* ktk_last_error(void);

/* Creates an empty problem on CUDA device `device`.  Fails with KTK_ECUDA when there is no CUDA device: there is no
 * CPU fallback.  device = -1 creates a host-only handle that answers ktk_get_structure / ktk_expand_static_rs
 * (problem-structure bookkeeping) and refuses every evaluation call with KTK_ECUDA. */
int ktk_problem_create(int device, ktk_problem** out);
void ktk_problem_destroy(ktk_problem* p);

/* The stream all device work of this problem is enqueued on (a cudaStream_t; NULL = the legacy default stream). */
int ktk_set_stream(ktk_problem* p, void* cuda_stream);

/* ktk_evaluate_device replays its kernel sequence as ONE CUDA graph when it is called again with the same buffers on a
 * non-default stream (default on; event-timed runs -- ktk_set_profiling -- are never captured). */
int ktk_set_graphs(ktk_problem* p, int32_t on);

/* UniformSE3SplineTrajectory(dt, t0) with n_knots control points (spline_base.h:30-62).  compat_zero_dB = 1 reproduces
 * the reference's Jet-path accelerometer on SE3 (dB left at zero, uniform_se3_spline_trajectory.h:138-141 vs :166-169). */
int ktk_set_se3_spline(ktk_problem* p, double dt, double t0, int32_t n_knots, int32_t compat_zero_dB);

/* SplitTrajectory(UniformR3SplineTrajectory(dt_r3, t0_r3), UniformSO3SplineTrajectory(dt_so3, t0_so3))
 * (split_trajectory.h:87-105; valid time = intersection of the two, :60-66). */
int ktk_set_split_spline(ktk_problem* p, double dt_r3, double t0_r3, int32_t n_r3, double dt_so3, double t0_so3, int32_t n_so3);

/* *Measurement::AddToEstimator, batched: returns the group id (>= 0) or a negative status.  Arrays are copied.
 *   gyroscope / accelerometer: t[n], y[3n], weight[n] (NULL = 1)   (gyroscope_measurement.h:18-20)
 *   static RS: obs_uv[2n], obs_t0[n] (view t0 of the observation), ref_uv[2n], ref_t0[n] (of the landmark's reference
 *   observation), lm_idx[n] into the rho array given to ktk_evaluate, weight[n] (NULL = 1), huber_c[n] (NULL = 5,
 *   static_rscamera_measurement.h:68-69). */
int ktk_add_gyroscope(ktk_problem* p, const ktk_sensor* imu, int64_t n, const double* t, const double* y, const double* weight);
int ktk_add_accelerometer(ktk_problem* p, const ktk_sensor* imu, int64_t n, const double* t, const double* y, const double* weight);
/* PositionMeasurement::AddToEstimator (measurements/position_measurement.h:58-79): r = weight (position[i] - trajectory.Position(t[i])),
 * 3 residuals, no sensor, no loss (the reference has no weight: pass NULL).  Rows as the IMU rows: SE3 J[4][3][7]; split trajectory
 * J[4 R3 knots][3][3] (36; the SO3 blocks of the residual are structurally present but zero), i0 = R3 knot, i0_c = SO3 segment start. */
int ktk_add_position(ktk_problem* p, int64_t n, const double* t, const double* position, const double* weight);
/* OrientationMeasurement::AddToEstimator (measurements/orientation_measurement.h:57-79): ONE residual per row,
 * r = q.angularDistance(trajectory.Orientation(t)) (:27-31; Eigen 3.3: 2 atan2(|vec(d)|, |d.w|), d = q conj(q_hat)), no sensor, no weight,
 * no loss.  q[4n] is (x, y, z, w) per row like every quaternion of this interface (the reference's constructor takes (w, x, y, z),
 * orientation_measurement.h:21-22); it need not be normalised.  Outputs: r[n]; SE3 rows J[4][1][7] (28; KTK_EVAL_LOCAL 24), i0; split
 * trajectory rows J[4 SO3 knots][1][4] (16; local 12), i0 = R3 segment start (structurally present, identically zero blocks), i0_c = SO3 knot.
 * The angle has no derivative at 0 and pi (the reference's Jets produce 0/0 there too): such rows come out NaN in J. */
int ktk_add_orientation(ktk_problem* p, int64_t n, const double* t, const double* q);
int ktk_add_static_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0,
                      const double* ref_uv, const double* ref_t0, const int32_t* lm_idx, const double* weight, const double* huber_c);
/* NewtonRsCameraMeasurement::AddToEstimator (measurements/newton_rscamera_measurement.h:201-262), same arrays as the static
 * measurement.  The row time is found by the reference's 5-step Newton iteration (:62-117) and the Jacobian is the
 * forward-mode derivative THROUGH that iteration, exactly as ceres::Jet produces it.  On a UniformSE3SplineTrajectory the camera's
 * relative pose may be unlocked (KTK_EVAL_SENSOR_JACOBIANS), its time offset not (KTK_EUNSUPPORTED).  KTK_EVAL_LOCAL rows:
 * [ref 4 x (2x6) | obs W x (2x6) | rho 2] (ktk_group_row_size_local). */
int ktk_add_newton_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0,
                      const double* ref_uv, const double* ref_t0, const int32_t* lm_idx, const double* weight, const double* huber_c);
/* LiftingRsCameraMeasurement::AddToEstimator (measurements/lifting_rscamera_measurement.h:151-229), same arrays as the static measurement.
 * The observation is evaluated at the LIFTED time t0_obs + time_offset + vt * readout (:34), vt in [0, 1] a parameter block of the
 * measurement (initially obs_uv.y / rows, :68); 3 residuals weight * [uv - y ; rows * (vt - vt_orig)] (:105-116) under the Huber loss.
 * ktk_set_group_vt feeds the current row times (caller order) like ktk_set_group_sensor feeds sensor parameters.  Packed row
 * [ref 4 x (3x7) | obs W x (3x7) | d r/d vt (3) | d r/d rho (3)] = 90 + 21 W doubles (ktk_group_row_size), W and i0_b as for the Newton
 * rows (whole observation span; the four active knot blocks sit at their place inside it, the others are zero); r is 3 per row.
 * On a UniformSE3SplineTrajectory the relative pose of the camera may be unlocked (Js: 24 per row), its time offset not; KTK_EVAL_LOCAL rows
 * [ref 4 x (3x6) | obs W x (3x6) | vt 3 | rho 3]; no matrix-free products (KTK_EUNSUPPORTED). */
int ktk_add_lifting_rs(ktk_problem* p, const ktk_camera* cam, int64_t n, const double* obs_uv, const double* obs_t0,
                       const double* ref_uv, const double* ref_t0, const int32_t* lm_idx, const double* weight, const double* huber_c);
int ktk_set_group_vt(ktk_problem* p, int32_t group, const double* vt);

/* The sensor parameters are part of the evaluation point when they are unlocked: update them between evaluations.
 * (Changing the time offset re-sorts the group on the next evaluation.)  ktk_set_group_bias: accelerometer / gyroscope bias of
 * a ConstantBiasImu, r = weight (y - (model + bias)). */
int ktk_set_group_sensor(ktk_problem* p, int32_t group, const ktk_sensor* sensor);
int ktk_set_group_bias(ktk_problem* p, int32_t group, const double* bias);

int32_t ktk_num_groups(const ktk_problem* p);
int64_t ktk_group_size(const ktk_problem* p, int32_t group);
int32_t ktk_group_kind(const ktk_problem* p, int32_t group);
int32_t ktk_group_row_size(const ktk_problem* p, int32_t group);   /* doubles per packed Jacobian row (84 / 114 / 48; Newton-RS 58 + 14 W) */
int32_t ktk_group_row_size_local(const ktk_problem* p, int32_t group);   /* ... with KTK_EVAL_LOCAL */
/* NewtonRs / LiftingRs groups: knots of the widest observation span on the current spline(s) -- W (SE3: *w_a, *w_b = 0) or Wa / Wb of the
 * R3 / SO3 splines of a split trajectory, whose rows are packed
 *   [ref R3 4 x (nres x 3) | ref SO3 4 x (nres x 4) | obs R3 Wa x (nres x 3) | obs SO3 Wb x (nres x 4) | (d r/d vt nres) | d r/d rho nres]
 * with i0 / i0_b = first knot of the reference window / of the observation span on the R3 spline and i0_c / i0_d the same on the SO3 spline
 * (newton_rscamera_measurement.h:201-262 and lifting_rscamera_measurement.h:151-229 instantiated on SplitTrajectory; forward mode,
 * csrc/newton_math.cuh).  KTK_EVAL_LOCAL: [ref R3 4 x (nres x 3) | ref SO3 4 x (nres x 3) | obs R3 Wa x (nres x 3) | obs SO3 Wb x (nres x 3) | tail]. */
int ktk_group_span_windows(const ktk_problem* p, int32_t group, int32_t* w_a, int32_t* w_b);
int64_t ktk_num_knot_doubles(const ktk_problem* p);                /* length of the `knots` argument of ktk_evaluate */

/* One batched evaluation at the parameter point (knots[n_knots*7], rho[n_rho]) with HOST buffers: uploads the point,
 * runs the kernels, downloads every non-NULL output of outs[0..num_groups), returns when they are complete.
 * Returns KTK_ERANGE if any measurement falls outside the trajectory (the reference throws at AddToEstimator,
 * trajectory_estimator.h:97-122, or inside Evaluate, spline_base.h:196-201). */
int ktk_evaluate(ktk_problem* p, const double* knots, const double* rho, int64_t n_rho, uint32_t flags, const ktk_group_out* outs);

/* Same with DEVICE buffers, asynchronous on the problem's stream; ktk_synchronize() waits and reports the status. */
int ktk_evaluate_device(ktk_problem* p, const double* d_knots, const double* d_rho, int64_t n_rho, uint32_t flags,
                        const ktk_group_out* d_outs);
int ktk_synchronize(ktk_problem* p);

/* Point queries on the whole trajectory -- trajectory.position(t) / velocity / acceleration / orientation /
 * angular_velocity of the reference's Python API (python/src/kontiki/trajectories/trajectory_helper.h:12-34 ->
 * trajectories/trajectory.h:98-132).  Host buffers; out[16 n] = position(3) | velocity(3) | acceleration(3) |
 * orientation x,y,z,w (4) | angular velocity in the world frame (3); status[n] per time (KTK_ERANGE where the reference
 * throws std::range_error, rows NaN).  Returns the worst status. */
int ktk_traj_evaluate(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status);
/* UniformSE3SplineTrajectory.evaluate(t) (python/src/kontiki/trajectories/py_uniform_se3_spline_trajectory.cc:53-60): out[48 n] =
 * the 4x4 matrices P | P' | P'' of EvaluateSpline (uniform_se3_spline_trajectory.h:101-194), row-major. */
int ktk_se3_evaluate_matrices(ktk_problem* p, const double* knots, int64_t n, const double* t, double* out, int32_t* status);

/* Number of kernel launches ktk_evaluate_device enqueued since the problem was created. */
int64_t ktk_launch_count(const ktk_problem* p);

/* Gauss-Newton contraction, matrix-free, on the packed rows a ktk_evaluate_device call left in device memory
 * (d_outs = the same array of device pointers, with J and every index array non-NULL).  The reference leaves this to
 * Ceres (SPARSE_SCHUR requested at trajectory_estimator.h:40); here the normal equations are applied, not formed.
 * Parameter vectors are AMBIENT and laid out [knots as in ktk_evaluate | rho (n_rho)], length ktk_num_parameters():
 *   ktk_j_apply      : d_u[g][row*nres + r]  = (J v)           for every group g  (nres = 3 IMU, 2 camera)
 *   ktk_jt_apply     : d_y                  += J^T u            (d_y is NOT cleared; fp64 atomics, order not fixed)
 *   ktk_jtj_diagonal : d_y                  += squared entries of the packed blocks, column by column (equals diag(J^T J)
 *                                              except where a camera row's two windows share a knot; the exact diagonal, in
 *                                              local coordinates, is ktk_jtj_diagonal_local)
 * All asynchronous on the problem's stream.  With rows sharded over several GPUs, the sum over ranks of d_y is the
 * full product: one ncclAllReduce of ktk_num_parameters() doubles per product (kontiki_b200/gn.py). */
int64_t ktk_num_parameters(const ktk_problem* p, int64_t n_rho);
int ktk_j_apply(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, const double* d_v, double* const* d_u);
int ktk_jt_apply(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, double* const* d_u, double* d_y);
int ktk_jtj_diagonal(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, double* d_y);
/* diag(P^T J^T J P) in LOCAL (tangent) coordinates, laid out [6 per SE3 knot | rho] or [3 per R3 knot | 3 per SO3 knot |
 * rho]: d_Pa = d Plus/d delta of every SE3 knot (n x 7 x 6, row-major; LocalParameterizationSE3,
 * uniform_se3_spline_trajectory.h:25-48), d_Pb = the same for SO3 knots (n_so3 x 4 x 3; ceres::EigenQuaternionParameterization,
 * uniform_so3_spline_trajectory.h:21); NULL where the trajectory has no such spline. */
int ktk_jtj_diagonal_local(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, const double* d_Pa, const double* d_Pb, double* d_y);
/* (flags: 0, or KTK_EVAL_DEVICE_ORDER if the rows were written with it.) */

/* ---- Gauss-Newton / Levenberg-Marquardt step on the device -----------------------------------------------------
 * What the reference leaves to ceres::Solve with SPARSE_SCHUR (trajectory_estimator.h:38-64), on the rows a ktk_evaluate_device call with
 * KTK_EVAL_DEVICE_ORDER left in device memory (csrc/gn_device.cuh): the landmark blocks (1x1: rho, static_rscamera_measurement.h:178-184)
 * are eliminated, the reduced knot system is solved by block-Jacobi preconditioned conjugate gradients with the Schur complement applied
 * implicitly, Plus() (uniform_se3_spline_trajectory.h:25-48, uniform_so3_spline_trajectory.h:21) and the rho >= 0 bound happen on the device.
 * Every reduction is a gather in a fixed order (no atomics): results are bit-reproducible.  All calls are asynchronous on the problem's
 * stream except ktk_gn_prepare and ktk_gn_pcg_status.  Local (tangent) vectors: 6 per SE3 knot; 3 per R3 and 3 per SO3 knot.
 * With rows sharded over ranks (every landmark's rows on ONE rank, kontiki_b200/sharding.py) the caller all-reduces, between the calls,
 * the buffers named below (ktk_gn_buffer): that is the only exchange.
 *   ktk_gn_prepare(flags, d_outs, n_rho, lm_locked, lock_a, lock_b, huber)   after one evaluation into d_outs; static row lists + vectors.
 *        lm_locked: n_rho bytes or NULL; lock_a / lock_b: SE3 (or R3) / SO3 spline constant; huber[g]: Huber constants of group g in
 *        caller order (NULL entries / NULL: no loss) for the cost 1/2 sum rho(s).
 *   ktk_gn_cost                       scal.cost = 1/2 sum rho(s) of this rank's rows
 *   ktk_gn_linearize_local(knots,rho) tangent bases, "c" = sum J_rho^2, "grho" = J_rho^T r, "blocks_a"/"blocks_b" = diagonal knot blocks
 *   ktk_gn_gradient_local             "z_a"/"z_b" = P^T J_k^T r
 *   ktk_gn_linearize_rhs(radius)      damping of c; "q_a"/"q_b" = reduced gradient P^T J_k^T (r - J_rho C^-1 grho)
 *   ktk_gn_pcg_begin(radius,tol,max)  b = -q, preconditioner (B_kk + D/radius)^-1 with D = clamp(diag B, 1e-6, 1e32) (Ceres' LM scaling)
 *   ktk_gn_product / ktk_gn_pcg_update   "q_a"/"q_b" = S p of this rank's rows (all-reduce), then one CG update; scalars stay on the device
 *   ktk_gn_pcg_status                 synchronises: iterations, convergence flag, |r|/|b|
 *   ktk_gn_finish_local / ktk_gn_finish_mask / ktk_gn_model_local   "drho" (all-reduce after the mask), scal.model_ur / model_uu: the
 *        model decrease is -(model_ur + model_uu / 2), scal.step2 = |delta_knots|^2
 *   ktk_gn_retract(knots_in, rho_in, knots_out, rho_out)
 * scal ("scal" buffer, doubles): [rz, pq, alpha, beta, rnorm2, bnorm2, tol2, (iter, done), (max_iter, pad), model_ur, model_uu, step2, cost]. */
int ktk_gn_prepare(ktk_problem* p, uint32_t flags, const ktk_group_out* d_outs, int64_t n_rho, const uint8_t* lm_locked, int32_t lock_a, int32_t lock_b,
                   const double* const* huber_caller_order);
int ktk_gn_cost(ktk_problem* p);
int ktk_gn_linearize_local(ktk_problem* p, const double* d_knots, const double* d_rho);
int ktk_gn_gradient_local(ktk_problem* p);
int ktk_gn_linearize_rhs(ktk_problem* p, double radius);
int ktk_gn_pcg_begin(ktk_problem* p, double radius, double tol, int32_t max_iter);
int ktk_gn_product(ktk_problem* p);
int ktk_gn_pcg_update(ktk_problem* p);
int ktk_gn_pcg_status(ktk_problem* p, int32_t* iterations, int32_t* done, double* rel_residual);
int ktk_gn_finish_local(ktk_problem* p);
int ktk_gn_finish_mask(ktk_problem* p);
int ktk_gn_model_local(ktk_problem* p);
int ktk_gn_retract(ktk_problem* p, const double* d_knots_in, const double* d_rho_in, double* d_knots_out, double* d_rho_out);
int64_t ktk_gn_buffer(ktk_problem* p, const char* name, double** ptr);

/* order[k] = insertion index (0-based, within the group) of the k-th row in device order.  Fixed once the group has been
 * uploaded (first evaluation, or this call); changes only if the spline grid or the sensor's time offset is changed. */
int ktk_get_row_order(ktk_problem* p, int32_t group, int32_t* order);

/* Kernel timing for bench.py's roofline: while on, every kernel launch of a measurement group is bracketed by CUDA
 * events on the problem's stream; ktk_read_profile synchronises, returns the summed device time (ms) and the number
 * of launches of that group's kernel since the last read, and clears the record. */
int ktk_set_profiling(ktk_problem* p, int32_t on);
int ktk_read_profile(ktk_problem* p, int32_t group, double* total_ms, int64_t* launches);

/* Page-locked host memory for ktk_evaluate's buffers (cudaHostAlloc); optional. */
void* ktk_host_alloc(int64_t bytes);
void ktk_host_free(void* ptr);

/* Structure of one residual block as *Measurement::AddToEstimator would have registered it
 * (spline_base.h:361-404 knot range / segment rule): knot ids in parameter-block order, -1 padded to `cap`.
 * n_ids[i] = number of knot blocks.  Host-side, no device work.  Returns KTK_EINVAL if cap is too small. */
int ktk_get_structure(const ktk_problem* p, int32_t group, int32_t cap, int32_t* knot_ids, int32_t* n_ids);
/* Split trajectory: ktk_get_structure answers for the R3 spline, this one for the SO3 spline (the residual's parameter
 * blocks are all R3 knots, then all SO3 knots, split_trajectory.h:117-123). */
int ktk_get_structure_so3(const ktk_problem* p, int32_t group, int32_t cap, int32_t* knot_ids, int32_t* n_ids);

/* Packed static-RS / Newton-RS rows -> the reference's structural blocks: out[n][cap][2][7] for the knot ids of
 * ktk_get_structure (zero for knots in the segment that are not active).  Host-side helper for Ceres-style consumers. */
int ktk_expand_static_rs(const ktk_problem* p, int32_t group, int32_t cap, const int32_t* knot_ids, const double* J_packed,
                         const int32_t* i0_ref, const int32_t* i0_obs, double* out);

#ifdef __cplusplus
}
#endif
#endif  /* KONTIKI_B200_H_ */
