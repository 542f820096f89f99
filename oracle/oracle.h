/* oracle/oracle.h — C interface of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load it.  See uav_oracle.hpp for what is restated.  The dynamics are pinned bit for bit against
 * the reference's own UavSystem sources compiled by ref_uavsystem.cpp (which exports the stepping
 * subset of this interface under the same names); the collision predicate is pinned against the
 * real vendored nanoflann by ref_nanoflann.cpp. */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_MOTORS 8

typedef struct orc_model_params {
  int32_t n_motors, ground_enabled, takeoff_patch_enabled, reserved_;
  double  g, mass, kf, km, prop_radius, arm_length, body_height, motor_time_constant, max_rpm, min_rpm, air_resistance_coeff, ground_z;
  double  J[9];                                /* row-major */
  double  allocation_matrix[4 * ORC_MAX_MOTORS]; /* row-major 4 x ORC_MAX_MOTORS, scaled */
} orc_model_params;

typedef struct orc_swarm orc_swarm;

void orc_model_params_default(orc_model_params* out); /* MM:26-66 */

orc_swarm* orc_create(int64_t n, int32_t n_types, const orc_model_params* types, const int32_t* type_of_uav, const double* spawn_xyz,
                      const double* spawn_heading);
void       orc_destroy(orc_swarm* s);

/* mode = UavSystem::INPUT_MODE value (US:19-32); payload rows as in include/mrsb.h */
void orc_set_input(orc_swarm* s, int32_t mode, int64_t n, const int32_t* idx, const double* payload, int32_t stride);
/* kind: 0 AccelerationHdgRate, 1 AccelerationHdg, 2 VelocityHdg, 3 VelocityHdgRate (US:254-272); rows [4] */
void orc_set_feedforward(orc_swarm* s, int32_t kind, int64_t n, const int32_t* idx, const double* payload);
void orc_make_step(orc_swarm* s, double dt, int32_t n_steps, int32_t n_threads);
void orc_get_state(orc_swarm* s, int64_t n, const int32_t* idx, double* x, double* v, double* R, double* omega, double* rpm, double* v_prev,
                   double* imu);
void orc_set_state(orc_swarm* s, int64_t n, const int32_t* idx, const double* x, const double* v, const double* R, const double* omega,
                   const double* rpm);
void orc_crash(orc_swarm* s, int64_t n, const int32_t* idx);
void orc_has_crashed(orc_swarm* s, int64_t n, const int32_t* idx, int32_t* out);
void orc_apply_force(orc_swarm* s, int64_t n, const int32_t* idx, const double* f);
void orc_get_force(orc_swarm* s, int64_t n, const int32_t* idx, double* f);
void orc_set_external_moment(orc_swarm* s, int64_t n, const int32_t* idx, const double* m);
void orc_set_params(orc_swarm* s, int64_t n, const int32_t* idx, const orc_model_params* p);
void orc_get_params(orc_swarm* s, int64_t uav, orc_model_params* out);
/* which: 0 mixer(desaturation=v[0]) 1 rate(kp,kd,ki) 2 attitude(kp,kd,ki,max_rp,max_yaw) 3 velocity(kp,kd,ki,max_acc) 4 position(kp,kd,ki,max_vel) */
void orc_set_controller_params(orc_swarm* s, int32_t which, int64_t n, const int32_t* idx, const double* v);
void orc_get_mixer_allocation(orc_swarm* s, int64_t uav, double* out /* [ORC_MAX_MOTORS*4] row-major */);
void orc_get_pid_state(orc_swarm* s, int64_t uav, double* out24); /* 12 x (last_error, integral): pos xyz, vel xyz, att xyz, rate xyz */

/* MultirotorSimulator::handleCollisions (SIM:295-359) on the swarm's current positions.
 * engine 0: cell-list port (this file's own candidate search);
 * engine 1: the real vendored nanoflann, via oracle/_ref/libref_nanoflann.so passed in as a
 *           function pointer by the caller (orc_collide_fn below) — NULL -> engine 0.
 * pairs: directed (i,j) in evaluation order; *count may exceed cap (then truncated). */
typedef int64_t (*orc_collide_fn)(int64_t n, const double* xyz, const double* arm, const double* prop, const double* mass, int32_t crash_mode,
                                  double rebounce, double* forces, uint8_t* crashed, int32_t* pairs, int64_t cap, int32_t n_threads);
void orc_handle_collisions(orc_swarm* s, int32_t enabled, int32_t crash, double rebounce, orc_collide_fn ref_engine, int32_t n_threads, int32_t* pairs,
                           int64_t cap, int64_t* count);

/* stand-alone collision port on a position snapshot (same signature as orc_collide_fn). */
int64_t orc_collide_port(int64_t n, const double* xyz, const double* arm, const double* prop, const double* mass, int32_t crash_mode, double rebounce,
                         double* forces, uint8_t* crashed, int32_t* pairs, int64_t cap, int32_t n_threads);

/* --- the ROS wrapper's arithmetic around the path (src/uav_system_ros.cpp), restated without ROS.
 * mrs_lib::AttitudeConverter (external dependency ctu-mrs/mrs_lib, not vendored, unpinned) is
 * restated from its published behaviour: Matrix3d -> Eigen::Quaterniond(R); getHeading() =
 * atan2 of the xy-projection of the body-x axis; (roll=0, pitch=0, yaw) -> Rz(yaw). */
/* UavSystemRos::timeoutInput (ROSW:474-647): replace the active command by its "hover" version */
void orc_timeout_input(orc_swarm* s, int64_t n, const int32_t* idx);
/* publishOdometry (ROSW:340-368): rows [13] = position xyz, orientation xyzw, body-frame velocity, angular velocity */
void orc_get_odometry(orc_swarm* s, int64_t n, const int32_t* idx, double* out13);
/* publishIMU (ROSW:374-395): rows [10] = angular velocity, linear acceleration, orientation xyzw */
void orc_get_imu(orc_swarm* s, int64_t n, const int32_t* idx, double* out10);
/* publishRangefinder (ROSW:401-420): rows [1] = range */
void orc_get_rangefinder(orc_swarm* s, int64_t n, const int32_t* idx, double* out1);
/* callbackSetMass (ROSW:1028-1054) / callbackSetGroundZ (ROSW:1060-1080) */
void orc_set_mass(orc_swarm* s, int64_t n, const int32_t* idx, const double* mass);
void orc_set_ground_z(orc_swarm* s, int64_t n, const int32_t* idx, const double* z);

/* PIDController::update (CTL/pid.hpp:67-96) on caller-held state {last_error, integral}; returns the output */
double orc_pid_update(double* state2, double kp, double kd, double ki, double saturation, double antiwindup, double error, double dt);

/* counter-based RNG shared by tests and bench: splitmix64(seed + GOLDEN*(stream*2^32 + index)) -> [0,1) */
double orc_u01(uint64_t seed, uint64_t stream, uint64_t index);

#ifdef __cplusplus
}
#endif
#endif
