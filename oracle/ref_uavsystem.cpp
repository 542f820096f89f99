// oracle/ref_uavsystem.cpp -> oracle/_ref/libref_uavsystem.so.  TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN UavSystem (uav_system.hpp, multirotor_model.hpp, controllers/*.hpp), compiled
// unmodified from where it lies under $(REFERENCE)/include — nothing is copied into this
// repository.  The image has neither Eigen nor Boost, so the two libraries the reference delegates
// leaf arithmetic to are replaced at compile time by the stand-ins under oracle/shim/ (see the
// headers there for the evaluation rules they follow); everything else — makeStep's cascade
// dispatch (US:304-380), the six controllers and the PID (CTL/*.hpp), the mixer and its
// desaturation (CTL/mixer.hpp:120-147), MultirotorModel::step / operator() (MM:220-366) with the
// ground and take-off patches and both NaN guards — is the reference's code.
//
// It exports the stepping subset of oracle/oracle.h under the same names, so oracle/binding.py can
// drive it and the restated oracle with one harness (tests/test_ref_uavsystem.py).  UavSystem keeps
// the model and the PIDs private; the three accessors the oracle's C interface offers beyond the
// public API (set_state, set_external_moment, get_pid_state) reach them through `#define private
// public`, which does not change any code on the stepping path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <optional>
#include <thread>
#include <vector>

#include <eigen3/Eigen/Core>

#include "oracle.h"

#define private public
#include <mrs_multirotor_simulator/uav_system/uav_system.hpp>
#undef private

using namespace mrs_multirotor_simulator;

struct orc_swarm {
  std::vector<UavSystem> uavs;
};

namespace {

MultirotorModel::ModelParams fromC(const orc_model_params& c) {
  MultirotorModel::ModelParams p;
  p.n_motors              = c.n_motors;
  p.g                     = c.g;
  p.mass                  = c.mass;
  p.kf                    = c.kf;
  p.km                    = c.km;
  p.prop_radius           = c.prop_radius;
  p.arm_length            = c.arm_length;
  p.body_height           = c.body_height;
  p.motor_time_constant   = c.motor_time_constant;
  p.max_rpm               = c.max_rpm;
  p.min_rpm               = c.min_rpm;
  p.air_resistance_coeff  = c.air_resistance_coeff;
  p.ground_enabled        = c.ground_enabled != 0;
  p.ground_z              = c.ground_z;
  p.takeoff_patch_enabled = c.takeoff_patch_enabled != 0;
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) p.J(r, k) = c.J[3 * r + k];
  p.allocation_matrix = Eigen::MatrixXd::Zero(4, c.n_motors);
  for (int r = 0; r < 4; r++)
    for (int m = 0; m < c.n_motors; m++) p.allocation_matrix(r, m) = c.allocation_matrix[r * ORC_MAX_MOTORS + m];
  return p;
}

void toC(const MultirotorModel::ModelParams& p, orc_model_params* c) {
  std::memset(c, 0, sizeof(*c));
  c->n_motors              = p.n_motors;
  c->g                     = p.g;
  c->mass                  = p.mass;
  c->kf                    = p.kf;
  c->km                    = p.km;
  c->prop_radius           = p.prop_radius;
  c->arm_length            = p.arm_length;
  c->body_height           = p.body_height;
  c->motor_time_constant   = p.motor_time_constant;
  c->max_rpm               = p.max_rpm;
  c->min_rpm               = p.min_rpm;
  c->air_resistance_coeff  = p.air_resistance_coeff;
  c->ground_enabled        = p.ground_enabled;
  c->ground_z              = p.ground_z;
  c->takeoff_patch_enabled = p.takeoff_patch_enabled;
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) c->J[3 * r + k] = p.J(r, k);
  for (int r = 0; r < 4; r++)
    for (int m = 0; m < p.n_motors; m++) c->allocation_matrix[r * ORC_MAX_MOTORS + m] = p.allocation_matrix(r, m);
}

inline int64_t at(const int32_t* idx, int64_t k) {
  return idx ? idx[k] : k;
}

inline Eigen::Vector3d vec3(const double* p) {
  return Eigen::Vector3d(p[0], p[1], p[2]);
}

}  // namespace

extern "C" {

// marks this library as the compiled reference (binding.py checks it before trusting the handle)
int32_t ref_uavsystem_flavour(void) {
#ifdef EIGSHIM_VECTORISED_REDUX
  return 2;
#else
  return 1;
#endif
}

void orc_model_params_default(orc_model_params* out) {
  toC(MultirotorModel::ModelParams(), out);
}

orc_swarm* orc_create(int64_t n, int32_t n_types, const orc_model_params* types, const int32_t* type_of_uav, const double* spawn_xyz,
                      const double* spawn_heading) {
  std::vector<MultirotorModel::ModelParams> tp;
  for (int t = 0; t < n_types; t++) tp.push_back(fromC(types[t]));
  orc_swarm* s = new orc_swarm();
  s->uavs.reserve(n);
  for (int64_t i = 0; i < n; i++) {
    const int             t   = type_of_uav ? type_of_uav[i] : 0;
    const Eigen::Vector3d pos = spawn_xyz ? vec3(spawn_xyz + 3 * i) : Eigen::Vector3d(0, 0, 0);
    s->uavs.emplace_back(tp[t], pos, spawn_heading ? spawn_heading[i] : 0.0);
  }
  return s;
}

void orc_destroy(orc_swarm* s) {
  delete s;
}

void orc_set_input(orc_swarm* s, int32_t mode, int64_t n, const int32_t* idx, const double* payload, int32_t stride) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem&    u = s->uavs[at(idx, k)];
    const double* p = payload ? payload + k * stride : nullptr;
    switch (mode) {
      case UavSystem::ACTUATOR_CMD: {
        reference::Actuators c;
        const int            nm = u.getParams().n_motors;
        c.motors                = Eigen::VectorXd::Zero(nm);
        for (int m = 0; m < nm; m++) c.motors(m) = m < stride ? p[m] : 0.0;
        u.setInput(c);
      } break;
      case UavSystem::CONTROL_GROUP_CMD: {
        reference::ControlGroup c;
        c.roll = p[0], c.pitch = p[1], c.yaw = p[2], c.throttle = p[3];
        u.setInput(c);
      } break;
      case UavSystem::ATTITUDE_RATE_CMD: {
        reference::AttitudeRate c;
        c.rate_x = p[0], c.rate_y = p[1], c.rate_z = p[2], c.throttle = p[3];
        u.setInput(c);
      } break;
      case UavSystem::ATTITUDE_CMD: {
        reference::Attitude c;
        for (int col = 0; col < 3; col++)
          for (int r = 0; r < 3; r++) c.orientation(r, col) = p[3 * col + r];
        c.throttle = p[9];
        u.setInput(c);
      } break;
      case UavSystem::TILT_HDG_RATE_CMD: {
        reference::TiltHdgRate c;
        c.tilt_vector  = vec3(p);
        c.heading_rate = p[3];
        c.throttle     = p[4];
        u.setInput(c);
      } break;
      case UavSystem::ACCELERATION_HDG_RATE_CMD:
        u.setInput(reference::AccelerationHdgRate(vec3(p), p[3]));
        break;
      case UavSystem::ACCELERATION_HDG_CMD:
        u.setInput(reference::AccelerationHdg(vec3(p), p[3]));
        break;
      case UavSystem::VELOCITY_HDG_RATE_CMD:
        u.setInput(reference::VelocityHdgRate(vec3(p), p[3]));
        break;
      case UavSystem::VELOCITY_HDG_CMD:
        u.setInput(reference::VelocityHdg(vec3(p), p[3]));
        break;
      case UavSystem::POSITION_CMD: {
        reference::Position c;
        c.position = vec3(p);
        c.heading  = p[3];
        u.setInput(c);
      } break;
      default:
        u.setInput();
        break;
    }
  }
}

void orc_set_feedforward(orc_swarm* s, int32_t kind, int64_t n, const int32_t* idx, const double* payload) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem&    u = s->uavs[at(idx, k)];
    const double* p = payload + 4 * k;
    switch (kind) {
      case 0:
        u.setFeedforward(reference::AccelerationHdgRate(vec3(p), p[3]));
        break;
      case 1:
        u.setFeedforward(reference::AccelerationHdg(vec3(p), p[3]));
        break;
      case 2:
        u.setFeedforward(reference::VelocityHdg(vec3(p), p[3]));
        break;
      case 3:
        u.setFeedforward(reference::VelocityHdgRate(vec3(p), p[3]));
        break;
    }
  }
}

void orc_make_step(orc_swarm* s, double dt, int32_t n_steps, int32_t n_threads) {
  const int64_t n = int64_t(s->uavs.size());
  auto          run = [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; i++)
      for (int k = 0; k < n_steps; k++) s->uavs[i].makeStep(dt);
  };
  if (n_threads <= 1 || n < 2) {
    run(0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t            chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    th.emplace_back(run, b, e);
  }
  for (auto& t : th) t.join();
}

void orc_get_state(orc_swarm* s, int64_t n, const int32_t* idx, double* x, double* v, double* R, double* omega, double* rpm, double* v_prev,
                   double* imu) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem&                   u   = s->uavs[at(idx, k)];
    const MultirotorModel::State st  = u.getState();
    const Eigen::Vector3d        acc = u.getImuAcceleration();
    const int                    nm  = u.getParams().n_motors;
    for (int c = 0; c < 3; c++) {
      if (x) x[3 * k + c] = st.x(c);
      if (v) v[3 * k + c] = st.v(c);
      if (omega) omega[3 * k + c] = st.omega(c);
      if (v_prev) v_prev[3 * k + c] = st.v_prev(c);
      if (imu) imu[3 * k + c] = acc(c);
      if (R)
        for (int r = 0; r < 3; r++) R[9 * k + 3 * c + r] = st.R(r, c);
    }
    if (rpm)
      for (int m = 0; m < ORC_MAX_MOTORS; m++) rpm[ORC_MAX_MOTORS * k + m] = m < nm ? st.motor_rpm(m) : 0.0;
  }
}

void orc_set_state(orc_swarm* s, int64_t n, const int32_t* idx, const double* x, const double* v, const double* R, const double* omega,
                   const double* rpm) {  // MultirotorModel::setState, MM:424-433
  for (int64_t k = 0; k < n; k++) {
    UavSystem&             u  = s->uavs[at(idx, k)];
    MultirotorModel::State st = u.multirotor_model_.getState();
    const int              nm = u.getParams().n_motors;
    for (int c = 0; c < 3; c++) {
      if (x) st.x(c) = x[3 * k + c];
      if (v) st.v(c) = v[3 * k + c];
      if (omega) st.omega(c) = omega[3 * k + c];
      if (R)
        for (int r = 0; r < 3; r++) st.R(r, c) = R[9 * k + 3 * c + r];
    }
    if (rpm)
      for (int m = 0; m < nm; m++) st.motor_rpm(m) = rpm[ORC_MAX_MOTORS * k + m];
    u.multirotor_model_.setState(st);
  }
}

void orc_crash(orc_swarm* s, int64_t n, const int32_t* idx) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].crash();
}
void orc_has_crashed(orc_swarm* s, int64_t n, const int32_t* idx, int32_t* out) {
  for (int64_t k = 0; k < n; k++) out[k] = s->uavs[at(idx, k)].hasCrashed();
}
void orc_apply_force(orc_swarm* s, int64_t n, const int32_t* idx, const double* f) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].applyForce(vec3(f + 3 * k));
}
void orc_get_force(orc_swarm* s, int64_t n, const int32_t* idx, double* f) {
  for (int64_t k = 0; k < n; k++) {
    const Eigen::Vector3d& e = s->uavs[at(idx, k)].multirotor_model_.getExternalForce();
    for (int c = 0; c < 3; c++) f[3 * k + c] = e(c);
  }
}
void orc_set_external_moment(orc_swarm* s, int64_t n, const int32_t* idx, const double* m) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].multirotor_model_.setExternalMoment(vec3(m + 3 * k));
}
void orc_set_params(orc_swarm* s, int64_t n, const int32_t* idx, const orc_model_params* p) {
  const MultirotorModel::ModelParams mp = fromC(*p);
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].setParams(mp);
}
void orc_get_params(orc_swarm* s, int64_t uav, orc_model_params* out) {
  toC(s->uavs[uav].getParams(), out);
}
void orc_set_controller_params(orc_swarm* s, int32_t which, int64_t n, const int32_t* idx, const double* v) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem& u = s->uavs[at(idx, k)];
    switch (which) {
      case 0: {
        Mixer::Params p;
        p.desaturation = v[0] != 0.0;
        u.setMixerParams(p);
      } break;
      case 1: {
        RateController::Params p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2];
        u.setRateControllerParams(p);
      } break;
      case 2: {
        AttitudeController::Params p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_rate_roll_pitch = v[3], p.max_rate_yaw = v[4];
        u.setAttitudeControllerParams(p);
      } break;
      case 3: {
        VelocityController::Params p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_acceleration = v[3];
        u.setVelocityControllerParams(p);
      } break;
      case 4: {
        PositionController::Params p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_velocity = v[3];
        u.setPositionControllerParams(p);
      } break;
    }
  }
}
void orc_get_mixer_allocation(orc_swarm* s, int64_t uav, double* out) {
  UavSystem&            u  = s->uavs[uav];
  const Eigen::MatrixXd a  = u.getMixerAllocation();
  const int             nm = u.getParams().n_motors;
  for (int m = 0; m < ORC_MAX_MOTORS; m++)
    for (int c = 0; c < 4; c++) out[4 * m + c] = m < nm ? a(m, c) : 0.0;
}
void orc_get_pid_state(orc_swarm* s, int64_t uav, double* o) {
  UavSystem&           u     = s->uavs[uav];
  const PIDController* p[12] = {&u.position_controller_.pid_x_, &u.position_controller_.pid_y_, &u.position_controller_.pid_z_,
                                &u.velocity_controller_.pid_x_, &u.velocity_controller_.pid_y_, &u.velocity_controller_.pid_z_,
                                &u.attitude_controller_.pid_x_, &u.attitude_controller_.pid_y_, &u.attitude_controller_.pid_z_,
                                &u.rate_controller_.pid_x_,     &u.rate_controller_.pid_y_,     &u.rate_controller_.pid_z_};
  for (int k = 0; k < 12; k++) {
    o[2 * k]     = p[k]->last_error_;
    o[2 * k + 1] = p[k]->integral_;
  }
}

double orc_pid_update(double* state2, double kp, double kd, double ki, double saturation, double antiwindup, double error, double dt) {
  PIDController pid;
  pid.setParams(kp, kd, ki, saturation, antiwindup);
  pid.last_error_ = state2[0];
  pid.integral_   = state2[1];
  const double r  = pid.update(error, dt);
  state2[0]       = pid.last_error_;
  state2[1]       = pid.integral_;
  return r;
}

}  // extern "C"
