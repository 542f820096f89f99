// oracle/oracle_capi.cpp — C interface + swarm runner + collision-loop port for the CPU oracle.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates the loop order of the reference node
// (src/multirotor_simulator.cpp:198-231: step every UAV, then handleCollisions) and the
// collision loop itself (src/multirotor_simulator.cpp:295-359).
#include "oracle.h"
#include "uav_oracle.hpp"

#include <algorithm>
#include <thread>

using namespace orc;

struct orc_swarm {
  std::vector<UavSystem> uavs;
};

static ModelParams fromC(const orc_model_params& c) {
  ModelParams p;
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
  p.allocation_matrix = MatX(4, c.n_motors);
  for (int r = 0; r < 4; r++)
    for (int m = 0; m < c.n_motors; m++) p.allocation_matrix(r, m) = c.allocation_matrix[r * ORC_MAX_MOTORS + m];
  return p;
}

static void toC(const ModelParams& p, orc_model_params* c) {
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

static inline int64_t at(const int32_t* idx, int64_t k) {
  return idx ? idx[k] : k;
}

template <class F>
static void parallelFor(int64_t n, int n_threads, F f) {
  if (n_threads <= 1 || n < 2) {
    f(0, int64_t(0), n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t            chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const int64_t b = t * chunk, e = std::min<int64_t>(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([=] { f(t, b, e); });
  }
  for (auto& t : th) t.join();
}

extern "C" {

void orc_model_params_default(orc_model_params* out) {
  toC(ModelParams(), out);
}

orc_swarm* orc_create(int64_t n, int32_t n_types, const orc_model_params* types, const int32_t* type_of_uav, const double* spawn_xyz,
                      const double* spawn_heading) {
  std::vector<ModelParams> tp;
  for (int t = 0; t < n_types; t++) tp.push_back(fromC(types[t]));
  orc_swarm* s = new orc_swarm();
  s->uavs.reserve(n);
  for (int64_t i = 0; i < n; i++) {
    const int t   = type_of_uav ? type_of_uav[i] : 0;
    const V3  pos = spawn_xyz ? v3(spawn_xyz[3 * i], spawn_xyz[3 * i + 1], spawn_xyz[3 * i + 2]) : v3(0, 0, 0);
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
    switch (mode) {  // US:175-248
      case ACTUATOR_CMD:
        for (int m = 0; m < kMaxMotors; m++) u.actuators_cmd.motors[m] = m < stride ? p[m] : 0.0;
        break;
      case CONTROL_GROUP_CMD:
        u.control_group_cmd.roll     = p[0];
        u.control_group_cmd.pitch    = p[1];
        u.control_group_cmd.yaw      = p[2];
        u.control_group_cmd.throttle = p[3];
        break;
      case ATTITUDE_RATE_CMD:
        u.attitude_rate_cmd.rate_x   = p[0];
        u.attitude_rate_cmd.rate_y   = p[1];
        u.attitude_rate_cmd.rate_z   = p[2];
        u.attitude_rate_cmd.throttle = p[3];
        break;
      case ATTITUDE_CMD:
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 3; r++) u.attitude_cmd.orientation(r, c) = p[3 * c + r];
        u.attitude_cmd.throttle = p[9];
        break;
      case TILT_HDG_RATE_CMD:
        u.tilt_hdg_rate_cmd.tilt_vector  = v3(p[0], p[1], p[2]);
        u.tilt_hdg_rate_cmd.heading_rate = p[3];
        u.tilt_hdg_rate_cmd.throttle     = p[4];
        break;
      case ACCELERATION_HDG_RATE_CMD:
        u.acceleration_hdg_rate_cmd.vec = v3(p[0], p[1], p[2]);
        u.acceleration_hdg_rate_cmd.s   = p[3];
        break;
      case ACCELERATION_HDG_CMD:
        u.acceleration_hdg_cmd.vec = v3(p[0], p[1], p[2]);
        u.acceleration_hdg_cmd.s   = p[3];
        break;
      case VELOCITY_HDG_RATE_CMD:
        u.velocity_hdg_rate_cmd.vec = v3(p[0], p[1], p[2]);
        u.velocity_hdg_rate_cmd.s   = p[3];
        break;
      case VELOCITY_HDG_CMD:
        u.velocity_hdg_cmd.vec = v3(p[0], p[1], p[2]);
        u.velocity_hdg_cmd.s   = p[3];
        break;
      case POSITION_CMD:
        u.position_cmd.vec = v3(p[0], p[1], p[2]);
        u.position_cmd.s   = p[3];
        break;
      default:
        mode = INPUT_UNKNOWN;
        break;
    }
    u.active_input = InputMode(mode);
  }
}

void orc_set_feedforward(orc_swarm* s, int32_t kind, int64_t n, const int32_t* idx, const double* payload) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem&    u = s->uavs[at(idx, k)];
    const double* p = payload + 4 * k;
    Vec3Scalar    c;
    c.vec = v3(p[0], p[1], p[2]);
    c.s   = p[3];
    switch (kind) {  // US:254-272
      case 0:
        u.acceleration_hdg_rate_ff     = c;
        u.has_acceleration_hdg_rate_ff = true;
        break;
      case 1:
        u.acceleration_hdg_ff     = c;
        u.has_acceleration_hdg_ff = true;
        break;
      case 2:
        u.velocity_hdg_ff     = c;
        u.has_velocity_hdg_ff = true;
        break;
      case 3:
        u.velocity_hdg_rate_ff     = c;
        u.has_velocity_hdg_rate_ff = true;
        break;
    }
  }
}

void orc_make_step(orc_swarm* s, double dt, int32_t n_steps, int32_t n_threads) {
  parallelFor(int64_t(s->uavs.size()), n_threads, [&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; i++)
      for (int k = 0; k < n_steps; k++) s->uavs[i].makeStep(dt);
  });
}

void orc_get_state(orc_swarm* s, int64_t n, const int32_t* idx, double* x, double* v, double* R, double* omega, double* rpm, double* v_prev,
                   double* imu) {
  for (int64_t k = 0; k < n; k++) {
    const UavSystem& u  = s->uavs[at(idx, k)];
    const State&     st = u.model.state;
    for (int c = 0; c < 3; c++) {
      if (x) x[3 * k + c] = st.x[c];
      if (v) v[3 * k + c] = st.v[c];
      if (omega) omega[3 * k + c] = st.omega[c];
      if (v_prev) v_prev[3 * k + c] = st.v_prev[c];
      if (imu) imu[3 * k + c] = u.model.imu_acceleration[c];
      if (R)
        for (int r = 0; r < 3; r++) R[9 * k + 3 * c + r] = st.R(r, c);
    }
    if (rpm)
      for (int m = 0; m < kMaxMotors; m++) rpm[kMaxMotors * k + m] = m < u.model.params.n_motors ? st.motor_rpm[m] : 0.0;
  }
}

void orc_set_state(orc_swarm* s, int64_t n, const int32_t* idx, const double* x, const double* v, const double* R, const double* omega,
                   const double* rpm) {  // MM:424-433
  for (int64_t k = 0; k < n; k++) {
    UavSystem& u  = s->uavs[at(idx, k)];
    State&     st = u.model.state;
    for (int c = 0; c < 3; c++) {
      if (x) st.x[c] = x[3 * k + c];
      if (v) st.v[c] = v[3 * k + c];
      if (omega) st.omega[c] = omega[3 * k + c];
      if (R)
        for (int r = 0; r < 3; r++) st.R(r, c) = R[9 * k + 3 * c + r];
    }
    if (rpm)
      for (int m = 0; m < u.model.params.n_motors; m++) st.motor_rpm[m] = rpm[kMaxMotors * k + m];
    u.model.pack();
  }
}

void orc_crash(orc_swarm* s, int64_t n, const int32_t* idx) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].crashed = true;
}
void orc_has_crashed(orc_swarm* s, int64_t n, const int32_t* idx, int32_t* out) {
  for (int64_t k = 0; k < n; k++) out[k] = s->uavs[at(idx, k)].crashed;
}
void orc_apply_force(orc_swarm* s, int64_t n, const int32_t* idx, const double* f) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].model.external_force = v3(f[3 * k], f[3 * k + 1], f[3 * k + 2]);
}
void orc_get_force(orc_swarm* s, int64_t n, const int32_t* idx, double* f) {
  for (int64_t k = 0; k < n; k++)
    for (int c = 0; c < 3; c++) f[3 * k + c] = s->uavs[at(idx, k)].model.external_force[c];
}
void orc_set_external_moment(orc_swarm* s, int64_t n, const int32_t* idx, const double* m) {
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].model.external_moment = v3(m[3 * k], m[3 * k + 1], m[3 * k + 2]);
}
void orc_set_params(orc_swarm* s, int64_t n, const int32_t* idx, const orc_model_params* p) {
  const ModelParams mp = fromC(*p);
  for (int64_t k = 0; k < n; k++) s->uavs[at(idx, k)].setParams(mp);
}
void orc_get_params(orc_swarm* s, int64_t uav, orc_model_params* out) {
  toC(s->uavs[uav].model.params, out);
}
void orc_set_controller_params(orc_swarm* s, int32_t which, int64_t n, const int32_t* idx, const double* v) {
  for (int64_t k = 0; k < n; k++) {
    UavSystem& u = s->uavs[at(idx, k)];
    switch (which) {
      case 0: {
        MixerParams p;
        p.desaturation = v[0] != 0.0;
        u.setMixerParams(p);
      } break;
      case 1: {
        RateParams p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2];
        u.setRateControllerParams(p);
      } break;
      case 2: {
        AttitudeParams p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_rate_roll_pitch = v[3], p.max_rate_yaw = v[4];
        u.setAttitudeControllerParams(p);
      } break;
      case 3: {
        VelocityParams p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_acceleration = v[3];
        u.setVelocityControllerParams(p);
      } break;
      case 4: {
        PositionParams p;
        p.kp = v[0], p.kd = v[1], p.ki = v[2], p.max_velocity = v[3];
        u.setPositionControllerParams(p);
      } break;
    }
  }
}
void orc_get_mixer_allocation(orc_swarm* s, int64_t uav, double* out) {
  const UavSystem& u = s->uavs[uav];
  for (int m = 0; m < kMaxMotors; m++)
    for (int c = 0; c < 4; c++) out[4 * m + c] = m < u.model.params.n_motors ? u.mixer.inv(m, c) : 0.0;
}
void orc_get_pid_state(orc_swarm* s, int64_t uav, double* o) {
  const UavSystem& u     = s->uavs[uav];
  const Pid*       p[12] = {&u.position.px, &u.position.py, &u.position.pz, &u.velocity.px, &u.velocity.py, &u.velocity.pz,
                            &u.attitude.px, &u.attitude.py, &u.attitude.pz, &u.rate.px,     &u.rate.py,     &u.rate.pz};
  for (int k = 0; k < 12; k++) {
    o[2 * k]     = p[k]->last_error;
    o[2 * k + 1] = p[k]->integral;
  }
}

// ------------------------------------------------------------------------------------------
// collision loop port.  Candidate search: sorted cell list (2 m cells, 27-cell stencil) instead
// of the KD-tree; metric and predicates exactly as nanoflann's L2_Adaptor for dim 3
// (include/nanoflann.hpp:452-486, 305-309) and src/multirotor_simulator.cpp:326-353.
// Neighbours are visited in ascending j (the reference visits them in KD-tree traversal order;
// the pair SET and crash flags are order-independent, the force sum is order-dependent only in
// its last bits when a UAV has >= 3 simultaneous neighbours).
// ------------------------------------------------------------------------------------------

static inline double nfDist2(const double* a, const double* b) {
  double       result = 0.0;
  const double d0     = a[0] - b[0];
  result += d0 * d0;
  const double d1 = a[1] - b[1];
  result += d1 * d1;
  const double d2 = a[2] - b[2];
  result += d2 * d2;
  return result;
}

static inline int64_t cellOf(double v) {
  double c = std::floor(v * 0.5);
  if (!(c > -1048000.0)) c = -1048000.0;  // also catches NaN
  if (c > 1048000.0) c = 1048000.0;
  return int64_t(c);
}
static inline uint64_t cellKey(int64_t cx, int64_t cy, int64_t cz) {
  const uint64_t off = 1u << 20;
  return (uint64_t(cz + off) << 42) | (uint64_t(cy + off) << 21) | uint64_t(cx + off);
}

int64_t orc_collide_port(int64_t n, const double* xyz, const double* arm, const double* prop, const double* mass, int32_t crash_mode, double rebounce,
                         double* forces, uint8_t* crashed, int32_t* pairs, int64_t cap, int32_t n_threads) {
  std::vector<std::pair<uint64_t, int32_t>> cells(n);
  for (int64_t i = 0; i < n; i++) cells[i] = {cellKey(cellOf(xyz[3 * i]), cellOf(xyz[3 * i + 1]), cellOf(xyz[3 * i + 2])), int32_t(i)};
  std::sort(cells.begin(), cells.end());

  for (int64_t i = 0; i < 3 * n; i++) forces[i] = 0.0;  // SIM:315-319

  const int                         nt = std::max(1, n_threads);
  std::vector<std::vector<int32_t>> found(nt);
  parallelFor(n, nt, [&](int t, int64_t b, int64_t e) {
    std::vector<int32_t> cand;
    for (int64_t i = b; i < e; i++) {
      const double* pi = xyz + 3 * i;
      const int64_t cx = cellOf(pi[0]), cy = cellOf(pi[1]), cz = cellOf(pi[2]);
      cand.clear();
      for (int64_t dz = -1; dz <= 1; dz++)
        for (int64_t dy = -1; dy <= 1; dy++) {
          const uint64_t lo  = cellKey(cx - 1, cy + dy, cz + dz);
          const uint64_t hi  = cellKey(cx + 1, cy + dy, cz + dz);
          auto           it  = std::lower_bound(cells.begin(), cells.end(), std::make_pair(lo, int32_t(-1)));
          for (; it != cells.end() && it->first <= hi; ++it) {
            if (nfDist2(pi, xyz + 3 * int64_t(it->second)) < 3.0) cand.push_back(it->second);  // SIM:326 + NF:305-309
          }
        }
      std::sort(cand.begin(), cand.end());
      for (int32_t j : cand) {
        if (j == i) continue;  // SIM:335
        const double dist = nfDist2(pi, xyz + 3 * int64_t(j));
        const double crit = arm[i] + prop[i] + arm[j] + prop[j];  // SIM:342
        if (dist < crit) {                                          // SIM:346
          found[t].push_back(int32_t(i));
          found[t].push_back(j);
          if (crash_mode) {
            __atomic_store_n(&crashed[j], uint8_t(1), __ATOMIC_RELAXED);  // SIM:348
          } else {
            const V3 rel = v3(pi[0] - xyz[3 * j], pi[1] - xyz[3 * j + 1], pi[2] - xyz[3 * j + 2]);
            const V3 nr  = normalized(rel);
            const double w = mass[j] / (mass[i] + mass[j]);
            for (int c = 0; c < 3; c++) forces[3 * i + c] += ((rebounce * nr[c]) * mass[i]) * w;  // SIM:350
          }
        }
      }
    }
  });
  int64_t count = 0;
  for (int t = 0; t < nt; t++) {
    for (size_t k = 0; k + 1 < found[t].size(); k += 2) {
      if (pairs && count < cap) {
        pairs[2 * count]     = found[t][k];
        pairs[2 * count + 1] = found[t][k + 1];
      }
      count++;
    }
  }
  return count;
}

void orc_handle_collisions(orc_swarm* s, int32_t enabled, int32_t crash, double rebounce, orc_collide_fn ref_engine, int32_t n_threads, int32_t* pairs,
                           int64_t cap, int64_t* count) {
  if (count) *count = 0;
  if (!(crash || enabled)) return;  // SIM:299-301
  const int64_t        n = int64_t(s->uavs.size());
  if (n == 0) return;
  std::vector<double>  xyz(3 * n), arm(n), prop(n), mass(n), forces(3 * n);
  std::vector<uint8_t> crashed(n);
  for (int64_t i = 0; i < n; i++) {
    const UavSystem& u = s->uavs[i];
    for (int c = 0; c < 3; c++) xyz[3 * i + c] = u.model.state.x[c];
    arm[i]     = u.model.params.arm_length;
    prop[i]    = u.model.params.prop_radius;
    mass[i]    = u.model.params.mass;
    crashed[i] = u.crashed;
  }
  orc_collide_fn fn = ref_engine ? ref_engine : orc_collide_port;
  const int64_t  c  = fn(n, xyz.data(), arm.data(), prop.data(), mass.data(), crash, rebounce, forces.data(), crashed.data(), pairs, cap, n_threads);
  if (count) *count = c;
  for (int64_t i = 0; i < n; i++) {
    if (crashed[i]) s->uavs[i].crashed = true;
    s->uavs[i].model.external_force = v3(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);  // SIM:356-358
  }
}

// ---- ROS-wrapper arithmetic (src/uav_system_ros.cpp) ------------------------------------------
static double headingOf(const M3& R) {  // mrs_lib::AttitudeConverter(R).getHeading()
  return std::atan2(R(1, 0), R(0, 0));
}
static void quaternionOf(const M3& m, double* q /* x y z w */) {  // Eigen::Quaterniond(Matrix3d)
  double t = red3(m(0, 0), m(1, 1), m(2, 2));
  if (t > 0.0) {
    t    = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t    = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t    = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t    = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}

void orc_timeout_input(orc_swarm* s, int64_t n, const int32_t* idx) {
  for (int64_t c = 0; c < n; c++) {
    UavSystem&   u  = s->uavs[at(idx, c)];
    const State& st = u.model.state;
    switch (u.active_input) {  // ROSW:480-646
      case POSITION_CMD:
        u.position_cmd.vec = st.x;
        u.position_cmd.s   = headingOf(st.R);
        break;
      case VELOCITY_HDG_CMD:
        u.velocity_hdg_cmd.vec = v3(0, 0, 0);
        u.velocity_hdg_cmd.s   = headingOf(st.R);
        break;
      case VELOCITY_HDG_RATE_CMD:
        u.velocity_hdg_rate_cmd.vec = v3(0, 0, 0);
        u.velocity_hdg_rate_cmd.s   = 0;
        break;
      case ACCELERATION_HDG_CMD:
        u.acceleration_hdg_cmd.vec = v3(0, 0, 0);
        u.acceleration_hdg_cmd.s   = headingOf(st.R);
        break;
      case ACCELERATION_HDG_RATE_CMD:
        u.acceleration_hdg_rate_cmd.vec = v3(0, 0, 0);
        u.acceleration_hdg_rate_cmd.s   = 0;
        break;
      case ATTITUDE_CMD: {
        const double h  = headingOf(st.R);
        const double ch = std::cos(h), sh = std::sin(h);
        M3           R  = identity3();  // AttitudeConverter(0, 0, heading)
        R(0, 0) = ch, R(0, 1) = -sh, R(1, 0) = sh, R(1, 1) = ch;
        u.attitude_cmd.orientation = R;
        u.attitude_cmd.throttle    = 0.0;
      } break;
      case TILT_HDG_RATE_CMD:
        u.tilt_hdg_rate_cmd              = TiltHdgRate();
        u.tilt_hdg_rate_cmd.tilt_vector = v3(0, 0, 1);
        break;
      case ATTITUDE_RATE_CMD:
        u.attitude_rate_cmd = AttitudeRate();
        break;
      case CONTROL_GROUP_CMD:
        u.control_group_cmd = ControlGroup();
        break;
      case ACTUATOR_CMD:
        u.actuators_cmd = Actuators();
        break;
      case INPUT_UNKNOWN:
        break;
    }
  }
}

void orc_get_odometry(orc_swarm* s, int64_t n, const int32_t* idx, double* out) {
  for (int64_t c = 0; c < n; c++) {
    const State& st = s->uavs[at(idx, c)].model.state;
    double*      o  = out + 13 * c;
    for (int k = 0; k < 3; k++) o[k] = st.x[k];
    quaternionOf(st.R, o + 3);
    const V3 vb = mul(transpose(st.R), st.v);
    for (int k = 0; k < 3; k++) {
      o[7 + k]  = vb[k];
      o[10 + k] = st.omega[k];
    }
  }
}

void orc_get_imu(orc_swarm* s, int64_t n, const int32_t* idx, double* out) {
  for (int64_t c = 0; c < n; c++) {
    const UavSystem& u = s->uavs[at(idx, c)];
    double*          o = out + 10 * c;
    for (int k = 0; k < 3; k++) {
      o[k]     = u.model.state.omega[k];
      o[3 + k] = u.model.imu_acceleration[k];
    }
    quaternionOf(u.model.state.R, o + 6);
  }
}

void orc_get_rangefinder(orc_swarm* s, int64_t n, const int32_t* idx, double* out) {
  for (int64_t c = 0; c < n; c++) {
    const UavSystem& u      = s->uavs[at(idx, c)];
    const V3         body_z = u.model.state.R.col(2);
    const V3         dir    = v3(-body_z[0], -body_z[1], -body_z[2]);
    const double     tilt   = std::acos(dot(dir, v3(0, 0, -1)));
    double           range;
    if (body_z[2] > 0) {
      range = (u.model.state.x[2] - u.model.params.ground_z) / std::cos(tilt) + 0.01;
    } else {
      range = 1.7976931348623157e308;
    }
    if (range > 40.0) range = 41.0;
    out[c] = range;
  }
}

void orc_set_mass(orc_swarm* s, int64_t n, const int32_t* idx, const double* mass) {
  for (int64_t c = 0; c < n; c++) {
    UavSystem&   u = s->uavs[at(idx, c)];
    ModelParams  p = u.model.params;
    const double original = p.mass;
    p.mass                = mass[c];
    for (int m = 0; m < p.n_motors; m++) p.allocation_matrix(2, m) = p.mass * (p.allocation_matrix(2, m) / original);
    p.J       = zero3();
    p.J(0, 0) = p.mass * (3.0 * p.arm_length * p.arm_length + p.body_height * p.body_height) / 12.0;
    p.J(1, 1) = p.mass * (3.0 * p.arm_length * p.arm_length + p.body_height * p.body_height) / 12.0;
    p.J(2, 2) = (p.mass * p.arm_length * p.arm_length) / 2.0;
    u.setParams(p);
  }
}

void orc_set_ground_z(orc_swarm* s, int64_t n, const int32_t* idx, const double* z) {
  for (int64_t c = 0; c < n; c++) {
    UavSystem&  u = s->uavs[at(idx, c)];
    ModelParams p = u.model.params;
    p.ground_z    = z[c];
    u.setParams(p);
  }
}

double orc_pid_update(double* state2, double kp, double kd, double ki, double saturation, double antiwindup, double error, double dt) {
  Pid p;
  p.setParams(kp, kd, ki, saturation, antiwindup);
  p.last_error    = state2[0];
  p.integral      = state2[1];
  const double u  = p.update(error, dt);
  state2[0]       = p.last_error;
  state2[1]       = p.integral;
  return u;
}

double orc_u01(uint64_t seed, uint64_t stream, uint64_t index) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((stream << 32) + index);
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return double(z >> 11) * (1.0 / 9007199254740992.0);
}

}  // extern "C"
