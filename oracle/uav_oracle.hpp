// oracle/uav_oracle.hpp — CPU restatement of the reference stepping path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
// build, load or call anything in oracle/.  The product (libmrsb) never does.
//
// What this restates (reference file:line, relative to the reference repository root):
//   US   include/mrs_multirotor_simulator/uav_system/uav_system.hpp          makeStep  US:304-380
//   MM   include/mrs_multirotor_simulator/uav_system/multirotor_model.hpp    step      MM:220-286
//                                                                            f(x)      MM:301-366
//   CTL  include/mrs_multirotor_simulator/uav_system/controllers/*.hpp       PID + 5 controllers + mixer
//   ODE  .../ode/boost/numeric/odeint/stepper/runge_kutta4.hpp:42-95 + algebra/default_operations.hpp:77-154
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES: oracle/_ref/libref_uavsystem.so is the
// reference's uav_system.hpp + multirotor_model.hpp + controllers/*.hpp compiled unmodified from
// where they lie (oracle/ref_uavsystem.cpp, `make -C oracle refsys`); this restatement equals it
// BIT FOR BIT in every input mode on 4/6/8-motor airframes over 10 s of flight, through every
// desaturation branch, patch, feed-forward, NaN guard and parameter reset
// (tests/test_ref_uavsystem.py), and the golden fixtures are generated from it.
// What that build cannot pin is Eigen itself: the image has neither Eigen nor Boost (no network),
// so the two libraries the reference delegates leaf arithmetic to are replaced by the stand-ins
// under oracle/shim, and both this file and the stand-in follow Eigen 3.3.7's (Ubuntu 20.04 / ROS
// Noetic) evaluation rules as far as they are known here.  tests/test_ref_uavsystem.py also builds
// the alternative reading of those rules and measures its effect after 10 s: <= 2e-12 m, three
// orders of magnitude below the stated tolerance.  The rules:
//   * fixed-size 3-term reductions (dot, squaredNorm, 3x3 product coefficients) are a + (b + c)
//     (redux_novec_unroller splits [0,1) | [1,3));
//   * dynamic small products are coefficient-based with a sequential inner sum;
//   * MatrixXd * vector goes through the column-major GEMV kernel: blocks of four columns
//     combined as (c0 + c1) + (c2 + c3), leftover columns added one at a time;
//   * VectorXd::mean()/sum() is the SSE2 two-lane packet reduction;
//   * Matrix3d::inverse() is the cofactor formula times the reciprocal determinant;
//   * dynamic inverse() is PartialPivLU (unblocked) + column-oriented triangular solves that
//     multiply by the reciprocal pivot;
//   * LLT<Matrix3d> is the unblocked lower Cholesky; normalized()/normalize() divide by sqrt(z)
//     only if z > 0.
// Any remaining summation-order difference is <= 1 ulp per operation and is covered by the stated
// state tolerance; it cannot affect collision pair lists, whose predicate is restated
// exactly and pinned against the real vendored nanoflann (oracle/ref_nanoflann.cpp).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

constexpr int kMaxMotors = 8;

// ------------------------------------------------------------------------------------------
// tiny fixed-size algebra with Eigen's evaluation order
// ------------------------------------------------------------------------------------------

inline double red3(double a, double b, double c) {
  return a + (b + c);
}

struct V3 {
  double c[3];
  double&       operator[](int i) { return c[i]; }
  const double& operator[](int i) const { return c[i]; }
};

inline V3 v3(double x, double y, double z) {
  V3 r;
  r[0] = x;
  r[1] = y;
  r[2] = z;
  return r;
}
inline V3     operator+(const V3& a, const V3& b) { return v3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3     operator-(const V3& a, const V3& b) { return v3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3     operator*(const V3& a, double s) { return v3(a[0] * s, a[1] * s, a[2] * s); }
inline V3     operator*(double s, const V3& a) { return v3(s * a[0], s * a[1], s * a[2]); }
inline V3     operator/(const V3& a, double s) { return v3(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const V3& a, const V3& b) { return red3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
inline double squaredNorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(squaredNorm(a)); }
inline V3     cross(const V3& a, const V3& b) {
  return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline V3 normalized(const V3& a) {
  const double z = squaredNorm(a);
  if (z > 0.0) {
    return a / std::sqrt(z);
  }
  return a;
}

struct M3 {
  double m[3][3];  // m[row][col]
  double&       operator()(int r, int c) { return m[r][c]; }
  const double& operator()(int r, int c) const { return m[r][c]; }
  V3            col(int c) const { return v3(m[0][c], m[1][c], m[2][c]); }
  void          setCol(int c, const V3& v) {
    m[0][c] = v[0];
    m[1][c] = v[1];
    m[2][c] = v[2];
  }
};

inline M3 zero3() {
  M3 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
inline M3 identity3() {
  M3 r = zero3();
  r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
  return r;
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
  return r;
}
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = red3(a(i, 0) * b(0, j), a(i, 1) * b(1, j), a(i, 2) * b(2, j));
  return r;
}
inline V3 mul(const M3& a, const V3& v) {
  V3 r;
  for (int i = 0; i < 3; i++) r[i] = red3(a(i, 0) * v[0], a(i, 1) * v[1], a(i, 2) * v[2]);
  return r;
}

// Matrix3d::inverse(): cofactors, det from column 0, multiply by 1/det.
inline double cofactor3(const M3& a, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return a(i1, j1) * a(i2, j2) - a(i1, j2) * a(i2, j1);
}
inline M3 inverse3(const M3& a) {
  const double c00 = cofactor3(a, 0, 0), c10 = cofactor3(a, 1, 0), c20 = cofactor3(a, 2, 0);
  const double det    = red3(c00 * a(0, 0), c10 * a(1, 0), c20 * a(2, 0));
  const double invdet = 1.0 / det;
  M3           r;
  r(0, 0) = c00 * invdet;
  r(0, 1) = c10 * invdet;
  r(0, 2) = c20 * invdet;
  for (int row = 1; row < 3; row++)
    for (int k = 0; k < 3; k++) r(row, k) = cofactor3(a, k, row) * invdet;
  return r;
}

// LLT<Matrix3d>(A).matrixL() as a dense matrix (strict upper triangle zero).  Unblocked lower
// Cholesky reading only the lower triangle; on a non-positive pivot the factorisation stops and
// the untouched entries keep A's values, like Eigen's in-place routine.
inline M3 cholL(const M3& A) {
  M3 w = A;
  for (int k = 0; k < 3; k++) {
    double x = w(k, k);
    if (k == 1) x -= w(1, 0) * w(1, 0);
    if (k == 2) x -= (w(2, 0) * w(2, 0) + w(2, 1) * w(2, 1));
    if (x <= 0.0) break;
    x       = std::sqrt(x);
    w(k, k) = x;
    if (k == 1) w(2, 1) -= w(2, 0) * w(1, 0);
    for (int r = k + 1; r < 3; r++) w(r, k) /= x;
  }
  w(0, 1) = w(0, 2) = w(1, 2) = 0.0;
  return w;
}

// R * chol(R^T R).inverse()   (MM:249-253 and MM:314-316)
inline M3 reorthonormalize(const M3& R) {
  const M3 P = cholL(mul(transpose(R), R));
  return mul(R, inverse3(P));
}

// ------------------------------------------------------------------------------------------
// small dynamic matrices (row-major storage, <= 8x8) with Eigen's dynamic-path arithmetic
// ------------------------------------------------------------------------------------------

struct MatX {
  int    rows = 0, cols = 0;
  double a[64];
  MatX() { std::memset(a, 0, sizeof(a)); }
  MatX(int r, int c) : rows(r), cols(c) { std::memset(a, 0, sizeof(a)); }
  double&       operator()(int r, int c) { return a[r * 8 + c]; }
  const double& operator()(int r, int c) const { return a[r * 8 + c]; }
};

inline MatX transposeX(const MatX& m) {
  MatX r(m.cols, m.rows);
  for (int i = 0; i < m.rows; i++)
    for (int j = 0; j < m.cols; j++) r(j, i) = m(i, j);
  return r;
}

// coefficient-based lazy product, sequential inner sum
inline MatX mulX(const MatX& x, const MatX& y) {
  MatX r(x.rows, y.cols);
  for (int i = 0; i < x.rows; i++)
    for (int j = 0; j < y.cols; j++) {
      double s = x(i, 0) * y(0, j);
      for (int k = 1; k < x.cols; k++) s = s + x(i, k) * y(k, j);
      r(i, j) = s;
    }
  return r;
}

// PartialPivLU(m).inverse()
inline MatX inverseX(const MatX& m) {
  const int n  = m.rows;
  MatX      lu = m;
  int       tr[8];
  for (int k = 0; k < n; k++) {
    int    piv  = k;
    double best = std::fabs(lu(k, k));
    for (int r = k + 1; r < n; r++) {
      if (std::fabs(lu(r, k)) > best) {
        best = std::fabs(lu(r, k));
        piv  = r;
      }
    }
    tr[k] = piv;
    if (best != 0.0) {
      if (piv != k) {
        for (int c = 0; c < n; c++) {
          const double t = lu(k, c);
          lu(k, c)       = lu(piv, c);
          lu(piv, c)     = t;
        }
      }
      for (int r = k + 1; r < n; r++) lu(r, k) /= lu(k, k);
    }
    for (int r = k + 1; r < n; r++)
      for (int c = k + 1; c < n; c++) lu(r, c) -= lu(r, k) * lu(k, c);
  }
  MatX x(n, n);
  for (int i = 0; i < n; i++) x(i, i) = 1.0;
  for (int k = 0; k < n; k++) {
    if (tr[k] != k) {
      for (int c = 0; c < n; c++) {
        const double t = x(k, c);
        x(k, c)        = x(tr[k], c);
        x(tr[k], c)    = t;
      }
    }
  }
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < n; i++) {  // unit lower
      const double b = x(i, j);
      for (int r = i + 1; r < n; r++) x(r, j) -= b * lu(r, i);
    }
    for (int i = n - 1; i >= 0; i--) {  // upper, reciprocal pivot
      const double a = 1.0 / lu(i, i);
      const double b = (x(i, j) *= a);
      for (int r = 0; r < i; r++) x(r, j) -= b * lu(r, i);
    }
  }
  return x;
}

// column-major GEMV of Eigen 3.3: res = M * v  (res starts at zero)
inline void gemv(const MatX& M, const double* v, double* res) {
  const int blocks = (M.cols / 4) * 4;
  for (int r = 0; r < M.rows; r++) {
    double acc = 0.0;
    for (int c = 0; c < blocks; c += 4) {
      acc = acc + ((M(r, c) * v[c] + M(r, c + 1) * v[c + 1]) + (M(r, c + 2) * v[c + 2] + M(r, c + 3) * v[c + 3]));
    }
    for (int c = blocks; c < M.cols; c++) acc = acc + M(r, c) * v[c];
    res[r] = acc;
  }
}

// VectorXd::sum() on SSE2 (two-lane packets, two accumulators)
inline double sumX(const double* v, int n) {
  if (n == 0) return 0.0;
  if (n == 1) return v[0];
  const int aligned2 = (n / 4) * 4;
  const int aligned  = (n / 2) * 2;
  double    p0[2]    = {v[0], v[1]};
  if (aligned > 2) {
    double p1[2] = {v[2], v[3]};
    for (int i = 4; i < aligned2; i += 4) {
      p0[0] += v[i];
      p0[1] += v[i + 1];
      p1[0] += v[i + 2];
      p1[1] += v[i + 3];
    }
    p0[0] += p1[0];
    p0[1] += p1[1];
    if (aligned > aligned2) {
      p0[0] += v[aligned2];
      p0[1] += v[aligned2 + 1];
    }
  }
  double res = p0[0] + p0[1];
  for (int i = aligned; i < n; i++) res += v[i];
  return res;
}
inline double meanX(const double* v, int n) {
  return sumX(v, n) / double(n);
}

// ------------------------------------------------------------------------------------------
// parameters  (MM:24-88; controller Params classes in CTL/*.hpp)
// ------------------------------------------------------------------------------------------

struct ModelParams {
  int    n_motors             = 4;
  double g                    = 9.81;
  double mass                 = 2.0;
  double kf                   = 0.00000027087;
  double km                   = 0.07;
  double prop_radius          = 0.15;
  double arm_length           = 0.25;
  double body_height          = 0.1;
  double motor_time_constant  = 0.03;
  double max_rpm              = 7800;
  double min_rpm              = 1170;
  double air_resistance_coeff = 0.30;
  M3     J;
  MatX   allocation_matrix;  // 4 x n_motors, scaled
  bool   ground_enabled        = false;
  double ground_z              = 0.0;
  bool   takeoff_patch_enabled = true;

  // header defaults = x500 (MM:26-66)
  ModelParams() {
    J       = zero3();
    J(0, 0) = mass * (3.0 * arm_length * arm_length + body_height * body_height) / 12.0;
    J(1, 1) = mass * (3.0 * arm_length * arm_length + body_height * body_height) / 12.0;
    J(2, 2) = (mass * arm_length * arm_length) / 2.0;

    allocation_matrix       = MatX(4, 4);
    const double quad[4][4] = {{-0.707, 0.707, 0.707, -0.707}, {-0.707, 0.707, -0.707, 0.707}, {-1, -1, 1, 1}, {1, 1, 1, 1}};
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) allocation_matrix(r, c) = quad[r][c];
    scaleAllocation();
  }

  // MM:59-62 / ROSW:100-103
  void scaleAllocation() {
    for (int c = 0; c < allocation_matrix.cols; c++) {
      allocation_matrix(0, c) *= arm_length * kf;
      allocation_matrix(1, c) *= arm_length * kf;
      allocation_matrix(2, c) *= km * (3.0 * prop_radius) * kf;
      allocation_matrix(3, c) *= kf;
    }
  }
};

struct MixerParams {
  bool desaturation = true;
};
struct RateParams {
  double kp = 4.0, kd = 0.04, ki = 0.0;
};
struct AttitudeParams {
  double kp = 6.0, kd = 0.05, ki = 0.01, max_rate_roll_pitch = 10.0, max_rate_yaw = 1.0;
};
struct VelocityParams {
  double kp = 2.0, kd = 0.05, ki = 0.01, max_acceleration = 4.0;
};
struct PositionParams {
  double kp = 2.0, kd = 0.15, ki = 0.2, max_velocity = 6.0;
};

// ------------------------------------------------------------------------------------------
// PID  (CTL/pid.hpp:67-96)
// ------------------------------------------------------------------------------------------

struct Pid {
  double kp = 0, kd = 0, ki = 0;
  double last_error = 0, integral = 0;
  double saturation = -1, antiwindup = -1;

  void reset() { last_error = integral = 0; }
  void setParams(double p, double d, double i, double sat, double aw) {
    kp         = p;
    kd         = d;
    ki         = i;
    saturation = sat;
    antiwindup = aw;
  }
  double update(double error, double dt) {
    const double difference = (error - last_error) / dt;
    last_error              = error;
    const double pc         = kp * error;
    const double dc         = kd * difference;
    const double ic         = ki * integral;
    double       sum        = pc + dc + ic;
    if (saturation > 0) {
      if (sum >= saturation) {
        sum = saturation;
      } else if (sum <= -saturation) {
        sum = -saturation;
      }
    }
    if (antiwindup > 0) {
      if (std::fabs(sum) < antiwindup) {
        integral += error * dt;
      }
    }
    return sum;
  }
};

// ------------------------------------------------------------------------------------------
// model state (MM:90-98) and commands (CTL/references.hpp)
// ------------------------------------------------------------------------------------------

struct State {
  V3     x, v, v_prev, omega;
  M3     R;
  double motor_rpm[kMaxMotors];
};

enum InputMode {
  INPUT_UNKNOWN = 0,
  ACTUATOR_CMD,
  CONTROL_GROUP_CMD,
  ATTITUDE_RATE_CMD,
  ATTITUDE_CMD,
  TILT_HDG_RATE_CMD,
  ACCELERATION_HDG_RATE_CMD,
  ACCELERATION_HDG_CMD,
  VELOCITY_HDG_RATE_CMD,
  VELOCITY_HDG_CMD,
  POSITION_CMD
};

struct Actuators {
  double motors[kMaxMotors] = {0, 0, 0, 0, 0, 0, 0, 0};
};
struct ControlGroup {
  double roll = 0, pitch = 0, yaw = 0, throttle = 0;
};
struct AttitudeRate {
  double rate_x = 0, rate_y = 0, rate_z = 0, throttle = 0;
};
struct Attitude {
  M3     orientation = identity3();
  double throttle    = 0;
};
struct TiltHdgRate {
  V3     tilt_vector  = v3(1, 0, 0);  // Eigen::Vector3d::Identity() (CTL/references.hpp:123)
  double heading_rate = 0, throttle = 0;
};
struct Vec3Scalar {  // AccelerationHdgRate / AccelerationHdg / VelocityHdgRate / VelocityHdg / Position
  V3     vec = v3(0, 0, 0);
  double s   = 0;  // heading or heading_rate
};

// ------------------------------------------------------------------------------------------
// MultirotorModel (MM)
// ------------------------------------------------------------------------------------------

class Model {
public:
  ModelParams params;
  State       state;
  V3          imu_acceleration;
  double      input[kMaxMotors];
  V3          external_force, external_moment, initial_pos;
  double      xs[18];  // InternalState (MM:204-214)

  void initializeState() {  // MM:183-198
    state.x = state.v = state.v_prev = state.omega = v3(0, 0, 0);
    state.R                                      = identity3();
    imu_acceleration                             = v3(0, 0, 0);
    for (int i = 0; i < kMaxMotors; i++) state.motor_rpm[i] = input[i] = 0.0;
    external_force = external_moment = v3(0, 0, 0);
  }

  void pack() {  // MM:204-214
    for (int i = 0; i < 3; i++) {
      xs[0 + i]  = state.x[i];
      xs[3 + i]  = state.v[i];
      xs[6 + i]  = state.R(i, 0);
      xs[9 + i]  = state.R(i, 1);
      xs[12 + i] = state.R(i, 2);
      xs[15 + i] = state.omega[i];
    }
  }

  void setStatePos(const V3& pos, double heading) {  // MM:439-446, AngleAxis(-heading, z).toRotationMatrix()
    initial_pos      = pos;
    state.x          = pos;
    const double ang = -heading;
    const double s = std::sin(ang), c = std::cos(ang);
    const double c1z = (1.0 - c) * 1.0;  // cos1_axis.z
    M3           R;
    R(0, 1) = 0.0 - s;
    R(1, 0) = 0.0 + s;
    R(0, 2) = 0.0 + 0.0;
    R(2, 0) = 0.0 - 0.0;
    R(1, 2) = 0.0 - 0.0;
    R(2, 1) = 0.0 + 0.0;
    R(0, 0) = 0.0 + c;
    R(1, 1) = 0.0 + c;
    R(2, 2) = c1z * 1.0 + c;
    state.R = R;
    pack();
  }

  void setInput(const Actuators& in) {  // MM:392-410
    for (int i = 0; i < params.n_motors; i++) {
      double val = in.motors[i];
      if (!std::isfinite(val)) val = 0;
      if (val < 0.0) {
        val = 0.0;
      } else if (val > 1.0) {
        val = 1.0;
      }
      input[i] = params.min_rpm + (params.max_rpm - params.min_rpm) * val;
    }
  }

  // MM:301-366
  void derivative(const double* x, double* dxdt) const {
    V3 cx, cv, cw;
    M3 cR;
    for (int i = 0; i < 3; i++) {
      cx[i]    = x[0 + i];
      cv[i]    = x[3 + i];
      cR(i, 0) = x[6 + i];
      cR(i, 1) = x[9 + i];
      cR(i, 2) = x[12 + i];
      cw[i]    = x[15 + i];
    }
    (void)cx;
    const M3 R = reorthonormalize(cR);

    M3 W    = zero3();
    W(2, 1) = cw[0];
    W(1, 2) = -cw[0];
    W(0, 2) = cw[1];
    W(2, 0) = -cw[1];
    W(1, 0) = cw[2];
    W(0, 1) = -cw[2];

    double sq[kMaxMotors];
    for (int i = 0; i < params.n_motors; i++) sq[i] = state.motor_rpm[i] * state.motor_rpm[i];
    double tt[4];
    gemv(params.allocation_matrix, sq, tt);
    const double thrust = tt[3];

    const double resistance = params.air_resistance_coeff * M_PI * (params.arm_length) * (params.arm_length) * norm(cv) * norm(cv);

    V3 vnorm = cv;
    if (norm(vnorm) != 0) {
      vnorm = normalized(vnorm);  // normalize(): z > 0 guard then divide by sqrt(z)
    }

    const V3 x_dot = cv;
    const V3 v_dot = ((v3(-0.0, -0.0, -params.g) + (thrust * R.col(2)) / params.mass) + external_force / params.mass) - (resistance * vnorm) / params.mass;
    const M3 R_dot = mul(R, W);

    const V3 Jw        = mul(params.J, cw);
    const V3 rhs       = (v3(tt[0], tt[1], tt[2]) - cross(cw, Jw)) + external_moment;
    const V3 omega_dot = mul(inverse3(params.J), rhs);

    for (int i = 0; i < 3; i++) {
      dxdt[0 + i]  = x_dot[i];
      dxdt[3 + i]  = v_dot[i];
      dxdt[6 + i]  = R_dot(i, 0);
      dxdt[9 + i]  = R_dot(i, 1);
      dxdt[12 + i] = R_dot(i, 2);
      dxdt[15 + i] = omega_dot[i];
    }
    for (int i = 0; i < 18; i++) {
      if (std::isnan(dxdt[i])) dxdt[i] = 0;
    }
  }

  // odeint runge_kutta4::do_step (ODE/stepper/detail/generic_rk_algorithm.hpp:190-234,
  // generic_rk_operations.hpp:30-68, algebra/default_operations.hpp:77-154); coefficients
  // ODE/stepper/runge_kutta4.hpp:42-95.  The zero coefficients are multiplied through.
  void rk4(double dt) {
    const double a21 = (1.0 / 2.0) * dt;
    const double a31 = 0.0 * dt, a32 = (1.0 / 2.0) * dt;
    const double a41 = 0.0 * dt, a42 = 0.0 * dt, a43 = 1.0 * dt;
    const double b1 = (1.0 / 6.0) * dt, b2 = (1.0 / 3.0) * dt, b3 = (1.0 / 3.0) * dt, b4 = (1.0 / 6.0) * dt;
    double       k1[18], k2[18], k3[18], k4[18], xt[18];
    derivative(xs, k1);
    for (int i = 0; i < 18; i++) xt[i] = 1.0 * xs[i] + a21 * k1[i];
    derivative(xt, k2);
    for (int i = 0; i < 18; i++) xt[i] = 1.0 * xs[i] + a31 * k1[i] + a32 * k2[i];
    derivative(xt, k3);
    for (int i = 0; i < 18; i++) xt[i] = 1.0 * xs[i] + a41 * k1[i] + a42 * k2[i] + a43 * k3[i];
    derivative(xt, k4);
    for (int i = 0; i < 18; i++) xs[i] = 1.0 * xs[i] + b1 * k1[i] + b2 * k2[i] + b3 * k3[i] + b4 * k4[i];
  }

  void step(double dt) {  // MM:220-286
    double save[18];
    std::memcpy(save, xs, sizeof(save));
    rk4(dt);
    for (int i = 0; i < 18; i++) {
      if (std::isnan(xs[i])) {
        std::memcpy(xs, save, sizeof(save));
        break;
      }
    }
    for (int i = 0; i < 3; i++) {
      state.x[i]     = xs[0 + i];
      state.v[i]     = xs[3 + i];
      state.R(i, 0)  = xs[6 + i];
      state.R(i, 1)  = xs[9 + i];
      state.R(i, 2)  = xs[12 + i];
      state.omega[i] = xs[15 + i];
    }

    const double filter_const = std::exp((-dt) / (params.motor_time_constant));
    for (int i = 0; i < params.n_motors; i++) state.motor_rpm[i] = filter_const * state.motor_rpm[i] + (1.0 - filter_const) * input[i];

    state.R = reorthonormalize(state.R);

    if (params.ground_enabled) {
      if (state.x[2] < params.ground_z && state.v[2] < 0) {
        state.x[2]  = params.ground_z;
        state.v     = v3(0, 0, 0);
        state.omega = v3(0, 0, 0);
      }
    }

    if (params.takeoff_patch_enabled) {
      const double hover_rpm = std::sqrt((params.mass * params.g) / (params.n_motors * params.kf));
      if (meanX(input, params.n_motors) <= 0.90 * hover_rpm) {
        if (state.x[2] < initial_pos[2] && state.v[2] < 0) {
          state.x[2]  = initial_pos[2];
          state.v     = v3(0, 0, 0);
          state.omega = v3(0, 0, 0);
        }
      } else {
        params.takeoff_patch_enabled = false;
      }
    }

    const V3 lin = ((state.v - state.v_prev) / dt) + v3(0, 0, params.g);
    imu_acceleration = mul(transpose(state.R), lin);
    state.v_prev     = state.v;

    pack();
  }
};

// ------------------------------------------------------------------------------------------
// controllers
// ------------------------------------------------------------------------------------------

struct Mixer {  // CTL/mixer.hpp
  ModelParams model;
  MixerParams params;
  MatX        inv;  // n x 4

  void calculateAllocation() {  // :72-101
    const MatX A  = model.allocation_matrix;
    const MatX At = transposeX(A);
    inv           = mulX(At, inverseX(mulX(A, At)));
    for (int i = 0; i < model.n_motors; i++) {
      const double z = inv(i, 0) * inv(i, 0) + inv(i, 1) * inv(i, 1);
      if (z > 0.0) {
        const double s = std::sqrt(z);
        inv(i, 0) /= s;
        inv(i, 1) /= s;
      }
    }
    for (int i = 0; i < model.n_motors; i++) {
      if (inv(i, 2) > 1e-2) {
        inv(i, 2) = 1.0;
      } else if (inv(i, 2) < -1e-2) {
        inv(i, 2) = -1.0;
      } else {
        inv(i, 2) = 0.0;
      }
    }
    for (int i = 0; i < model.n_motors; i++) inv(i, 3) = 1.0;
  }

  Actuators getControlSignal(const ControlGroup& ref) const {  // :107-144
    double    cg[4] = {ref.roll, ref.pitch, ref.yaw, ref.throttle};
    Actuators out;
    const int n = model.n_motors;
    gemv(inv, cg, out.motors);
    if (params.desaturation) {
      double mn = out.motors[0];
      for (int i = 1; i < n; i++) mn = out.motors[i] < mn ? out.motors[i] : mn;
      if (mn < 0.0) {
        for (int i = 0; i < n; i++) out.motors[i] += std::fabs(mn);
      }
      double mx = out.motors[0];
      for (int i = 1; i < n; i++) mx = out.motors[i] > mx ? out.motors[i] : mx;
      if (mx > 1.0) {
        if (ref.throttle > 1e-2) {
          for (int i = 0; i < 3; i++) cg[i] = cg[i] / (meanX(out.motors, n) / ref.throttle);
          gemv(inv, cg, out.motors);
        } else {
          for (int i = 0; i < n; i++) out.motors[i] /= mx;
        }
      }
    }
    return out;
  }
};

struct RateController {  // CTL/rate_controller.hpp
  ModelParams model;
  RateParams  params;
  Pid         px, py, pz;
  void        initializePIDs() {  // :56-65
    px.reset();
    py.reset();
    pz.reset();
    px.setParams(params.kp * model.J(0, 0), params.kd * model.J(0, 0), params.ki * model.J(0, 0), -1, 1.0);
    py.setParams(params.kp * model.J(1, 1), params.kd * model.J(1, 1), params.ki * model.J(1, 1), -1, 1.0);
    pz.setParams(params.kp * model.J(2, 2), params.kd * model.J(2, 2), params.ki * model.J(2, 2), -1, 1.0);
  }
  ControlGroup getControlSignal(const State& st, const AttitudeRate& ref, double dt) {  // :67-81
    const V3     e = v3(ref.rate_x, ref.rate_y, ref.rate_z) - st.omega;
    ControlGroup o;
    o.roll     = px.update(e[0], dt);
    o.pitch    = py.update(e[1], dt);
    o.yaw      = pz.update(e[2], dt);
    o.throttle = ref.throttle;
    return o;
  }
};

inline int signum(double val) {
  return (0.0 < val) - (val < 0.0);
}

struct AttitudeController {  // CTL/attitude_controller.hpp
  AttitudeParams params;
  Pid            px, py, pz;
  void           initializePIDs() {  // :162-171
    px.reset();
    py.reset();
    pz.reset();
    px.setParams(params.kp, params.kd, params.ki, params.max_rate_roll_pitch, 0.1);
    py.setParams(params.kp, params.kd, params.ki, params.max_rate_roll_pitch, 0.1);
    pz.setParams(params.kp, params.kd, params.ki, params.max_rate_yaw, 0.1);
  }

  static V3 errorVec(const M3& Rd, const M3& R) {  // :82-89
    const M3 a = mul(transpose(Rd), R);
    const M3 b = mul(transpose(R), Rd);
    M3       E;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) E(i, j) = 0.5 * (a(i, j) - b(i, j));
    return v3((E(1, 2) - E(2, 1)) / 2.0, (E(2, 0) - E(0, 2)) / 2.0, (E(0, 1) - E(1, 0)) / 2.0);
  }

  AttitudeRate getControlSignal(const State& st, const Attitude& ref, double dt) {  // :79-100
    const V3     e = errorVec(ref.orientation, st.R);
    AttitudeRate o;
    o.rate_x   = px.update(e[0], dt);
    o.rate_y   = py.update(e[1], dt);
    o.rate_z   = pz.update(e[2], dt);
    o.throttle = ref.throttle;
    return o;
  }

  static double intrinsicBodyRateToHeadingRate(const M3& R, const V3& w) {  // :177-206
    M3 W;
    W(0, 0) = 0;
    W(0, 1) = -w[2];
    W(0, 2) = w[1];
    W(1, 0) = w[2];
    W(1, 1) = 0;
    W(1, 2) = -w[0];
    W(2, 0) = -w[1];
    W(2, 1) = w[0];
    W(2, 2) = 0;
    const M3     R_d   = mul(R, W);
    const double rx    = R(0, 0);
    const double ry    = R(1, 0);
    const double denom = rx * rx + ry * ry;
    double       ax = 0, ay = 0;
    if (std::fabs(denom) <= 1e-5) {
      // reference only prints a warning
    } else {
      ax = -ry / denom;
      ay = rx / denom;
    }
    return ax * R_d(0, 0) + ay * R_d(1, 0);
  }

  static double getYawRateIntrinsic(const M3& R, double heading_rate) {  // :212-251
    if (std::fabs(heading_rate) < 1e-3) return 0;
    const V3 hv  = v3(R(0, 0), R(1, 0), 0);
    const V3 orb = cross(v3(0, 0, heading_rate), hv);
    V3       b   = cross(v3(0, 0, 1), hv);
    b            = normalized(b);
    M3 P;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) P(i, j) = b[i] * b[j];
    const V3     projected = mul(P, R.col(1));
    const double on        = norm(orb);
    const double pn        = norm(projected);
    if (std::fabs(pn) < 1e-5) return 0;
    const double direction = signum(dot(orb, projected));
    const double out       = direction * (on / pn);
    if (!std::isfinite(out)) return 0;
    return out;
  }

  AttitudeRate getControlSignal(const State& st, const TiltHdgRate& ref, double dt) {  // :106-145
    M3 Rd = zero3();
    Rd.setCol(2, normalized(ref.tilt_vector));
    Rd.setCol(1, normalized(cross(Rd.col(2), st.R.col(0))));
    Rd.setCol(0, normalized(cross(Rd.col(1), Rd.col(2))));
    const V3     e         = errorVec(Rd, st.R);
    double       rate_x    = px.update(e[0], dt);
    double       rate_y    = py.update(e[1], dt);
    double       rate_z    = pz.update(e[2], dt);
    const double parasitic = intrinsicBodyRateToHeadingRate(st.R, v3(rate_x, rate_y, rate_z));
    rate_z += getYawRateIntrinsic(st.R, ref.heading_rate - parasitic);
    AttitudeRate o;
    o.rate_x   = rate_x;
    o.rate_y   = rate_y;
    o.rate_z   = rate_z;
    o.throttle = ref.throttle;
    return o;
  }
};

struct AccelerationController {  // CTL/acceleration_controller.hpp
  ModelParams model;

  double throttleFor(const V3& fd, const State& st) const {  // :89-94 / :117-120
    const double thrust_force = dot(fd, st.R.col(2));
    return (std::sqrt(thrust_force / (model.kf * model.n_motors)) - model.min_rpm) / (model.max_rpm - model.min_rpm);
  }

  Attitude getControlSignalHdg(const State& st, const Vec3Scalar& ref) const {  // :44-97
    const V3 fd      = (ref.vec + v3(0, 0, model.g)) * model.mass;
    const V3 fd_norm = normalized(fd);
    const V3 bxd     = v3(std::cos(ref.s), std::sin(ref.s), 0.0);

    M3 Rd;
    Rd.setCol(2, fd_norm);

    M3 proj;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) proj(i, j) = (i == j ? 1.0 : 0.0) - fd_norm[i] * fd_norm[j];

    MatX A(3, 2), B(3, 2);
    for (int i = 0; i < 3; i++) {
      A(i, 0) = proj(i, 0);
      A(i, 1) = proj(i, 1);
    }
    B(0, 0) = 1;
    B(1, 1) = 1;

    const MatX Bt    = transposeX(B);
    const MatX BtA   = mulX(Bt, A);
    const MatX BtAt  = transposeX(BtA);
    const MatX pinv  = mulX(inverseX(mulX(BtAt, BtA)), BtAt);
    const MatX obl   = mulX(mulX(A, pinv), Bt);
    double     b3[3] = {bxd[0], bxd[1], bxd[2]};
    double     c0[3];
    gemv(obl, b3, c0);

    Rd.setCol(0, normalized(v3(c0[0], c0[1], c0[2])));
    Rd.setCol(1, normalized(cross(Rd.col(2), Rd.col(0))));

    Attitude out;
    out.orientation = Rd;
    out.throttle    = throttleFor(fd, st);
    return out;
  }

  TiltHdgRate getControlSignalHdgRate(const State& st, const Vec3Scalar& ref) const {  // :103-122
    const V3    fd = (ref.vec + v3(0, 0, model.g)) * model.mass;
    TiltHdgRate out;
    out.tilt_vector  = normalized(fd);
    out.heading_rate = ref.s;
    out.throttle     = throttleFor(fd, st);
    return out;
  }
};

struct Pid3Controller {  // VelocityController / PositionController share this shape
  Pid  px, py, pz;
  void init(double kp, double kd, double ki, double sat) {
    px.reset();
    py.reset();
    pz.reset();
    px.setParams(kp, kd, ki, sat, 1.0);
    py.setParams(kp, kd, ki, sat, 1.0);
    pz.setParams(kp, kd, ki, sat, 1.0);
  }
  V3 run(const V3& err, double dt) {
    V3 o;
    o[0] = px.update(err[0], dt);
    o[1] = py.update(err[1], dt);
    o[2] = pz.update(err[2], dt);
    return o;
  }
};

// ------------------------------------------------------------------------------------------
// UavSystem (US)
// ------------------------------------------------------------------------------------------

class UavSystem {
public:
  bool      crashed = false;
  Model     model;
  Mixer     mixer;
  RateController         rate;
  AttitudeController     attitude;
  AccelerationController acceleration;
  Pid3Controller         velocity;  // CTL/velocity_controller.hpp
  Pid3Controller         position;  // CTL/position_controller.hpp
  VelocityParams         velocity_params;
  PositionParams         position_params;
  InputMode              active_input = INPUT_UNKNOWN;

  Actuators    actuators_cmd;
  ControlGroup control_group_cmd;
  AttitudeRate attitude_rate_cmd;
  Attitude     attitude_cmd;
  TiltHdgRate  tilt_hdg_rate_cmd;
  Vec3Scalar   acceleration_hdg_rate_cmd, acceleration_hdg_cmd, velocity_hdg_rate_cmd, velocity_hdg_cmd, position_cmd;

  bool       has_velocity_hdg_rate_ff = false, has_velocity_hdg_ff = false, has_acceleration_hdg_rate_ff = false, has_acceleration_hdg_ff = false;
  Vec3Scalar velocity_hdg_rate_ff, velocity_hdg_ff, acceleration_hdg_rate_ff, acceleration_hdg_ff;

  // UavSystem(params, spawn_pos, spawn_heading)  US:144-153
  UavSystem(const ModelParams& p, const V3& spawn_pos, double spawn_heading) {
    model.params = p;
    model.initializeState();
    model.setStatePos(spawn_pos, spawn_heading);
    initializeControllers();
  }

  void initializeControllers() {  // US:159-169: fresh controllers, default gains, PIDs reset
    const ModelParams mp = model.params;
    mixer.model          = mp;
    mixer.params         = MixerParams();
    mixer.calculateAllocation();
    rate.model  = mp;
    rate.params = RateParams();
    rate.initializePIDs();
    attitude.params = AttitudeParams();
    attitude.initializePIDs();
    acceleration.model = mp;
    velocity_params    = VelocityParams();
    velocity.init(velocity_params.kp, velocity_params.kd, velocity_params.ki, velocity_params.max_acceleration);
    position_params = PositionParams();
    position.init(position_params.kp, position_params.kd, position_params.ki, position_params.max_velocity);
  }

  void setParams(const ModelParams& p) {  // US:404-409
    model.params = p;
    initializeControllers();
  }
  void setMixerParams(const MixerParams& p) {
    mixer.params = p;
    mixer.calculateAllocation();
  }
  void setRateControllerParams(const RateParams& p) {
    rate.params = p;
    rate.initializePIDs();
  }
  void setAttitudeControllerParams(const AttitudeParams& p) {
    attitude.params = p;
    attitude.initializePIDs();
  }
  void setVelocityControllerParams(const VelocityParams& p) {
    velocity_params = p;
    velocity.init(p.kp, p.kd, p.ki, p.max_acceleration);
  }
  void setPositionControllerParams(const PositionParams& p) {
    position_params = p;
    position.init(p.kp, p.kd, p.ki, p.max_velocity);
  }

  void makeStep(double dt) {  // US:304-380
    InputMode active = active_input;

    if (crashed || active_input == INPUT_UNKNOWN) {
      actuators_cmd = Actuators();
    } else {
      if (active == POSITION_CMD) {
        velocity_hdg_cmd.vec = position.run(position_cmd.vec - model.state.x, dt);
        velocity_hdg_cmd.s   = position_cmd.s;
        active               = VELOCITY_HDG_CMD;
        if (has_velocity_hdg_ff) {
          velocity_hdg_cmd.vec = velocity_hdg_cmd.vec + velocity_hdg_ff.vec;
        } else if (has_velocity_hdg_rate_ff) {
          velocity_hdg_cmd.vec = velocity_hdg_cmd.vec + velocity_hdg_rate_ff.vec;
        }
      }

      if (active == VELOCITY_HDG_CMD) {
        acceleration_hdg_cmd.vec = velocity.run(velocity_hdg_cmd.vec - model.state.v, dt);
        acceleration_hdg_cmd.s   = velocity_hdg_cmd.s;
        active                   = ACCELERATION_HDG_CMD;
        if (has_acceleration_hdg_ff) {
          acceleration_hdg_cmd.vec = acceleration_hdg_cmd.vec + acceleration_hdg_ff.vec;
        } else if (has_acceleration_hdg_rate_ff) {
          acceleration_hdg_cmd.vec = acceleration_hdg_cmd.vec + acceleration_hdg_rate_ff.vec;
        }
      } else if (active == VELOCITY_HDG_RATE_CMD) {
        acceleration_hdg_rate_cmd.vec = velocity.run(velocity_hdg_rate_cmd.vec - model.state.v, dt);
        acceleration_hdg_rate_cmd.s   = velocity_hdg_rate_cmd.s;
        active                        = ACCELERATION_HDG_RATE_CMD;
        if (has_acceleration_hdg_rate_ff) {
          acceleration_hdg_rate_cmd.vec = acceleration_hdg_rate_cmd.vec + acceleration_hdg_rate_ff.vec;
          acceleration_hdg_rate_cmd.s += acceleration_hdg_rate_ff.s;
        } else if (has_acceleration_hdg_ff) {
          acceleration_hdg_rate_cmd.vec = acceleration_hdg_rate_cmd.vec + acceleration_hdg_ff.vec;
        }
      }

      if (active == ACCELERATION_HDG_CMD) {
        attitude_cmd = acceleration.getControlSignalHdg(model.state, acceleration_hdg_cmd);
        active       = ATTITUDE_CMD;
      } else if (active == ACCELERATION_HDG_RATE_CMD) {
        tilt_hdg_rate_cmd = acceleration.getControlSignalHdgRate(model.state, acceleration_hdg_rate_cmd);
        active            = TILT_HDG_RATE_CMD;
      }

      if (active == ATTITUDE_CMD) {
        attitude_rate_cmd = attitude.getControlSignal(model.state, attitude_cmd, dt);
        active            = ATTITUDE_RATE_CMD;
      } else if (active == TILT_HDG_RATE_CMD) {
        attitude_rate_cmd = attitude.getControlSignal(model.state, tilt_hdg_rate_cmd, dt);
        active            = ATTITUDE_RATE_CMD;
      }

      if (active == ATTITUDE_RATE_CMD) {
        control_group_cmd = rate.getControlSignal(model.state, attitude_rate_cmd, dt);
        active            = CONTROL_GROUP_CMD;
      }

      if (active == CONTROL_GROUP_CMD) {
        actuators_cmd = mixer.getControlSignal(control_group_cmd);
        active        = ACTUATOR_CMD;
      }
    }

    model.setInput(actuators_cmd);
    model.step(dt);
  }
};

}  // namespace orc
