// oracle/shim/odeint_shim.hpp — NOT Boost.odeint.  TEST INFRASTRUCTURE ONLY.
//
// Pre-included (`g++ -include`) when the reference's own uav_system.hpp is compiled for
// oracle/_ref/libref_uavsystem.so.  The reference vendors odeint under
// include/mrs_multirotor_simulator/uav_system/ode/, but that tree needs the system Boost headers
// (mpl, fusion, range, utility ...) which this image does not have.  Defining the vendored forward
// header's include guard makes `#include "ode/boost/numeric/odeint.hpp"` (multirotor_model.hpp:13)
// a no-op; the three names the reference uses are supplied here:
//   boost::numeric::odeint::runge_kutta4<State>   multirotor_model.hpp:224
//   boost::numeric::odeint::integrate_n_steps     multirotor_model.hpp:226
//   boost::ref                                    multirotor_model.hpp:226
// The stepper restates the arithmetic of the vendored generic Runge-Kutta algorithm (paths below
// .../ode/boost/numeric/odeint/):
//   stepper/runge_kutta4.hpp:40-95                 Butcher tableau a = {1/2},{0,1/2},{0,0,1}, b = {1/6,1/3,1/3,1/6}
//   stepper/base/explicit_stepper_base.hpp         do_step(sys, x, t, dt): sys(x, dxdt, t); do_step_impl(sys, x, dxdt, t, x, dt)
//   stepper/detail/generic_rk_algorithm.hpp:172-215  stage s>1 evaluates sys(x_tmp, F[s-2], t + c*dt); then x_tmp (or out)
//                                                   = scale_sum{s+1}(1, a[0]*dt, ..., a[s-1]*dt) over (x, dxdt, F[0..])
//   stepper/detail/generic_rk_operations.hpp:28-60   coefficient products a[i]*dt, leading factor 1.0
//   algebra/default_operations.hpp:77-154            t1 = a1*t2 + a2*t3 + a3*t4 (+ ...), left to right, zero terms included
//   integrate/detail/integrate_n_steps.hpp           stepper_tag version: n calls of do_step with time = start + i*dt
#ifndef BOOST_NUMERIC_ODEINT_HPP_INCLUDED
#define BOOST_NUMERIC_ODEINT_HPP_INCLUDED

#include <cstddef>

namespace boost {

template <class T>
class reference_wrapper {
 public:
  explicit reference_wrapper(T& t) : p_(&t) {}
  T& get() const { return *p_; }
  operator T&() const { return *p_; }

 private:
  T* p_;
};
template <class T>
reference_wrapper<T> ref(T& t) {
  return reference_wrapper<T>(t);
}

namespace numeric {
namespace odeint {

template <class T>
T& unwrap(T& t) {
  return t;
}
template <class T>
T& unwrap(reference_wrapper<T>& w) {
  return w.get();
}

template <class State, class Value = double, class Deriv = State, class Time = Value>
class runge_kutta4 {
 public:
  template <class System>
  void do_step(System system, State& x, Time t, Time dt) {
    auto& sys = unwrap(system);
    const Value a21 = (static_cast<Value>(1) / static_cast<Value>(2)) * dt;
    const Value a31 = static_cast<Value>(0) * dt, a32 = (static_cast<Value>(1) / static_cast<Value>(2)) * dt;
    const Value a41 = static_cast<Value>(0) * dt, a42 = static_cast<Value>(0) * dt, a43 = static_cast<Value>(1) * dt;
    const Value b1 = (static_cast<Value>(1) / static_cast<Value>(6)) * dt, b2 = (static_cast<Value>(1) / static_cast<Value>(3)) * dt;
    const Value b3 = (static_cast<Value>(1) / static_cast<Value>(3)) * dt, b4 = (static_cast<Value>(1) / static_cast<Value>(6)) * dt;
    const Value c2 = static_cast<Value>(1) / static_cast<Value>(2), c3 = c2, c4 = static_cast<Value>(1);
    const std::size_t n = x.size();

    sys(x, dxdt_, t);
    for (std::size_t i = 0; i < n; i++) tmp_[i] = 1.0 * x[i] + a21 * dxdt_[i];
    sys(tmp_, f_[0], t + c2 * dt);
    for (std::size_t i = 0; i < n; i++) tmp_[i] = 1.0 * x[i] + a31 * dxdt_[i] + a32 * f_[0][i];
    sys(tmp_, f_[1], t + c3 * dt);
    for (std::size_t i = 0; i < n; i++) tmp_[i] = 1.0 * x[i] + a41 * dxdt_[i] + a42 * f_[0][i] + a43 * f_[1][i];
    sys(tmp_, f_[2], t + c4 * dt);
    for (std::size_t i = 0; i < n; i++) x[i] = 1.0 * x[i] + b1 * dxdt_[i] + b2 * f_[0][i] + b3 * f_[1][i] + b4 * f_[2][i];
  }

 private:
  Deriv dxdt_;
  State tmp_;
  Deriv f_[3];
};

template <class Stepper, class System, class State, class Time>
Time integrate_n_steps(Stepper stepper, System system, State& start_state, Time start_time, Time dt, std::size_t num_of_steps) {
  Time time = start_time;
  for (std::size_t step = 0; step < num_of_steps; ++step) {
    stepper.do_step(system, start_state, time, dt);
    time = start_time + static_cast<Time>(step + 1) * dt;
  }
  return time;
}

}  // namespace odeint
}  // namespace numeric
}  // namespace boost

#endif
