// C++ host driving libmrsb through the façade: BASELINE config 1 (one x500, PositionCmd
// hover-to-waypoint, RK4 dt = 0.005 s, 10 s) plus a two-UAV collision.  Prints the end state as
// JSON; tests/test_cpp_facade.py compares it with the golden fixture.
#include <cstdio>

#include "mrsb/uav_system.hpp"

int main() {
  try {
    mrsb_model_params x500 = mrsb::defaultModelParams();
    x500.takeoff_patch_enabled = 0;
    mrsb::Swarm     swarm({x500}, {}, {{0.0, 0.0, 1.0}, {50.0, 0.0, 1.0}, {50.3, 0.0, 1.0}}, {0.0, 0.0, 0.0});
    mrsb::UavSystem uav = swarm[0];
    mrsb::reference::Position cmd;
    cmd.position = {5.0, -3.0, 4.0};
    cmd.heading  = 1.0;
    uav.setInput(cmd);
    swarm.setCollisions(true, true, 100.0);
    for (int k = 0; k < 2000; k++) {
      swarm.makeStep(0.005);
      if (k == 0) swarm.handleCollisions();
    }
    const mrsb::State s = uav.getState();
    std::printf("{\"x\": [%.17g, %.17g, %.17g], \"v\": [%.17g, %.17g, %.17g], \"rpm\": [%.17g, %.17g, %.17g, %.17g], \"n_motors\": %zu, "
                "\"crashed\": [%d, %d, %d], \"pairs\": %zu}\n",
                s.x[0], s.x[1], s.x[2], s.v[0], s.v[1], s.v[2], s.motor_rpm[0], s.motor_rpm[1], s.motor_rpm[2], s.motor_rpm[3], s.motor_rpm.size(),
                int(swarm[0].hasCrashed()), int(swarm[1].hasCrashed()), int(swarm[2].hasCrashed()), swarm.collisionPairs().size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
}
