// C++ host driving libmrsb through the façade the way the reference's node does (multirotor_simulator.cpp:198-231): 256 x500 on a
// 2 m grid, collisions with rebounce, `iterate_without_input: false` (uav_system_ros.cpp:265) with every second UAV left without a
// command, 200 ticks through Swarm::run (one graph launch per tick).  Prints a few facts as JSON; tests/test_cpp_facade.py compares
// them with the same flight driven through the Python mirror.
#include <cstdio>
#include <vector>

#include "mrsb/uav_system.hpp"

int main() {
  try {
    const int         n    = 256;
    mrsb_model_params x500 = mrsb::defaultModelParams();
    x500.takeoff_patch_enabled = 0;
    x500.ground_enabled        = 1;
    x500.ground_z              = 0.0;
    std::vector<std::array<double, 3>> spawn;
    std::vector<double>                heading(n, 0.0);
    for (int i = 0; i < n; i++) spawn.push_back({2.0 * (i % 16), 2.0 * (i / 16), 3.0});
    mrsb::Swarm swarm({x500}, {}, spawn, heading);
    swarm.setIterateWithoutInput(false);
    swarm.setOutputs(true, true);
    swarm.setCollisions(true, false, 100.0);
    for (int i = 0; i < n; i += 2) {
      mrsb::reference::VelocityHdgRate cmd;
      cmd.velocity     = {0.5 * ((i % 7) - 3), 0.4 * ((i % 5) - 2), 0.1 * (i % 3)};
      cmd.heading_rate = 0.2;
      swarm[i].setInput(cmd);
    }
    swarm.run(0.01, 200);
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
      const mrsb::State s = swarm[i].getState();
      sum += s.x[0] + 2.0 * s.x[1] + 3.0 * s.x[2];
    }
    const mrsb::State idle = swarm[1].getState();
    std::printf("{\"sum\": %.17g, \"idle_z\": %.17g, \"idle_rpm0\": %.17g, \"pairs\": %zu}\n", sum, idle.x[2], idle.motor_rpm[0], swarm.collisionPairs().size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 3;
  }
}
