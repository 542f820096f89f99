"""B200-native batched multirotor stepping: the hot path of ctu-mrs/mrs_multirotor_simulator
(UavSystem::makeStep for N UAVs + MultirotorSimulator::handleCollisions) as sm_100a CUDA kernels
behind a C ABI (include/mrsb.h, libmrsb.so).  This package is the thin host-side mirror of the
reference's UavSystem interface; all arithmetic happens on the GPU and there is no CPU fallback.
"""
from .airframes import AIRFRAMES, CONTROLLER_DEFAULTS, SIMULATOR_DEFAULTS, airframe, load_airframe_yaml  # noqa: F401
from .batch import (ACCELERATION_HDG_CMD, ACCELERATION_HDG_RATE_CMD, ACTUATOR_CMD, ATTITUDE_CMD, ATTITUDE_RATE_CMD,  # noqa: F401
                    CONTROL_GROUP_CMD, INPUT_UNKNOWN, POSITION_CMD, STRIDE, TILT_HDG_RATE_CMD, VELOCITY_HDG_CMD,
                    VELOCITY_HDG_RATE_CMD, UavBatch, model_params)
from .scenario import Scenario, load_scenario, scenario_from_tree  # noqa: F401
from .uav_system import UavSystem  # noqa: F401
