"""Airframe and controller-gain data of the reference's shipped configuration files.

Data only (the numeric content of config/uavs/*.yaml, config/controllers/*.yaml and
config/multirotor_simulator.yaml of the reference); ``load_airframe_yaml`` reads the same YAML
schema from a user-supplied file so reference scenario files run unmodified.

`allocation` is the UNSCALED 4 x n_motors matrix exactly as written in the YAML; the scaling
(rows *= arm*kf, arm*kf, km*3*prop_radius*kf, kf — src/uav_system_ros.cpp:98-103) and the inertia
(src/uav_system_ros.cpp:664-671) are applied when the dict is converted to an mrsb_model_params.
"""

_QUAD = [[-0.707, 0.707, 0.707, -0.707], [-0.707, 0.707, -0.707, 0.707], [-1, -1, 1, 1], [1, 1, 1, 1]]
_HEXA = [[1, -1, -0.5, 0.5, 0.5, -0.5], [0, 0, -0.87, 0.87, -0.87, 0.87], [1, -1, 1, -1, -1, 1], [1, 1, 1, 1, 1, 1]]
_OCTA = [[-0.707, 0.707, 0.707, -0.707, 0.707, -0.707, -0.707, 0.707], [-0.707, 0.707, -0.707, 0.707, 0.707, -0.707, 0.707, -0.707],
         [-1, -1, 1, 1, 1, 1, -1, -1], [1, 1, 1, 1, 1, 1, 1, 1]]


def _af(n, mass, arm, body_h, tau, kf, km, prop_r, rpm_min, rpm_max, alloc, air=0.30):
    return dict(n_motors=n, mass=mass, arm_length=arm, body_height=body_h, motor_time_constant=tau, air_resistance_coeff=air, kf=kf, km=km,
                prop_radius=prop_r, min_rpm=float(rpm_min), max_rpm=float(rpm_max), allocation=[list(map(float, r)) for r in alloc])


# config/uavs/{x500,f330,f450,t650,a300,robofly,f550,naki}.yaml
AIRFRAMES = {
    "x500": _af(4, 2.0, 0.25, 0.10, 0.03, 0.00000027087, 0.07, 0.15, 1170, 7800, _QUAD),
    "f330": _af(4, 1.4, 0.165, 0.07, 0.03, 0.000000094268, 0.07, 0.09, 1459, 9722, _QUAD),
    "f450": _af(4, 1.7, 0.225, 0.10, 0.03, 0.00000012216, 0.07, 0.11, 1360, 9068, _QUAD),
    "t650": _af(4, 3.5, 0.325, 0.15, 0.03, 0.00000073385, 0.07, 0.19, 875, 5832, _QUAD),
    "a300": _af(4, 1.21, 0.15, 0.05, 0.05, 0.000000045, 0.012, 0.089, 3200, 21400, _QUAD),
    "robofly": _af(4, 0.8, 0.135, 0.07, 0.03, 0.00000000843, 0.012, 0.09, 2058, 41160, _QUAD),
    "f550": _af(6, 2.3, 0.27, 0.10, 0.03, 0.00000012216, 0.07, 0.11, 1360, 9068, _HEXA),
    "naki": _af(8, 7.5, 0.20, 0.20, 0.03, 0.00000057658, 0.07, 0.13, 956, 6376, _OCTA),
}

# config/multirotor_simulator.yaml
SIMULATOR_DEFAULTS = dict(simulation_rate=100.0, g=9.81, ground_enabled=True, ground_z=0.0, takeoff_patch_enabled=False,
                          collisions_enabled=True, collisions_crash=True, collisions_rebounce=100.0)

# config/controllers/*.yaml (== the header defaults of the controller Params classes)
CONTROLLER_DEFAULTS = dict(mixer_desaturation=True, rate_kp=4.0, rate_kd=0.04, rate_ki=0.0, att_kp=6.0, att_kd=0.05, att_ki=0.01,
                           att_max_rate_roll_pitch=10.0, att_max_rate_yaw=1.0, vel_kp=2.0, vel_kd=0.05, vel_ki=0.01, vel_max_acceleration=4.0,
                           pos_kp=2.0, pos_kd=0.15, pos_ki=0.2, pos_max_velocity=6.0)


def airframe(name, **overrides):
    """Copy of a shipped airframe with world parameters merged in (g, ground_*, takeoff_patch_enabled)."""
    d = dict(AIRFRAMES[name])
    d["allocation"] = [list(r) for r in d["allocation"]]
    d.setdefault("g", SIMULATOR_DEFAULTS["g"])
    d.setdefault("ground_enabled", False)
    d.setdefault("ground_z", 0.0)
    d.setdefault("takeoff_patch_enabled", False)
    d.update(overrides)
    return d


def load_airframe_yaml(path, type_name=None):
    """Read one airframe from a YAML file that follows the reference's config/uavs/<type>.yaml schema."""
    import yaml

    with open(path) as f:
        doc = yaml.safe_load(f)
    if type_name is None:
        type_name = next(iter(doc))
    t = doc[type_name]
    pr = t["propulsion"]
    n = int(t["n_motors"])
    flat = [float(v) for v in pr["allocation_matrix"]]
    alloc = [flat[r * n:(r + 1) * n] for r in range(4)]
    return _af(n, float(t["mass"]), float(t["arm_length"]), float(t["body_height"]), float(t["motor_time_constant"]), float(pr["force_constant"]),
               float(pr["moment_constant"]), float(pr["prop_radius"]), pr["rpm"]["min"], pr["rpm"]["max"], alloc, float(t["air_resistance_coeff"]))
