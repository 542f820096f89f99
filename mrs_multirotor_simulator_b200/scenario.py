"""Scenario loader: the reference's YAML configuration files -> a configured `UavBatch`.

The reference node reads its parameters from a stack of YAML files overlaid by the ROS parameter
server — `config/multirotor_simulator.yaml` (world), `config/uavs.yaml` (who flies and where),
`config/uavs/<type>.yaml` (airframes), `config/controllers/*.yaml` (gains), then a scenario's
`custom_configs/simulator.yaml` on top (e.g. `tmux/standalone_400_uavs`) — and each `UavSystemRos`
picks its values out of that tree (`src/uav_system_ros.cpp:30-163`, `src/multirotor_simulator.cpp:
98-160`).  `load_scenario` reproduces the overlay (later files win, key by key) and the parameter
names, so those files run unmodified; whatever a file does not say falls back to the shipped
defaults in `airframes.py` (the numeric content of the reference's own default files).

What is and is not modelled: everything that reaches the stepping path (world, airframes, gains,
spawn poses, collision knobs, rates, input time-out) is; the `frames/*` names and TF switches are
carried along untouched because nothing on the path uses them.  Spawn randomisation
(`randomization/*`, `uav_system_ros.cpp:89-94`) uses the reference's bounds and its +-3.14 heading
range, but draws from the seeded counter RNG of SURVEY §8d instead of `std::rand()`, so a scenario
is reproducible.
"""
from dataclasses import dataclass, field

import numpy as np

from .airframes import AIRFRAMES, CONTROLLER_DEFAULTS, SIMULATOR_DEFAULTS, _af

# config/multirotor_simulator.yaml keys that are not in SIMULATOR_DEFAULTS
_WORLD_EXTRA = dict(clock_rate=100.0, realtime_factor=1.0, iterate_without_input=True, input_timeout=1.0, randomization_enabled=False,
                    randomization_bounds=(15.0, 15.0, 15.0))


def _overlay(base, top):
    """ROS-parameter-server style overlay: dictionaries merge key by key, everything else is replaced."""
    for k, v in top.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _overlay(base[k], v)
        else:
            base[k] = v
    return base


def _u01(seed, stream, index):
    with np.errstate(over="ignore"):
        g = np.uint64(0x9E3779B97F4A7C15)
        z = np.uint64(seed) + g * ((np.uint64(stream) << np.uint64(32)) + np.asarray(index, dtype=np.uint64))
        z = z + g
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _airframe_from_tree(t):
    """One `<type>:` section of the reference's config/uavs/<type>.yaml schema (uav_system_ros.cpp:58-69, 98)."""
    pr = t["propulsion"]
    n = int(t["n_motors"])
    flat = [float(v) for v in pr["allocation_matrix"]]
    if len(flat) != 4 * n:
        raise ValueError(f"allocation_matrix has {len(flat)} entries, expected 4 x {n}")
    alloc = [flat[r * n:(r + 1) * n] for r in range(4)]
    return _af(n, float(t["mass"]), float(t["arm_length"]), float(t["body_height"]), float(t["motor_time_constant"]), float(pr["force_constant"]),
               float(pr["moment_constant"]), float(pr["prop_radius"]), pr["rpm"]["min"], pr["rpm"]["max"], alloc, float(t["air_resistance_coeff"]))


@dataclass
class Scenario:
    """Everything `MultirotorSimulator::onInit` + one `UavSystemRos` per UAV would have loaded."""
    uav_names: list
    type_names: list          # distinct airframe types, in order of first use
    types: list               # airframe dicts (world parameters merged in), parallel to type_names
    type_of_uav: np.ndarray   # int32 [n]
    spawn_xyz: np.ndarray     # float64 [n, 3]
    spawn_heading: np.ndarray  # float64 [n]
    simulation_rate: float
    clock_rate: float
    realtime_factor: float
    iterate_without_input: bool
    input_timeout: float
    collisions_enabled: bool
    collisions_crash: bool
    collisions_rebounce: float
    controllers: dict = field(default_factory=dict)  # keys of airframes.CONTROLLER_DEFAULTS
    frames: dict = field(default_factory=dict)

    @property
    def n(self):
        return len(self.uav_names)

    @property
    def dt(self):
        """simulation_step_size (multirotor_simulator.cpp:108-109)."""
        return 1.0 / self.simulation_rate

    def make_batch(self, device=0, warm_up=True, **kw):
        """The swarm on the GPU, configured as the reference node would have configured its UavSystems:
        per-UAV airframe and spawn pose, controller gains (`uav_system_ros.cpp:109-158`), collision knobs
        (`multirotor_simulator.cpp:124-126`), and — `warm_up` — the two 0.01 s zero-actuator steps every
        UavSystemRos makes before commands arrive (`uav_system_ros.cpp:223-232`)."""
        from .batch import ACTUATOR_CMD, UavBatch

        b = UavBatch(self.types, type_of_uav=self.type_of_uav, spawn_xyz=self.spawn_xyz, spawn_heading=self.spawn_heading, n=self.n, device=device, **kw)
        c = self.controllers
        b.set_controller_params("mixer", [1.0 if c["mixer_desaturation"] else 0.0])
        b.set_controller_params("rate", [c["rate_kp"], c["rate_kd"], c["rate_ki"]])
        b.set_controller_params("attitude", [c["att_kp"], c["att_kd"], c["att_ki"], c["att_max_rate_roll_pitch"], c["att_max_rate_yaw"]])
        b.set_controller_params("velocity", [c["vel_kp"], c["vel_kd"], c["vel_ki"], c["vel_max_acceleration"]])
        b.set_controller_params("position", [c["pos_kp"], c["pos_kd"], c["pos_ki"], c["pos_max_velocity"]])
        if warm_up:
            b.set_input(ACTUATOR_CMD, np.zeros((self.n, 8)))
            b.make_step(0.01)
            b.make_step(0.01)
        b.set_collisions(self.collisions_enabled, self.collisions_crash, self.collisions_rebounce)
        b.set_iterate_without_input(self.iterate_without_input)  # uav_system_ros.cpp:265
        return b


def scenario_from_tree(tree, seed=42):
    """Build a Scenario from an already overlaid parameter tree (nested dicts, the YAML documents' shape)."""
    w = dict(SIMULATOR_DEFAULTS)
    w.update(_WORLD_EXTRA)
    g = float(tree.get("g", w["g"]))
    ground = tree.get("ground", {})
    ground_enabled = bool(ground.get("enabled", w["ground_enabled"]))
    ground_z = float(ground.get("z", w["ground_z"]))
    patch = bool(tree.get("individual_takeoff_platform", {}).get("enabled", w["takeoff_patch_enabled"]))
    coll = tree.get("collisions", {})
    rnd = tree.get("randomization", {})
    bounds = rnd.get("bounds", {})
    rb = (float(bounds.get("x", w["randomization_bounds"][0])), float(bounds.get("y", w["randomization_bounds"][1])),
          float(bounds.get("z", w["randomization_bounds"][2])))

    names = list(tree.get("uav_names", []))
    if not names:
        raise ValueError("the configuration names no UAVs (uav_names)")
    type_names, types, tou = [], [], []
    xyz = np.zeros((len(names), 3))
    hdg = np.zeros(len(names))
    for i, name in enumerate(names):
        if name not in tree:
            raise ValueError(f"uav_names lists '{name}' but there is no '{name}:' section")  # param_loader.loadedSuccessfully() (uav_system_ros.cpp:160)
        sec = tree[name]
        tname = str(sec["type"])
        if tname not in type_names:
            if isinstance(tree.get(tname), dict) and "n_motors" in tree[tname]:
                af = _airframe_from_tree(tree[tname])
            elif tname in AIRFRAMES:
                af = dict(AIRFRAMES[tname])
                af["allocation"] = [list(r) for r in af["allocation"]]
            else:
                raise ValueError(f"unknown UAV type '{tname}' (no '{tname}:' section and not a shipped airframe)")
            af.update(g=g, ground_enabled=ground_enabled, ground_z=ground_z, takeoff_patch_enabled=patch)
            type_names.append(tname)
            types.append(af)
        tou.append(type_names.index(tname))
        sp = sec["spawn"]
        xyz[i] = (float(sp["x"]), float(sp["y"]), float(sp["z"]))
        hdg[i] = float(sp["heading"])
    if bool(rnd.get("enabled", w["randomization_enabled"])):  # uav_system_ros.cpp:89-94
        k = np.arange(len(names))
        for c in range(3):
            xyz[:, c] += -rb[c] + 2.0 * rb[c] * _u01(seed, c, k)
        hdg += -3.14 + 6.28 * _u01(seed, 3, k)

    ctl = dict(CONTROLLER_DEFAULTS)
    for sect, prefix, keys in (("rate_controller", "rate_", ("kp", "kd", "ki")), ("attitude_controller", "att_", ("kp", "kd", "ki", "max_rate_roll_pitch", "max_rate_yaw")),
                               ("velocity_controller", "vel_", ("kp", "kd", "ki", "max_acceleration")), ("position_controller", "pos_", ("kp", "kd", "ki", "max_velocity"))):
        for key in keys:
            if key in tree.get(sect, {}):
                ctl[prefix + key] = float(tree[sect][key])
    if "desaturation" in tree.get("mixer", {}):
        ctl["mixer_desaturation"] = bool(tree["mixer"]["desaturation"])

    return Scenario(uav_names=names, type_names=type_names, types=types, type_of_uav=np.asarray(tou, dtype=np.int32), spawn_xyz=xyz, spawn_heading=hdg,
                    simulation_rate=float(tree.get("simulation_rate", w["simulation_rate"])), clock_rate=float(tree.get("clock_rate", w["clock_rate"])),
                    realtime_factor=float(tree.get("realtime_factor", w["realtime_factor"])),
                    iterate_without_input=bool(tree.get("iterate_without_input", w["iterate_without_input"])),
                    input_timeout=float(tree.get("input_timeout", w["input_timeout"])),
                    collisions_enabled=bool(coll.get("enabled", w["collisions_enabled"])), collisions_crash=bool(coll.get("crash", w["collisions_crash"])),
                    collisions_rebounce=float(coll.get("rebounce", w["collisions_rebounce"])), controllers=ctl, frames=dict(tree.get("frames", {})))


def load_scenario(*yaml_paths, seed=42):
    """Overlay the given YAML files in order (later files win, like `rosparam load` in the reference's
    launch file) and build the Scenario.  A scenario's custom config alone is enough: the reference's
    default files are represented by the shipped defaults."""
    import yaml

    tree = {}
    for path in yaml_paths:
        with open(path) as f:
            doc = yaml.safe_load(f)
        if doc:
            _overlay(tree, doc)
    return scenario_from_tree(tree, seed=seed)
