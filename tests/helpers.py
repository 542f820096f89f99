"""Shared helpers for the parity tests: build the SAME swarm in the CPU oracle and in libmrsb,
drive both with the same seeded commands, compare state within the stated tolerances.

Tolerances: after 10 s of simulated flight
    |dx| <= 1e-9 m, |dv| <= 1e-9 m/s, |dR| <= 1e-10, |domega| <= 1e-8 rad/s, |drpm| <= 1e-6 RPM.
"""
import numpy as np

from oracle import binding as O

TOL = {"x": 1e-9, "v": 1e-9, "R": 1e-10, "omega": 1e-8, "motor_rpm": 1e-6, "v_prev": 1e-9, "imu": 1e-6}


def grid_spawn(n, pitch=4.0, z=0.0):
    side = int(np.ceil(np.sqrt(n)))
    i = np.arange(n)
    return np.stack([pitch * (i % side), pitch * (i // side), np.full(n, z)], axis=1).astype(np.float64)


def rand(seed, stream, n, lo, hi):
    return lo + (hi - lo) * O.u01(seed, stream, np.arange(n))


def make_pair(types, type_of_uav, spawn_xyz, spawn_heading=None, device=0):
    from mrs_multirotor_simulator_b200 import UavBatch

    n = len(spawn_xyz)
    if spawn_heading is None:
        spawn_heading = np.zeros(n)
    orc = O.OracleSwarm(types, type_of_uav=type_of_uav, spawn_xyz=spawn_xyz, spawn_heading=spawn_heading, n=n)
    gpu = UavBatch(types, type_of_uav=type_of_uav, spawn_xyz=spawn_xyz, spawn_heading=spawn_heading, n=n, device=device)
    return orc, gpu


def max_diffs(orc, gpu, fields=("x", "v", "R", "omega", "motor_rpm", "v_prev", "imu")):
    so = orc.get_state()
    sg = gpu.get_full_state()
    return {k: float(np.max(np.abs(so[k] - sg[k]))) if so[k].size else 0.0 for k in fields}


def assert_parity(orc, gpu, tol=None, scale=1.0, what=""):
    tol = tol or TOL
    so = orc.get_state()
    sg = gpu.get_full_state()
    bad = []
    for k, t in tol.items():
        assert np.all(np.isfinite(so[k])), f"oracle {k} not finite {what}"
        d = float(np.max(np.abs(so[k] - sg[k]))) if so[k].size else 0.0
        if not d <= t * scale:
            bad.append(f"{k}: {d:.3e} > {t * scale:.1e}")
    assert not bad, f"parity {what}: " + "; ".join(bad)
