"""CPU-only checks of the host-side logic of the product: parameter derivation, airframe data,
YAML loader, sharding helpers, the command encoders of the UavSystem mirror."""
import ctypes as C

import numpy as np
import pytest

from mrs_multirotor_simulator_b200 import AIRFRAMES, airframe, load_airframe_yaml, model_params
from mrs_multirotor_simulator_b200 import _lib
from mrs_multirotor_simulator_b200.sharding import gather_layout, shard_range
from oracle import binding as O


@pytest.mark.parametrize("name", sorted(AIRFRAMES))
def test_param_derivation_matches_the_reference_formulas(name):
    """J (uav_system_ros.cpp:664-671) and allocation scaling (:98-103): library == oracle, bit for bit."""
    a = model_params(airframe(name))
    b = O.params_from_dict(airframe(name))
    assert list(a.J) == list(b.J)
    assert list(a.allocation_matrix) == list(b.allocation_matrix)
    for f in ("n_motors", "mass", "kf", "km", "prop_radius", "arm_length", "body_height", "motor_time_constant", "max_rpm", "min_rpm", "g"):
        assert getattr(a, f) == getattr(b, f), f


def test_default_model_params_are_the_x500_header_defaults():
    p = _lib.ModelParams()
    _lib.lib().mrsb_model_params_default(C.byref(p))
    q = model_params(airframe("x500", takeoff_patch_enabled=True))
    o = O.OrcModelParams()
    O.lib().orc_model_params_default(C.byref(o))
    assert bytes(p) == bytes(q) == bytes(o)  # multirotor_model.hpp:26-66
    assert p.takeoff_patch_enabled == 1 and p.ground_enabled == 0


@pytest.mark.parametrize("name", sorted(AIRFRAMES))
def test_mixer_allocation_host_vs_oracle(name):
    """Mixer::calculateAllocation (mixer.hpp:72-101): independent implementations agree to rounding."""
    out = np.zeros((8, 4))
    p = model_params(airframe(name))
    _lib.lib().mrsb_mixer_allocation_of(C.byref(p), out.ctypes.data_as(C.c_void_p))
    ref = O.OracleSwarm([airframe(name)], spawn_xyz=[[0, 0, 0]], n=1).get_mixer_allocation()
    n = p.n_motors
    assert np.max(np.abs(out - ref)) < 1e-12
    assert np.allclose(np.linalg.norm(out[:n, :2], axis=1), 1.0) and set(np.unique(out[:n, 2])) <= {-1.0, 0.0, 1.0}
    assert np.all(out[:n, 3] == 1.0) and not out[n:].any()


def test_yaml_loader_reads_the_reference_schema(tmp_path):
    y = tmp_path / "f550.yaml"
    y.write_text("""
f550:
  n_motors: 6
  mass: 2.3
  arm_length: 0.27
  body_height: 0.1
  motor_time_constant: 0.03
  air_resistance_coeff: 0.30
  propulsion:
    force_constant: 0.00000012216
    moment_constant: 0.07
    prop_radius: 0.11
    allocation_matrix: [
      1, -1, -0.5,  0.5,  0.5,   -0.5,
      0, 0,  -0.87, 0.87, -0.87, 0.87,
      1, -1, 1,     -1,   -1,    1,
      1, 1,  1,     1,    1,     1
    ]
    rpm:
      min: 1360
      max: 9068
""")
    assert load_airframe_yaml(str(y)) == AIRFRAMES["f550"]


def test_shard_ranges_tile_the_swarm():
    for n, w in ((1 << 20, 8), (1000, 3), (7, 8), (5, 1)):
        rs = [shard_range(n, w, r) for r in range(w)]
        assert rs[0][0] == 0 and sum(c for _, c in rs) == n
        assert all(rs[r][0] + rs[r][1] == rs[r + 1][0] for r in range(w - 1))
        assert max(c for _, c in rs) - min(c for _, c in rs) <= 1
    assert gather_layout(10, 2) == [(0, 15), (15, 15)]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_command_encoders_follow_references_hpp():
    from mrs_multirotor_simulator_b200 import uav_system as U

    assert U._encode(U.Position(position=np.array([1.0, 2, 3]), heading=0.5)) == (10, [1.0, 2.0, 3.0, 0.5])
    assert U._encode(U.VelocityHdgRate(velocity=np.array([1.0, 2, 3]), heading_rate=0.5)) == (8, [1.0, 2.0, 3.0, 0.5])
    mode, pl = U._encode(U.Attitude(orientation=np.arange(9.0).reshape(3, 3), throttle=0.4))
    assert mode == 4 and pl == [0.0, 3.0, 6.0, 1.0, 4.0, 7.0, 2.0, 5.0, 8.0, 0.4]  # column-major
    assert U._encode(U.TiltHdgRate())[1] == [1.0, 0.0, 0.0, 0.0, 0.0]  # Vector3d::Identity() default (references.hpp:123)
    with pytest.raises(TypeError):
        U._encode(object())


def test_neighbour_list_validity_argument_on_a_numpy_model():
    """The invariant behind collide.cu's neighbour lists, checked on a host model: lists built with radius
    sqrt(3) + skin stay a superset of the true neighbours (d^2 < 3) for as long as twice the SUM of the
    per-tick maximum displacements stays <= skin — whatever the individual UAVs do."""
    rng = np.random.default_rng(3)
    n, skin = 400, 1.2
    r_list = np.sqrt(3.0) + skin
    x = rng.uniform(0, 40, (n, 3)) * [1, 1, 0.2]
    def within(p, r2):
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d2, np.inf)
        return d2 < r2
    lists, D, rebuilds, hits = within(x, r_list ** 2), 0.0, 0, 0
    for tick in range(300):
        step = rng.normal(0, 0.02, (n, 3)) + rng.choice([0.0, 0.1], (n, 1), p=[0.999, 0.001]) * rng.normal(0, 1, (n, 3))
        x = x + step
        D += float(np.sqrt((step ** 2).sum(-1)).max())
        if 2.0 * D > skin:
            lists, D = within(x, r_list ** 2), 0.0
            rebuilds += 1
        true = within(x, 3.0)
        assert not np.any(true & ~lists), tick  # every true neighbour is still listed
        hits += int(true.sum())
    assert 0 < rebuilds < 100 and hits > 0, (rebuilds, hits)


def test_pull_exchange_buffer_protocol_on_a_scheduler_model():
    """Host model of the pull exchange's hazard argument (api.cu `pass_no` / collide.cu `decide_kernel`): every rank owns
    two position buffers; collision pass k READS buffer k & 1 of every peer after the hand-shake of pass k (it has sent its own
    pass number and seen k from every peer), and everything a rank writes between its passes k - 1 and k goes to buffer k & 1.
    Ranks are driven by a random scheduler (any interleaving a set of independent streams could produce, with any number of
    writes between passes); the model checks that no buffer is ever written while a peer reads it, and that a reader always sees
    the owner's latest complete positions of the pass it is in."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        G = int(rng.integers(2, 5))
        n_pass = 6
        # per-rank program: for each pass k = 1..n_pass: some writes (0..3, each bumps a version), then signal(k), wait, read, done
        prog = []
        for r in range(G):
            ops = []
            for k in range(1, n_pass + 1):
                ops += [("write", k)] * int(rng.integers(0, 4)) + [("signal", k), ("wait", k), ("read_begin", k), ("read_end", k)]
            prog.append(ops)
        pc = [0] * G
        flags = [[0] * G for _ in range(G)]      # flags[r][s]: last pass number rank s told rank r
        buf = [[0, 0] for _ in range(G)]         # version stored in buffer b of rank r
        version = [0] * G                        # latest position version of rank r (what `st` holds)
        wrote = [False] * G                      # wrote_since_pass
        reading = [[None] * G for _ in range(G)]  # reading[a][b] = buffer index rank a currently reads of rank b
        at_signal = [dict() for _ in range(G)]    # version rank r held when it entered pass k
        while any(pc[r] < len(prog[r]) for r in range(G)):
            ready = []
            for r in range(G):
                if pc[r] >= len(prog[r]):
                    continue
                op, k = prog[r][pc[r]]
                if op == "wait" and any(flags[r][s] < k for s in range(G) if s != r):
                    continue  # blocked in the hand-shake
                ready.append(r)
            assert ready, "deadlock"
            r = int(rng.choice(ready))
            op, k = prog[r][pc[r]]
            pc[r] += 1
            w = k & 1  # buffer written before pass k and read during pass k
            if op == "write":
                assert all(reading[a][r] != w for a in range(G)), "a peer is reading the buffer being written"
                version[r] += 1
                buf[r][w] = version[r]
                wrote[r] = True
            elif op == "signal":
                if not wrote[r]:  # nothing wrote positions since the last pass: publish_positions into the buffer of this pass
                    assert all(reading[a][r] != w for a in range(G))
                    buf[r][w] = version[r]
                at_signal[r][k] = version[r]
                for s in range(G):
                    if s != r:
                        flags[s][r] = k
            elif op == "read_begin":
                for s in range(G):
                    if s != r:
                        reading[r][s] = w
                        # the owner has signalled pass k, so its buffer k & 1 holds the positions it had when it entered pass k;
                        # it may have moved on since, but only into the OTHER buffer
                        assert flags[r][s] >= k
                        assert buf[s][w] == at_signal[s][k], "the reader does not see the owner's positions of this pass"
            elif op == "read_end":
                for s in range(G):
                    if s != r:
                        assert buf[s][w] == at_signal[s][k], "the buffer changed under the reader"
                    reading[r][s] = None
                wrote[r] = False


def test_bucket_layout_of_mixed_batches():
    """mrsb_bucket_layout (what mrsb_create uses): a batch with 2..8 airframe types is stored sorted stably by type, every type
    padded to whole 128-slot tiles — each tile of the device arrays then holds ONE airframe; callers keep their own indices."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    for n, n_types in ((1, 1), (300, 1), (64, 2), (3000, 3), (5000, 8), (777, 9), (129, 2), (3 * 58001, 3)):
        tou = (np.arange(n) * 7 % n_types).astype(np.int32) if n_types <= 3 else rng.integers(0, n_types, n).astype(np.int32)
        slot = np.full(n, -1, dtype=np.int32)
        first = np.zeros(n_types, dtype=np.int64)
        count = np.zeros(n_types, dtype=np.int64)
        n_slots = C.c_int64(0)
        nb = L.mrsb_bucket_layout(n, n_types, tou.ctypes.data_as(C.c_void_p), slot.ctypes.data_as(C.c_void_p), C.byref(n_slots),
                                  first.ctypes.data_as(C.c_void_p), count.ctypes.data_as(C.c_void_p))
        present = len(np.unique(tou))
        assert n_slots.value % 128 == 0 and n_slots.value >= n
        assert np.array_equal(count, np.bincount(tou, minlength=n_types))
        if present < 2 or present > 8:
            assert nb == 1 and np.array_equal(slot, np.arange(n))  # one airframe, or too many for a launch each: the caller's order
            continue
        assert nb == present
        assert len(np.unique(slot)) == n and slot.min() >= 0 and slot.max() < n_slots.value  # a permutation into the slots
        type_of_slot = np.full(n_slots.value, -1)
        type_of_slot[slot] = tou
        tiles = type_of_slot.reshape(-1, 128)
        for row in tiles:  # every tile: one airframe (plus padding at the end of a bucket)
            assert len(set(row[row >= 0])) <= 1
        for t in range(n_types):
            mine = np.flatnonzero(tou == t)
            if len(mine) == 0:
                assert first[t] == -1
                continue
            assert first[t] % 128 == 0
            assert np.array_equal(slot[mine], first[t] + np.arange(len(mine)))  # stable: the caller's order inside a bucket
    bad = np.array([0, 5], dtype=np.int32)
    assert L.mrsb_bucket_layout(2, 2, bad.ctypes.data_as(C.c_void_p), None, C.byref(C.c_int64(0)), None, None) == -1
