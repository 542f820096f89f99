#!/usr/bin/env python
"""bench.py — UAV-steps/s of the stepping path (RK4 + controller cascade + collisions).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port +
                                                             # the real vendored nanoflann), all host threads

Workload (BASELINE.json config 4 = the configuration the metric is quoted on):
  1,048,576 x500 UAVs on a 1024 x 1024 grid, 4 m pitch, spawned at z = 0, heading 0; world of
  config/multirotor_simulator.yaml (dt = 0.01 s, g = 9.81, ground plane at z = 0, collisions on,
  crash:false, rebounce 100 as in tmux/standalone_400_uavs); two 0.01 s zero-actuator warm-up steps
  (uav_system_ros.cpp:223-232); seeded VelocityHdgRate commands v_xy~U(-2,2), v_z~U(0,2),
  hdg_rate~U(-1,1) (tmux/standalone_400_uavs/velocity_cmd.py:33-39) from the counter RNG of SURVEY
  §8d; K = 1 (collisions every step).  One "step" = one tick of the reference node's loop
  (multirotor_simulator.cpp:198-231): makeStep for every UAV, then handleCollisions.
  With N GPUs the SAME 1 Mi swarm is sharded by contiguous index ranges (strong scaling) and every
  tick exchanges the packed positions (fused peer stores over NVLink, or an NCCL all-gather).

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes
through the public C ABI with pinned HOST buffers (commands in, positions out, every step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UAVS = 1 << 20
DT = 0.01
SEED = 42
METRIC = "UAV-steps/s (RK4+control+collisions) at 1M UAVs"
UNIT = "UAV-steps/s"
L2_BYTES = 126 * (1 << 20)
# algorithmic HBM bytes per UAV-step of the stepping kernel on this workload (VelocityHdgRate, quad,
# K=1): SURVEY §8d  R = 144 + 8n + P + 24 + 8 + 4 + C,  W = 144 + 8n + P  with n=4, P=144, C=32
STEP_BYTES_PER_UAV = 708
# as-written FP64 census of the reference for this mode
STEP_FLOP_PER_UAV = 2550
COLLIDE_BYTES_PER_UAV = 52


def u01(seed, stream, index):
    with np.errstate(over="ignore"):
        g = np.uint64(0x9E3779B97F4A7C15)
        z = np.uint64(seed) + g * ((np.uint64(stream) << np.uint64(32)) + np.asarray(index, dtype=np.uint64))
        z = z + g
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def workload(begin, count, side=1024):
    """Spawn positions and commands of UAVs [begin, begin+count) of the 1 Mi swarm."""
    k = np.arange(begin, begin + count)
    spawn = np.stack([4.0 * (k % side), 4.0 * (k // side), np.zeros(count)], axis=1).astype(np.float64)
    cmd = np.stack([-2 + 4 * u01(SEED, 1, k), -2 + 4 * u01(SEED, 2, k), 2 * u01(SEED, 3, k), -1 + 2 * u01(SEED, 4, k)], axis=1)
    return spawn, np.ascontiguousarray(cmd)


def x500_world():
    from mrs_multirotor_simulator_b200 import airframe

    return airframe("x500", ground_enabled=True, ground_z=0.0, takeoff_patch_enabled=False, g=9.81)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(power))}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference(n_sample, ticks, warmup, threads):
    """Times `ticks` ticks (makeStep for all + handleCollisions) of the CPU path on a sample of the
    workload: oracle port for UavSystem::makeStep (the reference itself needs Eigen+Boost, absent),
    the REAL vendored nanoflann for handleCollisions.  Returns (UAV-steps/s, description)."""
    from oracle import binding as O

    spawn, cmd = workload(0, n_sample)
    sw = O.OracleSwarm([x500_world()], spawn_xyz=spawn, n=n_sample)
    sw.set_input(O.ACTUATOR_CMD, np.zeros((n_sample, 8)))
    sw.make_step(DT, 2, threads)
    sw.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
    sw.set_collisions(True, False, 100.0)
    engine = "nanoflann" if O.ref_lib() is not None else "port"
    for _ in range(warmup):
        sw.make_step(DT, 1, threads)
        sw.handle_collisions(engine=engine, n_threads=threads, cap=1 << 16)
    t0 = time.perf_counter()
    for _ in range(ticks):
        sw.make_step(DT, 1, threads)
        sw.handle_collisions(engine=engine, n_threads=threads, cap=1 << 16)
    dt = time.perf_counter() - t0
    desc = (f"{n_sample} UAVs (first rows of the 1 Mi grid) x {ticks} ticks, {threads} threads; stepping = oracle port (-O2 -ffp-contract=off), "
            f"collisions = {'real vendored nanoflann (KD-tree build 1 thread, queries threaded)' if engine == 'nanoflann' else 'cell-list port'}")
    desc += "; " + port_vs_reference_sources()
    return n_sample * ticks / dt, dt / ticks * 1e3, desc


_PORT_CHECK = None


def port_vs_reference_sources(n=2048, ticks=20):
    """The timed port against the reference's OWN UavSystem sources compiled against the Eigen/odeint stand-ins
    (oracle/_ref/libref_uavsystem.so, built where /root/reference exists and shipped): same bits, and how fast each is."""
    global _PORT_CHECK
    if _PORT_CHECK is None:
        from oracle import binding as O

        if O.refsys_lib() is None:
            _PORT_CHECK = "reference-sources build (oracle/_ref/libref_uavsystem.so) not shipped: port not re-verified in this run"
        else:
            spawn, cmd = workload(0, n)
            rate = {}
            state = {}
            for name, cls in (("port", O.OracleSwarm), ("reference sources", O.RefSwarm)):
                sw = cls([x500_world()], spawn_xyz=spawn, n=n)
                sw.set_input(O.VELOCITY_HDG_RATE_CMD, cmd)
                t0 = time.perf_counter()
                sw.make_step(DT, ticks, 1)
                rate[name] = n * ticks / (time.perf_counter() - t0)
                state[name] = sw.get_state()
            same = all(np.array_equal(state["port"][k], state["reference sources"][k]) for k in state["port"])
            _PORT_CHECK = (f"port {'bit-identical to' if same else 'DIFFERS from'} the reference's own UavSystem sources compiled against Eigen/odeint stand-ins "
                           f"on {n} UAVs x {ticks} ticks (1 thread: port {rate['port'] / 1e6:.2f} M, reference sources {rate['reference sources'] / 1e6:.2f} M UAV-steps/s; "
                           f"the faster one is the baseline)")
    return _PORT_CHECK


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: the largest slab of the swarm whose (steps + warmup) ticks finish in ~2 minutes
    speed, _, _ = cpu_reference(65536, 2, 1, threads)
    n_sample = 16384
    for cand in (262144, 131072, 65536, 32768):
        if cand * (args.steps + args.warmup) / speed <= 120.0:
            n_sample = cand
            break
    value, ms, desc = cpu_reference(n_sample, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, None),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def collision_report(tick_ms, step_ms, rebuild_ms, info0, info1):
    """The collision pass inside the timed ticks: its average cost is the tick minus the stepping kernel timed alone;
    the spatial hash is rebuilt only on the fraction of passes the device-side displacement bound demands."""
    passes = max(1, info1["passes"] - info0["passes"])
    out = {"ms_avg_in_tick": tick_ms - step_ms, "share_of_tick": (tick_ms - step_ms) / tick_ms, "n_hashed": N_UAVS, "cell_m": info1["cell"],
           "neighbour_lists": info1["neighbour_lists"], "ms_rebuild_pass_alone": rebuild_ms}
    if info1["neighbour_lists"]:
        out.update({"list_radius_m": info1["list_radius"], "skin_m": info1["skin"], "rebuild_fraction": (info1["rebuilds"] - info0["rebuilds"]) / passes,
                    "crowded_uavs_at_last_rebuild": info1["crowded_uavs"]})
    return out


EXCHANGE = {0: "none (single shard)", 1: "NCCL all-gather of packed xyz per tick", 2: "fused: stepping kernel stores positions into all peers over NVLink (CUDA IPC), flag hand-shake per tick"}


def workload_config(n_gpus, l2_note):
    cfg = {"workload": "C4: 1,048,576 x500 UAVs, 1024x1024 grid 4 m pitch, VelocityHdgRate commands, dt=0.01, K=1, ground plane + mutual collisions "
                       "(rebounce 100) every tick", "n_uavs": N_UAVS, "dt": DT, "k_substeps": 1, "collisions": "enabled, crash=false, rebounce=100",
           "sharding": f"{n_gpus} contiguous index shards, packed xyz of the whole swarm exchanged every tick" if n_gpus > 1 else "single shard"}
    if l2_note:
        cfg["l2"] = l2_note
    return cfg


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    from mrs_multirotor_simulator_b200 import ACTUATOR_CMD, VELOCITY_HDG_RATE_CMD, UavBatch, _lib
    from mrs_multirotor_simulator_b200.sharding import connect, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    begin, n_local = shard_range(N_UAVS, world, rank)

    spawn, cmd = workload(begin, n_local)
    batch = UavBatch([x500_world()], spawn_xyz=spawn, n=n_local, device=local, n_global=N_UAVS, shard_begin=begin)
    if world > 1:
        connect(batch, dist)
    batch.set_input(ACTUATOR_CMD, np.zeros((n_local, 8)))
    batch.make_step(DT)
    batch.make_step(DT)
    batch.set_collisions(True, False, 100.0)
    batch.set_input(VELOCITY_HDG_RATE_CMD, cmd)
    batch.sync()

    stream = torch.cuda.ExternalStream(batch.stream, device=torch.device("cuda", local))
    L = _lib.lib()

    # working set per GPU: state the step kernel touches + collision workspace
    working_set = n_local * (STEP_BYTES_PER_UAV // 2 + 200) + N_UAVS * 64
    flush = working_set < 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}") if flush else None
    l2_note = ("L2 flushed between timed steps (256 MiB memset, outside the event pairs)" if flush else
               f"per-GPU working set {working_set / 1e6:.0f} MB > 126 MB L2: inputs larger than L2, no flush")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def tick():
        batch.make_step(DT, 1)
        batch.handle_collisions()

    def timed_loop(fn, steps):
        """K steps timed with CUDA events on the handle's stream; returns total ms (max over ranks).  When the
        per-GPU working set exceeds L2 there is nothing to flush and ONE event pair brackets all K steps;
        otherwise every step has its own pair and the L2 flush runs between the pairs."""
        barrier()
        if not flush:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                fn()
            b.record(stream)
            barrier()
            ms = a.elapsed_time(b)
        else:
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
                a.record(stream)
                fn()
                b.record(stream)
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in ev)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- headline: device-resident ticks ---------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        tick()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    c0 = batch.counters()["launches"]
    info0 = batch.collision_info()
    total_ms = timed_loop(tick, args.steps)
    info1 = batch.collision_info()
    launches = batch.counters()["launches"] - c0
    if sampler:
        clocks = sampler.stop()
    value = N_UAVS * args.steps / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (uav_step_kernel), timed alone -------------------
    step_only = lambda: batch.make_step(DT, 1)
    n_roof = min(args.steps, 200)
    step_ms = timed_loop(step_only, n_roof) / n_roof
    # a pass that is not preceded by exactly one stepping launch rebuilds the spatial hash: this times the rebuild pass
    coll_rebuild_ms = timed_loop(batch.handle_collisions, n_roof) / n_roof

    # ---- e2e: commands from pinned host memory in, positions to host out, every step -------
    import ctypes as C

    cmd_host = torch.from_numpy(cmd).pin_memory()
    pos_host = [torch.empty((n_local, 3), dtype=torch.float64).pin_memory() for _ in range(2)]

    def e2e_loop(n_ticks, pipelined):
        """Every tick: VelocityHdgRate rows H2D, makeStep, handleCollisions, positions D2H — through the C ABI.
        pipelined: the upload of tick t+1 and the download of tick t-1 overlap tick t (mrsb_*_async);
        otherwise the blocking setInput/getState pair."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for t in range(n_ticks):
            if pipelined:
                _lib.check(L.mrsb_set_input_async(batch.h, VELOCITY_HDG_RATE_CMD, C.c_void_p(cmd_host.data_ptr()), 4))
            else:
                _lib.check(L.mrsb_set_input_velocity_hdg_rate(batch.h, n_local, None, C.c_void_p(cmd_host.data_ptr())))
            _lib.check(L.mrsb_make_step(batch.h, DT, 1))
            _lib.check(L.mrsb_handle_collisions(batch.h))
            if pipelined:
                _lib.check(L.mrsb_get_positions_async(batch.h, C.c_void_p(pos_host[t & 1].data_ptr())))
            else:
                _lib.check(L.mrsb_get_state(batch.h, n_local, None, C.c_void_p(pos_host[0].data_ptr()), None, None, None, None))
        batch.sync()  # uploads, compute and downloads have all landed
        e1.record(stream)
        barrier()
        sec = e0.elapsed_time(e1) * 1e-3
        if dist is not None:
            t_ = torch.tensor([sec], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            sec = float(t_.item())
        return sec

    n_e2e = min(args.steps, 200)
    e2e_loop(3, True)
    e2e_s = e2e_loop(n_e2e, True)
    e2e_loop(3, False)
    e2e_blocking_s = e2e_loop(n_e2e, False)
    # the downloaded positions are the simulation's: compare the last snapshot with a blocking read
    check_pos = batch.get_state(fields=("x",))["x"]
    assert np.array_equal(pos_host[0].numpy(), check_pos), "e2e positions differ from a blocking read"
    e2e_value = N_UAVS * n_e2e / e2e_s

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    fp64 = C.c_double()
    copy = C.c_double()
    L.mrsb_microbench_fp64(local, C.byref(fp64))
    L.mrsb_microbench_copy(local, C.byref(copy))
    achieved = n_local * STEP_BYTES_PER_UAV / (step_ms * 1e-3) / 1e9
    roofline = {"kernel": "uav_step_kernel<4, VELOCITY_HDG_RATE>", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                "algorithmic_bytes_per_uav_step": STEP_BYTES_PER_UAV, "launch_ms": step_ms, "uavs_per_launch": n_local,
                "fp64": {"achieved_tflops_as_written_census": n_local * STEP_FLOP_PER_UAV / (step_ms * 1e-3) / 1e12,
                         "peak_tflops_measured_dfma": fp64.value, "flop_per_uav_step_as_written": STEP_FLOP_PER_UAV},
                "copy_gbs_measured_here": copy.value,
                "collision_pass": collision_report(total_ms / args.steps, step_ms, coll_rebuild_ms, info0, info1)}
    fp64_file = os.path.join(ROOT, "profiles", "step_kernel_fp64.json")
    if os.path.exists(fp64_file):  # executed FP64 work of the same kernel (ncu), reported beside the as-written census
        with open(fp64_file) as f:
            ex = json.load(f)["executed_fp64_flop_per_uav_step"]
        tf = n_local * ex / (step_ms * 1e-3) / 1e12
        roofline["fp64"].update({"executed_flop_per_uav_step": ex, "achieved_tflops_executed": tf,
                                 "frac_of_measured_dfma_peak": min(tf, roofline["fp64"]["achieved_tflops_as_written_census"]) / fp64.value})
    traffic_file = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            roofline["traffic"] = json.load(f).get("dram_bytes_per_uav_step") * n_local  # ncu dram read+write per launch, scaled to this shard

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = 262144
        v1, _, _ = cpu_reference(n_sample, 2, 1, threads)
        ticks = int(max(3, min(200, 12.0 * v1 / n_sample)))
        v, _, desc = cpu_reference(n_sample, ticks, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload_config(world, l2_note), exchange=EXCHANGE[batch.exchange_mode()]), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(cmd_host.numel() * 8), "d2h_bytes_per_step": int(pos_host[0].numel() * 8),
                    "steps": n_e2e, "blocking_api_value": N_UAVS * n_e2e / e2e_blocking_s,
                    "h2d_gbs_achieved": cmd_host.numel() * 8 * n_e2e / e2e_s / 1e9, "d2h_gbs_achieved": pos_host[0].numel() * 8 * n_e2e / e2e_s / 1e9,
                    "note": "per rank, every tick, via the C ABI: VelocityHdgRate rows H2D from pinned memory (mrsb_set_input_async), makeStep, "
                            "handleCollisions, positions D2H to pinned memory (mrsb_get_positions_async); upload of tick t+1 and download of tick "
                            "t-1 overlap tick t on separate streams (the tick is then as long as its PCIe upload: see h2d_gbs_achieved); "
                            "blocking_api_value = same with mrsb_set_input + mrsb_get_state"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
