#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing hot path.

Metric (BASELINE.json): Msamples/s at 1920x1080x1024 spp, 8 bounces, path sampler, on the
1M-triangle instanced scene. One *step* = one trace_samples call that
advances every pixel by `--spp-per-step` samples (default 128, so 8 steps = one 1024-spp image;
cost per sample does not depend on the batch size, only the per-batch tail does, which makes
small batches conservative). With --gpus N the image is row-tiled over N ranks (one process per
GPU, replicated scene) and finished images are combined with ONE ncclAllGather inside the timed
region.

  python bench.py                                   # N=1, 8 steps, 3 warm-up steps
  torchrun --nproc-per-node N bench.py --gpus N     # tile-parallel
  python bench.py --impl reference                  # the reference CPU renderer on the host cores

Output: one JSON line (see the keys in `main`). `roofline` is for the dominant kernel (k_extend,
closest-hit traversal): algorithmic bytes (SURVEY.md §8d formula with traversal counters measured
on this workload by the kernel's counting variant, untimed) / CUDA-event time of its launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in ("yocto-gl_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, _p))

import numpy as np  # noqa: E402

CPU_BASELINE = None

WORKLOADS = {
    # name: (scene factory, resolution, spp, bounces, description)
    "c3": ("instanced_spheres", dict(n=10), 1920, 1024, 8,
           "C3: 1000x 1024-tri sphere instances + floor + area light + env (1,024,004 instanced tris), "
           "1920x1080, 1024 spp, 8 bounces, path"),
    "c1": ("cornellbox", {}, 256, 16, 4, "C1: Cornell box 256x256, 16 spp, 4 bounces, path"),
    "c2": ("bunny_like", dict(subdiv=6), 1280, 256, 8,
           "C2: 81,920-tri blob (bunny stand-in) + env map, 1280x720, 256 spp, 8 bounces, path"),
    "c5": ("hair_scene", {}, 1920, 512, 12,
           "C5: 524,288 line segments + 83,970 tris, glossy/subsurface/refractive, 1920x1080, 512 spp, 12 bounces"),
}

# bytes per element in the reference layout
B_NODE, B_INST, B_RAY_IO = 32, 56, 32 + 24
B_PRIM = dict(triangle_tests=48, quad_tests=64, line_tests=40, point_tests=20)


def algorithmic_bytes_per_ray(c):
    rays = max(1, c["scene_rays"])
    total = B_NODE * (c["top_nodes"] + c["bottom_nodes"]) + B_INST * c["instance_visits"]
    total += sum(B_PRIM[k] * c[k] for k in B_PRIM)
    return total / rays + B_RAY_IO


class ClockSampler:
    """SM clock and throttle reasons of ONE GPU sampled at 5 Hz during the timed region, on rank 0 only, through
    NVML in-process (no process is forked inside the timed region); falls back to polling nvidia-smi if the NVML
    binding is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, enabled=True):
        self.rows, self.stop, self.index, self.enabled = [], False, index, enabled
        self.nvml = self.handle = None
        self.sm_max = None
        if enabled:
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nvml = pynvml
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
                self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            except Exception:
                self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        try:
            power = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
        except Exception:
            power = None
        return [sm, self.sm_max, power] + [bool(mask & b) for b in bits]

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        return [float(r[0]), float(r[1]), float(r[2])] + [x.lower().startswith("active") for x in r[3:7]]

    def run(self):
        while not self.stop:
            try:
                self.rows.append(self.sample_nvml() if self.nvml else self.sample_smi())
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        if self.enabled:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.enabled:
            self.t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[3 + k] for r in self.rows)]
        power = [r[2] for r in self.rows if r[2] is not None]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(power) if power else None,
                "source": "NVML in-process, rank 0 only, 5 Hz" if self.nvml else "nvidia-smi polled, rank 0 only"}


def host_cpu_info():
    """What the CPU arm really had: logical CPUs, the affinity mask and cgroup quota of THIS process, CPU model."""
    info = {"logical_cpus": os.cpu_count()}
    try:
        info["affinity_cpus"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            info["cgroup_cpu_max"] = open(path).read().strip()
            break
        except Exception:
            pass
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                info["model"] = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return info


def scene_bytes(scene):
    n = 0
    for s in scene.shapes:
        n += sum(a.nbytes for a in s.values())
    n += sum(t["pixels"].nbytes for t in scene.textures)
    n += 56 * len(scene.instances) + 84 * len(scene.materials) + 72 * len(scene.cameras) + 64 * len(scene.environments)
    return n


def make_workload(name):
    from ygl_b200 import abi, scenes
    factory, kw, res, spp, bounces, desc = WORKLOADS[name]
    scene = getattr(scenes, factory)(**kw)
    params = abi.trace_params(resolution=res, samples=spp, bounces=bounces)
    return scene, params, desc


def run_reference(args):
    """The reference's own CPU renderer (oracle/_ref, the unmodified sources compiled here) on the
    host cores: same scene_data, same params; each step is a bounded spp sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refbind
    scene, params, desc = make_workload(args.workload)
    kind = "reference"
    if not refbind.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libyocto_ref.so not built"}))
        return
    rs = refbind.Ref().scene(scene)
    threads = rs.lib.ref_hardware_concurrency()  # what the reference's parallel_for starts (yocto_trace.cpp:58)
    t1 = rs.trace_image(params, samples=1)["seconds"]
    spp_step = int(max(1, min(64, 2.5 / max(t1, 1e-3))))
    w = h = 0
    times = []
    cpu0, wall0 = None, None
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            cpu0, wall0 = sum(os.times()[:2]), time.perf_counter()
        out = rs.trace_image(params, samples=spp_step)
        w, h = out["width"], out["height"]
        if i >= args.warmup:
            times.append(out["seconds"])
    busy = (sum(os.times()[:2]) - cpu0) / max(time.perf_counter() - wall0, 1e-9)  # mean runnable threads that got a CPU
    total = sum(times)
    value = w * h * spp_step * len(times) / total / 1e6
    # one thread (params.noparallel), on a quarter-resolution frame of the same scene: the per-core rate
    p1 = type(params).from_buffer_copy(params)
    p1.noparallel, p1.resolution = 1, max(64, params.resolution // 4)
    o1 = rs.trace_image(p1, samples=1)
    one_thread = o1["width"] * o1["height"] / max(o1["seconds"], 1e-9) / 1e6
    sample = f"{spp_step} spp per step of the {w}x{h} frame (cost is linear in spp), trace_samples loop only"
    host = host_cpu_info()
    print(json.dumps({
        "impl": "reference",
        "metric": "Msamples/s (rays shaded/s) at 1920x1080x1024spp, 8 bounces" if args.workload == "c3" else "Msamples/s",
        "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "spp_per_step": spp_step},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": sample,
                         "threads_started": threads, "busy_cores_measured": round(busy, 1),
                         "one_thread_value": one_thread,
                         "one_thread_sample": f"1 spp of {o1['width']}x{o1['height']}, params.noparallel",
                         "host": host},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=0, help="samples per pixel per step (default spp/8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    global CPU_BASELINE
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # The reference CPU renderer on this box's host cores, in its own process and BEFORE any CUDA
        # work (a clean affinity mask: inside this process its thread pool was seen pinned to one core).
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                  "--warmup", "1", "--workload", args.workload], capture_output=True, text=True,
                                 timeout=600, env={k: v for k, v in os.environ.items() if not k.startswith("OMP_")})
            CPU_BASELINE = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # the baseline is reported, never required
            CPU_BASELINE = {"error": str(e)}

    import torch
    from ygl_b200 import abi, lib

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, params, desc = make_workload(args.workload)
    spp_total = params.samples
    spp_step = args.spp_per_step or max(1, spp_total // 8)
    params.batch = spp_step

    ctx = lib.Context(local)
    ds = lib.DeviceScene(ctx, scene)
    W, H = lib.state_size(scene, params)
    rows = lib.tile_rows(H, rank, world)
    if world > 1:
        idb = torch.zeros(lib.load().ygl_comm_id_size(), dtype=torch.uint8, device="cuda")
        if rank == 0:
            blob = np.zeros(idb.numel(), np.uint8)
            lib._check(lib.load().ygl_comm_create_id(blob.ctypes.data))
            idb.copy_(torch.from_numpy(blob))
        dist.broadcast(idb, 0)
        blob = idb.cpu().numpy()
        lib._check(lib.load().ygl_comm_init(ctx.h, blob.ctypes.data, rank, world))

    def mk_state(p):
        # N > 1: interleaved rows (row j -> rank j % N) balance sky against geometry; values are unaffected
        return ds.make_state(p, interleave=(rank, world)) if world > 1 else ds.make_state(p)

    # ---- untimed counting pass: traversal statistics of THIS workload for the roofline ----
    ctx.set_profiling(False, True)
    cp = abi.trace_params(resolution=params.resolution, samples=2, bounces=params.bounces, batch=2)
    cstate = mk_state(cp)
    ds.trace_samples(cstate, cp)
    cc = ctx.counters()
    bytes_per_ray = algorithmic_bytes_per_ray(cc)
    del cstate

    # ---- timed steps ----
    ctx.set_profiling(True, False)
    n_total = args.warmup + args.steps
    steps_per_image = (spp_total + spp_step - 1) // spp_step
    n_states = (n_total + steps_per_image - 1) // steps_per_image
    states = [mk_state(params) for _ in range(n_states)]
    stream = torch.cuda.ExternalStream(ctx.stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    agg = dict(extend_ms=0.0, extend_launches=0, scene_rays=0, instance_rays=0, launches=0, samples=0, gathers=0)
    clocks = None

    def do_step(i, timed):
        st = states[i // steps_per_image]
        ds.trace_samples(st, params)
        if timed:
            t, c = ctx.timings(), ctx.counters()
            agg["extend_ms"] += t["extend_ms"]
            agg["extend_launches"] += t["extend_launches"]
            agg["scene_rays"] += c["scene_rays"]
            agg["instance_rays"] += c["instance_rays"]
            agg["launches"] += c["launches"]
            agg["samples"] += c["camera_samples"]
        if (i + 1) % steps_per_image == 0 and world > 1:
            st.gather_image()  # the single collective of the path: ncclAllGather of the tiles
            if timed:
                agg["gathers"] += 1

    for i in range(args.warmup):
        do_step(i, False)
    if world > 1:
        # NCCL sets its connections up lazily at the first collective: do one un-timed gather so the timed region
        # pays for the gathers of the path only (round 1 billed the 8-rank connect to the render)
        states[0].gather_image()
    sync_all()
    with ClockSampler(local, enabled=(rank == 0)) as cs:
        ev0.record(stream)
        t0 = time.perf_counter()
        for i in range(args.warmup, n_total):
            do_step(i, True)
        ev1.record(stream)
        sync_all()
        wall = time.perf_counter() - t0
        clocks = cs.summary()
    dev_ms = ev0.elapsed_time(ev1)
    elapsed = max(dev_ms / 1e3, 0.0)
    t_all = torch.tensor([elapsed, wall], device="cuda", dtype=torch.float64)
    samples_all = torch.tensor([float(agg["samples"])], device="cuda", dtype=torch.float64)
    launches_all = torch.tensor([float(agg["launches"])], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(samples_all)
        dist.all_reduce(launches_all)
    elapsed, wall = float(t_all[0]), float(t_all[1])
    total_samples = float(samples_all[0])
    value = total_samples / elapsed / 1e6

    # ---- end-to-end through the public API with HOST buffers (scene upload + bvh build + render +
    # image download / gather inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        sync_all()
        t0 = time.perf_counter()
        ep = abi.trace_params(resolution=params.resolution, samples=spp_total, bounces=params.bounces, batch=spp_total)
        eds = lib.DeviceScene(ctx, scene)          # host BVH/lights build + H2D of the whole scene arena
        est = eds.make_state(ep, interleave=(rank, world)) if world > 1 else eds.make_state(ep)  # rng table + H2D
        eds.trace_samples(est, ep)
        img = est.gather_image() if world > 1 else est.download()["image"]  # D2H (after ncclAllGather)
        sync_all()
        te = time.perf_counter() - t0
        if dist:
            tt = torch.tensor([te], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt[0])
        assert np.isfinite(img).all()
        h2d = scene_bytes(scene) + est.num_rows * W * 16
        d2h = W * H * 16
        e2e = {"value": W * H * spp_total / te / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "seconds": te,
               "what": "DeviceScene(host scene)+make_state+trace_samples(all spp)+image download, wall clock"}
        del est, eds

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    ext_s = agg["extend_ms"] / 1e3
    achieved = bytes_per_ray * agg["scene_rays"] / max(ext_s, 1e-9) / 1e9
    traffic, traffic_src = None, None
    try:
        # NOT measured in this run (DRAM counters need ncu): bytes per scene ray of one `ncu --set full` capture of
        # k_extend on this workload, scaled to this run's rays per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "extend_traffic.json")))
        traffic = tj["dram_bytes_per_ray"] * agg["scene_rays"] / max(1, agg["extend_launches"])
        traffic_src = "static: " + tj.get("source", "profiles/extend_traffic.json") + " x this run's rays per launch"
    except Exception:
        pass
    roofline = {
        "kernel": "k_extend (closest-hit two-level BVH traversal)", "bound": "hbm",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_ray": bytes_per_ray,
        "algorithmic_bytes_checked": "kernel counters == instrumented reference (tests/test_gpu_parity.py::"
                                     "test_traversal_counters_match_instrumented_oracle)",
        "rays_per_launch": agg["scene_rays"] / max(1, agg["extend_launches"]),
        "avg_launch_ms": agg["extend_ms"] / max(1, agg["extend_launches"]),
        "kernel_share_of_step": ext_s / max(elapsed, 1e-9), "scene_Mrays_per_s_in_kernel": agg["scene_rays"] / max(ext_s, 1e-9) / 1e6,
        "traversal_per_ray": {k: cc[k] / max(1, cc["scene_rays"]) for k in
                              ("top_nodes", "bottom_nodes", "instance_visits", "triangle_tests", "quad_tests",
                               "line_tests", "point_tests")},
    }

    cpu = CPU_BASELINE
    line = {
        "metric": "Msamples/s (rays shaded/s) at 1920x1080x1024spp, 8 bounces" if args.workload == "c3" else "Msamples/s",
        "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "spp_per_step": spp_step, "steps_per_image": steps_per_image,
                   "image": [W, H], "tiling": "full image" if world == 1 else f"interleaved rows j % {world} == rank", "l2": "per-step working set (path state 200 B/pixel "
                   "x 2.07M pixels = 415 MB) exceeds the 126 MB L2", "parallelism": f"tiles{world}"},
        "Mrays_per_s_rank0": (agg["scene_rays"] + agg["instance_rays"]) / elapsed / 1e6,
        "wall_seconds": wall, "device_seconds": elapsed,
        "gpu_launches": int(launches_all[0]), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "e2e": e2e,
    }
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
