#!/usr/bin/env python
"""bench.py — paths/s (depth 4) and ms/frame at 1080p on the named BASELINE.json workload.

  python bench.py --gpus N --steps K --warmup W            # CUDA arm (libvoxelrt, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement of
                                                           # the reference path on the host cores

Workload (config.workload = "config3"): BASELINE.json configs[2] — synthetic dense random 256^3
voxel grid (~50 % occupancy), 1920x1080, depth 4, physical sky + clouds, sun (1,1,1); this is the
configuration the north-star target (>= 1 Gpaths/s per B200) is quoted on and it fits one GPU.
A "step" is one accumulate() batch of --spp samples per pixel over the whole frame.

N > 1 (launched by torch.distributed.run, one rank per GPU): sample sharding — rank r renders
sample indices r, r+N, ... of every pixel (weak scaling: per-GPU work fixed), then ONE NCCL
all-reduce of the float4 accumulation buffer per step.

Timing: W untimed warm-up steps, then exactly K steps between barrier + synchronize, CUDA events
on the launching stream, max over ranks. The sky tables (2 x 236 MB as float4) exceed the 126 MB
L2 and every step touches ~100 MB of colour/occupancy/accumulation data, so inputs are larger than
L2 (config.l2 = "inputs>L2").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SUN = ((1, 1, 1), 0.025, (1.0 * 1.3, 0.949 * 1.3, 0.937 * 1.3))  # example6.py:10 colour, default direction


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=8, help="samples per pixel per step (per GPU)")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--sky-res", type=int, default=3840)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config3", choices=["config3", "config2", "config4"],
                    help="config3 (default, the bench line): dense random grid; config2: example6 fixture scene, sky + clouds; "
                         "config4: ReSTIR mode (render + spatial GRIS per frame) on the example6 scene")
    ap.add_argument("--cpu-spp", type=int, default=48, help="samples per pixel of the bounded CPU-baseline sample")
    return ap.parse_args()


WORKLOAD = "config3"


def build_scene(R):
    import scenes

    if WORKLOAD == "config3":
        return scenes.random_grid(R, 0.5, 1234)
    z = np.load(os.path.join(ROOT, "tests", "golden", "example6_seed0.npz"))  # example6.py run through the shim, seed 0
    return z["material"], z["color"]


def configure(r, R, mat, col, sky=True):
    r.set_voxels(mat, col)
    if WORKLOAD == "config3":
        r.set_floor(-1e5, (1.0, 1.0, 1.0))  # floor disabled as example9.py:4 does
        r.set_directional_light(*SUN)
    else:
        r.set_floor(-0.85, (1.0, 1.0, 1.0))                   # example6.py:8
        r.set_directional_light((1, 1, -1), SUN[1], SUN[2])    # example6.py:10
    if sky:
        r.set_use_physical_sky(True, True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_sample(args, W, H, R, mat, col, sky_tables, cpu_spp):
    """Oracle on all host cores over a bounded sample of the same workload: the full 1080p frame,
    same camera / scene / sky tables (downloaded from the GPU: the 3840^2 precompute is ~1e12
    inner iterations, not a CPU job), `cpu_spp` samples per pixel."""
    from oracle.binding import OracleRenderer, load
    from voxel_rt2_b200.materials import material_table

    lib = load()
    o = OracleRenderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=args.sky_res if sky_tables else 0, exposure=2.0, seed=1,
                       materials=material_table())
    configure(o, R, mat, col, sky=sky_tables is not None)
    if sky_tables is not None:
        o.set_sky_tables(*sky_tables)
    o.prepare_data()
    o.accumulate(cpu_spp, stats=True)
    c = o.counters()
    ms = o.last_ms()
    return {"paths": c["paths"], "ms": ms, "paths_per_s": c["paths"] / (ms * 1e-3), "counters": c, "cores": int(lib.orc_num_threads()),
            "spp": cpu_spp}


def _upsample_table(tab, S):
    """Bilinear resampling (wrap-around, texel centres) of an s x s x 3 sky table to S x S x 3."""
    s = tab.shape[0]
    x = (np.arange(S) + 0.5) * s / S - 0.5
    i0 = np.floor(x).astype(np.int64)
    f = (x - i0).astype(np.float32)
    a = tab[i0 % s] * (1.0 - f)[:, None, None] + tab[(i0 + 1) % s] * f[:, None, None]
    return np.ascontiguousarray(a[:, i0 % s] * (1.0 - f)[None, :, None] + a[:, (i0 + 1) % s] * f[None, :, None], dtype=np.float32)


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement (the reference itself is Taichi/Vulkan and cannot run here)."""
    if rank != 0:
        return
    W, H = [int(x) for x in args.res.split("x")]
    R = args.grid
    mat, col = build_scene(R)
    from oracle.binding import OracleRenderer, load
    from voxel_rt2_b200.materials import material_table

    lib = load()
    cores = int(lib.orc_num_threads())
    # sky tables: the CPU cannot run the 3840^2 precompute in bounded time (~1e12 inner iterations), so
    # the oracle computes a 64^2 table itself and it is resampled to the GPU arm's table size: every
    # lookup then walks the same 2 x 177 MB footprint (same cache behaviour) through the same code.
    S = 64
    o = OracleRenderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=S, cloud_passes=2, exposure=2.0, seed=1, materials=material_table())
    configure(o, R, mat, col, sky=True)
    o.prepare_data()
    if args.sky_res > S:
        tabs = [_upsample_table(t, args.sky_res) for t in o.get_sky_tables()]
        S = o.sky_res = args.sky_res
        o.set_sky_tables(*tabs)
    # each step = 1 spp over every 4th 8x4 tile (1/4 of the frame) so K+W steps end in minutes
    n = 4
    o.set_tile_shard(0, n)
    times, paths = [], 0
    for i in range(args.warmup + args.steps):
        before = o.counters()["paths"]
        o.accumulate(1, stats=True)
        if i >= args.warmup:
            times.append(o.last_ms())
            paths += o.counters()["paths"] - before
    tot = sum(times) * 1e-3
    v = paths / tot
    line = {"impl": "reference", "metric": "paths_per_sec_depth4_1080p", "value": v, "unit": "paths/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(len(times), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config3: dense random %d^3 (50%% occupancy), %dx%d, depth 4, physical sky + clouds" % (R, W, H),
                       "spp_per_step": 1, "sample": "every 4th 8x4 tile of the frame per step (1/4 frame, 1 spp)", "sky_res": S},
            "cpu_baseline": {"value": v, "unit": "paths/s", "cores": cores, "kind": "port",
                             "sample": "oracle C++/OpenMP restatement; per step 1 spp over every 4th 8x4 tile of the 1080p frame"},
            "e2e": {"value": v, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """Everything that libraries print to fd 1 during the run (NCCL's version banner, for one) goes to
    stderr; the one JSON line is written to the real stdout at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global WORKLOAD
    args = parse()
    _protect_stdout()
    WORKLOAD = args.workload
    if WORKLOAD != "config3":
        args.grid = 128
    if WORKLOAD == "config4":
        args.spp = 1  # ReSTIR mode renders one sample per frame
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import voxel_rt2_b200 as vrt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = [int(x) for x in args.res.split("x")]
    R, spp = args.grid, args.spp
    mat, col = build_scene(R)

    r = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=args.sky_res, exposure=2.0, seed=1, device=local_rank)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    configure(r, R, mat, col)
    r.set_sample_shard(rank, world)
    t_prep = time.time()
    r.prepare_data()
    t_prep = time.time() - t_prep
    sky_ms = r.stats()["sky_precompute_ms"]
    accum = r.accum_tensor()

    # untimed counter pass (algorithmic bytes per path)
    r.accumulate(1, stats=True)
    st = r.stats()
    per = {k: st[k] / max(st["paths"], 1) for k in ("rays", "steps", "queries", "hits", "sky_escapes", "nee_visible", "vertices")}
    b_path = 4 * per["queries"] + 4 * per["hits"] + 96 * per["sky_escapes"] + 48 * per["nee_visible"] + 32
    r.reset_framebuffer()

    def step():
        with torch.cuda.stream(stream):
            if world > 1:
                r.reset_framebuffer()
            if WORKLOAD == "config4":
                r.accumulate_restir(spp)
            else:
                r.accumulate(spp)
            if world > 1:
                dist.all_reduce(accum)  # one NCCL all-reduce of the accumulation buffer per step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = [], 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        s = r.stats()
        kernel_ms.append(s["last_render_ms"] + (s["last_gris_ms"] if WORKLOAD == "config4" else 0.0))
        launches += s["kernel_launches"]
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    paths_total = W * H * spp * args.steps * world
    value = paths_total / (ms_total * 1e-3)

    # ---- e2e: the call a user makes per displayed frame (scene.py:233-262): camera upload,
    # accumulate(spp), fetch_image into host memory. Host buffers, copies inside the timed region.
    pos, view, proj = vrt.default_camera_matrices(W, H)
    host_imgs = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(2)]

    def e2e_step():
        with torch.cuda.stream(stream):
            r.set_view_proj(pos, view, proj)
            if world > 1:
                r.reset_framebuffer()
            if WORKLOAD == "config4":
                r.accumulate_restir(spp)
            else:
                r.accumulate(spp)
            if world > 1:
                dist.all_reduce(accum)
            if rank == 0:
                # the displayed frame (merged over all ranks) goes to pinned host memory through the pipelined
                # fetch: tonemap on the render stream, D2H on the copy engine while the next step renders; the
                # call first waits for the previous step's image, the last one is waited for before the clock stops
                r.fetch_image_async(host_imgs[launches_e2e[0] // 2 % 2])
            launches_e2e[0] += 2

    launches_e2e = [0]
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_step()
    if rank == 0:
        r.wait_image()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert rank != 0 or float(host_imgs[0][..., :3].max()) > 0.0 and float(host_imgs[1][..., :3].max()) > 0.0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = W * H * spp * n_e2e * world / e2e_s
    h2d = 3 * 4 + 2 * 64 + 8 * spp
    d2h = W * H * 16

    line = None
    if rank == 0:
        peak, how = measured_peak()
        k_ms = float(np.mean(kernel_ms))
        achieved = (W * H * spp) * b_path / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_k_path_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("spp") == spp and tj.get("sky_res") == args.sky_res:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": "paths_per_sec_depth4_1080p", "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_frame": ms_total / args.steps / spp,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": {"config3": "config3: dense random %d^3 (50%% occupancy), %dx%d, depth 4, physical sky + clouds, sun (1,1,1)" % (R, W, H),
                                    "config2": "config2: example6 scene (shim seed 0) %d^3, %dx%d, depth 4, physical sky + clouds" % (R, W, H),
                                    "config4": "config4: ReSTIR mode (render + spatial GRIS 32 taps) on the example6 scene %d^3, %dx%d" % (R, W, H)}[WORKLOAD],
                       "spp_per_step_per_gpu": spp, "sky_res": args.sky_res, "parallelism": "sample-shard x%d + 1 all-reduce/step" % world,
                       "l2": "inputs>L2 (sky tables 2x%d MB + colour %d MB)" % (args.sky_res ** 2 * 16 // 2 ** 20, R ** 3 * 4 // 2 ** 20),
                       "sky_precompute_ms": sky_ms, "prepare_s": t_prep},
            "rays_per_sec": value * per["rays"], "per_path": per,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs)" if how == "measured" else "fallback 6650",
                         "bytes_per_path": b_path, "kernel": "k_path", "kernel_ms": k_ms,
                         "note": "latency/issue bound by construction: working set is L2-resident except the sky tables"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "paths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / n_e2e, "call": "set_view_proj + accumulate(spp) [+ all-reduce] + fetch_image_async -> pinned host on rank 0 (D2H of step k overlaps step k+1; every image is complete before the clock stops)"},
            "gpu_launches": launches,
        }
        if WORKLOAD == "config4":
            line["roofline"]["kernel"] = "k_path<restir> + k_gris"
            line["roofline"]["note"] = "bytes_per_path counts the path kernel only; GRIS adds 32 x 80 B of (L2-resident) tap reads per pixel"
        if not args.no_cpu_baseline and world == 1 and WORKLOAD != "config4":
            try:
                sky_tables = r.get_sky_tables()
                cb = cpu_sample(args, W, H, R, mat, col, sky_tables, args.cpu_spp)
                line["cpu_baseline"] = {"value": cb["paths_per_s"], "unit": "paths/s", "cores": cb["cores"], "kind": "port",
                                        "sample": "oracle (C++/OpenMP restatement), full 1080p frame of the same workload, %d spp, %d paths, %.1f s"
                                                  % (cb["spp"], cb["paths"], cb["ms"] * 1e-3),
                                        "per_path": {k: cb["counters"][k] / max(cb["paths"], 1) for k in
                                                     ("rays", "steps", "queries", "hits", "sky_escapes", "nee_visible")}}
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "paths/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
