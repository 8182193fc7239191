#!/usr/bin/env python3
"""bench.py -- Mrays/s (shadow + AO any-hit rays) and ms/frame of the lighting pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

A step is one frame of the hot path through the reference-facing host calls
(GPUScene::UpdateResources[GPU] for animated configs, DeferredRenderer::LightPass, ::TAAPass
[+ the NCCL frame gather when N > 1], ::SwapLightHistory) on synthetic data of BASELINE.json's shape.
Default workload: configs[2] = "4K synthetic 10M-tri scene, 10k instances, 1 shadow ray/light + 16 AO
spp" (the config the 4K metric and the north-star target are quoted on; it fits one GPU).
N > 1: one process per GPU (torchrun), the frame is cut into bands of ~54 rows dealt round-robin to the
ranks (strong scaling: the frame is fixed), each rank shades its bands + the halo rows TAA reads, and one
ncclAllGather assembles the resolved frame (next frame's TAA history) on every GPU.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference's GLSL
(oracle/, BVH2 traverser, OpenMP over all host cores) on a bounded sample of the same workload: the
reference's Vulkan path cannot be built here (no Vulkan loader / glslang / lavapipe), which is the
substitution BASELINE.json allows; the line says so.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s (shadow+AO any-hit rays) of the deferred lighting + TAA frame"


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_workload(args, rt, tmp):
    """The workload as a Luz project, loaded through the host mirror (luz_b200/workloads.py)."""
    from luz_b200 import workloads
    return workloads.Workload(rt, args.config, variant=args.variant, width=args.width, height=args.height,
                              shadow_type=args.shadow_type, volumetric=args.volumetric, tmp=tmp,
                              light_samples=args.light_samples, ao_samples=args.ao_samples)


def blue_noise(scenes):
    p = os.path.join(ROOT, "tests", "golden", "blue_noise_256.rgba")
    if os.path.exists(p):  # a 256x256 crop of the reference's own texture
        return np.fromfile(p, dtype=np.uint8).reshape(256, 256, 4)
    return scenes.synthetic_blue_noise(1024)


def sysfs_bus_id(smi_csv, index):
    """PCI address of GPU `index` as sysfs spells it, from `nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader`
    (nvidia-smi prints an 8-digit PCI domain and upper case, sysfs a 4-digit one and lower case)."""
    for line in smi_csv.splitlines():
        parts = [t.strip() for t in line.split(",")]
        if len(parts) >= 2 and parts[0].isdigit() and int(parts[0]) == index:
            bus = parts[1].lower()
            return bus[4:] if len(bus.split(":")[0]) == 8 else bus
    return None


def cpulist_to_set(text):
    """'0-3,8,10-11' -> {0,1,2,3,8,10,11} (the format of /sys/devices/system/node/nodeN/cpulist)."""
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local):
    """Multi-GPU runs: keep this rank's threads (and so the page-locked buffers it first touches) on the NUMA node its GPU
    hangs off.  Eight ranks staging 50 MB per step each through buffers on one socket is what bounds the end-to-end
    number at 8 GPUs (profiles/r1_scaling.md).  Best effort: any failure leaves the affinity alone.  Returns a note for
    the JSON line.  LUZ_BENCH_NO_NUMA=1 disables it."""
    if os.environ.get("LUZ_BENCH_NO_NUMA"):
        return "disabled"
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local]) if visible and all(t.strip().isdigit() for t in visible.split(",")) else local
        bus = sysfs_bus_id(out, index)
        if bus is None:
            return "gpu not listed"
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "no NUMA information"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = cpulist_to_set(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return "no allowed CPU on node %d" % node
        os.sched_setaffinity(0, cpus)
        return "node %d (%d CPUs)" % (node, len(cpus))
    except Exception as e:  # noqa: BLE001 -- never fail the bench over placement
        return "unavailable (%s)" % type(e).__name__


def run_ours(args):
    import torch
    import torch.distributed as dist
    from luz_b200 import host as H
    from luz_b200 import rt as R
    from luz_b200 import scenes
    from luz_b200 import strips

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("--gpus %d needs torchrun (one process per GPU)" % args.gpus)
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the lighting path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "single GPU: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rt = R.LuzRT(device=local, rank=rank, world=world)
    tmp = tempfile.mkdtemp(prefix="luzbench_r%d_" % rank)
    wl = make_workload(args, rt, tmp)
    app, cfg = wl.app, wl.cfg
    W, Hh = wl.width, wl.height
    wl.upload(blue_noise(scenes))
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.from_numpy(rt.comm_unique_id()))
        dist.broadcast(idt, 0)
        rt.comm_init(idt.cpu().numpy())
    bvh_checked = False

    animate = cfg["animate"]

    def step(frame, first=False):
        wl.step(first)

    stream = torch.cuda.ExternalStream(rt.stream(), device=torch.device("cuda", local))
    # frame 0 builds the TLAS, produces the G-buffer on the device (input producer) and seeds the history
    step(0, first=True)
    rt.sync()
    if world > 1:  # every rank built its own replica of the BLASes / TLAS: check they are bitwise the same (collective)
        rt.comm_check_bvh()
        bvh_checked = True

    # one instrumented frame (outside the timed region): rays, nodes, triangles, instances per frame
    rt.set_debug(R.DEBUG_STATS)
    rt.light_pass(app.frame_count)
    st = rt.read(R.STATS)
    rt.set_debug(0)
    if os.environ.get("LUZRT_TILE_OCCL") and rank == 0:
        # diagnostic (stderr): how coherent shadow-ray occlusion is over the 8x4 pixel tile a warp traces
        sm = rt.read(R.SHADOW_MASK)[..., 0]
        lit = np.linalg.norm(rt.read(R.GBUF_NORMAL)[..., :3], axis=-1) > 0
        hh, ww = (sm.shape[0] // 4) * 4, (sm.shape[1] // 8) * 8
        tl = lit[:hh, :ww].reshape(hh // 4, 4, ww // 8, 8).sum(axis=(1, 3))
        for l in range(min(int(app.scene_block().num_lights), 8)):
            oc = (((sm >> l) & 1).astype(bool) & lit)[:hh, :ww].reshape(hh // 4, 4, ww // 8, 8).sum(axis=(1, 3))
            t = tl > 0
            sys.stderr.write("tile occlusion light %d: lit tiles %d, fully occluded %.3f, fully unoccluded %.3f, mixed %.3f, "
                             "rays occluded %.3f\n" % (l, t.sum(), ((oc == tl) & t).sum() / t.sum(), ((oc == 0) & t).sum() / t.sum(),
                                                        ((oc > 0) & (oc < tl)).sum() / t.sum(), oc.sum() / max(tl.sum(), 1)))

    for i in range(args.warmup):
        step(1 + i)
    rt.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = rt.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(1 + args.warmup + i)
    rt.device_ptr(R.IMG_LIGHT)  # orders the ctx stream after the last frame's gather (it runs on its own stream)
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = rt.launch_count() - launches0
    clocks = sampler.finish() if sampler else None

    # per-kernel times, measured live with CUDA events on the launch stream (a few extra frames)
    kt = {"light_ms": [], "taa_ms": [], "gather_ms": [], "tlas_ms": [], "gbuffer_ms": [], "volumetric_ms": [],
          "shadow_map_ms": [], "light_rays_ms": []}
    for i in range(min(args.steps, 8)):
        step(1 + args.warmup + args.steps + i)
        t = rt.read(R.TIMINGS)
        for k in kt:
            kt[k].append(getattr(t, k))
    kavg = {k: float(np.mean(v)) for k, v in kt.items()}
    t_hint = rt.read(R.TIMINGS)
    temporal = {"on": bool(t_hint.temporal_on), "settled_fraction": float(t_hint.temporal_settled)}
    # the same frames without the per-ray temporal occluder hints (what the first frame after a cut costs; same bits)
    rt.set_debug(R.DEBUG_NO_TEMPORAL)
    cold = []
    for i in range(min(args.steps, 6)):
        step(0)
        t = rt.read(R.TIMINGS)
        cold.append((t.light_ms, t.light_rays_ms))
    rt.set_debug(0)
    for i in range(3):  # hints warm again before the end-to-end runs
        step(0)
    temporal["light_ms_without"] = float(np.mean([c[0] for c in cold[1:]]))
    temporal["light_rays_ms_without"] = float(np.mean([c[1] for c in cold[1:]]))

    # ---- end to end through the C ABI with HOST buffers (G-buffer in, resolved frame out) ----
    own_rows = Hh // world
    e2e = None
    if not args.no_e2e:
        px = W * Hh
        gbufs = {}
        for sel, shape, dt in ((R.GBUF_ALBEDO, (Hh, W, 4), torch.uint8), (R.GBUF_NORMAL, (Hh, W, 4), torch.float32),
                               (R.GBUF_MATERIAL, (Hh, W, 4), torch.uint8), (R.GBUF_EMISSION, (Hh, W, 4), torch.uint8),
                               (R.GBUF_DEPTH, (Hh, W), torch.float32)):
            t = torch.empty(shape, dtype=dt, pin_memory=True)
            rt.read(sel, out=t.numpy())
            gbufs[sel] = t
        out_host = torch.empty((own_rows, W, 4), dtype=torch.float32, pin_memory=True)
        sb = app.scene_block()
        extra = app.extra_lights()

        need_maps = args.shadow_type == 2 or args.volumetric == 2

        def e2e_step(frame):
            rt.set_scene(sb, extra)
            if need_maps:
                rt.shadow_map_pass(1024)
            rt.set_gbuffer(gbufs[R.GBUF_ALBEDO].numpy(), gbufs[R.GBUF_NORMAL].numpy(), gbufs[R.GBUF_MATERIAL].numpy(),
                           gbufs[R.GBUF_EMISSION].numpy(), gbufs[R.GBUF_DEPTH].numpy())
            rt.light_pass(frame)
            if args.volumetric:
                rt.volumetric_pass(frame)
            rt.taa_pass(True)
            rt.gather()
            rt.read_owned(R.IMG_LIGHT, out_host.numpy())
            rt.swap_light_history()

        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        for i in range(n_e2e):
            e2e_step(2 + i)
        rt.sync()
        barrier()
        serial_ms = (time.perf_counter() - t0) * 1e3 / n_e2e

        # The same step with the copies of successive steps overlapped (the reference keeps 3 frames in flight,
        # VulkanWrapper.cpp:178): the next step's G-buffer is uploaded on a copy stream while this step is shaded,
        # and the resolved rows come back on a second copy stream, one step behind.  Every step still uploads its
        # own inputs from pinned memory and reads its own result.
        out_host2 = torch.empty_like(out_host).pin_memory()
        outs = [out_host, out_host2]

        def prefetch():
            rt.prefetch_gbuffer(gbufs[R.GBUF_ALBEDO].numpy(), gbufs[R.GBUF_NORMAL].numpy(), gbufs[R.GBUF_MATERIAL].numpy(),
                                gbufs[R.GBUF_EMISSION].numpy(), gbufs[R.GBUF_DEPTH].numpy())

        def pipelined(n, first_frame):
            prefetch()
            for i in range(n):
                rt.set_scene(sb, extra)
                if need_maps:
                    rt.shadow_map_pass(1024)
                rt.flip_gbuffer()
                prefetch()      # inputs of step i+1 go up (into the set step i-1 used) while step i is shaded
                rt.light_pass(first_frame + i)
                if args.volumetric:
                    rt.volumetric_pass(first_frame + i)
                rt.taa_pass(True)
                rt.gather()
                rt.read_wait()  # result of step i-1 has landed in outs[(i-1) % 2]
                rt.read_owned_async(R.IMG_LIGHT, outs[i % 2].numpy())
                rt.swap_light_history()
            rt.read_wait()
            rt.flip_gbuffer()   # retire the last prefetch
            rt.sync()

        pipelined(2, 20)
        barrier()
        # the timed run includes the fill (first upload) and the drain (last read-back) of the pipeline, which do
        # not overlap anything: enough steps that they amortise (one upload + one read-back ~ 8 ms at 4K)
        n_pipe = max(3, min(3 * args.steps, 120))
        t0 = time.perf_counter()
        pipelined(n_pipe, 30)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_pipe
        t_last = rt.read(R.TIMINGS)
        e2e = {"ms": e2e_ms, "serial_ms": serial_ms, "n_pipe": n_pipe, "light_ms": t_last.light_ms, "taa_ms": t_last.taa_ms, "h2d": strips.h2d_bytes(rank, world, W, Hh), "d2h": own_rows * W * 16}

        # The frame as Luz's own RenderFrame crosses the boundary (SURVEY 8b "data crossing"): the host hands over the
        # scene block, the model blocks and the instance transforms (GPUScene::UpdateResources[GPU]), the G-buffer is
        # produced on the device by the opaque pass, and the host reads the composed BGRA8 rows it owns (4 B/px).  Every
        # step uploads its inputs and reads its result; the TLAS is refit (the transforms do not change in this workload).
        comp_host = torch.empty((own_rows, W, 4), dtype=torch.uint8, pin_memory=True)
        n_inst = len(app.instances())

        def scene_step():
            wl.move(wl.frame)
            app.render_frame(H.FRAME_OPAQUE | H.FRAME_COMPOSE | (0 if animate == "rebuild" else H.FRAME_TLAS_REFIT))
            wl.frame += 1
            rt.read_owned(R.IMG_COMPOSE, comp_host.numpy())

        for i in range(3):
            scene_step()
        barrier()
        n_scene = max(3, min(args.steps, 30))
        t0 = time.perf_counter()
        for i in range(n_scene):
            scene_step()
        rt.sync()
        barrier()
        t_scene = rt.read(R.TIMINGS)
        e2e["scene_ms"] = (time.perf_counter() - t0) * 1e3 / n_scene
        e2e["scene_h2d"] = 31200 + 128 * n_inst + 72 * n_inst  # SceneBlock + ModelBlock[] + luzrt_instance[]
        e2e["scene_d2h"] = own_rows * W * 4
        e2e["scene_gbuffer_ms"] = t_scene.gbuffer_ms

    ms_per_step = total_ms / args.steps
    vals = torch.tensor([ms_per_step, float(st.rays), e2e["ms"] if e2e else 0.0, kavg["light_ms"], kavg["taa_ms"],
                         kavg["gather_ms"], kavg["tlas_ms"], float(st.nodes_visited), float(st.triangles_tested),
                         float(st.instances_entered), float(st.lit_pixels), e2e["scene_ms"] if e2e else 0.0],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    per_rank_light = [kavg["light_ms"]]
    if world > 1:
        lt = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(lt, torch.tensor([kavg["light_ms"]], dtype=torch.float64, device="cuda"))
        per_rank_light = [float(t.item()) for t in lt]
    mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
    ms_per_step = float(mx[0])
    rays_frame = float(sm[1])

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_sample(app, rt, R, W, Hh, budget_s=args.cpu_budget)

    # parity of the frames that were just timed: product kernels vs the oracle on a seeded row sample (checker, untimed)
    parity = None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import bench_parity as BP
        rt.set_debug(0)
        parity = BP.check_workload(wl, blue_noise(scenes), n_rows=args.parity_rows,
                                   candidate_rows=strips.owned_rows(rank, world, Hh) if world > 1 else None,
                                   taa=(world == 1))
        parity["pass"] = bool(parity["agree"] >= BP.MASK_AGREE
                              and (parity["max_abs_mask_identical_pixels"] <= BP.RADIANCE_TOL or
                                   parity["max_rel_mask_identical_pixels"] <= BP.REL_TOL)
                              and (parity["max_abs"] <= BP.RADIANCE_TOL or parity["psnr_db"] >= 50.0)
                              and parity.get("taa_max_abs", 0.0) <= BP.TAA_TOL
                              and parity.get("bvh2_check", {"rays_differ": 0})["rays_differ"] == 0)

    if rank == 0:
        hbm_peak, hbm_src = read_peaks()
        l2_gbs = rt.probe_read_bandwidth(32 << 20, 200)
        hbm_probe = rt.probe_read_bandwidth(4 << 30, 4)
        light_ms, taa_ms = float(mx[3]), float(mx[4])
        # per-GPU figures: the slowest rank's kernel time against the mean rank's work
        own_px = W * own_rows
        nodes, tris, insts, rays_r = float(sm[7]) / world, float(sm[8]) / world, float(sm[9]) / world, rays_frame / world
        trav_bytes = 80.0 * nodes + 48.0 * tris + 64.0 * insts
        shade_rows = len(strips.shaded_rows(rank, world, Hh))
        rays_ms = max(kavg["light_rays_ms"], 1e-6)
        shade_ms = max(kavg["light_ms"] - kavg["light_rays_ms"], 1e-6)
        # The ray kernel's bound: the BVH of every instanced configuration (a few MB) is cache resident, so the memory
        # roofline that applies to traversal is the L2 read bandwidth ("vs measured L2 or HBM peak as
        # applicable"), probed in this run; the unique-BLAS variant (~2 GB of BVH) streams from HBM.
        unique = args.variant == "unique"
        # (the unique-BLAS variant too: ncu shows its 1.9 GB of BVH served at 83 % L2 hit rate with 0.36 GB of DRAM traffic
        # per launch -- a frame touches the visible surface regions only -- profiles/r2_unique_blas.md)
        trav_peak = l2_gbs
        words = max((app.light_count() * cfg["light_samples"] + 31) // 32, 1) + max((cfg["ao_samples"] + 31) // 32, 1)
        rays_stream = (20.0 + 4.0 * words) * W * shade_rows     # normal 16 + depth 4 in, mask words out
        shade_stream = (48.0 + 4.0 * words) * W * shade_rows    # G-buffer 32 in, masks in, radiance 16 out
        src_hash = kernel_source_hash()
        out = {
            "metric": METRIC, "value": rays_frame / ms_per_step / 1e3, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s%s: %dx%d, %d instances, %d lights x %d shadow + %d AO rays/px%s" % (
                args.config, "-unique" if unique else "", W, Hh, len(app.instances()), app.light_count(),
                cfg["light_samples"], cfg["ao_samples"], (", TLAS %s per frame" % animate) if animate else "") +
                (", shadowType 2 (shadow maps, %d^2)" % 1024 if args.shadow_type == 2 else "") +
                (", volumetricType %d on every light" % args.volumetric if args.volumetric else ""),
                "parallelism": ("round-robin bands of %d rows x%d + ncclAllGather" % (strips.band_rows(Hh, world), world))
                if world > 1 else "1 GPU",
                "rays_per_frame": rays_frame, "lit_pixels": float(sm[10]),
                "l2_policy": ("inputs larger than L2 (G-buffer + 3 light buffers = %.0f MB > 126 MB)" if W * Hh * 80 > 126e6 else
                              "NOT flushed: G-buffer + 3 light buffers = %.0f MB fit the 126 MB L2 (reference-size config, "
                              "reported for parity, not a roofline claim)") % (W * Hh * 80 / 1e6)},
            "kernels_ms": {"light": light_ms, "taa": taa_ms, "gather": float(mx[5]), "tlas": float(mx[6]) if animate else 0.0,
                           "light_rays": kavg["light_rays_ms"], "light_shade": kavg["light_ms"] - kavg["light_rays_ms"],
                           "light_per_rank": [round(v, 4) for v in per_rank_light],
                           "volumetric": kavg["volumetric_ms"] if args.volumetric else 0.0,
                           "shadow_map": kavg["shadow_map_ms"] if (args.shadow_type == 2 or args.volumetric == 2) else 0.0},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # the dominant kernel (ray generation + any-hit traversal) against the memory level its working set lives in
            "roofline": {"kernel": "k_shadow_hints + the ray kernels (k_shadow_rays_temporal, k_ao_rays_compact: ray generation + "
                                   "any-hit traversal, the dominant kernels), timed with CUDA events on the launch stream", "bound": "l2",
                         "achieved": (trav_bytes + rays_stream) / (rays_ms * 1e6), "peak": trav_peak, "unit": "GB/s",
                         "frac": (trav_bytes + rays_stream) / (rays_ms * 1e6) / trav_peak if trav_peak else None,
                         "traffic": (lambda a, b, c: (a + b + c) if None not in (a, b, c) else None)(
                             ncu_traffic(args, "k_shadow_rays_temporal", src_hash),
                             ncu_traffic(args, "k_ao_rays_compact", src_hash), ncu_traffic(args, "k_shadow_hints", src_hash)),
                         "peak_source": "luzrt_probe_read_bandwidth, 32 MiB L2-resident buffer, measured in this run",
                         "ncu": ncu_counters(args, ("k_shadow_rays_temporal", "k_ao_rays_compact", "k_shadow_hints"), src_hash),
                         "algorithmic_bytes": "SURVEY 8(d): per ray 80 B/node + 48 B/triangle + 64 B/instance (counted by the "
                                              "statistics variant of the same kernel on the same BVH and rays) + 20 B/px "
                                              "normal+depth in + 4 B/px per mask word out",
                         "stream_bytes": rays_stream, "traversal_bytes": trav_bytes, "ms": rays_ms,
                         "bytes_per_ray": trav_bytes / max(rays_r, 1.0),
                         "nodes_per_ray": nodes / max(rays_r, 1.0),
                         "tris_per_ray": tris / max(rays_r, 1.0),
                         "instances_per_ray": insts / max(rays_r, 1.0),
                         "occluded_fraction": float(st.rays_occluded) / max(float(st.rays), 1.0),
                         "grays_per_s_per_gpu": rays_r / (rays_ms * 1e6),
                         "hbm_equivalent_frac": (trav_bytes + rays_stream) / (rays_ms * 1e6) / hbm_peak,
                         "note": "rays of pixels whose AO candidate list is empty are resolved by the per-pixel TLAS box query "
                                 "and count as traced (they are any-hit rays of the frame, answered exactly)"},
            "roofline_shade": {"kernel": "k_light_shade", "bound": "hbm", "achieved": shade_stream / (shade_ms * 1e6),
                               "peak": hbm_peak, "unit": "GB/s", "frac": shade_stream / (shade_ms * 1e6) / hbm_peak,
                               "traffic": ncu_traffic(args, "k_light_shade", src_hash), "peak_source": hbm_src, "ms": shade_ms,
                               "algorithmic_bytes": "48 B/px (32 B G-buffer + 16 B radiance) + 4 B/px per mask word"},
            "roofline_taa": {"kernel": "k_taa", "bound": "hbm", "achieved": 52.0 * own_px / (taa_ms * 1e6),
                             "peak": hbm_peak, "unit": "GB/s", "frac": 52.0 * own_px / (taa_ms * 1e6) / hbm_peak,
                             "traffic": ncu_traffic(args, "k_taa", src_hash), "peak_source": hbm_src, "hbm_read_probe_gbs": hbm_probe,
                             "ms": taa_ms},
            "l2_read_probe_gbs": l2_gbs,
            # per-ray temporal occluder hints (light_pass.cu k_shadow_rays_temporal): every shadow ray first tries the
            # triangles that occluded the same ray in the last frames; exact (same bits), adaptive (off when few rays are
            # settled); *_without = the same frames with LUZRT_DEBUG_NO_TEMPORAL
            "temporal_hints": temporal,
        }
        if bvh_checked:
            out["bvh_identical_across_ranks"] = True  # luzrt_comm_check_bvh raised otherwise
        if parity:
            out["parity"] = parity
        if e2e:
            out["e2e"] = {"value": rays_frame / float(mx[2]) / 1e3, "unit": "Mrays/s", "ms_per_step": float(mx[2]),
                          "mode": "host G-buffer in, resolved rows out every step; copies of successive steps overlapped "
                                  "(luzrt_prefetch_gbuffer / luzrt_read_owned_async), result one step behind; "
                                  "wall clock over %d steps incl. pipeline fill and drain" % e2e["n_pipe"],
                          "serial_ms_per_step": e2e["serial_ms"],
                          "kernels_ms_under_copies": {"light": e2e["light_ms"], "taa": e2e["taa_ms"]},
                          "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                          "host_numa_binding": numa_note}
            out["e2e_scene"] = {"value": rays_frame / float(mx[11]) / 1e3, "unit": "Mrays/s", "ms_per_step": float(mx[11]),
                                "mode": "the frame as Luz's RenderFrame crosses the boundary: scene block + model blocks + instance "
                                        "transforms in (UpdateResources[GPU], TLAS refit), G-buffer produced on the device (opaque pass, "
                                        "inside the timed region, not part of the ray count), composed BGRA8 rows of this rank out; "
                                        "blocking read-back every step",
                                "h2d_bytes_per_step": int(e2e["scene_h2d"]), "d2h_bytes_per_step": int(e2e["scene_d2h"]),
                                "gbuffer_ms": e2e["scene_gbuffer_ms"]}
        if cpu_base:
            out["cpu_baseline"] = cpu_base
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_source_hash():
    """sha1 over the CUDA sources: ties an ncu capture (profiles/traffic.json) to the code it was taken from."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "luz_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(args, kernel, src_hash):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture
    (profiles/traffic.json, written from the .ncu-rep by profiles/summarize.py).  The capture carries the hash of the
    kernel sources it was taken from: a capture of other code is stale and reported as null."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = args.config + ("-" + args.variant if args.variant else "")
        if args.width or args.gpus != 1 or t.get("source_hash") != src_hash:
            return None
        v = t.get(key, {}).get(kernel)
        return v.get("dram_bytes") if isinstance(v, dict) else v
    except Exception:
        return None


def ncu_counters(args, kernels, src_hash):
    """The committed ncu counters (active lanes, issue, L1 / L2 / DRAM throughput) of the named kernels, same staleness rule."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = args.config + ("-" + args.variant if args.variant else "")
        if args.width or args.gpus != 1 or t.get("source_hash") != src_hash:
            return None
        return {k: t[key][k] for k in kernels if k in t.get(key, {})} or None
    except Exception:
        return None


def oracle_world(app):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    return O, O.World(app.meshes(), app.instances())


def cpu_baseline_sample(app, rt, R, W, Hh, budget_s=15.0):
    """The oracle (CPU restatement, BVH2 traverser, all host cores) on a row sample of the same frame."""
    O, world = oracle_world(app)
    gb = O.GBuffer(W, Hh)
    rt.read(R.GBUF_ALBEDO, out=gb.albedo)
    rt.read(R.GBUF_NORMAL, out=gb.normal)
    rt.read(R.GBUF_MATERIAL, out=gb.material)
    rt.read(R.GBUF_EMISSION, out=gb.emission)
    rt.read(R.GBUF_DEPTH, out=gb.depth)
    return time_oracle_rows(O, world, app.scene_block(), app.extra_lights(), gb, W, Hh, budget_s, "port")


def time_oracle_rows(O, world, sb, extra, gb, W, Hh, budget_s, kind):
    from luz_b200 import scenes
    bn = blue_noise(scenes)
    O.lib().orc_set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = O.lib().orc_get_threads()
    mid = Hh // 2
    probe = min(cores, Hh)  # one row per thread (rows are the unit of the oracle's OpenMP loop)
    t0 = time.perf_counter()
    rc, _, _, _, st = O.light_pass(sb, gb, 0, bn, world, extra_lights=extra, exhaustive=False, rows=(mid, mid + probe))
    dt = max(time.perf_counter() - t0, 1e-4)
    rows = int(max(min(4 * cores, Hh), min(Hh, probe * budget_s / dt)))
    y0 = max(0, min(Hh - rows, mid - rows // 2))
    t0 = time.perf_counter()
    rc, _, _, _, st = O.light_pass(sb, gb, 0, bn, world, extra_lights=extra, exhaustive=False, rows=(y0, y0 + rows))
    dt = time.perf_counter() - t0
    return {"value": st.rays / dt / 1e6, "unit": "Mrays/s", "cores": O.lib().orc_get_threads(), "kind": kind,
            "sample": "light.frag restatement over rows [%d,%d) of the %dx%d frame (%d rays, %.1f s); "
                      "CPU restatement, not lavapipe (no Vulkan/glslang in the image)" % (y0, y0 + rows, W, Hh, st.rays, dt),
            "seconds": dt, "rays": int(st.rays)}


def run_reference(args):
    """CPU arm: the oracle port of the reference's lighting path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from luz_b200 import scenes
    tmp = tempfile.mkdtemp(prefix="luzbench_ref_")
    wl = make_workload(args, None, tmp)
    app, cfg = wl.app, wl.cfg
    W, Hh = wl.width, wl.height
    wl.upload(None)
    app.update_resources()
    O, world = oracle_world(app)
    # all host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.lib().orc_set_threads(cores)
    sb, extra = app.scene_block(), app.extra_lights()
    models, n_models = app.models()
    # bounded sample: a band of rows in the middle of the frame, at least 4 rows per core so that every core works
    # (rows are the unit of the oracle's OpenMP loop); its G-buffer is produced by the oracle's own generator (untimed)
    rows = min(Hh - 2, max(args.ref_rows if args.ref_rows else 4 * cores, 64 if not args.ref_rows else 2))
    y0 = Hh // 2 - rows // 2
    gb = O.gbuffer_pass(sb, world, models, n_models, app.textures(), W, Hh, exhaustive=False, rows=(y0, y0 + rows))
    bn = blue_noise(scenes)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc, light, _, _, st = O.light_pass(sb, gb, i, bn, world, extra_lights=extra, exhaustive=False, rows=(y0, y0 + rows))
        O.taa_pass(sb, light, light, gb.depth, True, rows=(y0 + 1, y0 + rows - 1))
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = st.rays
    ms = float(np.mean(times)) * 1e3
    val = rays / (ms * 1e3)
    threads = O.lib().orc_get_threads()
    sample = ("rows [%d,%d) of the %dx%d frame per step (%d rays, %d threads); CPU restatement of light.frag/taa.comp with a "
              "BVH2 traverser, not the Vulkan/lavapipe path (no Vulkan loader, glslang or lavapipe in the image)" % (
                  y0, y0 + rows, W, Hh, rays, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s%s: %dx%d, %d instances, %d lights x %d shadow + %d AO rays/px%s" % (
            args.config, "-unique" if args.variant == "unique" else "", W, Hh, len(app.instances()), app.light_count(),
            cfg["light_samples"], cfg["ao_samples"], (", TLAS %s per frame" % cfg["animate"]) if cfg["animate"] else ""),
            "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    # stdout carries exactly one JSON line: libraries that print banners to fd 1 (NCCL's version line) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--variant", default=None, choices=[None, "unique"])
    ap.add_argument("--shadow-type", type=int, default=1, choices=[0, 1, 2], help="scene.shadowType (2 = shadow maps)")
    ap.add_argument("--volumetric", type=int, default=0, choices=[0, 1, 2], help="volumetricType of every light")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-rows", type=int, default=0, help="rows per step of the CPU arm (0: 4 x host cores, >= 64)")
    ap.add_argument("--light-samples", type=int, default=None, help="override lightSamples (experiments; shown in config)")
    ap.add_argument("--ao-samples", type=int, default=None, help="override aoSamples (experiments; shown in config)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-rows", type=int, default=64)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
