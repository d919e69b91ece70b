#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Hald-CLUT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload generate|apply]

One JSON line on stdout (rank 0). BASELINE.json's metric has two parts and both are
measured in every run:

  * top level  : "LUT entries/s (level-16 Gaussian RBF)". One step = one complete level-16
                 CLUT (16,777,216 entries) of GaussianRemapper(catppuccin-mocha, shape 128,
                 nearest 16, lum 1.0, preserve off) -- BASELINE config 1's remapper at the level
                 the metric names. With N > 1 the rows are sharded across ranks and one in-place
                 NCCL all-gather assembles the CLUT on every rank (strong scaling: the work is one
                 CLUT however many GPUs share it).
  * "apply"    : "LUT apply Gpix/s". One step = the level-8 CLUT applied in place to a batch of
                 synthetic 3840x2160 RGBA8 frames resident in HBM (BASELINE configs 2/5; frames are
                 sharded across ranks, no collective, weak scaling).

`--workload apply` swaps the two (apply on top, generation under "generate").

`value` is device-resident throughput (CUDA events on the launching stream, L2 flushed
between generation steps; the apply batch is larger than L2). `e2e` is the same metric through
the public host API with page-locked HOST buffers, copies inside the timed region.
`cpu_baseline` / `--impl reference`: the reference is pure Rust and no Rust toolchain exists in
this environment, so its CPU path is timed as the C restatement in oracle/ ("port"), on the
box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# stdout carries exactly one JSON line: keep NCCL's version banner off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

LEVEL_GEN = 16
LEVEL_APPLY = 8
FRAME_W, FRAME_H = 3840, 2160
FRAME_PIX = FRAME_W * FRAME_H
GEN_PARAMS = dict(shape=128.0, nearest=16, lum=1.0, preserve=False)
# flops per entry = 8N + 9k + 77 (N palette colours, k neighbours), FMA = 2
GEN_N, GEN_K = 26, 16
GEN_FLOPS_PER_ENTRY = 8 * GEN_N + 9 * GEN_K + 77
APPLY_BYTES_PER_PIXEL = 8  # 4 read + 4 written, in place; CLUT traffic is cache resident
# config.workload: the SAME string in both arms (b200 and reference); how a run executes it goes into other config keys
GEN_WORKLOAD = ("generate level-16 Gaussian RBF CLUT (16,777,216 entries), catppuccin-mocha 26 colours, shape 128, nearest 16, "
                "lum 1.0, preserve off")
APPLY_WORKLOAD = "apply level-8 CLUT in place to 3840x2160 RGBA8 frames, uniform random pixels"


def host_threads():
    """Host cores this process may run on. NOT omp_get_max_threads(): torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would make the CPU arm single-threaded for N >= 2."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_near_gpu(torch, local_rank):
    """One process per GPU: run this process on the cores of the memory node its GPU hangs off (sysfs), so that the
    host pages it touches first are allocated there and its PCIe copies do not cross the socket link. Returns a short
    description for the JSON line. Only used for N > 1 (at N = 1 the CPU baseline wants every core)."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        dev = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "no memory-node information for the GPU"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"node {node} has none of this process's cores"
        os.sched_setaffinity(0, cpus)
        return f"node {node}, {len(cpus)} cores"
    except (OSError, ValueError, AttributeError) as e:
        return f"unavailable ({type(e).__name__})"


def load_palette():
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        return np.array(json.load(f)["catppuccin_mocha"], np.uint8)


def profiled_traffic():
    """DRAM bytes per launch of the two headline kernels from this round's ncu --set full captures."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm / CPU baseline
def cpu_generate(steps, warmup, threads):
    """Oracle port of par_generate_lut (lib.rs:143-147): the WHOLE level-16 CLUT per step, `threads` OpenMP
    threads passed explicitly (num_threads clause: the environment's OMP_NUM_THREADS does not apply)."""
    import oracle_lib as O
    pal = load_palette()
    r = O.Remapper(O.GAUSSIAN, pal, param=GEN_PARAMS["shape"], nearest=GEN_PARAMS["nearest"], lum_factor=GEN_PARAMS["lum"],
                   preserve=GEN_PARAMS["preserve"])
    side = LEVEL_GEN ** 3
    out = np.zeros((side, side, 3), np.uint8)
    entries = side * side

    def step():
        O.lib().lutgen_oracle_generate_lut_rows(r._h, LEVEL_GEN, out.ctypes.data, 0, side, threads)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return entries / dt, dt, f"the whole level-16 CLUT (16,777,216 entries) per step, {steps} steps after {warmup} warm-up", out


def cpu_apply(steps, warmup, threads):
    """Oracle port of correct_image on one 4K frame; threads=1 is what the reference does
    (identity.rs:71-78 is a serial loop)."""
    import oracle_lib as O
    from conftest import random_rgba
    pal = load_palette()
    lut = O.Remapper(O.GAUSSIAN, pal, param=128.0, nearest=16).generate_lut(LEVEL_APPLY)
    img = random_rgba(FRAME_PIX, 42080085)

    def step():
        buf = img.copy()
        t = time.perf_counter()
        O.lib().lutgen_oracle_correct_image(buf.ctypes.data, FRAME_PIX, lut.ctypes.data, LEVEL_APPLY, threads)
        return time.perf_counter() - t

    for _ in range(warmup):
        step()
    dt = float(np.mean([step() for _ in range(steps)]))
    return FRAME_PIX / dt / 1e9, dt, "one 3840x2160 RGBA8 frame per step"


def run_reference(args):
    """The reference's own CPU path for the same metric/config, timed on this box's host cores: the C restatement in
    oracle/ ("port": the reference is Rust and cannot be built here). Runs exactly the steps / warm-up it prints."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    gen_v, gen_dt, gen_sample, _ = cpu_generate(steps, warmup, cores)
    app_v, app_dt, app_sample = cpu_apply(steps, warmup, 1)
    app_all_v, _, _ = cpu_apply(min(steps, 5), 1, cores)
    gen = {"metric": "LUT entries/s (level-16 Gaussian RBF)", "value": gen_v, "unit": "entries/s", "ms_per_step": gen_dt * 1e3,
           "cpu_baseline": {"value": gen_v, "unit": "entries/s", "cores": cores, "kind": "port", "sample": gen_sample},
           "e2e": {"value": gen_v, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "config": {"workload": GEN_WORKLOAD, "level": LEVEL_GEN, "host_threads": cores}}
    app = {"metric": "LUT apply Gpix/s", "value": app_v, "unit": "Gpix/s", "ms_per_step": app_dt * 1e3,
           "cpu_baseline": {"value": app_v, "unit": "Gpix/s", "cores": 1, "kind": "port", "sample": app_sample,
                            "all_cores_value": app_all_v, "all_cores": cores},
           "e2e": {"value": app_v, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "config": {"workload": APPLY_WORKLOAD, "level": LEVEL_APPLY, "frames_per_step": 1, "host_threads": 1}}
    top, other, other_key = (gen, app, "apply") if args.workload == "generate" else (app, gen, "generate")
    line = dict(top)
    line.update({"impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "higher_is_better": True,
                 "scaling": "strong" if args.workload == "generate" else "weak", "vs_baseline": None,
                 "dtype": "f32/f64" if args.workload == "generate" else "u8", "data": "synthetic",
                 "note": "CPU restatement of the reference algorithm (oracle/), not the Rust binary: no Rust toolchain here; "
                         f"{cores} OpenMP threads passed explicitly (schedule(dynamic), as rayon's par_pixels_mut); apply is one thread as "
                         "identity.rs:71-78 is",
                 other_key: other})
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from lutgen_rs_b200 import _native, device, identity, interpolation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    binding = bind_near_gpu(torch, local_rank) if world > 1 else None
    device.init(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; stdout must
        # carry exactly one JSON line, so it is pointed at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    peaks, peak_kind = measured_peaks()
    L = _native.lib()
    pal = load_palette()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.fill_(1)

    def timed(fn, n, flush, prepare=None):
        """n timed calls of fn, each bracketed by CUDA events on the launching stream.
        `prepare` (untimed) restores the inputs an in-place step consumed."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            if prepare is not None:
                prepare()
            if flush:
                flush_l2()
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) * 1e-3 for a, b in evs]

    # ---------------- generation ----------------
    remapper = interpolation.GaussianRemapper(pal, GEN_PARAMS["shape"], GEN_PARAMS["nearest"], GEN_PARAMS["lum"], GEN_PARAMS["preserve"])
    side = LEVEL_GEN ** 3
    entries = side * side
    r0, r1, per = device.shard_rows(side, world, rank)
    lut16 = torch.empty(per * world * side * 3, dtype=torch.uint8, device="cuda")

    fused = world > 1 and not os.environ.get("LUTGEN_BENCH_ALLGATHER")
    if fused:
        # the fused form needs the ranks' buffers mapped into each other (symmetric memory or CUDA IPC); if a box
        # cannot do that, every rank falls back to the all-gather form together
        try:
            device.generate_lut_fused(remapper, LEVEL_GEN)
            torch.cuda.synchronize()
            ok = 1
        except Exception as e:  # noqa: BLE001
            print(f"rank {rank}: fused generation unavailable ({e!r}); using the all-gather form", file=sys.stderr)
            ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        fused = bool(int(t_ok.item()))

    def gen_step():
        # N > 1: every rank computes its share of the tiles and stores it into all ranks' CLUT buffers
        # over NVLink while it computes; a barrier through peer memory replaces the all-gather
        if fused:
            device.generate_lut_fused(remapper, LEVEL_GEN)
        else:
            device.generate_lut(remapper, LEVEL_GEN, out=lut16)

    def gen_allgather_step():
        device.generate_lut(remapper, LEVEL_GEN, out=lut16)

    def gen_kernel_only():
        device.generate_lut_rows_(remapper, LEVEL_GEN, lut16, r0, r1)

    # ---------------- apply ----------------
    lut8_host = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).generate_lut(LEVEL_APPLY)
    lut8 = torch.from_numpy(lut8_host).cuda().reshape(-1)
    frames_per_gpu = args.frames
    g = torch.Generator(device="cuda")
    g.manual_seed(42080085 + rank)
    shape = (frames_per_gpu, FRAME_H, FRAME_W, 4)
    # (a) uniform random RGBA: the worst case for the CLUT gather (no two neighbours share a cell)
    pristine_random = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g)
    # (b) photo-like: smooth gradients + a few LSB of noise per channel, opaque alpha
    xs = torch.arange(FRAME_W, device="cuda", dtype=torch.float32).view(1, 1, FRAME_W)
    ys = torch.arange(FRAME_H, device="cuda", dtype=torch.float32).view(1, FRAME_H, 1)
    ph = torch.arange(frames_per_gpu, device="cuda", dtype=torch.float32).view(-1, 1, 1) * 7.0
    chans = [(xs * (255.0 / FRAME_W) + ph) % 256.0 + 0 * ys, (ys * (255.0 / FRAME_H) + 2 * ph) % 256.0 + 0 * xs,
             ((xs + ys) * (255.0 / (FRAME_W + FRAME_H)) + 3 * ph) % 256.0]
    noise = torch.randint(-4, 4, (frames_per_gpu, FRAME_H, FRAME_W, 3), device="cuda", generator=g, dtype=torch.int16)
    pristine_photo = torch.empty(shape, dtype=torch.uint8, device="cuda")
    for c in range(3):
        pristine_photo[..., c] = (chans[c].to(torch.int16) + noise[..., c]).clamp_(0, 255).to(torch.uint8)
    pristine_photo[..., 3] = 255
    del noise, chans
    frames = torch.empty(shape, dtype=torch.uint8, device="cuda")
    apply_pix = frames_per_gpu * FRAME_PIX

    def apply_step():
        device.apply_hald_(frames, lut8, LEVEL_APPLY)

    # e2e buffers (page-locked host memory)
    lut16_host = torch.empty((side, side, 3), dtype=torch.uint8).pin_memory()
    e2e_frames = 4
    host_frames = [torch.randint(0, 256, (FRAME_H, FRAME_W, 4), dtype=torch.uint8).pin_memory() for _ in range(e2e_frames)]
    host_frames_np = [t.numpy() for t in host_frames]

    # N > 1, CLUT wanted on the host: ONE page-locked image in POSIX shared memory, mapped by every rank; each rank
    # generates its share of the row groups and copies it to its place over its own PCIe link (no GPU-to-GPU exchange)
    shared_img = None
    if world > 1:
        shm_path = f"/dev/shm/lutgen_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        if rank == 0:
            with open(shm_path, "wb") as f:
                f.truncate(3 * entries)
        barrier()
        shared_img = np.memmap(shm_path, dtype=np.uint8, mode="r+", shape=(side, side, 3))
        # every rank touches the rows it will write first (the process runs next to its GPU, bind_near_gpu): the pages
        # are allocated on that memory node, and the rank's copies stay on its own socket
        for r0_, r1_ in interpolation.InterpolatedRemapper.part_rows(LEVEL_GEN, rank, world):
            shared_img[r0_:r1_] = 0
        barrier()
        rc = torch.cuda.cudart().cudaHostRegister(shared_img.ctypes.data, 3 * entries, 0)
        if int(rc) != 0:
            print(f"rank {rank}: cudaHostRegister of the shared image failed ({rc}); copies will be pageable", file=sys.stderr)
        barrier()
        if rank == 0:
            os.unlink(shm_path)

    def gen_e2e():
        if world > 1:
            rm = interpolation.GaussianRemapper(pal, GEN_PARAMS["shape"], GEN_PARAMS["nearest"], GEN_PARAMS["lum"], GEN_PARAMS["preserve"])
            rm.generate_lut_part(LEVEL_GEN, shared_img, rank, world)
            del rm
            return
        if world == 1:
            # as `lutgen generate` does (crates/cli/src/main.rs:334-348): the remapper is built per invocation, so its
            # construction (palette upload + Oklab conversion on the device) is inside the timed region
            rm = interpolation.GaussianRemapper(pal, GEN_PARAMS["shape"], GEN_PARAMS["nearest"], GEN_PARAMS["lum"], GEN_PARAMS["preserve"])
            rm.generate_lut(LEVEL_GEN, out=lut16_host.numpy())
            del rm


    def apply_e2e():
        identity.correct_images(host_frames_np, lut8_host, LEVEL_APPLY)

    # warm-up
    for _ in range(warmup):
        gen_step()
        apply_step()
    gen_e2e()
    apply_e2e()
    barrier()

    # ---------------- output checks (outside every timed region) ----------------
    # N > 1: every rank's exchanged CLUT (fused and all-gather forms) must equal, byte for byte, what ONE GPU generates
    checks = {}
    ref16 = torch.empty(side * side * 3, dtype=torch.uint8, device="cuda")
    device.generate_lut_rows_(remapper, LEVEL_GEN, ref16, 0, side)
    torch.cuda.synchronize()
    if world > 1:
        same_fused = 1
        if fused:
            got = device.generate_lut_fused(remapper, LEVEL_GEN)
            torch.cuda.synchronize()
            same_fused = int(torch.equal(got.reshape(-1), ref16))
        got = device.generate_lut(remapper, LEVEL_GEN, out=lut16)
        torch.cuda.synchronize()
        same_ag = int(torch.equal(got.reshape(-1), ref16))
        t_same = torch.tensor([same_fused, same_ag], device="cuda")
        dist.all_reduce(t_same, op=dist.ReduceOp.MIN)
        if fused:
            checks["fused_matches_single_gpu"] = bool(int(t_same[0].item()))
        checks["allgather_matches_single_gpu"] = bool(int(t_same[1].item()))
        if not all(checks.values()):
            if rank == 0:
                print(json.dumps({"error": "multi-GPU CLUT differs from the one-GPU CLUT", **checks}), flush=True)
            dist.destroy_process_group()
            sys.exit(2)
    ref16_host = ref16.cpu().numpy().reshape(side, side, 3) if rank == 0 else None
    exchange_route = "n/a"
    if fused:
        peers = next(iter(device._peer_luts.values()), None)
        exchange_route = "NVSwitch multicast address of torch symmetric memory, multimem.st" if (peers and peers._mc[0]) else (
            "peer mappings of torch symmetric memory" if (peers and peers._symm) else "CUDA IPC peer mappings")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()

    # generation, device resident (whole step incl. all-gather), max over ranks
    barrier()
    gen_times = timed(gen_step, steps, flush=True)
    barrier()
    gen_launches = _native.launch_count() - launches0
    gen_t = max_over_ranks(float(np.mean(gen_times)))
    gen_kernel_times = timed(gen_kernel_only, steps, flush=True)
    gen_kernel_t = float(np.mean(gen_kernel_times))
    gen_ag_t = None
    if world > 1 and fused:
        barrier()
        gen_ag_t = max_over_ranks(float(np.mean(timed(gen_allgather_step, steps, flush=True))))
        barrier()
    # apply, device resident
    launches1 = _native.launch_count()
    barrier()
    apply_times = timed(apply_step, steps, flush=False, prepare=lambda: frames.copy_(pristine_random))
    barrier()
    apply_launches = _native.launch_count() - launches1
    apply_t = max_over_ranks(float(np.mean(apply_times)))
    photo_times = timed(apply_step, steps, flush=False, prepare=lambda: frames.copy_(pristine_photo))
    barrier()
    photo_t = max_over_ranks(float(np.mean(photo_times)))
    # end to end through the host API
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        gen_e2e()
    barrier()
    gen_e2e_t = max_over_ranks((time.perf_counter() - t0) / steps)
    if rank == 0:
        host_img = shared_img if world > 1 else lut16_host.numpy()
        checks["e2e_host_image_matches_device_clut"] = bool((np.asarray(host_img).reshape(-1) == ref16_host.reshape(-1)).all())
        assert checks["e2e_host_image_matches_device_clut"], "the CLUT delivered to the host differs from the device-resident one"
    host_pristine = [t.clone() for t in host_frames]
    acc = 0.0
    for _ in range(steps):
        for t, p in zip(host_frames, host_pristine):
            t.copy_(p)
        barrier()
        t0 = time.perf_counter()
        apply_e2e()
        acc += time.perf_counter() - t0
    barrier()
    apply_e2e_t = max_over_ranks(acc / steps)
    clocks = sampler.stop() if rank == 0 else None

    # roofline denominators
    fp32 = C_double()
    mufu = C_double()
    _native.check(L.lutgen_cuda_measure_peaks(fp32.ref, mufu.ref))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    shard_entries = (r1 - r0) * side
    gen_achieved = shard_entries * GEN_FLOPS_PER_ENTRY / gen_kernel_t / 1e12
    gen_line = {
        "metric": "LUT entries/s (level-16 Gaussian RBF)", "value": entries / gen_t, "unit": "entries/s",
        "ms_per_step": gen_t * 1e3, "scaling": "strong", "dtype": "f32 (f64 on flagged near-ties)",
        "gpu_launches": int(gen_launches),
        "e2e": {"value": entries / gen_e2e_t, "unit": "entries/s", "h2d_bytes_per_step": 4 * len(pal), "d2h_bytes_per_step": 3 * entries,
                "ms_per_step": gen_e2e_t * 1e3,
                "what": ("GaussianRemapper(palette, ...) constructed (palette upload: 4 bytes per colour), generate_lut(16) into a page-locked "
                         "host buffer, remapper dropped -- every step") if world == 1 else
                        (f"every rank: GaussianRemapper constructed, lutgen_cuda_generate_lut_part -- its share of the row groups generated and copied "
                         f"over its own PCIe link into ONE page-locked image in shared memory ({world} ranks, no GPU-to-GPU exchange), remapper dropped -- "
                         "every step; h2d bytes are per rank"),
                **({"host_binding": f"each rank runs on the cores of its GPU's memory node and touched its own rows of the image first (rank 0: {binding})"}
                   if world > 1 else {})},
        "roofline": {"bound": "fp32", "achieved": gen_achieved, "peak": fp32.value, "unit": "TFLOP/s",
                     "frac": gen_achieved / fp32.value if fp32.value else None,
                     "traffic": profiled_traffic().get("k_remap_warp_level16_bytes_per_launch") if world == 1 else None,
                     "traffic_note": "DRAM bytes of one k_remap_warp launch (ncu --set full, profiles/r2_traffic.json): the CLUT it writes stays in L2",
                     "peak_source": "on-box FFMA micro-benchmark (lutgen_cuda_measure_peaks), this run",
                     "mufu_peak_tops": mufu.value, "kernel_ms": gen_kernel_t * 1e3,
                     "algorithmic_flops_per_entry": GEN_FLOPS_PER_ENTRY,
                     "step_ms_each": [round(t * 1e3, 4) for t in gen_times],
                     "kernel_ms_each": [round(t * 1e3, 4) for t in gen_kernel_times]},
        "config": {"workload": GEN_WORKLOAD, "level": LEVEL_GEN,
                   "execution": ("one GPU" if world == 1 else
                                 (f"tiles interleaved over {world} ranks; the kernel stores its results into every rank's buffer over NVLink "
                                  f"({exchange_route}) and signals a peer-memory barrier" if fused else
                                  "rows sharded over ranks + in-place NCCL all-gather")),
                   "l2": "flushed between timed iterations (512 MiB fill)"},
    }
    if gen_ag_t is not None:
        gen_line["allgather_variant"] = {"ms_per_step": gen_ag_t * 1e3, "value": entries / gen_ag_t,
                                         "what": "same CLUT as contiguous row shards + in-place NCCL all-gather (the unfused scheme)"}
    total_pix = apply_pix * world
    apply_bw = apply_pix * APPLY_BYTES_PER_PIXEL / apply_t / 1e9
    apply_line = {
        "metric": "LUT apply Gpix/s", "value": total_pix / apply_t / 1e9, "unit": "Gpix/s", "ms_per_step": apply_t * 1e3,
        "scaling": "weak", "dtype": "u8", "gpu_launches": int(apply_launches),
        "e2e": {"value": e2e_frames * FRAME_PIX * world / apply_e2e_t / 1e9, "unit": "Gpix/s",
                "h2d_bytes_per_step": e2e_frames * FRAME_PIX * 4 + lut8_host.size, "d2h_bytes_per_step": e2e_frames * FRAME_PIX * 4,
                "ms_per_step": apply_e2e_t * 1e3},
        "roofline": {"bound": "hbm", "achieved": apply_bw, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": apply_bw / peaks["hbm_gbs"],
                     "traffic": (profiled_traffic().get("k_apply_bytes_per_pixel") or 0) * apply_pix or None,
                     "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"},
        "config": {"workload": APPLY_WORKLOAD, "level": LEVEL_APPLY, "frames_per_gpu": frames_per_gpu,
                   "note": "uniform random pixels are the worst case: every gather is its own 32 B L2 sector",
                   "l2": f"batch is {apply_pix * 4 >> 20} MiB per GPU, larger than the 126 MB L2; frames restored (untimed) before every step"},
    }
    photo_bw = apply_pix * APPLY_BYTES_PER_PIXEL / photo_t / 1e9
    apply_line["photo_like"] = {
        "value": total_pix / photo_t / 1e9, "unit": "Gpix/s", "ms_per_step": photo_t * 1e3,
        "roofline": {"bound": "hbm", "achieved": photo_bw, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": photo_bw / peaks["hbm_gbs"]},
        "config": {"workload": "same batch shape, photo-like pixels: per-frame colour gradients + 3 bits of noise per channel, alpha 255"},
    }
    # CPU baseline (rank 0, N = 1 only): oracle port on the host cores, bounded sample
    if world == 1 and not args.no_cpu:
        import oracle_lib as O
        cores = host_threads()
        v, _, sample, oracle_lut = cpu_generate(3, 1, cores)
        gen_line["cpu_baseline"] = {"value": v, "unit": "entries/s", "cores": cores, "kind": "port", "sample": sample}
        diff = np.abs(ref16_host.astype(np.int16) - oracle_lut.astype(np.int16))
        checks["generate_vs_oracle"] = {"max_lsb": int(diff.max()), "bytes_off_by_one": int((diff == 1).sum()), "bytes": int(diff.size),
                                        "checked": "the whole level-16 CLUT of the timed remapper against the CPU oracle (contract: +-1 LSB)"}
        assert diff.max() <= 1, "generated CLUT differs from the oracle by more than 1 LSB"
        one = frames[0].cpu().numpy().reshape(-1, 4)   # a frame the timed apply step produced
        src = pristine_photo[0].cpu().numpy().reshape(-1, 4)
        want = O.correct_image(src.reshape(FRAME_H, FRAME_W, 4), lut8_host, LEVEL_APPLY).reshape(-1, 4)
        checks["apply_vs_oracle"] = {"bit_exact": bool((one == want).all()), "checked": "one photo-like 4K frame of the timed batch against the CPU oracle"}
        assert checks["apply_vs_oracle"]["bit_exact"], "applied frame differs from the oracle"
        v, _, sample = cpu_apply(5, 1, 1)
        apply_line["cpu_baseline"] = {"value": v, "unit": "Gpix/s", "cores": 1, "kind": "port",
                                      "sample": sample + " (single thread, as identity.rs:71-78)"}
        v, _, _ = cpu_apply(5, 1, cores)
        apply_line["cpu_baseline"]["all_cores_value"] = v
        apply_line["cpu_baseline"]["all_cores"] = cores

    # the other BASELINE.json configurations, device resident on one GPU (rank 0; a few repetitions
    # each, after the timed regions above): evidence next to the headline, not part of `value`
    # N = 1 only: with peer mappings in place (N > 1) a cudaMalloc on one rank has to be mapped into the peers,
    # which cannot answer while they wait inside an NCCL kernel -- a rank must not work alone between collectives
    other_configs = None
    if world == 1:
        from conftest import synthetic_palette
        pal256 = synthetic_palette(256, 1)

        def ms_of(make, level, reps=5, evict_level=None):
            side_ = level ** 3
            buf = torch.empty(3 * side_ * side_, dtype=torch.uint8, device="cuda")
            rm = make()
            ts = []
            for i in range(reps + 2):
                if evict_level:
                    # the blur remapper caches the blurred volume of the last level: ask for another one first
                    device.generate_lut_rows_(rm, evict_level, buf, 0, evict_level ** 3)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush_l2()
                a.record(stream)
                device.generate_lut_rows_(rm, level, buf, 0, side_)
                b.record(stream)
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(b))
            return round(float(np.median(ts)), 4)

        other_configs = {
            "shepard_level16_256colours_nearest16_ms": ms_of(lambda: interpolation.ShepardRemapper(pal256, 4.0, 16, 1.0, False), 16),
            "gaussian_sampling_level12_512iter_256colours_ms": ms_of(
                lambda: interpolation.GaussianSamplingRemapper(pal256, 0.0, 20.0, 512, 1.0, 42080085, False), 12),
            "nearest_neighbor_level16_ms": ms_of(lambda: interpolation.NearestNeighborRemapper(pal, 1.0, False), 16),
            "gaussian_blur_level16_radius8_ms": ms_of(lambda: interpolation.GaussianBlurRemapper(pal, 8.0, 1.0, False), 16, reps=3, evict_level=4),
            "gaussian_rbf_level8_ms": ms_of(lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), 8),
            "gaussian_rbf_level12_ms": ms_of(lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), 12),
            "what": "ms per whole CLUT on ONE GPU, device resident, L2 flushed (BASELINE.json configs 3, 4, 1 and SURVEY 8f#1)",
        }
        # BASELINE config 5, second half: direct remap_image (no CLUT file in between) over 4K frames. The remapper is
        # tabulated over the sRGB cube once per handle (its level-16 CLUT, built on first use) and the frames gather from it.
        rm = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
        nfr = min(8, frames_per_gpu)
        ts = []
        for i in range(5):
            frames[:nfr].copy_(pristine_photo[:nfr])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            device.remap_image_(rm, frames[:nfr])
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        other_configs["direct_remap_image_gpix_per_s"] = round(nfr * FRAME_PIX / (float(np.median(ts[1:])) * 1e-3) / 1e9, 1)
        other_configs["direct_remap_image_first_call_ms"] = round(ts[0], 3)
        other_configs["direct_remap_image_what"] = (f"GaussianRemapper.remap_image on {nfr} photo-like 4K RGBA frames in one call, device resident; "
                                                    "the first call includes tabulating the remapper over the sRGB cube")
        # the PNG writer around the path (`lut.save(path)`): the headline CLUT encoded on the device,
        # `lutgen generate` end to end as generate + encode + download of the PNG bytes, and a CPU PNG encoder beside it.
        from lutgen_rs_b200 import png as lpng
        side16 = 16 ** 3
        clut = torch.empty((side16, side16, 3), dtype=torch.uint8, device="cuda")
        device.generate_lut_rows_(rm, 16, clut.view(-1), 0, side16)
        png_out = torch.empty(lpng.bound(side16, side16, 3), dtype=torch.uint8, device="cuda")
        ts = []
        for i in range(6):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            view = device.encode_png_(clut, png_out, stream)
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        other_configs["png_encode_level16_clut_ms"] = round(float(np.median(ts[1:])), 3)
        other_configs["png_encode_level16_clut_bytes"] = int(view.numel())
        other_configs["png_encode_level16_clut_gb_per_s_of_pixels"] = round(clut.numel() / (float(np.median(ts[1:])) * 1e-3) / 1e9, 1)
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            data = lpng.generate_lut_png(rm, 16)
            ts.append((time.perf_counter() - t0) * 1e3)
        other_configs["generate_lut_png_level16_host_to_host_ms"] = round(float(np.median(ts[1:])), 2)
        # `lutgen apply` on a still image: upload a 4K RGBA frame, correct it with the level-8 CLUT, download its PNG
        frame_host = pristine_photo[0].cpu().numpy().reshape(FRAME_H, FRAME_W, 4)
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            fdata = lpng.correct_image_png(frame_host, lut8_host, LEVEL_APPLY)
            ts.append((time.perf_counter() - t0) * 1e3)
        other_configs["correct_image_png_4k_frame_host_to_host_ms"] = round(float(np.median(ts[1:])), 2)
        other_configs["correct_image_png_4k_frame_bytes"] = len(fdata)
        if not args.no_cpu:
            try:
                import io

                from PIL import Image
                host = clut.cpu().numpy()
                assert (np.array(Image.open(io.BytesIO(data))) == host).all(), "PNG does not decode to the CLUT"
                t0 = time.perf_counter()
                bio = io.BytesIO()
                Image.fromarray(host).save(bio, "PNG", compress_level=1)
                other_configs["cpu_png_encode_level16_clut_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
                other_configs["cpu_png_encode_level16_clut_bytes"] = len(bio.getvalue())
            except ImportError:
                pass
        other_configs["png_what"] = ("PNG file bytes of the level-16 headline CLUT: device resident (L2 flushed), then generate + encode + "
                                     "download through lutgen_cuda_generate_lut_png into a page-locked buffer, copied into a Python bytes object; "
                                     "correct_image_png_* = a photo-like 4K RGBA frame from pageable host memory through lutgen_cuda_apply_hald_png; cpu_* = PIL (zlib level 1, one "
                                     "thread) on the same CLUT, checked to decode to the same pixels -- the reference's `image` crate "
                                     "encoder cannot run here")
        # palette extraction (k-means in Oklab) on a 1080p crop of a photo-like frame, host to host;
        # first measured by this run -- a failure here must not take the line down
        try:
            from lutgen_rs_b200 import extract as lextract
            crop = np.ascontiguousarray(frame_host[:1080, :1920, :3])
            ts = []
            for i in range(4):
                t0 = time.perf_counter()
                pal16 = lextract.extract_palette(crop, 16, 16)
                ts.append((time.perf_counter() - t0) * 1e3)
            other_configs["extract_palette_1080p_16colours_16iters_host_to_host_ms"] = round(float(np.median(ts[1:])), 2)
            other_configs["extract_palette_colours"] = int(len(pal16))
            if not args.no_cpu:
                import oracle_lib as O_ext  # the checker / CPU baseline leg only
                t0 = time.perf_counter()
                want16 = O_ext.extract_palette(crop, 16, 16)
                other_configs["cpu_extract_palette_1080p_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
                other_configs["extract_palette_matches_cpu_restatement"] = bool(pal16.shape == want16.shape and (pal16 == want16).all())
            other_configs["extract_what"] = ("deterministic k-means in Oklab (csrc/extract_core.h), NOT the reference's quantette palette (parity "
                                             "unpinned); cpu_* = the oracle's one-thread restatement of the same specification")
        except Exception as e:  # noqa: BLE001
            other_configs["extract_error"] = f"{type(e).__name__}: {e}"
    top, other, other_key = (gen_line, apply_line, "apply") if args.workload == "generate" else (apply_line, gen_line, "generate")
    line = dict(top)
    line.update({"n_gpus": world, "steps": steps, "warmup": warmup, "higher_is_better": True, "vs_baseline": None,
                 "data": "synthetic", "clocks": clocks, "checks": checks, other_key: other})
    line.update({k: v for k, v in checks.items() if isinstance(v, bool)})
    if other_configs is not None:
        line["other_configs"] = other_configs
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class C_double:
    def __init__(self):
        import ctypes
        self._v = ctypes.c_double(0.0)
        self.ref = ctypes.byref(self._v)

    @property
    def value(self):
        return float(self._v.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="generate", choices=["generate", "apply"])
    ap.add_argument("--frames", type=int, default=32, help="4K frames per GPU in the device-resident apply batch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
