#!/usr/bin/env python
"""Benchmark of the Vlaser-VLA `infer_action` hot path (BASELINE.json metric: policy actions/s,
Vlaser-2B-VLA, bs=64 per GPU, bf16, 10 flow steps; configs[1] at N=1, data-parallel weak scaling at N>1).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host cores (oracle port)

One "step" = one policy call on one batch of synthetic observations (64 per GPU). `value` is timed with
CUDA events around K calls whose inputs are already resident in HBM; `e2e` is the same call through
the host-buffer C-ABI entry (H2D of the inputs from pinned memory and D2H of the actions inside the
timed region). Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vlaser_b200.config import PRESETS  # noqa: E402
from vlaser_b200.synth import make_inputs  # noqa: E402
from vlaser_b200.weights import random_state_dict, state_dict_spec  # noqa: E402

WORKLOAD = "Vlaser-2B-VLA WidowX/Bridge bs=64 bf16, 10 flow steps, single B200 (BASELINE configs[1])"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d["bf16_tflops_sustained"]), "tflops_burst": float(d["bf16_tflops"]),
                "hbm_gbs": float(d["hbm_gbs"]), "which": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "which": "fallback (B200_PROFILING.md)"}


def _ncu_traffic(kernel: str):
    """DRAM bytes (read + write) per launch of a kernel family from the committed ncu pass over one policy call
    (tools/summarize_traffic.py -> profiles/r02_traffic.json, else the round-1 capture); None when no capture holds the
    family. ncu cannot run inside the timed bench, so this is the same workload measured once under the profiler.
    "gemm_tcgen05_pair_kernel@full_grid" = the pair-kernel launches with one CTA pair per SM pair, i.e. exactly the
    large-M launches the roofline line's algorithmic bytes are about."""
    for tag in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", f"{tag}_traffic.json")
        try:
            with open(p) as f:
                k = json.load(f)["kernels"][kernel]
            return k["traffic_bytes_per_launch"]
        except Exception:
            continue
    return None


def _gpu_port_timing(cfg, dev, B: int, steps: int, compiled: bool):
    """Baseline leg only (never the product path): the reference ALGORITHM -- the oracle port, dense padded S = 384
    prefix, dense masks, eager attention, as `PiZero.infer_action` executes it -- on the same GPU in bf16, eager or under
    torch.compile(mode="default") (what Vlaser_VLA/Simpler/src/agent/eval.py:46-53 does to the reference model).
    CUDA events around `steps` calls after 2 warm-up calls; returns ms per call."""
    from oracle import vlaser_oracle as O  # baseline legs of bench.py only
    sd = {k: t for k, t in _device_state_dict_stream(cfg, 0, dev)}
    inp = make_inputs(cfg, B, 1)
    ids = inp["input_ids"].to(dev)
    pix, prop, noise = (inp[k].to(torch.bfloat16).to(dev) for k in ("pixel_values", "proprios", "noise"))
    fn = O.infer_action
    if compiled:
        fn = torch.compile(O.infer_action, mode="default")
    with torch.inference_mode():
        for _ in range(2):
            fn(cfg, sd, ids, pix, prop, noise)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(cfg, sd, ids, pix, prop, noise)
        e1.record()
        torch.cuda.synchronize()
    del sd
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / steps


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def _cpu_oracle_timing(cfg, sd, steps: int, warmup: int):
    """Times the reference algorithm (oracle port, fp32, all host threads) on B=1 observations of the workload."""
    from oracle import vlaser_oracle as O  # bench.py's cpu_baseline / --impl reference legs only
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            inp = make_inputs(cfg, 1, 100 + i)
            t0 = time.perf_counter()
            O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], inp["noise"])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, torch.get_num_threads()


def _device_state_dict_stream(cfg, seed: int, device):
    """Random-init weights generated directly on the GPU (bench only; shapes/architecture as the model)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for key, (shape, kind) in state_dict_spec(cfg).items():
        if kind == "n02":
            t = torch.randn(shape, generator=g, device=device) * 0.02
        elif kind == "randn":
            t = torch.randn(shape, generator=g, device=device)
        elif kind == "one":
            t = 1.0 + torch.randn(shape, generator=g, device=device) * 0.02
        elif kind == "ls":
            t = 0.1 + torch.randn(shape, generator=g, device=device) * 0.002
        else:
            bound = 1.0 / float(kind[3:]) ** 0.5
            t = (torch.rand(shape, generator=g, device=device) * 2 - 1) * bound
        yield key, t.to(torch.bfloat16)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = PRESETS[args.config]()
    sd = random_state_dict(cfg, 0)
    times, cores = _cpu_oracle_timing(cfg, sd, args.steps, args.warmup)
    per_obs = sum(times) / len(times)
    value = cfg.horizon_steps / per_obs
    line = {
        "impl": "reference", "metric": "policy actions/s", "value": value, "unit": "actions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_obs * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "each step = 1 observation (B=1) of the workload, fp32, padded S=384",
                   "kind": "port (oracle/vlaser_oracle.py: the reference algorithm restated in plain torch and pinned to the "
                           "unmodified reference at 2e-5; /root/reference does not exist on the GPU box and has no installable package)"},
        "cpu_baseline": {"value": value, "unit": "actions/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} single-observation calls of the oracle port (reference algorithm, fp32)"},
        "e2e": {"value": value, "unit": "actions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from vlaser_b200.dp import gather_actions
    from vlaser_b200.policy import PiZeroInference

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = PRESETS[args.config]()
    Bg = args.batch_per_gpu
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the GPU count")
        Bg = args.global_batch // world
    peaks = _peaks()

    # ---- weights (random-init, bf16). N=1 rank 0 builds them on the CPU so the cpu_baseline leg can share them.
    want_cpu = (world == 1 and not args.no_cpu_baseline)
    pol = PiZeroInference(cfg, max_batch=Bg, device=local)
    sd_cpu = None
    if want_cpu:
        sd_cpu = random_state_dict(cfg, 0)
        pol.load_state_dict(sd_cpu, strict=True)
    else:
        for k, t in _device_state_dict_stream(cfg, 0, dev):
            pol.load_state_dict({k: t}, strict=False)
    pol.finalize()

    inp = make_inputs(cfg, Bg, 1 + rank)
    valid = inp["valid_len"]
    ids_h = inp["input_ids"]
    pix_h = inp["pixel_values"].to(torch.bfloat16).pin_memory()
    prop_h = inp["proprios"].to(torch.bfloat16).pin_memory()
    noise_h = inp["noise"].to(torch.bfloat16).pin_memory()
    ids_d, pix_d, prop_d, noise_d = ids_h.to(dev), pix_h.to(dev), prop_h.to(dev), noise_h.to(dev)
    vl_d = valid.to(dev)
    global_B = Bg * world

    def step_device():
        out = pol.infer_action(ids_d, pix_d, proprios=prop_d, noise=noise_d)
        return gather_actions(out, global_B)

    def step_host():
        out = pol.infer_action_host(ids_h, pix_h, prop_h, noise_h)
        if world > 1:
            return gather_actions(out.to(dev), global_B).cpu()
        return out

    # widened boundary: u8 camera frames cross PCIe, resize + normalisation run on the device
    from vlaser_b200.adapter import FrameStager
    stager = FrameStager(cfg.image_size)
    gf = torch.Generator().manual_seed(100 + rank)
    frames_h = torch.randint(0, 256, (Bg * cfg.num_images, 480, 640, 3), generator=gf, dtype=torch.uint8).pin_memory()
    frames_d = torch.empty_like(frames_h, device=dev)
    ids_s, prop_s, noise_s = torch.empty_like(ids_d), torch.empty_like(prop_d), torch.empty_like(noise_d)
    ids_hp = ids_h.pin_memory()

    def step_staged():
        frames_d.copy_(frames_h, non_blocking=True)
        ids_s.copy_(ids_hp, non_blocking=True)
        prop_s.copy_(prop_h, non_blocking=True)
        noise_s.copy_(noise_h, non_blocking=True)
        out = pol.infer_action(ids_s, stager(frames_d), proprios=prop_s, noise=noise_s)
        out = gather_actions(out, global_B)
        return out.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (builds the CUDA graph)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed region: K calls, inputs resident in HBM, CUDA events on the launch stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pol.last_launch_count() * args.steps
    # ---- e2e: host buffers through the C ABI
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    for _ in range(2):
        step_staged()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_staged()
    barrier()
    staged_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([ms, e2e_s * 1e3, staged_s * 1e3], dtype=torch.float64, device=dev)
    per_rank_ms = [ms / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x[0]) / args.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, staged_ms = float(t[0]), float(t[1]), float(t[2])

    # ---- in-situ kernel timing for the roofline figures (rank 0).
    # Dominant kernel = the CTA-pair tcgen05 GEMM at large M (ViT + prefill projections): direct launches, CUDA events on
    # the launch stream around every GEMM launch (each is 40-190 us, far above the event resolution).
    # The small-M launches of the denoise steps are a dependent chain of a-few-microsecond kernels: bracketing each with
    # events would time the events, so that class is timed from the graph: whole call minus the call stopped after the
    # prefill (option "stop_after" = 3).
    roof = roof_small = roof_expert = None
    lat = None
    if rank == 0 and not args.no_profile:
        def timed_graph(n):
            for _ in range(2):
                pol.infer_action(ids_d, pix_d, proprios=prop_d, noise=noise_d)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                pol.infer_action(ids_d, pix_d, proprios=prop_d, noise=noise_d)
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n
        t_full = timed_graph(10)
        pol.set_option("stop_after", 3)
        t_prefill = timed_graph(10)
        pol.set_option("stop_after", 0)
        pol.set_option("use_graph", 0)
        pol.infer_action(ids_d, pix_d, proprios=prop_d, noise=noise_d)  # warm direct path
        torch.cuda.synchronize()
        pol.set_option("profile_mtot", int(valid.sum()))
        pol.set_option("profile_gemm", 1)
        reps = 3
        for _ in range(reps):
            pol.infer_action(ids_d, pix_d, proprios=prop_d, noise=noise_d)
        cls = pol.profile_read_classes()
        pol.set_option("profile_gemm", 0)
        pol.set_option("use_graph", 1)
        step_ms = ms / args.steps
        pr, sg = cls["pair"], cls["single"]
        if pr["launches"]:
            ach = pr["flops"] / (pr["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_tcgen05_pair_kernel at M > 512 (cta_group::2 256x256 tiles: ViT + LLM-prefill projections)",
                    "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                    "frac_of_burst": ach / peaks["tflops_burst"],
                    "traffic": _ncu_traffic("gemm_tcgen05_pair_kernel@full_grid"),
                    "traffic_unit": "bytes per launch (ncu dram read+write, the same full-grid launches)",
                    "algorithmic_mbytes_per_launch": pr["bytes"] / pr["launches"] / 1e6, "peak_source": peaks["which"],
                    "launches_per_call": int(pr["launches"] / reps), "avg_launch_us": 1e3 * pr["ms"] / pr["launches"],
                    "algorithmic_gflop_per_launch": pr["flops"] / pr["launches"] / 1e9,
                    "ms_per_call": pr["ms"] / reps, "share_of_step": (pr["ms"] / reps) / step_ms}
        # the denoise steps (10 x 28 expert layers x {q|k|v, o, gate/up, down} + the action-encoder GEMMs), from the graph
        he, f, qd = cfg.hidden_e, cfg.mlp_e, cfg.n_heads * cfg.head_dim
        kvd = cfg.n_kv_heads * cfg.head_dim
        w_layer = he * (qd + 2 * kvd) + qd * he + 3 * he * f
        w_step = cfg.layers * w_layer + 2 * he * he
        rows = Bg * cfg.num_action_tokens
        den_ms = max(t_full - t_prefill, 1e-6)
        den_bytes = cfg.num_inference_steps * 2.0 * w_step
        den_flops = cfg.num_inference_steps * 2.0 * rows * w_step
        roof_small = {"bound": "hbm", "kernel": "the denoise steps: small-M (M = %d) tcgen05 GEMM chain of the action expert with its "
                                               "norm / attention / Euler launches; weights streamed once per step" % rows,
                      "achieved": den_bytes / (den_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                      "frac": den_bytes / (den_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "algorithmic_mbytes_per_call": den_bytes / 1e6, "ms_per_call": den_ms,
                      "how": "graph replay: whole call %.3f ms minus the call stopped after the prefill %.3f ms" % (t_full, t_prefill),
                      "us_per_expert_layer": 1e3 * den_ms / (cfg.num_inference_steps * cfg.layers),
                      "share_of_step": den_ms / t_full, "achieved_tflops": den_flops / (den_ms * 1e-3) / 1e12,
                      "gemm_launches_per_call_small_m": int(sg["launches"] / reps)}
        xp = cls["expert"]
        if xp["launches"]:
            ach = xp["bytes"] / (xp["ms"] * 1e-3) / 1e9
            roof_expert = {"bound": "hbm", "kernel": "expert_pass_kernel (persistent: 28 expert layers of one proprio / "
                           "denoise pass per launch, weights streamed once per pass)",
                           "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                           "traffic": _ncu_traffic("expert_pass_kernel"), "launches_per_call": int(xp["launches"] / reps),
                           "avg_launch_us": 1e3 * xp["ms"] / xp["launches"], "ms_per_call": xp["ms"] / reps,
                           "algorithmic_mbytes_per_launch": xp["bytes"] / xp["launches"] / 1e6,
                           "share_of_step": (xp["ms"] / reps) / step_ms,
                           "achieved_tflops": xp["flops"] / (xp["ms"] * 1e-3) / 1e12}
        # ---- p50 latency of a single-observation call (B=1), same handle, graph replay, CUDA events per call
        one = make_inputs(cfg, 1, 999)
        o_ids, o_pix = one["input_ids"].to(dev), one["pixel_values"].to(torch.bfloat16).to(dev)
        o_prop, o_noise = one["proprios"].to(torch.bfloat16).to(dev), one["noise"].to(torch.bfloat16).to(dev)
        for _ in range(3):
            pol.infer_action(o_ids, o_pix, proprios=o_prop, noise=o_noise)
        torch.cuda.synchronize()
        ts = []
        for _ in range(21):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pol.infer_action(o_ids, o_pix, proprios=o_prop, noise=o_noise)
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        lat = {"p50_ms": statistics.median(ts), "min_ms": min(ts), "calls": len(ts),
               "hbm_floor_ms": 16.3e9 / (peaks["hbm_gbs"] * 1e9) * 1e3}
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    obs_per_s = global_B * args.steps / (ms * 1e-3)
    value = obs_per_s * cfg.horizon_steps
    e2e_value = global_B * args.steps / (e2e_ms * 1e-3) * cfg.horizon_steps
    flops_call = pol.flops_per_call(valid)
    es = 2
    h2d = ids_h.numel() * 8 + (pix_h.numel() + prop_h.numel() + noise_h.numel()) * es
    d2h = Bg * cfg.horizon_steps * cfg.action_dim * es
    line = {
        "metric": "policy actions/s", "value": value, "unit": "actions/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.config == "vlaser_2b_bridge" and Bg == 64) else f"{args.config}, {Bg} observations per GPU",
                   "batch_per_gpu": Bg, "global_batch": global_B, "parallelism": f"dp{world}",
                   "valid_len": f"{int(valid.min())}-{int(valid.max())} of {cfg.max_image_text_tokens} (var-len packed prefix)",
                   "obs_per_s": obs_per_s, "actions_per_obs": cfg.horizon_steps,
                   "l2": "no explicit flush: weights (5 GB) + activations per call >> 126 MB L2",
                   "algorithmic_gflop_per_obs": flops_call / Bg / 1e9,
                   "frac_of_bf16_roofline_whole_call": (flops_call / (ms / args.steps * 1e-3) / 1e12) / peaks["tflops"]},
        "clocks": sampler.result(),
        "ms_per_step_by_rank": per_rank_ms,
        "e2e": {"value": e2e_value, "unit": "actions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "e2e_staged": {"value": global_B * args.steps / (staged_ms * 1e-3) * cfg.horizon_steps, "unit": "actions/s",
                       "what": "u8 640x480 camera frames from pinned host memory -> device Lanczos4 resize + normalise "
                               "(vlaser_op_stage_frames) -> infer_action -> actions to host",
                       "h2d_bytes_per_step": frames_h.numel() + ids_h.numel() * 8 + (prop_h.numel() + noise_h.numel()) * 2,
                       "d2h_bytes_per_step": d2h, "ms_per_step": staged_ms / args.steps},
        "gpu_launches": launches,
    }
    if roof is not None:
        line["roofline"] = roof
    if roof_small is not None:
        line["roofline_small_m"] = roof_small
    if roof_expert is not None:
        line["roofline_expert"] = roof_expert
    if lat is not None:
        line["latency_b1"] = lat
    if args.gpu_baseline != "none" and world == 1:
        # the reference algorithm on this GPU (bf16, dense padded prefix): what the hand-written kernels have to beat
        legs = []
        for kind in (["port-eager", "port-compile"] if args.gpu_baseline == "compile" else ["port-eager"]):
            for bb in (Bg, 1):
                try:
                    t_ms = _gpu_port_timing(cfg, dev, bb, 5 if bb > 1 else 10, kind == "port-compile")
                    legs.append({"value": bb * cfg.horizon_steps / (t_ms * 1e-3), "unit": "actions/s", "kind": kind,
                                 "batch": bb, "ms_per_call": t_ms, "dtype": "bf16"})
                except Exception as ex:  # the baseline leg must never take the bench line down
                    legs.append({"kind": kind, "batch": bb, "error": repr(ex)[:200]})
        line["gpu_eager_baseline"] = legs
    if want_cpu:
        times, cores = _cpu_oracle_timing(cfg, sd_cpu, 3, 1)
        per_obs = statistics.median(times)
        line["cpu_baseline"] = {"value": cfg.horizon_steps / per_obs, "unit": "actions/s", "cores": cores, "kind": "port",
                                "sample": "3 single-observation calls (B=1, fp32, padded S=384) of the oracle port after 1 warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="vlaser_2b_bridge", choices=sorted(PRESETS))
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong-scaling form of BASELINE configs[2]: the global batch is split evenly over the GPUs "
                         "(512 -> 256 / 128 / 64 per GPU at N = 2 / 4 / 8); overrides --batch-per-gpu")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--gpu-baseline", default="eager", choices=["none", "eager", "compile"],
                    help="N=1: also time the reference algorithm (oracle port) on the GPU in bf16, eager / + torch.compile")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
