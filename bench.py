#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Silero-VAD stream engine.

Metric (BASELINE.json): audio-seconds per wall-second over N concurrent 16 kHz streams
(Silero v5), plus p99 latency of one batched frame step.  Workload = configs[1]:
4,096 concurrent 16 kHz streams per GPU, every step advances every stream by
`--frames-per-step` 512-sample frames (hop 512).  Weak scaling: streams per GPU fixed.

  python bench.py [--gpus N] [--steps K] [--warmup W]              our arm
  python bench.py --impl reference [...]                            CPU arm (oracle port, all host threads)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))

FLOP_FRAME = 1_127_936          # v5/16k nominal FLOP per 512-sample frame
FLOP_FRONTEND = 865_536         # STFT + encoder.0-3 share (432,768 MAC)
FLOP_RECURRENT = 262_400        # LSTM + decoder share (131,200 MAC)
SMS, FP32_LANES = 148, 128


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def ncu_traffic_bytes(kernel: str):
    """DRAM read+write bytes per launch of `kernel` from the newest committed ncu --set full summary
    (profiles/*.json, written by tools/ncu_summary.py from the same bench command); None if absent."""
    best = None
    for f in sorted((ROOT / "profiles").glob("*.json")):
        if "warm_caches" in f.name:      # --cache-control none captures re-read the replayed inputs from L2: not a DRAM figure
            continue
        try:
            d = json.loads(f.read_text())
            for k in d.get("full_capture", []):
                if kernel in k.get("kernel", ""):
                    tot = 0.0
                    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                        val, _, unit = k.get(key, "0 byte").partition(" ")
                        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit.strip(), 1)
                        tot += float(val) * mult
                    best = {"bytes": tot, "source": f"profiles/{f.name}"}
        except Exception:
            continue
    return best


def synth_audio(n_streams: int, n_samples: int, seed: int) -> np.ndarray:
    """Vectorised version of the tests' signal recipe: per-stream noise floor plus gated
    harmonic 'voice' bursts (reference examples/probability_demo.py:60-67), float32."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float32) / 16000.0
    sigma = np.array([0.005, 0.02, 0.1], np.float32)[np.arange(n_streams) % 3][:, None]
    level = np.array([0.3, 0.7], np.float32)[(np.arange(n_streams) // 3) % 2][:, None]
    f0 = rng.uniform(110, 220, size=(n_streams, 1)).astype(np.float32)
    out = rng.standard_normal((n_streams, n_samples), dtype=np.float32) * sigma
    ph = 2 * np.pi * f0 * t[None, :]
    voice = level * (0.4 * np.sin(ph) + 0.3 * np.sin(2 * ph) + 0.2 * np.sin(4 * ph))
    period = rng.uniform(1.0, 4.0, size=(n_streams, 1)).astype(np.float32)
    phase = rng.uniform(0, 1, size=(n_streams, 1)).astype(np.float32)
    gate = (((t[None, :] / period) + phase) % 1.0) < 0.5
    out += np.where(gate, voice, 0.0).astype(np.float32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm

def cpu_port(native=True):
    sys.path.insert(0, str(ROOT / "oracle"))
    from vad_oracle import RefLib, RefV5, v5_blob
    lib = RefLib(native=native)
    blob = v5_blob(str(ROOT / "cutter-vad_b200" / "real_time_vad" / "models" / "silero_vad_v5.onnx"))
    return lib, RefV5(blob, lib)


def cpu_baseline_sample(n_streams: int, target_s: float = 12.0):
    """Oracle port (plain C + OpenMP, every host thread) on a bounded sample of the workload."""
    lib, ref = cpu_port()
    cores = host_threads()
    n = min(n_streams, 1024)
    audio = synth_audio(n, 512 * 8, seed=99)
    t0 = time.perf_counter()
    ref.run(audio, 8, nthreads=cores)
    dt = time.perf_counter() - t0
    rate = n * 8 / dt
    frames = int(max(8, min(4000, rate * target_s / n)))
    audio = synth_audio(n, 512 * frames, seed=100)
    t0 = time.perf_counter()
    ref.run(audio, frames, nthreads=cores)
    dt = time.perf_counter() - t0
    fps = n * frames / dt
    return {"value": fps * 0.032, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "frames_per_s": fps,
            "sample": f"{n} streams x {frames} frames (hop 512) of the same synthetic recipe, {dt:.1f} s, "
                      f"oracle/silero_ref.c -O3 -march=native + OpenMP; CPU restatement, not onnxruntime"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, ref = cpu_port()
    cores = host_threads()
    n, F = args.streams, args.frames_per_step
    pool = [synth_audio(n, 512 * F, seed=1000 + i) for i in range(4)]
    h = np.zeros((n, 128), np.float32)
    c = np.zeros((n, 128), np.float32)
    for i in range(args.warmup):
        ref.run(pool[i % 4], F, h=h, c=c, nthreads=cores)
    lat = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        t1 = time.perf_counter()
        ref.run(pool[i % 4], F, h=h, c=c, nthreads=cores)
        lat.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    value = n * F * args.steps * 0.032 / dt
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, pool_mb=None),
        "p99_step_ms": 1e3 * float(np.percentile(lat, 99)),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"every step = full workload ({n} streams x {F} frames); oracle/silero_ref.c "
                                   f"(plain C, OpenMP, {cores} threads); CPU restatement of the reference's "
                                   "onnxruntime path, onnxruntime itself is absent from this image"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, pool_mb):
    model = getattr(args, "model", "v5")
    rate = getattr(args, "src_rate", 16000)
    return {"workload": f"Silero {model}, {args.streams} concurrent {rate // 1000} kHz streams per GPU batched per "
                        f"frame step" + (" (BASELINE.json configs[1])" if (model, rate) == ("v5", 16000) else
                                         ", mixed 24/48 kHz by stream parity, resampled to 16 kHz on the GPU in one step "
                                         "(BASELINE.json configs[3] per-GPU share)" if getattr(args, "mixed_rates", False) else
                                         ", resampled to 16 kHz on the GPU" if rate != 16000 else ""),
            "model": model, "src_rate": rate,
            "streams_per_gpu": args.streams, "frames_per_step": args.frames_per_step, "hop": 512,
            "frame_len": 512, "denoise": True, "state_machine": True, "math": getattr(args, "math", "fp32"),
            "l2": (f"inputs cycle through a pool of distinct step buffers totalling {pool_mb} MB (> 126 MB L2)"
                   if pool_mb else "n/a (CPU arm)")}


# ----------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from real_time_vad.engine import capi
    from real_time_vad.engine.stream_engine import StreamEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, F = args.streams, args.frames_per_step
    n_in = args.src_rate * 512 // 16000          # source samples per model frame
    mixed = bool(getattr(args, "mixed_rates", False))
    if mixed:                                      # BASELINE.json configs[3]: 24 / 48 kHz by stream parity, one step
        args.src_rate = 48000
        n_in = 1536                                # row length; 24 kHz streams use the first half of their row
        rates_np = np.where(np.arange(n) & 1, 48000, 24000).astype(np.int32)
    step_samples = n_in * F
    step_bytes = n * step_samples * 4
    pool_n = max(2, int(np.ceil(300e6 / step_bytes)))  # > 2x L2 worth of distinct inputs
    pool_n = min(pool_n, 64)
    peaks, peaks_src = load_peaks()

    eng = StreamEngine(args.model, max_streams=n, device=local)
    math = args.math
    eng.set_math(math)
    flop_frame = FLOP_FRAME if args.model == "v5" else 1_379_280
    flop_fe = FLOP_FRONTEND if args.model == "v5" else 1_379_280 - 2 * 65_600
    flop_rec = FLOP_RECURRENT if args.model == "v5" else 2 * 65_600
    if args.src_rate != 16000:
        rs_flop = 2 * 512 * (n_in if not mixed else (768 + 1536) // 2)
        flop_fe += rs_flop                         # the resampling GEMM runs ahead of the front end
        flop_frame += rs_flop
    eng.configure(enable_denoising=True)
    stream = torch.cuda.Stream(device=local)
    eng.set_stream(stream.cuda_stream)

    # ---- synthetic inputs: pinned host pool (e2e) and device pool (kernel-only)
    host_pool = []
    for i in range(pool_n):
        t = torch.empty((n, step_samples), dtype=torch.float32).pin_memory()
        t.numpy()[:] = synth_audio(n, step_samples, seed=rank * 7919 + i)
        host_pool.append(t)
    dev_pool = [t.to(f"cuda:{local}", non_blocking=False) for t in host_pool]
    d_probs = torch.zeros((n, F), dtype=torch.float32, device=f"cuda:{local}")
    d_flags = torch.zeros((n, F), dtype=torch.uint8, device=f"cuda:{local}")
    d_events = torch.zeros((max(16, 2 * n * F) * 24,), dtype=torch.uint8, device=f"cuda:{local}")
    d_nev = torch.zeros((1,), dtype=torch.int32, device=f"cuda:{local}")

    d_rates = torch.from_numpy(rates_np).to(f"cuda:{local}") if mixed else None
    rate_kw = {"src_rates": rates_np, "max_frames": F} if mixed else {"src_rate": args.src_rate}

    def dev_args(buf):
        a = capi.StepArgs()
        a.n_streams = n
        a.audio = buf.data_ptr()
        a.pcm_format = capi.PCM_F32
        a.stream_stride = step_samples
        a.max_frames = F
        a.frame_len = n_in
        a.hop = n_in
        a.src_rate = args.src_rate
        if mixed:
            a.src_rates = d_rates.data_ptr()
        a.probs_out = d_probs.data_ptr()
        a.flags_out = d_flags.data_ptr()
        a.events_out = d_events.data_ptr()
        a.max_events = max(16, 2 * n * F)
        a.n_events_out = d_nev.data_ptr()
        return a

    dargs = [dev_args(b) for b in dev_pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ================= kernel-only (`value`): inputs resident in HBM, CUDA events on the engine's stream
    eng.reset()
    for i in range(args.warmup):
        eng.step_device(dargs[i % pool_n])
    eng.sync()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            eng.step_device(dargs[(args.warmup + i) % pool_n])
        e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    # per-kernel durations for the roofline: the same K steps again with an event between the two kernels
    # (kept out of the pass above: the events would sit between the kernels and serialise their overlap)
    eng.set_timing(True)
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            eng.step_device(dargs[(args.warmup + i) % pool_n])
    barrier()
    fe_ms, rec_ms, n_timed = eng.read_timing()
    eng.set_timing(False)

    # per-step latency distribution (device side), one step at a time
    lat = []
    for i in range(min(args.steps, 200)):
        a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        eng.step_device(dargs[i % pool_n])
        a1.record(stream)
        a1.synchronize()
        lat.append(a0.elapsed_time(a1))

    # ================= end to end (`e2e`): public host-buffer call, H2D + D2H inside the timed region
    eng.reset()
    host_np = [t.numpy() for t in host_pool]
    for i in range(args.warmup):
        eng.step(host_np[i % pool_n], **rate_kw)
    barrier()
    # (a) blocking calls: one step at a time -> per-call latency distribution
    e2e_lat = []
    t0 = time.perf_counter()
    for i in range(min(args.steps, 300)):
        t1 = time.perf_counter()
        r = eng.step(host_np[(args.warmup + i) % pool_n], **rate_kw)
        e2e_lat.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    e2e_blocking_s = (time.perf_counter() - t0) / min(args.steps, 300)
    # (b) pipelined calls (submit steps i+1, i+2 before collecting step i): the throughput figure.
    #     Every step's H2D of its inputs and D2H of its results is inside the timed region.
    eng.reset()
    barrier()
    from collections import deque
    depth = 3                                      # steps in flight (the engine allows 4)
    inflight = deque()
    t0 = time.perf_counter()
    for i in range(args.steps):
        inflight.append(eng.submit(host_np[(args.warmup + i) % pool_n], **rate_kw))
        if len(inflight) == depth:
            r = inflight.popleft().collect()
    while inflight:
        r = inflight.popleft().collect()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_events = len(r.events)
    # (c) the same pipelined loop with 16-bit PCM in the host buffers (the websocket service's wire format,
    #     vad_websocket_server.py:341; converted with /32767.0f in the frame loader): half the PCIe bytes
    e2e_s16_s = None
    if args.src_rate == 16000:
        s16_pool = []
        for i in range(min(pool_n, 8)):
            t = torch.empty((n, step_samples), dtype=torch.int16).pin_memory()
            t.numpy()[:] = np.clip(np.rint(host_np[i] * 32767.0), -32768, 32767).astype(np.int16)
            s16_pool.append(t.numpy())
        eng.reset()
        for i in range(args.warmup):
            eng.step(s16_pool[i % len(s16_pool)], pcm_format=capi.PCM_S16_32767)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            inflight.append(eng.submit(s16_pool[i % len(s16_pool)], pcm_format=capi.PCM_S16_32767))
            if len(inflight) == depth:
                inflight.popleft().collect()
        while inflight:
            inflight.popleft().collect()
        torch.cuda.synchronize()
        e2e_s16_s = time.perf_counter() - t0
    # raw pinned host -> device copy rate of one step's input, for context (e2e is transfer-bound)
    h0 = torch.cuda.Event(enable_timing=True)
    h1 = torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(10):
        dev_pool[i % pool_n].copy_(host_pool[i % pool_n], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = 10 * step_bytes / (h0.elapsed_time(h1) * 1e-3) / 1e9
    clocks = sampler.stop()

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(times[0]), float(times[1])

    audio_s_per_step = n * F * 0.032
    total_audio = audio_s_per_step * args.steps * world
    value = total_audio / (dev_ms_max * 1e-3)
    e2e_value = total_audio / (e2e_ms_max * 1e-3)
    frames_per_step = n * F
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    peak_fp32 = SMS * FP32_LANES * 2 * sm_max * 1e6 / 1e12
    fe_avg_s = fe_ms / max(n_timed, 1) * 1e-3
    rec_avg_s = rec_ms / max(n_timed, 1) * 1e-3
    fe_tflops = frames_per_step * flop_fe / fe_avg_s / 1e12 if fe_avg_s > 0 else 0.0
    step_tflops = frames_per_step * flop_frame / (dev_ms / args.steps * 1e-3) / 1e12

    fused = math in ("tc", "tc16") and args.model == "v5" and F == 1 and os.environ.get("CVAD_FUSE", "1") != "0"
    h16 = fused and math == "tc16"      # FP16 two-way split: 3 tensor-core products per MAC (BF16 split: 6)
    # chained one-frame steps (cvad_step_device on 16 kHz input): the timed region holds K launches of ONE kernel and
    # nothing else -- no memset, no event -- each scheduled by programmatic dependent launch while its predecessor
    # drains, so the kernel's average launch duration over the timed region is region / launches.  The duration of a
    # launch bracketed by its own events (no overlap with its neighbours) is reported beside it.
    chained = (math in ("tc", "tc16") and args.model == "v5" and F == 1 and args.src_rate == 16000 and not mixed
               and os.environ.get("CVAD_FUSE", "1") != "0" and os.environ.get("CVAD_CHAIN", "1") != "0"
               and launches == args.steps)
    isolated_ms = fe_avg_s * 1e3
    if fused:
        # one kernel does the whole step (front end + LSTM step + state machine): its FLOPs are the frame's
        flop_fe = flop_frame
        if chained:
            fe_avg_s = dev_ms / launches * 1e-3
        fe_tflops = frames_per_step * flop_fe / fe_avg_s / 1e12 if fe_avg_s > 0 else 0.0
    if rank == 0:
        tc = math in ("tc", "tc16")
        fe_kernel = (("v4tc_stft_kernel+v4_frontend_kernel" if args.model == "v4" else
                      "v5tc_frontend_kernel<FUSED,H16>" if h16 else
                      "v5tc_frontend_kernel<FUSED>" if fused else "v5tc_frontend_kernel") if tc
                     else f"{args.model}_frontend_kernel") + ("+resample_kernel" if args.src_rate != 16000 else "")
        # the ncu capture of exactly this kernel variant (template arguments <DBG, FUSED, H16>)
        variant = ("v5tc_frontend_kernel<0, 1, 1>" if h16 else "v5tc_frontend_kernel<0, 1>" if fused else
                   "v5tc_frontend_kernel<0, 0" if tc else "v5_frontend_kernel")
        traffic = (ncu_traffic_bytes(variant)
                   if (n == 4096 and F == 1 and args.model == "v5" and args.src_rate == 16000) else None)
        peak_bf16 = float(peaks.get("bf16_tflops", 1645.0))
        rec_tflops = frames_per_step * flop_rec / rec_avg_s / 1e12 if rec_avg_s else 0.0
        fe_weight_bytes = ((622_592 + 524_288) if h16 else (933_888 + (786_432 if fused else 0))) if tc else 156032 * 4
        roofline = {
            # tc: every algorithmic MAC is executed as 6 BF16 tensor-core products (3-way operand split) -- 3 FP16
            # products (2-way split, per-stream scaling) in the tc16 build of the fused kernel -- so the executed
            # rate is 6x / 3x `achieved`; the path is bound by shared-memory operand bandwidth and by the serial
            # loader -> MMA -> epilogue chain of one tile per SM, not by the tensor pipe
            "bound": "tensor" if tc else "fp32_ffma", "kernel": fe_kernel,
            "achieved": fe_tflops, "peak": peak_bf16 if tc else peak_fp32, "unit": "TFLOP/s",
            "frac": fe_tflops / (peak_bf16 if tc else peak_fp32),
            "traffic": (traffic or {}).get("bytes"), "traffic_source": (traffic or {}).get("source"),
            "algorithmic_bytes_per_launch": frames_per_step * 2048 + fe_weight_bytes +
            frames_per_step * ((2 * 1024 + 5) if fused else 768 if tc else 512),
            "peak_source": (f"{peaks_src} MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if tc else
                            f"148 SMs x 128 FP32 lanes x 2 x sm_max_mhz {sm_max:.0f} ({peaks_src} "
                            "MEASURED_PEAKS.json clock); the path is FP32-FFMA bound, not HBM or tensor bound"),
            "algorithmic_flop_per_frame": flop_fe, "frames_per_launch": frames_per_step,
            "avg_launch_ms": fe_avg_s * 1e3,
            "isolated_launch_ms": isolated_ms,
            "kernel_timing": ("avg_launch_ms = timed region (CUDA events on the engine's stream) / launches: the region holds "
                              "nothing but this kernel's K launches, chained by programmatic dependent launch; "
                              "isolated_launch_ms = second pass with events around every launch (no overlap)" if chained else
                              "second pass over the same K steps with a CUDA event between the two kernels"),
            "recurrent_kernel": ({"kernel": "(fused into the kernel above for one-frame steps)", "avg_launch_ms": 0.0,
                                  "achieved": 0.0, "algorithmic_flop_per_frame": 0} if fused else
                                 {"kernel": "v5tc_recurrent_kernel" if tc else f"{args.model}_recurrent_kernel",
                                  "avg_launch_ms": rec_avg_s * 1e3, "achieved": rec_tflops,
                                  "algorithmic_flop_per_frame": flop_rec}),
            "whole_step": {"achieved": step_tflops, "frac": step_tflops / (peak_bf16 if tc else peak_fp32),
                           "algorithmic_flop_per_frame": flop_frame},
            "vs_fp32_ffma_peak": {"peak": peak_fp32, "kernel_frac": fe_tflops / peak_fp32,
                                  "whole_step_frac": step_tflops / peak_fp32},
            "hbm": {"algorithmic_bytes_per_frame": 2048 + 2 * 1024 + 4 + 1,
                    "achieved_gbs": frames_per_step * (2048 + 2048 + 5) / (dev_ms / args.steps * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs")},
        }
        if tc:
            products = 3 if h16 else 6
            roofline["tensor_products_per_mac"] = products
            roofline["executed_bf16_tflops"] = products * fe_tflops
            roofline["executed_frac"] = products * fe_tflops / peak_bf16
        cpu = None if args.skip_cpu else cpu_baseline_sample(n)
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16x2 operands with per-stream scaling, f32 accumulate (FP32-equivalent)" if h16 else
                      "bf16x3 operands, f32 accumulate (FP32-equivalent)" if tc else "f32"),
            "data": "synthetic", "config": workload_config(args, pool_mb=round(pool_n * step_bytes / 1e6)),
            "p99_step_ms": float(np.percentile(lat, 99)), "p50_step_ms": float(np.percentile(lat, 50)),
            "frames_per_s": value / 0.032,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": step_bytes,
                    "d2h_bytes_per_step": n * F * 5 + n * 4 + 4, "ms_per_step": e2e_ms_max / args.steps,
                    "p99_step_ms": 1e3 * float(np.percentile(e2e_lat, 99)),
                    "blocking_ms_per_step": 1e3 * e2e_blocking_s,
                    "blocking_value": audio_s_per_step / e2e_blocking_s,
                    "pinned_h2d_gbs": h2d_gbs, "transfer_bound_value": audio_s_per_step * world / (step_bytes / (h2d_gbs * 1e9)),
                    "s16": (None if e2e_s16_s is None else
                            {"value": audio_s_per_step * args.steps * world / e2e_s16_s, "h2d_bytes_per_step": step_bytes // 2,
                             "pcm": "int16 / 32767.0f in the frame loader (CVAD_PCM_S16_32767)"}),
                    "api": "StreamEngine.submit/collect -> cvad_step_submit/cvad_step_collect, three steps in flight, "
                           "float32 PCM in pinned host buffers; p99_step_ms and blocking_* are the one-call-at-a-time "
                           "StreamEngine.step figures", "events_last_step": e2e_events},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line there.
    Send fd 1 to stderr for the duration of the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores it)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--streams", type=int, default=4096, help="concurrent streams per GPU")
    ap.add_argument("--frames-per-step", type=int, default=1, help="512-sample frames per stream per step")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--mixed-rates", action="store_true",
                    help="BASELINE.json configs[3] per-GPU share: 24 / 48 kHz streams by parity, resampled in one step")
    ap.add_argument("--math", choices=["tc16", "tc", "fp32"], default=None,
                    help="GEMM arithmetic: tc16 (v5 default) = tcgen05 tensor cores, FP16 2-way split with per-stream scaling for "
                         "one-frame steps; tc = BF16 3-way split; fp32 (v4 default) = packed FP32 FMA")
    ap.add_argument("--model", choices=["v5", "v4"], default="v5", help="v5 = headline (configs[1]); v4 = configs[2]")
    ap.add_argument("--src-rate", type=int, default=16000, choices=[8000, 16000, 24000, 48000],
                    help="source rate of the synthetic streams; != 16000 adds the GPU resampler (configs[2..3])")
    args = ap.parse_args()
    if args.math is None:
        args.math = "tc16" if args.model == "v5" else "fp32"
    args.warmup = max(args.warmup, 3)
    _protect_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
