#!/usr/bin/env python
"""Headline benchmark: res15 utterances/second, 1 s @ 16 kHz clips, waveform -> MFCC -> logits
(BASELINE.json `metric`; workload = `configs[1]`, res15 at batch 8192 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model res15] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one batch of synthetic waveforms per GPU.
  value   whole-job utterances/s with the waveforms already resident in HBM (device-timed,
          CUDA events, max over ranks); N > 1 is weak scaling (8192 utterances per GPU) and
          includes the per-step NCCL all-gather of logits + all-reduce of accuracy counts.
  e2e     same metric through the public API (model.forward_wave) with HOST buffers: pinned
          waveforms copied host->device and logits copied device->host inside the timed region.
  roofline  the dominant kernel (the C->C 3x3 convolution): algorithmic FLOPs per launch /
          its average launch duration, measured with CUDA events around every launch in a
          second pass over the same K steps.
  cpu_baseline  the oracle port (oracle/mfcc_ref.py + oracle/model_ref.py, the reference
          algorithm in numpy / PyTorch-CPU fp32) timed on this host's cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "res15 utterances/sec (1s 16kHz, MFCC+forward)"
UNIT = "utterances/s"
N_SAMPLES = 16000
FLOPS_PER_UTT = {  # SURVEY.md section 8d / BASELINE.md section 3 (2*MAC, padded taps counted)
    "res15": 1.917627e9, "res8": 0.074351e9, "res26": 0.878073e9, "res15_narrow": 0.342657e9,
    "res8_narrow": 0.014053e9, "res26_narrow": 0.157334e9, "cnn-trad-fpool3": 0.249187e9,
}


def committed_traffic(kernel_path, batch):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same command
    (profiles/*traffic*.json; the newest record whose kernel and batch match), or None."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if kernel_path in d.get("kernel", "") and d.get("batch_per_launch") == batch:
            best = d
    return None if best is None else {"dram_bytes_per_launch": best["dram_bytes_per_launch"], "source": best.get("source")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        # started BEFORE the warm-up (nvidia-smi needs ~0.3 s to produce its first row) and sampled every 20 ms; every
        # row is tagged with the host time it arrived at, stop() keeps the rows that fall inside the timed region
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, window=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows, note = [r for _, r in self.rows], None
        if window is not None:
            t0, t1 = window
            inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.03]
            if inside:
                rows = inside
            else:   # region shorter than the sampling period: nearest rows (the GPU is under the same load in the warm-up)
                rows = [r for t, r in self.rows if t0 - 0.25 <= t <= t1 + 0.25]
                note = "no sample fell inside the timed region; rows within 0.25 s of it"
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def cpu_reference_step(model_name, n_utts, seed):
    """The reference algorithm on the CPU: per-sample compute_mfccs + cat (collate_fn,
    audio_data_loader.py:26-29) then model(x) under no_grad (run/test.py:25-26)."""
    from honk2_b200 import synth
    from oracle import mfcc_ref, model_ref
    waves = synth.broadband(n_utts, N=N_SAMPLES, seed=seed)
    t0 = time.perf_counter()
    feats = torch.from_numpy(mfcc_ref.compute_mfccs_batch(waves))
    t1 = time.perf_counter()
    kind, cfg, sd = cpu_reference_step.model
    with torch.no_grad():
        logits = model_ref.forward(kind, sd, cfg, feats)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, logits


def cpu_setup(model_name):
    import honk2_b200
    from honk2_b200.zoo import model_config
    kind, cfg = model_config(model_name)
    m = honk2_b200.build_model(model_name)
    cpu_reference_step.model = (kind, cfg, {k: v.clone() for k, v in m.state_dict().items()})
    torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_setup(args.model)
    batch = args.ref_batch
    for i in range(args.warmup):
        cpu_reference_step(args.model, batch, seed=1000 + i)
    t_fe = t_model = 0.0
    t0 = time.perf_counter()
    for i in range(args.steps):
        a, b, _ = cpu_reference_step(args.model, batch, seed=2000 + i)
        t_fe += a; t_model += b
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    sample = (f"{args.steps} steps x {batch} synthetic 1 s clips: numpy restatement of compute_mfccs per sample + "
              f"PyTorch-CPU fp32 restatement of {args.model} forward; front-end {t_fe:.2f} s, model {t_model:.2f} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, batch):
    secs, frames = N_SAMPLES / 16000.0, 1 + N_SAMPLES // 160
    return {"workload": f"{args.model} inference, {batch} synthetic {secs:g} s / 16 kHz clips per GPU per step, "
                        f"waveform -> {frames}x40 MFCC -> logits (12 GSC classes), random-init weights (seed of the config)",
            "model_config": args.model, "batch_per_gpu": batch, "clip_samples": N_SAMPLES,
            "l2": "inputs larger than L2 (512 KB.. per step: %.0f MB of waveforms per GPU)" % (batch * N_SAMPLES * 4 / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="res15")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--clip-samples", type=int, default=16000,
                    help="samples per clip (16000 = the headline 1 s clips; 144000 = the hey_snips-shaped 9 s clips of config 5)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="concurrent chunk streams of the bf16 path (0 = library default)")
    ap.add_argument("--ref-batch", type=int, default=64)
    ap.add_argument("--e2e-sub-batch", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    global N_SAMPLES
    N_SAMPLES = args.clip_samples

    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.lanes:
        os.environ["HONK2_TC_LANES"] = str(args.lanes)

    import honk2_b200
    from honk2_b200 import AudioProcessor, synth
    from honk2_b200 import dist as kdist
    from honk2_b200.metric import Acc
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    precision = args.precision or os.environ.get("HONK2_BENCH_PRECISION", "bf16")
    model = honk2_b200.build_model(args.model, precision=precision).to(dev)
    if args.chunk:
        model.chunk = {"fp32": args.chunk, "bf16": args.chunk}
    fe = AudioProcessor()
    B = args.batch
    K, W = args.steps, max(args.warmup, 3)

    # synthetic data: a few distinct batches (each 524 MB > L2), device resident for `value`
    n_sets = 2
    host_sets = [torch.from_numpy(synth.broadband(B, N=N_SAMPLES, seed=100 * rank + s)).pin_memory() for s in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    targets = torch.randint(0, model.n_labels, (B,), device=dev)
    acc = Acc()

    def step_device(i):
        logits = model.forward_wave(dev_sets[i % n_sets], fe)
        acc.accumulate(logits, targets)
        if world > 1:
            full = kdist.all_gather_rows(logits, B * world)
            acc.all_reduce()
            return full
        return logits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for i in range(W):
            step_device(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        t_host0 = time.perf_counter()
        e0.record()
        for i in range(K):
            step_device(i)
            launches += model.last_launches(dev) + 1 + (2 if world > 1 else 0)
        e1.record()
        barrier()
        t_host1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop((t_host0, t_host1)) if rank == 0 else None
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        value = B * world * K / (ms / 1e3)

        # ---- e2e: host buffers, copies inside the timed region
        from honk2_b200.evaluate import HostPipeline
        host_logits = torch.empty((B, model.n_labels), dtype=torch.float32).pin_memory()
        pipe = HostPipeline(model, fe, N_SAMPLES, sub_batch=args.e2e_sub_batch, device=dev)

        def step_e2e(i):
            pipe(host_sets[i % n_sets], host_logits)

        for i in range(2):
            step_e2e(i)
        barrier()
        e0.record()
        for i in range(K):
            step_e2e(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
        e2e_value = B * world * K / (ms_e2e / 1e3)

        # ---- roofline of the dominant kernel: per-launch events in a second pass over K steps
        prof = honk2_b200.profile_layers(model, fe, dev_sets, K) if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    roof = None
    if prof is not None and prof["conv_launches"] > 0:
        flops_per_launch = prof["conv_flops_per_launch"]
        avg_s = prof["conv_ms"] / prof["conv_launches"] / 1e3
        achieved = flops_per_launch / avg_s / 1e12
        peak = pk["bf16_tflops_sustained"]
        tr = committed_traffic(prof.get("kernel_path", "?"), B)
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None if tr is None else tr["dram_bytes_per_launch"],
                "traffic_source": None if tr is None else tr["source"],
                "frac_of_burst_peak": achieved / pk["bf16_tflops"],
                "kernel": prof["conv_kernel"], "launches": prof["conv_launches"],
                "avg_launch_ms": prof["conv_ms"] / prof["conv_launches"], "peak_source": pk["source"] + " (sustained bf16)",
                "share_of_step": prof["conv_ms"] / max(prof["total_ms"], 1e-9),
                "frontend_ms_per_step": prof["frontend_ms"] / K, "other_ms_per_step": prof["other_ms"] / K}

    cpu = None
    if not args.no_cpu_baseline:
        cores = cpu_setup(args.model)
        cpu_reference_step(args.model, 8, seed=1)
        n, t_cpu, fe_s, mo_s = 0, 0.0, 0.0, 0.0
        while t_cpu < args.cpu_seconds and n < 64 * 64:
            t0 = time.perf_counter()
            a, b, _ = cpu_reference_step(args.model, args.ref_batch, seed=3000 + n)
            t_cpu += time.perf_counter() - t0
            fe_s += a; mo_s += b
            n += args.ref_batch
        cpu = {"value": n / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} synthetic 1 s clips in batches of {args.ref_batch}: numpy compute_mfccs restatement per sample "
                         f"({fe_s:.1f} s) + PyTorch-CPU fp32 {args.model} forward ({mo_s:.1f} s)"}

    # ---- streaming windows, reported next to the headline: B windows of 1 s at a 10 ms shift
    # (gsc_dev_config.json:62-63) from one resident stream; front-end alone and front-end + network.
    streaming = None
    if world == 1 and N_SAMPLES == 16000:
        with torch.no_grad():
            stream = torch.from_numpy(synth.broadband(1, N=B * 160 + N_SAMPLES, seed=77)[0]).to(dev)
            feats_s = torch.empty((B, fe.n_frames(N_SAMPLES), fe.n_mels), dtype=torch.float32, device=dev)

            def timed(fn, reps=5):
                for _ in range(2):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps

            ms_fe_stream = timed(lambda: fe.compute_mfccs_stream(stream, N_SAMPLES, 160, out=feats_s))
            ms_fe_batch = timed(lambda: fe.compute_mfccs_batch(dev_sets[0], out=feats_s))
            ms_all = timed(lambda: model(fe.compute_mfccs_stream(stream, N_SAMPLES, 160, out=feats_s)))
            streaming = {"windows_per_step": B, "window_samples": N_SAMPLES, "shift_samples": 160,
                         "frontend_ms_shared_frames": ms_fe_stream, "frontend_ms_per_window_batch": ms_fe_batch,
                         "windows_per_s_frontend_plus_network": B / (ms_all / 1e3)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": dict(workload_config(args, B), precision=precision, parallelism=f"dp{world}"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * N_SAMPLES * 4,
                    "d2h_bytes_per_step": B * model.n_labels * 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "streaming_windows": streaming,
            "tensor_frac_of_burst_peak_whole_step": FLOPS_PER_UTT.get(args.model, 0) * value / world / 1e12 / pk["bf16_tflops"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
