#!/usr/bin/env python
"""Headline benchmark: decoded audio-seconds per second of the VITS waveform decoder (HiFi-GAN Generator).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a decoder
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch eager port)

One "step" = one decode of a batch of synthetic latents (BASELINE.json config 3: 16 utterances x 10 s,
T = 862 frames, hop 256, 22.05 kHz, random-init weights of configs/finetune_speaker.json).  N > 1: one
process per GPU (torchrun), every rank decodes its own 16 utterances (utterance sharding, no collective
on the data path; "weak" scaling), time = max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, HOP = 22050, 256
FLOP_PER_FRAME = 614907904       # SURVEY.md section 8d / BASELINE.md section 3 (2 x 307 453 952 MAC)
FLOP_PER_UTT = 262144            # cond(g)


def load_traffic(launches_per_step):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 conv launches of one 16 x 10 s step,
    from the committed ncu pass (profiles/r01_launches_final.csv -> profiles/r01_traffic.json), per launch."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("conv_launches_per_step") != launches_per_step:
        return None  # the schedule changed since the capture: do not quote a stale number
    return d["conv_dram_bytes_per_step"] / launches_per_step


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.05):
                continue  # only samples taken DURING the timed region
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference(batch, frames, steps, warmup, threads=None):
    """The reference's CPU implementation of the path: torch eager conv1d/conv_transpose1d/leaky_relu/tanh with the
    weight-norm recompute per forward (oracle/generator_torch.py restates models.py:270-289 op for op)."""
    import numpy as np
    import torch
    import oracle
    from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the
    # reference to one thread)
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    hp = oracle.FINETUNE_SPEAKER
    sd = to_torch_state_dict(oracle.synth_state_dict(hp, 0, gain=2.0))
    rs = np.random.RandomState(1)
    z = torch.from_numpy(rs.standard_normal((batch, hp.initial_channel, frames)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((batch, hp.gin_channels, 1)).astype(np.float32))
    for _ in range(warmup):
        generator_forward_torch(hp, sd, z, g)
    t0 = time.perf_counter()
    for _ in range(steps):
        generator_forward_torch(hp, sd, z, g)
    dt = (time.perf_counter() - t0) / steps
    return batch * frames * HOP / SR / dt, dt, torch.get_num_threads()


def main():
    # stdout carries exactly ONE line, the JSON record.  Native libraries print there too (NCCL's "NCCL version ..."
    # banner is a C-level printf): point fd 1 at stderr for the whole run and keep the real stdout for emit().
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(record):
        os.write(real_stdout, (json.dumps(record) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true",
                    help="profiling runs (ncu): warm-up + the device-resident timed loop only, no JSON line")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    frames = -(-int(round(args.seconds * SR)) // HOP)  # ceil(sec * 22050 / 256): 10 s -> 862
    B = args.batch
    workload = "Generator decode, batch %d x %.0f s synthetic latents per GPU (T=%d frames, hop 256, 22.05 kHz), " \
               "finetune_speaker.json hyper-parameters" % (B, args.seconds, frames)
    config = {"workload": workload, "batch_per_gpu": B, "frames": frames,
              "l2": "working set ~%.1f GB of activations per step >> 126 MB L2; no explicit flush" %
                    (7 * B * frames * 16384 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: ONE utterance of the same workload per step (the CPU needs seconds per utterance)
        steps = max(1, min(args.steps, 3))
        val, dt, cores = cpu_reference(1, frames, steps, 1)
        line = {"impl": "reference", "metric": "decoded audio-sec/sec, VITS Generator", "value": val,
                "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                 "sample": "1 of the %d utterances (B=1, T=%d) per step, torch CPU eager fp32, "
                                           "weight-norm recomputed per forward like the reference" % (B, frames)},
                "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import vitsdec

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # random-init weights of the shipped architecture (no network for checkpoints): torch default init like the
    # reference's constructor, with weight_g perturbed so the weight-norm fold is exercised
    cargs, ckw = vitsdec.generator_args()  # configs/finetune_speaker.json model block
    torch.manual_seed(1234)
    G = vitsdec.Generator(*cargs, **ckw)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith("weight_g"):
                p.mul_(torch.empty_like(p).uniform_(0.5, 1.5))
    G = G.to(dev).eval()
    for kv in filter(None, os.environ.get("VITSDEC_OPTS", "").split(",")):  # experiment knob, e.g. "fold=0,fuse_pairs=0"
        k, v = kv.split("=")
        G.set_option(k, int(v))
    G.assume_frozen = True
    rs = np.random.RandomState(1 + rank)
    z_host = torch.from_numpy(rs.standard_normal((B, cargs[0], frames)).astype(np.float32)).pin_memory()
    g_host = torch.from_numpy(rs.standard_normal((B, ckw["gin_channels"], 1)).astype(np.float32)).pin_memory()
    z = z_host.to(dev)
    g = g_host.to(dev)
    out_host = torch.empty((B, 1, frames * HOP), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to start: launch it before the warm-up, keep only timed-region samples
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = G(z, g)
        barrier()
        # ---- device-resident timing (value) + live conv-kernel timing (roofline)
        G.set_option("profile", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_begin = sampler.mark()
        e0.record()
        for _ in range(args.steps):
            y = G(z, g)
        e1.record()
        barrier()
        t_end = sampler.mark()
        clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        conv_ms, conv_launches = G.profile_read()
        G.set_option("profile", 0)
        launches = G.last_launch_count() * args.steps
        if args.quick:
            sys.stderr.write("quick: %.3f ms/step device-resident (%d steps)\n" % (ms_total / args.steps, args.steps))
            if world > 1:
                dist.destroy_process_group()
            return 0

        # ---- end to end through the public API with HOST buffers (pinned): H2D + decode + D2H every step, two batches
        # in flight (vitsdec.HostPipeline: the copies of neighbouring steps overlap the decode of the current one)
        pipe = vitsdec.HostPipeline(G, depth=2)
        outs_host = [out_host, torch.empty_like(out_host).pin_memory()]
        for i in range(8):   # each slot's plan reaches its graph (captured at the third use) before the timed region
            pipe.submit(z_host, g_host, outs_host[i % 2])
        pipe.wait_all()
        barrier()
        e0.record()
        for i in range(args.steps):
            pipe.submit(z_host, g_host, outs_host[i % 2])
        pipe.join()
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))

        # the stage before the decoder: flow(z_p, reverse) on the same batch, timed separately -- it is
        # not part of the metric, which is the Generator decode alone
        Fl = vitsdec.ResidualCouplingBlock(cargs[0], 192, 5, 1, 4, gin_channels=ckw["gin_channels"])
        for name, p in Fl.named_parameters():
            if name.endswith("post.weight"):
                p.uniform_(-0.07, 0.07)  # the reference zero-initialises post: give the couplings something to do
        Fl = Fl.to(dev).eval()
        Fl.assume_frozen = True
        ymask = torch.ones((B, 1, frames), device=dev)
        for _ in range(3):
            zf = Fl(z, ymask, g=g, reverse=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            zf = Fl(z, ymask, g=g, reverse=True)
        e1.record()
        barrier()
        flow_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps

        # BASELINE config 2 (batch 1, 2 s latent): latency of one decode, launch- and prologue-bound (extra key)
        z1, g1 = z[:1, :, :173].contiguous(), g[:1]
        for _ in range(5):
            G(z1, g1)
        barrier()
        e0.record()
        for _ in range(50):
            G(z1, g1)
        e1.record()
        barrier()
        lat_ms = max_over_ranks(e0.elapsed_time(e1)) / 50

        # option "fp16" (fp16 instead of bf16 operands / stored activations; same kernels, same FLOPs): a short timed
        # loop as an extra key.  The headline above is the bf16 mode BASELINE.json names.
        G.set_option("fp16", 1)
        for _ in range(4):
            G(z, g)
        barrier()
        e0.record()
        for _ in range(5):
            G(z, g)
        e1.record()
        barrier()
        fp16_ms = max_over_ranks(e0.elapsed_time(e1)) / 5
        G.set_option("fp16", 0)

        gather_ms = None
        if world > 1:  # the optional final waveform gather (north_star): timed separately, not on the data path
            full = torch.empty((world * B, 1, frames * HOP), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full, y)
            barrier()
            e0.record()
            dist.all_gather_into_tensor(full, y)
            e1.record()
            barrier()
            gather_ms = max_over_ranks(e0.elapsed_time(e1))

    audio_s = world * B * frames * HOP / SR
    ms_step = ms_total / args.steps
    value = audio_s / (ms_step / 1e3)
    e2e_val = audio_s / (ms_e2e / args.steps / 1e3)
    burst, sustained, peak_src = load_peaks()
    conv_flops_step = B * frames * FLOP_PER_FRAME  # per rank; every conv of the path runs in a tcgen05 kernel (conv_post too)
    conv_ms_step = conv_ms / args.steps
    achieved = conv_flops_step / (conv_ms_step / 1e3) / 1e12
    launches_conv = conv_launches // max(1, args.steps)
    line = {
        "metric": "decoded audio-sec/sec, VITS Generator", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
        "tflops": (world * B * (frames * FLOP_PER_FRAME + FLOP_PER_UTT)) / (ms_step / 1e3) / 1e12,
        "e2e": {"value": e2e_val, "unit": "audio-s/s",
                "h2d_bytes_per_step": int(z_host.numel() * 4 + g_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor",
                     "kernel": "conv_tc_kernel + conv_pair_kernel (tcgen05 implicit-GEMM convs: %d launches/step, "
                               ">98%% of step time; figures are per launch, averaged over them)" % launches_conv,
                     "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                     "frac_of_sustained": achieved / sustained, "peak_source": peak_src + ", bf16 dense burst",
                     "flop_per_launch": conv_flops_step / max(1, launches_conv),
                     "ms_per_launch": conv_ms_step / max(1, launches_conv),
                     "conv_ms_per_step": conv_ms_step,
                     "traffic": load_traffic(launches_conv) if (B, frames) == (16, 862) else None,
                     "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"},
        "clocks": clocks,
    }
    line["flow_reverse_ms_per_step"] = flow_ms
    line["latency_b1_2s_ms"] = lat_ms
    line["fp16_mode_ms_per_step"] = fp16_ms
    if gather_ms is not None:
        line["waveform_gather_ms"] = gather_ms
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            val, dt, cores = cpu_reference(1, frames, 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                    "sample": "1 of the %d utterances (B=1, T=%d), 2 timed passes after 1 warm-up, "
                                              "torch CPU eager fp32 (oracle/generator_torch.py)" % (B, frames)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
