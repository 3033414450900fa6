#!/usr/bin/env python
"""Headline benchmark: decoded audio-seconds per second of the VITS waveform decoder (HiFi-GAN Generator).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a decoder
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference Generator on the host cores

One "step" = one decode of a batch of synthetic latents (BASELINE.json config 3: 16 utterances x 10 s,
T = 862 frames, hop 256, 22.05 kHz, random-init weights of configs/finetune_speaker.json).  N > 1: one
process per GPU (torchrun), every rank decodes its own 16 utterances (utterance sharding, no collective
on the data path; "weak" scaling), time = max over ranks.  Prints ONE JSON line on rank 0.

Extra keys of the same line (outside the headline metric): BASELINE config 4 as stated (256 x 10 s sharded 256/N per
GPU, strong scaling, final waveform gather inside the timed region: "config4_strong"), config 5 (60 s utterance decoded
in 512-frame chunks with recompute halos: "chunked_60s_audio_s_per_s"), config 2 latency with the module's DEFAULT
settings, config 1 (the reference's SynthesizerTrn.infer on the host cores with the decoder's share) and the survey's
"kernel to beat": the unmodified reference Generator through torch-eager cuDNN on the same B200 ("cudnn_eager_*").
The reference itself is the byte-for-byte install under baseline/_ref (baseline/install_ref.py).
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
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 conv launches of one 16 x 10 s step, per
    launch, from the committed ncu pass (tools/final_profiles.sh -> profiles/r02_traffic.json).  An ncu pass cannot run
    inside the timed run, so the capture carries the digest of the kernel sources it was taken on (build._digest()) and
    its launch count: a number from other kernels or another schedule is not quoted (null)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("conv_launches_per_step") != launches_per_step:
        return None  # the schedule changed since the capture: do not quote a stale number
    try:
        import importlib
        digest = importlib.import_module("personalized_text-to-speech_b200.build")._digest()
    except Exception:
        return None
    if d.get("csrc_digest") != digest:
        return None  # the kernels changed since the capture
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


def reference_dir():
    d = os.path.join(ROOT, "baseline", "_ref")
    return d if os.path.exists(os.path.join(d, "models.py")) else None


def import_reference(name):
    d = reference_dir()
    if d not in sys.path:
        sys.path.insert(0, d)
    import importlib
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)   # old-style weight_norm deprecation notice
    return importlib.import_module(name)


def reference_generator(device="cpu"):
    """The reference's own decoder, unmodified (baseline/_ref/models.py:244-296), built like models.py:447 builds it,
    with the SAME weights bench.py gives the B200 decoder (seed 1234, weight_g perturbed)."""
    import torch
    import vitsdec
    models = import_reference("models")
    cargs, ckw = vitsdec.generator_args()
    torch.manual_seed(1234)
    G = models.Generator(*cargs, **ckw)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith("weight_g"):
                p.mul_(torch.empty_like(p).uniform_(0.5, 1.5))
    return G.to(device).eval()


def cpu_reference(batch, frames, steps, warmup, threads=None, budget_s=None):
    """The reference's CPU implementation of the path on the host cores.  kind "reference": the unmodified
    models.Generator from baseline/_ref; kind "port" (only when that install is missing): oracle/generator_torch.py,
    the op-for-op torch restatement that tests/test_oracle.py pins bit-identical to it.
    Returns (audio-s/s, seconds per pass, threads, kind, passes timed)."""
    import numpy as np
    import torch
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the
    # reference to one thread)
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    rs = np.random.RandomState(1)
    z = torch.from_numpy(rs.standard_normal((batch, 192, frames)).astype(np.float32))
    g = torch.from_numpy(rs.standard_normal((batch, 256, 1)).astype(np.float32))
    if reference_dir() is not None:
        G = reference_generator("cpu")
        kind = "reference"

        def run():
            with torch.no_grad():
                return G(z, g=g)
    else:
        import oracle
        from oracle.generator_torch import generator_forward_torch, to_torch_state_dict
        hp = oracle.FINETUNE_SPEAKER
        sd = to_torch_state_dict(oracle.synth_state_dict(hp, 0, gain=2.0))
        kind = "port"

        def run():
            return generator_forward_torch(hp, sd, z, g)
    t0 = time.perf_counter()
    for _ in range(warmup):
        run()
    t_warm = (time.perf_counter() - t0) / max(1, warmup)
    if budget_s is not None and warmup > 0:   # bounded: as many of the requested passes as fit the time budget
        steps = max(1, min(steps, int(budget_s / max(t_warm, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return batch * frames * HOP / SR / dt, dt, torch.get_num_threads(), kind, steps


def cudnn_eager(z, g, reps=3):
    """The survey's "kernel to beat" (SURVEY.md 2.1 / 8d): the unmodified reference Generator through torch-eager cuDNN
    on this B200, same weights and batch.  ms per step for strict fp32, torch's default fp32 (TF32 convs allowed) and
    bf16 autocast; cudnn.benchmark on so that the library path gets its best algorithms."""
    import torch
    G = reference_generator(z.device)
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for key, tf32, autocast in (("cudnn_eager_fp32_ms_per_step", False, False),
                                    ("cudnn_eager_tf32_ms_per_step", True, False),
                                    ("cudnn_eager_bf16_autocast_ms_per_step", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                for _ in range(2):
                    y = G(z, g=g)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    y = G(z, g=g)
                e1.record()
                torch.cuda.synchronize()
            out[key] = e0.elapsed_time(e1) / reps
            out[key.replace("_ms_per_step", "_out_dtype")] = str(y.dtype).replace("torch.", "")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
    del G
    torch.cuda.empty_cache()
    return out


def infer_cpu(reps=2):
    """BASELINE config 1: the reference's SynthesizerTrn.infer on the host cores (fp32, finetune_speaker.json
    hyper-parameters, random init, batch 1, 50 synthetic symbols, sid 0) and the share of it spent in ``dec``."""
    import torch
    models = import_reference("models")
    cfg = json.load(open(os.path.join(reference_dir(), "configs", "finetune_speaker.json")))
    torch.manual_seed(1234)
    net = models.SynthesizerTrn(68, cfg["data"]["filter_length"] // 2 + 1,
                                cfg["train"]["segment_size"] // cfg["data"]["hop_length"],
                                n_speakers=cfg["data"]["n_speakers"], **cfg["model"]).eval()
    x = torch.randint(1, 68, (1, 50))
    xl = torch.tensor([50])
    sid = torch.tensor([0])
    t_dec = [0.0]
    t_in = [0.0]
    h0 = net.dec.register_forward_pre_hook(lambda m, a: t_in.__setitem__(0, time.perf_counter()))
    h1 = net.dec.register_forward_hook(lambda m, a, o: t_dec.__setitem__(0, t_dec[0] + time.perf_counter() - t_in[0]))
    best = None
    with torch.no_grad():
        net.infer(x, xl, sid=sid, noise_scale=.667, noise_scale_w=0.8, length_scale=1)   # warm-up
        for _ in range(reps):
            t_dec[0] = 0.0
            t0 = time.perf_counter()
            o = net.infer(x, xl, sid=sid, noise_scale=.667, noise_scale_w=0.8, length_scale=1)[0]
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, t_dec[0], o.shape[-1] / SR)
    h0.remove()
    h1.remove()
    return {"infer_cpu_s": best[0], "dec_share": best[1] / best[0], "infer_cpu_audio_s": best[2]}


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
    ap.add_argument("--total-batch", type=int, default=256,
                    help="BASELINE config 4: utterances sharded over all GPUs for the strong-scaling extra key (0 = skip)")
    ap.add_argument("--micro-batch", type=int, default=64, help="utterances per decode call of the config-4 leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (A/B runs)")
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
        # the WHOLE batch of the workload per step (16 x 10 s), every host thread; as many of the K requested passes as
        # fit ~150 s of CPU time (a pass takes ~6 s on 16 cores)
        val, dt, cores, kind, steps = cpu_reference(B, frames, max(1, args.steps), 1, budget_s=150.0)
        line = {"impl": "reference", "metric": "decoded audio-sec/sec, VITS Generator", "value": val,
                "unit": "audio-s/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                 "ref_install": "baseline/_ref (unmodified models.Generator)" if kind == "reference"
                                 else "missing: oracle/generator_torch.py port",
                                 "sample": "the whole batch (B=%d, T=%d) per step, %d timed passes after 1 warm-up (as "
                                           "many of the requested %d as fit 150 s), torch CPU eager fp32"
                                           % (B, frames, steps, args.steps)},
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
        # N > 1: the final waveform gather (north_star: the one collective of the path) is part of every e2e step: an
        # all_gather_into_tensor of the rank's fp32 waveforms.
        # The gather is a copy-engine push over NVLink peer memory (sharding.PeerGather: no kernel beside the decode); where
        # the symmetric-memory rendezvous is not available it falls back to NCCL, in the slot's stream order
        # (decode -> gather -> D2H) with the NEXT batch's decode waiting for it (HostPipeline decode_after): an NCCL kernel
        # spinning for its peer while a decode runs holds a few SMs, and the decoder's persistent one-CTA-per-SM launches
        # then need two rounds each (measured at N = 2: 18.4 ms per step overlapped, 9.6 serialised, 8.5 device-only).
        step_no = [0]
        gdone = [None]
        peer = None
        fulls = None
        if world > 1:
            try:
                peer = vitsdec.PeerGather((B, 1, frames * HOP), torch.float32, dev)
            except Exception as e:
                sys.stderr.write("bench: peer-memory gather unavailable (%r), using NCCL\n" % (e,))
                fulls = [torch.empty((world * B, 1, frames * HOP), dtype=torch.float32, device=dev) for _ in range(2)]

        def gather_hook(y, stream):
            if peer is not None:
                peer.push(y, step_no[0], after=stream)
            else:
                dist.all_gather_into_tensor(fulls[step_no[0] % 2], y)   # the slot's stream waits for NCCL's
                ev = torch.cuda.Event()
                ev.record(stream)
                gdone[0] = ev
            step_no[0] += 1

        hook = gather_hook if world > 1 else None
        for i in range(8):   # each slot's plan reaches its graph (captured at the third use) before the timed region
            pipe.submit(z_host, g_host, outs_host[i % 2], on_device=hook, decode_after=gdone[0])
        pipe.wait_all()
        if peer is not None:
            peer.finish()
        barrier()
        e0.record()
        for i in range(args.steps):
            pipe.submit(z_host, g_host, outs_host[i % 2], on_device=hook, decode_after=gdone[0])
        pipe.join()               # the current stream waits for every slot (decode, gather and D2H of the last steps)
        if peer is not None:
            peer.finish()         # ... and for every rank's pushes: the gathered batches are complete inside the region
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        extras = {}
        gather_ms = None

        def timed(fn, reps, warm):
            for _ in range(warm):
                fn()
            barrier()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)) / reps

        if not args.no_extras:
            # the stage before the decoder: flow(z_p, reverse) on the same batch, timed separately -- it
            # is not part of the metric, which is the Generator decode alone
            Fl = vitsdec.ResidualCouplingBlock(cargs[0], 192, 5, 1, 4, gin_channels=ckw["gin_channels"])
            for name, p in Fl.named_parameters():
                if name.endswith("post.weight"):
                    p.uniform_(-0.07, 0.07)  # the reference zero-initialises post: give the couplings something to do
            Fl = Fl.to(dev).eval()
            Fl.assume_frozen = True
            ymask = torch.ones((B, 1, frames), device=dev)
            extras["flow_reverse_ms_per_step"] = timed(lambda: Fl(z, ymask, g=g, reverse=True), args.steps, 3)
            del Fl

            # BASELINE config 2 (batch 1, 2 s latent): latency of one decode, launch- and prologue-bound.  With the
            # module's DEFAULT settings (per-call parameter-version check on) and with assume_frozen.
            z1, g1 = z[:1, :, :173].contiguous(), g[:1]
            G.assume_frozen = False
            extras["latency_b1_2s_ms"] = timed(lambda: G(z1, g1), 50, 5)
            G.assume_frozen = True
            extras["latency_b1_2s_frozen_ms"] = timed(lambda: G(z1, g1), 50, 5)

            # BASELINE config 5: one 60 s utterance (T = 5168) decoded in 512-frame chunks with recompute halos
            # (vitsdec.decode_chunked; the exact receptive field, 13 frames; interior chunks run as one batch)
            T60 = -(-int(round(60.0 * SR)) // HOP)
            z60 = torch.from_numpy(np.random.RandomState(60 + rank).standard_normal((1, cargs[0], T60)).astype(np.float32)).to(dev)
            ms60 = timed(lambda: vitsdec.decode_chunked(G, z60, g1, chunk_frames=512, hop=HOP), 10, 4)
            extras["chunked_60s_audio_s_per_s"] = world * T60 * HOP / SR / (ms60 / 1e3)
            extras["chunked_60s_ms"] = ms60
            ms60u = timed(lambda: G(z60, g1), 10, 4)
            extras["unchunked_60s_ms"] = ms60u
            extras["chunked_60s_config"] = "T=%d, 512-frame chunks, halo %d frames each side (halo recompute %.1f %%), " \
                                           "1 utterance per GPU" % (T60, G.receptive_halo(), 100.0 * 2 * G.receptive_halo() / 512)
            del z60

            # option "fp16" (fp16 instead of bf16 operands / stored activations; same kernels, same FLOPs): a short timed
            # loop as an extra key.  The headline above is the bf16 mode BASELINE.json names.
            G.set_option("fp16", 1)
            extras["fp16_mode_ms_per_step"] = timed(lambda: G(z, g), 5, 4)
            G.set_option("fp16", 0)
            G(z, g)

            # BASELINE config 4 as stated: 256 x 10 s sharded by utterance, 256/N per GPU (strong scaling), decoded in
            # micro-batches, with the final fp32 waveform gather INSIDE the timed region (started after the last
            # micro-batch; all ranks receive all 256 waveforms).  Device-resident inputs; max over ranks.
            if args.total_batch > 0 and args.total_batch % world == 0:
                n = args.total_batch // world
                mb = min(n, max(1, args.micro_batch))
                rs4 = np.random.RandomState(100 + rank)
                z4 = torch.from_numpy(rs4.standard_normal((n, cargs[0], frames)).astype(np.float32)).to(dev)
                g4 = torch.from_numpy(rs4.standard_normal((n, ckw["gin_channels"], 1)).astype(np.float32)).to(dev)
                y4 = torch.empty((n, 1, frames * HOP), dtype=torch.float32, device=dev)
                full4 = torch.empty((args.total_batch, 1, frames * HOP), dtype=torch.float32, device=dev) if world > 1 else None

                def step4():
                    for lo in range(0, n, mb):
                        y4[lo:lo + mb] = G(z4[lo:lo + mb], g4[lo:lo + mb])
                    if world > 1:
                        dist.all_gather_into_tensor(full4, y4)

                ms4 = timed(step4, max(2, min(args.steps, 5)), 3)
                extras["config4_strong"] = {
                    "workload": "Generator decode, %d x %.0f s latents sharded by utterance, %d per GPU in micro-batches "
                                "of %d, final waveform all-gather inside the timed region" % (args.total_batch, args.seconds, n, mb),
                    "scaling": "strong", "total_batch": args.total_batch, "per_gpu": n, "micro_batch": mb,
                    "ms_per_step": ms4, "value": args.total_batch * frames * HOP / SR / (ms4 / 1e3), "unit": "audio-s/s",
                    "tflops": args.total_batch * (frames * FLOP_PER_FRAME + FLOP_PER_UTT) / (ms4 / 1e3) / 1e12,
                    "gather_bytes_per_rank": int(y4.numel() * 4) if world > 1 else 0}
                del z4, g4, y4, full4
                torch.cuda.empty_cache()

            if world > 1:  # the gather alone, for reference
                full = torch.empty((world * B, 1, frames * HOP), dtype=torch.float32, device=dev)
                gather_ms = timed(lambda: dist.all_gather_into_tensor(full, y), 3, 1)
                del full
            extras["graph_failed_plans"] = G.get_option("graph_failed")

            # the survey's "kernel to beat" and BASELINE config 1, rank 0 of a 1-GPU run only (they are baselines)
            if world == 1 and reference_dir() is not None:
                try:
                    extras.update(cudnn_eager(z, g))
                    ms_ref = extras["cudnn_eager_bf16_autocast_ms_per_step"]
                    extras["speedup_vs_cudnn_eager_bf16_autocast"] = ms_ref / (ms_total / args.steps)
                    extras["speedup_vs_cudnn_eager_fp32"] = extras["cudnn_eager_fp32_ms_per_step"] / (ms_total / args.steps)
                except Exception as e:   # a baseline leg must never take the headline down with it
                    extras["cudnn_eager_error"] = repr(e)[:200]

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
                     "kernel": "conv_tc / conv_pair / conv_pairf / conv_mrfp kernels (tcgen05 implicit-GEMM convs: %d launches/step, "
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
    line.update(extras)
    if gather_ms is not None:
        line["waveform_gather_ms"] = gather_ms
    if world > 1:
        line["e2e"]["includes_waveform_gather"] = True
        line["e2e"]["gather"] = "copy-engine pushes over NVLink peer memory (sharding.PeerGather)" if peer is not None \
            else "NCCL all_gather_into_tensor, serialised with the decodes"
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            val, dt, cores, kind, n = cpu_reference(1, frames, 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                    "sample": "1 of the %d utterances (B=1, T=%d), %d timed passes after 1 warm-up, "
                                              "torch CPU eager fp32, %s" % (B, frames, n,
                                              "the unmodified reference Generator (baseline/_ref)" if kind == "reference"
                                              else "oracle/generator_torch.py port")}
            if not args.no_extras and reference_dir() is not None:
                try:
                    line.update(infer_cpu())
                except Exception as e:
                    line["infer_cpu_error"] = repr(e)[:200]
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
