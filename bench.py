#!/usr/bin/env python
"""bench.py -- BigVGAN decode audio-seconds/second on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one decode of 32 utterances x 10 s (T0 = 235 latent frames, 281-frame reference mel) per
GPU in bf16 (BASELINE config 3; at N = 8 this is config 4: 256 utterances sharded by utterance,
32 per GPU) followed, for N > 1, by one NCCL all_gather of the waveforms.  Weak scaling.
Prints ONE JSON line on rank 0.  `--impl reference` times the reference's CPU algorithm (the
oracle port, torch fp32 on all host threads) on a bounded sample of the same workload."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T0 = 235                    # "10 s": 240 640 samples = 10.027 s at 24 kHz
TM = 281                    # 3 s reference mel (hop 256 @ 24 kHz)
B_PER_GPU = 32
SR = 24000
UP = 1024
AUDIO_S_PER_UTT = T0 * UP / SR


def workload_config(batch, precision, world):
    """The `config` object of the GPU arm."""
    return {"workload": f"IndexTTS-1.5 BigVGAN decode (random init), {batch} x 10 s utterances per GPU "
                        f"(T0={T0} latent frames, Tm={TM} mel frames), {precision} storage, utterance-sharded, output = "
                        f"int16 PCM (the caller's clamp/int16 epilogue of infer.py:206-212,234 fused into conv_post); "
                        f"N>1 adds one NCCL all_gather of the int16 waveforms per step",
            "global_batch": world * batch, "audio_s_per_step": world * batch * AUDIO_S_PER_UTT,
            "l2_policy": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


FUSED_MAX_C = 128      # AMPBlock1 act->conv pairs with C <= 128 run as one fused kernel (conv_umma_fused.cu)


def measured_traffic(kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu capture of THIS
    build committed under profiles/ (tools/traffic_from_ncu.py writes profiles/traffic.json next to the csv it read);
    (None, note) when no capture of the current library exists -- never a number typed in by hand."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        e = d[kernel_class]
        return float(e["bytes_per_launch"]), f"{e['source']} ({e['launches']} launches, library {d.get('library', '?')})"
    except Exception:
        return None, "no ncu dram__bytes capture of this build under profiles/ (profiles/traffic.json)"


def algorithmic_work(h, B, T0_, es=2, fused=True):
    """Algorithmic work of one step: FLOPs of the dense convs, bytes of the Activation1d passes
    (2*B*C*T*es each) and of the fused Activation1d->conv launches (input + output [+ residuals] + weights)."""
    C0 = h.upsample_initial_channel
    conv_mac = C0 * h.gpt_dim * 7 * T0_            # conv_pre
    fused_mac = 0
    convtr_mac = 0
    act_elems = 0                                   # standalone Activation1d launches only
    fused_bytes = 0.0
    fused_launches = 0
    T = T0_
    nk = len(h.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        cin, cout = C0 >> i, C0 >> (i + 1)
        convtr_mac += cin * cout * k * T
        T *= u
        for j, rk in enumerate(h.resblock_kernel_sizes):
            mac = 6 * cout * cout * rk * T
            if fused and cout <= FUSED_MAX_C:
                fused_mac += mac
                fused_launches += 6
                # 3 x c1 (read x, write xt) + 3 x c2 (read xt, read the residual, write); the last c2 of resblocks
                # 2.. also reads the running sum
                passes = 3 * 2 + 3 * 3 + (1 if j > 0 else 0)
                fused_bytes += passes * cout * T * es * B + 6 * cout * cout * rk * 2
            else:
                conv_mac += mac
                act_elems += 6 * cout * T
    cp = C0 >> len(h.upsample_rates)
    act_elems += cp * T
    post_mac = cp * 7 * T
    return dict(conv_flops=2.0 * conv_mac * B, fused_flops=2.0 * fused_mac * B, convtr_flops=2.0 * convtr_mac * B,
                post_flops=2.0 * post_mac * B, act_elems=float(act_elems) * B, fused_bytes=fused_bytes,
                fused_launches=fused_launches, nk=nk)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
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
                "reasons": sorted(reasons), "samples": len(sm)}


def act1d_sweep(pkg, torch, dev, pk):
    """BASELINE config 5b, bounded: standalone Activation1d through the op boundary (plain [B,C,T] tensors,
    `anti_alias_activation_cuda.forward` seam) on 3 shapes x {fp32, bf16} x {precise (libdevice sinf, the fp32 default),
    fast (MUFU)}, in+out >= 512 MB per call (>> 126 MB L2); plus the decode path's own c8t bf16 kernels (CUDA-core
    stencil and tensor-core FIRs) on the generator's stage shapes.  GB/s = 2*B*C*T*sizeof / CUDA-event time."""
    out = {"op_boundary": [], "c8t_bf16": []}
    L = pkg.capi.lib()

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for Cn, T, Bn in ((768, 4096, 44), (96, 65536, 22), (24, 1048576, 6)):
        a = torch.randn(Cn, device=dev) * 0.5
        b = torch.randn(Cn, device=dev) * 0.5
        for dt, name, es in ((torch.float32, "fp32", 4), (torch.bfloat16, "bf16", 2)):
            x = torch.randn(Bn, Cn, T, device=dev).to(dt)
            for precise in (True, False):
                ms = timed(lambda: pkg.anti_alias_activation_forward(x, None, None, a, b, precise=precise))
                gbs = 2.0 * Bn * Cn * T * es / (ms * 1e-3) / 1e9
                out["op_boundary"].append({"C": Cn, "T": T, "B": Bn, "dtype": name, "precise": precise, "ms": ms,
                                           "GBps": gbs, "frac_of_hbm": gbs / pk["hbm"]})
            del x
    # the "existing GPU kernel to beat" (BASELINE.md section 4): the reference's own fused CUDA op
    # (alias_free_activation/cuda/anti_alias_activation_cuda.cu:43-181) built for sm_100 by baseline/build_ref_kernel.py,
    # on the same shapes and tensors.  Bench-only: nothing in the product imports it.
    out["reference_cuda_kernel"] = ref_kernel_times(torch, dev, pk, timed)
    st = torch.cuda.current_stream().cuda_stream
    for Cn, T, Bn in ((768, 940, 32), (384, 3760, 32), (192, 15040, 32), (96, 60160, 32), (24, 240640, 32)):
        x = (torch.randn(Bn, Cn, T, device=dev) * 1.5).to(torch.bfloat16)
        y = torch.empty_like(x)
        a = torch.randn(Cn, device=dev) * 0.5
        b = torch.randn(Cn, device=dev) * 0.5
        for impl, name in ((1, "cuda_core_stencil"), (2, "tensor_core_fir")):
            for _ in range(2):
                pkg.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), Bn, Cn, T, impl, st))
            torch.cuda.synchronize()
            pkg.capi.profile_begin()                    # (the entry point converts plain <-> c8t around the kernel: time the
            for _ in range(5):                          #  Activation1d kernel class only)
                pkg.capi.check(L.bvg_act1d_c8t_impl_fwd(y.data_ptr(), x.data_ptr(), a.data_ptr(), b.data_ptr(), Bn, Cn, T, impl, st))
            ms = pkg.capi.profile_end()["act1d"][0] / 5
            gbs = 4.0 * Bn * Cn * T / (ms * 1e-3) / 1e9
            out["c8t_bf16"].append({"C": Cn, "T": T, "B": Bn, "impl": name, "ms": ms, "GBps": gbs, "frac_of_hbm": gbs / pk["hbm"]})
        del x, y
    return out


def ref_kernel_times(torch, dev, pk, timed):
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import build_ref_kernel
        mod = build_ref_kernel.load()
    except Exception as e:                                      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    if mod is None:
        return {"unavailable": "baseline/_ref/anti_alias_activation_cuda not built (python baseline/build_ref_kernel.py)"}
    from oracle import bigvgan_oracle as O
    taps = torch.tensor(O.act1d_taps(), dtype=torch.float32, device=dev).view(1, 1, 12)
    rows = []
    for Cn, T, Bn in ((768, 4096, 44), (96, 65536, 22), (24, 1048576, 6)):
        a = torch.randn(Cn, device=dev) * 0.5
        b = torch.randn(Cn, device=dev) * 0.5
        for dt, name, es in ((torch.float32, "fp32", 4), (torch.bfloat16, "bf16", 2)):
            x = torch.randn(Bn, Cn, T, device=dev).to(dt)
            try:
                ms = timed(lambda: mod.forward(x, taps, taps, a, b))
                gbs = 2.0 * Bn * Cn * T * es / (ms * 1e-3) / 1e9
                rows.append({"C": Cn, "T": T, "B": Bn, "dtype": name, "ms": ms, "GBps": gbs, "frac_of_hbm": gbs / pk["hbm"]})
            except Exception as e:                              # noqa: BLE001
                rows.append({"C": Cn, "T": T, "B": Bn, "dtype": name, "error": f"{type(e).__name__}: {e}"})
            del x
    return {"kernel": "reference anti_alias_activation_cuda.forward, -gencode arch=compute_100,code=sm_100", "rows": rows}


def latency_b1(m, O, h, torch, dev):
    """BASELINE config 2: one 10 s utterance (B = 1, T0 = 235), ms per decode for the three precisions (CUDA events)."""
    lat, mel = O.synthetic_inputs(h, 1, T0, TM, seed=3)
    lat, mel = lat.to(dev), mel.to(dev)
    res = {}
    keep = m.precision
    try:
        for prec in ("fp32", "fp32x3", "bf16"):
            m.precision = prec
            for _ in range(2):
                m.decode(lat, mel_ref=mel)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3 if prec == "fp32" else 10
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                m.decode(lat, mel_ref=mel)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            res[prec] = {"ms": ms, "x_realtime": AUDIO_S_PER_UTT * 1e3 / ms}
        # the same bf16 decode replayed from a CUDA graph (BigVGAN.make_graphed_decode: the serving form for one utterance,
        # ~200 launches without their host-side cost)
        m.precision = "bf16"
        run = m.make_graphed_decode(1, T0, TM, device=dev)
        for _ in range(2):
            run(lat, mel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            run(lat, mel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res["bf16_cuda_graph"] = {"ms": ms, "x_realtime": AUDIO_S_PER_UTT * 1e3 / ms}
        # a voice whose embedding is cached (SpeakerEmbeddingCache, the serving case of webui.py:199-221): the decode takes spk
        spk = m.speaker_embed(mel)
        for _ in range(2):
            m.decode(lat, spk=spk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            m.decode(lat, spk=spk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["bf16_cached_voice"] = {"ms": ms, "x_realtime": AUDIO_S_PER_UTT * 1e3 / ms}
    finally:
        m.precision = keep
    return res


def ragged_batch(m, O, h, torch, dev):
    """32 utterances of DISTINCT lengths (3.4 .. 10 s) in one bvg_decode_varlen call, against decoding them
    one by one (what equal-length grouping degenerates to) and against padding all of them to the longest."""
    lens = [80 + 5 * i for i in range(32)]
    lat, mel = O.synthetic_inputs(h, 32, max(lens), TM, seed=7)
    lat, mel = lat.to(dev), mel.to(dev)
    spk = m.speaker_embed(mel)
    audio_s = sum(lens) * UP / SR

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    seqs = [lat[i, :n].contiguous() for i, n in enumerate(lens)]
    ms_var = timed(lambda: m.decode_varlen(lat, spk=spk, lens=lens, pcm16=True))
    ms_one = timed(lambda: [m.decode(s[None], spk=spk[i:i + 1], pcm16=True) for i, s in enumerate(seqs)], n=2)
    ms_pad = timed(lambda: m.decode(lat, spk=spk, pcm16=True))
    return {"lens_frames": [lens[0], lens[-1]], "audio_s": audio_s,
            "one_call_varlen": {"ms": ms_var, "audio_s_per_s": audio_s * 1e3 / ms_var},
            "one_by_one": {"ms": ms_one, "audio_s_per_s": audio_s * 1e3 / ms_one},
            "padded_to_longest_wrong_edges": {"ms": ms_pad, "audio_s_per_s": audio_s * 1e3 / ms_pad}}


def cpu_reference_rate(frames, threads=None, warm_frames=16):
    """audio-s/s of the oracle port (the reference's algorithm in torch fp32 on the host cores)."""
    import torch
    from oracle import bigvgan_oracle as O
    if threads:
        torch.set_num_threads(threads)
    h = O.indextts15_config()
    sd = O.fold_weight_norm(O.make_state_dict(h, 0, "tame"))
    with torch.no_grad():
        lat, mel = O.synthetic_inputs(h, 1, warm_frames, TM, seed=1)
        O.bigvgan_forward(lat, mel, sd, h)
        lat, mel = O.synthetic_inputs(h, 1, frames, TM, seed=1)
        t = time.perf_counter()
        O.bigvgan_forward(lat, mel, sd, h)
        dt = time.perf_counter() - t
    return frames * UP / SR / dt, dt, torch.get_num_threads()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port; the Python
    reference itself cannot travel to the GPU box) on all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import bigvgan_oracle as O
    h = O.indextts15_config()
    sd = O.fold_weight_norm(O.make_state_dict(h, 0, "tame"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    budget = 150.0
    with torch.no_grad():
        lat, mel = O.synthetic_inputs(h, 1, 16, TM, seed=1)
        t = time.perf_counter(); O.bigvgan_forward(lat, mel, sd, h); t16 = time.perf_counter() - t
        t = time.perf_counter(); O.bigvgan_forward(lat, mel, sd, h); t16 = min(t16, time.perf_counter() - t)
        n_steps = args.steps + args.warmup
        frames = int(max(16, min(T0, 16 * budget / max(t16 * n_steps, 1e-6))))
        lat, mel = O.synthetic_inputs(h, 1, frames, TM, seed=1)
        for _ in range(args.warmup):
            O.bigvgan_forward(lat, mel, sd, h)
        t = time.perf_counter()
        for _ in range(args.steps):
            O.bigvgan_forward(lat, mel, sd, h)
        dt = time.perf_counter() - t
    val = args.steps * frames * UP / SR / dt
    sample = f"1 utterance x {frames} latent frames ({frames * UP / SR:.2f} s audio) per step, fp32, torch CPU"
    line = {
        "impl": "reference", "metric": "bigvgan_decode_audio_seconds_per_second", "value": val,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IndexTTS-1.5 BigVGAN decode (random init), CPU oracle port (kind: port), fp32, "
                               f"1 utterance x {frames} latent frames (Tm={TM}) per step on {torch.get_num_threads()} host "
                               f"threads: a bounded sample of the GPU arm's workload ({args.batch} x 10 s utterances per GPU)",
                   "global_batch": 1, "audio_s_per_step": frames * UP / SR},
        "reference_note": "the reference's algorithm (oracle port, torch fp32) on the host CPU; NOT the same batch as the "
                          "GPU arm: audio-s/s is a rate, so the bounded sample stands for the workload",
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import index_tts_ipex_b200 as pkg
    from oracle import bigvgan_oracle as O      # only for synthetic weights/inputs and the CPU baseline leg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    if args.total_utts:                      # strong scaling (BASELINE config 4: a fixed job of N utterances split over the ranks)
        assert args.total_utts % world == 0, "--total-utts must divide by the number of GPUs"
        B = args.total_utts // world
    h = O.indextts15_config()
    sd = O.make_state_dict(h, 0, "tame")
    m = pkg.BigVGAN(h, use_cuda_kernel=True)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    m.remove_weight_norm()
    m.precision = args.precision
    lat_h, mel_h = O.synthetic_inputs(h, B, T0, TM, seed=1 + rank)
    lat_h, mel_h = lat_h.pin_memory(), mel_h.pin_memory()
    lat, mel = lat_h.to(dev), mel_h.to(dev)
    L = T0 * UP
    gathered = torch.empty(world * B, L, dtype=torch.int16, device=dev) if world > 1 else None
    m._ensure_plan(dev)
    ws = torch.empty(m.workspace_bytes(B, T0, TM), dtype=torch.uint8, device=dev)      # reused across steps

    def step():
        # int16 PCM out of the decode (fused caller epilogue); the bytes travel as uint8 (NCCL has no int16 datatype)
        wav = m.decode(lat, mel_ref=mel, pcm16=True, workspace=ws)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.uint8), wav.view(torch.uint8))
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pkg.capi.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = pkg.capi.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    nl = torch.tensor([float(launches)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    value = world * B * AUDIO_S_PER_UTT * args.steps / (total_ms / 1e3)

    # ---- end to end through the public host-buffer API: pinned H2D + decode + D2H every step
    out_h = torch.empty(B, L, dtype=torch.int16, pin_memory=True)          # reused result buffer, like the inputs
    for _ in range(2):
        m.decode_host(lat_h, mel_h, dev, pcm16=True, out=out_h)
    barrier()
    t = time.perf_counter()
    n_e2e = max(2, min(args.steps, 5))
    for _ in range(n_e2e):
        m.decode_host(lat_h, mel_h, dev, pcm16=True, out=out_h)       # synchronises the stream before returning
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t) * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * B * AUDIO_S_PER_UTT * n_e2e / (float(e2e_ms.item()) / 1e3)
    h2d = lat_h.numel() * 4 + mel_h.numel() * 4
    d2h = out_h.numel() * out_h.element_size()

    # ---- per-kernel-class device time inside one more step (CUDA events on the launch stream)
    # (the AMP blocks of a stage normally run on three streams; for this pass they stay on one, so that the events around a
    # launch time that kernel alone -- the per-class sums therefore exceed ms_per_step, where the kernels overlap)
    barrier()
    pkg.capi.lib().bvg_debug_set_multi_stream(0)
    try:
        m.decode(lat, mel_ref=mel, pcm16=True)
        pkg.capi.profile_begin()
        n_prof = 2
        for _ in range(n_prof):
            m.decode(lat, mel_ref=mel, pcm16=True)
        prof = pkg.capi.profile_end()
    finally:
        pkg.capi.lib().bvg_debug_set_multi_stream(1)
    if rank == 0:
        es = 2 if args.precision == "bf16" else 4
        fused_on = args.precision == "bf16" and os.environ.get("BVG_FUSE", "1") != "0"
        work = algorithmic_work(h, B, T0, es, fused_on)
        pk = peaks()
        per_step = {k: (v[0] / n_prof, v[1] // n_prof) for k, v in prof.items()}
        kernel_ms = sum(v[0] for v in per_step.values())
        dom = max(per_step, key=lambda k: per_step[k][0])
        conv_ms, conv_n = per_step["conv1d"]
        act_ms, act_n = per_step["act1d"]
        fus_ms, fus_n = per_step["actconv"]
        conv_tf = work["conv_flops"] / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        act_gbs = 2.0 * work["act_elems"] * es / (act_ms / 1e3) / 1e9 if act_ms > 0 else 0.0
        fus_gbs = work["fused_bytes"] / (fus_ms / 1e3) / 1e9 if fus_ms > 0 else 0.0
        if fus_n and fus_n != work["fused_launches"]:
            print(f"bench: warning: {fus_n} fused launches, accounting expects {work['fused_launches']}", file=sys.stderr)
        if dom == "actconv":
            # algorithmic bytes per launch / average launch time; traffic = ncu dram bytes per launch (same step)
            roof = {"kernel": "actconv_tc_kernel: fused Activation1d -> Conv1d of the narrow stages (C = 96 / 48 / 24), both anti-alias "
                              "FIRs and the conv on tcgen05", "bound": "hbm",
                    "achieved": fus_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": fus_gbs / pk["hbm"],
                    "traffic": measured_traffic("actconv")[0], "algorithmic_bytes_per_launch": work["fused_bytes"] / max(fus_n, 1),
                    "traffic_note": measured_traffic("actconv")[1],
                    "limiter": "instruction issue of the snake / store / epilogue warps (85 % of the issue slots on the three SM "
                               "sub-partitions that own the FIR lanes), not HBM: see profiles/README.md, round 2"}
        elif dom == "act1d":
            roof = {"kernel": "act1d_c8t_kernel", "bound": "hbm", "achieved": act_gbs, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": act_gbs / pk["hbm"], "traffic": None}
        else:
            roof = {"kernel": "conv_umma_kernel (dense generator convs outside the fused stages)", "bound": "tensor",
                    "achieved": conv_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": conv_tf / pk["tf_sustained"], "traffic": None,
                    "traffic_note": "per-launch DRAM bytes of representative launches are in profiles/r01_ncu_full_summary.txt"}
        roof["peak_source"] = pk["src"] + (" sustained" if roof["bound"] == "tensor" else " copy")
        roof["launches_per_step"] = per_step[dom][1]
        roof["avg_launch_ms"] = per_step[dom][0] / max(per_step[dom][1], 1)
        roof["share_of_step_kernel_time"] = per_step[dom][0] / kernel_ms if kernel_ms else None
        extras = None
        if not args.no_extras and world == 1:
            ex_sampler = ClockSampler(local)
            ex_sampler.start()
            extras = {"latency_b1_10s": latency_b1(m, O, h, torch, dev), "act1d_sweep": act1d_sweep(pkg, torch, dev, pk),
                      "ragged_batch": ragged_batch(m, O, h, torch, dev)}
            extras["clocks"] = ex_sampler.stop()
        cpu_val, cpu_dt, cores = cpu_reference_rate(args.cpu_frames)
        line = {
            "metric": "bigvgan_decode_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.total_utts else "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": workload_config(B, args.precision, world),
            "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": n_e2e, "api": "BigVGAN.decode_host -> bvg_decode_host (pinned host buffers)"},
            "gpu_launches": int(nl.item()),
            "clocks": clocks,
            "roofline": roof,
            "kernel_classes_ms_per_step": {k: {"ms": v[0], "launches": v[1]} for k, v in per_step.items()},
            "kernel_classes_note": "CUDA events around every launch of one extra step with the three AMP blocks of a stage "
                                   "serialised on one stream (each kernel timed alone); in the timed steps the blocks run on "
                                   "three streams and overlap, so ms_per_step is below the sum of the classes",
            "roofline_actconv": {"bound": "hbm", "achieved": fus_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": fus_gbs / pk["hbm"], "bytes_per_step": work["fused_bytes"],
                                 "flops_per_step": work["fused_flops"], "launches_per_step": fus_n},
            "roofline_act1d": {"bound": "hbm", "achieved": act_gbs, "peak": pk["hbm"], "unit": "GB/s",
                               "frac": act_gbs / pk["hbm"], "bytes_per_step": 2.0 * work["act_elems"] * es,
                               "launches_per_step": act_n},
            "roofline_conv": {"bound": "tensor", "achieved": conv_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                              "frac": conv_tf / pk["tf_sustained"], "flops_per_step": work["conv_flops"],
                              "launches_per_step": conv_n},
            "cpu_baseline": {"value": cpu_val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": f"1 utterance x {args.cpu_frames} latent frames "
                                       f"({args.cpu_frames * UP / SR:.2f} s audio), fp32 oracle port, {cpu_dt:.1f} s"},
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="utterances per GPU (default: the benchmark's 32)")
    ap.add_argument("--total-utts", type=int, default=0, help="strong scaling: total utterances per step, split evenly over the GPUs")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32x3"])
    ap.add_argument("--cpu-frames", type=int, default=200, help="latent frames of the CPU-baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the B = 1 latencies and the Activation1d sweep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
