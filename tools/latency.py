"""BASELINE config 2: B=1, 10 s utterance latency (fp32 parity path and bf16 path), device-resident inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P
from oracle import bigvgan_oracle as O

h = O.indextts15_config()
m = P.BigVGAN(h, use_cuda_kernel=True)
m.load_state_dict(O.make_state_dict(h, 0, "tame"), strict=True)
m = m.to("cuda").eval(); m.remove_weight_norm()
lat, mel = O.synthetic_inputs(h, 1, 235, 281, seed=1)
lat, mel = lat.cuda(), mel.cuda()
for prec in ("fp32", "fp32x3", "bf16"):
    m.precision = prec
    for _ in range(3):
        m.decode(lat, mel_ref=mel)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.decode(lat, mel_ref=mel); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    P.capi.profile_begin(); m.decode(lat, mel_ref=mel); prof = P.capi.profile_end()
    print(f"{prec}: B=1 x 10.03 s  median {ts[len(ts)//2]:.2f} ms  min {ts[0]:.2f} ms  -> {10.027/ (ts[len(ts)//2]/1e3):.0f} x real time; classes {prof}")

for prec in ("fp32", "fp32x3", "bf16"):
    m.precision = prec
    run = m.make_graphed_decode(1, 235, 281)
    for _ in range(3):
        run(lat, mel)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(lat, mel); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{prec} CUDA graph: median {ts[len(ts)//2]:.2f} ms  min {ts[0]:.2f} ms")

# steady-state serving: the speaker embedding of a voice is computed once and reused (decode(..., spk=...))
m.precision = "fp32"
spk = m.speaker_embed(mel)
for prec in ("fp32x3", "bf16"):
    m.precision = prec
    for _ in range(3):
        m.decode(lat, spk=spk)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.decode(lat, spk=spk); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{prec} with a cached speaker embedding: median {ts[len(ts)//2]:.2f} ms  min {ts[0]:.2f} ms")
