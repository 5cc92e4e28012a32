"""Speaker encoder alone (bvg_speaker_embed, fp32 path) and the kernel-class breakdown of one bf16 decode: time per call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P
from oracle import bigvgan_oracle as O

h = O.indextts15_config()
sd = O.make_state_dict(h, 0, "tame")
m = P.BigVGAN(h, use_cuda_kernel=True)
m.load_state_dict(sd, strict=True)
m = m.to("cuda").eval()
m.remove_weight_norm()
for Bm in (32, 1):
    latent, mel = O.synthetic_inputs(h, Bm, 235, 281, seed=1)
    mel = mel.cuda()
    for _ in range(3):
        m.speaker_embed(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    P.capi.launch_count_reset()
    e0.record()
    for _ in range(20):
        m.speaker_embed(mel)
    e1.record()
    torch.cuda.synchronize()
    print(f"speaker_embed fp32 path Bm={Bm} Tm=281: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per call, {P.capi.launch_count() // 20} launches")
    m.precision = "bf16"
    lat = latent.cuda()
    for _ in range(3):
        m.decode(lat, mel_ref=mel)
    torch.cuda.synchronize()
    P.capi.launch_count_reset()
    P.capi.profile_begin()
    for _ in range(5):
        m.decode(lat, mel_ref=mel)
    torch.cuda.synchronize()
    pr = P.capi.profile_end()
    print(f"decode bf16 B={Bm}: " + "  ".join(f"{k}={v[0] / 5 * 1e3:.0f}us/{v[1] // 5}" for k, v in pr.items()) + f"  launches={P.capi.launch_count() // 5}")
    m.precision = None
