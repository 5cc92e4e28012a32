import sys, os
sys.path.insert(0, "/root/repo")
import torch
import index_tts_ipex_b200 as pkg
from oracle import bigvgan_oracle as O
dev = torch.device("cuda:0")
h = O.indextts15_config()
for wseed in (11, 0):
    sd = O.make_state_dict(h, wseed, "tame")
    m = pkg.BigVGAN(h, use_cuda_kernel=True); m.load_state_dict(sd, strict=True); m = m.to(dev).eval(); m.remove_weight_norm()
    m.precision = "bf16"
    sdc = {k: v.to(dev) for k, v in O.fold_weight_norm(sd).items()}
    for (B, T0, Tm, seed) in ((2, 64, 40, 3), (2, 235, 281, 2), (2, 64, 281, 3)):
        latent, mel = O.synthetic_inputs(h, B, T0, Tm, seed=seed)
        with torch.no_grad():
            ref = O.bigvgan_forward(latent.to(dev), mel.to(dev), sdc, h).cpu()
        for thr in (-1, 0):
            pkg.capi.lib().bvg_debug_set_tc_min_melems(thr)
            w = m.decode(latent.to(dev), mel_ref=mel.to(dev)).cpu()
            print(f"wseed {wseed} B {B} T0 {T0} Tm {Tm} thr {thr}: SNR {O.snr_db(ref, w):.2f} dB  per-utt {[round(O.snr_db(ref[i], w[i]),1) for i in range(B)]}  refmax {float(ref.abs().max()):.3f}")
        os.environ["BVG_FUSE_TC"] = "0"
    pkg.capi.lib().bvg_debug_set_tc_min_melems(-1)
