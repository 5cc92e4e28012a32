import sys, numpy as np, torch, os
sys.path.insert(0, ".")
import index_tts_ipex_b200 as P
from oracle import bigvgan_oracle as O
for name in ["full15_tame_T12", "full15_wild_T9", "small_wild_T17_bcast"]:
    g = np.load(os.path.join("tests/golden", name + ".npz"))
    h = O.small_config() if str(g["config"]) == "small" else O.indextts15_config()
    m = P.BigVGAN(h, use_cuda_kernel=True); m.load_state_dict(O.make_state_dict(h, int(g["wseed"]), str(g["mode"])), strict=True)
    m = m.to("cuda").eval(); m.remove_weight_norm(); m.precision = "fp32x3"
    lat, mel = O.synthetic_inputs(h, int(g["B"]), int(g["T0"]), int(g["Tm"]), seed=int(g["iseed"]), Bm=int(g["Bm"]))
    y = m.decode(lat.cuda(), mel_ref=mel.cuda())
    print(name, "max-abs", float(np.abs(y.cpu().numpy() - g["wav"]).max()), "abs-max ref", float(np.abs(g["wav"]).max()))
h = O.indextts15_config(); m = P.BigVGAN(h, use_cuda_kernel=True); m.load_state_dict(O.make_state_dict(h, 0, "tame"), strict=True); m = m.to("cuda").eval(); m.remove_weight_norm()
lat, mel = O.synthetic_inputs(h, 1, 235, 281, seed=1)
m.precision = "fp32"; ref = m.decode(lat.cuda(), mel_ref=mel.cuda()); m.precision = "fp32x3"; y = m.decode(lat.cuda(), mel_ref=mel.cuda())
print("10 s vs fp32 path: max-abs", float((y - ref).abs().max()))
