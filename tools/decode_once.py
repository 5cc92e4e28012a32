"""Two bf16 decodes of the benchmark shape (B utterances of 10 s); for `ncu --metrics gpu__time_duration.sum` launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import index_tts_ipex_b200 as P
from oracle import bigvgan_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
h = O.indextts15_config()
sd = O.make_state_dict(h, 0, "tame")
m = P.BigVGAN(h, use_cuda_kernel=True)
m.load_state_dict(sd, strict=True)
m = m.to("cuda").eval()
m.remove_weight_norm()
latent, mel = O.synthetic_inputs(h, B, 235, 281, seed=1)
m.precision = "bf16"
for _ in range(2):
    torch.cuda.nvtx.range_push("decode")
    m.decode(latent.cuda(), mel_ref=mel.cuda())
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok")
