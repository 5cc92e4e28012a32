"""2+ GPU functional check of decode_sharded over NCCL: the gathered waveforms must equal a single-GPU decode of the
whole batch, including a ragged split (B not divisible by the world size) and the broadcast reference mel.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import index_tts_ipex_b200 as P  # noqa: E402
from oracle import bigvgan_oracle as O  # noqa: E402  (synthetic weights / inputs only)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    h = O.indextts15_config()
    m = P.BigVGAN(h, use_cuda_kernel=True)
    m.load_state_dict(O.make_state_dict(h, 0, "tame"), strict=True)
    m = m.to(dev).eval()
    m.remove_weight_norm()
    ok = True
    for precision in ("fp32", "bf16"):
        m.precision = precision
        for B, Bm in ((2 * world + 1, 2 * world + 1), (world + 1, 1)):
            lat, mel = O.synthetic_inputs(h, B, 12, 60, seed=7, Bm=Bm)
            lat, mel = lat.to(dev), mel.to(dev)
            full = m.decode(lat, mel_ref=mel)
            got = P.decode_sharded(lambda x, c: m.decode(x, mel_ref=c), lat, mel)
            same = bool(torch.equal(full, got))
            ok &= same
            if rank == 0:
                print(f"{precision} B={B} Bm={Bm} world={world}: gathered == single-GPU decode: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
