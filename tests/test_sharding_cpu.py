"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): utterance partitioning and the final gather.
The decode itself needs a GPU, so a deterministic stand-in `decode_fn` is injected."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_decode(latent, mel):
    # wav[b,0,l] depends on the utterance's own latent and its (possibly broadcast) mel only
    B, T0, D = latent.shape
    L = T0 * 4
    base = latent.sum(dim=(1, 2)).view(B, 1, 1) + mel.sum(dim=(1, 2)).view(-1, 1, 1)
    return base + torch.arange(L, dtype=latent.dtype).view(1, 1, L)


def _fake_decode_pcm(latent, mel):
    # the int16 PCM layout of the fused epilogue: [b, L]
    return _fake_decode(latent, mel).squeeze(1).clamp(-3e4, 3e4).to(torch.int16)


def _worker(rank, world, port, B, Bm, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import index_tts_ipex_b200 as P
    g = torch.Generator().manual_seed(5)
    latent = torch.randn(B, 3, 4, generator=g)
    mel = torch.randn(Bm, 5, 2, generator=g)
    calls = []

    def fn(l, m):
        calls.append(l.shape[0])
        return _fake_decode(l, m)

    out = P.decode_sharded(fn, latent, mel)
    ref = _fake_decode(latent, mel)
    ok = out.shape == ref.shape and torch.equal(out, ref)
    lo, hi = P.shard_bounds(B, world)[rank]
    ok = ok and (calls == ([hi - lo] if hi > lo else []))
    # int16 PCM shards travel as int16 (half the gather bytes), ranks with an empty shard included
    out16 = P.decode_sharded(_fake_decode_pcm, latent, mel)
    ref16 = _fake_decode_pcm(latent, mel)
    ok = ok and out16.dtype == torch.int16 and out16.shape == ref16.shape and torch.equal(out16, ref16)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B,Bm", [(2, 8, 8), (2, 5, 1), (3, 2, 2), (2, 1, 1)])
def test_decode_sharded_gloo(world, B, Bm):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, Bm, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_shard_bounds_properties():
    import index_tts_ipex_b200 as P
    for n in (0, 1, 7, 32, 255, 256):
        for w in (1, 2, 3, 4, 8):
            b = P.shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1
    assert P.shard_bounds(256, 8) == [(32 * r, 32 * (r + 1)) for r in range(8)]     # BASELINE config 4
    with pytest.raises(ValueError):
        P.shard_bounds(4, 0)
