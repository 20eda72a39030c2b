"""Host logic of the multi-GPU path on CPU: world_size-2 gloo process group, a deterministic
fake infer_fn; checks contiguous sharding, ordered gather, tail padding and empty shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yourmt3_b200.audio_utils import slice_padded_array
from yourmt3_b200.sharding import shard_range, transcribe_sharded


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1758):
        for w in (1, 2, 4, 8):
            got = []
            for r in range(w):
                a, b, per = shard_range(n, w, r)
                assert 0 <= a <= b <= n and b - a <= per
                got += list(range(a, b))
            assert got == list(range(n))


def test_slice_padded_array_semantics():
    x = np.arange(70000, dtype=np.float32)
    s = slice_padded_array(x, 32767, 32767)
    assert s.shape == (3, 1, 32767)
    assert (s[0, 0] == x[:32767]).all() and (s[1, 0] == x[32767:65534]).all()
    assert (s[2, 0, : 70000 - 65534] == x[65534:]).all() and (s[2, 0, 70000 - 65534:] == 0).all()
    assert slice_padded_array(torch.zeros(1, 100), 32767, 32767).shape == (1, 1, 32767)
    assert slice_padded_array(np.zeros(32767 * 2, np.float32)).shape == (2, 1, 32767)
    assert slice_padded_array(np.zeros(100, np.float32), pad=False).shape == (0, 1, 32767)
    # 1 hour of 16 kHz audio -> 1758 segments (BASELINE.json configs[4])
    assert slice_padded_array(np.zeros(57_600_000, np.float32)).shape[0] == 1758


def _fake_infer(x):
    # tokens derived from the segment content so that ordering errors are visible: (b, 3, 5)
    base = x[:, 0, 0].round().to(torch.int64)
    return base[:, None, None] * 100 + torch.arange(15).view(1, 3, 5)


def _worker(rank, world, port, n_seg, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    audio = torch.zeros(n_seg, 1, 64)
    audio[:, 0, 0] = torch.arange(n_seg, dtype=torch.float32)
    out = transcribe_sharded(_fake_infer, audio, bsz=2, device=torch.device("cpu"), pad_id=0, token_shape=(3, 5))
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_seg", [7, 1, 4])
def test_gloo_world2_gather(n_seg):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_seg, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (np.arange(n_seg)[:, None, None] * 100 + np.arange(15).reshape(1, 3, 5)).astype(np.int32)
    for r in range(2):
        assert res[r].dtype == np.int32 and res[r].shape == (n_seg, 3, 5)
        assert (res[r] == expect).all()


def test_single_process_path():
    audio = torch.zeros(5, 1, 64)
    audio[:, 0, 0] = torch.arange(5, dtype=torch.float32)
    out = transcribe_sharded(_fake_infer, audio, bsz=2, device=torch.device("cpu"))
    assert out.shape == (5, 3, 5) and int(out[4, 0, 0]) == 400


def test_empty_shard_without_token_shape_is_an_error():
    with pytest.raises(ValueError, match="token_shape"):
        transcribe_sharded(_fake_infer, torch.zeros(0, 1, 64), bsz=2, device=torch.device("cpu"))
    out = transcribe_sharded(_fake_infer, torch.zeros(0, 1, 64), bsz=2, device=torch.device("cpu"), token_shape=(3, 5))
    assert out.shape == (0, 3, 5)
