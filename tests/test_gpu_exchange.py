"""K6: the peer-memory mix-bus exchange (mxb_exchange). With one GPU only the world-of-one path can run; with two or
more (gpurun --gpus 2) two processes, one per GPU, shard a bank and must both end every block with the same bus
-- the rank-ordered sum of the two local buses -- bit for bit, equal to the oracle's full-bank mix within fp64
reassociation, with no NCCL involved (the IPC handles travel through a multiprocessing queue)."""
import numpy as np
import pytest

from maximilian_b200 import capi
from maximilian_b200 import workloads as W

pytestmark = pytest.mark.gpu

V, B, NBLK = 6000, 256, 5


def test_world_of_one_is_the_plain_mix():
    p = W.voice_params(V, seed=3)
    a = capi.Bank(V, osc="saw", filt="biquad", max_frames=B); W.configure_bank(a, "biquad", p)
    b = capi.Bank(V, osc="saw", filt="biquad", max_frames=B); W.configure_bank(b, "biquad", p)
    ex = capi.Exchange(b.ctx, 0, 1, max_doubles=2 * B)
    ex.attach(b)
    for _ in range(3):
        _, ma = a.process(B, want_out=False, want_mix=True)
        _, mb = b.process(B, want_out=False, want_mix=True)
        assert np.array_equal(ma, mb)


def _rank(rank, world, q_in, q_out, results):
    import torch
    torch.cuda.set_device(rank)
    from maximilian_b200 import shard
    ctx = capi.Context(rank, 48000)
    p = W.voice_params(V, seed=3)
    lo, hi = shard.shard_range(V, rank, world)
    bank = capi.Bank(hi - lo, osc="saw", filt="biquad", max_frames=B, ctx=ctx)
    W.configure_bank(bank, "biquad", {k: v[lo:hi] for k, v in p.items()})
    ex = capi.Exchange(ctx, rank, world, max_doubles=2 * B)
    q_out.put((rank, ex.local_handle()))
    handles = q_in.get(timeout=120)
    ex.connect(handles)
    ex.attach(bank)
    mixes = []
    for _ in range(NBLK):
        _, m = bank.process(B, want_out=False, want_mix=True)
        mixes.append(m.copy())
    results.put((rank, np.stack(mixes)))


def test_two_ranks_share_one_bus(port):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q_ins = [ctx.Queue() for _ in range(world)]
    q_out, results = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, world, q_ins[r], q_out, results)) for r in range(world)]
    [p.start() for p in procs]
    hs = dict(q_out.get(timeout=120) for _ in range(world))
    for r in range(world):
        q_ins[r].put([hs[k] for k in range(world)])
    res = dict(results.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert np.array_equal(res[0], res[1])                       # same bits on every rank
    p = W.voice_params(V, seed=3)
    o = port.Bank(V, osc="saw", filt="biquad"); W.configure_bank(o, "biquad", p)
    for k in range(NBLK):
        _, mo = o.process(B, want_out=False, want_mix=True)
        np.testing.assert_allclose(res[0][k], mo, rtol=1e-9, atol=1e-10)
