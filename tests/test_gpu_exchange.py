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


def test_two_ranks_share_one_bus(port):
    import torch
    from maximilian_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    res = shard.run_exchange_ranks(world, V, B, NBLK, seed=3)
    for r in range(world):
        buses, err, status = res[r]
        assert err is None and status == 0, (r, err, status)
        assert np.array_equal(buses, res[0][0])                 # same bits on every rank
    p = W.voice_params(V, seed=3)
    o = port.Bank(V, osc="saw", filt="biquad"); W.configure_bank(o, "biquad", p)
    for k in range(NBLK):
        _, mo = o.process(B, want_out=False, want_mix=True)
        np.testing.assert_allclose(res[0][0][k], mo, rtol=1e-9, atol=1e-10)


def test_a_silent_peer_times_out_instead_of_hanging():
    """Rank 1 connects and never launches a block: rank 0's exchange kernel gives up after the bounded wait
    (MXB_EXCHANGE_TIMEOUT_MS), mxb_bank_process returns MXB_ERR_STATE and the status mask names the missing rank."""
    import torch
    from maximilian_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    res = shard.run_exchange_ranks(2, V, B, 2, seed=3, silent_rank=1, timeout_ms=300)
    buses, err, status = res[0]
    assert err is not None and "timed out" in err, err
    assert status == 0b10


def _tables():
    import os
    t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tables.npz"))
    return t["sine"], t["transition"], float(t["sine_before"])


def test_patch_world_of_one_is_the_plain_mix():
    """mxb_patch_set_exchange: a voice patch's bus through the exchange-fused reduce kernel; alone in its world it is the local bus."""
    Vp, Bp = 1500, 192
    capi.set_tables(*_tables())
    pat = W.note_pattern(Vp, seed=9); prm = W.polysynth_params(Vp, seed=9)
    a = capi.Patch(W.polysynth_patch("u8"), Vp, max_frames=Bp); b = capi.Patch(W.polysynth_patch("u8"), Vp, max_frames=Bp)
    for k, v in prm.items():
        a.set(k, v); b.set(k, v)
    ex = capi.Exchange(b.ctx, 0, 1, max_doubles=2 * Bp)
    ex.attach(b)
    for blk in range(3):
        tr = W.note_triggers(pat, Bp, blk)
        _, ma = a.process(Bp, {"trigger": tr}, want_out=False, want_mix=True)
        _, mb = b.process(Bp, {"trigger": tr}, want_out=False, want_mix=True)
        assert np.array_equal(ma, mb) and np.abs(ma).max() > 0
    assert ex.status() == 0
    small = capi.Exchange(b.ctx, 0, 1, max_doubles=16)
    small.attach(b)
    with pytest.raises(capi.MxbError):
        b.process(Bp, {"trigger": W.note_triggers(pat, Bp, 3)}, want_out=False, want_mix=True)      # the bus does not fit: refused before any kernel


def test_two_ranks_share_one_patch_bus(port):
    """The polysynth patch sharded over the GPUs: every rank ends each block with the same bus, the rank-ordered sum of the local
    buses, equal to the oracle's whole-patch mix to fp64 reassociation (the lores stage designs per sample: 1e-9)."""
    import torch
    from maximilian_b200 import shard
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = min(torch.cuda.device_count(), 4)
    Vp, Bp, nb = 3000, 256, 4
    res = shard.run_exchange_ranks(world, Vp, Bp, nb, seed=9, what="patch", tables=_tables())
    for r in range(world):
        buses, err, status = res[r]
        assert err is None and status == 0, (r, err, status)
        assert np.array_equal(buses, res[0][0])
    port.set_tables(*_tables(), "port")
    o = port.Patch(W.polysynth_patch("u8"), Vp, kind="port")
    for k, v in W.polysynth_params(Vp, seed=9).items():
        o.set(k, v)
    pat = W.note_pattern(Vp, seed=9)
    for blk in range(nb):
        _, mo = o.process(Bp, {"trigger": W.note_triggers(pat, Bp, blk).astype(np.float64)}, want_out=False, want_mix=True)
        np.testing.assert_allclose(res[0][0][blk], mo, rtol=1e-9, atol=1e-10)
