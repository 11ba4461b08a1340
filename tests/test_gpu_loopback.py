"""-m gpu: the fused loss all-reduce (mbx_match_loss_allreduce) with several ranks emulated on ONE GPU.

The multi-GPU tests (tests/test_gpu_dist.py, profiles/dist_check.py) need a box with >= 2 GPUs.  The
exchange protocol itself -- outboxes, the relay kernel that forwards them into every rank's table, the pull
route when it has not, the lag under programmatic dependent launch, the blocking mode, the rings, the sticky
timeout flag --
only sees a table of buffer pointers, so `multibox_b200.dist.LoopbackPeers` runs it unchanged with every
"rank" on the same device.  Semantics checked: SUM of reference loss.py:100-101 over the ranks, added in
rank order (bit-exact), for the step the result block names (results[14])."""
import numpy as np
import pytest
import torch

from multibox_b200 import _lib, loss, synth
from multibox_b200 import dist as mdist
from gpu_util import dev

pytestmark = pytest.mark.gpu
NSETS = 3


def _shards(world, B=12, K=5, M=20):
    """NSETS different global batches of world*B images, sharded by rank: sets[s][r] = device tensors."""
    sets = []
    for s in range(NSETS):
        d = synth.make_train_inputs(K=K, B=world * B, M=M, seed=100 + s, edge_cases=(s == 0))
        per_rank = []
        for r in range(world):
            sl = slice(r * B, (r + 1) * B)
            per_rank.append((dev(d["locations"][sl]), dev(d["confidences"][sl]).view(B, -1), dev(d["gt"][sl]),
                             dev(d["num_gt"][sl])))
        sets.append(per_rank)
    return d, sets


def _local_sums(d, sets, world, B):
    """fp64 local sums of every (set, rank), from a plain single-rank step (no all-reduce)."""
    solo = loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"])
    out = []
    for s in range(NSETS):
        row = []
        for r in range(world):
            solo.step(*sets[s][r])
            torch.cuda.synchronize()
            row.append(solo.out["results"].cpu()[4:8].view(torch.float64).numpy().copy())
        out.append(row)
    return out


def _expect(local, s, world):
    g = np.zeros(2, dtype=np.float64)
    for r in range(world):          # rank order, like the kernel
        g = g + local[s][r]
    return g


def _read(step):
    res = step.out["results"].cpu()
    return (res[4:8].view(torch.float64).numpy().copy(), res[8:12].view(torch.float64).numpy().copy(),
            int(res[14].item()), int(res[2].item()))


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("mode", ["deferred", "deferred_pdl", "deferred_generic"])
def test_deferred_allreduce_loopback(cuda_device, world, mode):
    B = 12
    d, sets = _shards(world, B)
    local = _local_sums(d, sets, world, B)
    peers = mdist.LoopbackPeers(world)
    lag = 12 if mode == "deferred_pdl" else 1
    ranks = []
    for r in range(world):
        st = loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"], peer=peers.rank(r),
                                   deferred_allreduce=True, pdl=(mode == "deferred_pdl"))
        if mode == "deferred_generic":
            st.flags |= _lib.FLAG_GENERIC     # the shared-memory kernel family, same tail
        ranks.append(st)
    # prepare() runs one step per (rank, set): steps 0 .. NSETS-1, all ranks in step order
    launches = [[None] * NSETS for _ in range(world)]
    for s in range(NSETS):
        for r in range(world):
            launches[r][s] = ranks[r].prepare(*sets[s][r])
    torch.cuda.synchronize()
    done = NSETS                      # steps every rank has run so far; step k used set k % NSETS
    # (a) synchronised steps: every result block names its step and carries that step's global sums
    for it in range(lag + 4):
        s = done % NSETS
        for r in range(world):
            launches[r][s]()
        torch.cuda.synchronize()
        for r in range(world):
            loc, glob, gstep, status = _read(ranks[r])
            assert status == 0
            assert np.array_equal(loc, local[s][r])
            assert gstep == max(done - lag, -1)
            want = _expect(local, gstep % NSETS, world) if gstep >= 0 else np.zeros(2)   # (-1: nothing to complete yet)
            assert np.array_equal(glob, want), (mode, world, r, it)
        done += 1
    # (b) 40 back-to-back steps without any host synchronisation (overlapping launches under PDL)
    for it in range(40):
        s = done % NSETS
        for r in range(world):
            launches[r][s]()
        done += 1
    torch.cuda.synchronize()
    for r in range(world):
        loc, glob, gstep, status = _read(ranks[r])
        assert status == 0
        assert gstep == done - 1 - lag
        assert np.array_equal(glob, _expect(local, gstep % NSETS, world))
    # (c) flush completes the newest step
    for r in range(world):
        g = ranks[r].flush()
        assert ranks[r].global_step() == done - 1
        assert np.array_equal(np.array(g), _expect(local, (done - 1) % NSETS, world))


def test_blocking_allreduce_loopback_two_streams(cuda_device):
    """Blocking mode: a rank's kernel waits for every rank's post of the SAME step, so the two emulated
    ranks run on two streams (concurrent kernels).  Then a mode switch on the same buffers: deferred steps
    pull what blocking steps left in the outboxes, and the other way round."""
    world, B = 2, 12
    d, sets = _shards(world, B)
    local = _local_sums(d, sets, world, B)
    peers = mdist.LoopbackPeers(world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    ranks, launches = [], [[None] * NSETS for _ in range(world)]
    for r in range(world):
        ranks.append(loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"], peer=peers.rank(r)))
        with torch.cuda.stream(streams[r]):     # outputs + this stream's workspace exist before any kernel spins
            loss.match_loss_raw(*sets[0][r], ranks[r].priors, d["alpha"], out=ranks[r].out)
    torch.cuda.synchronize()
    for s in range(NSETS):
        for r in range(world):      # (prepare launches one step: both ranks' kernels must be in flight together)
            with torch.cuda.stream(streams[r]):
                launches[r][s] = ranks[r].prepare(*sets[s][r])
    torch.cuda.synchronize()
    done = NSETS
    for it in range(6):
        s = done % NSETS
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                launches[r][s]()
        torch.cuda.synchronize()
        for r in range(world):
            loc, glob, gstep, status = _read(ranks[r])
            assert status == 0 and gstep == done
            assert np.array_equal(glob, _expect(local, s, world))
        done += 1
    # deferred steps after blocking ones (one stream is enough now)
    defer = [loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"], peer=peers.rank(r),
                                   deferred_allreduce=True) for r in range(world)]
    for it in range(3):
        s = done % NSETS
        for r in range(world):
            defer[r].step(*sets[s][r])
        torch.cuda.synchronize()
        for r in range(world):
            loc, glob, gstep, status = _read(defer[r])
            assert status == 0 and gstep == done - 1
            assert np.array_equal(glob, _expect(local, (done - 1) % NSETS, world))
        done += 1
    # and blocking again
    s = done % NSETS
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            launches[r][s]()
    torch.cuda.synchronize()
    for r in range(world):
        loc, glob, gstep, status = _read(ranks[r])
        assert status == 0 and gstep == done
        assert np.array_equal(glob, _expect(local, s, world))


def test_allreduce_timeout_is_sticky_until_reset(cuda_device):
    """A rank whose peer never launches: MBX_STATUS_AR_TIMEOUT after ~2 s instead of a hang, the local
    losses are still produced, later steps fail AT ONCE (sticky flag), and a reset recovers."""
    import time
    world, B = 2, 12
    d, sets = _shards(world, B)
    local = _local_sums(d, sets, world, B)
    peers = mdist.LoopbackPeers(world)
    r0 = loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"], peer=peers.rank(0))   # blocking
    r0.step(*sets[0][0])
    torch.cuda.synchronize()
    loc, glob, gstep, status = _read(r0)
    assert status & _lib.STATUS_AR_TIMEOUT
    assert np.array_equal(loc, local[0][0])
    t0 = time.perf_counter()
    r0.step(*sets[1][0])
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 1.0
    assert _read(r0)[3] & _lib.STATUS_AR_TIMEOUT
    with pytest.raises(RuntimeError):
        loss.raise_for_status(_read(r0)[3])
    peers.reset()
    both = [loss.MultiboxLossStep(B, d["P"], d["M"], d["priors"], d["alpha"], peer=peers.rank(r),
                                  deferred_allreduce=True) for r in range(world)]
    for it in range(2):
        for r in range(world):
            both[r].step(*sets[it][r])
    torch.cuda.synchronize()
    for r in range(world):
        loc, glob, gstep, status = _read(both[r])
        assert status == 0 and gstep == 0
        assert np.array_equal(glob, _expect(local, 0, world))
