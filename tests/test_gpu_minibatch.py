"""GPU: the device-side minibatch generator (epde_minibatch_indices_device, csrc/epde_mt.cuh) is bit-exact with the
reference's `random.choices(range(K), cum_weights=1..K, k=B)` stream (src/data_set.py:31,67) and hands Python's
`random` back exactly the state the reference would have left."""
import ctypes as C
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _mb(K):
    from eigenpde_nn_b200.engine import DeviceMinibatch
    return DeviceMinibatch(K, "cuda:0")


def _host_draw(K, B):
    """The host C generator (pinned bit-exact against random.choices by the CPU tests), continuing Python's state."""
    from eigenpde_nn_b200 import _lib
    version, state, gauss = random.getstate()
    mt = (C.c_uint32 * 624)(*state[:624]); pos = C.c_int32(state[624])
    out = np.empty(B, dtype=np.int64)
    _lib.check(_lib.lib().epde_minibatch_indices(mt, C.byref(pos), K, B, out.ctypes.data_as(C.c_void_p)), "host draw")
    random.setstate((version, tuple(mt) + (pos.value,), gauss))
    return out


def test_golden_index_streams_on_device():
    z = np.load(GOLDEN + "/index_streams.npz")
    n = 0
    while f"s{n}_meta" in z:
        seed, K, B, reps = (int(v) for v in z[f"s{n}_meta"])
        random.seed(seed)
        mb = _mb(K).seed_from_python()
        for r in range(reps):
            got = mb.draw(B).cpu().numpy()
            assert np.array_equal(got, z[f"s{n}_idx"][r]), (seed, K, B, r)
        # the state handed back continues the stream exactly where random.choices would have left it
        mb.sync_to_python()
        after_dev = [random.random() for _ in range(3)]
        random.seed(seed)
        for r in range(reps):
            random.choices(range(K), cum_weights=np.arange(1, K + 1), k=B)
        assert after_dev == [random.random() for _ in range(3)]
        n += 1
    assert n >= 4


@pytest.mark.parametrize("K,B,burn", [(5_000_000, 1 << 17, 0), (12_345_677, 100_003, 311), (20_000, 5000, 623), (7, 1, 5),
                                      (1 << 24, 312, 0), (1000, 312, 0), (1000, 311, 2)])
def test_live_random_choices(K, B, burn):
    """Against random.choices itself, from arbitrary positions inside a 624-word block (burn draws of random.random()
    consume two words each) and across block boundaries (312 draws = exactly one block)."""
    random.seed(90210 + K % 1000 + B)
    for _ in range(burn):
        random.random()
    saved = random.getstate()
    mb = _mb(K).seed_from_python()
    got = [mb.draw(B).cpu().numpy() for _ in range(3)]
    mb.sync_to_python()
    nxt = random.random()
    random.setstate(saved)
    cw = np.arange(1, K + 1)
    for r in range(3):
        want = np.asarray(random.choices(range(K), cum_weights=cw, k=B), dtype=np.int64)
        assert np.array_equal(got[r], want), (K, B, burn, r)
    assert random.random() == nxt


def test_full_size_batches_and_rank_slices():
    """B = 2^24 (BASELINE config 4's global batch) against the host generator, and the 8 rank slices of a sharded step:
    every rank advances the whole stream and keeps its contiguous slice; the union is the single-rank draw."""
    from eigenpde_nn_b200.parallel import shard_bounds
    K, B = 5_000_000, 1 << 24
    random.seed(424242)
    saved = random.getstate()
    want = _host_draw(K, B)
    want2 = _host_draw(K, 4096)
    state_after = random.getstate()
    random.setstate(saved)
    mb = _mb(K).seed_from_python()
    got = mb.draw(B)
    got2 = mb.draw(4096)
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    assert np.array_equal(got2.cpu().numpy(), want2)
    mb.sync_to_python()
    assert random.getstate() == state_after
    for rank in (0, 3, 7):
        random.setstate(saved)
        mr = _mb(K).seed_from_python()
        lo, hi = shard_bounds(B, 8, rank)
        part = mr.draw(B, lo=lo, n=hi - lo)
        assert torch.equal(part.cpu(), torch.from_numpy(want[lo:hi]))
        assert np.array_equal(mr.draw(4096).cpu().numpy(), want2)         # same global stream position on every rank


def test_jump_ahead_slices_equal_the_sequential_stream():
    """Sharded draws by jump-ahead (x^J mod the characteristic polynomial applied to the next 20 560 words): each rank
    generates only its slice, consecutive steps chain through the jumped base state, and the stream handed back to Python
    continues exactly like the sequential one (the jumped state is the 624-word window at the same stream position)."""
    from eigenpde_nn_b200.parallel import shard_bounds
    K, B, world, steps = 4_194_304, 300_007, 8, 3
    random.seed(1729)
    for _ in range(17):
        random.random()                                     # start inside a block
    saved = random.getstate()
    want = [_host_draw(K, B) for _ in range(steps)]
    follow = [random.random() for _ in range(4)]
    for rank in (0, 1, 5, 7):
        random.setstate(saved)
        mb = _mb(K).seed_from_python()
        lo, hi = shard_bounds(B, world, rank)
        for s in range(steps):
            got = mb.draw(B, lo=lo, n=hi - lo).cpu().numpy()
            assert np.array_equal(got, want[s][lo:hi]), (rank, s)
        mb.sync_to_python()
        assert [random.random() for _ in range(4)] == follow, rank


def test_jump_polynomial_identity_and_small_jumps():
    """x^0 = 1 (the state itself), and a slice that starts at 0 but is shorter than the batch."""
    K, B = 1000, 5000
    random.seed(5)
    saved = random.getstate()
    want = _host_draw(K, B)
    want2 = _host_draw(K, B)
    random.setstate(saved)
    mb = _mb(K).seed_from_python()
    assert np.array_equal(mb.draw(B, lo=0, n=1234).cpu().numpy(), want[:1234])
    assert np.array_equal(mb.draw(B, lo=4000, n=1000).cpu().numpy(), want2[4000:])


def test_empty_and_tiny_draws():
    random.seed(3)
    saved = random.getstate()
    mb = _mb(10).seed_from_python()
    assert mb.draw(0).numel() == 0
    mb.sync_to_python()
    assert random.getstate() == saved                                      # nothing consumed
    one = mb.draw(1).cpu().numpy()
    random.setstate(saved)
    assert one[0] == random.choices(range(10), cum_weights=np.arange(1, 11), k=1)[0]


def test_draw_ahead_pipeline_matches_sequential_draws():
    from eigenpde_nn_b200.engine import DrawAhead
    K, B = 1 << 22, 1 << 18
    random.seed(77)
    saved = random.getstate()
    want = [_host_draw(K, B) for _ in range(5)]
    random.setstate(saved)
    mb = _mb(K).seed_from_python()
    ahead = DrawAhead(mb, B).start()
    for i in range(5):
        idx = ahead.take()
        got = idx.clone()                       # the "step" reading the indices
        ahead.done_with(idx)
        assert torch.equal(got.cpu(), torch.from_numpy(want[i])), i
    ahead.finish()


def test_draw_ahead_with_two_draws_in_flight():
    """DrawAhead(depth=2): the next step's indices are complete (peek) while the current step runs and one more draw is
    being generated - the same stream as sequential draws, in order."""
    from eigenpde_nn_b200.engine import DrawAhead
    K, B, steps = 1_000_003, 70_001, 7
    random.seed(31337)
    saved = random.getstate()
    want = [_host_draw(K, B) for _ in range(steps + 2)]
    random.setstate(saved)
    mb = _mb(K).seed_from_python()
    ahead = DrawAhead(mb, B, depth=2).start()
    for i in range(steps):
        idx = ahead.take(defer=True)
        nxt, drawn = ahead.peek(1)
        drawn.synchronize()
        assert np.array_equal(nxt.cpu().numpy(), want[i + 1])
        assert np.array_equal(idx.cpu().numpy(), want[i])
        ahead.done_with(idx)
        ahead.enqueue_next()
    ahead.finish()
