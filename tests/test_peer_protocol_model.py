"""CPU: the PROTOCOL of the peer-memory gradient exchange (nerf-tf2_b200/csrc/peer.cu) as an executable model.

Every CTA of every rank is a coroutine that performs the kernel's memory events one at a time -- flag stores, flag
polls, the loads and stores of its part of the rank's slice, the done-counter, the epoch update -- and a scheduler
interleaves all CTAs of all ranks at random (sequentially consistent memory: the model checks the protocol, the
fences are the kernel's business). A rank launches its next exchange, with a NEW local gradient written into its
buffer, as soon as its own kernel has finished, whatever the other ranks are doing -- which is how a fast rank runs
ahead in the real step. Checked over many random schedules, several consecutive exchanges, with and without the fused
Adam epilogue, unicast and multicast data paths, grids wider than the slice:

  * no schedule deadlocks;
  * after exchange e every rank's buffer holds the sum over the ranks of the e-th gradients (in rank order for the
    unicast path), identical on all ranks, and the Adam epilogue only ever reads fully summed values;
  * the in-place update is safe: an element is written only after it was read from every rank, by its owner alone;
  * epochs: flags of exchange e+1 arriving while a rank is still in exchange e do no harm.
"""
import numpy as np
import pytest

class World:
    def __init__(self, W, n, rng):
        self.W, self.n, self.rng = W, n, rng
        self.buf = [np.zeros(n) for _ in range(W)]
        self.hdr = [dict(arriveA=[0] * W, arriveB=[0] * W, epoch=0, done=0) for _ in range(W)]
        self.reads_of = {}       # (exchange, element) -> set of ranks whose copy the owner has read
        self.adam_seen = [dict() for _ in range(W)]      # rank -> {(exchange, element): value read by the Adam epilogue}


def cta_program(w, rank, b, grid, threads, adam, multicast, exchange, barrier_a=True):
    """Mirrors peer_allreduce_kernel: one `yield` per memory event; `yield pred` blocks until pred() holds."""
    W, n, me = w.W, w.n, w.hdr[rank]
    epoch = me["epoch"] + 1
    yield
    if b == 0:                                                   # barrier A: "my gradient is complete"
        for p in range(W):
            w.hdr[p]["arriveA"][rank] = epoch
            yield
    if barrier_a:
        yield (lambda: all(me["arriveA"][p] >= epoch for p in range(W)))
    per = -(-n // W)
    lo, hi = per * rank, min(n, per * rank + per)
    for i in range(lo + b * threads, hi, grid * threads):        # this CTA's elements (a "thread" per element, in turn)
        for k in range(i, min(i + threads, hi)):
            if multicast:                                        # the switch reads every copy, sums, replicates
                s = 0.0
                for p in w.rng.permutation(W):                   # ... in an order of its own
                    s += w.buf[p][k]
                    w.reads_of.setdefault((exchange, k), set()).add(int(p))
                yield
                for p in range(W):
                    assert w.reads_of[(exchange, k)] == set(range(W)), "written before it was read everywhere"
                    w.buf[p][k] = s
                yield
            else:
                vals = []
                for p in range(W):
                    vals.append(w.buf[p][k])
                    w.reads_of.setdefault((exchange, k), set()).add(p)
                    yield
                s = vals[0]
                for v in vals[1:]:
                    s = s + v                                    # rank order
                for p in range(W):
                    assert w.reads_of[(exchange, k)] == set(range(W)), "written before it was read everywhere"
                    w.buf[p][k] = s
                    yield
    working = max(1, min(grid, -(-(hi - lo) // threads)))
    last = False
    if b < working:                                              # only the CTAs that hold a part of the slice are counted
        c = me["done"]
        me["done"] = c + 1                                       # (one atomic event)
        last = c == working - 1
        yield
    if last:                                                     # barrier B: "my slice has landed everywhere"
        for p in range(W):
            w.hdr[p]["arriveB"][rank] = epoch
            yield
    if not adam:
        if not last:
            return
        yield (lambda: all(me["arriveB"][p] >= epoch for p in range(W)))
        me["done"], me["epoch"] = 0, epoch
        return
    yield (lambda: all(me["arriveB"][p] >= epoch for p in range(W)))
    for i in range(b * threads, n, grid * threads):              # Adam epilogue: this CTA's share of ALL elements
        for k in range(i, min(i + threads, n)):
            w.adam_seen[rank][(exchange, k)] = w.buf[rank][k]
        yield
    if last:
        me["done"], me["epoch"] = 0, epoch


def run(W, n, grid, threads, exchanges, adam, multicast, seed, barrier_a=True):
    rng = np.random.default_rng(seed)
    w = World(W, n, rng)
    grads = rng.integers(-8, 9, size=(exchanges, W, n)).astype(np.float64)      # exact in any summation order
    active = {}          # rank -> list of live coroutines
    blocked = {}         # coroutine -> predicate
    nxt = [0] * W        # next exchange of each rank
    done_ex = [0] * W

    def launch(rank):
        e = nxt[rank]
        w.buf[rank][:] = grads[e, rank]                  # the rank's backward pass of step e (stream order: kernel e-1 is over)
        active[rank] = [cta_program(w, rank, b, grid, threads, adam, multicast, e, barrier_a) for b in range(grid)]
        nxt[rank] += 1

    for r in range(W):
        launch(r)
    steps = 0
    while any(active.values()):
        runnable = [(r, g) for r, gs in active.items() for g in gs if g not in blocked or blocked[g]()]
        assert runnable, f"deadlock: W={W} grid={grid} adam={adam} multicast={multicast} seed={seed}"
        r, g = runnable[rng.integers(len(runnable))]
        blocked.pop(g, None)
        try:
            out = next(g)
            if callable(out):
                blocked[g] = out
        except StopIteration:
            active[r].remove(g)
            if not active[r]:                            # the kernel of rank r is over: its buffer must hold the full sum
                e = nxt[r] - 1
                want = grads[e].sum(axis=0)
                assert np.array_equal(w.buf[r], want), (r, e)
                assert w.hdr[r]["epoch"] == e + 1 and w.hdr[r]["done"] == 0
                done_ex[r] += 1
                if nxt[r] < exchanges:
                    launch(r)                            # runs ahead of slower ranks
        steps += 1
        assert steps < 5_000_000
    assert done_ex == [exchanges] * W
    if adam:
        for r in range(W):
            for (e, k), v in w.adam_seen[r].items():
                assert v == grads[e][:, k].sum(), "the Adam epilogue read a partial sum"
            assert len(w.adam_seen[r]) == exchanges * n
    return steps


@pytest.mark.parametrize("W", [1, 2, 3, 8])
@pytest.mark.parametrize("adam", [False, True])
@pytest.mark.parametrize("multicast", [False, True])
def test_exchange_protocol_under_random_schedules(W, adam, multicast):
    n, threads = 37, 2                                   # slices of unequal length (37 = 8*5 - 3), the last one short
    for seed in range(12):
        grid = [1, 2, 3, 7, 25][seed % 5]                # narrower and wider than the slice needs (with Adam: wider)
        run(W, n, grid, threads, exchanges=3, adam=adam, multicast=multicast, seed=seed)


def test_model_catches_a_broken_protocol():
    """The model is only worth something if it fails when the protocol is wrong: without the wait of barrier A a fast rank
    sums a peer's buffer before that peer has written its new gradient."""
    failures = 0
    for seed in range(20):
        try:
            run(3, 37, 3, 2, exchanges=3, adam=False, multicast=False, seed=seed, barrier_a=False)
        except AssertionError:
            failures += 1
    assert failures >= 10
