"""CPU: the placement algorithm of sample_fine_fast_kernel (csrc/sampler.cu), restated in NumPy and checked against
np.sort on oracle-generated samples. The kernel replaces sort(concat(t_coarse, t_fine)) by

  rank among the fine samples  = (#fine samples in lower u-buckets) + (rank by t inside the bucket, ties by arrival)
  coarse samples below a fine  = its bin index + (t_coarse[bin] < t_fine)
  coarse samples               = the slots the fine samples left empty, in order

followed by a CHECK (exactly N_coarse slots were left empty, output ascending) whose failure sends the ray to a full
sort. This file pins the two properties the GPU parity tests rely on: whenever the check passes the result IS the exact
sort, and on realistic inputs (flat ... nearly one-hot pdfs) the check passes for essentially every ray, so the
fallback is a correctness net, not the common path. The GPU tests (tests/test_gpu_kernels.py) compare the kernel itself
with np.sort bit for bit, fallback cases included."""
import numpy as np
import pytest

from oracle import ray_march as rm

F32 = np.float32


def placement(t_c, t_f, idx, u, presorted=False):
    """One ray. Returns the merged samples, or None where the kernel's self-check would fail."""
    Nc, Nf = len(t_c), len(t_f)
    if presorted:                       # in-kernel uniforms arrive ascending: rank = index
        frank = np.arange(Nf)
    else:
        key = np.clip((u * F32(Nf)).astype(np.int64), 0, Nf - 1)
        cnt = np.bincount(key, minlength=Nf)
        base = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        slot = np.zeros(Nf, int)
        seen = {}
        for j, k in enumerate(key):     # arrival order inside a bucket is arbitrary on the GPU; any order must work
            slot[j] = seen.get(k, 0)
            seen[k] = slot[j] + 1
        r = np.zeros(Nf, int)
        for j in range(Nf):
            for o in np.where(key == key[j])[0]:
                if t_f[o] < t_f[j] or (t_f[o] == t_f[j] and slot[o] < slot[j]):
                    r[j] += 1
        frank = base[key] + r
    below = idx + (t_c[idx] < t_f)
    out = np.full(Nc + Nf, np.nan, F32)
    out[frank + below] = t_f
    holes = np.isnan(out)
    if holes.sum() != Nc:
        return None
    out[holes] = t_c
    if not np.all(out[:-1] <= out[1:]):
        return None
    return out


def _case(Nc, Nf, sharp, seed, B=96):
    rng = np.random.default_rng(seed)
    near = np.full((B, 1), 0.425, F32); far = np.full((B, 1), 1.275, F32)
    z = np.zeros((B, 3), F32)
    o = rm.create_input_batch_coarse_model(Nc, True, True, z, z + 1, near, far, rng.random((B, Nc), dtype=F32))
    w = (rng.random((B, Nc), dtype=F32) ** sharp).astype(F32)
    w[rng.random((B, Nc)) < 0.3] = 0.0
    u = rng.random((B, Nf), dtype=F32)
    ref = rm.create_input_batch_fine_model(z, z + 1, w, o["bin_data"], o["t_vals"], u, return_debug=True)
    return o["t_vals"], ref["t_vals_fine"], ref["piece_idxs"], u, ref["t_vals"]


@pytest.mark.parametrize("Nc,Nf,sharp", [(64, 128, 1), (64, 128, 8), (64, 128, 64), (128, 256, 4)])
def test_placement_equals_sort_and_rarely_falls_back(Nc, Nf, sharp):
    t_c, t_f, idx, u, t_sorted = _case(Nc, Nf, sharp, seed=Nc + sharp)
    fell_back = 0
    for i in range(t_c.shape[0]):
        out = placement(t_c[i], t_f[i], idx[i], u[i])
        if out is None:
            fell_back += 1
        else:
            assert np.array_equal(out, t_sorted[i])      # the oracle's own sorted concat
    assert fell_back <= 1


def test_placement_with_presorted_uniforms():
    """The in-kernel draw produces u ascending, so the rank among the fine samples is the index."""
    t_c, t_f, idx, u, _ = _case(64, 128, 4, seed=5)
    order = np.argsort(u, axis=1, kind="stable")
    for i in range(t_c.shape[0]):
        tf, ix = t_f[i][order[i]], idx[i][order[i]]
        out = placement(t_c[i], tf, ix, u[i][order[i]], presorted=True)
        if out is not None:                              # t is monotone in u only up to rounding across bins
            assert np.array_equal(out, np.sort(np.concatenate([t_c[i], t_f[i]])))
    # and the check does catch a wrong order: swap two fine samples that differ
    tf = t_f[0][order[0]].copy(); ix = idx[0][order[0]].copy()
    a, b = 3, 90
    assert tf[a] != tf[b]
    tf[[a, b]] = tf[[b, a]]; ix[[a, b]] = ix[[b, a]]
    assert placement(t_c[0], tf, ix, None, presorted=True) is None


def test_placement_adversarial_inputs_are_exact_or_rejected():
    rng = np.random.default_rng(9)
    t_c, t_f, idx, u, _ = _case(64, 128, 2, seed=9, B=8)
    # every sample in one bucket with identical t (ties by arrival slot)
    out = placement(t_c[0], np.full(128, t_c[0][10], F32), np.full(128, 10), np.zeros(128, F32))
    assert out is not None and np.array_equal(out, np.sort(np.concatenate([t_c[0], np.full(128, t_c[0][10], F32)])))
    # bin indices that do not match the samples (a corrupted hint) are rejected, never silently misplaced
    bad = placement(t_c[1], t_f[1], rng.permutation(idx[1]), u[1])
    assert bad is None or np.array_equal(bad, np.sort(np.concatenate([t_c[1], t_f[1]])))
