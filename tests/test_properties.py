"""Property tests (hypothesis) of the host-side logic that has no GPU dependency: batch sharding, the flat gradient
layout, the optimizer chunk table, and the oracle's tuple plumbing."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import octic_oracle as O
from octic_vits_b200.optim import build_chunk_table
from octic_vits_b200.parallel import FlatGrads, shard_batch


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 5000), st.integers(1, 16))
def test_shard_batch_partitions_the_batch(batch, world):
    spans = [shard_batch(batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == batch
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.integers(1, 7), min_size=1, max_size=3), min_size=1, max_size=8))
def test_flat_grads_views_are_aligned_and_disjoint(shapes):
    params = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes]
    fg = FlatGrads(params, fuse_accumulation=False)
    assert all(o % 4 == 0 for o in fg.offsets)
    ends = [o + p.numel() for o, p in zip(fg.offsets, params)]
    assert all(e <= o2 for e, o2 in zip(ends, fg.offsets[1:])) and ends[-1] <= fg.flat.numel()
    for i, p in enumerate(params):
        p.grad.fill_(i + 1)
    for i, (o, p) in enumerate(zip(fg.offsets, params)):
        assert bool((fg.flat[o:o + p.numel()] == i + 1).all())
    assert float(fg.flat.sum()) == sum((i + 1) * p.numel() for i, p in enumerate(params))     # padding stays zero


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(1, 40000), min_size=1, max_size=12), st.sampled_from([64, 1000, 8192]))
def test_chunk_table_covers_each_element_once(numels, chunk):
    offsets, off = [], 0
    for n in numels:
        offsets.append(off)
        off += (n + 3) // 4 * 4
    ptrs = [1 << 20 | (i << 24) for i in range(len(numels))]
    chunks, segs = build_chunk_table(numels, offsets, ptrs, [0] * len(numels), [(0.1, 1.0)] * len(numels), chunk)
    covered = torch.zeros(off, dtype=torch.int32)
    for i, (p, e, o, n, seg) in enumerate(chunks):
        assert 0 < n <= chunk and segs[seg][2] <= i < segs[seg][2] + segs[seg][3]
        assert p - ptrs[seg] == 4 * (o - offsets[seg])
        covered[o:o + n] += 1
    assert int(covered.sum()) == sum(numels) and int(covered.max()) == 1
    assert [s[3] for s in segs] == [(n + chunk - 1) // chunk for n in numels]


@settings(max_examples=50, deadline=None)
@given(st.integers(1, 3), st.integers(1, 5), st.integers(1, 6), st.integers(0, 10 ** 6))
def test_oracle_tuple_plumbing_round_trips(B, N, C, seed):
    g = torch.Generator().manual_seed(seed)
    xs = tuple(torch.randn(B, N, C, generator=g) for _ in range(4)) + (torch.randn(B, N, 2, 2 * C, generator=g),)
    for a, b in zip(O.unpack_rows(O.pack_rows(xs)), xs):
        assert torch.equal(a, b)
    for a, b in zip(O.eight_to_five(O.five_to_eight(xs)), xs):
        assert torch.equal(a, b)
    packed = O.pack_rows(xs)
    eight = O.five_to_eight(xs)
    order = (0, 1, 2, 3, 4, 6, 5, 7)                      # packed column block of 8-tuple component k (include/octic_b200.h)
    for k in range(8):
        assert torch.equal(packed[..., order[k] * C:(order[k] + 1) * C], eight[k])
    # the D8 GELU acts on the 8 components of one channel: it commutes with any permutation of the channels
    perm = torch.randperm(C, generator=g)
    permuted = tuple(t[..., perm] for t in xs[:4]) + (torch.cat((xs[4][..., :C][..., perm], xs[4][..., C:][..., perm]), -1),)
    y, yp = O.gelu_d8(xs), O.gelu_d8(permuted)
    for i in range(4):
        torch.testing.assert_close(yp[i], y[i][..., perm])
    torch.testing.assert_close(yp[4][..., :C], y[4][..., :C][..., perm])
