"""Property tests (hypothesis) of the host-side logic on either side of the hot path: the batch layout
of instance_loader.py:29-80, the dense-EV <-> incidence conversion, the instance sharder and the
parameter blob.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import tspgnn_oracle as orc
from tsp_gnn_b200 import instances as inst
from tsp_gnn_b200 import params as P
from tsp_gnn_b200 import sharding

sizes_st = st.lists(st.integers(min_value=3, max_value=12), min_size=1, max_size=6)


@settings(max_examples=25, deadline=None)
@given(sizes=sizes_st, seed=st.integers(0, 10_000), conn=st.sampled_from([1.0, 0.7, 0.4]))
def test_batch_builder_equals_reference_loops(sizes, seed, conn):
    insts = inst.synth_instances(sizes, seed=seed, connectivity=conn, two_opt_sweeps=0)
    EV, W, C, y, nv, ne = inst.create_batch(insts, dev=0.05)
    EVr, Wr, Cr, yr, nvr, ner = orc.create_batch_ref(insts, dev=0.05)
    assert np.array_equal(EV.toarray(), EVr) and np.array_equal(W, Wr) and np.allclose(C, Cr, rtol=0, atol=1e-15)
    assert list(y) == list(yr) and list(nv) == list(nvr) and list(ne) == list(ner)
    # two non-zeros per row, src < dst, block-diagonal (instance_loader.py:56-66)
    assert np.all(EV.src < EV.dst)
    voff = np.concatenate([[0], np.cumsum(nv)])
    owner = np.repeat(np.arange(len(ne)), ne)
    assert np.all(EV.src >= voff[owner]) and np.all(EV.dst < voff[owner + 1])
    back = inst.Incidence.from_dense(EVr)
    assert np.array_equal(back.src, EV.src) and np.array_equal(back.dst, EV.dst)
    s2, d2 = orc.ev_to_coo(EVr)
    assert np.array_equal(s2, EV.src) and np.array_equal(d2, EV.dst)


@settings(max_examples=40, deadline=None)
@given(n_edges=st.lists(st.integers(1, 2000), min_size=1, max_size=40), world=st.integers(1, 8))
def test_partition_covers_every_instance_once_and_is_lpt_balanced(n_edges, world):
    parts = sharding.partition_instances(n_edges, world)
    assert len(parts) == world
    flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
    assert sorted(flat.tolist()) == list(range(len(n_edges)))
    loads = np.array([int(np.sum(np.asarray(n_edges)[p])) if len(p) else 0 for p in parts])
    # greedy longest-processing-time bound: no rank exceeds the mean by more than the largest item
    assert loads.max() <= np.sum(n_edges) / world + max(n_edges)


@settings(max_examples=20, deadline=None)
@given(sizes=st.lists(st.integers(3, 9), min_size=2, max_size=6), seed=st.integers(0, 1000), world=st.integers(2, 3))
def test_shards_reassemble_the_batch(sizes, seed, world):
    EV, W, C, y, nv, ne = inst.synth_batch(sizes, seed=seed)
    parts = sharding.partition_instances(ne, world)
    seen_edges = 0
    for idx in parts:
        s, d, w, c, pv, pe = sharding.take_instances(idx, EV.src, EV.dst, W, C, nv, ne)
        assert list(pv) == [int(nv[k]) for k in idx] and list(pe) == [int(ne[k]) for k in idx]
        assert len(s) == int(np.sum(pe)) and (len(s) == 0 or (s.min() >= 0 and d.max() < int(np.sum(pv))))
        # local ids keep the edge structure of every instance
        eoff = np.concatenate([[0], np.cumsum(ne)])
        voff = np.concatenate([[0], np.cumsum(nv)])
        lo = 0
        vacc = 0
        for k in idx:
            m = int(ne[k])
            assert np.array_equal(s[lo:lo + m] - vacc, EV.src[eoff[k]:eoff[k + 1]] - voff[k])
            assert np.array_equal(d[lo:lo + m] - vacc, EV.dst[eoff[k]:eoff[k + 1]] - voff[k])
            lo += m
            vacc += int(nv[k])
        seen_edges += len(s)
    assert seen_edges == int(np.sum(ne))


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 1000))
def test_parameter_blob_roundtrip(seed):
    params = P.init_params(64, seed=seed)
    blob = P.flatten(params)
    assert blob.shape == (115529,) and blob.dtype == np.float32
    back = P.unflatten(blob)
    assert sorted(back) == sorted(params) and all(np.array_equal(back[k], params[k]) for k in params)
    table, total = P.param_offsets(64)
    assert total == 115529 and sum(n for _, n, _ in table.values()) == total
