"""CPU, world_size 2 over gloo: chain partitioning + per-generation all-gather of the driver
(bart_b200/driver.py) with a stand-in evaluator, and the BandModel input converter against a
literal restatement of BARTfunc.py:333-347."""
import os
import numpy as np
import pytest

from bart_b200 import driver


def test_partition_covers_all_chains():
    for n in (1, 3, 10, 17, 1000):
        for w in (1, 2, 3, 8):
            spans = [driver.partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_eval(params):
    # deterministic function of the parameters, 4 "filters"
    return np.stack([params.sum(axis=1), params[:, 0] * 2, np.cos(params[:, -1]), params.prod(axis=1)], axis=1)


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from util import TorchComm
    comm = TorchComm(dist)
    rng = np.random.default_rng(7)
    for nchains in (10, 7, 2, 1):
        params = rng.uniform(-1, 1, (nchains, 5))
        got = driver.evaluate_generation(_fake_eval, params, comm)
        np.save(os.path.join(tmp, "out_%d_%d.npy" % (nchains, rank)), got)
    dist.destroy_process_group()


def test_generation_allgather_gloo(tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(7)
    for nchains in (10, 7, 2, 1):
        params = rng.uniform(-1, 1, (nchains, 5))
        want = _fake_eval(params)
        for rank in range(2):
            got = np.load(str(tmp_path / ("out_%d_%d.npy" % (nchains, rank))))
            assert got.shape == want.shape and np.array_equal(got, want)


class _StubTransit:
    eclipse = True
    nfilters = 3

    def __init__(self):
        self.calls = []

    def set_batch_knobs(self, n, **kw):
        self.calls.append(("knobs", n, sorted(kw)))

    def bandflux_batch(self, prof):
        self.calls.append(("batch", prof.shape))
        return np.tile(prof[:, :1], (1, 3)), np.zeros(prof.shape[0], dtype=np.int32)


def test_bandmodel_input_converter():
    species = ["H", "He", "C", "N", "O", "H2", "CO", "CO2", "CH4", "H2O"]
    nl = 6
    press = np.logspace(2, -5, nl)
    base = np.tile([1e-9, 0.15, 1e-9, 1e-9, 1e-9, 0.8496, 1e-4, 1e-4, 1e-4, 1e-4], (nl, 1))
    tr = _StubTransit()
    pt = lambda p, x: np.full(len(p), x[0])              # isothermal
    bm = driver.BandModel(tr, press, species, base, ["H2O", "CH4"], pt, npt=1, fit_radius=False)
    params = np.array([[1500.0, 1.0, -2.0],      # fine
                       [3500.0, 0.0, 0.0],       # T above Tmax -> rejected
                       [1200.0, 4.5, 0.0]])      # metals sum > 1 -> rejected
    prof, rejected, knobs = bm.profiles(params)
    assert list(rejected) == [False, True, True] and knobs == {}
    # literal BARTfunc.py:333-347 for model 0
    ap = base.T.copy()
    ap[species.index("H2O")] = base[:, 9] * 10.0 ** 1.0
    ap[species.index("CH4")] = base[:, 8] * 10.0 ** -2.0
    imet = [i for i, s in enumerate(species) if s not in ("H2", "He")]
    q = 1.0 - ap[imet].sum(axis=0)
    ratio = base[:, 5] / base[:, 1]
    ap[5] = ratio * q / (1.0 + ratio)
    ap[1] = q / (1.0 + ratio)
    want = np.concatenate([np.full(nl, 1500.0), ap.ravel()])
    assert np.allclose(prof[0], want, rtol=0, atol=0)
    assert np.allclose(prof[0, nl:].reshape(len(species), nl).sum(axis=0), 1.0)
    flux = bm.evaluate(params)
    assert flux.shape == (3, 3) and (flux[1:] == -1).all() and (flux[0] == 1500.0).all()
    assert tr.calls == [("batch", (1, (len(species) + 1) * nl))]


def test_library_chain_block_matches_partition(built):
    """The device DE-MC loop's chain -> rank map (bart_chain_block, pure host arithmetic) is the
    driver's `partition`; blocks tile [0, nchains) for every world size."""
    import ctypes as C
    from bart_b200 import api
    L = api.lib()
    for nchains in (3, 10, 17, 4096):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for rank in range(world):
                lo, hi = C.c_int(), C.c_int()
                L.bart_chain_block(nchains, world, rank, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == driver.partition(nchains, world, rank)
                assert lo.value == nxt
                nxt = hi.value
            assert nxt == nchains
