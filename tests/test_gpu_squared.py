"""Device Squared + vectoriser vs the reference's own trajectories (tests/golden/squared_*.npz): bit-exact obs,
rewards, terminals, truncations, masks; EpisodeStats infos to 1e-12."""
import numpy as np
import pytest
import torch

import pufferlib_b200.vector as pvec
from pufferlib_b200.environments import ocean
from pufferlib_b200.exceptions import APIUsageError

pytestmark = pytest.mark.gpu

CASES = ['squared_c1', 'squared_n5_seed42', 'squared_n1_seed7', 'squared_d2_n33', 'squared_d5_n8']


def to_np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def run_case(g, backend, actions_on_device):
    n, seed, h, d = int(g['num_envs']), int(g['seed']), int(g['horizon']), int(g['distance_to_target'])
    vec = pvec.make(ocean.env_creator('squared'), env_kwargs={'distance_to_target': d}, num_envs=n, backend=backend)
    vec.async_reset(seed)
    info_rows = []
    for t in range(h + 1):
        o, r, term, trunc, infos, ids, mask = vec.recv()
        assert np.array_equal(to_np(o), g['obs_i8'][t].astype(np.float32)), f'obs step {t}'
        assert np.array_equal(to_np(r).view(np.uint32), g['rewards'][t].view(np.uint32)), f'reward bits step {t}'
        assert np.array_equal(to_np(term), g['terminals'][t]), f'terminals step {t}'
        assert np.array_equal(to_np(trunc), g['truncations'][t])
        assert np.array_equal(to_np(mask), g['masks'][t])
        assert np.array_equal(ids, np.arange(n))
        for k, i in enumerate(infos):
            info_rows.append((t, k, i['episode_return'], i['episode_length'], i['score']))
        if t < h:
            a = g['actions'][t]
            vec.send(torch.as_tensor(a, device='cuda') if actions_on_device else a)
    got = np.asarray(info_rows, dtype=np.float64).reshape(-1, 5)
    assert got.shape == g['infos'].shape
    assert np.array_equal(got[:, [0, 1, 3, 4]], g['infos'][:, [0, 1, 3, 4]])
    assert np.allclose(got[:, 2], g['infos'][:, 2], rtol=1e-12, atol=1e-15)   # fp64 sum vs python's compensated sum
    vec.close()


@pytest.mark.parametrize('case', CASES)
def test_squared_device_tensors(golden, case):
    run_case(golden(case), pvec.B200.options(exact_infos=True), actions_on_device=True)


@pytest.mark.parametrize('case', CASES)
def test_squared_host_buffers(golden, case):
    """host_buffers=True: numpy in / numpy out like vector.Serial."""
    run_case(golden(case), pvec.B200.options(host_buffers=True), actions_on_device=False)


def test_squared_matches_oracle_at_scale():
    """N beyond the goldens: device vs the oracle restatement (several twists of the MT stream, ragged warps)."""
    from oracle.squared import SquaredSerial
    n, seed, h = 1000, 77, 60
    tape = np.random.default_rng(1).integers(0, 8, size=(h, n))
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=pvec.B200.options(exact_infos=True))
    ora = SquaredSerial(n)
    vec.async_reset(seed)
    ora.async_reset(seed)
    for t in range(h + 1):
        o, r, term, trunc, infos, ids, mask = vec.recv()
        oo, orr, ot, otr, oinf, _, om = ora.recv()
        assert np.array_equal(to_np(o), oo) and np.array_equal(to_np(r).view(np.uint32), orr.view(np.uint32))
        assert np.array_equal(to_np(term), ot)
        assert len(infos) == len(oinf)
        if t < h:
            vec.send(tape[t])
            ora.send(tape[t])
    vec.close()


def test_device_episode_stats_reduction(golden):
    g = golden('squared_c1')
    n, seed, h = int(g['num_envs']), int(g['seed']), int(g['horizon'])
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=pvec.B200)
    vec.async_reset(seed)
    for t in range(h):
        vec.recv()
        vec.send(torch.as_tensor(g['actions'][t], device='cuda'))
    vec.recv()
    means, count = vec.episode_stats()
    ref = g['infos']
    assert count == len(ref)
    assert np.isclose(means['episode_return'], ref[:, 2].mean(), rtol=1e-9, atol=1e-12)
    assert np.isclose(means['episode_length'], ref[:, 3].mean()) and np.isclose(means['score'], ref[:, 4].mean())
    assert vec.episode_stats() == ({}, 0)
    vec.close()


def test_api_order_errors():
    """tests/test_api.py conventions of the reference: recv before reset, send before recv, bad actions."""
    vec = pvec.make(ocean.env_creator('squared'), num_envs=4, backend=pvec.B200)
    with pytest.raises(APIUsageError):
        vec.recv()
    with pytest.raises(APIUsageError):
        vec.send(np.zeros(4, dtype=np.int64))
    vec.async_reset(1)
    with pytest.raises(APIUsageError):
        vec.send(np.zeros(4, dtype=np.int64))     # send before recv
    vec.recv()
    with pytest.raises(APIUsageError):
        vec.recv()                                # double recv
    vec2 = pvec.make(ocean.env_creator('squared'), num_envs=4, backend=pvec.B200)
    vec2.reset(seed=1)
    with pytest.raises(APIUsageError):
        vec2.send(np.full(4, 9, dtype=np.int64))  # outside Discrete(8)
    o, r, d, t, i = vec.step(np.zeros(4, dtype=np.int64))
    assert o.shape == (4, 7, 7)
    with pytest.raises(APIUsageError):
        pvec.make(ocean.env_creator('squared'), num_envs=0, backend=pvec.B200)
    with pytest.raises(APIUsageError):
        pvec.make(ocean.env_creator('squared'), num_envs=4, backend=pvec.B200, bogus=1)
    vec.close(); vec2.close()


@pytest.mark.parametrize('case', ['squared_mp_n8_w2', 'squared_mp_n12_w4'])
@pytest.mark.parametrize('bound', [False, True])
def test_squared_multiprocessing_golden_sync(golden, case, bound):
    """backend=B200 with num_workers=W against the reference's own Multiprocessing backend run with W real worker
    processes (tests/golden/generate.py::run_squared_multiprocessing): every worker has its own process-global MT19937
    stream (vector.py:168-190, 424-428), which changes the targets of every auto-reset relative to Serial."""
    g = golden(case)
    n, w, seed, h = int(g['num_envs']), int(g['num_workers']), int(g['seed']), int(g['horizon'])
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, num_workers=w, batch_size=n,
                    backend=pvec.B200.options(exact_infos=True))
    assert isinstance(vec, pvec.B200) and len(vec._shards) == w
    if bound:
        from pufferlib_b200 import clean_pufferl
        exp = clean_pufferl.Experience(n * 8, 4, n * 8, (7, 7), np.float32, ())
        vec.bind_rollout(exp)
    vec.async_reset(seed)
    infos_all = []
    for t in range(h + 1):
        o, r, d, tr, infos, ids, m = vec.recv()
        assert np.array_equal(to_np(o), g['obs_i8'][t].astype(np.float32)), f'obs step {t}'
        assert np.array_equal(to_np(r).view(np.uint32), g['rewards'][t].view(np.uint32)), f'reward bits step {t}'
        assert np.array_equal(to_np(d), g['terminals'][t])
        for k, i in enumerate(infos):
            infos_all.append((t, k, i['episode_return'], i['episode_length'], i['score']))
        if t < h:
            a = torch.as_tensor(g['actions'][t], device='cuda')
            if bound:
                z = torch.zeros(n, device='cuda')
                exp.store(o, z, a, z, r, d, ids, m)
                if exp.full:
                    exp.sort_training_data()
            vec.send(a)
    ref = g['infos']
    got = np.asarray(infos_all, dtype=np.float64).reshape(-1, 5)
    assert got.shape == ref.shape and np.array_equal(got[:, [0, 1, 3, 4]], ref[:, [0, 1, 3, 4]])
    assert np.allclose(got[:, 2], ref[:, 2], rtol=1e-12, atol=0)
    vec.close()


def test_squared_multiprocessing_golden_pool():
    """Pool mode (batch_size = one worker's envs): groups come back round-robin instead of first-ready, but every env's
    trajectory under its own action sequence is the reference Multiprocessing trajectory (per-worker streams)."""
    from conftest import load_golden
    g = load_golden('squared_mp_n12_w4')
    n, w, seed, h = int(g['num_envs']), int(g['num_workers']), int(g['seed']), int(g['horizon'])
    b = n // w
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, num_workers=w, batch_size=b, backend=pvec.B200)
    assert isinstance(vec, pvec.B200Pool) and len(vec.groups) == w
    vec.async_reset(seed)
    for t in range(h + 1):
        for k in range(w):
            o, r, d, tr, infos, ids, m = vec.recv()
            lo, hi = k * b, (k + 1) * b
            assert np.array_equal(ids, np.arange(lo, hi))
            assert np.array_equal(to_np(o), g['obs_i8'][t, lo:hi].astype(np.float32)), (t, k)
            assert np.array_equal(to_np(d), g['terminals'][t, lo:hi])
            vec.send(torch.as_tensor(g['actions'][min(t, h - 1), lo:hi], device='cuda'))
    # two workers per batch: each group holds two RNG shards
    vec2 = pvec.make(ocean.env_creator('squared'), num_envs=n, num_workers=w, batch_size=2 * b, backend=pvec.B200)
    assert len(vec2.groups) == 2 and all(len(v._shards) == 2 for v in vec2.groups)
    vec2.async_reset(seed)
    for t in range(h + 1):
        for k in range(2):
            o, r, d, tr, infos, ids, m = vec2.recv()
            lo, hi = k * 2 * b, (k + 1) * 2 * b
            assert np.array_equal(to_np(o), g['obs_i8'][t, lo:hi].astype(np.float32)), (t, k)
            vec2.send(torch.as_tensor(g['actions'][min(t, h - 1), lo:hi], device='cuda'))
    vec.close(); vec2.close()


def test_seed_lists(golden):
    """make_seeds list form (vector.py:639-650): a run of consecutive integers is what Multiprocessing hands a worker."""
    g = golden('squared_n5_seed42')
    vec = pvec.make(ocean.env_creator('squared'), num_envs=5, backend=pvec.B200)
    vec.async_reset([42, 43, 44, 45, 46])
    assert np.array_equal(to_np(vec.recv()[0]), g['obs_i8'][0].astype(np.float32))
    with pytest.raises(APIUsageError):
        vec.async_reset([1, 2, 3])               # wrong length
    with pytest.raises(APIUsageError):
        vec.async_reset([5, 4, 3, 2, 1])         # not representable on the device
    with pytest.raises(APIUsageError):
        vec.async_reset('seed')
    vec.close()
