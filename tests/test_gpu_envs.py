"""Device breakout / snake / pong vs the oracle's scalar C restatement of oracle/SPEC.md: bit-exact obs, rewards,
terminals, EpisodeStats infos, through the vector API (own buffers and bound-rollout rows)."""
import numpy as np
import pytest
import torch

import pufferlib_b200.vector as pvec
from pufferlib_b200 import clean_pufferl
from pufferlib_b200.environments import ocean
from oracle.envs import OracleVec, NUM_ACTIONS, OBS

pytestmark = pytest.mark.gpu

FAST_END = {'breakout': dict(max_ticks=150), 'snake': dict(max_ticks=40), 'pong': dict(max_score=1, max_ticks=120)}
IPARAM = {'breakout': lambda k: [k.get('max_ticks', 0)], 'snake': lambda k: [k.get('max_ticks', 0)],
          'pong': lambda k: [k.get('max_score', 0), k.get('max_ticks', 0)]}


def cpu(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def tape_for(kind, h, n, seed=0):
    return np.random.default_rng(seed).integers(0, NUM_ACTIONS[kind], size=(h, n), dtype=np.int64)


def compare_run(kind, n, h, seed, env_kwargs, bound=False, offset=0, check_infos=True):
    vec = pvec.make(ocean.env_creator(kind), env_kwargs=env_kwargs, num_envs=n,
                    backend=pvec.B200.options(exact_infos=check_infos, env_index_offset=offset))
    ora = OracleVec(kind, n, env_index_offset=offset, iparam=IPARAM[kind](env_kwargs))
    tape = tape_for(kind, h, n, seed)
    vec.async_reset(seed)
    ora.async_reset(seed)
    exp = None
    if bound:
        shape, dtype = OBS[kind]
        rows = 8
        exp = clean_pufferl.Experience(n * rows, 4, n * rows, shape, dtype, ())
        vec.bind_rollout(exp)
    n_eps = 0
    for t in range(h + 1):
        o, r, term, trunc, infos, ids, mask = vec.recv()
        oo, orr, ot, otr, oinf, _, om = ora.recv()
        assert np.array_equal(cpu(o), oo), f'{kind}: obs differ at step {t}'
        assert np.array_equal(cpu(r).view(np.uint32), orr.view(np.uint32)), f'{kind}: reward bits differ at step {t}'
        assert np.array_equal(cpu(term), ot), f'{kind}: terminals differ at step {t}'
        assert not cpu(trunc).any() and cpu(mask).all()
        if check_infos:
            assert len(infos) == len(oinf)
            for a, b in zip(infos, oinf):
                assert a['episode_length'] == b['episode_length'] and a['score'] == b['score']
                assert np.isclose(a['episode_return'], b['episode_return'], rtol=1e-12, atol=0)
            n_eps += len(infos)
        if t < h:
            a = torch.as_tensor(tape[t], device='cuda')
            if bound:
                z = torch.zeros(n, device='cuda')
                exp.store(o, z, a, z, r, term, ids, mask)
                if exp.full:
                    exp.sort_training_data()
            vec.send(a)
            ora.send(tape[t])
    vec.close()
    return n_eps


@pytest.mark.parametrize('kind', ['breakout', 'snake', 'pong'])
@pytest.mark.parametrize('n', [1, 37, 1024])
def test_env_vs_oracle_fast_episodes(kind, n):
    """Short episodes (small max_ticks / max_score) so auto-resets, reset rows and infos are all exercised."""
    eps = compare_run(kind, n, h=200 if n < 1024 else 160, seed=3 + n, env_kwargs=FAST_END[kind])
    assert eps > 0


@pytest.mark.parametrize('kind', ['breakout', 'snake', 'pong'])
def test_env_vs_oracle_default_params(kind):
    compare_run(kind, 64, h=400, seed=11, env_kwargs={})


@pytest.mark.parametrize('kind', ['breakout', 'snake', 'pong'])
def test_env_bound_rollout_rows(kind):
    """Step outputs written straight into Experience rows (and the carry-over at rollout boundaries)."""
    compare_run(kind, 33, h=50, seed=5, env_kwargs=FAST_END[kind], bound=True)


@pytest.mark.parametrize('kind', ['breakout', 'snake', 'pong'])
def test_env_shards_seed_by_global_index(kind):
    """Multi-GPU sharding: a shard created with env_index_offset=k is bit-identical to envs [k, k+n) of the
    unsharded run -- checked here against the oracle given the same offset."""
    compare_run(kind, 16, h=60, seed=9, env_kwargs=FAST_END[kind], offset=1000, check_infos=False)


def test_breakout_c2_full_size():
    """BASELINE config C2 width: N = 16384, a few hundred steps, every row compared."""
    compare_run('breakout', 16384, h=64, seed=1, env_kwargs={}, check_infos=False)


def test_snake_c3_full_size():
    compare_run('snake', 65536, h=48, seed=1, env_kwargs={}, check_infos=False)


def test_pong_c4_full_size():
    compare_run('pong', 8192, h=6, seed=1, env_kwargs={}, check_infos=False)


def test_observation_spaces():
    for kind in ('breakout', 'snake', 'pong'):
        vec = pvec.make(ocean.env_creator(kind), num_envs=2, backend=pvec.B200)
        shape, dtype = OBS[kind]
        assert vec.single_observation_space.shape == shape and vec.single_observation_space.dtype == dtype
        assert vec.single_action_space.n == NUM_ACTIONS[kind]
        vec.close()
