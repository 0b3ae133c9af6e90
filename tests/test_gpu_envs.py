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


@pytest.mark.parametrize('kind', ['breakout', 'snake'])
@pytest.mark.parametrize('groups', [2, 4])
def test_pool_mode_round_robin_groups(kind, groups):
    """batch_size < num_envs (reference pool mode, vector.py:345-390): G groups returned round-robin, each stepped on its
    own stream.  Envs are independent, so every env's trajectory must equal the oracle's for the same per-env action
    sequence, and ids / shapes follow the reference (agents_per_batch = batch_size)."""
    n, h = 64, 30
    b = n // groups
    vec = pvec.make(ocean.env_creator(kind), env_kwargs=FAST_END[kind], num_envs=n, backend=pvec.B200, batch_size=b)
    assert isinstance(vec, pvec.B200Pool) and vec.agents_per_batch == b and vec.num_agents == n and vec.num_envs == b
    ora = OracleVec(kind, n, iparam=IPARAM[kind](FAST_END[kind]))
    tape = tape_for(kind, h, n, seed=4)
    vec.async_reset(7)
    ora.async_reset(7)
    for t in range(h):
        oo, orr, ot, _, _, _, _ = ora.recv()
        for g in range(groups):
            o, r, term, trunc, infos, ids, mask = vec.recv()
            lo, hi = g * b, (g + 1) * b
            assert np.array_equal(ids, np.arange(lo, hi))
            assert np.array_equal(cpu(o), oo[lo:hi]) and np.array_equal(cpu(r), orr[lo:hi]) and np.array_equal(cpu(term), ot[lo:hi])
            vec.send(torch.as_tensor(tape[t, lo:hi], device='cuda'))
        ora.send(tape[t])
    vec.close()


def test_pool_mode_through_evaluate_train():
    """Pool-mode vecenv through create/evaluate/train: rows land in arrival order t*N + e, so the stored rollout replays
    through the oracle exactly like the non-pool one."""
    import pufferlib_b200
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    n, h, b = 64, 16, 16
    vec = pvec.make(ocean.env_creator('breakout'), num_envs=n, backend=pvec.B200, batch_size=b)
    torch.manual_seed(0)
    pol = cleanrl.Policy(models.Default(vec.driver_env), fused_sample=True, seed=1).cuda()
    cfg = pufferlib_b200.namespace(
        seed=1, torch_deterministic=True, env='breakout', batch_size=n * h, bptt_horizon=8, minibatch_size=n * h // 2,
        cpu_offload=False, device='cuda', compile=False, learning_rate=2.5e-4, gamma=0.99, gae_lambda=0.95,
        update_epochs=1, norm_adv=True, clip_coef=0.1, clip_vloss=True, vf_clip_coef=0.1, vf_coef=0.5, ent_coef=0.01,
        max_grad_norm=0.5, target_kl=None, anneal_lr=False, total_timesteps=10 ** 9)
    data = clean_pufferl.create(cfg, vec, pol)
    ora = OracleVec('breakout', n)
    ora.async_reset(1)
    for it in range(2):
        clean_pufferl.evaluate(data)
        exp = data.experience
        acts, obs = cpu(exp.actions).reshape(h, n), cpu(exp.obs).reshape(h, n, 128)
        for t in range(h):
            o, r, d, _, _, _, _ = ora.recv()
            assert np.array_equal(o, obs[t]), (it, t)
            ora.send(acts[t])
        clean_pufferl.train(data)
        assert np.isfinite(data.losses.policy_loss)
        assert data.global_step == (it + 1) * n * h
    clean_pufferl.close(data)


def test_snake_kernel_variants():
    """The 16-lanes-per-env kernel of round 1 stays selectable (A/B measurements) and bit-exact against the oracle too."""
    from pufferlib_b200 import _native
    lib = _native.lib()
    try:
        _native.check(lib.pb_snake_set_variant(16))
        compare_run('snake', 37, h=120, seed=21, env_kwargs=FAST_END['snake'])
    finally:
        lib.pb_snake_set_variant(4)
    compare_run('snake', 37, h=120, seed=21, env_kwargs=FAST_END['snake'])
    compare_run('snake', 3, h=300, seed=22, env_kwargs={})          # n * 4 lanes not a multiple of 32
