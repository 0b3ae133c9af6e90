"""BASELINE.json configs C3 (snake, MLP) and C4 (pong, NatureCNN) THROUGH create / evaluate / train, not just their env
kernels: per iteration the stored rollout (obs / reward / done rows) is replayed through the oracle vectoriser with the
actions the policy sampled, GAE / returns are checked against the oracle's C restatement of c_gae.pyx on the stored
values, and one full PPO update runs on it.  Reference: clean_pufferl.py:30-292 (create / evaluate / train),
pufferlib/models.py:12-62 (Default), :113-157 (Convolutional), environments/atari/torch.py:8-18."""
import numpy as np
import pytest
import torch

import pufferlib_b200
import pufferlib_b200.vector as pvec
from pufferlib_b200 import clean_pufferl, models
from pufferlib_b200.environments import ocean
from pufferlib_b200.frameworks import cleanrl
from oracle import gae as ogae
from oracle.envs import OracleVec, OBS

pytestmark = pytest.mark.gpu


def cpu(x):
    return x.detach().cpu().numpy()


def ppo_config(env, n, h, bptt, minibatches, cuda_graph):
    return pufferlib_b200.namespace(
        seed=1, torch_deterministic=True, env=env, batch_size=n * h, bptt_horizon=bptt,
        minibatch_size=n * h // minibatches, cpu_offload=False, device='cuda', compile=False, learning_rate=2.5e-4,
        gamma=0.99, gae_lambda=0.95, update_epochs=2, norm_adv=True, clip_coef=0.1, clip_vloss=True, vf_clip_coef=0.1,
        vf_coef=0.5, ent_coef=0.01, max_grad_norm=0.5, target_kl=None, anneal_lr=False, total_timesteps=10 ** 10,
        cuda_graph=cuda_graph)


def run_config(kind, n, h, bptt, minibatches, cuda_graph, iterations, replay_envs=None):
    """evaluate/train `iterations` times; every stored rollout must replay bit-exactly through the oracle (all envs, or
    the first `replay_envs` of them when the oracle would be too slow) and its GAE must match the oracle within 1e-5."""
    vec = pvec.make(ocean.env_creator(kind), num_envs=n, backend=pvec.B200)
    torch.manual_seed(0)
    net = models.Convolutional(vec.driver_env) if kind == 'pong' else models.Default(vec.driver_env)
    pol = cleanrl.Policy(net, fused_sample=True, seed=3).cuda()
    cfg = ppo_config(kind, n, h, bptt, minibatches, cuda_graph)
    data = clean_pufferl.create(cfg, vec, pol)
    m = n if replay_envs is None else replay_envs
    ora = OracleVec(kind, m)
    ora.collect_infos = False
    ora.async_reset(1)
    shape, _ = OBS[kind]
    before = [p.detach().clone() for p in pol.parameters()]
    for it in range(iterations):
        stats, _ = clean_pufferl.evaluate(data)
        exp = data.experience
        acts = cpu(exp.actions).reshape(h, n)
        rew, done = cpu(exp.rewards).reshape(h, n), cpu(exp.dones).reshape(h, n)
        assert acts.min() >= 0 and acts.max() < vec.single_action_space.n
        obs = exp.obs.view(h, n, *shape)
        for t in range(h):
            o, r, d, _, _, _, _ = ora.recv()
            assert np.array_equal(o, cpu(obs[t, :m])), (kind, it, t)
            assert np.array_equal(r.view(np.uint32), rew[t, :m].view(np.uint32)), (kind, it, t)
            assert np.array_equal(d.astype(np.float32), done[t, :m]), (kind, it, t)
            ora.send(acts[t, :m])
        values = cpu(exp.values)
        assert np.isfinite(values).all() and np.isfinite(cpu(exp.logprobs)).all()
        clean_pufferl.train(data)
        # GAE on the stored (sorted) batch vs the oracle's restatement of c_gae.pyx:11-32
        idx = np.asarray(clean_pufferl._LazyIdxs(n, h))
        ref = ogae.compute_gae(done.reshape(-1)[idx], values[idx], rew.reshape(-1)[idx], 0.99, 0.95)
        got = cpu(exp.advantages)
        assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-5
        # Experience.returns is the reference's returns_np, literally (clean_pufferl.py:476): SORTED-order advantages plus
        # ARRIVAL-order values at the same flat index (it only feeds the explained variance)
        assert np.allclose(cpu(exp.returns), got + values, rtol=1e-5, atol=1e-5)
        for k in ('policy_loss', 'value_loss', 'entropy', 'approx_kl', 'clipfrac'):
            assert np.isfinite(getattr(data.losses, k)), k
        assert data.global_step == (it + 1) * n * h
    moved = sum(float((p.detach() - b).abs().max()) for p, b in zip(pol.parameters(), before))
    assert moved > 0, 'the update must change the parameters'
    clean_pufferl.close(data)
    return data


def test_snake_c3_shape_through_evaluate_train():
    """C3 horizon (H = 256, bptt 16, 4 minibatches) at 2048 envs: eager, then captured rollout + update graphs."""
    data = run_config('snake', 2048, 256, 16, 4, cuda_graph=True, iterations=4)
    assert data.graph_replays >= 2 and data.train_graph_state == 2
    assert data.manual_update is not None, 'models.Default on snake must take the hand-written update'


def test_snake_c3_full_size_through_evaluate_train():
    """BASELINE configs[2] at full size: 65536 envs x 256 steps (16.7 M agent-steps per rollout); the oracle replays the
    first 512 envs (envs are independent), GAE is checked on the whole batch."""
    run_config('snake', 65536, 256, 16, 4, cuda_graph=False, iterations=1, replay_envs=512)


def test_pong_c4_cnn_through_evaluate_train():
    """C4: pong (4,84,84) uint8 frames + the NatureCNN policy (models.Convolutional, cuDNN) through the same loop."""
    data = run_config('pong', 256, 32, 8, 2, cuda_graph=False, iterations=2)
    assert data.manual_update is None


def test_pong_c4_cnn_graphed():
    data = run_config('pong', 128, 16, 8, 2, cuda_graph=True, iterations=3)
    assert data.graph_replays >= 1
