"""Experience (store -> GAE -> flatten_batch -> adv-norm) on the device vs the reference's own outputs
(tests/golden/experience_*.npz) and the numpy oracle; plus the standalone train-prep kernels at larger sizes."""
import numpy as np
import pytest
import torch

import pufferlib_b200.vector as pvec
from pufferlib_b200 import _native, clean_pufferl
from pufferlib_b200.environments import ocean
from oracle import experience as oexp
from oracle import gae as ogae

pytestmark = pytest.mark.gpu


def cpu(x):
    return x.detach().cpu().numpy()


@pytest.mark.parametrize('case', ['experience_c1', 'experience_small', 'experience_one_mb'])
@pytest.mark.parametrize('bound', [True, False])
def test_experience_pipeline_vs_reference(golden, case, bound):
    g = golden(case)
    n, h = int(g['num_envs']), int(g['horizon'])
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=pvec.B200)
    vec.async_reset(int(g['seed']))
    exp = clean_pufferl.Experience(n * h, int(g['bptt']), int(g['minibatch_size']), (7, 7), np.float32, ())
    if bound:
        vec.bind_rollout(exp)
    dev = torch.device('cuda')
    t = 0
    while not exp.full:
        o, r, d, tr, infos, env_id, mask = vec.recv()
        a = torch.as_tensor(g['tape'][t], device=dev)
        exp.store(o, torch.as_tensor(g['values_in'][t], device=dev), a, torch.as_tensor(g['logprobs_in'][t], device=dev),
                  r, d, env_id, mask)
        vec.send(a)
        t += 1
    assert t == h
    # stored rollout == the reference's Experience arrays, bit for bit
    assert np.array_equal(cpu(exp.obs), g['stored_obs_i8'].astype(np.float32))
    for k in ('actions', 'logprobs', 'rewards', 'dones', 'values'):
        assert np.array_equal(cpu(getattr(exp, k)), g['stored_' + k]), k
    idxs = exp.sort_training_data()
    assert np.array_equal(np.asarray(idxs), g['idxs'])
    adv = cpu(exp.compute_gae(float(g['gamma']), float(g['gae_lambda'])))
    ref = g['advantages']
    assert np.max(np.abs(adv - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-5
    exp.flatten_batch()
    assert np.array_equal(cpu(exp.b_obs), g['b_obs_i8'].astype(np.float32))
    for k in ('b_actions', 'b_logprobs', 'b_dones', 'b_values'):
        assert np.array_equal(cpu(getattr(exp, k)), g[k]), k
    for k in ('b_advantages', 'b_returns'):
        assert np.allclose(cpu(getattr(exp, k)), g[k], rtol=1e-5, atol=1e-5), k
    assert np.allclose(cpu(exp.returns), g['returns_np'], rtol=1e-5, atol=1e-5)      # the literal :476 quantity
    norm = cpu(exp.normalize_advantages())
    assert np.allclose(norm, g['b_advantages_normalized'], rtol=1e-5, atol=2e-5)
    # a second rollout through the same buffers: carry-over row 0 must be the step that closed the first rollout
    o2 = vec.recv()[0]
    assert cpu(o2).shape == (n, 7, 7)
    vec.close()


@pytest.mark.parametrize('n,h,mb,bptt,obs_shape,dtype', [
    (64, 128, 2048, 16, (7, 7), np.float32), (33, 12, 36, 4, (5,), np.float32), (16, 64, 256, 8, (128,), np.float32),
    (8, 32, 64, 16, (4, 84, 84), np.uint8), (100, 10, 250, 5, (3,), np.uint8), (256, 128, 4096, 32, (116,), np.float32)])
def test_flatten_and_gather_vs_numpy_oracle(n, h, mb, bptt, obs_shape, dtype):
    rng = np.random.default_rng(n * h)
    b = n * h
    dev = torch.device('cuda')
    exp = clean_pufferl.Experience(b, bptt, mb, obs_shape, dtype, ())
    ora = oexp.Experience(b, bptt, mb, obs_shape, dtype)
    if np.dtype(dtype) == np.uint8:
        obs = rng.integers(0, 256, size=(b, *obs_shape), dtype=np.uint8)
    else:
        obs = rng.standard_normal((b, *obs_shape)).astype(np.float32)
    fields = dict(actions=rng.integers(0, 6, size=b).astype(np.int64), logprobs=-rng.random(b).astype(np.float32),
                  rewards=rng.standard_normal(b).astype(np.float32), dones=(rng.random(b) < 0.05).astype(np.float32),
                  values=rng.standard_normal(b).astype(np.float32))
    exp.obs.copy_(torch.as_tensor(obs, device=dev))
    ora.obs[:] = obs
    for k, v in fields.items():
        getattr(exp, k).copy_(torch.as_tensor(v, device=dev))
        getattr(ora, k)[:] = v
    exp.num_envs = n
    ora.sort_keys = [(e, t) for t in range(h) for e in range(n)]
    idxs = ora.sort_training_data()
    assert np.array_equal(np.asarray(exp.sort_training_data()), idxs)
    adv_ref = ogae.compute_gae(ora.dones[idxs], ora.values[idxs], ora.rewards[idxs], 0.99, 0.95)
    adv = cpu(exp.compute_gae(0.99, 0.95))
    assert np.max(np.abs(adv - adv_ref) / np.maximum(1.0, np.abs(adv_ref))) <= 1e-5
    # feed the oracle's advantages so the remaining comparisons are exact byte movement
    exp.advantages.copy_(torch.as_tensor(adv_ref, device=dev))
    exp.flatten_batch()
    ora.flatten_batch(adv_ref)
    assert np.array_equal(cpu(exp.b_obs), ora.b_obs)
    for k in ('b_actions', 'b_logprobs', 'b_dones', 'b_values', 'b_advantages', 'b_returns'):
        assert np.array_equal(cpu(getattr(exp, k)), getattr(ora, k)), k
    assert np.array_equal(cpu(exp.returns), ora.returns_np)
    norm = cpu(exp.normalize_advantages())
    for m in range(exp.num_minibatches):
        t_ref = torch.as_tensor(ora.b_advantages[m])
        t_ref = ((t_ref - t_ref.mean()) / (t_ref.std() + 1e-8)).numpy()      # clean_pufferl.py:213 on CPU fp32
        assert np.allclose(norm[m], t_ref, rtol=1e-5, atol=2e-5)
    # zero-copy minibatch form: same row sets as the reference minibatches, slab-major order inside a minibatch
    nm, s_per_env = exp.num_minibatches, h // bptt
    ok = exp.flatten_batch_slabs()
    assert ok == (h % bptt == 0 and s_per_env % nm == 0)
    if ok:
        g_ = s_per_env // nm
        sl = exp._slabs
        for name in ('actions', 'logprobs', 'values', 'advantages', 'returns'):
            s_x, b_x = cpu(getattr(sl, name)), cpu(getattr(exp, 'b_' + name))
            for m in range(nm):
                assert np.array_equal(s_x[m].reshape(g_, bptt, n).transpose(2, 0, 1), b_x[m].reshape(n, g_, bptt)), name
        for m in range(nm):
            so = exp.slab_obs(m)
            assert so.data_ptr() == exp.obs.data_ptr() + m * bptt * n * exp.obs_row_bytes     # a view, not a copy
            assert np.array_equal(cpu(so).reshape(g_, bptt, n, -1).transpose(2, 0, 1, 3),
                                  cpu(exp.b_obs[m]).reshape(n, g_, bptt, -1))
        assert np.array_equal(cpu(exp.returns), ora.returns_np)
        norm_s = cpu(exp.normalize_advantages(slabs=True))
        for m in range(nm):
            assert np.allclose(norm_s[m].reshape(g_, bptt, n).transpose(2, 0, 1).reshape(-1), norm[m].reshape(-1),
                               rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('n_mb,mb_size', [(1, 2), (1, 2048), (4, 65536), (2, 1048576), (3, 1000), (128, 16384)])
def test_adv_norm_vs_torch(n_mb, mb_size):
    dev = torch.device('cuda')
    g = torch.Generator(device='cpu').manual_seed(n_mb * 7 + mb_size)
    a = (torch.randn(n_mb, mb_size, generator=g) * 3 + 0.5).to(dev)
    out = torch.empty_like(a)
    lib = _native.lib()
    ws = torch.zeros(max(16, lib.pb_adv_norm_workspace_bytes(n_mb, mb_size)), dtype=torch.uint8, device=dev)
    _native.check(lib.pb_adv_norm(_native.ptr(a), _native.ptr(out), n_mb, mb_size, _native.ptr(ws), ws.numel(),
                                  _native.stream_ptr()))
    ref = torch.stack([(x - x.mean()) / (x.std() + 1e-8) for x in a.double()]).float()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    # idempotence-style property at any size: the output has mean 0 and unbiased std 1
    assert torch.allclose(out.double().mean(1), torch.zeros(n_mb, device=dev, dtype=torch.float64), atol=1e-5)
    if mb_size > 2:
        assert torch.allclose(out.double().std(1), torch.ones(n_mb, device=dev, dtype=torch.float64), atol=1e-4)
    # in-place
    _native.check(lib.pb_adv_norm(_native.ptr(a), _native.ptr(a), n_mb, mb_size, _native.ptr(ws), ws.numel(),
                                  _native.stream_ptr()))
    assert torch.equal(a, out)


def test_copy_rows_and_store():
    dev = torch.device('cuda')
    lib = _native.lib()
    for row_bytes, n_rows in [(196, 64), (512, 1000), (7, 33), (28224, 16), (16, 1)]:
        src = torch.randint(0, 256, (n_rows, row_bytes + 16), dtype=torch.uint8, device=dev)
        dst = torch.zeros(n_rows, row_bytes + 32, dtype=torch.uint8, device=dev)
        _native.check(lib.pb_copy_rows(_native.ptr(src), row_bytes + 16, _native.ptr(dst), row_bytes + 32, row_bytes,
                                       n_rows, _native.stream_ptr()))
        assert torch.equal(dst[:, :row_bytes], src[:, :row_bytes]) and int(dst[:, row_bytes:].sum()) == 0
    n = 1000
    v, lp = torch.randn(n, device=dev), torch.randn(n, device=dev)
    a = torch.randint(0, 8, (n,), device=dev)
    vr, lr, ar = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, dtype=torch.int64, device=dev)
    _native.check(lib.pb_rollout_store(_native.ptr(v), _native.ptr(lp), _native.ptr(a), _native.ptr(vr), _native.ptr(lr),
                                       _native.ptr(ar), n, _native.stream_ptr()))
    assert torch.equal(v, vr) and torch.equal(lp, lr) and torch.equal(a, ar)


def make_config(n, h, **kw):
    import pufferlib_b200
    cfg = dict(seed=1, torch_deterministic=True, env='squared', batch_size=n * h, bptt_horizon=8, minibatch_size=n * h // 2,
               cpu_offload=False, device='cuda', compile=False, learning_rate=2.5e-4, gamma=0.99, gae_lambda=0.95,
               update_epochs=2, norm_adv=True, clip_coef=0.1, clip_vloss=True, vf_clip_coef=0.1, vf_coef=0.5,
               ent_coef=0.01, max_grad_norm=0.5, target_kl=None, anneal_lr=False, total_timesteps=10 ** 9)
    cfg.update(kw)
    return pufferlib_b200.namespace(**cfg)


def test_evaluate_train_loop_eager_graph_and_host_modes():
    """create / evaluate / train with the reference signatures: eager device rollout, CUDA-graph rollout and the
    host-buffer (numpy in/out) mode all produce finite losses, and the rollout they store replays bit-exactly
    through the oracle (same actions -> same obs / rewards / dones)."""
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    from oracle.squared import SquaredSerial
    n, h = 64, 32
    for mode in ('eager', 'graph', 'host', 'host_graph'):
        # (per-step info dicts need the host after every env step: the captured host loop runs with the device-side statistics)
        backend = pvec.B200.options(host_buffers=mode.startswith('host'), exact_infos=(False if mode == 'host_graph' else None))
        vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=backend)
        torch.manual_seed(0)
        pol = cleanrl.Policy(models.Default(vec.driver_env), fused_sample=(mode != 'host'), seed=1).cuda()
        data = clean_pufferl.create(make_config(n, h, cuda_graph=mode.endswith('graph')), vec, pol)
        ora = SquaredSerial(n)
        ora.async_reset(1)
        for it in range(3):                      # iteration 2 is the first graph replay
            stats, infos = clean_pufferl.evaluate(data)
            exp = data.experience
            acts = cpu(exp.actions).reshape(h, n)
            obs, rew, don = cpu(exp.obs).reshape(h, n, 7, 7), cpu(exp.rewards).reshape(h, n), cpu(exp.dones).reshape(h, n)
            for t in range(h):
                o, r, d, _, _, _, _ = ora.recv()
                assert np.array_equal(o, obs[t]) and np.array_equal(r, rew[t]) and np.array_equal(d, don[t] > 0), (mode, it, t)
                ora.send(acts[t])
            clean_pufferl.train(data)
            assert np.isfinite(data.losses.policy_loss) and np.isfinite(data.losses.value_loss)
            assert data.global_step == (it + 1) * n * h
            assert 'episode_return' in stats
            if mode.startswith('host'):
                # the pinned host arrays hold the step that closed the rollout, the action array the last actions sent
                # (host_graph: every env step of the captured loop copies through them, no host code in between)
                hobs = vec.host_sync()[0]
                assert hobs.shape == (n, 7, 7) and np.isfinite(hobs).all()
                assert np.array_equal(vec._host_np.actions, acts[-1])
                assert vec.d2h_bytes >= (it + 1) * h * n * 49 and vec.h2d_bytes == (it + 1) * h * n * 8
        if mode.endswith('graph'):
            assert data.graph_replays == 2 and data.graph_launches > 0
        clean_pufferl.close(data)


def test_lstm_policy_path():
    """Recurrent policies (LSTMWrapper + RecurrentPolicy): state carried across env steps in evaluate and across
    bptt segments in train (clean_pufferl.py:100-105, 188-191)."""
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    n, h = 32, 16
    vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=pvec.B200)
    torch.manual_seed(0)
    net = models.LSTMWrapper(vec.driver_env, models.Default(vec.driver_env), input_size=128, hidden_size=128)
    pol = cleanrl.RecurrentPolicy(net).cuda()
    data = clean_pufferl.create(make_config(n, h), vec, pol)
    assert data.experience.lstm_h.shape == (1, n, 128)
    for _ in range(2):
        clean_pufferl.evaluate(data)
        assert float(data.experience.lstm_h.abs().sum()) > 0
        clean_pufferl.train(data)
        assert np.isfinite(data.losses.policy_loss) and np.isfinite(data.losses.entropy)
    clean_pufferl.close(data)


def test_graphed_training_matches_eager_training():
    """CUDA-graph rollout + CUDA-graph train update are the same computation as the eager loop.  Same config in both runs
    (so the same capturable Adam), graphs on vs off.  The comparison is made where chaos has not set in yet: the GAE
    look-back composes tile aggregates in a timing-dependent order (1e-7 differences run to run), and action sampling
    amplifies any parameter difference over iterations, so: the first graph-replayed rollout (iteration 2) samples the
    same actions as the eager one, and after the first graphed update the parameters agree to 2e-5."""
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    n, h = 64, 32
    params, rollouts = {}, {}
    for mode in ('eager', 'graph'):
        vec = pvec.make(ocean.env_creator('breakout'), num_envs=n, backend=pvec.B200)
        torch.manual_seed(0)
        pol = cleanrl.Policy(models.Default(vec.driver_env), fused_sample=True, seed=7).cuda()
        g = mode == 'graph'
        data = clean_pufferl.create(make_config(n, h, env='breakout', cuda_graph=True, cuda_graph_rollout=g,
                                                cuda_graph_train=g, anneal_lr=True, total_timesteps=20 * n * h), vec, pol)
        rollouts[mode], params[mode] = [], []
        for it in range(3):
            clean_pufferl.evaluate(data)
            rollouts[mode].append(cpu(data.experience.actions).copy())
            clean_pufferl.train(data)
            params[mode].append([p.detach().cpu().clone() for p in pol.parameters()])
        if g:
            assert data.train_graph_state == 2 and data.train_graph_replays == 2 and data.graph_replays == 2, data.msg
        else:
            assert data.train_graph_state != 2 and data.graph_replays == 0
        clean_pufferl.close(data)
    agree = [float((a == b).mean()) for a, b in zip(rollouts['eager'], rollouts['graph'])]
    diffs = [max(float((a - b).abs().max()) for a, b in zip(pa, pb)) for pa, pb in zip(params['eager'], params['graph'])]
    assert agree[0] == 1.0 and agree[1] > 0.9995 and agree[2] > 0.98, (agree, diffs)
    assert diffs[0] <= 2e-6 and diffs[1] <= 2e-5, (agree, diffs)


def test_zero_copy_minibatches_match_gathered_minibatches():
    """train() on zero-copy slab minibatches (Experience.flatten_batch_slabs; observations never gathered) is the same
    update as train() on the gathered, sorted minibatches of the reference layout: same rollout, parameters after the
    first update agree to 2e-5 (row order inside a minibatch only changes the summation order)."""
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    n, h = 64, 32
    params, acts, used = {}, {}, {}
    for zc in (True, False):
        vec = pvec.make(ocean.env_creator('breakout'), num_envs=n, backend=pvec.B200)
        torch.manual_seed(0)
        pol = cleanrl.Policy(models.Default(vec.driver_env), fused_sample=True, seed=7).cuda()
        data = clean_pufferl.create(make_config(n, h, env='breakout', zero_copy_minibatches=zc), vec, pol)
        clean_pufferl.evaluate(data)
        acts[zc] = cpu(data.experience.actions).copy()
        clean_pufferl.train(data)
        params[zc] = [p.detach().cpu().clone() for p in pol.parameters()]
        used[zc] = data.experience._slabs is not None
        assert np.isfinite(data.losses.policy_loss) and np.isfinite(data.losses.explained_variance)
        clean_pufferl.close(data)
    assert used[True] and not used[False]
    assert np.array_equal(acts[True], acts[False])
    diff = max(float((a - b).abs().max()) for a, b in zip(params[True], params[False]))
    assert diff <= 2e-5, diff


def test_lstm_parity_vs_reference(golden):
    """The recurrent path against the REFERENCE's own run (tests/golden/lstm_squared.npz: unmodified
    clean_pufferl.create/evaluate/train + models.LSTMWrapper + cleanrl.RecurrentPolicy on CPU, generate.py::gen_lstm).
    Same initial weights, same seed, the reference's sampled actions replayed: evaluate must store the same
    observations / values / logprobs and leave the same LSTM state (lstm_h[:, env_id] carry, clean_pufferl.py:100-105);
    train must give the same losses and parameters (bptt segments [rows, bptt, *obs], state carried across the
    minibatches of an epoch and reset per epoch, :176-191; models.py:64-111)."""
    import pufferlib_b200
    from pufferlib_b200 import models
    from pufferlib_b200.frameworks import cleanrl
    g = golden('lstm_squared')
    n, h, bptt, mbs, hid = (int(g[k]) for k in ('num_envs', 'horizon', 'bptt', 'minibatch_size', 'hidden'))
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False     # the golden is CPU fp32
    try:
        vec = pvec.make(ocean.env_creator('squared'), num_envs=n, backend=pvec.B200)
        net = models.LSTMWrapper(vec.driver_env, models.Default(vec.driver_env, hidden_size=hid), input_size=hid,
                                 hidden_size=hid)
        net.policy.fast_path = False

        class TapePolicy(cleanrl.RecurrentPolicy):
            """RecurrentPolicy that replays the reference's sampled actions in evaluate (sampling itself is
            torch.multinomial on the CPU generator there; the given-action branch of sample_logits is the same code)."""
            tape, t = torch.as_tensor(g['actions'].reshape(h, n), device='cuda'), 0

            def forward(self, x, state=None, action=None):
                if action is None:
                    action = self.tape[self.t]
                    self.t += 1
                return super().forward(x, state, action)

        pol = TapePolicy(net).cuda()
        sd = {k[len('init/'):]: torch.as_tensor(g[k]) for k in g.files if k.startswith('init/')}
        assert set(sd) == set(pol.state_dict()), 'module / parameter names follow the reference'
        pol.load_state_dict(sd)
        cfg = pufferlib_b200.namespace(
            seed=int(g['seed']), torch_deterministic=True, env='squared', batch_size=n * h, bptt_horizon=bptt,
            minibatch_size=mbs, cpu_offload=False, device='cuda', compile=False, learning_rate=float(g['learning_rate']),
            gamma=0.99, gae_lambda=0.95, update_epochs=int(g['update_epochs']), norm_adv=True, clip_coef=0.1,
            clip_vloss=True, vf_clip_coef=0.1, vf_coef=0.5, ent_coef=0.01, max_grad_norm=0.5, target_kl=None,
            anneal_lr=False, total_timesteps=10 ** 9)
        data = clean_pufferl.create(cfg, vec, pol)
        clean_pufferl.evaluate(data)
        exp = data.experience
        assert np.array_equal(cpu(exp.obs), g['obs_i8'].astype(np.float32))
        assert np.array_equal(cpu(exp.actions), g['actions'])
        assert np.array_equal(cpu(exp.rewards), g['rewards']) and np.array_equal(cpu(exp.dones), g['dones'])
        assert np.allclose(cpu(exp.values), g['values'], rtol=1e-4, atol=2e-6)
        assert np.allclose(cpu(exp.logprobs), g['logprobs'], rtol=1e-4, atol=2e-6)
        assert np.allclose(cpu(exp.lstm_h), g['lstm_h'], rtol=1e-4, atol=2e-6)
        assert np.allclose(cpu(exp.lstm_c), g['lstm_c'], rtol=1e-4, atol=2e-6)
        clean_pufferl.train(data)
        assert np.array_equal(cpu(exp.b_obs), g['b_obs_i8'].astype(np.float32))
        assert np.allclose(cpu(exp.b_advantages), g['advantages'], rtol=1e-4, atol=1e-5)
        for k in ('policy_loss', 'value_loss', 'entropy', 'old_approx_kl', 'approx_kl', 'clipfrac'):
            assert np.isclose(getattr(data.losses, k), float(g['loss_' + k]), rtol=2e-3, atol=2e-5), \
                (k, getattr(data.losses, k), float(g['loss_' + k]))
        assert np.isclose(data.losses.explained_variance, float(g['loss_explained_variance']), rtol=1e-3, atol=1e-4)
        # parameters after update_epochs x num_minibatches Adam steps (lr 2.5e-3): Adam normalises the gradient, so fp32
        # noise on near-zero gradients moves single elements by a fraction of lr -- the bulk must agree tightly
        diffs = []
        for k, v in pol.state_dict().items():
            d = np.abs(cpu(v) - g['after/' + k])
            assert d.max() < 1e-3, (k, d.max())
            diffs.append(d.ravel())
        diffs = np.concatenate(diffs)
        assert np.mean(diffs) < 2e-5 and np.quantile(diffs, 0.99) < 2e-4
        moved = np.concatenate([np.abs(g['after/' + k] - g['init/' + k]).ravel() for k in sd])
        assert np.mean(moved) > 50 * np.mean(diffs), 'the update itself is much larger than the disagreement'
        clean_pufferl.close(data)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
