"""Hardware check of the EXPERIMENTAL fused heads + PPO loss + tail-backward kernel
(csrc/experimental/heads_loss_tail.cu) against the validated chain it replaces:
    out = hidden @ W_heads^T + b  ->  pb_ppo_loss  ->  pb_mlp_tail_backward.
Not collected by pytest.  Run by hand on a B200 box under a timeout:

    gpurun --timeout 300 -- 'python -m pufferlib_b200.build --experimental && timeout 150 python tests/experimental/check_heads_loss_tail.py'
"""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from pufferlib_b200 import _native  # noqa: E402

exp = C.CDLL(os.path.join(REPO, 'pufferlib_b200', 'libpuffer_b200_exp.so'))
exp.pbx_heads_loss_tail.restype = C.c_int
exp.pbx_heads_loss_tail.argtypes = [C.c_void_p] * 8 + [C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_float,
                                                       C.c_float] + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
exp.pbx_heads_loss_tail_workspace_bytes.restype = C.c_size_t
exp.pbx_heads_loss_tail_workspace_bytes.argtypes = [C.c_int64]
exp.pbx_hlt_last_error.restype = C.c_char_p


def ptr(t):
    return C.c_void_p(t.data_ptr())


def chain(hidden, w_cat, b_cat, act, olp, adv, ret, oval, n_act, cfg):
    """The validated path: cuBLAS head GEMM + pb_ppo_loss + pb_mlp_tail_backward."""
    lib, s = _native.lib(), _native.stream_ptr()
    m = hidden.shape[0]
    out = torch.addmm(b_cat, hidden, w_cat.t())
    dout = torch.empty_like(out)
    stats = torch.empty(8, dtype=torch.float64, device=hidden.device)
    o, d = out.data_ptr(), dout.data_ptr()
    _native.check(lib.pb_ppo_loss(C.c_void_p(o), 8, C.c_void_p(o + 4 * n_act), 8, ptr(act), ptr(olp), ptr(adv), ptr(ret),
                                  ptr(oval), m, n_act, C.c_float(cfg[0]), int(cfg[1]), C.c_float(cfg[2]),
                                  C.c_float(cfg[3]), C.c_float(cfg[4]), C.c_void_p(d), 8, C.c_void_p(d + 4 * n_act), 8,
                                  ptr(stats), s))
    dpre = torch.empty_like(hidden)
    grads = torch.empty(8 * 128 + 128 + 8, device=hidden.device)
    ws = torch.empty(lib.pb_mlp_tail_workspace_bytes(m, 128), dtype=torch.uint8, device=hidden.device)
    _native.check(lib.pb_mlp_tail_backward(ptr(dout), 8, ptr(w_cat), ptr(hidden), m, 128, ptr(dpre), ptr(grads), ptr(ws),
                                           ws.numel(), s))
    return dpre, grads, stats


def fused(hidden, w_cat, b_cat, act, olp, adv, ret, oval, n_act, cfg):
    m = hidden.shape[0]
    dpre = torch.empty_like(hidden)
    grads = torch.empty(8 * 128 + 128 + 8, device=hidden.device)
    stats = torch.empty(8, dtype=torch.float64, device=hidden.device)
    ws = torch.empty(exp.pbx_heads_loss_tail_workspace_bytes(m), dtype=torch.uint8, device=hidden.device)
    rc = exp.pbx_heads_loss_tail(ptr(hidden), ptr(w_cat), ptr(b_cat), ptr(act), ptr(olp), ptr(adv), ptr(ret), ptr(oval), m,
                                 n_act, cfg[0], int(cfg[1]), cfg[2], cfg[3], cfg[4], ptr(dpre), ptr(grads), ptr(stats),
                                 ptr(ws), ws.numel(), _native.stream_ptr())
    assert rc == 0, exp.pbx_hlt_last_error().decode()
    return dpre, grads, stats


def main():
    dev = torch.device('cuda')
    torch.backends.cuda.matmul.allow_tf32 = True
    cfg = (0.1, True, 0.1, 0.5, 0.01)
    for m, n_act in ((64, 4), (1, 4), (1000, 7), (512 * 3 + 17, 1), (524288, 4)):
        torch.manual_seed(m)
        hidden = torch.relu(torch.randn(m, 128, device=dev))
        w_cat = torch.zeros(8, 128, device=dev)
        w_cat[:n_act + 1] = torch.randn(n_act + 1, 128, device=dev) * 0.1
        b_cat = torch.zeros(8, device=dev)
        b_cat[:n_act + 1] = torch.randn(n_act + 1, device=dev) * 0.1
        act = torch.randint(0, n_act, (m,), device=dev)
        olp = -torch.rand(m, device=dev) - 0.5
        adv, ret, oval = torch.randn(m, device=dev), torch.randn(m, device=dev), torch.randn(m, device=dev)
        args = (hidden, w_cat, b_cat, act, olp, adv, ret, oval, n_act, cfg)
        a, b = chain(*args), fused(*args)
        torch.cuda.synchronize()
        names = ('dpre', 'grads', 'stats')
        for name, x, y in zip(names, a, b):
            scale = float(x.abs().max()) + 1e-30
            err = float((x.double() - y.double()).abs().max()) / scale
            print(f'm={m:7d} n_act={n_act} {name:6s} max rel err {err:.3e}')
            assert err < 5e-3, (name, err)         # TF32 head products on both sides, different rounding of hidden
    # timing at the bench minibatch
    for name, fn in (('chain (GEMM + pb_ppo_loss + pb_mlp_tail_backward)', chain), ('fused', fused)):
        for _ in range(3):
            fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(*args)
        e1.record()
        torch.cuda.synchronize()
        print(f'{name:52s} {e0.elapsed_time(e1) * 100:8.1f} us')
    print('ok')


if __name__ == '__main__':
    main()
