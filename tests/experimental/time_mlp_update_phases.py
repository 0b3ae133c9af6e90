"""Phase timing of the fused update kernel: PB_MUF_SKIP masks remove parts of the epilogue (results are wrong then; only
the time matters).  Run on a B200:  python tests/experimental/time_mlp_update_phases.py"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import check_mlp_update_fused as c
dev = torch.device('cuda'); m, n_act = 524288, 4
torch.manual_seed(0)
xbuf = torch.randn(4 * m, 128, device=dev)
w_enc = torch.randn(128, 128, device=dev) * 0.1; b_enc = torch.randn(128, device=dev) * 0.1
w_cat = torch.zeros(8, 128, device=dev); w_cat[:5] = torch.randn(5, 128, device=dev) * 0.1; b_cat = torch.zeros(8, device=dev)
act = torch.randint(0, n_act, (m,), device=dev); olp = -torch.rand(m, device=dev) - 0.5
adv, ret, oval = torch.randn(m, device=dev), torch.randn(m, device=dev), torch.randn(m, device=dev)
dpre = torch.empty(m, 128, device=dev)
for name, dpo in (('kernel', None), ('hbm', dpre)):
    fn = lambda k: c.fused(xbuf[(k %% 4) * (m // 2):], 128, m // 2, 2 * m, 2, w_enc, b_enc, w_cat, b_cat, act, olp, adv, ret, oval, n_act, False, dpre_out=dpo)
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(16): fn(k)
    e1.record(); torch.cuda.synchronize()
    print('  dW %%-6s %%7.1f us' %% (name, e0.elapsed_time(e1) * 1000 / 16), flush=True)
''' % HERE
NAMES = {1: 'head FFMAs', 2: 'loss', 4: 'dPre FFMAs', 8: 'mma.sync+staging', 16: 'dPre stores', 32: 'column sums', 64: 'exchange'}
for mask in (0, 1, 2, 4, 8, 16, 32, 64, 1 | 4, 1 | 2 | 4 | 8 | 32, 127):
    label = ' + '.join(v for k, v in NAMES.items() if mask & k) or 'nothing'
    print(f'skip {mask:3d} ({label})', flush=True)
    env = dict(os.environ, PB_MUF_SKIP=str(mask))
    subprocess.run([sys.executable, '-c', CODE], env=env, check=False)
