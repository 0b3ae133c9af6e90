"""Micro-benchmark: ways to compute dW_enc = dPre^T @ x  ([128 x M] @ [M x 128], M = 524288, TF32) with library GEMMs."""
import torch
torch.set_float32_matmul_precision('high')
M, H = 524288, 128
x = torch.randn(M, H, device='cuda'); d = torch.randn(M, H, device='cuda')


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ref = d.t() @ x          # TF32 library result as the common yardstick (fp64 GEMMs of this shape are very slow here)
variants = {
    'd.t() @ x': lambda: d.t() @ x,
    '(x.t() @ d).t()': lambda: (x.t() @ d).t(),
    'einsum mi,mj->ij': lambda: torch.einsum('mi,mj->ij', d, x),
}
for S in (16, 64, 128, 256, 512, 1024):
    variants[f'bmm split-K S={S}'] = (lambda S=S: torch.bmm(d.view(S, M // S, H).transpose(1, 2), x.view(S, M // S, H)).sum(0))
for name, fn in variants.items():
    out = fn()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f'{name:28s} {timeit(fn):8.1f} us   rel.diff vs d.t()@x {err:.2e}', flush=True)
