"""Hardware check of the EXPERIMENTAL tcgen05 encoder GEMM (csrc/experimental/enc_gemm_tcgen05.cu).  Not collected
by pytest (no test_ prefix): run it by hand on a B200 box under a timeout, e.g.

    gpurun --timeout 300 -- 'python -m pufferlib_b200.build --experimental && timeout 120 python tests/experimental/check_enc_gemm_tcgen05.py'

Checks hidden = relu(x @ W^T + b) against torch (TF32 tolerance), row counts that are / are not multiples of the
128-row tile, a strided x (rollout slab view), then times the bench shape (M = 524288) against cuBLAS."""
import ctypes as C
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
lib = C.CDLL(os.path.join(REPO, 'pufferlib_b200', 'libpuffer_b200_exp.so'))
lib.pbx_enc_gemm_tf32.restype = C.c_int
lib.pbx_enc_gemm_tf32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.pbx_last_error.restype = C.c_char_p


def run(x, w, b, out):
    rc = lib.pbx_enc_gemm_tf32(x.data_ptr(), x.stride(0), x.shape[0], w.data_ptr(), b.data_ptr(), out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.pbx_last_error().decode()


def main():
    dev = torch.device('cuda')
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    w = torch.randn(128, 128, device=dev) * 0.1
    b = torch.randn(128, device=dev) * 0.1
    for m in (128, 1, 1000, 4096 + 37, 148 * 128 * 3):
        x = torch.randn(m, 128, device=dev)
        out = torch.full((m, 128), -7.0, device=dev)
        run(x, w, b, out)
        torch.cuda.synchronize()
        ref = torch.relu(x @ w.t() + b)
        err = float((out - ref).abs().max())
        print(f'm={m:8d} max abs err {err:.3e}')
        assert err < 1e-2, err        # TF32: both operands truncated to 10 mantissa bits, K = 128 products
    # strided rows (a [R, 128] window of a wider buffer)
    big = torch.randn(4096, 256, device=dev)
    x = big[:, 64:192]
    out = torch.empty(4096, 128, device=dev)
    run(x, w, b, out)
    torch.cuda.synchronize()
    assert float((out - torch.relu(x @ w.t() + b)).abs().max()) < 5e-3
    # timing at the bench minibatch
    m = 524288
    x = torch.randn(m, 128, device=dev)
    out = torch.empty(m, 128, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    for name, fn in (('tcgen05 (ours)', lambda: run(x, w, b, out)),
                     ('cuBLASLt addmm+relu epilogue', lambda: torch._addmm_activation(b, x, w.t(), use_gelu=False, out=out))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        print(f'{name:32s} {us:8.1f} us   {2 * m * 128 * 4 / us / 1e3:8.1f} GB/s (x read + hidden write)')
    print('ok')


if __name__ == '__main__':
    main()
