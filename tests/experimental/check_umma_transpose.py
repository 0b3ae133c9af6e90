"""Hardware probe (csrc/experimental/umma_probe.cu, pbx_umma_probe2) of the two tensor-core steps the single-x-layout
fused update kernel needs:

  stage 1  x^T[feat][row] = I . x^T as an SS MMA: A = an 8 KB no-swizzle K-major "sliding identity" image (31 8-row groups
           per 4-column strip, only group 15 non-zero; MMA k starts (15 - k) groups into it, so row group k sees the 8x8
           identity and every other row group sees zeros), B = the K-major SWIZZLE_128B x tile the forward product reads
  stage 2  dW^T[feat][hid] = x^T . dPre as a TS MMA: A = stage 1's accumulator straight from tensor memory (lanes = feat,
           columns = rows), B = dPre as K-major SWIZZLE_128B blocks [128 hid][32 rows]

    python -m pufferlib_b200.build --experimental && python tests/experimental/check_umma_transpose.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_umma_probe_lib import image_kmajor, desc, idesc, tf32_trunc      # noqa: E402

lib = C.CDLL(os.path.join(REPO, 'pufferlib_b200', 'libpuffer_b200_exp.so'))
lib.pbx_umma_probe2.restype = C.c_int
lib.pbx_umma_probe2.argtypes = ([C.c_void_p] * 3 + [C.c_uint32] * 3 + [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int] +
                                [C.c_uint32] * 5 + [C.c_uint64, C.c_uint32, C.c_int, C.c_int] + [C.c_uint32] * 4 +
                                [C.c_void_p, C.c_void_p])

GROUP = 128                      # 8 rows x 16 B core matrix
STRIP = 32 * GROUP               # one 4-column strip of the sliding identity (31 groups used)


def sliding_identity():
    img = np.zeros(2 * STRIP // 4, dtype=np.float32)
    for r in range(8):           # core matrix of strip r // 4: row r holds the 1 at column r % 4
        img[((r // 4) * STRIP + 15 * GROUP + r * 16 + (r % 4) * 4) // 4] = 1.0
    return img


def main():
    rng = np.random.default_rng(0)
    x = tf32_trunc(rng.standard_normal((128, 128)).astype(np.float32))       # [row][feat]
    dp = rng.standard_normal((128, 128)).astype(np.float32)                  # [row][hid]
    a_img, b_img, c_img = sliding_identity(), image_kmajor(x, 16384), image_kmajor(dp.T.copy(), 16384)
    dev = torch.device('cuda')
    ta, tb, tc = (torch.from_numpy(i).to(dev) for i in (a_img, b_img, c_img))
    out = torch.full((128, 256), -777.0, device=dev)
    neg = lambda v: (-v) & 0xFFFFFFFF
    rc = lib.pbx_umma_probe2(
        ta.data_ptr(), tb.data_ptr(), tc.data_ptr(), ta.numel() * 4, tb.numel() * 4, tc.numel() * 4,
        desc(STRIP, GROUP, 0), desc(16, 1024, 2), idesc(128, 128, 0, 0), 16, 4, neg(GROUP), 32, neg(4 * GROUP), 16384, 15 * GROUP,
        desc(16, 1024, 2), idesc(128, 128, 0, 0), 16, 4, 32, 16384, 0, 8, out.data_ptr(), None)
    assert rc == 0, rc
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    e1 = np.abs(got[:, :128] - x.T).max()
    ref = x.astype(np.float64).T @ tf32_trunc(dp).astype(np.float64)
    e2 = np.abs(got[:, 128:] - ref).max() / np.abs(ref).max()
    print(f'stage 1 (x^T by the sliding identity): max abs err {e1:.3e}  {"OK" if e1 == 0 else "MISMATCH"}')
    print(f'stage 2 (TS MMA, A = stage 1 accumulator): max rel err {e2:.3e}  {"OK" if e2 < 1e-5 else "MISMATCH"}')
    if e1 != 0:
        print('stage 1 sample rows:\n', got[:3, :8], '\nexpected\n', x.T[:3, :8])
        nz = np.count_nonzero(got[:, :128])
        print('nonzeros', nz, 'of', 128 * 128)


if __name__ == '__main__':
    main()
