"""Hardware probe of the tcgen05 descriptor encodings the fused update kernel relies on (csrc/experimental/
umma_probe.cu).  Builds shared-memory images of both operands in numpy for each layout hypothesis and compares the raw
TMEM accumulator with a float64 matmul of TF32-truncated operands.  Not collected by pytest; run on a B200 box:

    gpurun --timeout 300 -- 'python -m pufferlib_b200.build --experimental && timeout 120 python tests/experimental/check_umma_probe.py'
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
lib = C.CDLL(os.path.join(REPO, 'pufferlib_b200', 'libpuffer_b200_exp.so'))
lib.pbx_umma_probe.restype = C.c_int
lib.pbx_umma_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int,
                               C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]


def swz128(off):
    return off ^ (((off >> 7) & 7) << 4)


def image_kmajor(mat, kblk_stride):
    """mat [MN][K] fp32 -> SWIZZLE_128B K-major image: K-block kb (32 floats) = [MN rows][128 B], 8-row atoms of 1 KiB."""
    mn, k = mat.shape
    img = np.zeros(max(kblk_stride * (k // 32), mn * 128) // 4, dtype=np.float32)
    r, c = np.meshgrid(np.arange(mn), np.arange(k), indexing='ij')
    off = (c // 32) * kblk_stride + r * 128 + (c % 32) * 4
    img[swz128(off) // 4] = mat
    return img


def image_mnmajor(mat, lbo, sbo):
    """mat [MN][K] fp32 -> SWIZZLE_128B MN-major image: 32 MN elements contiguous (128 B), 8 k's per 1 KiB atom (stride
    128 B), k-groups at SBO, MN-groups (of 32) at LBO."""
    mn, k = mat.shape
    size = ((mn - 1) // 32) * lbo + ((k - 1) // 8) * sbo + 1024
    img = np.zeros(size // 4, dtype=np.float32)
    r, c = np.meshgrid(np.arange(mn), np.arange(k), indexing='ij')
    off = (r % 32) * 4 + (r // 32) * lbo + (c % 8) * 128 + (c // 8) * sbo
    img[swz128(off) // 4] = mat
    return img


def desc(lbo, sbo, layout=2):
    return (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def swz128_32(off):
    """SWIZZLE_128B_BASE32B (UMMA layout type 1; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32-byte chunks within a 128-byte
    row, chunk index ^= row index & 3  (byte-address bits [5,7) ^= bits [7,9)) -- the only MN-major layout for 32-bit types."""
    return off ^ (((off >> 7) & 3) << 5)


def image_mnmajor32(mat, lbo, sbo, swz=swz128_32):
    """mat [MN][K] fp32 -> MN-major image with 4-k atoms: 32 MN elements contiguous (128 B), 4 k's per 512-byte atom (stride
    128 B), k-groups (of 4) at SBO, MN-groups (of 32) at LBO."""
    mn, k = mat.shape
    size = ((mn - 1) // 32) * lbo + ((k - 1) // 4) * sbo + 512
    img = np.zeros(size // 4, dtype=np.float32)
    r, c = np.meshgrid(np.arange(mn), np.arange(k), indexing='ij')
    off = (r % 32) * 4 + (r // 32) * lbo + (c % 4) * 128 + (c // 4) * sbo
    img[swz(off) // 4] = mat
    return img


def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def run(a_img, b_img, a_desc, b_desc, idesc_, n_mma, inner, steps, ncols):
    dev = torch.device('cuda')
    a = torch.from_numpy(a_img).to(dev)
    b = torch.from_numpy(b_img).to(dev)
    out = torch.full((128, ncols), -777.0, device=dev)
    rc = lib.pbx_umma_probe(a.data_ptr(), b.data_ptr(), a.numel() * 4, b.numel() * 4, a_desc, b_desc, idesc_, n_mma, inner,
                            steps[0], steps[1], steps[2], steps[3], out.data_ptr(), ncols,
                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out.cpu().numpy()


def report(name, got, ref):
    err = float(np.abs(got - ref).max()) / (float(np.abs(ref).max()) + 1e-30)
    print(f'{name:70s} max rel err {err:.3e}  {"OK" if err < 2e-3 else "MISMATCH"}', flush=True)
    return err < 2e-3


def main():
    rng = np.random.default_rng(0)
    ok = {}
    # ---- 1. K-major x K-major, K = 32 (one swizzle row): 4 MMAs stepping 32 B inside the row
    a = rng.standard_normal((128, 32)).astype(np.float32)
    b = rng.standard_normal((128, 32)).astype(np.float32)
    ref = tf32_trunc(a).astype(np.float64) @ tf32_trunc(b).astype(np.float64).T
    got = run(image_kmajor(a, 16384), image_kmajor(b, 16384), desc(16, 1024), desc(16, 1024), idesc(128, 128, 0, 0), 4, 4,
              (32, 32, 0, 0), 128)
    ok['k32'] = report('K-major x K-major  M=N=128 K=32  (4 MMAs, +32 B)', got, ref)
    # ---- 2. K = 128: 4 K-blocks of 16 KiB
    a = rng.standard_normal((128, 128)).astype(np.float32)
    b = rng.standard_normal((128, 128)).astype(np.float32)
    ref = tf32_trunc(a).astype(np.float64) @ tf32_trunc(b).astype(np.float64).T
    got = run(image_kmajor(a, 16384), image_kmajor(b, 16384), desc(16, 1024), desc(16, 1024), idesc(128, 128, 0, 0), 16, 4,
              (32, 32, 16384, 16384), 128)
    ok['k128'] = report('K-major x K-major  M=N=128 K=128 (16 MMAs, 4 K-blocks)', got, ref)
    # ---- 3. MN-major x MN-major: D[f][j] = sum_r x[r][f] * dp[r][j]; the x image is the SAME bytes as a K-major tile
    x = rng.standard_normal((128, 128)).astype(np.float32)          # [row][feature]
    dp = rng.standard_normal((128, 128)).astype(np.float32)         # [row][hidden]
    ref = tf32_trunc(x).astype(np.float64).T @ tf32_trunc(dp).astype(np.float64)
    x_img = image_kmajor(x, 16384)                                  # what TMA writes for the forward GEMM
    assert np.array_equal(x_img, image_mnmajor(x.T.copy(), 16384, 1024))
    dp_img = image_mnmajor(dp.T.copy(), 16384, 1024)
    for name, (lbo, sbo) in (('LBO=16K SBO=1K', (16384, 1024)), ('LBO=1K SBO=16K (swapped)', (1024, 16384))):
        got = run(x_img, dp_img, desc(lbo, sbo), desc(lbo, sbo), idesc(128, 128, 1, 1), 16, 16, (1024, 1024, 0, 0), 128)
        ok['mn ' + name] = report(f'MN-major x MN-major M=N=128 K=128 rows, {name}', got, ref)
    # ---- 4. MN-major B chunk with N = 32 (the dPre chunk buffer: [128 rows][32 floats], no LBO use)
    for n in (32, 64):
        dpc = dp[:, :n].copy()
        ref = tf32_trunc(x).astype(np.float64).T @ tf32_trunc(dpc).astype(np.float64)
        got = run(x_img, image_mnmajor(dpc.T.copy(), 16384, 1024), desc(16384, 1024), desc(16384, 1024),
                  idesc(128, n, 1, 1), 16, 16, (1024, 1024, 0, 0), n)
        ok[f'mn chunk {n}'] = report(f'MN-major x MN-major M=128 N={n} K=128 rows (dPre chunk)', got, ref)
    # ---- 5. accumulate flag / partial K: only the first 64 rows (8 MMAs)
    ref = tf32_trunc(x[:64]).astype(np.float64).T @ tf32_trunc(dp[:64]).astype(np.float64)
    got = run(x_img, dp_img, desc(16384, 1024), desc(16384, 1024), idesc(128, 128, 1, 1), 8, 8, (1024, 1024, 0, 0), 128)
    ok['mn half'] = report('MN-major x MN-major, first 64 rows only (8 MMAs)', got, ref)
    # ---- 6. the 32-bit MN-major layout: SWIZZLE_128B_BASE32B (layout type 1), 4-k atoms of 512 B.  x tile = what TMA writes
    #         with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B for a [128 rows][32 floats] box: row r at 128 r, 32-byte chunk c at c ^ (r & 3)
    for name, (sbo, k_step) in (('SBO=512, +1024 B per K=8', (512, 1024)),):
        xi = image_mnmajor32(x.T.copy(), 16384, sbo)
        di = image_mnmajor32(dp.T.copy(), 16384, sbo)
        # what the TMA tile looks like, computed independently: element (row, f) at kb*16384 + row*128 + ((f%32)*4 ^ ((row&3)<<5))
        rr, ff = np.meshgrid(np.arange(128), np.arange(128), indexing='ij')
        tma = np.zeros(16384, dtype=np.float32)
        tma[((ff // 32) * 16384 + rr * 128 + (((ff % 32) * 4) ^ ((rr & 3) << 5))) // 4] = x
        assert np.array_equal(tma, xi), 'BASE32B image == TMA ATOM_32B tile'
        ref = tf32_trunc(x).astype(np.float64).T @ tf32_trunc(dp).astype(np.float64)
        for lay in (1, 2):
            got = run(xi, di, desc(16384, sbo, lay), desc(16384, sbo, lay), idesc(128, 128, 1, 1), 16, 16, (k_step, k_step, 0, 0), 128)
            ok[f'mn32 L{lay} {name}'] = report(f'MN-major BASE32B image, layout type {lay}, {name}', got, ref)
            print('      got stats: absmax %.3e, zeros %.2f' % (np.abs(got).max(), float((got == 0).mean())))
        got = run(xi, di, desc(512, 16384, 1), desc(512, 16384, 1), idesc(128, 128, 1, 1), 16, 16, (k_step, k_step, 0, 0), 128)
        ok['mn32 swapped'] = report('MN-major BASE32B image, layout 1, LBO/SBO swapped', got, ref)
        for n in (32, 64):
            dpc = dp[:, :n].copy()
            refc = tf32_trunc(x).astype(np.float64).T @ tf32_trunc(dpc).astype(np.float64)
            got = run(xi, image_mnmajor32(dpc.T.copy(), 16384, sbo), desc(16384, sbo, 1), desc(16384, sbo, 1),
                      idesc(128, n, 1, 1), 16, 16, (k_step, k_step, 0, 0), n)
            ok[f'mn32 chunk {n}'] = report(f'MN-major BASE32B, dPre chunk N={n}', got, refc)
        # mixed: A MN-major (x, BASE32B) with B K-major is not needed; A K-major (forward) unaffected
    # ---- 7. can ONE x tile serve the forward (K-major) and the dW product (MN-major)?
    # 7a. K-major descriptor with layout type 1 on the BASE32B-swizzled row-major tile (what TMA ATOM_32B writes)
    a = rng.standard_normal((128, 128)).astype(np.float32)       # [row][feature]
    w = rng.standard_normal((128, 128)).astype(np.float32)       # [hidden][feature]
    ref = tf32_trunc(a).astype(np.float64) @ tf32_trunc(w).astype(np.float64).T
    rr, ff = np.meshgrid(np.arange(128), np.arange(128), indexing='ij')
    a32 = np.zeros(16384, dtype=np.float32)
    a32[((ff // 32) * 16384 + rr * 128 + (((ff % 32) * 4) ^ ((rr & 3) << 5))) // 4] = a
    if 'kmajor1' in sys.argv:     # faults with "misaligned address" on sm_100a: K-major operands cannot use layout type 1
        for sbo in (1024, 512):
            got = run(a32, image_kmajor(w, 16384), desc(16, sbo, 1), desc(16, 1024, 2), idesc(128, 128, 0, 0), 16, 4,
                      (32, 32, 16384, 16384), 128)
            ok[f'k-major layout1 sbo{sbo}'] = report(f'K-major A with layout type 1 (BASE32B tile), SBO={sbo}', got, ref)
    # 7b. no-swizzle core-matrix image: strip s = features 4s..4s+3 as [128 rows][16 B]; K-major and MN-major views
    def image_cm(mat):       # mat [row][col] -> strips of 4 columns, [128 rows][16 B] each
        img = np.zeros(mat.shape[0] * mat.shape[1], dtype=np.float32)
        r_, c_ = np.meshgrid(np.arange(mat.shape[0]), np.arange(mat.shape[1]), indexing='ij')
        img[((c_ // 4) * (mat.shape[0] * 16) + r_ * 16 + (c_ % 4) * 4) // 4] = mat
        return img
    a_cm, w_cm = image_cm(a), image_cm(w)
    cases = [x for x in sys.argv[1:] if x.startswith('cm')]
    for name, (lbo, sbo) in (('LBO=2048 SBO=128', (2048, 128)), ('LBO=128 SBO=2048', (128, 2048))):
        if f'cmk{lbo}' not in cases:
            continue
        got = run(a_cm, w_cm, desc(lbo, sbo, 0), desc(lbo, sbo, 0), idesc(128, 128, 0, 0), 16, 16, (4096, 4096, 0, 0), 128)
        ok['cm k-major ' + name] = report(f'no-swizzle core matrices, K-major x K-major, {name}', got, ref)
    dp2 = rng.standard_normal((128, 128)).astype(np.float32)     # [row][hidden]
    ref_dw = tf32_trunc(a).astype(np.float64).T @ tf32_trunc(dp2).astype(np.float64)
    d_cm = image_cm(dp2)
    for name, (lbo, sbo) in (('LBO=128 SBO=2048', (128, 2048)), ('LBO=2048 SBO=128', (2048, 128))):
        if f'cmm{lbo}' not in cases:
            continue
        got = run(a_cm, d_cm, desc(lbo, sbo, 0), desc(lbo, sbo, 0), idesc(128, 128, 1, 1), 16, 16, (128, 128, 0, 0), 128)
        ok['cm mn-major ' + name] = report(f'no-swizzle core matrices, MN-major x MN-major (same x image), {name}', got, ref_dw)
    # mixed: A = x no-swizzle MN-major, B = dPre BASE32B MN-major chunk (N = 32)
    dpc = dp2[:, :32].copy()
    refc = tf32_trunc(a).astype(np.float64).T @ tf32_trunc(dpc).astype(np.float64)
    for name, (lbo, sbo) in (('LBO=128 SBO=2048', (128, 2048)), ('LBO=2048 SBO=128', (2048, 128))):
        if f'cmx{lbo}' not in cases:
            continue
        got = run(a_cm, image_mnmajor32(dpc.T.copy(), 16384, 512), desc(lbo, sbo, 0), desc(16384, 512, 1), idesc(128, 32, 1, 1), 16, 16,
                  (128, 1024, 0, 0), 32)
        ok['cm/32 mixed ' + name] = report(f'A no-swizzle MN-major ({name}), B BASE32B chunk N=32', got, refc)
    # the failing SW128 (16-byte) MN-major cases above: how do they fail?
    got = run(x_img, dp_img, desc(16384, 1024), desc(16384, 1024), idesc(128, 128, 1, 1), 16, 16, (1024, 1024, 0, 0), 128)
    print('SW128 MN-major result stats: absmax %.3e, zeros %.2f, nan %.2f' % (np.nanmax(np.abs(got)), float((got == 0).mean()),
                                                                            float(np.isnan(got).mean())))
    print({k: bool(v) for k, v in ok.items()})
    print('probe done')


if __name__ == '__main__':
    main()
