"""Operand-image and descriptor helpers shared by the tcgen05 probes (see check_umma_probe.py)."""
import numpy as np


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
