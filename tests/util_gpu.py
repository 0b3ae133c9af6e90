"""Helpers shared by the -m gpu parity tests: thin callers of the C ABI with torch tensors as device memory."""
import ctypes as C

import numpy as np
import torch

from pufferlib_b200 import _native


def gae_device(rewards_tm, values_tm, dones_tm, gamma, lam, want_returns=True):
    """rewards/values/dones: numpy [H, N] (arrival order).  Returns (advantages_sorted, returns_sorted) numpy."""
    h, n = rewards_tm.shape
    dev = torch.device('cuda')
    r = torch.as_tensor(np.ascontiguousarray(rewards_tm), device=dev)
    v = torch.as_tensor(np.ascontiguousarray(values_tm), device=dev)
    d = torch.as_tensor(np.ascontiguousarray(dones_tm), device=dev)
    adv = torch.full((n * h,), float('nan'), device=dev)
    ret = torch.full((n * h,), float('nan'), device=dev) if want_returns else None
    lib = _native.lib()
    ws = torch.zeros(max(16, lib.pb_gae_workspace_bytes(n, h)), dtype=torch.uint8, device=dev)
    _native.check(lib.pb_gae(_native.ptr(r), _native.ptr(v), _native.ptr(d), _native.ptr(adv), _native.ptr(ret),
                             n, h, C.c_float(gamma), C.c_float(lam), _native.ptr(ws), ws.numel(),
                             _native.stream_ptr()))
    torch.cuda.synchronize()
    assert int(ws.to(torch.int32).abs().sum()) == 0, 'workspace must be left zeroed'
    return adv.cpu().numpy(), (ret.cpu().numpy() if want_returns else None)


def sorted_from_time_major(x_tm):
    """[H, N] arrival order -> flat sorted order f = e*H + t."""
    return np.ascontiguousarray(x_tm.T).reshape(-1)


def gae_tolerance_check(adv, ref32, ref64):
    """North-star tolerance: fp32 GAE within 1e-5 relative of the reference.  Errors are measured against the float64
    chain on the scale max(1, |A|).  Where the chain is so long and undamped (gamma*lambda ~ 1, no dones) that the
    reference's OWN fp32 rounding exceeds 1e-5, the bar is "no less accurate than 2x the reference's error": two
    fp32 evaluations of an ill-conditioned sum cannot agree better than either is accurate."""
    scale = np.maximum(1.0, np.abs(ref64))
    err = (np.abs(adv.astype(np.float64) - ref64) / scale).max()
    ref_err = (np.abs(ref32.astype(np.float64) - ref64) / scale).max()
    assert err <= max(1e-5, 2 * ref_err), f'cuda err {err:.3e} vs reference fp32 err {ref_err:.3e}'
    direct = (np.abs(adv - ref32) / np.maximum(1.0, np.abs(ref32))).max()
    assert direct <= max(1e-5, err + ref_err), f'direct diff {direct:.3e}'
    return err, ref_err
