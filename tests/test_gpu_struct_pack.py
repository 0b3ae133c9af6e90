"""pb_struct_pack / pb_struct_unpack vs numpy's own aligned structured assignment (the reference's `emulate`)."""
import numpy as np
import pytest
import torch

from pufferlib_b200 import emulation, spaces
from pufferlib_b200.namespace import namespace

pytestmark = pytest.mark.gpu


def spaces_under_test():
    yield spaces.Dict({'x': spaces.Box(-1.0, 1.0, (1, 2), np.float32),
                       'y': spaces.Dict({'a': spaces.Box(0, 255, (7, 7), np.uint8),
                                         'b': spaces.Box(-1024, 1024, (2, 3), np.int32)})})
    yield spaces.Dict({'screen': spaces.Box(0, 255, (18, 20), np.uint8)})
    yield spaces.Tuple([spaces.Box(0, 255, (3,), np.uint8), spaces.Box(-5, 5, (2, 2), np.int16),
                        spaces.Box(-1, 1, (5,), np.float64)])
    yield spaces.Dict({'xx': spaces.Box(-1, 1, (1, 2), np.float32), 'yy': spaces.Box(-1, 1, (4, 5), np.float32)})


def fill(rec, dt, rng, n):
    """Random content for every leaf of a structured array; returns the nested sample of torch tensors."""
    sample = {} if dt.names and not dt.names[0].startswith('f') else []
    for name in dt.names:
        fdt = dt.fields[name][0]
        if fdt.fields is not None:
            sub = fill(rec[name], fdt, rng, n)
        else:
            base, shape = fdt.subdtype if fdt.subdtype else (fdt, ())
            if base.kind == 'f':
                arr = rng.standard_normal((n, *shape)).astype(base)
            else:
                info = np.iinfo(base)
                arr = rng.integers(info.min, int(info.max) + 1, (n, *shape), dtype=base)
            rec[name] = arr
            sub = torch.as_tensor(arr, device='cuda')
        if isinstance(sample, dict):
            sample[name] = sub
        else:
            sample.append(sub)
    return sample if isinstance(sample, dict) else tuple(sample)


@pytest.mark.parametrize('idx', range(4))
@pytest.mark.parametrize('n', [1, 33, 4096])
def test_struct_pack_unpack(idx, n):
    sp = list(spaces_under_test())[idx]
    dt = emulation.dtype_from_space(sp)
    rng = np.random.default_rng(idx * 100 + n)
    rec = np.zeros(n, dtype=dt)
    sample = fill(rec, dt, rng, n)
    packed = emulation.emulate_batch(dt, sample)
    ref = rec.view(np.uint8).reshape(n, dt.itemsize)
    assert np.array_equal(packed.cpu().numpy(), ref)                        # incl. zero padding bytes
    # typed zero-copy views of the packed batch (nativize_tensor) see the original leaves
    native = emulation.nativize_dtype(namespace(observation_dtype=np.dtype(np.uint8), emulated_observation_dtype=dt))
    views = emulation.nativize_tensor(packed, native)
    back = emulation.nativize_batch(dt, packed)
    for path, base, shape, off, nb in emulation.leaf_layout(dt):
        v, b, r = views, back, rec
        for k in path:
            v, b, r = v[k], b[k], r[k]
        assert np.array_equal(v.cpu().numpy(), r) and np.array_equal(b.cpu().numpy(), r)


def test_struct_pack_rejects_bad_layout():
    import ctypes as C
    from pufferlib_b200 import _native
    from pufferlib_b200.exceptions import APIUsageError
    lay = emulation.Layout(n_leaves=1, record_bytes=8)
    lay.offset[0], lay.nbytes[0] = 4, 8          # runs past the record
    ptrs = (C.c_void_p * 32)()
    x = torch.zeros(64, dtype=torch.uint8, device='cuda')
    ptrs[0] = x.data_ptr()
    with pytest.raises(APIUsageError):
        _native.check(_native.lib().pb_struct_pack(C.byref(lay), ptrs, _native.ptr(x), 8, 4, _native.stream_ptr()))
