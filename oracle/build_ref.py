"""Compile the reference's OWN ``c_gae.pyx`` from where it lies under /root/reference.  TEST INFRASTRUCTURE.

Recipe (no reference build system involved): ``cython`` translates /root/reference/c_gae.pyx to C in a temp
dir, ``gcc`` compiles that one file into ``oracle/_ref/c_gae*.so`` (git-ignored, travels to the GPU box with
the snapshot).  Nothing is copied into the repo history.  ``load()`` imports the built module if present.
"""
import glob
import importlib.util
import os
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, '_ref')


def build(force=False):
    existing = glob.glob(os.path.join(OUT, 'c_gae*.so'))
    if existing and not force:
        return existing[0]
    src = os.path.join(REF, 'c_gae.pyx')
    if not os.path.exists(src):
        return None
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    so = os.path.join(OUT, 'c_gae' + ext)
    with tempfile.TemporaryDirectory() as tmp:
        c_file = os.path.join(tmp, 'c_gae.c')
        subprocess.check_call([sys.executable, '-m', 'cython', '-3', src, '-o', c_file])
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fwrapv', '-fno-strict-aliasing',
                               '-I', sysconfig.get_paths()['include'], '-I', np.get_include(),
                               '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION', c_file, '-o', so])
    return so


def load():
    """Return the reference's compiled ``c_gae`` module, or None if it was never built."""
    found = glob.glob(os.path.join(OUT, 'c_gae*.so'))
    if not found:
        return None
    spec = importlib.util.spec_from_file_location('c_gae', found[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force=True))
