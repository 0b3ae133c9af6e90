"""Build / load the oracle's plain-C restatements (gcc, OpenMP).  TEST INFRASTRUCTURE.

``python -m oracle.build`` compiles ``oracle/csrc/*.c`` into ``oracle/_build/liboracle.so``.
``load()`` returns the ctypes handle, building on demand (gcc only; no GPU, no torch).
"""
import ctypes
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'liboracle.so')
_lib = None


def sources():
    return sorted(glob.glob(os.path.join(HERE, 'csrc', '*.c')))


def build(force=False):
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(HERE, 'csrc', '*.h'))
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-std=c11', '-ffp-contract=off', '-fopenmp', '-Wall',
           '-o', SO] + srcs + ['-lm']
    subprocess.check_call(cmd)
    return SO


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


if __name__ == '__main__':
    print(build(force=True))
