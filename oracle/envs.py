"""ctypes front-end of ``oracle/csrc/envs.c``: a Serial-like vectoriser over the spec'd breakout/snake/pong.
TEST INFRASTRUCTURE (see oracle/__init__.py).  API mirrors vector.Serial: async_reset / send / recv."""
import ctypes as C

import numpy as np

from . import build as _build

KINDS = {'breakout': 1, 'snake': 2, 'pong': 3}
OBS = {'breakout': ((128,), np.float32), 'snake': ((16, 16), np.uint8), 'pong': ((4, 84, 84), np.uint8)}
NUM_ACTIONS = {'breakout': 4, 'snake': 4, 'pong': 6}


def _lib():
    lib = _build.load()
    if not getattr(lib, '_envs_typed', False):
        lib.oracle_vec_create.restype = C.c_void_p
        lib.oracle_vec_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int)]
        lib.oracle_vec_destroy.argtypes = [C.c_void_p]
        lib.oracle_vec_reset.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        lib.oracle_vec_step.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        lib.oracle_set_threads.argtypes = [C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        lib._envs_typed = True
    return lib


class OracleVec:
    def __init__(self, kind, num_envs, env_index_offset=0, iparam=(), threads=None):
        self.lib = _lib()
        self.kind, self.n = kind, num_envs
        ip = (C.c_int * 8)(*(list(iparam) + [0] * (8 - len(iparam))))
        self.h = C.c_void_p(self.lib.oracle_vec_create(KINDS[kind], num_envs, env_index_offset, ip))
        assert self.h
        if threads:
            self.lib.oracle_set_threads(threads)
        shape, dtype = OBS[kind]
        self.observations = np.zeros((num_envs, *shape), dtype=dtype)
        self.rewards = np.zeros(num_envs, dtype=np.float32)
        self.terminals = np.zeros(num_envs, dtype=bool)
        self.truncations = np.zeros(num_envs, dtype=bool)
        self.masks = np.ones(num_envs, dtype=bool)
        self.info_return = np.zeros(num_envs, dtype=np.float64)
        self.info_length = np.zeros(num_envs, dtype=np.int32)
        self.info_score = np.zeros(num_envs, dtype=np.float32)
        self.agent_ids = np.arange(num_envs)
        self.infos = []
        self.collect_infos = True

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def async_reset(self, seed=42):
        self.lib.oracle_vec_reset(self.h, C.c_uint64(seed % (1 << 64)), self._p(self.observations),
                                  self._p(self.rewards), self._p(self.terminals), self._p(self.truncations),
                                  self._p(self.masks))
        self.infos = []

    def send(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int64)
        self.lib.oracle_vec_step(self.h, self._p(a), self._p(self.observations), self._p(self.rewards),
                                 self._p(self.terminals), self._p(self.truncations), self._p(self.masks),
                                 self._p(self.info_return), self._p(self.info_length), self._p(self.info_score))
        if self.collect_infos:
            self.infos = [{'episode_return': float(self.info_return[i]), 'episode_length': int(self.info_length[i]),
                           'score': float(self.info_score[i])} for i in np.nonzero(self.terminals)[0]]

    def recv(self):
        return (self.observations, self.rewards, self.terminals, self.truncations, self.infos, self.agent_ids,
                self.masks)

    def close(self):
        if self.h:
            self.lib.oracle_vec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
