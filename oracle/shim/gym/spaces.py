"""Legacy `gym.spaces` stand-in: separate (never-instantiated) classes mirroring gymnasium's names."""


class Box: pass
class Dict: pass
class Discrete: pass
class MultiBinary: pass
class MultiDiscrete: pass
class Tuple: pass
