"""Import shim (test infrastructure): the reference's utils.py only needs
gymnasium.Wrapper as a base class and gymnasium.spaces.Box (utils.py:11,238,244)."""


class Wrapper:
    def __init__(self, env):
        self.env = env


class _Spaces:
    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


spaces = _Spaces()
