"""Stub: the reference imports tornadox (pnkraemer/tornadox, not vendored) in pde/*.py for `to_tornadox_ivp` only."""


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


ivp = _Anything()
init = _Anything()
ek0 = ek1 = step = _Anything()


def __getattr__(name):
    return _Anything()
