"""Test-only stand-in for matplotlib (absent in the build image): hydroscatt's graph.py imports it at module load.
Every attribute resolves to a no-op so the scripts' plotting calls do nothing."""
import sys
import types


class _Noop:
    def __getattr__(self, name):
        return _Noop()

    def __call__(self, *a, **k):
        return _Noop()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Noop()


def use(*a, **k):
    pass


def __getattr__(name):
    return _Noop()


for _m in ("pyplot", "cm", "colors", "dates", "ticker"):
    mod = types.ModuleType("matplotlib." + _m)
    mod.__getattr__ = lambda name: _Noop()
    sys.modules["matplotlib." + _m] = mod
