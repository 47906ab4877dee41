"""Developer aid: run the test-suite the way the GPU box sees it — with the reference checkout hidden.

/root/reference exists only in the build container; nothing the GPU box runs may need it. This wrapper makes every
path below /root/reference look absent to os.path.isdir / exists / isfile and open() and then hands over to pytest:

    python tools/run_tests_without_reference.py -m "not gpu"          # reference-dependent tests must SKIP, not fail
    python tools/run_tests_without_reference.py -m gpu --collect-only  # GPU tests must collect without it
"""
import builtins
import os
import sys

_isdir, _exists, _isfile, _open = os.path.isdir, os.path.exists, os.path.isfile, builtins.open


def _hidden(p):
    return isinstance(p, (str, bytes, os.PathLike)) and os.fsdecode(os.fspath(p)).startswith("/root/reference")


def _open_guarded(p, *a, **k):
    if _hidden(p):
        raise FileNotFoundError(p)
    return _open(p, *a, **k)


if __name__ == "__main__":
    os.path.isdir = lambda p: False if _hidden(p) else _isdir(p)
    os.path.exists = lambda p: False if _hidden(p) else _exists(p)
    os.path.isfile = lambda p: False if _hidden(p) else _isfile(p)
    builtins.open = _open_guarded
    import pytest
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.exit(pytest.main([os.path.join(here, "tests"), "-q", "-p", "no:cacheprovider"] + sys.argv[1:]))
