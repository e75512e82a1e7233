"""TRIQS availability switches (python/triqs_support.py.in:31-78 in the reference).

This package never links against TRIQS: Green-function containers (``GfImTime``, ``GfReFreq`` ...) are outside the
B200 hot path, so ``if_no_triqs()`` is always true and everything marked ``@require_triqs`` raises
``NotImplementedError`` with the reference's message -- the same behaviour as a reference build configured with
``USE_TRIQS=OFF``."""
import functools


def if_triqs_1():
    return False


def if_triqs_2():
    return False


def if_no_triqs():
    return True


def require_triqs(func):
    """Mark ``func`` as needing TRIQS: calling it raises NotImplementedError."""
    @functools.wraps(func)
    def needs_triqs(*args, **kwargs):
        raise NotImplementedError(
            "The functionality provided by {} is only available with TRIQS.".format(func.__name__))
    doc = func.__doc__ or ""
    needs_triqs.__doc__ = ".. warning::\n\n    This requires TRIQS support!\n\n" + doc
    return needs_triqs


def assert_text_files_equal(a, b):
    """Line-by-line comparison of two text files ignoring trailing whitespace (used by the reference's text goldens,
    e.g. test/python/logtaker.py:60-61)."""
    with open(a) as fa, open(b) as fb:
        la = [l.rstrip() for l in fa.read().rstrip().splitlines()]
        lb = [l.rstrip() for l in fb.read().rstrip().splitlines()]
    assert la == lb, "text files %s and %s differ" % (a, b)
