"""Tiny stand-in for more_itertools (absent from the image): only the functions the reference's Python layer
calls, with the documented semantics.  Golden-generation infrastructure only (see README.md)."""
from collections import defaultdict, deque
from itertools import chain, repeat, starmap

flatten = chain.from_iterable


def unique_everseen(it, key=None):
    seen = set()
    for x in it:
        k = x if key is None else key(x)
        if k not in seen:
            seen.add(k)
            yield x


def duplicates_everseen(it, key=None):
    seen = set()
    for x in it:
        k = x if key is None else key(x)
        if k in seen:
            yield x
        else:
            seen.add(k)


def all_unique(it, key=None):
    seen = set()
    for x in it:
        k = x if key is None else key(x)
        if k in seen:
            return False
        seen.add(k)
    return True


def map_reduce(it, keyfunc, valuefunc=None, reducefunc=None):
    d = defaultdict(list)
    for x in it:
        d[keyfunc(x)].append(x if valuefunc is None else valuefunc(x))
    if reducefunc is not None:
        for k in d:
            d[k] = reducefunc(d[k])
    d.default_factory = None
    return d


_marker = object()


def first(it, default=_marker):
    for x in it:
        return x
    if default is _marker:
        raise ValueError('first() was called on an empty iterable')
    return default


def locate(it, pred=bool):
    return (i for i, x in enumerate(it) if pred(x))


def repeatfunc(f, times=None, *args):
    return starmap(f, repeat(args) if times is None else repeat(args, times))


def transpose(it):
    return zip(*it)


def ilen(it):
    return sum(1 for _ in it)


def consume(it, n=None):
    if n is None:
        deque(it, maxlen=0)
    else:
        for _ in zip(range(n), it):
            pass


def numeric_range(*args):
    start, stop, step = (0, args[0], 1) if len(args) == 1 else (args[0], args[1], 1) if len(args) == 2 else args
    x = start
    while (x < stop) if step > 0 else (x > stop):
        yield x
        x = x + step
