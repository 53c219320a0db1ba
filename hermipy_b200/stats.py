"""Call statistics and debug printing (reference: hermipy/stats.py:29-93).

`stats[<function>-<module>] = {'Calls': n, 'Time': seconds}` accumulates over every call of a function
decorated with `log_stats()`; `settings['trails']` prints an indented trail, `settings['debug']` (through
`debug()`) prints the arguments.  Times are host wall-clock around the whole call, so for the CUDA-backed
functions they include the host<->device copies and the final stream synchronisation."""
import functools
import time

import numpy as np

from .settings import settings

stats = {}
threshold = 0      # trails: only report calls slower than this many seconds
_depth = [0]


def _key(function):
    return function.__name__ + '-' + function.__module__


def _short(arg):
    if isinstance(arg, np.ndarray):
        return "numpy ndarray " + str(arg.shape)
    if isinstance(arg, list):
        return str([_short(a) for a in arg])
    return str(arg)


def debug(force=False):
    def decorator(function):
        @functools.wraps(function)
        def wrapper(*args, **kwargs):
            if force or settings['debug']:
                print("▶ Entering function " + _key(function))
                print("-▲ args")
                for a in args:
                    print("--● A " + _short(a))
                print("-▲ kwargs")
                for name, value in kwargs.items():
                    print("--● KW " + str(name) + ": " + _short(value))
            return function(*args, **kwargs)
        return wrapper
    return decorator


def log_stats(force=False):
    def decorator(function):
        key = _key(function)

        @functools.wraps(function)
        def wrapper(*args, **kwargs):
            verbose = force or settings['trails']
            prefix = " " * _depth[0] + str(_depth[0]) + " "
            if verbose:
                print(prefix + "Entering " + key)
            _depth[0] += 1
            t0 = time.time()
            try:
                return function(*args, **kwargs)
            finally:
                elapsed = time.time() - t0
                _depth[0] -= 1
                entry = stats.setdefault(key, {'Calls': 0, 'Time': 0})
                entry['Calls'] += 1
                entry['Time'] += elapsed
                if verbose and elapsed > threshold:
                    print(prefix + "Function " + key + ": " + str(elapsed))
        return wrapper
    return decorator


def print_stats():
    print('\n-- [statistics] --')
    for key, entry in stats.items():
        print("| '" + key + "': Calls: " + str(entry['Calls']) + ", Time: " + str(entry['Time']))
    print('--')
