"""``pool=`` duck-type (utils.py:2358-2381): an object with ``.map`` and ``.size``.

dynesty only allows ``queue_size > 1`` when a pool is given (utils.py:2366-2367) and takes
the default queue size from ``pool.size``.  With the B200 samplers the queue is evaluated
by ONE kernel launch inside ``prepare_sampler``; ``map`` is then the in-process builtin
(the per-item "task" is the identity), and ``size`` is the number of chains per launch.
The same object also serves ``Sampler``'s other uses of the pool (initial live points,
sampler.py:148-158): plain in-process map of the host callables.
"""


class B200Pool:
    def __init__(self, n_chains=2048):
        self.size = int(n_chains)

    def map(self, fn, iterable):
        return list(map(fn, iterable))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def join(self):
        pass
