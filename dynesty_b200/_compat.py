"""Base classes of the two plug-in seams.

When the reference package is importable (``import dynesty``) the B200 classes
subclass ITS ``Bound`` / ``InternalSampler`` -- ``dynesty.NestedSampler`` checks
``isinstance(bound, bounding.Bound)`` (dynesty.py:496) and
``isinstance(sample, InternalSampler)`` (dynesty.py:155, 501) -- so they drop in.
On a machine without dynesty (the GPU box) the same classes sit on the minimal
mirrors below, which reproduce only the interface contract (names, argument
meaning, return tuple) documented in bounding.py:76-122 and
internal_samplers.py:23-203; ``dynesty_b200.nested`` drives them the same way.
"""
from collections import namedtuple
import warnings

try:  # pragma: no cover - depends on the environment
    from dynesty.bounding import Bound as BoundBase
    from dynesty.internal_samplers import (InternalSampler as InternalSamplerBase,
                                           SamplerReturn)
    HAVE_DYNESTY = True
except ImportError:
    HAVE_DYNESTY = False

    SamplerReturn = namedtuple('SamplerReturn', [
        'u', 'v', 'logl', 'ncalls', 'evaluation_history', 'tuning_info', 'proposal_stats'])

    class BoundBase:
        """Interface of a bounding distribution (bounding.py:76-122)."""

        def __init__(self, ndim):
            self.logvol = 0
            self.need_centers = False
            self.ndim = ndim

    class InternalSamplerBase:
        """Interface of an inner sampler (internal_samplers.py:36-203): holds the
        proposal ``scale`` and the per-call ``sampler_kwargs``; ``prepare_sampler``
        builds one argument per queue slot, the static ``sample`` turns one
        argument into a ``SamplerReturn``, ``tune`` feeds the statistics back."""

        def __init__(self, **kwargs):
            self.scale = 1
            self.input_kwargs = kwargs
            self.ndim = kwargs.get('ndim')
            self.sampler_kwargs = {k: kwargs.get(k) for k in ('nonbounded', 'periodic', 'reflective')}

        @property
        def update_bound_interval_ratio(self):
            return 1

        def _new_from_template(self, template_kwargs):
            merged = dict(self.input_kwargs)
            for k, v in template_kwargs.items():
                merged.setdefault(k, v)
            return self.__class__(**merged)

        def tune(self, tuning_info, update=False):
            pass

        @property
        def citations(self):
            return []
