"""Engine-side `pll` module: the two classes the reference's scoring path uses from pll.py
(`dataModel`, pll.py:12-73; `partitionModel`, pll.py:75-144) with the same constructor arguments,
method names, attributes (`model`, `instance`, `handle`, `length`, `isprotein`) and side effects
(`alignment._pllmodel`, `alignment.paths['pll']`), on top of pylogeny_b200.libpllWrapper.

Not a transcription: the partition file is produced by one formatter shared by both creators and
validated before it is written (the engine scores one partition spanning the alignment; a file
with several partitions is refused at `libpllWrapper.new` with EngineError, not mis-scored), the
model of an alignment is resolved in one helper that the batched entry points of scoring.py share,
and both classes are context managers.  A dataModel MUST BE CLOSED after use, as in the reference.
"""
import os
import tempfile

from . import alignment as _alignment
from . import libpllWrapper as _wrapper


def partition_lines(models, names, ranges):
    """libpll partition syntax, one "<MODEL>, <name> = <from>-<to>" line per partition."""
    if not (len(models) == len(names) == len(ranges)):
        raise ValueError("models, partition names and ranges must have the same length")
    out = []
    for model, name, (lo, hi) in zip(models, names, ranges):
        if int(lo) < 1 or int(hi) < int(lo):
            raise ValueError("partition %s: bad site range %s-%s" % (name, lo, hi))
        out.append("%s, %s = %d-%d\n" % (model, name, int(lo), int(hi)))
    return out


def model_for(ali):
    """The partition model an alignment carries, made on first use (what pll.py:35-43 does inline):
    a simple whole-alignment model, remembered on the alignment and in its temp-file registry."""
    model = getattr(ali, '_pllmodel', None)
    if model is None:
        model = partitionModel(ali)
        model.createSimpleModel()
        ali._pllmodel = model
        ali.paths['pll'] = model.getFileName()
    return model


class partitionModel(object):
    """A partition model file in libpll's syntax (pll.py:75-144)."""

    def __init__(self, ali):
        self.handle = tempfile.NamedTemporaryFile(mode='w', delete=False)
        self.length = len(ali)
        self.isprotein = ali.getDataType() == 'protein'

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def getFileName(self):
        return self.handle.name

    def _write(self, lines):
        self.handle.writelines(lines)
        self.handle.close()

    def createSimpleModel(self, pmodel='WAG'):
        """One partition over all sites: the protein model name (engine: WAG or JTT) or DNA."""
        self._write(partition_lines([pmodel if self.isprotein else 'DNA'], ['p1'], [(1, self.length)]))

    def createModel(self, models, partnames, ranges):
        """Several partitions (pll.py:118-137).  The file is written as asked; the engine refuses
        more than one partition when a problem is created on it."""
        self._write(partition_lines(models, partnames, ranges))

    def close(self):
        if not self.handle.closed:
            self.handle.close()
        if os.path.exists(self.handle.name):
            os.unlink(self.handle.name)


class dataModel(object):
    """A topology + phylipFriendlyAlignment bound into an engine problem (one per tree, like the
    reference); scores log-likelihood after branch-length smoothing and returns the optimised tree."""

    def __init__(self, topo, alignm, model=None):
        if type(alignm) is not _alignment.phylipFriendlyAlignment:
            raise TypeError('libpll interfacing needs phylipFriendlyAlignment.')
        self.model = model if model else model_for(alignm)
        self.instance = _wrapper.new(alignm.getPhylip(), topo.toUnrootedNewick(), self.model.getFileName())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def getNewickString(self):
        return _wrapper.getNewickString(self.instance)

    def getLogLikelihood(self):
        return _wrapper.getLogLikelihood(self.instance)

    def close(self):
        if self.instance is not None:
            _wrapper.destroy(self.instance)
            self.instance = None
