"""Mirror of the reference's pll.py (pll.py:12-144): `dataModel` (one problem instance per tree,
MUST BE CLOSED AFTER USE) and `partitionModel` (temp partition file "DNA, p1 = 1-N" /
"WAG, p1 = 1-N"), on top of pylogeny_b200.libpllWrapper."""
import os
from tempfile import NamedTemporaryFile as NTempFile

from . import alignment
from .libpllWrapper import destroy, getLogLikelihood, getNewickString, new


class dataModel(object):
    """A tree (as topology) + alignment bound into an engine problem; allows log-likelihood scoring."""

    def __init__(self, topo, alignm, model=None):
        if type(alignm) != alignment.phylipFriendlyAlignment:
            raise TypeError('libpll interfacing needs phylipFriendlyAlignment.')
        if not model:
            if hasattr(alignm, '_pllmodel'):
                self.model = alignm._pllmodel
            else:
                self.model = partitionModel(alignm)
                self.model.createSimpleModel()
                alignm._pllmodel = self.model
                alignm.paths['pll'] = self.model.getFileName()
        else:
            self.model = model
        modf = self.model.getFileName()
        self.instance = new(alignm.getPhylip(), topo.toUnrootedNewick(), modf)

    def getNewickString(self):
        return getNewickString(self.instance)

    def getLogLikelihood(self):
        return getLogLikelihood(self.instance)

    def close(self):
        destroy(self.instance)


class partitionModel(object):
    """A partition model file in libpll's syntax (pll.py:75-144)."""

    def __init__(self, ali):
        self.handle = NTempFile(mode='w', delete=False)
        self.length = len(ali)
        self.isprotein = (ali.getDataType() == 'protein')

    def getFileName(self):
        return self.handle.name

    def createSimpleModel(self, pmodel='WAG'):
        if self.isprotein:
            simplemodel = "%s, p1 = 1-%d\n" % (pmodel, self.length)
        else:
            simplemodel = "DNA, p1 = 1-%d\n" % (self.length)
        self.handle.write(simplemodel)
        self.handle.close()

    def createModel(self, models, partnames, ranges):
        for m in range(len(models)):
            rl, ru = ranges[m]
            self.handle.write("%s, %s = %d-%d\n" % (models[m], partnames[m], rl, ru))
        self.handle.close()

    def close(self):
        os.unlink(self.handle.name)
