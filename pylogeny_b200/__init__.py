"""pylogeny_b200 -- B200-native tree-scoring engine behind Pylogeny's scoring API.

Host-side mirror of the reference modules on the hot path (same names, argument
meaning and error behaviour):

    pylogeny_b200.fitch          <- fitch.cpp          (calculateCost)
    pylogeny_b200.libpllWrapper  <- libpllWrapper.c    (new/destroy/getLogLikelihood/...)
    pylogeny_b200.parsimony      <- parsimony.py       (profile_set, site_profile)
    pylogeny_b200.scoring        <- scoring.py         (getParsimony*, getLogLikelihood)
    pylogeny_b200.pll            <- pll.py             (dataModel, partitionModel)

All compute goes through the C-ABI library built from pylogeny_b200/csrc (hand-written
sm_100a CUDA); see include/pylogeny_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
