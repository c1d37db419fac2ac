"""parla_b200 -- B200-native implementation of PARLA's randomized-sketching hot path.

Package-level names follow parla/__init__.py:8-19 for the components on the path.
"""
__version__ = '0.1.0'

from .utils.sketching import (gaussian_operator, sjlt_operator, srct_operator, generate_srct, apply_srct,
                              sparse_sign_operator, orthonormal_operator, sampling_operator, as_device_operator)
from .utils.linalg_wrappers import orth
from .comps.sketchers.oblivious import SkOpGA, SkOpSJ, SkOpTC, SkOpSS, SkOpON, SkOpIN, SketchOpGen
from .comps.sketchers.aware import RS1, RowSketcher
from .comps.qb import QB1, QB2, QB3, QBDecomposer
from .comps.rangefinders import RF1, RangeFinder
from .comps.determiter.logging import SketchAndPrecondLog
from .comps.determiter.saddle import PcSS1, PcSS2, PrecondSaddleSolver, pcss1, pcss2
from .comps.determiter.pcg import pcg
from .drivers.least_squares import SPO, SAP1, SAP2, SSO1, OverLstsqSolver, SPU1, UnderLstsqSolver
from .drivers.saddlesys import SPS1, SPS2, SaddleSolver, sps
from .drivers.svd import SVD1, SVDecomposer
from .drivers.evd import EVD1, EVD2, EVDecomposer
from .parallel import RowSharded
from .comps.interpolative import ROCS1, RowOrColSelection, qrcp_osid
from .drivers.interpolative import OSID1, OSID2, OneSidedID, TSID1, TwoSidedID, CUR1, CURDecomposition
