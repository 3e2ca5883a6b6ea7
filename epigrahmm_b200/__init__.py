"""epigrahmm_b200 -- the EM hot path of epigraHMM on B200 (sm_100a).

The product is ``libehmm_b200.so`` (C ABI in ``include/ehmm.h``); this package is the
host-side mirror of the reference's R/Rcpp interface for that path:

* ``epigrahmm_b200.api``     the Rcpp exports by name (expStep, maxStepProb, ...), keyed by
                             the hdf5 path string exactly like R/RcppExports.R
* ``epigrahmm_b200.em``      the EM drivers of R/base-EM.R (emConsensus, outerEMDifferential,
                             innerEMDifferential, emInitialization)
* ``epigrahmm_b200.Session`` the device-resident state that replaces the HDF5 file

Importing the package loads the CUDA library; there is no CPU fallback.
"""
from ._lib import EhmmError, SO_PATH  # noqa: F401  (raises ImportError when the library is missing)
from .session import Session, aggregate, reweight  # noqa: F401

__all__ = ["Session", "EhmmError", "aggregate", "reweight", "SO_PATH"]
