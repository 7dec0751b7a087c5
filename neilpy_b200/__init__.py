"""neilpy_b200 -- the SMRF ground-classification path of thomaspingel/neilpy on B200.

Drop-in for `from neilpy import smrf, create_dem, progressive_filter,
inpaint_nans_by_springs` (neilpy/__init__.py:1), plus `read_las` (the step before the path) and the
classification write-back of the reference's laspy notebook, and `slope` / `aspect` / `hillshade` / `pssm`
(the step after it); everything else in neilpy is out of scope.
Compute lives in libsmrf_b200.so (hand-written CUDA for sm_100a behind a C ABI,
include/smrf_b200.h); build it with `python -m neilpy_b200.build`.
"""
from .affine import Affine
from .api import create_dem, inpaint_nans_by_fda, inpaint_nans_by_springs, progressive_filter, smrf
from .las import classify_las, read_las, read_las_device
from .terrain import aspect, hillshade, pssm, slope

__all__ = ['smrf', 'create_dem', 'progressive_filter', 'inpaint_nans_by_springs', 'inpaint_nans_by_fda', 'Affine',
           'read_las', 'read_las_device', 'classify_las', 'slope', 'aspect', 'hillshade', 'pssm']
__version__ = '0.1.0'
