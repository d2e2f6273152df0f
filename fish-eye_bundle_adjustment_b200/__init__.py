"""B200-native Gauss-Newton hot path of the equidistant fish-eye bundle adjustment.

Host-side mirror of the reference's entry points (``main``, ``BatchRun``, ``ReadFiles``,
``findSetting``, ``Buildxhat``, ``BuildRSD``) over a C-ABI CUDA library (``include/feba.h``).
The directory name carries a hyphen, so it is imported through the ``feba_b200`` shim at the
repo root (``import feba_b200``).
"""
from .formats import ReadFiles, findSetting, read_string_table          # noqa: F401
from .problem import Settings, Problem, Buildxhat, load_problem, save_problem  # noqa: F401
