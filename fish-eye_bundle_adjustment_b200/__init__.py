"""B200-native Gauss-Newton hot path of the equidistant fish-eye bundle adjustment.

Host-side mirror of the reference's entry points (``main``, ``BatchRun``, ``ReadFiles``,
``findSetting``, ``Buildxhat``) over a C-ABI CUDA library (``include/feba.h``, ``libfeba.so``).
The directory name carries a hyphen, so it is imported through the ``feba_b200`` shim at the
repo root (``import feba_b200``).  The CUDA library is loaded lazily by ``lib.load()`` and there
is no CPU fallback for the hot path.
"""
from .formats import ReadFiles, findSetting, read_string_table          # noqa: F401
from .problem import Settings, Problem, Buildxhat, load_problem, save_problem  # noqa: F401
from .lib import Handle, FebaError                                       # noqa: F401
from .main import main, BatchRun, adjust, findfiles, write_rsd, write_par, covariance_outputs  # noqa: F401
from .batch import adjust_batch                                          # noqa: F401
from .pack import load_problem_native, pack_files, PackError            # noqa: F401
from . import synth, lib, build, pack, report                                    # noqa: F401
