"""maveric-slam tracking hot path on B200 (sm_100a).

Host-side mirror of the reference's interface for this path over the C ABI of
``libmaveric_b200.so`` (``include/maveric_b200.h``, ``include/maveric_slam_compat.h``):

* :mod:`.lib`       -- ctypes binding of the shared library (fails loudly if missing)
* :mod:`.tracking`  -- ``Frame``, ``compute_softmax``, ``compute_top_N``, ``track`` ... with the
                       reference's names and argument meaning, plus the batched sequence API
* :mod:`.synth`     -- synthetic KITTI-shaped frames (numpy twin of the CUDA generator)
* :mod:`.build`     -- nvcc recipe (sm_100a)

No module here imports anything from ``oracle/``: that directory is the test checker.
"""
from . import synth  # noqa: F401  (numpy only)

__all__ = ["synth", "lib", "tracking", "build"]
