"""ipm_b200 -- B200-native barrier interior-point engine (drop-in for fdeguire03/InteriorPoint-GPU's Python API).

The directory is named ``interiorpoint-gpu_b200`` (not an importable identifier); import it as ``ipm_b200``
(the repo-root shim package) or put this directory on ``sys.path`` and use the reference's flat module names
(``from LPSolver import LPSolver`` ...), exactly like the reference's own flat layout.
"""

from . import _abi  # noqa: F401

__all__ = ["_abi"]
