"""B200-native batched NMPC (PANOC + ALM/PM) solver — drop-in for the OpEn-generated
solver of Woodenonez/DyObAv-MPCnWTA-Warehouse (``trajectory_tracker.py:61-62,362``)."""
from .problem import Dims, RobotSpec, SolverSettings, MpcConfig, EXIT_STATUS_NAMES  # noqa: F401

__version__ = "0.1.0"
