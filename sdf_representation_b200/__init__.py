"""Importable alias of the ``sdf-representation_b200/`` package directory (a hyphen is not a valid module name).

All code lives in ``sdf-representation_b200/``; this shim only extends ``__path__`` so that
``import sdf_representation_b200`` and ``from sdf_representation_b200.networks import ImplicitNet`` work.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "sdf-representation_b200"))

from . import _lib, checkpoint, config, distributed, engine, feed, inference, losses, networks, optim  # noqa: E402,F401
from .networks import ImplicitNet  # noqa: E402,F401
from .losses import IGRLOSS, PointCloudLoss, WeightedSmoothL2Loss, compute_normal, get_points, gradient  # noqa: E402,F401
from .config import Configuration  # noqa: E402,F401
from .optim import FusedAdam  # noqa: E402,F401
from .feed import DeviceFeed, PinnedPrefetcher  # noqa: E402,F401
