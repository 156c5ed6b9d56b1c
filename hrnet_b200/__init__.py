"""Importable alias of the `hrnet-hand-pose-estimation_b200/` package directory.

The product package lives in a directory whose name is not a Python identifier; this shim makes it
importable as `hrnet_b200` (``hrnet_b200.models.pose_hrnet`` etc.).
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "hrnet-hand-pose-estimation_b200")
__path__.insert(0, _PKG_DIR)
PACKAGE_DIR = _PKG_DIR
