"""kbnet_b200 -- B200-native (sm_100a) implementation of KBNet's SparseToDensePool and Calibrated Backprojection
layers behind the reference's own nn.Module interface, plus the encoder/decoder wiring around them.

Importable as ``kbnet_b200`` (see kbnet_b200.py at the repository root; this directory keeps the name the build
contract asks for).  The compute lives in libkbnet_b200.so (csrc/, C ABI in include/kbnet_b200.h).
"""
from . import _lib, build as _build, dist, io, ops
from .modules import (CalibratedBackprojectionBlock, OutlierRemoval, SparseToDensePool, install, invalidate_packs, set_precision,
                      get_precision)
from .model import GraphRunner, KBNetEncoder, KBNetModel, MultiScaleDecoder

__all__ = ['SparseToDensePool', 'CalibratedBackprojectionBlock', 'OutlierRemoval', 'KBNetEncoder', 'MultiScaleDecoder', 'KBNetModel',
           'install', 'invalidate_packs', 'set_precision', 'get_precision', 'set_autotune', 'ops', 'dist', 'io', 'build_library', 'library_path']


def build_library(force=False, verbose=False):
    return _build.build(force=force, verbose=verbose)


def set_autotune(on=True):
    """Measured tile selection for the tensor-core engine: the first call for each new layer shape times the best candidate
    tilings (synchronising once) and caches the winner.  Off by default."""
    _lib.load().kbnet_set_autotune(1 if on else 0)


def library_path():
    return _lib.LIB_PATH
