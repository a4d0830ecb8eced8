"""Import shim: the product package lives in ``mars5-tts_b200/`` (a directory name Python cannot import directly).

``import mars5_tts_b200`` resolves sub-modules (``capi``, ``engine``, ``weights`` ...) from that directory.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PKG_DIR = _os.path.join(_ROOT, "mars5-tts_b200")
__path__.append(PKG_DIR)
