"""rwkv.cpp_b200 -- host-side mirror of rwkv.cpp's Python interface over the B200-native librwkv.so.

The directory name contains a dot, so it is imported through ``__graft_entry__.load_package()``
(which registers it as the module ``rwkv_cpp_b200``). Public surface mirrors the reference package
python/rwkv_cpp/ (rwkv_cpp_shared_library.py, rwkv_cpp_model.py):

    RWKVSharedLibrary, load_rwkv_shared_library, RWKVModel
"""
from .shared_library import RWKVSharedLibrary, RWKVContext, load_rwkv_shared_library, library_path
from .model import RWKVModel
from . import pipeline

__all__ = ["RWKVSharedLibrary", "RWKVContext", "load_rwkv_shared_library", "library_path", "RWKVModel", "pipeline"]
