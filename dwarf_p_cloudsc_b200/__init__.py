"""cloudsc-b200: B200-native CLOUDSC hot path behind the dwarf-cloudsc-c driver contract.

Sub-modules:
  params   ctypes mirrors of TECLDP / kernel constants / the 56-entry field table
  state    NumPy restatement of load_state/expand for tests and benchmarks
  h5lite   dependency-free reader/writer for the dwarf's HDF5 subset
  api      ctypes binding of libcloudsc_b200.so (imports fail loudly if the library is not built)
  validate the dwarf's error table (cloudsc_validate) with a real pass/fail tolerance
  build    in-tree nvcc/gcc build
"""
__all__ = ["params", "state", "h5lite", "api", "validate", "build"]
