"""dad_b200 -- B200-native (sm_100a) implementation of the DAD cross-domain training step.

Import as `import dad_b200` (see /dad_b200.py; the directory name is not a Python identifier).
The CUDA library (csrc/ -> libdad_sm100a.so, C ABI in /include/dad_sm100a.h) is loaded lazily by
dad_b200._lib; every compute entry point raises if it is missing -- there is no CPU fallback."""
__version__ = "0.1.0"
