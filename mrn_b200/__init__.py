"""mrn_b200: B200-native (sm_100a) implementation of MRN's multiplexed-routing train / infer step.

Host-side mirror of the reference API lives in mrn_b200.modules.{model,dm_router} and
mrn_b200.il_modules.mrn; the compute is hand-written CUDA behind the C ABI of include/mrn_b200.h."""
__version__ = "0.1.0"
