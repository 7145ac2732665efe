"""metabo_b200 -- B200-native (sm_100a) hot path of MetaBO behind MetaBO's own Python API.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/metabo_b200.h) -> libmetabo_b200.so
  _lib.py          ctypes binding of the C ABI (fails loudly when the library is missing)
  ops.py           torch-tensor wrappers of the C ABI entry points
  engine.py        batched BO-episode engine (B episodes in lock-step on one GPU)
  policies.py      NeuralAF drop-in (metabo/policies/policies.py API)
  environment.py   MetaBO gym-style env drop-in + vectorised env
  ppo/             BatchRecorder / PPO drop-ins (metabo/ppo API), torch.distributed sharding

There is no CPU fallback: every numeric entry point requires the CUDA library and a CUDA device.
"""
__version__ = "0.1.0"
