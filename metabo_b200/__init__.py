"""metabo_b200 -- B200-native (sm_100a) hot path of MetaBO behind MetaBO's own Python API.

Layout
  csrc/                      hand-written CUDA kernels + the C ABI (include/metabo_b200.h) -> libmetabo_b200.so
                             gp.cu (GP fit / posterior), neural_af_tc.cu (tcgen05 NeuralAF), mlp_simt.cu, reduce.cu,
                             ppo_update.cu / ppo_update_simt.cu (PPO optimiser step), comm.cu (peer-memory gradient exchange)
  _lib.py                    ctypes binding of the C ABI (fails loudly when the library is missing)
  ops.py                     torch-tensor wrappers of the C ABI entry points
  engine.py                  AFEngine: one AF-optimisation round for B episodes in lock-step on one GPU
  environment/vec_env.py     VecMetaBO: B device-resident episodes; environment/metabo_gym.py: MetaBO gym-style drop-in
  environment/functions.py   objective families evaluated on the device (objectives.py)
  policies/policies.py       NeuralAF, EI, UCB, PI drop-ins; policies/taf.py: TAF on the batched GP kernel
  ppo/                       BatchRecorder / PPO drop-ins (metabo/ppo API), data-parallel update over torch.distributed ranks
  eval/evaluate.py           eval_experiment drop-in
  gym_compat.py, sobol.py    the slices of gym / sobol_seq the reference uses

There is no CPU fallback: every numeric entry point requires the CUDA library and a CUDA device.
"""
__version__ = "0.2.0"
