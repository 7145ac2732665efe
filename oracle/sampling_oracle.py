"""ORACLE (test infrastructure): numpy restatement of the counter-based Gumbel-max sampler of
metabo_b200/csrc/reduce.cu (the CUDA replacement of Categorical(logits).sample(),
metabo/policies/policies.py:144-146).  The reference draws from torch's process-global CPU
generator, which cannot be reproduced on a GPU; parity for the stochastic
policy is therefore (a) this bit-level restatement of the noise and (b) a chi-square test of the
empirical action distribution against softmax(logits) (tests/test_reduce_gpu.py)."""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def gumbel_noise(seed, b, step, M):
    """fp32 Gumbel noise for episode b (scalar) at `step`, candidates 0..M-1."""
    with np.errstate(over="ignore"):
        key = mix64(mix64(np.uint64(seed) + GOLD * np.uint64(b + 1)) ^ (np.uint64(step) + GOLD))
        r = mix64(key + GOLD * (np.arange(M, dtype=np.uint64) + np.uint64(1)))
    u = ((r >> np.uint64(41)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)   # 23 bits: exact, never 1
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def sample(logits, seed, step):
    """actions [B] for logits [B, M] (first-max tie rule)."""
    logits = np.asarray(logits, dtype=np.float32)
    out = np.zeros(logits.shape[0], dtype=np.int64)
    for b in range(logits.shape[0]):
        out[b] = int(np.argmax(logits[b] + gumbel_noise(seed, b, step, logits.shape[1])))
    return out
