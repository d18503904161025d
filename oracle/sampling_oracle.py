"""TEST INFRASTRUCTURE (only tests/ may import this). CPU restatement of the reference's token sampling
(python/sampling.py:10-52) with the one random draw made explicit: `u` replaces the uniform number numpy.random.choice takes
from the global RandomState (legacy RandomState.choice with p: cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, 'right')).
Pinned by tests/golden/sampling_cases.json, produced by running the reference module itself (tests/golden/make_sampling_golden.py)."""
import numpy as np


def softmax(x: np.ndarray) -> np.ndarray:                      # sampling.py:5-8 (float32 throughout for float32 logits)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def sample_probs_u(probs: np.ndarray, temperature: float, top_p: float, u: float, logit_bias=None) -> int:
    if not (0.0 <= temperature):                               # :20-23
        raise ValueError("temperature")
    if not (0.0 <= top_p <= 1.0):
        raise ValueError("top_p")
    probs = probs.copy()
    if top_p == 0.0:                                           # :25-26
        top_p = 1.0
    if logit_bias:                                             # :28-37
        logits = np.log(probs)
        ids, values = zip(*logit_bias.items())
        logits[list(ids)] += values
        logits -= logits.max(axis=-1, keepdims=True)
        probs = np.exp(logits) / np.sum(np.exp(logits))
    if temperature == 0.0:                                     # :39-40
        return int(np.argmax(probs))
    if top_p < 1.0:                                            # :42-45
        sorted_probs = np.sort(probs)[::-1]
        cumulative = np.cumsum(sorted_probs)
        cutoff = float(sorted_probs[np.argmax(cumulative > top_p)])
        probs[probs < cutoff] = 0
    if temperature != 1.0:                                     # :47-48
        probs = np.power(probs, 1.0 / temperature)
    probs = probs / np.sum(probs)                              # :50
    cdf = probs.astype(np.float64).cumsum()                    # :52 numpy.random.choice(a, p=probs)
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def sample_logits_u(logits: np.ndarray, temperature: float, top_p: float, u: float, logit_bias=None) -> int:   # :10-16
    return sample_probs_u(softmax(np.asarray(logits, dtype=np.float32).copy()), temperature, top_p, u, logit_bias)


def boundary_distance(logits: np.ndarray, temperature: float, top_p: float, u: float, logit_bias=None) -> float:
    """How far u is from the nearest edge of the chosen token's CDF interval (tests skip knife-edge draws)."""
    probs = softmax(np.asarray(logits, dtype=np.float32).copy())
    if temperature == 0.0:
        s = np.sort(probs)
        return float(s[-1] - s[-2])
    p = probs.copy()
    if logit_bias:
        lg = np.log(p); ids, values = zip(*logit_bias.items()); lg[list(ids)] += values; lg -= lg.max(); p = np.exp(lg) / np.sum(np.exp(lg))
    tp = 1.0 if top_p == 0.0 else top_p
    if tp < 1.0:
        sp = np.sort(p)[::-1]; cutoff = float(sp[np.argmax(np.cumsum(sp) > tp)]); p[p < cutoff] = 0
    if temperature != 1.0:
        p = np.power(p, 1.0 / temperature)
    p = p / p.sum()
    cdf = p.astype(np.float64).cumsum(); cdf /= cdf[-1]
    return float(np.min(np.abs(cdf - u)))
