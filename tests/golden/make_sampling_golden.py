"""Generates tests/golden/sampling_cases.json by running the REFERENCE's python/sampling.py (imported from /root/reference): for
seeded synthetic logits, numpy.random.seed(s) then sample_logits(...) gives the token; the same seed's first random_sample() is the
`u` our oracle (oracle/sampling_oracle.py) and the GPU kernel take explicitly."""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("ref_sampling", "/root/reference/python/sampling.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def make_logits(seed, n_vocab, scale):
    return (np.random.RandomState(seed).standard_normal(n_vocab) * scale).astype(np.float32)


cases = []
configs = [(1.0, 0.8), (0.8, 0.5), (1.0, 1.0), (1.5, 0.9), (0.0, 0.8), (0.7, 0.0), (2.0, 0.3), (1.0, 0.05)]
for n_vocab, scale in ((256, 2.0), (50277, 3.0), (65536, 4.0), (1000, 8.0)):
    for ci, (temperature, top_p) in enumerate(configs):
        for k in range(6):
            lseed = 1000 * ci + k + n_vocab
            rseed = 77 * ci + k
            bias = {3: 2.5, 17: -4.0, n_vocab - 1: 1.0} if k == 5 else None
            logits = make_logits(lseed, n_vocab, scale)
            np.random.seed(rseed)
            token = int(ref.sample_logits(logits.copy(), temperature, top_p, bias))
            u = float(np.random.RandomState(rseed).random_sample())
            cases.append({"n_vocab": n_vocab, "scale": scale, "logits_seed": lseed, "temperature": temperature, "top_p": top_p,
                          "bias": bias and {str(i): v for i, v in bias.items()}, "rng_seed": rseed, "u": u, "token": token})
json.dump(cases, open(os.path.join(ROOT, "tests", "golden", "sampling_cases.json"), "w"), indent=0)
print(len(cases), "cases")
