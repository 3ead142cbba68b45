"""Run the UNMODIFIED reference on the 24 seeded random configurations of oracle/random_configs.py (build container
only) and store its results in tests/golden/random_configs.npz: final factors, losses and relative errors of
`cmf_aoadmm(..., n_iter_max=10, return_errors=True)` per seed.

    python oracle/gen_golden_random.py

This closes the parity chain of the randomised differential test: CUDA path == oracle (GPU test) and oracle ==
reference (tests/test_oracle.py::test_oracle_matches_reference_on_random_configs, CPU) on the SAME 24 problems.
Test infrastructure, like gen_golden.py (same stand-ins for the absent `tensorly` / `condat_tv` packages)."""
import os
import sys

sys.dont_write_bytecode = True
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache_golden")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_tl_standin"))
sys.path.insert(0, os.path.join(HERE, "_condat_standin"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import matcouply.decomposition as D  # noqa: E402  (the reference)
from matcouply import penalties as P  # noqa: E402

from oracle import aoadmm_oracle as O  # noqa: E402
from oracle.random_configs import N_RANDOM_CONFIGS, random_config, reference_safe  # noqa: E402

if __name__ == "__main__":
    out = {}
    for seed in range(N_RANDOM_CONFIGS):
        X, R, kw = random_config(seed)
        kw = reference_safe(kw)
        call = dict(kw)
        call["regs"] = O.regs_from_spec(call.pop("regs_spec"), P)
        cmf, diag = D.cmf_aoadmm([x.copy() for x in X], R, return_errors=True, **call)
        _, (A, Bs, C) = cmf
        out[f"s{seed}_A"], out[f"s{seed}_B"], out[f"s{seed}_C"] = A, np.concatenate(Bs, 0), C
        out[f"s{seed}_loss"] = np.asarray(diag.regularized_loss, dtype=np.float64)
        out[f"s{seed}_rec"] = np.asarray(diag.rec_errors, dtype=np.float64)
        print(seed, [[p[0] for p in m] for m in kw["regs_spec"]], "loss", float(diag.regularized_loss[-1]))
    path = os.path.join(ROOT, "tests", "golden", "random_configs.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")
