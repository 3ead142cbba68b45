"""Run oracle/host_cases.py against the UNMODIFIED reference (build container only) and write the outcomes — return
value summaries and "<ExceptionType>: <message>" strings — to tests/golden/host_contract.json, and the signatures of the public callables to
tests/golden/host_signatures.json.

    python oracle/gen_golden_host.py

Test infrastructure, like gen_golden.py: the reference is imported as-is from /root/reference/src against the NumPy
stand-in for `tensorly` (oracle/_tl_standin)."""
import json
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

from matcouply import coupled_matrices, data, decomposition, penalties, random  # noqa: E402  (the reference)

from oracle.host_cases import host_cases, public_signatures  # noqa: E402

if __name__ == "__main__":
    out = host_cases(coupled_matrices, random, decomposition, penalties)
    path = os.path.join(ROOT, "tests", "golden", "host_contract.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    n_err = sum(isinstance(v, str) and ": " in v and v.split(":")[0].endswith(("Error", "Exception")) for v in out.values())
    print(f"{len(out)} cases ({n_err} raising) -> {path}")
    sigs = public_signatures(dict(coupled_matrices=coupled_matrices, data=data, decomposition=decomposition,
                                  penalties=penalties, random=random))
    path = os.path.join(ROOT, "tests", "golden", "host_signatures.json")
    with open(path, "w") as f:
        json.dump(sigs, f, indent=1, sort_keys=True)
    print(f"{len(sigs)} signatures -> {path}")
