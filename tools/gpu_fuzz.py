#!/usr/bin/env python
"""Extended randomised GPU parity run (the generator and the bars of tests/test_zzz_gpu_fuzz.py over many more seeds):
`python tools/gpu_fuzz.py [n_box] [n_general] [first_seed]` -> one JSON line with the counts and the failing cases, if any."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as ge
    ge.build()
    import fuzz_cases
    from test_zzz_gpu_fuzz import _run

    n_box = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    n_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    s0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    t0 = time.time()
    done = pairs = skipped = 0
    bad = []
    jobs = [("box3d", lambda s: fuzz_cases.box_case(s, False), n_box), ("box2d", lambda s: fuzz_cases.box_case(s, True), n_box)]
    jobs += [(k, (lambda kk: (lambda s: fuzz_cases.general_case(s, kk)))(k), n_gen) for k in ("skew2d", "skew3d", "prism2d", "mixed3d")]
    for name, make, n in jobs:
        for seed in range(s0, s0 + n):
            try:
                p = _run(make(seed))
                pairs += p
                skipped += p == 0
            except AssertionError as ex:
                bad.append({"kind": name, "seed": seed, "error": str(ex)[:300]})
            except Exception as ex:      # anything else (library error) is a failure too
                bad.append({"kind": name, "seed": seed, "error": repr(ex)[:300]})
            done += 1
    print(json.dumps({"cases": done, "pairs_compared": int(pairs), "cases_without_pairs_or_non_finite_reference": int(skipped),
                      "failures": len(bad), "failing": bad[:20], "first_seed": s0, "seconds": round(time.time() - t0, 1)}))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
