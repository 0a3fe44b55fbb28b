"""Writes tests/golden/oracle_trees.json from the CPU oracle.

The reference is Rust and cannot be built or imported here (no cargo/rustc), and its own tests hold
no search-tree fixtures, so these digests are outputs of the oracle itself, taken after it agreed with
the independent Python restatement and the reference's known-answer tests.  Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle  # noqa: E402
from test_oracle_tree import run_golden_case  # noqa: E402

CASES = [
    dict(n=19, b=8, steps=200, seed=0, first_root=0),
    dict(n=19, b=4, steps=800, seed=1, first_root=4096),
    dict(n=8, b=16, steps=100, seed=2, first_root=0),
    dict(n=33, b=4, steps=100, seed=3, first_root=0),
    dict(n=64, b=2, steps=60, seed=4, first_root=7),
]

if __name__ == "__main__":
    oracle.build()
    out = dict(note="oracle outputs (lambda_1 by multisection); see make_golden.py",
               cases=[dict(config=c, expect=run_golden_case(oracle, c)) for c in CASES])
    with open(os.path.join(HERE, "oracle_trees.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(CASES), "cases")
