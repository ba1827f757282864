"""Extracts the reference's own golden vectors of the two-segment data source into a fixture that travels to the GPU box
(/root/reference does not exist there).  Source: /root/reference/cc/env/test_make_env.py:146-335 -- DYNAMIC_INPUTS (40 actions)
and `expected_positions` of test_dynamic_environments for two_segments_v1/v2/v3 (40 float32 observations each; the test
feeds DYNAMIC_INPUTS[i] to env.step and compares observation i, so observation 0 is the reset observation and action 0 is
ignored by dm_env's reset-on-first-step).  Data only -- no reference code is copied.

    python tests/golden/make_two_segments_golden.py   ->  tests/golden/two_segments_golden.json
"""
import json
import os
import re

SRC = "/root/reference/cc/env/test_make_env.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def floats(block):
    return [float(x) for x in block.replace("\n", " ").split(",") if x.strip()]


def main():
    src = open(SRC).read()
    inputs = floats(re.search(r"DYNAMIC_INPUTS = \[(.*?)\]", src, re.S).group(1))
    tail = src[src.index("expected_positions"):]
    out = {"source": "cc/env/test_make_env.py:146-335", "dynamic_inputs": inputs, "expected_positions": {}}
    for env in ("two_segments_v1", "two_segments_v2", "two_segments_v3"):
        out["expected_positions"][env] = floats(re.search(r'"%s",\s*\[(.*?)\]' % env, tail, re.S).group(1))
        assert len(out["expected_positions"][env]) == 40
    assert len(inputs) == 40
    with open(os.path.join(HERE, "two_segments_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
