"""the pasted solver variants of fd_discrete_barrier_option.py:425-2421 (Rannacher start-up,
discounted rebate, double-out, Richardson over N and N/2 ...) have no runnable reference.

``oracle/probe_fragments.py`` cuts every solver fragment out by its line range, compiles it with the imports the file
lacks and grafts it onto the two classes it could belong to; ``tests/golden/fragments.json`` records what happened.
None executes without a repair (undefined helpers, a misspelt attribute, indentation errors inside the fragment), so
the engine options that follow them (``barrier_options.DiscreteBarrierCrankNicolsonLogEx``) stay held to the extended
oracle -- "parity unpinned", by demonstration rather than by omission."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _records():
    with open(os.path.join(HERE, "golden", "fragments.json")) as fh:
        return json.load(fh)["records"]


def test_every_listed_fragment_was_probed():
    recs = _records()
    starts = sorted(r["lines"][0] for r in recs)
    assert starts == [428, 912, 1009, 1133, 1261, 1460, 1841, 2238]


def test_no_fragment_runs_unrepaired():
    for r in _records():
        if not r["compiles"]:
            assert "IndentationError" in r["error"] or "SyntaxError" in r["error"], r
            continue
        assert r["hosts"], r
        for host, h in r["hosts"].items():
            assert h["ran"] is False, (r["fragment"], host)
            assert h["error"].split(":")[0] in ("AttributeError", "TypeError", "NameError"), (r["fragment"], host, h["error"])


def test_probe_reproduces_here_when_the_reference_is_present():
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("/root/reference is not present (GPU box)")
    from oracle import probe_fragments
    now = {r["fragment"]: r for r in probe_fragments.probe()}
    for r in _records():
        n = now[r["fragment"]]
        assert n["compiles"] == r["compiles"]
        for host, h in r["hosts"].items():
            assert n["hosts"][host]["ran"] == h["ran"] and n["hosts"][host]["error"] == h["error"]
