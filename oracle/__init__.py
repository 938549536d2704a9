"""CPU oracle (torch fp32 restatement of the reference's algorithm) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from
llamarec_b200/.  Parity is pinned against the reference executed by oracle/make_golden.py
(fixtures in tests/golden/)."""
