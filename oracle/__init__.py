"""TEST INFRASTRUCTURE — CPU oracle of the Uni-Adapter hot path (not shipped, never on the product path).

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and __graft_entry__.smoke() may import this package.
``oracle.tokenizer`` wraps the plain-C restatement (tokenizer_oracle.c); ``oracle.adapters`` is a numpy restatement of
the head, DOTA, MODE-DOTA and fusion arithmetic. Parity pinning: the reference ships no tests or golden vectors
(SURVEY §4), so the oracle is pinned against outputs of the reference's own Python code, generated in the build
container by oracle/make_golden.py and committed under tests/golden/.
"""
