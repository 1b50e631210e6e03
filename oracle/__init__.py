"""TEST INFRASTRUCTURE ONLY.  CPU oracles for the ANNP hot path.

Nothing in the product package (`meng_zhang_b200/`) may import from here; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs do.
"""
