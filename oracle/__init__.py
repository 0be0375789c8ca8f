"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithm for the DREAM belief-map hot path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this package, and only as the checker.  The product (`dream_b200/`) never imports it and has no CPU
fallback.

Pinning: the restatement is checked (tests/test_oracle_golden.py) against golden vectors produced by
the *reference's own code* imported from /root/reference in the build container by
`oracle/make_golden.py` (committed under tests/golden/), and against the reference's own known-answer
test for peak extraction (test/test_image_proc.py:94-120).  The reference has no test, fixture or
golden tensor for the network forward/backward itself, so for that part the
pin is "outputs of the reference run here".
"""
