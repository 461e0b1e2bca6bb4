"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's algorithm for the hot path (flow decoder
step, relative-position attention encoder, monotonic alignment search).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- and there only as the checker or the
timed CPU baseline, never as the thing shipped.  ``glow_tts_b200`` never
imports this package (tests/test_no_oracle_in_product.py enforces it).
"""
