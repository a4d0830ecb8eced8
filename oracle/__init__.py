"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU fp32 restatements of the reference algorithms on the hot path, each function citing the reference file:line it
follows.  Pinned against the unmodified reference by tests/golden/make_golden.py (AR + NAR; the Vocos vocoder is
"parity unpinned", see vocos_oracle.py) and tests/golden/make_bpe_golden.py (minbpe tokenisers, bpe_oracle.py) and
tests/golden/make_trim_golden.py (silence trim, trim_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product package mars5_tts_b200 never does.
"""
