"""CPU oracle for the contrastors hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``contrastors_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / CPU baseline, never as the product path.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the *unmodified* reference
(``/root/reference/src/contrastors/{loss,distributed}.py`` and the reference's own
pure-torch encoder ``models/huggingface/modeling_hf_nomic_bert.py``) in the build
container, runs it on seeded inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors and
against the one known-answer test the reference ships (``tests/test_loss.py:5-17``,
loss = 1.0940139293670654 with an identity scale).
"""
