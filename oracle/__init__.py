"""CPU oracle for the YourMT3 inference hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package ``yourmt3_b200``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the reported CPU baseline -- never as the thing shipped.

Pinning status: the mounted reference
(/root/reference) contains only README.md + LICENSE, so there are no reference
tests or golden vectors to pin against.  The arithmetic of the path lives in
third-party dependencies that ARE installed in this image and on the GPU box:
torchaudio 2.11.0+cu128, torch 2.11.0+cu128, transformers 5.5.0.  Each oracle
function is pinned against those libraries' own CPU outputs (live in tests, and
through the committed fixtures under tests/golden/ made by
tools/make_golden.py).  Architecture wiring that exists only upstream
(Perceiver-TF, MoE, multi-channel decoder) is restated from the YourMT3+ paper
and memory of mimbres/YourMT3 and is marked "parity unpinned".
"""
