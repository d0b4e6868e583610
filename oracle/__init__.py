"""CPU oracle for the mdctGAN hot path  --  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy / torch-CPU) of the
algorithms the reference (neoncloud/mdctGAN @ 0e2063cf) runs on the
MDCT -> generator -> IMDCT path.  It exists to *check* the CUDA product in
``mdctgan_b200/``; it is never the thing that is shipped or measured.

Who may import this package:
  * ``tests/``                       (parity checker)
  * ``__graft_entry__.smoke()``      (one-shot checker on cuda:0)
  * ``bench.py``                     (``cpu_baseline`` leg and ``--impl reference`` only)

``mdctgan_b200`` never imports ``oracle`` (tests/test_boundary.py greps for it)
and has no CPU fallback: it raises when ``libmdctgan_b200.so`` is missing.

Pinning: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the *reference itself*, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``
(committed) -> ``tests/golden/*.npz`` (committed).  ``tests/test_oracle_golden.py``
re-checks every oracle function against those fixtures on CPU.
The one exception is ``oracle/bottlestack_ref.py`` (third-party
``bottleneck_transformer_pytorch==0.1.4``, absent from /root/reference and from
this image): parity unpinned -- restated from the package's published semantics.
"""
