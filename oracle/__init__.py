"""CPU oracle for the RAD-MMM flow-decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``rad-mmm_b200/``)
imports this; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline, never as the thing shipped.

The oracle is a functional restatement (plain torch-CPU / numpy, one function
per op, state passed as a flat ``state_dict`` with the reference's key names)
of the algorithm in NVIDIA/RAD-MMM's ``decoders.py``, ``models/radmmm.py``,
``common.py``, ``partialconv1d.py``, ``splines.py``, ``maskedbatchnorm1d.py``,
``loss.py`` and ``audio_processing.py``.  Each function cites the reference
file:line it follows.

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference itself, imported unmodified from /root/reference in the build
container by ``tests/golden/make_golden.py``; the resulting vectors live in
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks the oracle
against them.  The one part that stays "parity unpinned" is the librosa-0.8.0
Slaney mel filterbank (librosa is not vendored in the reference and not
installed here): ``oracle/frontend.py`` restates its published algorithm and
is cross-checked against torchaudio's Slaney filterbank instead.
"""
