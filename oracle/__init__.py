"""CPU oracle for the spectral hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker or as the timed CPU baseline -- never as the thing
shipped.  The product package (``speech_enhancement_pytorch_b200``) never imports this module
and raises when its CUDA library is missing.

Two restatements live here:

* ``spectral_oracle``  -- torch-CPU fp32 restatement of the reference's call sites.  The
  reference's arithmetic for this path lives in a third-party dependency (PyTorch:
  ``torch.stft`` / ``torch.istft`` / ``F.conv1d`` / ``F.conv_transpose1d``; the reference pins
  torch 1.7.1+cu110, this image has 2.11.0+cu128), so the restatement calls the same
  library entry points with the reference's arguments (reference file:line cited per function).
* ``spectral_np64``    -- numpy float64 restatement from first principles (explicit reflect
  padding, framing, DFT, overlap-add, adjoints), used for error budgeting and to make sure the
  fp32 oracle is not hiding a library quirk.

Parity pin: ``tests/golden/*.npz`` were produced by importing the REAL reference from
``/root/reference`` in the build container (``tests/golden/make_golden.py``); the oracle is
checked against every one of them in ``tests/test_oracle_golden.py``.  The MR-STFT loss does not
exist in the reference; its definition is the pin, and its
golden vectors are produced with the reference's own ``stft_custom`` underneath:
"parity unpinned by the reference's tests" for that one function.
"""
