"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (AI4XDE -> DeepXDE PyTorch backend):
``ddelite`` re-states the DeepXDE semantics listed in SURVEY.md Appendix B using
plain ``torch.autograd`` on the CPU, ``cases`` re-states the reference ``pde``
closures of the BASELINE configs, ``sampling`` re-states NumPy's legacy
``RandomState.choice`` and the RAD / RAR_D / R3 selections.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package.  The product
package ``ai4xde_b200`` never does (tests/test_boundary.py checks that).

Parity pin: the third-party module that holds the arithmetic (``deepxde``,
unpinned in /root/reference/setup.py:41-46) is absent from this image, so the
restatement is pinned by (i) running the UNMODIFIED reference case/algorithm
sources from /root/reference on top of ``ddelite`` (tests/golden/make_golden.py,
fixtures committed under tests/golden/), (ii) analytic known-answer residuals of
the reference's own ``sol()`` functions and (iii) ``numpy.random.choice`` itself
for the sampler.  DeepXDE's own outputs are not available: "parity pinned to the
reference's sources + PyTorch autograd, DeepXDE internals restated".
"""
