"""CPU oracle for the uncertainty-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the reference's algorithm for the one path this
repository accelerates (conditional reverse-diffusion sampling + MPV / gx reduction).
It exists so the CUDA path can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it -- never the product package, which has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * nsdiff_oracle, sigma_oracle, mpv_oracle, windows_oracle: PINNED against outputs of
    the reference itself, run in the build container through the arithmetic-free import
    stubs in ``oracle/_stubs`` (``oracle/make_golden.py`` is the committed generator,
    fixtures live in ``tests/golden/``).
  * tmdm_oracle sampler: PINNED the same way (reference ``tmdm_model.py`` /
    ``tmdm_diffusion_utils.py`` import without third-party code).
  * fx_oracle (ns-Transformer condition encoder): "parity unpinned" -- its blocks live in
    the un-vendored dependency torch-timeseries==0.1.10 and no shipped checkpoint holds
    weights for it.
"""
