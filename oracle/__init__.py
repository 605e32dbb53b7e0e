"""CPU oracle: float64 NumPy restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline.  The product path (``5g_based_..._b200``) never imports it and fails
loudly when the CUDA library is missing.

PARITY UNPINNED.  The reference (xds0112/5G_based_System_level_Integrated_
Sensing_and_Communication_Simulator, commit f16d1fb) is MATLAB + closed
MathWorks toolboxes, ships no tests / golden vectors, and neither MATLAB nor
Octave exists in the build image.  Every function below restates the cited
reference lines (file:line, relative to the reference root); where the
arithmetic lives in a closed toolbox (``phased.CFARDetector2D``, ``kaiser``,
``findpeaks``, ``nrOFDMDemodulate``, ``nrPUSCHCodebook``, ``nrCDLChannel`` ...)
the function restates the *published* behaviour (MathWorks docs / 3GPP TS
38.211 / 38.214 / TR 38.901) and is tagged ``PARITY-UNPINNED`` in its docstring.
The oracle itself is pinned by analytic known-answer tests (tests/test_oracle_*.py)
and by independent SciPy implementations where one exists.
"""
