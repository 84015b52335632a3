"""FoKL drop-in package backed by hand-written sm_100a CUDA kernels (libfokl_b200.so).

Importable as `FoKL` so that pickled models keep the class identity `FoKL.FoKLRoutines.FoKL`
(reference: src/FoKL/FoKLRoutines.py:1840-1842)."""
__all__ = ['FoKLRoutines', 'getKernels']
