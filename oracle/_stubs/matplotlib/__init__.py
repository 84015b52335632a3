"""Import stub so the unmodified reference (which imports matplotlib.pyplot at module top,
FoKLRoutines.py:18, getKernels.py:4) can be imported in a container without matplotlib.
Test infrastructure only."""
