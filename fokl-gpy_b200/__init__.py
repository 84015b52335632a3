"""fokl-gpy_b200: B200-native implementation of the FoKL-GPy `fit` hot path.

The directory name is not a valid Python identifier; put it on sys.path and `import FoKL`
(the drop-in package keeps the reference's module names so pickled models stay loadable):

    sys.path.insert(0, '<repo>/fokl-gpy_b200'); from FoKL import FoKLRoutines
"""
