"""Version of the reference interface this package mirrors (gpuSolve/_version.py: 1.3.2)."""
__version__ = ['1', '3', '2']


def version():
    return '.'.join(__version__)
