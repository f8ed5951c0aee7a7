"""B200-native AdaBins head + loss + external-info path (drop-in for the reference's Python surface).

Importing the package is cheap and works on a CPU-only box; every *operator* loads the sm_100a shared
library on first use and raises if it (or a GPU) is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
