"""alayalite_b200: AlayaLite's batched k-NN hot path on B200, behind the reference's Python API.

    from alayalite_b200 import Client, Index, calc_gt, calc_recall

The package needs its two native pieces, built in-tree by `python -m alayalite_b200.build`:
libalaya_b200.so (sm_100a kernels + C ABI) and the _alayalitepy pybind module. There is no CPU
fallback: the first use of the API fails loudly if they are missing.
"""

import importlib as _importlib

__all__ = ["Client", "Collection", "Index", "IndexParams", "load_fvecs", "load_ivecs", "calc_recall", "calc_gt"]
__version__ = "0.1.0"

_WHERE = {"Client": ".client", "Collection": ".collection", "Index": ".index", "IndexParams": ".schema", "calc_gt": ".utils", "calc_recall": ".utils",
          "load_fvecs": ".utils", "load_ivecs": ".utils"}


def __getattr__(name):  # PEP 562: keeps `python -m alayalite_b200.build` importable before the build
    if name in _WHERE:
        try:
            return getattr(_importlib.import_module(_WHERE[name], __name__), name)
        except (ImportError, OSError) as e:
            raise ImportError(f"alayalite_b200 native libraries are missing or broken ({e}); run "
                              "`python -m alayalite_b200.build` (needs nvcc). There is no CPU fallback.") from e
    raise AttributeError(name)
