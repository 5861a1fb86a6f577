"""Native Wigner-3j tables: the drop-in for pre_calc_Wigner (core_deism.py:1842-1913)."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib


def wigner_tables(N: int, V: int):
    """{'W_1_all': complex64[N+1,V+1,N+V+1], 'W_2_all': complex64[N+1,V+1,N+V+1,2N+1,2V+1]}
    with W_2[n,v,l,m,u] = (n v l; -m u m-u), negative m/u wrapped like NumPy indices."""
    W1 = np.zeros((N + 1, V + 1, N + V + 1), dtype=np.complex64)
    W2 = np.zeros((N + 1, V + 1, N + V + 1, 2 * N + 1, 2 * V + 1), dtype=np.complex64)
    check(lib().deism_wigner_tables(int(N), int(V), W1.ctypes.data, W2.ctypes.data),
          "deism_wigner_tables")
    return {"W_1_all": W1, "W_2_all": W2}


def pre_calc_Wigner(params, timeit=True):
    """Same contract as the reference: fills params['Wigner'] (and 'updated_where')."""
    params["Wigner"] = wigner_tables(int(params["sourceOrder"]), int(params["receiverOrder"]))
    if params.get("track_updated_where"):
        params["updated_where"]["Wigner"] = ["pre_calc_Wigner"]
    return params
