"""Native Wigner-3j tables: the drop-in for pre_calc_Wigner (core_deism.py:1842-1913)."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib


_tables = {}


def wigner_tables(N: int, V: int):
    """{'W_1_all': complex64[N+1,V+1,N+V+1], 'W_2_all': complex64[N+1,V+1,N+V+1,2N+1,2V+1]}
    with W_2[n,v,l,m,u] = (n v l; -m u m-u), negative m/u wrapped like NumPy indices.

    The tables depend on the two orders only: they are generated once per process and handed out as
    read-only arrays (the same objects, so the ORG coupling plan built from them is found again too)."""
    key = (int(N), int(V))
    hit = _tables.get(key)
    if hit is None:
        W1 = np.zeros((N + 1, V + 1, N + V + 1), dtype=np.complex64)
        W2 = np.zeros((N + 1, V + 1, N + V + 1, 2 * N + 1, 2 * V + 1), dtype=np.complex64)
        check(lib().deism_wigner_tables(int(N), int(V), W1.ctypes.data, W2.ctypes.data),
              "deism_wigner_tables")
        W1.flags.writeable = False
        W2.flags.writeable = False
        if len(_tables) >= 16:
            _tables.clear()
        hit = _tables[key] = (W1, W2)
    return {"W_1_all": hit[0], "W_2_all": hit[1]}


def pre_calc_Wigner(params, timeit=True):
    """Same contract as the reference: fills params['Wigner'] (and 'updated_where')."""
    params["Wigner"] = wigner_tables(int(params["sourceOrder"]), int(params["receiverOrder"]))
    if params.get("track_updated_where"):
        params["updated_where"]["Wigner"] = ["pre_calc_Wigner"]
    return params
