"""Oracle for ``Agent_State.update_global_goal`` (nav/agent/agent_state.py:376-416) - SURVEY.md section 8(f), N1b.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the geodesic distance: the reference computes it with scikit-fmm
(``skfmm.distance``, pinned at 2019.1.30 by peanut.Dockerfile:8), which is not installable here; ``oracle/fmm.c`` restates
its published second-order fast-marching algorithm (see that file's header) and this module restates the method around it
line by line.  scikit-image is absent too: ``skimage.morphology.binary_dilation(image, selem)`` is
``scipy.ndimage.binary_dilation(image, structure=selem)`` and ``skimage.morphology.disk(r)`` is the published
``x**2 + y**2 <= r**2`` mask (the same two wrappers oracle/goal_map.py uses).
"""
import ctypes
import os

import numpy as np
import numpy.ma as ma
from scipy import ndimage

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libpn_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _LIB = ctypes.CDLL(path)
        _LIB.pn_oracle_fmm_distance.restype = ctypes.c_int
        _LIB.pn_oracle_fmm_distance.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _LIB


def disk(radius):
    """skimage.morphology.disk"""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=np.uint8)


def skfmm_distance(phi_ma):
    """skfmm.distance(phi, dx=1) for a 2-D masked array, as pfmm.py wraps the C++ marcher: masked / unreached cells come
    back masked (their data set to 0)."""
    phi = np.ascontiguousarray(ma.getdata(phi_ma), dtype=np.float64)
    mask = np.ascontiguousarray(ma.getmaskarray(phi_ma), dtype=np.uint8)
    out = np.empty_like(phi)
    err = _lib().pn_oracle_fmm_distance(phi.ctypes.data, mask.ctypes.data, phi.shape[0], phi.shape[1], out.ctypes.data)
    if err:
        raise RuntimeError("Negative discriminant in distance marcher quadratic.")
    big = out == np.finfo(float).max
    if big.any():
        out[big] = 0
        return ma.MaskedArray(out, big)
    return out


def traversible(full_map0, selem, collision_map, visited_vis):
    """agent_state.py:382-386."""
    trav = ndimage.binary_dilation(np.rint(full_map0), structure=selem) != True  # noqa: E712
    trav[collision_map == 1] = 0
    trav[visited_vis == 1] = 1
    return trav


def distance_field(trav, agent_r, agent_c):
    """agent_state.py:388-393 -> dd [full_w, full_h] float64, np.inf where not traversible / unreachable."""
    traversible_ma = ma.masked_values(trav * 1, 0)
    traversible_ma[agent_r, agent_c] = 0
    dd = skfmm_distance(traversible_ma)
    dd = ma.filled(dd, np.max(dd) + 1)
    dd[np.where(dd == np.max(dd))] = np.inf
    return dd


def update_global_goal(full_map0, collision_map, visited_vis, lmb, loc_r, loc_c, target_pred, col_rad=4,
                       dist_weight_temperature=500.0, map_resolution=5, prev_dd_wt=None, global_goals=None,
                       last_global_goal=None):
    """The whole method on plain arrays.  Returns dict(dd, dd_wt, value, global_goals, last_global_goal)."""
    full_w, full_h = full_map0.shape
    trav = traversible(full_map0, disk(col_rad), collision_map, visited_vis)
    ar = int(np.clip(loc_r + lmb[0], a_min=0, a_max=full_w - 1))
    ac = int(np.clip(loc_c + lmb[2], a_min=0, a_max=full_h - 1))
    dd = distance_field(trav, ar, ac)
    temperature = dist_weight_temperature / map_resolution
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        dd_wt = np.exp(-dd / temperature)[lmb[0]:lmb[1], lmb[2]:lmb[3]]
    if np.sum(dd_wt) < 10 and prev_dd_wt is not None:  # stuck inside obstacle, use last dd_wt
        dd_wt = prev_dd_wt
    if dist_weight_temperature == -1:
        value = target_pred
    elif dist_weight_temperature == 0:
        dd = dd.copy()
        dd[np.where(dd < 60)] = np.inf
        value = np.exp(-dd / 100.)[lmb[0]:lmb[1], lmb[2]:lmb[3]]
    else:
        value = target_pred * dd_wt
    new_global_goal = [np.unravel_index(value.argmax(), value.shape)]
    if new_global_goal != last_global_goal:  # avoid repeating the last goal
        last_global_goal = global_goals
        global_goals = new_global_goal
    return dict(dd=dd, dd_wt=dd_wt, value=value, global_goals=global_goals, last_global_goal=last_global_goal)
