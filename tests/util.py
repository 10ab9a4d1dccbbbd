import numpy as np


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    m = np.abs(b) > 0
    if not m.any():
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def tau_relerr(tau, tau_ref, last):
    """Compare optical-depth columns only down to `last` (the reference leaves the rest at 0)."""
    worst = 0.0
    for w in range(tau_ref.shape[0]):
        L = int(last[w])
        worst = max(worst, relerr(tau[w, :L + 1], tau_ref[w, :L + 1]))
    return worst


def apply_setters(obj, setters):
    if "radius" in setters:
        obj.set_radius(setters["radius"])
    if "cloudtop" in setters:
        obj.set_cloudtop(setters["cloudtop"])
    if "scattering" in setters:
        obj.set_scattering(1, setters["scattering"])
